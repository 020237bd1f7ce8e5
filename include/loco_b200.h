/* loco_b200.h -- C ABI of the B200-native LOCO interpolation hot path.
 *
 * The reference (mit-gfx/LOCOInterpolate) has no FFI of its own: its boundary is the C++ class surface
 * of LCSampledFunction / PrecomputedParametricShapeConverter, linked statically.  Each entry point below
 * names the reference interface it stands in for (paths relative to the reference root); INTEGRATION.md
 * shows the few lines a maintainer adds on the reference side to route a batch of queries through it.
 *
 * Conventions (mirroring LCError, LCUtils/LCError.h:7-20): every call returns 0 on success and a
 * non-zero code otherwise; loco_last_error() then gives the description, using the reference's own
 * message texts where the reference has one.  The caller owns every buffer it passes in.  A model handle
 * is bound to ONE GPU and is not re-entrant (one call at a time per handle).  There is no CPU fallback:
 * without a CUDA device every constructor fails with LOCO_ERR_CUDA.
 *
 * Numbering is canonical: leaf index = rank in LCAdaptiveGridCell::getAllLeafCells order
 * (LCAdaptiveGridCell.cpp:247-258), sample index = position in LCSampledFunction::samples_ (= proto id,
 * LCProtoConverter.cpp:147-152), tree cells in pre-order with child1 first (:131-133).
 */
#ifndef LOCO_B200_H
#define LOCO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct loco_model loco_model_t;

enum {
	LOCO_OK = 0,
	LOCO_ERR_INVALID = 1,  /* bad argument / inconsistent graph (message has the detail) */
	LOCO_ERR_IO = 2,       /* "could not open file ..." / "could not parse the file" */
	LOCO_ERR_CUDA = 3,     /* no device, allocation or launch failure */
	LOCO_ERR_RANGE = 4     /* index out of range / buffer too small */
};

/* `device` value that builds the host side of a model only (graph, flattening, support sets, proto I/O) for
 * inspection on machines without a GPU.  Every evaluation call on such a handle fails with LOCO_ERR_CUDA. */
#define LOCO_HOST_ONLY (-1)

/* LCBasisFunction::LCBasisFunctionType (LCAdaptiveSampling/LCBasisFunction.h:19) */
enum { LOCO_LINEAR_BSPLINE = 0, LOCO_CUBIC_BSPLINE = 1 };

/* per-query status byte; the texts are the reference's LCError descriptions */
enum {
	LOCO_ST_OK = 0,
	LOCO_ST_OUTSIDE_CELL = 1,     /* "parameter ouside cell"            LCAdaptiveGridCell.cpp:438-454 */
	LOCO_ST_BAD_ARITY = 2,        /* "invalid number of parameters"     (raised by the C++ facade only) */
	LOCO_ST_NO_COMMON_SAMPLE = 3, /* "could not find closest sample"    LCAdaptiveGridCell.cpp:533-538 */
	LOCO_ST_EMPTY_SUPPORT = 4     /* "cannot create new from empty list of weighted samples" LCFunctionValue.cpp:14-17 */
};

/* Plain-array image of an LCSampledFunction object graph: what a reference-side binding fills from
 * LCSampledFunction::{getRoot,getSamples,getBorderCells,getRanges} (LCSampledFunction.h:11-36).
 * Cells: tree cells in pre-order (child1 = index+1), each with ranges_[2N]; a leaf has split_dir = -1. */
typedef struct loco_graph {
	int32_t n_params;            /* N */
	int32_t value_dim;           /* D: 1 for LCRealFunctionValue, 3V for a V-vertex mesh value */
	int32_t basis_type;          /* LOCO_LINEAR_BSPLINE / LOCO_CUBIC_BSPLINE (basisType_, LCSampledFunction.cpp:15) */
	const double *domain;        /* [2N] ranges_ (min,max per parameter) */

	int64_t n_samples;           /* S */
	const double *sample_center; /* [S*N]  LCSample::center_ (cube coordinates) */
	const double *sample_value;  /* [S*D]  LCSample::shapeInfo_ */
	const int32_t *sample_value_group; /* [S] or NULL: samples sharing one LCFunctionValue object carry the same id */
	const int64_t *bf_offset;    /* [S+1]  CSR: basisFunctions_ of each sample */
	const double *bf_center;     /* [B*N]  LCBasisFunction::center_ */
	const double *bf_support;    /* [B*N]  LCBasisFunction::support_ */
	const double *bf_weight;     /* [B]    LCBasisFunction::weight_ */

	int64_t n_cells;             /* C tree cells */
	const double *cell_ranges;   /* [C*2N] */
	const int32_t *cell_split_dir;   /* [C] splitDirection_, -1 for a leaf */
	const double *cell_split_val;    /* [C] splitVal_ */
	const int32_t *cell_child2;      /* [C] pre-order index of child2_, -1 for a leaf */
	const double *cell_center_value; /* [C*D] or NULL: centerShapeInfo_ of leaves (true f at the cell centre) */
	const int64_t *cell_sample_offset; /* [C+1] CSR: samples_ map of each leaf */
	const int32_t *cell_sample;        /* sample index */
	const double *cell_sample_value;   /* [E*D] or NULL: per-cell value copies (NULL = the sample's own value) */

	int64_t n_border_cells;      /* border cells (cubic only; level -1, all leaves) */
	const double *border_ranges; /* [Bc*2N] */
	const double *border_center_value;   /* [Bc*D] or NULL */
	const int64_t *border_sample_offset; /* [Bc+1] */
	const int32_t *border_sample;
	const double *border_sample_value;   /* or NULL */
} loco_graph_t;

typedef struct loco_info {
	int32_t n_params, value_dim, basis_type, exact_rcp;
	int32_t depth, max_support, max_terms, device;
	int64_t n_samples, n_value_rows, n_leaves, n_cells, n_border_cells, n_basis_functions;
	int64_t n_support_entries, n_terms, device_bytes;
	int64_t n_fold_entries;  /* value rows over all tensor leaves after folding samples that share a value object (0: no tensor leaves) */
	int64_t n_eval_cells;    /* evaluation cells of the sorted scalar path (0: sorted by leaf) */
	int64_t n_eval_terms;    /* term-list entries over all evaluation cells */
	int64_t n_poly_cells;    /* evaluation cells that have a single-polynomial form */
	int64_t n_reused_leaves; /* leaves whose evaluation cells were taken over from the previous state by loco_model_update_from_graph */
} loco_info_t;

/* f : original parameter space [N] -> value [D]   (LCFunction::evalShapeInfo, LCFunction/LCFunction.h:32) */
typedef void (*loco_value_fn)(const double *params, double *out, void *user);

/* ---- construction ------------------------------------------------------------------------------ */

/* PrecomputedParametricShapeConverter::loadFromProto(filename, useAscii, &shape)
 * (LCAdaptiveSampling/LCProtoConverter.h:14, .cpp:26-33 -> convertToCplusplus :356-403 ->
 * addAllNeighboursTopDown + LCSampledFunction::addBorderCells), then flatten + upload to `device`. */
int loco_model_from_proto(const char *path, int ascii, int device, loco_model_t **out);

/* new LCSampledFunction(root, ranges, samples, borderCells) + addBorderCells()
 * (LCAdaptiveSampling/LCSampledFunction.cpp:9-47), from a plain-array graph. */
int loco_model_from_graph(const loco_graph_t *graph, int device, loco_model_t **out);

/* The same handle after the grid changed -- the re-flatten of the refinement loop (LCAdaptiveGrid::adaptiveSample,
 * LCAdaptiveGrid.cpp:481-545: after refineCell only the split cell's doubly-adjacent cells change, getDoubleAdjacentCells :174).
 * `graph` is the NEW state of the same function (plain arrays, as for loco_model_from_graph).  Incremental where it pays:
 * every leaf carries a signature of what its evaluation cells and polynomial cells depend on (box, merged basis functions);
 * leaves whose signature is unchanged take that derived data -- 90 % of the flatten time on the reference's adaptive trees --
 * over from the previous state; the device image is re-uploaded, streams / workspaces / counters of the handle stay.
 * Results are identical to a handle created from `graph` (tests/test_refine.py). */
int loco_model_update_from_graph(loco_model_t *m, const loco_graph_t *graph);

/* LCAdaptiveGrid(f, basis, LCAdaptiveSamplingParams(maxDepth <= N*level, ., level, .), evalCorners)
 * restricted to its bootstrap phase (LCAdaptiveGrid.cpp:16-39 -> bootstrapSample :256-456), wrapped in an
 * LCSampledFunction.  fn == NULL leaves the value table unset (use loco_model_fill_values_synthetic or
 * loco_model_set_values); the true centre values needed by loco_center_error are then unavailable. */
int loco_model_bootstrap(int32_t n_params, int32_t basis_type, int32_t level, int32_t value_dim, int32_t eval_corners,
                         const double *domain, loco_value_fn fn, void *user, int device, loco_model_t **out);

/* PrecomputedParametricShapeConverter::outputToProto(filename, useAscii, shape) (LCProtoConverter.cpp:14-24) */
int loco_model_save_proto(const loco_model_t *m, const char *path, int ascii);

void loco_free(loco_model_t *m);

/* ---- introspection ------------------------------------------------------------------------------- */
int loco_model_info(const loco_model_t *m, loco_info_t *info);

/* Sorted sample indices of LCAdaptiveGridCell::getAllInfluencingSamples for a leaf
 * (LCAdaptiveGridCell.cpp:558-607).  *k receives the count; LOCO_ERR_RANGE if cap is too small. */
int loco_support_set(const loco_model_t *m, int32_t leaf, int32_t *ids, int32_t cap, int32_t *k);

/* LCSampledFunction::getMinRange/getMaxRange (LCSampledFunction.cpp:52-59): ranges [2N] = min,max per parameter */
int loco_model_domain(const loco_model_t *m, double *ranges);

/* Introspection of the compacted term lists of a scalar model on generic leaves (no reference counterpart: the lists are
 * derived data, see DESIGN.md "merged terms" and "evaluation cells"; available on handles created with LOCO_HOST_ONLY,
 * device handles release the host copies after upload).
 * loco_merged_terms: the distinct basis functions of a leaf after merging equal (centre, support) pairs over its influencing
 *   samples (LCBasisFunctionKey equality, LCBasisFunction.h:85-104): centre[n*N], support[n*N]; *first_id = index of the
 *   leaf's first merged term in the model-wide numbering used by loco_eval_cell_terms.
 * loco_eval_cells: the leaf's sub-grid, 2^bits[d] cells per dimension (dimension 0 fastest), cells first_cell .. first_cell+n_cells-1.
 * loco_eval_cell_terms: model-wide merged-term ids listed for one evaluation cell. */
int loco_merged_terms(const loco_model_t *m, int32_t leaf, double *centre, double *support, int32_t cap, int32_t *n, int64_t *first_id);
int loco_eval_cells(const loco_model_t *m, int32_t leaf, int32_t *bits, int64_t *first_cell, int64_t *n_cells);
int loco_eval_cell_terms(const loco_model_t *m, int64_t cell, int64_t *term_ids, int32_t cap, int32_t *n);
/* polynomial form of an evaluation cell (SURVEY 8f.4 applied per cell, csrc/loco_cellpoly.h): *n_coef = F^N (0: the cell keeps
 * its term list), geom [2N] = mid_d, 1/halfwidth_d, coef [F^N] with dimension 0 fastest: value(x) = sum_j coef[j] prod_d t_d^(j_d),
 * t_d = (x_d - mid_d) * geom[2d+1] in cube coordinates.  Computed on the host from the handle's current values with the
 * same code the device kernel runs (host-only handles, scalar functions). */
int loco_eval_cell_poly(const loco_model_t *m, int64_t cell, double *geom, double *coef, int32_t cap, int32_t *n_coef);

/* LCAdaptiveGridCell::ranges_ of a leaf in cube coordinates: box [2N] = lo,hi per dimension (LCAdaptiveGridCell.h:89) */
int loco_leaf_box(const loco_model_t *m, int32_t leaf, double *box);

/* splitDirection_ of a leaf's parent cell (LCAdaptiveGridCell::getParent()->getSplitDirection(), what
 * LCAdaptiveGrid::decideSplitDirection's CYCLIC policy reads, LCAdaptiveGrid.cpp:553); -1 when the root is a leaf */
int loco_leaf_parent_split(const loco_model_t *m, int32_t leaf, int32_t *direction);

/* number of leaf / border-cell neighbours of a leaf (LCAdaptiveGridCell::getNeighbors, :93-96) */
int loco_leaf_neighbors(const loco_model_t *m, int32_t leaf, int32_t *n_leaf, int32_t *n_border);

/* description of the last failure on this handle (or, with m == NULL, of the last failed constructor
 * on this thread) -- LCError::internalDescription (LCUtils/LCError.cpp:13-16) */
const char *loco_last_error(const loco_model_t *m);
const char *loco_status_string(int status);

/* ---- value table ----------------------------------------------------------------------------------- */
/* row of the value table holding LCSample::shapeInfo_ of every sample: rows [n_samples].  Samples that share one
 * LCFunctionValue object (cubic border samples built with evalCorners=false, LCAdaptiveGrid.cpp:345-355) share a row;
 * rows not named here hold per-cell copies that differ from the sample's value and the leaves' centerShapeInfo_. */
int loco_sample_value_rows(const loco_model_t *m, int32_t *rows);
/* replace the value table: rows [n_value_rows * value_dim], host memory.  The handle's host copy (what
 * loco_model_save_proto writes) follows: tables up to 1 GiB are mirrored, larger ones are read back on demand. */
int loco_model_set_values(loco_model_t *m, const double *values);
/* fill the table on the device with v[row][j] = sin(0.37*j + 0.61*row + seed) + 0.001*(j % 97)
 * (synthetic mesh tables for benchmarks; never materialised on the host) */
int loco_model_fill_values_synthetic(loco_model_t *m, double seed);
/* copy rows [row0, row0+n) to host */
int loco_model_get_values(const loco_model_t *m, int64_t row0, int64_t n, double *out);
/* copy components [col0, col0+n) of EVERY row to host: out [n_value_rows * n] (spot checks of tables too large to read back) */
int loco_model_get_value_columns(const loco_model_t *m, int64_t col0, int64_t n, double *out);

/* ---- the hot path ------------------------------------------------------------------------------------
 * LCSampledFunction::evalShapeInfo(params, &result) (LCSampledFunction.cpp:69-78) for a batch:
 *   mapToStandardHypercube (LCFunction.cpp:59-73) -> getCointainingLeaf (LCAdaptiveGridCell.cpp:260-272)
 *   -> evalPametersAreInCell (:438-454) -> LCAdaptiveGridCell::evalShapeInfo (:611-676):
 *   getInterpolationWeight (LCSample.cpp:76-89) over LCLinearBSpline/LCCubicBSpline::eval
 *   (LCBasisFunction.cpp:273-293, 380-419) -> LCFunctionValue::newFromWeightedSum (LCFunctionValue.cpp:11-32).
 * q: [Q*N] row-major in ORIGINAL parameter space.  out: [Q*D] (NaN where status != 0).
 * leaf: [Q] containing leaf (or NULL).  status: [Q] LOCO_ST_* (or NULL). */

/* host buffers; host<->device copies are pipelined with the kernels inside the call */
int loco_eval_batch(loco_model_t *m, const double *q, int64_t n_queries, double *out, int32_t *leaf, uint8_t *status);

/* device buffers on the model's GPU; asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream) */
int loco_eval_batch_device(loco_model_t *m, const double *q, int64_t n_queries, double *out, int32_t *leaf,
                           uint8_t *status, void *cuda_stream);

/* mesh-valued batches too large to keep: evaluate but store only a per-query checksum
 *   sum_j out[q][j] * (1 + (j % 7))   -> checksum [Q] (NaN where status != 0); `ring` is a device scratch of
 * ring_queries * (D rounded up to even) doubles that the results are streamed through.  The WHOLE batch is sorted by
 * leaf once and then evaluated in ring-sized runs of full leaf tiles, so the ring holds results in leaf order; the call
 * synchronises the stream once (to split the sorted batch into runs). */
int loco_eval_batch_ring_device(loco_model_t *m, const double *q, int64_t n_queries, double *ring, int64_t ring_queries,
                                double *checksum, int32_t *leaf, uint8_t *status, void *cuda_stream);

/* the same from HOST buffers (q in, checksum / leaf / status out; copies inside the call): the ring is an internal device
 * scratch of ring_queries results (0 = a default sized for ~16 GiB, at least one 64-query tile per leaf run) */
int loco_eval_batch_ring(loco_model_t *m, const double *q, int64_t n_queries, int64_t ring_queries, double *checksum, int32_t *leaf,
                         uint8_t *status);

/* ---- derivative ------------------------------------------------------------------------------------------
 * LCSampledFunction::evalDeriv(params, direction, &result) (LCSampledFunction.cpp:61-67) for a batch:
 *   LCAdaptiveGridCell::evalDeriv (LCAdaptiveGridCell.cpp:678-711) -> LCSample::getDerivInterpolationWeight
 *   (LCSample.cpp:92-105) -> LCLinearBSpline/LCCubicBSpline::evalDeriv (LCBasisFunction.cpp:295-332, 421-483)
 *   -> LCFunctionDeriv::newFromWeightedSum (LCFunctionDeriv.cpp:12-35; LCRealFunctionValue only => D == 1).
 * The derivative is taken along CUBE coordinate `direction`, as in the reference.  The linear basis follows the
 * reference's formula literally, including its use of the signed distance (see loco_basis.cuh). */
int loco_eval_deriv_batch(loco_model_t *m, const double *q, int64_t n_queries, int32_t direction, double *out, int32_t *leaf,
                          uint8_t *status);
int loco_eval_deriv_batch_device(loco_model_t *m, const double *q, int64_t n_queries, int32_t direction, double *out, int32_t *leaf,
                                 uint8_t *status, void *cuda_stream);

/* ---- refinement-loop error check ---------------------------------------------------------------------
 * The data-parallel part of LCAdaptiveGrid::decideSplit (LCAdaptiveGrid.cpp:578-626) with
 * LCRealFunctionValue::computeErrorVector (LCRealFunctionValue.cpp:37-47): evaluate the interpolant at P
 * points, err = |approx - truth| component-wise, per containing leaf keep the max error and
 * split = !(max(err - threshold) <= 0) OR-reduced (NaN => split).
 * q [P*N] original space, truth [P*D]; err_max [L] and split [L] are OVERWRITTEN (leaves without points:
 * 0 / 0).  Points whose status != 0 are skipped and counted in *n_bad (may be NULL). */
int loco_error_check(loco_model_t *m, const double *q, const double *truth, int64_t n_points, double threshold,
                     double *err_max, uint8_t *split, int64_t *n_bad);
int loco_error_check_device(loco_model_t *m, const double *q, const double *truth, int64_t n_points, double threshold,
                            double *err_max, uint8_t *split, int64_t *n_bad_dev, void *cuda_stream);

/* CENTER_BASED metric (LCAdaptiveGrid.cpp:594-602, LCAdaptiveGridCell::getApproximationError :1088-1100):
 * interpolant at every leaf centre against the stored centerShapeInfo_.  err_max/split: [L] host. */
int loco_center_error(loco_model_t *m, double threshold, double *err_max, uint8_t *split);

/* RANDOM_SAMPLES metric at scale (LCAdaptiveGrid.cpp:603-619 with getJitterPoint :564-575): points_per_leaf
 * jitter points per leaf are generated ON THE DEVICE from a counter-based generator (leaf, i, seed), the
 * truth is the reference's test function family f(x) = 1 + sum_i sum_j a[i][j] * sin(3*j*x_i + p[i][j])
 * (LCRealFunction.cpp:50-63) with coefficients amp/phase [N*n_terms] supplied by the caller.
 * err_max/split: [L] device.
 * The points are a documented function of (seed, leaf, i) so that a host can regenerate them: point p = leaf*points_per_leaf + i,
 * dimension d:  h = splitmix64(seed ^ splitmix64(p*N + d));  r = (float)(h >> 40) / 16777215.0f;
 * pos_d = lo_d + (double)r * (hi_d - lo_d) in cube coordinates (unfused double operations, lo/hi = loco_leaf_box).
 * Scalar functions on tensor-product leaves run as ONE kernel without intermediates in HBM (option "fused_errgen", default 1). */
int loco_error_check_generated_device(loco_model_t *m, int64_t points_per_leaf, uint64_t seed, const double *amp,
                                      const double *phase, int32_t n_terms, double threshold, double *err_max,
                                      uint8_t *split, void *cuda_stream);

/* ---- multi-GPU result exchange ---------------------------------------------------------------------------
 * One process per GPU, the model replicated, the query batch sharded.  The per-shard results travel to the
 * consuming GPU over NVLink through the COPY ENGINES, not through SM kernels: the consumer exports a device buffer, every
 * producer maps it (CUDA IPC) and pushes its shard with an asynchronous peer copy on a side stream while its SMs evaluate the
 * next batch.  (locointerpolate_b200/shard.py: RotatingExchange drives these; NCCL stays for the completion signal and the
 * error-maxima all-reduce.)  No model handle: errors are reported through loco_last_error(NULL).
 *   create : cudaMalloc on `device` + export; handle = 64 opaque bytes to send to the peers
 *   open   : map a peer's buffer into this process (peer access is enabled lazily by the runtime); *dptr is valid on `device`
 *   copy   : cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault) on `cuda_stream` -- device-to-device, either side may be a mapped peer buffer */
int loco_peer_buffer_create(int device, size_t bytes, void **dptr, unsigned char handle[64]);
int loco_peer_buffer_open(int device, const unsigned char handle[64], void **dptr);
int loco_peer_buffer_close(int device, void *dptr);
int loco_peer_buffer_free(int device, void *dptr);
int loco_peer_copy_async(void *dst, const void *src, size_t bytes, void *cuda_stream);

/* ---- tuning / accounting -------------------------------------------------------------------------------- */
/* options: "lut" 1 locate grid before the k-d descent (default) | 0 descent from the root only;
 *          "path" 0 auto | 1 direct (thread per query) | 2 leaf-sorted TMA tiles | 3 direct without the tensor-leaf shortcut |
 *                 4 L2-resident leaf-sorted chunks (exact counting sort) | 5 single-pass fixed-capacity leaf buckets (tensor leaves);  "chunk" queries per chunk of path 4 (0 = auto);
 *          "overlap" 1..4 (default 2): internal streams the chunks of path 4 are dealt over (forked from / joined to the caller's);
 *          "fp32" 1: optional fp32 mode of the throughput path for scalar functions on tensor-product leaves (path 0/4): point
 *                 location stays fp64 and bit-exact, basis polynomials and the weighted sum run in fp32 (1e-5 relative);
 *          "uniform" 1 (default): cubic tensor leaves with equally spaced centres form their four factors per dimension from one parameter;
 *          "mesh_simt" 1: mesh-valued tiles with DFMA register tiles instead of the fp64 MMA pipe (default 0);
 *          "ecells" 1 (default): path 4 on generic leaves sorts by evaluation cell (a regular sub-grid of the leaves with many terms,
 *                 each cell listing only the basis functions that reach it) | 0 by leaf;
 *          "leaf_poly" 1 (default): scalar functions on tensor-product leaves that lie inside one polynomial piece of every basis function
 *                 (every bootstrap grid) evaluate the leaf's single tensor-product polynomial by nested Horner (changes the
 *                 summation order, within the 1e-12 bar) | 0: N*F basis polynomials + sum-factorised contraction;
 *          "fast_scatter" 1 (default): path 5's scatter kernel specialised for models with power-of-two domain widths, a uniform dyadic
 *                 locate grid and split-plane-consistent leaf boxes (every bootstrap grid); same arithmetic | 0: the general kernel;
 *          "bulk_eval" 1: the evaluation kernel of path 5 works on items of up to 96 records of one leaf that the TMA engine
 *                 prefetches into shared memory an item ahead (cp.async.bulk + mbarrier) | 0 (default; measured equal): per-lane record loads one group ahead;
 *          "aggregate" 1 (default): the sort kernels of paths 4 / 5 issue one slot atomic per group of lanes of a warp that share a leaf | 0: one per query;
 *          "adaptive" 1 (default): with "path" 0 big scalar batches on tensor-product leaves take the leaf buckets (path 5) while the batches
 *                 seen so far fit the buckets' uniform-batch capacity, and the exact counting sort (path 4) once more than 4 % of the queries
 *                 overflow (skewed batches); decided from counters the kernels keep, without synchronisation | 0: always the buckets;
 *          "fused_errgen" 1 (default): loco_error_check_generated_device as one fused kernel where supported | 0: generate / evaluate / reduce kernels;
 *          "ecell_poly" 1 (default): evaluation cells that no knot crosses are evaluated from their single tensor-product polynomial
 *                 (changes the summation order, within the 1e-12 bar) | 0: always the term list */
int loco_set_option(loco_model_t *m, const char *name, int64_t value);
/* kernels launched by this handle since creation (for bench accounting) */
int64_t loco_kernel_launches(const loco_model_t *m);

/* per-kernel device times (CUDA events on the launching stream around every kernel this library launches, from any
 * handle of this process) between begin and end; end synchronises the device and writes a JSON object
 * {"kernel name": {"launches": n, "total_ms": t}, ...} into `json`. */
int loco_profile_begin(loco_model_t *m);
int loco_profile_end(loco_model_t *m, char *json, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* LOCO_B200_H */
