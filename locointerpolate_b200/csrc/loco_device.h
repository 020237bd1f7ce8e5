// loco_device.h -- device image of the flat model and the kernel launch interface.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "loco_flat.h"

namespace loco {

const int kMaxTemplateN = 8;   // N = 1..8 get fully unrolled kernels
const int kMaxN = 16;          // larger N use the runtime-N instantiation (coordinates in local memory)

// Pointers into HBM; passed by value to every kernel.  Layout documented in DESIGN.md ("HBM layout").
struct DeviceModel {
	int32_t N, D, basis, exact_rcp;
	int32_t L, C, n_inner, max_support, max_terms;
	int64_t R, value_stride;          // rows and row stride (doubles) of the value table
	const double *dom_min, *dom_max;  // [N]
	// the same in the kernel parameter bank: min, max-min (the reference's own subtraction) and, where max-min is a
	// power of two, its exact reciprocal: (p-min)*(1/w) is then bit-identical to (p-min)/w
	double dom_lo[kMaxN], dom_w[kMaxN], dom_rcp[kMaxN];
	uint32_t dom_exact_mask;
	const TreeCell *cells;            // [n_inner] inner cells of the k-d tree in pre-order, 16 B each
	const double *leaf_lo_eps, *leaf_hi_eps;  // [L*N]
	const double *leaf_lo, *leaf_hi;          // [L*N]
	const uint8_t *leaf_status;       // [L]
	const int32_t *leaf_center_row;   // [L]
	const int32_t *sup_off;           // [L+1]
	const int32_t *sup_row;           // [sum K]
	const int32_t *term_off;          // [L+1]
	const double *term_cs;            // [sum T * 2N]
	const double *term_w;             // [sum T]
	const int32_t *term_slot;         // [sum T]
	const int32_t *term_row;          // [sum T]
	const double *term_wv;            // [sum T + 2] weight * value of the owning sample (scalar functions only)
	const double *values;             // [R * value_stride]
	// merged terms (Flat::mterm_*; mterm_off == nullptr: not built).  mterm_wv[t] = sum over the run of weight * value.
	const int32_t *mterm_off;
	const double *mterm_cs, *mterm_wv;
	const int32_t *mrun_off, *mrun_term;
	int64_t n_mterms;
	// evaluation cells (Flat::ecell_*; ecell_off == nullptr: not built).  eterm_wv[t] = mterm_wv[eterm_src[t]].
	const int32_t *ecell_off;
	const uint8_t *ecell_bits;
	const int32_t *eterm_off;
	const double *eterm_cs, *eterm_wv;
	const int32_t *eterm_src;
	int32_t n_ecells;
	int64_t n_eterms;
	// polynomial evaluation cells (loco_cellpoly.h).  epoly_idx == nullptr: not built, or switched off (option "ecell_poly").
	// Block of a cell: [2N geometry: mid_d, 1/halfwidth_d | F^N coefficients, dimension 0 fastest], epoly_stride doubles.
	const int32_t *epoly_idx;   // [E] block index or -1
	const double *epoly;
	int32_t epoly_stride;
	// derivative-only extra terms of the linear basis (CSR by leaf; see Flat::dterm_*)
	const int32_t *dterm_off;
	const double *dterm_cs, *dterm_w;
	const int32_t *dterm_row;
	// tensor-product leaves (tensor_F > 0): one block of tl_stride doubles per leaf,
	//   [N*F centres | N supports-or-reciprocals | pad to even | F^N sample values in tensor order (D == 1 only)]
	int32_t tensor_F, tensor_K, tl_stride, tl_voff, tensor_uniform;
	const double *tl_block;           // [L * tl_stride]
	const float *tl_vals_f32;         // [L * F^N] fp32 copy of the leaf values in tensor order (optional fp32 mode, D == 1)
	// per-leaf polynomial form (D == 1, Flat::tensor_poly; nullptr: not built or switched off by option "leaf_poly"):
	//   [N x (mid_d, 1/halfwidth_d) | F^N coefficients, dimension 0 fastest], tp_stride doubles per leaf;
	//   value(x) = sum_j coef[j] prod_d t_d^(j_d),  t_d = (x_d - mid_d) / halfwidth_d in [-1, 1] inside the leaf
	const double *tl_poly;
	int32_t tp_stride;
	const int32_t *tl_row;            // [L * F^N] value rows in tensor order
	// locate grid (Flat::lut): lut == nullptr disables it
	const uint32_t *lut;
	const double *lut_bounds;
	int32_t lut_bits[kMaxN], lut_boff[kMaxN];
	double lut_lo[kMaxN], lut_inv[kMaxN];   // estimate of the grid cell: (x - lo) * inv, then exact comparisons
	// dimensions whose grid boundaries are EXACTLY lo + i*h with h a power of two (every bootstrap grid): the boundary next to
	// the estimate is formed by one exact fma instead of two table loads
	double lut_h[kMaxN];
	uint32_t lut_uniform_mask;
	// folded tensor leaves (Flat::fold_*): value rows after merging digits that share rows
	const int32_t *fold_off;          // [L+1]
	const int32_t *fold_row;
	const uint8_t *fold_map;          // [L*N*F]
	const uint8_t *fold_n;            // [L*N]
	// containment against the root box only (valid when every leaf box is implied by its split planes)
	int32_t path_consistent, any_leaf_status;
	double root_lo_eps[kMaxN], root_hi_eps[kMaxN];
};


// ---- per-kernel timing (bench.py's roofline leg): CUDA events around every launch of a named kernel, recorded on
// the launching stream while profiling is switched on (loco_profile_begin / loco_profile_end)
struct KernelProfile {
	bool on = false;
	struct Rec { const char *name; cudaEvent_t a, b; };
	Rec *recs = nullptr;
	int n = 0, cap = 0;
	void push(const char *name, cudaEvent_t a, cudaEvent_t b);
};
KernelProfile &kernel_profile();
struct ScopedKernelTime {
	cudaEvent_t a = nullptr, b = nullptr;
	cudaStream_t st;
	const char *name;
	ScopedKernelTime(const char *name_, cudaStream_t st_) : st(st_), name(name_)
	{
		if (!kernel_profile().on) return;
		cudaEventCreate(&a); cudaEventCreate(&b);
		cudaEventRecord(a, st);
	}
	~ScopedKernelTime()
	{
		if (!a) return;
		cudaEventRecord(b, st);
		kernel_profile().push(name, a, b);
	}
};
#define LOCO_KERNEL_CAT2(a, b) a##b
#define LOCO_KERNEL_CAT(a, b) LOCO_KERNEL_CAT2(a, b)
#define LOCO_KERNEL(name, st) ::loco::ScopedKernelTime LOCO_KERNEL_CAT(_loco_kt_, __LINE__)(name, st)

// ---- launchers (loco_kernels.cu).  All asynchronous on `st`; return the number of kernels launched.
// fused map -> locate -> weights -> weighted sum for scalar functions (D == 1), one thread per query
int launch_eval_scalar_direct(const DeviceModel &m, const double *q, int64_t Q, double *out, int32_t *leaf,
                              uint8_t *status, cudaStream_t st);
// same, but q holds CUBE coordinates and leaf[] the cell to evaluate in (evaluation on a known cell)
int launch_eval_scalar_in_cube(const DeviceModel &m, const double *x, int64_t Q, double *out, int32_t *leaf,
                               uint8_t *status, cudaStream_t st);
// d/dx_direction of the interpolant in CUBE coordinates (LCSampledFunction::evalDeriv), scalar functions
int launch_eval_deriv_direct(const DeviceModel &m, const double *q, int64_t Q, int direction, double *out, int32_t *leaf,
                             uint8_t *status, cudaStream_t st);
// locate only: cube coordinates xs [Q*N] (may be null), leaf, status
int launch_locate(const DeviceModel &m, const double *q, int64_t Q, double *xs, int32_t *leaf, uint8_t *status, cudaStream_t st);
// per-sample weights of each query's leaf: W [Q * max_support] (slots beyond K are zero)
int launch_weights(const DeviceModel &m, const double *q, int64_t Q, double *W, int32_t *leaf, uint8_t *status, cudaStream_t st);
int launch_weights_in_cube(const DeviceModel &m, const double *x, int64_t Q, double *W, int32_t *leaf, uint8_t *status, cudaStream_t st);
// out[q][:] = sum_k W[q][k] * values[sup_row[leaf][k]][:]
int launch_mesh_gather(const DeviceModel &m, const double *W, const int32_t *leaf, const uint8_t *status, int64_t Q,
                       double *out, int64_t out_stride, cudaStream_t st);
// checksum[q] = sum_j out[q][j] * (1 + j % 7)
int launch_checksum(const double *out, int64_t out_stride, int32_t D, int64_t Q, double *checksum, cudaStream_t st);
// error check reduction: per-leaf max |out - truth| and split flag
int launch_error_reduce(const DeviceModel &m, const double *out, const double *truth, const int32_t *leaf, const uint8_t *status,
                        int64_t P, double threshold, double *err_max, uint8_t *split, unsigned long long *n_bad, cudaStream_t st);
int launch_fill_values(double *values, int64_t R, int64_t D, int64_t stride, double seed, cudaStream_t st);
int launch_leaf_centers(const DeviceModel &m, double *q_out, cudaStream_t st);
int launch_gather_rows(const DeviceModel &m, const int32_t *rows, int64_t n, double *out, cudaStream_t st);
int launch_jitter_points(const DeviceModel &m, int64_t points_per_leaf, uint64_t seed, int64_t first, int64_t count,
                         const double *amp, const double *phase, int32_t n_terms, double *q_out, int32_t *leaf_out,
                         double *truth_out, cudaStream_t st);

// RANDOM_SAMPLES error check in one kernel (loco_errgen.cu): generate -> evaluate in the cell -> |approx - truth| -> per-leaf max
bool error_check_generated_fused_supported(const DeviceModel &m, int n_terms);
int launch_error_check_generated_fused(const DeviceModel &m, int64_t ppl, uint64_t seed, const double *amp, const double *phase, int n_terms, double threshold,
                                       double *err_max, uint8_t *split, int sm_count, cudaStream_t st);

// ---- leaf-sorted tile path (loco_tiles.cu)
struct SortWork {
	int32_t *count;     // [L] queries per leaf
	int32_t *offset;    // [L+1] bucket offsets
	int32_t *tile_off;  // [L+1] tile offsets
	int32_t *rank;      // [Q] arrival rank of each query inside its leaf bucket
	int32_t *idx;       // unused (kept for layout compatibility)
	int32_t *tile_leaf; // [Q / 32 + L + 1] leaf of every tile (so that a tile's leaf is one load, not a binary search over tile_off)
	double *xs;         // [Q * row] sorted query records: N cube coordinates + original index, 32-byte sectors
};
size_t sort_work_bytes(const DeviceModel &m, int64_t Q);
SortWork carve_sort_work(const DeviceModel &m, int64_t Q, void *base);
int launch_update_term_wv(const DeviceModel &m, double *term_wv, int64_t T, cudaStream_t st);
int launch_update_mterm_wv(const DeviceModel &m, double *mterm_wv, cudaStream_t st);
int launch_update_eterm_wv(const DeviceModel &m, double *eterm_wv, cudaStream_t st);
int launch_update_epoly(const DeviceModel &m, const int32_t *epoly_cell, int32_t n_epoly, double *epoly, cudaStream_t st);
int launch_update_tensor_vals(const DeviceModel &m, double *tl_block, float *tl_vals_f32, cudaStream_t st);
int launch_update_tensor_poly(const DeviceModel &m, double *tl_poly, int32_t tp_stride, cudaStream_t st);
int launch_eval_scalar_tensor_direct(const DeviceModel &m, bool in_cube, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status,
                                     cudaStream_t st);
int launch_sort_by_leaf(const DeviceModel &m, const double *q, int64_t Q, int32_t *leaf, uint8_t *status, const SortWork &w, int tile_q,
                        double *scalar_out, cudaStream_t st);
int mesh_tile_queries();
int launch_mesh_tensor_tiles(const DeviceModel &m, const SortWork &w, const uint8_t *status, int64_t Q, double *out, int64_t out_stride, int sm_count,
                             cudaStream_t st);
int launch_mesh_nan_rows(const DeviceModel &m, const uint8_t *status, int64_t Q, double *out, int64_t out_stride, cudaStream_t st);
// fp64 MMA version of the mesh tile kernel (loco_mesh_mma.cu); same sorted input, tiles of mesh_mma_tile_queries()
int mesh_mma_tile_queries();
bool mesh_mma_supported(const DeviceModel &m);
int launch_mesh_tensor_mma(const DeviceModel &m, const SortWork &w, int64_t Q, double *out, int64_t out_stride, int sm_count, int64_t tile0, int64_t n_tiles,
                           int64_t ring_base, cudaStream_t st);
// checksum[query index of sorted position p] = sum_j ring[p - p0][j] * (1 + j % 7) for p in [p0, p0 + n); queries with an error status get NaN
int launch_checksum_sorted(const DeviceModel &m, const double *ring, int64_t ring_stride, const SortWork &w, int64_t p0, int64_t n, double *checksum, cudaStream_t st);
int launch_checksum_bad(const uint8_t *status, int64_t Q, double *checksum, cudaStream_t st);
// fp64 MMA tiles for mesh-valued functions on GENERIC leaves (loco_mesh_gen.cu): weights from the term lists into global memory,
// then the same TMA / DMMA tile pipeline; same sorted input, tiles of mesh_generic_tile_queries()
int mesh_generic_tile_queries();
size_t mesh_generic_weight_doubles(const DeviceModel &m, int64_t Q);
int launch_mesh_generic_mma(const DeviceModel &m, const SortWork &w, const int32_t *leaf_of, int64_t Q, int64_t *woff, double *W, double *out,
                            int64_t out_stride, int sm_count, int64_t tile0, int64_t n_tiles, int64_t ring_base, bool weights_ready, cudaStream_t st);
int launch_eval_scalar_tiles(const DeviceModel &m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, const SortWork &w,
                             int sm_count, cudaStream_t st);

// ---- L2-resident sorted chunks (loco_sorted.cu)
// Skew statistics of a workspace slot (host-mapped, written by ONE thread per chunk, read by the host between calls):
// queries that did not fit / would not have fit their leaf's fixed-capacity bucket, and queries seen.  eval_device() uses
// them to pick the leaf-bucket path (uniform batches) or the exact counting sort (skewed batches) for the NEXT call.
struct SkewStats { long long overflow, total; };

struct SortedWork {
	bool aggregate;    // warp-aggregated histogram / cursor atomics (option "aggregate")
	SkewStats *stats;  // may be null
	int32_t stats_cap; // bucket capacity the overflow is counted against
	int32_t *count;    // [L] histogram (zero between chunks)
	int32_t *cursor;   // [L+1] bucket cursors; [L] = number of records of the chunk
	unsigned *done;    // arrival counter of the count kernel
	double *xs;        // [chunk * record] leaf-ordered query records
};
size_t sorted_work_bytes(const DeviceModel &m, int64_t chunk);
SortedWork carve_sorted_work(const DeviceModel &m, void *base);
int launch_sorted_reset(const DeviceModel &m, const SortedWork &w, cudaStream_t st);
int launch_eval_scalar_sorted_chunk(const DeviceModel &m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, const SortedWork &w,
                                    int sm_count, bool fp32, cudaStream_t st, int deriv_dir = -1);

// ---- single-pass fixed-capacity leaf buckets (loco_bucket.cu)
struct BucketWork {
	SkewStats *stats;  // may be null
	bool aggregate;    // warp-aggregated slot atomics (option "aggregate")
	bool bulk;         // evaluation with TMA-prefetched work items (option "bulk_eval")
	bool fast;         // specialised scatter kernel where the model allows it (option "fast_scatter")
	int32_t *count;    // [L] queries that asked for a slot in the leaf's bucket (zero between chunks)
	int32_t *n_ovf;    // number of overflow entries
	int32_t *ovf;      // [chunk] query indices whose bucket was full
	double *rec;       // [L * cap * record] records, bucket-major
	int32_t cap;       // slots per bucket
};
int bucket_capacity(const DeviceModel &m, int64_t chunk);
bool bucket_path_supported(const DeviceModel &m);
size_t bucket_work_bytes(const DeviceModel &m, int64_t chunk);
BucketWork carve_bucket_work(const DeviceModel &m, int64_t chunk, void *base);
int launch_bucket_reset(const DeviceModel &m, const BucketWork &w, cudaStream_t st);
int launch_eval_scalar_bucket_chunk(const DeviceModel &m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, const BucketWork &w,
                                    int sm_count, bool fp32, int parity, cudaStream_t st, int deriv_dir = -1);

} // namespace loco
