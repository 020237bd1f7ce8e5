// loco_flat.h -- the flattened, GPU-ready model (host image) and the flattener.
//
// Everything the reference recomputes per query from its pointer graph
//   * the leaf's influencing-sample set   (LCAdaptiveGridCell::getAllInfluencingSamples, LCAdaptiveGridCell.cpp:558-607)
//   * which basis functions can be non-zero on the leaf (LCSample::getInterpolationWeight, LCSample.cpp:76-89)
//   * the +-EPSILON containment bounds     (LCAdaptiveGridCell::evalPametersAreInCell, :438-454)
// is a pure function of the leaf and is therefore computed ONCE here, into SoA arrays that are
// uploaded verbatim.  Numbering is canonical: cells in pre-order, leaf index = rank
// in getAllLeafCells order (:247-258), sample index = position in samples_.
#pragma once
#include "loco_graph.h"

namespace loco {

const double kEpsilon = 0.000001;             // LCMathHelper::EPSILON (LCMathHelper.cpp:84)
const double kContainingEpsilon = 0.0000001;  // LCAdaptiveGridCell::CONTAINING_EPSILON (LCAdaptiveGridCell.h:109)

// per-query status byte
enum QueryStatus : uint8_t {
	ST_OK = 0,
	ST_OUTSIDE_CELL = 1,      // "parameter ouside cell"                   LCAdaptiveGridCell.cpp:450
	ST_BAD_ARITY = 2,         // "invalid number of parameters"           LCAdaptiveGridCell.cpp:441-444 (host facade only)
	ST_NO_COMMON_SAMPLE = 3,  // "could not find closest sample"          LCAdaptiveGridCell.cpp:533-538
	ST_EMPTY_SUPPORT = 4,     // "cannot create new from empty list ..."  LCFunctionValue.cpp:14-17
};
const char *status_string(int code);

// One INNER k-d tree cell as the kernels read it: one 16-byte load per level.  Children are stored as
// codes so that reaching a leaf needs no further load: code = kLeafFlag | leaf index, or the index of the
// inner cell in `cells` (inner cells only, in pre-order).  a = (splitDirection_ << 28) | code(child1).
const uint32_t kLeafFlag = 0x08000000u;
const uint32_t kCodeMask = 0x0fffffffu;
struct alignas(16) TreeCell {
	double split;  // splitVal_
	uint32_t a;    // dim << 28 | code(child1)
	uint32_t b;    // code(child2)
};

struct Flat {
	int32_t N = 0, D = 1, basis = 0;
	int64_t S = 0;  // samples
	int64_t R = 0;  // value rows
	int64_t L = 0;  // leaves
	int64_t C = 0;  // tree cells
	int32_t depth = 0;       // deepest leaf level
	int32_t max_support = 0; // max K over leaves
	int32_t max_terms = 0;   // max T over leaves
	bool exact_rcp = false;  // every support is a power of two: term_cs holds 1/s and u=(x-c)*(1/s) is bit-identical to (x-c)/s

	std::vector<double> dom_min, dom_max;  // [N] ranges_[2i], ranges_[2i+1]
	std::vector<TreeCell> cells;           // [#inner cells] pre-order; empty when the root is a leaf
	std::vector<double> leaf_lo, leaf_hi;          // [L*N] leaf box
	std::vector<double> leaf_lo_eps, leaf_hi_eps;  // [L*N] lo-EPSILON, hi+EPSILON as the reference computes them
	std::vector<int32_t> leaf_center_row;          // [L] value row of centerShapeInfo_ (true f at the centre)
	std::vector<uint8_t> leaf_status;              // [L] ST_OK, or the error every query in this leaf gets

	// influencing-sample sets, CSR by leaf, sorted by sample index (the "support set" parity object)
	std::vector<int64_t> sup_off;      // [L+1]
	std::vector<int32_t> sup_sample;   // [sum K]
	std::vector<int32_t> sup_row;      // [sum K] value row used for that (leaf,sample) pair

	// terms, CSR by leaf: every basis function of every influencing sample whose box can reach the leaf
	std::vector<int64_t> term_off;     // [L+1]
	std::vector<double> term_cs;       // [sum T * 2N]  (c_0, s_0, c_1, s_1, ...) ; s or 1/s, see exact_rcp
	std::vector<double> term_w;        // [sum T] weight_
	std::vector<int32_t> term_slot;    // [sum T] position of the owning sample in the leaf's support list (ascending)
	std::vector<int32_t> term_row;     // [sum T] = sup_row[sup_off[leaf] + slot]

	// merged terms (scalar functions): within a leaf, basis functions with bit-identical (centre, support) owned by different
	// samples are one term  (sum_e weight_e * val_e) * B(x)  -- refineCell re-attaches equal functions to several samples
	// (LCAdaptiveGrid.cpp:825-862), so adaptive trees carry each function ~3 times.  Built only when it saves >= 25 %.
	std::vector<int64_t> mterm_off;    // [L+1] CSR by leaf into the merged terms
	std::vector<double> mterm_cs;      // [M * 2N]
	std::vector<int64_t> mrun_off;     // [M+1] CSR: the original terms merged into each
	std::vector<int32_t> mrun_term;    // [sum T] indices into the term arrays above

	// evaluation cells (scalar functions on generic leaves): in the reference's adaptive trees most basis functions that
	// overlap a leaf are much NARROWER than the leaf (refineCell refines in all directions), so at any one point only a
	// fraction of the leaf's terms is non-zero.  Leaves with many merged terms are cut into a regular grid of
	// 2^bits[d] cells per dimension; every cell lists the merged terms whose box reaches it (with a margin that covers
	// the +-1e-6 acceptance band and the rounding of the cell index, so the choice of cell never changes a result).
	// The throughput path sorts queries by evaluation cell instead of by leaf.
	std::vector<int64_t> ecell_off;    // [L+1] first evaluation cell of each leaf
	std::vector<uint8_t> ecell_bits;   // [L * N]
	std::vector<int64_t> eterm_off;    // [E+1] CSR by evaluation cell
	std::vector<double> eterm_cs;      // [sum * 2N]
	std::vector<int32_t> eterm_src;    // [sum] index of the merged term (for weight * value)

	// polynomial evaluation cells (loco_cellpoly.h): cells no knot of a listed basis function crosses and that do not touch
	// the root box (where the +-1e-6 band would leave the piece).  The coefficients are formed from the current values
	// (device: epoly_kernel; host: loco_eval_cell_poly for the tests); here only the selection and the geometry.
	std::vector<int32_t> epoly_idx;    // [E] index of the cell's polynomial block, -1: term list
	std::vector<int32_t> epoly_cell;   // [Pc]
	std::vector<double> epoly_geom;    // [Pc * 2N] mid_d, 1 / halfwidth_d

	// derivative-only extra terms (linear basis; see loco_flatten.cpp): hats outside the leaf that the reference's
	// LCLinearBSpline::evalDeriv still counts because it compares the SIGNED distance with 1
	std::vector<int64_t> dterm_off;    // [L+1]
	std::vector<double> dterm_cs;      // [sum * 2N]
	std::vector<double> dterm_w;       // [sum]
	std::vector<int32_t> dterm_row;    // [sum]

	std::vector<int32_t> sample_row;   // [S]
	std::vector<double> values;        // [R*D] (may be empty when the table is filled on the device)

	// ---- tensor-product leaves ---------------------------------------------------------------------
	// On bootstrap-uniform grids every leaf's term list is a full tensor product: one unit-weight basis
	// function per influencing sample, F distinct centres per dimension (F = 2 linear / 4 cubic), one
	// common support per dimension, all F^N combinations present.  Then w_k = prod_d b((x_d - c_d[j_d]) / s_d)
	// needs only N*F one-dimensional evaluations instead of N*F^N.  tensor_F > 0 iff EVERY leaf has this
	// structure with the same F; the generic term lists above stay valid either way.
	int32_t tensor_F = 0;
	int32_t tensor_K = 0;                 // F^N
	// ---- incremental re-flatten.  Per-leaf signature of everything a leaf's evaluation cells and polynomial
	// cells are a function of (leaf box, root box, the leaf's merged term list, the global cell scale): flatten(g, out, prev)
	// copies that derived data -- 90 % of the flatten time on the reference's adaptively refined trees -- from the leaf of
	// `prev` with the same signature instead of recomputing it.  After LCAdaptiveGrid::refineCell only the leaves in the
	// split cell's doubly-adjacent set (LCAdaptiveGrid.cpp:174) change their term lists, so everything else is reused.
	std::vector<uint64_t> leaf_sig;       // [2*L]
	double ecell_scale = 0.0;             // cell scale the evaluation cells were built with (0: none)
	int64_t reused_leaves = 0;            // leaves whose derived data came from `prev` in the last flatten
	bool tensor_poly = false;             // every leaf lies inside ONE polynomial piece of each of its F per-dimension basis functions (no knot strictly
	                                      // inside the leaf box): the interpolant on the leaf is a single tensor-product polynomial
	bool tensor_uniform = false;          // cubic: in every leaf and dimension the 4 centres are spaced by exactly support/2 (one uniform knot vector)
	std::vector<double> tl_center;        // [L * N * F] ascending per dimension
	std::vector<double> tl_rcp;           // [L * N]     support (or 1/support when exact_rcp)
	std::vector<int32_t> tl_slot;         // [L * F^N]   slot in the leaf's sorted support list, tensor order (dimension 0 fastest)
	std::vector<int32_t> tl_row;          // [L * F^N]   value row, same order

	// ---- folded tensor leaves (mesh-valued path) --------------------------------------------------------------------
	// Samples that share one LCFunctionValue object (cubic border samples built with evalCorners=false share the value
	// of their clamped in-domain sample, LCAdaptiveGrid.cpp:345-355) carry the same value row.  On a tensor leaf that
	// sharing is itself a tensor product: in dimension d, digits j and j' are equivalent when swapping them never changes
	// the row.  Then  sum_k w_k V[row_k]  =  sum over CLASSES of (prod_d g_d[class_d]) V[row],  g_d[c] = sum of the
	// factors of the digits in class c -- the same sum re-associated, with prod_d n_d instead of F^N rows.
	std::vector<int64_t> fold_off;        // [L+1] CSR into fold_row
	std::vector<int32_t> fold_row;        // value rows of the folded tensor, class index of dimension 0 fastest
	std::vector<uint8_t> fold_map;        // [L * N * F] digit -> class index in its dimension
	std::vector<uint8_t> fold_n;          // [L * N] classes per dimension (1..F)
	int32_t max_fold_K = 0;

	// every leaf box equals the box implied by the split planes on its root path: the +-EPSILON containment
	// test can then only fail at the root box, and no per-leaf bounds need to be read per query
	bool path_consistent = false;
	bool any_leaf_status = false;

	// ---- locate grid ---------------------------------------------------------------------------------------
	// A regular grid of prod_d 2^lut_bits[d] cells over the root box; entry = code (see TreeCell) of the DEEPEST
	// tree cell whose region -- as defined by the split comparisons on its root path -- contains the whole grid
	// cell.  A query finds its grid cell with exact comparisons against the stored boundaries and resumes the
	// descent from that tree cell: on trees whose top levels are balanced (every bootstrap grid) the k-d descent
	// of getCointainingLeaf collapses to one load.  Outer grid cells extend to +-infinity, NaN coordinates map
	// to the last cell (the reference's `x < split` is false for NaN: always child2).
	int32_t lut_bits[16] = {0};
	std::vector<uint32_t> lut;         // [prod 2^bits], dimension 0 fastest; empty when the root is a leaf
	std::vector<double> lut_bounds;    // per dimension 2^bits+1 boundaries, concatenated
	int32_t lut_boff[16] = {0};        // offset of dimension d in lut_bounds

	// leaf adjacency kept for introspection / tests
	std::vector<int32_t> n_leaf_neighbors, n_border_neighbors;  // [L]
};

// prev: the flat model of an earlier state of the same grid (may be null): leaves whose signature is unchanged reuse its derived data
Status flatten(const Graph &g, Flat *out, const Flat *prev = nullptr);

} // namespace loco
