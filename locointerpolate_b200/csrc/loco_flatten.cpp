// loco_flatten.cpp -- Graph -> Flat.
//
// Mirrors what the reference does between loading a file and answering the first query:
//   * leaf<->leaf neighbours        LCAdaptiveGridCell::addAllNeighboursTopDown (LCAdaptiveGridCell.cpp:139-164)
//                                   with the box test of checkIsNeighbor (:910-925, tolerance 1e-7)
//   * border cell<->leaf neighbours LCSampledFunction::addBorderCells (LCSampledFunction.cpp:24-47) through
//                                   LCAdaptiveGridCell::getAdjacentLeaves (:310-333)
//   * influencing-sample sets       LCAdaptiveGridCell::getAllInfluencingSamples (:558-607)
// and then pre-filters each leaf's basis functions to those whose box can reach the leaf (terms that
// are identically zero on the leaf only ever add +0.0 to LCSample::getInterpolationWeight's sum).
#include "loco_flat.h"
#include "loco_cellpoly.h"

#include <algorithm>
#include <unordered_map>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <numeric>

namespace loco {

const char *status_string(int code)
{
	switch (code) {
	case ST_OK: return "no error";
	case ST_OUTSIDE_CELL: return "parameter ouside cell";  // sic, LCAdaptiveGridCell.cpp:450
	case ST_BAD_ARITY: return "invalid number of parameters";
	case ST_NO_COMMON_SAMPLE: return "could not find closest sample";
	case ST_EMPTY_SUPPORT: return "cannot create new from empty list of weighted samples";
	default: return "unknown status";
	}
}

namespace {

struct Ctx {
	const Graph &g;
	Flat &f;
	int N;
	std::vector<int32_t> leaf_of_cell;  // pre-order cell -> leaf rank or -1
	std::vector<int32_t> cell_of_leaf;
	explicit Ctx(const Graph &g_, Flat &f_) : g(g_), f(f_), N(g_.N) {}

	const double *box(const Graph::CellSet &cs, int64_t i) const { return &cs.ranges[(size_t)i * 2 * N]; }

	// LCAdaptiveGridCell::checkIsNeighbor (LCAdaptiveGridCell.cpp:910-925)
	bool touches(const double *a, const double *b) const
	{
		for (int i = 0; i < N; i++) {
			double minA = a[2 * i], maxA = a[2 * i + 1], minB = b[2 * i], maxB = b[2 * i + 1];
			if (((minB - maxA) > kContainingEpsilon) || ((minA - maxB) > kContainingEpsilon)) return false;
		}
		return true;
	}

	// all leaves whose box touches `a` (the fixed point of addAllNeighboursTopDown: a cell only ever
	// inherits neighbours that touch its parent, so the final lists are exactly the touching leaves)
	void touching_leaves(int32_t cell, const double *a, int32_t self_leaf, std::vector<int32_t> *out) const
	{
		if (!touches(a, box(g.tree, cell))) return;
		if (g.tree.split_dir[cell] < 0) {
			if (leaf_of_cell[cell] != self_leaf) out->push_back(leaf_of_cell[cell]);
			return;
		}
		touching_leaves(cell + 1, a, self_leaf, out);
		touching_leaves(g.tree.child2[cell], a, self_leaf, out);
	}

	// LCAdaptiveGridCell::evalPametersAreInCell (:438-454)
	bool in_cell_eps(int32_t cell, const double *x) const
	{
		const double *r = box(g.tree, cell);
		for (int i = 0; i < N; i++)
			if ((x[i] < r[2 * i] - kEpsilon) || (x[i] > r[2 * i + 1] + kEpsilon)) return false;
		return true;
	}

	// LCAdaptiveGridCell::getAdjacentLeaves (:310-333), including its early return on the first leaf
	// that rejects the point (leaves inserted before that stay in the result)
	bool adjacent_leaves(int32_t cell, const double *x, std::vector<int32_t> *out) const
	{
		int32_t dir = g.tree.split_dir[cell];
		if (dir < 0) {
			bool ok = in_cell_eps(cell, x);
			if (ok) out->push_back(leaf_of_cell[cell]);
			return ok;
		}
		double sv = g.tree.split_val[cell];
		if (x[dir] <= sv + kContainingEpsilon)
			if (!adjacent_leaves(cell + 1, x, out)) return false;
		if (x[dir] >= sv - kContainingEpsilon)
			if (!adjacent_leaves(g.tree.child2[cell], x, out)) return false;
		return true;
	}
};

bool is_pow2(double s)
{
	if (!(s > 0) || !std::isfinite(s)) return false;
	int e;
	return std::frexp(s, &e) == 0.5 && std::isnormal(s) && std::isfinite(1.0 / s) && std::isnormal(1.0 / s);
}

} // namespace

namespace {
// phase timer of flatten(): printed when LOCO_FLATTEN_TIMING is set (used to decide what an incremental re-flatten would save)
struct PhaseTimer {
	bool on = getenv("LOCO_FLATTEN_TIMING") != nullptr;
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	void lap(const char *what)
	{
		if (!on) return;
		const auto t1 = std::chrono::steady_clock::now();
		fprintf(stderr, "[flatten] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
		t0 = t1;
	}
};
} // namespace

namespace {
// 128-bit signature of a byte range (two independent 64-bit FNV-1a style streams with different offsets and multipliers)
struct Sig128 {
	uint64_t a = 0xcbf29ce484222325ull, b = 0x84222325cbf29ce4ull;
	void add(const void *p, size_t n)
	{
		const unsigned char *c = static_cast<const unsigned char *>(p);
		size_t i = 0;
		for (; i + 8 <= n; i += 8) {
			uint64_t w;
			memcpy(&w, c + i, 8);
			a = (a ^ w) * 0x100000001b3ull; a ^= a >> 29;
			b = (b ^ (w + 0x9e3779b97f4a7c15ull)) * 0xc2b2ae3d27d4eb4full; b ^= b >> 31;
		}
		for (; i < n; i++) { a = (a ^ c[i]) * 0x100000001b3ull; b = (b ^ c[i]) * 0xc2b2ae3d27d4eb4full; }
	}
};
struct SigKey {
	uint64_t a, b;
	bool operator==(const SigKey &o) const { return a == o.a && b == o.b; }
};
struct SigHash { size_t operator()(const SigKey &k) const { return (size_t)(k.a ^ (k.b * 0x9e3779b97f4a7c15ull)); } };
} // namespace

Status flatten(const Graph &g, Flat *out, const Flat *prev)
{
	PhaseTimer timer;
	LOCO_RETURN_IF_ERROR(validate_graph(g));
	Flat &f = *out;
	f = Flat();
	Ctx cx(g, f);
	const int N = g.N;
	f.N = N; f.D = g.D; f.basis = g.basis;
	f.S = g.n_samples(); f.R = g.n_rows(); f.C = g.tree.size();
	for (int i = 0; i < N; i++) { f.dom_min.push_back(g.domain[2 * i]); f.dom_max.push_back(g.domain[2 * i + 1]); }

	// ---- tree cells, leaf numbering (getAllLeafCells order == pre-order rank of leaves) -----------
	cx.leaf_of_cell.assign((size_t)f.C, -1);
	std::vector<int32_t> level((size_t)f.C, 0), inner_of_cell((size_t)f.C, -1);
	int32_t n_inner = 0;
	for (int64_t c = 0; c < f.C; c++) {
		if (g.tree.split_dir[c] < 0) {
			cx.leaf_of_cell[(size_t)c] = (int32_t)cx.cell_of_leaf.size();
			cx.cell_of_leaf.push_back((int32_t)c);
			f.depth = std::max(f.depth, level[(size_t)c]);
		} else {
			inner_of_cell[(size_t)c] = n_inner++;
			level[(size_t)c + 1] = level[(size_t)c] + 1;
			level[(size_t)g.tree.child2[c]] = level[(size_t)c] + 1;
		}
	}
	if ((uint64_t)n_inner >= kLeafFlag || cx.cell_of_leaf.size() >= kLeafFlag) return Status("tree too large");
	if (N > 16) return Status("more than 16 parameters are not supported");
	f.cells.resize((size_t)n_inner);
	auto code = [&](int64_t c) -> uint32_t {
		return g.tree.split_dir[c] < 0 ? (kLeafFlag | (uint32_t)cx.leaf_of_cell[(size_t)c]) : (uint32_t)inner_of_cell[(size_t)c];
	};
	for (int64_t c = 0; c < f.C; c++) {
		if (g.tree.split_dir[c] < 0) continue;
		TreeCell &tc = f.cells[(size_t)inner_of_cell[(size_t)c]];
		tc.split = g.tree.split_val[c];
		tc.a = ((uint32_t)g.tree.split_dir[c] << 28) | code(c + 1);
		tc.b = code(g.tree.child2[c]);
	}
	f.L = (int64_t)cx.cell_of_leaf.size();
	const int64_t L = f.L;
	f.leaf_lo.resize((size_t)L * N); f.leaf_hi.resize((size_t)L * N);
	f.leaf_lo_eps.resize((size_t)L * N); f.leaf_hi_eps.resize((size_t)L * N);
	f.leaf_center_row.resize((size_t)L);
	f.leaf_status.assign((size_t)L, ST_OK);
	for (int64_t l = 0; l < L; l++) {
		const double *r = cx.box(g.tree, cx.cell_of_leaf[(size_t)l]);
		for (int i = 0; i < N; i++) {
			f.leaf_lo[(size_t)l * N + i] = r[2 * i];
			f.leaf_hi[(size_t)l * N + i] = r[2 * i + 1];
			f.leaf_lo_eps[(size_t)l * N + i] = r[2 * i] - kEpsilon;      // same double op as :447
			f.leaf_hi_eps[(size_t)l * N + i] = r[2 * i + 1] + kEpsilon;  // same double op as :448
		}
		f.leaf_center_row[(size_t)l] = g.tree.center_row[cx.cell_of_leaf[(size_t)l]];
	}

	timer.lap("tree cells + leaf boxes");
	// ---- neighbours ---------------------------------------------------------------------------------
	const bool cubic = g.basis == CUBIC_BSPLINE;
	std::vector<std::vector<int32_t>> leaf_nb((size_t)L), border_nb((size_t)L);
	// the neighbour relation exists for both basis types (split()/addAllNeighboursTopDown maintain it
	// regardless); only the cubic influencing set consumes it
	for (int64_t l = 0; l < L; l++) {
		cx.touching_leaves(0, cx.box(g.tree, cx.cell_of_leaf[(size_t)l]), (int32_t)l, &leaf_nb[(size_t)l]);
		std::sort(leaf_nb[(size_t)l].begin(), leaf_nb[(size_t)l].end());
	}
	{
		// addBorderCells: every leaf adjacent to any sample centre of the border cell
		std::vector<int32_t> stamp((size_t)L, -1), found;
		for (int64_t b = 0; b < g.border.size(); b++) {
			found.clear();
			for (int64_t e = g.border.ls_off[b]; e < g.border.ls_off[b + 1]; e++)
				cx.adjacent_leaves(0, &g.sample_center[(size_t)g.border.ls_sample[e] * N], &found);
			for (int32_t l : found) {
				if (stamp[(size_t)l] == (int32_t)b) continue;
				stamp[(size_t)l] = (int32_t)b;
				border_nb[(size_t)l].push_back((int32_t)b);
			}
		}
	}
	f.n_leaf_neighbors.resize((size_t)L); f.n_border_neighbors.resize((size_t)L);
	for (int64_t l = 0; l < L; l++) {
		f.n_leaf_neighbors[(size_t)l] = (int32_t)leaf_nb[(size_t)l].size();
		f.n_border_neighbors[(size_t)l] = (int32_t)border_nb[(size_t)l].size();
	}

	timer.lap("neighbours");
	// ---- influencing-sample sets (getAllInfluencingSamples) -----------------------------------------------
	f.sup_off.assign(1, 0);
	std::vector<int64_t> seen((size_t)f.S, -1), own((size_t)f.S, -1);
	std::vector<std::pair<int32_t, int32_t>> sup;  // (sample, row)
	for (int64_t l = 0; l < L; l++) {
		sup.clear();
		int32_t cell = cx.cell_of_leaf[(size_t)l];
		for (int64_t e = g.tree.ls_off[cell]; e < g.tree.ls_off[cell + 1]; e++) {
			int32_t s = g.tree.ls_sample[e];
			if (seen[(size_t)s] == l) continue;
			seen[(size_t)s] = l; own[(size_t)s] = l;
			sup.push_back({s, g.tree.ls_row[e]});
		}
		if (cubic) {
			auto take = [&](const Graph::CellSet &cs, int32_t nb) {
				bool checked = false;
				for (int64_t e = cs.ls_off[nb]; e < cs.ls_off[nb + 1]; e++) {
					int32_t s = cs.ls_sample[e];
					if (seen[(size_t)s] == l) continue;
					if (!checked) {
						// getNeightbourShapeInfoMappingMapping (:507-556) needs a sample common to both cells
						bool common = false;
						for (int64_t e2 = cs.ls_off[nb]; e2 < cs.ls_off[nb + 1] && !common; e2++) common = own[(size_t)cs.ls_sample[e2]] == l;
						if (!common) f.leaf_status[(size_t)l] = ST_NO_COMMON_SAMPLE;
						checked = true;
					}
					seen[(size_t)s] = l;
					sup.push_back({s, cs.ls_row[e]});
				}
			};
			for (int32_t nb : leaf_nb[(size_t)l]) take(g.tree, cx.cell_of_leaf[(size_t)nb]);
			for (int32_t nb : border_nb[(size_t)l]) take(g.border, nb);
		}
		if (sup.empty() && f.leaf_status[(size_t)l] == ST_OK) f.leaf_status[(size_t)l] = ST_EMPTY_SUPPORT;
		std::sort(sup.begin(), sup.end());
		for (auto &p : sup) { f.sup_sample.push_back(p.first); f.sup_row.push_back(p.second); }
		f.sup_off.push_back((int64_t)f.sup_sample.size());
		f.max_support = std::max<int32_t>(f.max_support, (int32_t)sup.size());
	}

	timer.lap("influencing-sample sets");
	// ---- terms ------------------------------------------------------------------------------------------
	f.exact_rcp = true;
	for (double s : g.bf_support) if (!is_pow2(s)) { f.exact_rcp = false; break; }
	f.term_off.assign(1, 0);
	f.dterm_off.assign(1, 0);
	for (int64_t l = 0; l < L; l++) {
		const double *lo = &f.leaf_lo_eps[(size_t)l * N], *hi = &f.leaf_hi_eps[(size_t)l * N];
		for (int64_t k = f.sup_off[(size_t)l]; k < f.sup_off[(size_t)l + 1]; k++) {
			int32_t s = f.sup_sample[(size_t)k];
			for (int64_t b = g.bf_off[s]; b < g.bf_off[s + 1]; b++) {
				const double *c = &g.bf_center[(size_t)b * N], *sp = &g.bf_support[(size_t)b * N];
				bool reach = true;
				for (int i = 0; i < N && reach; i++) {
					// keep unless the basis box is provably disjoint from every accepted point of the leaf
					// ([lo-EPS, hi+EPS]); the margin only ever keeps extra (zero) terms
					double m = 1e-9 + 1e-12 * (std::fabs(c[i]) + std::fabs(sp[i]) + std::fabs(lo[i]) + std::fabs(hi[i]));
					if ((c[i] + sp[i] < lo[i] - m) || (c[i] - sp[i] > hi[i] + m)) reach = false;
				}
				if (!reach) {
					// The reference's LINEAR evalDeriv tests the SIGNED d = (x-c)/s against 1 (LCBasisFunction.cpp:295-332),
					// so a hat lying entirely to the right of the leaf in some dimension (d < -1 there) still
					// contributes to a derivative.  Such terms are kept in a derivative-only list; a hat entirely
					// to the LEFT in any dimension has d > 1 there and is zero in the reference too.
					if (!cubic) {
						bool left = false;
						for (int i = 0; i < N && !left; i++) {
							double m = 1e-9 + 1e-12 * (std::fabs(c[i]) + std::fabs(sp[i]) + std::fabs(lo[i]) + std::fabs(hi[i]));
							left = c[i] + sp[i] < lo[i] - m;
						}
						if (!left) {
							for (int i = 0; i < N; i++) {
								f.dterm_cs.push_back(c[i]);
								f.dterm_cs.push_back(f.exact_rcp ? 1.0 / sp[i] : sp[i]);
							}
							f.dterm_w.push_back(g.bf_weight[b]);
							f.dterm_row.push_back(f.sup_row[(size_t)k]);
						}
					}
					continue;
				}
				for (int i = 0; i < N; i++) {
					f.term_cs.push_back(c[i]);
					f.term_cs.push_back(f.exact_rcp ? 1.0 / sp[i] : sp[i]);
				}
				f.term_w.push_back(g.bf_weight[b]);
				f.term_slot.push_back((int32_t)(k - f.sup_off[(size_t)l]));
				f.term_row.push_back(f.sup_row[(size_t)k]);
			}
		}
		f.term_off.push_back((int64_t)f.term_w.size());
		f.dterm_off.push_back((int64_t)f.dterm_w.size());
		f.max_terms = std::max<int32_t>(f.max_terms, (int32_t)(f.term_off[(size_t)l + 1] - f.term_off[(size_t)l]));
	}

	timer.lap("term lists");
	// ---- merged terms for scalar functions (see loco_flat.h) ---------------------------------------------------------
	if (g.D == 1) {
		std::vector<int64_t> order;
		const size_t row_bytes = sizeof(double) * 2 * (size_t)N;
		f.mterm_off.assign(1, 0);
		f.mrun_off.assign(1, 0);
		for (int64_t l = 0; l < L; l++) {
			const int64_t t0 = f.term_off[(size_t)l], t1 = f.term_off[(size_t)l + 1];
			order.resize((size_t)(t1 - t0));
			std::iota(order.begin(), order.end(), t0);
			std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
				const int c = memcmp(&f.term_cs[(size_t)a * 2 * N], &f.term_cs[(size_t)b * 2 * N], row_bytes);
				return c < 0 || (c == 0 && a < b);
			});
			for (size_t i = 0; i < order.size(); i++) {
				const bool same = i > 0 && memcmp(&f.term_cs[(size_t)order[i] * 2 * N], &f.term_cs[(size_t)order[i - 1] * 2 * N], row_bytes) == 0;
				if (!same) {
					if (i > 0) f.mrun_off.push_back((int64_t)f.mrun_term.size());
					f.mterm_cs.insert(f.mterm_cs.end(), f.term_cs.begin() + order[i] * 2 * N, f.term_cs.begin() + (order[i] + 1) * 2 * N);
				}
				f.mrun_term.push_back((int32_t)order[i]);
			}
			if (!order.empty()) f.mrun_off.push_back((int64_t)f.mrun_term.size());
			f.mterm_off.push_back((int64_t)f.mterm_cs.size() / (2 * N));
		}
		const int64_t M = f.mterm_off.back(), T = f.term_off.back();
		if (T >= (int64_t)0x7fffffff || M * 4 > T * 3) {  // not worth a second copy of the term list
			f.mterm_off.clear(); f.mterm_cs.clear(); f.mrun_off.clear(); f.mrun_term.clear();
		}
		timer.lap("merged terms");
		// ---- evaluation cells (see loco_flat.h): only where a leaf carries many merged terms
		if (!f.mterm_off.empty()) {
			// signatures of the leaves (incremental re-flatten, see loco_flat.h) and where equal leaves of `prev` are
			f.leaf_sig.resize((size_t)2 * L);
			for (int64_t l = 0; l < L; l++) {
				Sig128 h;
				const int32_t hdr[3] = {N, g.basis, f.exact_rcp ? 1 : 0};
				h.add(hdr, sizeof hdr);
				h.add(&g.tree.ranges[0], sizeof(double) * 2 * (size_t)N);
				h.add(&f.leaf_lo[(size_t)l * N], sizeof(double) * (size_t)N);
				h.add(&f.leaf_hi[(size_t)l * N], sizeof(double) * (size_t)N);
				const int64_t m0 = f.mterm_off[(size_t)l], m1 = f.mterm_off[(size_t)l + 1];
				h.add(&f.mterm_cs[(size_t)m0 * 2 * N], sizeof(double) * 2 * (size_t)N * (size_t)(m1 - m0));
				f.leaf_sig[(size_t)2 * l] = h.a; f.leaf_sig[(size_t)2 * l + 1] = h.b;
			}
			std::unordered_map<SigKey, int64_t, SigHash> prev_leaf;
			const bool can_reuse = prev && prev->ecell_scale > 0.0 && !prev->ecell_off.empty() && prev->N == N && prev->basis == g.basis &&
			                       (int64_t)prev->leaf_sig.size() == 2 * prev->L && (int64_t)prev->eterm_src.size() == prev->eterm_off.back() &&
			                       (int64_t)prev->eterm_cs.size() == prev->eterm_off.back() * 2 * N;
			if (can_reuse)
				for (int64_t pl = 0; pl < prev->L; pl++) prev_leaf[SigKey{prev->leaf_sig[(size_t)2 * pl], prev->leaf_sig[(size_t)2 * pl + 1]}] = pl;
			std::vector<int64_t> reuse_from((size_t)L, -1);  // leaf of `prev` whose evaluation cells this leaf copies
			if (can_reuse)
				for (int64_t l = 0; l < L; l++) {
					auto it = prev_leaf.find(SigKey{f.leaf_sig[(size_t)2 * l], f.leaf_sig[(size_t)2 * l + 1]});
					if (it != prev_leaf.end() &&
					    prev->mterm_off[(size_t)it->second + 1] - prev->mterm_off[(size_t)it->second] == f.mterm_off[(size_t)l + 1] - f.mterm_off[(size_t)l])
						reuse_from[(size_t)l] = it->second;
				}
			timer.lap("  signatures");
			// cells of about a quarter of the narrowest support, up to 32 per dimension and 1,024 per leaf; coarser if the
			// lists would outgrow 32x the merged list (or 1 GB)
			const int kMinTerms = 24, kMaxBitsPerDim = 5, kMaxBits = 10;
			const int64_t kMaxEterms = std::min<int64_t>(((int64_t)1 << 30) / ((2 * N + 2) * 8), 32 * f.mterm_off.back());
			bool any = false, ok = false;
			double used_scale = 0.0;
			// an update starts at the scale of the previous state (a refined grid does not get by with finer cells than its
			// predecessor needed; starting below would recompute every leaf once for nothing)
			const double first_scale = can_reuse ? prev->ecell_scale : 0.25;
			for (double kCellScale = first_scale; kCellScale <= 2.0 && !ok; kCellScale *= 2.0) {
			ok = true; any = false;
			used_scale = kCellScale;
			f.eterm_cs.clear(); f.eterm_src.clear();
			if (can_reuse) {  // about the size of the previous lists: no regrowth while copying
				f.eterm_cs.reserve(prev->eterm_cs.size() + prev->eterm_cs.size() / 8);
				f.eterm_src.reserve(prev->eterm_src.size() + prev->eterm_src.size() / 8);
				f.eterm_off.reserve(prev->eterm_off.size() + prev->eterm_off.size() / 8);
			}
			f.ecell_off.assign(1, 0);
			f.ecell_bits.assign((size_t)L * N, 0);
			f.eterm_off.assign(1, 0);
			std::vector<double> smin((size_t)N);
			f.reused_leaves = 0;
			for (int64_t l = 0; l < L && ok; l++) {
				const int64_t m0 = f.mterm_off[(size_t)l], m1 = f.mterm_off[(size_t)l + 1];
				const double *lo = &f.leaf_lo[(size_t)l * N], *hi = &f.leaf_hi[(size_t)l * N];
				uint8_t *bits = &f.ecell_bits[(size_t)l * N];
				if (reuse_from[(size_t)l] >= 0 && prev->ecell_scale == kCellScale) {
					// same box, same merged terms, same scale: the cells and their term lists are those of the previous model
					const int64_t pl = reuse_from[(size_t)l];
					const int64_t pc0 = prev->ecell_off[(size_t)pl], pc1 = prev->ecell_off[(size_t)pl + 1];
					const int64_t shift = m0 - prev->mterm_off[(size_t)pl];
					int total = 0;
					for (int i = 0; i < N; i++) { bits[i] = prev->ecell_bits[(size_t)pl * N + i]; total += bits[i]; }
					any = any || total > 0;
					const int64_t e0 = prev->eterm_off[(size_t)pc0], e1 = prev->eterm_off[(size_t)pc1];
					const int64_t base = (int64_t)f.eterm_src.size();
					f.eterm_cs.insert(f.eterm_cs.end(), prev->eterm_cs.begin() + e0 * 2 * N, prev->eterm_cs.begin() + e1 * 2 * N);
					{
						const size_t at = f.eterm_src.size();
						f.eterm_src.resize(at + (size_t)(e1 - e0));
						const int32_t *ps = &prev->eterm_src[(size_t)e0];
						int32_t *ns = &f.eterm_src[at];
						for (int64_t e = 0; e < e1 - e0; e++) ns[e] = (int32_t)(ps[e] + shift);
					}
					for (int64_t c = pc0; c < pc1; c++) f.eterm_off.push_back(base + prev->eterm_off[(size_t)c + 1] - e0);
					f.ecell_off.push_back(f.ecell_off.back() + (pc1 - pc0));
					f.reused_leaves++;
					if ((int64_t)f.eterm_src.size() > kMaxEterms) { ok = false; break; }
					continue;
				}
				if (m1 - m0 >= kMinTerms) {
					for (int i = 0; i < N; i++) smin[(size_t)i] = 1e300;
					for (int64_t t = m0; t < m1; t++)
						for (int i = 0; i < N; i++) {
							const double v = f.mterm_cs[(size_t)t * 2 * N + 2 * i + 1];
							smin[(size_t)i] = std::min(smin[(size_t)i], f.exact_rcp ? 1.0 / v : v);
						}
					int total = 0;
					for (int i = 0; i < N; i++) {
						const double cells = (hi[i] - lo[i]) / (kCellScale * smin[(size_t)i]);  // cell ~ full width of the narrowest function
						int b = 0;
						while (b < kMaxBitsPerDim && (double)(2 << b) <= cells) b++;
						bits[i] = (uint8_t)b;
						total += b;
					}
					while (total > kMaxBits) {  // trim the finest dimension first
						int best = 0;
						for (int i = 1; i < N; i++) if (bits[i] > bits[best]) best = i;
						bits[best]--; total--;
					}
					any = any || total > 0;
				}
				int64_t n_cells = 1;
				for (int i = 0; i < N; i++) n_cells <<= bits[i];
				std::vector<double> clo((size_t)N), chi((size_t)N);
				for (int64_t c = 0; c < n_cells; c++) {
					int64_t r = c;
					for (int i = 0; i < N; i++) {
						const int64_t g = (int64_t)1 << bits[i], k = r % g;
						r /= g;
						const double h = (hi[i] - lo[i]) / (double)g;
						// margin: the acceptance band (1e-6), the rounding of the cell index, and a relative guard
						const double mg = 2.5e-6 + 1e-9 * (std::fabs(lo[i]) + std::fabs(hi[i])) + 1e-6 * h;
						clo[(size_t)i] = lo[i] + h * (double)k - mg;
						chi[(size_t)i] = lo[i] + h * (double)(k + 1) + mg;
					}
					for (int64_t t = m0; t < m1; t++) {
						bool reach = true;
						for (int i = 0; i < N && reach; i++) {
							const double cc = f.mterm_cs[(size_t)t * 2 * N + 2 * i], v = f.mterm_cs[(size_t)t * 2 * N + 2 * i + 1];
							const double sp = f.exact_rcp ? 1.0 / v : v;
							if (cc + sp < clo[(size_t)i] || cc - sp > chi[(size_t)i]) reach = false;
						}
						if (!reach) continue;
						f.eterm_cs.insert(f.eterm_cs.end(), f.mterm_cs.begin() + t * 2 * N, f.mterm_cs.begin() + (t + 1) * 2 * N);
						f.eterm_src.push_back((int32_t)t);
					}
					f.eterm_off.push_back((int64_t)f.eterm_src.size());
					if ((int64_t)f.eterm_src.size() > kMaxEterms) { ok = false; break; }
				}
				f.ecell_off.push_back(f.ecell_off.back() + n_cells);
			}
			}
			if (!any || !ok || f.ecell_off.back() >= (int64_t)0x7fffffff) {
				f.ecell_off.clear(); f.ecell_bits.clear(); f.eterm_off.clear(); f.eterm_cs.clear(); f.eterm_src.clear();
				f.reused_leaves = 0;
			} else {
				f.ecell_scale = used_scale;
				timer.lap("  evaluation cells");
				// ---- polynomial cells (loco_cellpoly.h)
				const int F = cubic ? 4 : 2;
				int64_t P = 1;
				for (int i = 0; i < N; i++) P *= F;
				if (N <= 8 && P <= 64) {
					f.epoly_idx.assign((size_t)f.ecell_off.back(), -1);
					const double *rb = &g.tree.ranges[0];  // the root box
					std::vector<double> clo((size_t)N), chi((size_t)N);
					for (int64_t l = 0; l < L; l++) {
						const double *lo = &f.leaf_lo[(size_t)l * N], *hi = &f.leaf_hi[(size_t)l * N];
						const uint8_t *bits = &f.ecell_bits[(size_t)l * N];
						const int64_t c0 = f.ecell_off[(size_t)l], c1 = f.ecell_off[(size_t)l + 1];
						if (c1 - c0 < 2) continue;  // undivided leaves keep their term list
						if (reuse_from[(size_t)l] >= 0 && prev->ecell_scale == used_scale) {
							// reused cells: which of them are polynomial cells, and their geometry, are those of the previous model
							const int64_t pc0 = prev->ecell_off[(size_t)reuse_from[(size_t)l]];
							for (int64_t c = c0; c < c1; c++) {
								const int32_t pi = prev->epoly_idx.empty() ? -1 : prev->epoly_idx[(size_t)(pc0 + (c - c0))];
								if (pi < 0) continue;
								f.epoly_idx[(size_t)c] = (int32_t)f.epoly_cell.size();
								f.epoly_cell.push_back((int32_t)c);
								f.epoly_geom.insert(f.epoly_geom.end(), prev->epoly_geom.begin() + (size_t)pi * 2 * N, prev->epoly_geom.begin() + (size_t)(pi + 1) * 2 * N);
							}
							continue;
						}
						for (int64_t c = c0; c < c1; c++) {
							int64_t r = c - c0;
							bool good = true;
							for (int i = 0; i < N; i++) {
								const int64_t gdim = (int64_t)1 << bits[i], k = r % gdim;
								r /= gdim;
								const double h = (hi[i] - lo[i]) / (double)gdim;
								clo[(size_t)i] = lo[i] + h * (double)k;
								chi[(size_t)i] = lo[i] + h * (double)(k + 1);
								if ((k == 0 && lo[i] <= rb[2 * i]) || (k == gdim - 1 && hi[i] >= rb[2 * i + 1])) good = false;  // touches the root box
							}
							for (int64_t t = f.eterm_off[(size_t)c]; good && t < f.eterm_off[(size_t)c + 1]; t++)
								for (int i = 0; i < N && good; i++) {
									const double cc = f.eterm_cs[(size_t)t * 2 * N + 2 * i], v = f.eterm_cs[(size_t)t * 2 * N + 2 * i + 1];
									const double sp = f.exact_rcp ? 1.0 / v : v;
									if (cubic ? knot_inside<true>(cc, sp, clo[(size_t)i], chi[(size_t)i]) : knot_inside<false>(cc, sp, clo[(size_t)i], chi[(size_t)i]))
										good = false;
								}
							if (!good) continue;
							f.epoly_idx[(size_t)c] = (int32_t)f.epoly_cell.size();
							f.epoly_cell.push_back((int32_t)c);
							for (int i = 0; i < N; i++) {
								f.epoly_geom.push_back(0.5 * (clo[(size_t)i] + chi[(size_t)i]));
								f.epoly_geom.push_back(2.0 / (chi[(size_t)i] - clo[(size_t)i]));
							}
						}
					}
					if (f.epoly_cell.empty()) f.epoly_idx.clear();
					timer.lap("  polynomial cells");
				}
			}
		}
	}

	timer.lap("merged terms + evaluation cells");
	// ---- tensor-product structure (see loco_flat.h) ---------------------------------------------------
	{
		const int F = cubic ? 4 : 2;
		int64_t KT = 1;
		for (int i = 0; i < N && KT <= (1 << 20); i++) KT *= F;
		bool all = L > 0 && KT <= (1 << 16);
		std::vector<double> cen((size_t)N * F);
		std::vector<int32_t> hit;
		for (int64_t l = 0; l < L && all; l++) {
			const int64_t t0 = f.term_off[(size_t)l], t1 = f.term_off[(size_t)l + 1];
			if (t1 - t0 != KT || f.sup_off[(size_t)l + 1] - f.sup_off[(size_t)l] != KT) { all = false; break; }
			// distinct centres per dimension, common support per dimension, unit weights
			std::vector<int> nc((size_t)N, 0);
			for (int64_t t = t0; t < t1 && all; t++) {
				if (f.term_w[(size_t)t] != 1.0) all = false;
				for (int i = 0; i < N && all; i++) {
					const double c = f.term_cs[(size_t)t * 2 * N + 2 * i], r = f.term_cs[(size_t)t * 2 * N + 2 * i + 1];
					if (r != f.term_cs[(size_t)t0 * 2 * N + 2 * i + 1]) { all = false; break; }
					bool found = false;
					for (int j = 0; j < nc[(size_t)i]; j++) found |= cen[(size_t)i * F + j] == c;
					if (!found) {
						if (nc[(size_t)i] == F) { all = false; break; }
						cen[(size_t)i * F + nc[(size_t)i]++] = c;
					}
				}
			}
			for (int i = 0; i < N && all; i++) {
				if (nc[(size_t)i] != F) all = false;
				std::sort(cen.begin() + (size_t)i * F, cen.begin() + (size_t)(i + 1) * F);
			}
			if (!all) break;
			hit.assign((size_t)KT, -1);
			for (int64_t t = t0; t < t1 && all; t++) {
				int64_t idx = 0, mul = 1;
				for (int i = 0; i < N; i++) {
					const double c = f.term_cs[(size_t)t * 2 * N + 2 * i];
					int j = 0;
					while (cen[(size_t)i * F + j] != c) j++;
					idx += j * mul;
					mul *= F;
				}
				if (hit[(size_t)idx] >= 0) all = false;
				hit[(size_t)idx] = (int32_t)(t - t0);
			}
			if (!all) break;
			f.tl_center.insert(f.tl_center.end(), cen.begin(), cen.end());
			for (int i = 0; i < N; i++) f.tl_rcp.push_back(f.term_cs[(size_t)t0 * 2 * N + 2 * i + 1]);
			for (int64_t j = 0; j < KT; j++) {
				f.tl_slot.push_back(f.term_slot[(size_t)(t0 + hit[(size_t)j])]);
				f.tl_row.push_back(f.term_row[(size_t)(t0 + hit[(size_t)j])]);
			}
		}
		if (all) {
			f.tensor_F = F; f.tensor_K = (int32_t)KT;
			// uniform knots: c[j+1] - c[j] == support / 2 exactly, in every leaf and dimension (cubic only)
			bool uni = cubic;
			for (int64_t l = 0; l < L && uni; l++)
				for (int i = 0; i < N && uni; i++) {
					const double *c = &f.tl_center[((size_t)l * N + i) * F];
					const double r = f.tl_rcp[(size_t)l * N + i], sup = f.exact_rcp ? 1.0 / r : r;
					for (int j = 0; j + 1 < F; j++) uni = uni && (c[j + 1] - c[j] == 0.5 * sup);
					// the four segment polynomials are evaluated at t = (x - c[1]) / h, valid on [c[1], c[2]] only:
					// the leaf box must BE that knot interval (true for bootstrap grids, not for an arbitrary graph)
					uni = uni && f.leaf_lo[(size_t)l * N + i] == c[1] && f.leaf_hi[(size_t)l * N + i] == c[2];
				}
			f.tensor_uniform = uni;
			// single polynomial piece per leaf: no knot of any of the leaf's one-dimensional basis functions strictly inside the leaf box
			bool poly = true;
			for (int64_t l = 0; l < L && poly; l++)
				for (int i = 0; i < N && poly; i++) {
					const double *c = &f.tl_center[((size_t)l * N + i) * F];
					const double r = f.tl_rcp[(size_t)l * N + i], sup = f.exact_rcp ? 1.0 / r : r;
					const double lo = f.leaf_lo[(size_t)l * N + i], hi = f.leaf_hi[(size_t)l * N + i];
					if (!(hi > lo)) poly = false;
					for (int j = 0; j < F && poly; j++) poly = cubic ? !knot_inside<true>(c[j], sup, lo, hi) : !knot_inside<false>(c[j], sup, lo, hi);
				}
			f.tensor_poly = poly;
		}
		else { f.tl_center.clear(); f.tl_rcp.clear(); f.tl_slot.clear(); f.tl_row.clear(); }
		// ---- fold digits whose exchange never changes the value row (see loco_flat.h)
		if (all) {
			f.fold_off.assign(1, 0);
			f.fold_map.resize((size_t)L * N * F);
			f.fold_n.resize((size_t)L * N);
			std::vector<int64_t> stride((size_t)N);
			for (int i = 0; i < N; i++) stride[(size_t)i] = i ? stride[(size_t)i - 1] * F : 1;
			for (int64_t l = 0; l < L; l++) {
				const int32_t *rows = &f.tl_row[(size_t)l * KT];
				uint8_t *map = &f.fold_map[(size_t)l * N * F];
				int rep[16][4];  // representative digit of each class
				for (int d = 0; d < N; d++) {
					int n = 0;
					for (int j = 0; j < F; j++) {
						int cls = -1;
						for (int c = 0; c < n && cls < 0; c++) {
							// j ~ rep[d][c]  iff  rows agree for every setting of the other digits
							const int64_t dj = (int64_t)(j - rep[d][c]) * stride[(size_t)d];
							bool same = true;
							for (int64_t k = 0; k < KT && same; k++)
								if ((k / stride[(size_t)d]) % F == rep[d][c]) same = rows[k] == rows[k + dj];
							if (same) cls = c;
						}
						if (cls < 0) { cls = n; rep[d][n++] = j; }
						map[d * F + j] = (uint8_t)cls;
					}
					f.fold_n[(size_t)l * N + d] = (uint8_t)n;
				}
				// rows of the folded tensor (representatives), class index of dimension 0 fastest
				int64_t KF = 1;
				for (int d = 0; d < N; d++) KF *= f.fold_n[(size_t)l * N + d];
				for (int64_t kf = 0; kf < KF; kf++) {
					int64_t r = kf, k = 0;
					for (int d = 0; d < N; d++) {
						const int n = f.fold_n[(size_t)l * N + d];
						k += rep[d][r % n] * stride[(size_t)d];
						r /= n;
					}
					f.fold_row.push_back(rows[k]);
				}
				f.fold_off.push_back((int64_t)f.fold_row.size());
				f.max_fold_K = std::max<int32_t>(f.max_fold_K, (int32_t)KF);
			}
		}
	}

	timer.lap("tensor structure + folding");
	// ---- are the leaf boxes exactly the boxes implied by the split planes? -----------------------------------
	{
		bool ok = true;
		std::vector<double> lo((size_t)f.C * N), hi((size_t)f.C * N);
		for (int i = 0; i < N; i++) { lo[i] = g.tree.ranges[2 * i]; hi[i] = g.tree.ranges[2 * i + 1]; }
		for (int64_t c = 0; c < f.C && ok; c++) {
			const int32_t dir = g.tree.split_dir[c];
			if (dir < 0) {
				const double *r = cx.box(g.tree, c);
				for (int i = 0; i < N; i++) ok &= (r[2 * i] == lo[(size_t)c * N + i]) && (r[2 * i + 1] == hi[(size_t)c * N + i]);
				continue;
			}
			const int64_t c1 = c + 1, c2 = g.tree.child2[c];
			const double sv = g.tree.split_val[c];
			if (!(sv >= lo[(size_t)c * N + dir] && sv <= hi[(size_t)c * N + dir])) ok = false;
			for (int i = 0; i < N; i++) {
				lo[(size_t)c1 * N + i] = lo[(size_t)c2 * N + i] = lo[(size_t)c * N + i];
				hi[(size_t)c1 * N + i] = hi[(size_t)c2 * N + i] = hi[(size_t)c * N + i];
			}
			hi[(size_t)c1 * N + dir] = sv;
			lo[(size_t)c2 * N + dir] = sv;
		}
		f.path_consistent = ok;
	}
	for (uint8_t s : f.leaf_status) f.any_leaf_status |= s != ST_OK;

	timer.lap("path consistency");
	// ---- locate grid (see loco_flat.h) -------------------------------------------------------------------------
	if (n_inner > 0) {
		const int kMaxLutBits = 17;
		// one bit per tree level, given to the dimension most cells of that level split
		std::vector<std::vector<int64_t>> by_level;
		for (int64_t c = 0; c < f.C; c++) {
			if (g.tree.split_dir[c] < 0) continue;
			if ((size_t)level[(size_t)c] >= by_level.size()) by_level.resize((size_t)level[(size_t)c] + 1, std::vector<int64_t>((size_t)N, 0));
			by_level[(size_t)level[(size_t)c]][(size_t)g.tree.split_dir[c]]++;
		}
		int total_bits = 0;
		for (size_t lv = 0; lv < by_level.size() && total_bits < kMaxLutBits; lv++) {
			int best = (int)(std::max_element(by_level[lv].begin(), by_level[lv].end()) - by_level[lv].begin());
			f.lut_bits[best]++;
			total_bits++;
		}
		std::vector<int64_t> stride((size_t)N);
		int64_t entries = 1;
		for (int d = 0; d < N; d++) {
			stride[(size_t)d] = entries;
			entries <<= f.lut_bits[d];
			f.lut_boff[d] = (int32_t)f.lut_bounds.size();
			const int64_t G = (int64_t)1 << f.lut_bits[d];
			const double lo = g.tree.ranges[2 * d], hi = g.tree.ranges[2 * d + 1];
			for (int64_t i = 0; i <= G; i++) f.lut_bounds.push_back(i == G ? hi : lo + (hi - lo) * ((double)i / (double)G));
		}
		f.lut.assign((size_t)entries, 0u);
		std::vector<int64_t> a((size_t)N, 0), b((size_t)N);
		for (int d = 0; d < N; d++) b[(size_t)d] = (int64_t)1 << f.lut_bits[d];
		// fill the (non-empty) index box [a, b) with `code`
		auto fill = [&](const std::vector<int64_t> &lo_i, const std::vector<int64_t> &hi_i, uint32_t cd) {
			std::vector<int64_t> it(lo_i);
			for (;;) {
				int64_t idx = 0;
				for (int d = 0; d < N; d++) idx += it[(size_t)d] * stride[(size_t)d];
				// dimension 0 is contiguous
				for (int64_t i0 = lo_i[0]; i0 < hi_i[0]; i0++) f.lut[(size_t)(idx - it[0] + i0)] = cd;
				int d = 1;
				for (; d < N; d++) {
					if (++it[(size_t)d] < hi_i[(size_t)d]) break;
					it[(size_t)d] = lo_i[(size_t)d];
				}
				if (d >= N) break;
			}
		};
		// recursive split of the index box along the tree (explicit stack: trees may be deep)
		struct Item { int64_t cell; std::vector<int64_t> a, b; };
		std::vector<Item> stack;
		stack.push_back({0, a, b});
		while (!stack.empty()) {
			Item it = std::move(stack.back());
			stack.pop_back();
			bool empty = false;
			for (int d = 0; d < N; d++) empty |= it.a[(size_t)d] >= it.b[(size_t)d];
			if (empty) continue;
			const int32_t dir = g.tree.split_dir[it.cell];
			if (dir < 0) { fill(it.a, it.b, code(it.cell)); continue; }
			const double sv = g.tree.split_val[it.cell];
			const double *B = &f.lut_bounds[(size_t)f.lut_boff[dir]];
			const int64_t G = (int64_t)1 << f.lut_bits[dir];
			// grid cell i covers [B[i], B[i+1]) with B[0] = -inf and B[G] = +inf.
			// child1 (x < sv) contains it iff upper(i) <= sv; child2 (x >= sv, or NaN) iff lower(i) >= sv.
			int64_t j1 = it.a[(size_t)dir];
			while (j1 < it.b[(size_t)dir] && j1 + 1 < G && B[j1 + 1] <= sv) j1++;   // cells [a, j1) lie in child1
			int64_t j2 = j1;
			while (j2 < it.b[(size_t)dir] && !(j2 > 0 && B[j2] >= sv)) j2++;          // cells [j2, b) lie in child2
			if (j2 > j1) {  // straddling cells: the descent resumes at this tree cell
				Item mid{it.cell, it.a, it.b};
				mid.a[(size_t)dir] = j1; mid.b[(size_t)dir] = j2;
				fill(mid.a, mid.b, code(it.cell));
			}
			Item left{it.cell + 1, it.a, it.b}, right{g.tree.child2[it.cell], it.a, it.b};
			left.b[(size_t)dir] = j1;
			right.a[(size_t)dir] = j2;
			stack.push_back(std::move(left));
			stack.push_back(std::move(right));
		}
	}

	f.sample_row = g.sample_row;
	timer.lap("locate grid");
	f.values = g.values;
	return Status();
}

} // namespace loco
