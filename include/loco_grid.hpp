// loco_grid.hpp -- the HOST side of the refinement loop: the builder's grid state and the topology
// surgery of LCAdaptiveGrid::refineCell, so that the loop "batched decideSplit on the GPU -> split the flagged cells ->
// loco_model_update_from_graph" runs without the reference library.  Header-only C++ above the C ABI (loco_b200.h); it
// never evaluates the interpolant (there is no CPU evaluation path in this build: the error check of the loop is
// loco_center_error / loco_error_check* on the device).
//
//   reference (LCAdaptiveSampling/)                                     here
//   LCAdaptiveGrid::refineCell               LCAdaptiveGrid.cpp:629-914    Grid::refineCell
//   LCAdaptiveGrid::getReplacingSamples      :1002-1122                    Grid::replacingSamples
//   LCAdaptiveGrid::getAdjacentCells / getDoubleAdjacentCells :141-203     Grid::adjacentCells / doubleAdjacentCells
//   LCAdaptiveGrid::setLeafInformation       :212-231                      Grid::setLeafInformation
//   LCAdaptiveGridCell::split, checkIsNeighbor, replaceNeighborWithChildren LCAdaptiveGridCell.cpp:107-137, 910-944
//   LCAdaptiveGridCell::getAdjacentLeaves    :310-333                      Grid::adjacentLeaves
//   LCAdaptiveGridCell::getMidValues, getSample, getDoubleAdjacentSamples :715-793
//   LCSample::refineBasisFunctions, addBasisFunction, extractGroupBasisFunctions LCSample.cpp:106-226
//   LC{Linear,Cubic}BSpline::refine, LCBasisFunction::refineAllDirections, checkRegionOverlap LCBasisFunction.cpp:64-86, 209-228, 334-360, 485-509
//   LCBasisFunctionKey equality (centre and support within EPSILON)     LCBasisFunction.h:76-104
//
// Representation: indices instead of pointers (cells, samples, value rows); a cell keeps its neighbour LIST and its
// sample map as the reference does.  Where the reference iterates an unordered container (its basis-function maps hash
// every key to 0, LCBasisFunction.h:107-121), the order here is a fixed one (insertion / index order): every such loop of
// refineCell is order-independent except for the numbering of newly created samples and the order of floating-point
// sums of dyadic weights, so grids agree with the reference's as SETS (samples by centre, basis functions by key) with
// weights to rounding; tests/test_native_refine.py checks exactly that against the compiled reference.
//
// Neighbour relations of a grid taken over from a plain-array graph (Grid::fromGraph) are those of the LIVE builder, not
// of the proto loader: border cells are linked to the leaves and to the border cells they touch (bootstrapSample,
// LCAdaptiveGrid.cpp:374-451; the loader, LCSampledFunction.cpp:24-47, links border cells to leaves only, and refineCell
// on such a grid would treat the ring samples' functions as locality violations).
#ifndef LOCO_GRID_HPP
#define LOCO_GRID_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "loco_b200.h"

namespace loco_grid {

constexpr double kEpsilon = 0.000001;             // LCMathHelper::EPSILON (LCMathHelper.cpp:84)
constexpr double kContainingEpsilon = 0.0000001;  // LCAdaptiveGridCell::CONTAINING_EPSILON (LCAdaptiveGridCell.h:108)

struct Status {
	bool ok = true;
	std::string msg = "no error";
	Status() {}
	explicit Status(const std::string &m) : ok(false), msg(m) {}
};
#define LOCO_GRID_RETURN_IF_ERROR(expr) do { ::loco_grid::Status _s = (expr); if (!_s.ok) return _s; } while (0)

class Grid {
public:
	struct Basis {  // LCBasisFunction: center_, support_, weight_
		std::vector<double> c, s;
		double w = 0;
	};
	struct Sample {  // LCSample: center_, shapeInfo_ (a row of the value table), basisFunctions_
		std::vector<double> x;
		int32_t row = -1;
		std::vector<Basis> bf;
	};
	struct Cell {  // LCAdaptiveGridCell
		int32_t level = 0, parent = -1, child1 = -1, child2 = -1, split_dir = -1;
		double split_val = 0;
		bool leaf = true;
		std::vector<double> box;                           // ranges_ [2N]
		std::vector<int32_t> nb;                           // neighbors_ (a list: order kept)
		int32_t center_row = -1;                           // centerShapeInfo_
		std::vector<std::pair<int32_t, int32_t>> samples;  // samples_: sample -> value row of the per-cell copy
	};

	Grid() {}

	// The grid state of an LCSampledFunction given as plain arrays (what loco_model_from_graph takes; e.g. a bootstrap grid).
	// fn evaluates the true function at ORIGINAL parameters (LCFunction::evalShapeInfo); it is called for new samples and
	// for the centres of new leaves only.
	static Status fromGraph(const loco_graph_t &g, loco_value_fn fn, void *user, Grid *out)
	{
		Grid &G = *out;
		G = Grid();
		if (g.n_params < 1 || g.value_dim < 1 || !g.domain || g.n_cells < 1) return Status("invalid graph");
		G.N_ = g.n_params;
		G.D_ = g.value_dim;
		G.basis_ = g.basis_type;
		G.fn_ = fn;
		G.user_ = user;
		G.domain_.assign(g.domain, g.domain + 2 * G.N_);
		const int N = G.N_, D = G.D_;
		const int64_t S = g.n_samples;
		if (!g.sample_value) return Status("the builder needs the sample values");
		G.values_.assign(g.sample_value, g.sample_value + S * D);
		G.samples_.resize((size_t)S);
		std::vector<int32_t> first_of_group;
		for (int64_t i = 0; i < S; i++) {
			Sample &s = G.samples_[(size_t)i];
			s.x.assign(g.sample_center + i * N, g.sample_center + (i + 1) * N);
			s.row = (int32_t)i;
			if (g.sample_value_group) {  // samples that share one LCFunctionValue object share a row (LCAdaptiveGrid.cpp:345-355)
				const int32_t grp = g.sample_value_group[i];
				if (grp >= 0) {
					if ((size_t)grp >= first_of_group.size()) first_of_group.resize((size_t)grp + 1, -1);
					if (first_of_group[(size_t)grp] < 0) first_of_group[(size_t)grp] = (int32_t)i;
					s.row = first_of_group[(size_t)grp];
				}
			}
			for (int64_t b = g.bf_offset[i]; b < g.bf_offset[i + 1]; b++) {
				Basis f;
				f.c.assign(g.bf_center + b * N, g.bf_center + (b + 1) * N);
				f.s.assign(g.bf_support + b * N, g.bf_support + (b + 1) * N);
				f.w = g.bf_weight[b];
				s.bf.push_back(std::move(f));
			}
		}
		auto take_cells = [&](int64_t C, const double *ranges, const double *center_value, const int64_t *off, const int32_t *smp,
		                      const double *copies, bool border) {
			const size_t base = G.cells_.size();
			G.cells_.resize(base + (size_t)C);
			for (int64_t i = 0; i < C; i++) {
				Cell &c = G.cells_[base + (size_t)i];
				c.box.assign(ranges + i * 2 * N, ranges + (i + 1) * 2 * N);
				c.level = border ? -1 : 0;
				const bool is_leaf = border || g.cell_split_dir[i] < 0;
				if (center_value && is_leaf) c.center_row = G.newRow(center_value + i * D);
				if (is_leaf)
					for (int64_t e = off[i]; e < off[i + 1]; e++) {
						const int32_t s = smp[e];
						int32_t row = G.samples_[(size_t)s].row;
						if (copies && !std::equal(copies + e * D, copies + (e + 1) * D, G.values_.begin() + (size_t)row * D)) row = G.newRow(copies + e * D);
						c.samples.emplace_back(s, row);
					}
				if (border) G.border_.push_back((int32_t)(base + (size_t)i));
			}
		};
		take_cells(g.n_cells, g.cell_ranges, g.cell_center_value, g.cell_sample_offset, g.cell_sample, g.cell_sample_value, false);
		for (int64_t i = 0; i < g.n_cells; i++) {
			Cell &c = G.cells_[(size_t)i];
			if (g.cell_split_dir[i] < 0) continue;
			c.leaf = false;
			c.split_dir = g.cell_split_dir[i];
			c.split_val = g.cell_split_val[i];
			c.child1 = (int32_t)i + 1;
			c.child2 = g.cell_child2[i];
			if (c.child2 <= c.child1 || c.child2 >= g.n_cells) return Status("invalid graph: child index");
			for (int32_t ch : {c.child1, c.child2}) {
				G.cells_[(size_t)ch].parent = (int32_t)i;
				G.cells_[(size_t)ch].level = c.level + 1;
			}
		}
		if (g.n_border_cells > 0)
			take_cells(g.n_border_cells, g.border_ranges, g.border_center_value, g.border_sample_offset, g.border_sample, g.border_sample_value, true);
		G.rebuildNeighbors();
		return Status();
	}

	// LCAdaptiveGrid::bootstrapSample (LCAdaptiveGrid.cpp:256-456): 2^level cells per dimension over the cube [-1,1]^N, one
	// sample with ONE unit-weight basis function per lattice point (support = cell size, cubic: twice that), the uniform
	// k-d tree of bootstrapUniformSplit (:234-249: directions 0..N-1 cyclically, `level` rounds); cubic: the ring of border
	// samples (value object shared with the clamped in-domain sample unless eval_corners) and of border cells of level -1.
	static Status bootstrap(int n_params, int basis_type, int level, int value_dim, bool eval_corners, const double *domain, loco_value_fn fn,
	                        void *user, Grid *out)
	{
		Grid &G = *out;
		G = Grid();
		if (n_params < 1 || value_dim < 1 || !domain || !fn) return Status("invalid argument");
		if (level < 0) return Status("number of levels must be positive");
		if (basis_type != LOCO_LINEAR_BSPLINE && basis_type != LOCO_CUBIC_BSPLINE) return Status("basis type unknown");
		if ((double)level * n_params > 40) return Status("bootstrap grid too large");
		G.N_ = n_params;
		G.D_ = value_dim;
		G.basis_ = basis_type;
		G.fn_ = fn;
		G.user_ = user;
		G.domain_.assign(domain, domain + 2 * n_params);
		const int N = n_params;
		const int64_t n = (int64_t)1 << level;
		const double cell = 2 / (double)n;
		auto lattice = [&](int64_t i) {  // shapeMin * (1 - alpha) + shapeMax * alpha with alpha = i / n (:274-275)
			const double alpha = (double)i / (double)n;
			return (-1.0) * (1.0 - alpha) + 1.0 * alpha;
		};
		auto unit_function = [&](const std::vector<double> &x) {
			Basis f;
			f.c = x;
			f.s.assign((size_t)N, basis_type == LOCO_CUBIC_BSPLINE ? 2 * cell : cell);
			f.w = 1.0;
			return f;
		};
		// in-domain samples, dimension 0 fastest (computeCombinations, LCMathHelper.cpp:4-27)
		int64_t n_lattice = 1;
		for (int d = 0; d < N; d++) n_lattice *= n + 1;
		for (int64_t k = 0; k < n_lattice; k++) {
			std::vector<double> x((size_t)N);
			int64_t num = k;
			for (int d = 0; d < N; d++) {
				x[(size_t)d] = lattice(num % (n + 1));
				num /= n + 1;
			}
			const int32_t s = G.newSample(x);
			G.samples_[(size_t)s].bf.push_back(unit_function(x));
		}
		// the uniform tree
		Cell root;
		root.box.resize((size_t)2 * N);
		for (int d = 0; d < N; d++) {
			root.box[2 * d] = -1;
			root.box[2 * d + 1] = 1;
		}
		G.cells_.push_back(root);
		struct Todo { int32_t cell; int dir, levels; };
		std::vector<Todo> todo(1, Todo{0, 0, level});
		while (!todo.empty()) {
			Todo t = todo.back();
			todo.pop_back();
			if (t.levels <= 0) continue;
			const int next = (t.dir + 1) % N;
			if (next == 0) t.levels--;
			G.split(t.cell, t.dir);
			todo.push_back(Todo{G.cells_[(size_t)t.cell].child2, next, t.levels});
			todo.push_back(Todo{G.cells_[(size_t)t.cell].child1, next, t.levels});
		}
		// leaves: the lattice samples at their corners (what cointainsPoint selects among all samples, :212-221) + centre value
		auto lattice_index = [&](double x) { return (int64_t)std::llround((x + 1.0) / cell); };
		for (int32_t leaf : G.leaves()) {
			std::vector<int32_t> corners;
			const std::vector<double> box = G.cells_[(size_t)leaf].box;
			for (int64_t k = 0; k < ((int64_t)1 << N); k++) {
				int64_t idx = 0, stride = 1;
				for (int d = 0; d < N; d++) {
					idx += (lattice_index(box[2 * d]) + ((k >> d) & 1)) * stride;
					stride *= n + 1;
				}
				corners.push_back((int32_t)idx);
			}
			std::sort(corners.begin(), corners.end());
			G.setLeafInformation(leaf, corners);
		}
		if (basis_type == LOCO_CUBIC_BSPLINE) {
			// border samples: the lattice points of [-1, n+1]^N with a coordinate at -1 or n+1 (computeCornerCombinations(N, n+3), :322-360)
			int64_t n_ring = 1;
			for (int d = 0; d < N; d++) n_ring *= n + 3;
			for (int64_t k = 0; k < n_ring; k++) {
				std::vector<double> x((size_t)N);
				int64_t num = k, ref = 0, stride = 1;
				bool on_ring = false;
				for (int d = 0; d < N; d++) {
					const int64_t c = num % (n + 3);
					num /= n + 3;
					on_ring = on_ring || c == 0 || c == n + 2;
					x[(size_t)d] = lattice(c - 1);
					ref += std::min(std::max(c - 1, (int64_t)0), n) * stride;
					stride *= n + 1;
				}
				if (!on_ring) continue;
				Sample s;
				s.x = x;
				s.row = eval_corners ? G.evalTrueValue(x) : G.samples_[(size_t)ref].row;
				s.bf.push_back(unit_function(x));
				G.samples_.push_back(std::move(s));
			}
			// border cells: the cells of the lattice [-1, n]^N with an index at -1 or n (computeCornerCombinations(N, n+2), :362-418)
			int64_t n_cells = 1;
			for (int d = 0; d < N; d++) n_cells *= n + 2;
			for (int64_t k = 0; k < n_cells; k++) {
				Cell c;
				c.level = -1;
				c.box.resize((size_t)2 * N);
				int64_t num = k;
				bool on_ring = false;
				for (int d = 0; d < N; d++) {
					const int64_t i = num % (n + 2);
					num /= n + 2;
					on_ring = on_ring || i == 0 || i == n + 1;
					c.box[2 * d] = lattice(i - 1);
					c.box[2 * d + 1] = c.box[2 * d] + (1.0 - (-1.0)) / (double)n;
				}
				if (!on_ring) continue;
				G.cells_.push_back(std::move(c));
				const int32_t ci = (int32_t)G.cells_.size() - 1;
				G.border_.push_back(ci);
				std::vector<double> center((size_t)N);
				for (int d = 0; d < N; d++) {
					const std::vector<double> &b = G.cells_[(size_t)ci].box;
					center[(size_t)d] = std::min(1.0, std::max(-1.0, 0.5 * (b[2 * d + 1] + b[2 * d])));  // clamped into the domain (:391-396)
				}
				G.cells_[(size_t)ci].center_row = G.evalTrueValue(center);
				for (size_t s = 0; s < G.samples_.size(); s++)
					if (G.containsPoint(ci, G.samples_[s].x)) G.addSample(ci, (int32_t)s);
			}
			G.linkBorderCells();
		}
		return Status();
	}

	// LCAdaptiveGridCell::getOptimalSplitDirection (LCAdaptiveGridCell.cpp:977-1041, the ERROR_BASED split policy): the direction
	// with the largest summed difference between the values at the two ends of the cell's edges (computeDifference: |a - b|,
	// summed over the components of a vector value).  Needs a sample at every corner of the leaf.
	Status optimalSplitDirection(int32_t leaf, int *direction) const
	{
		const std::vector<int32_t> lv = leaves();
		if (leaf < 0 || (size_t)leaf >= lv.size()) return Status("leaf index out of range");
		const Cell &c = cells_[(size_t)lv[(size_t)leaf]];
		std::vector<int32_t> corner_row((size_t)1 << N_, -1);
		for (const auto &e : c.samples) {
			int id = 0;
			bool is_corner = true;
			for (int d = 0; d < N_ && is_corner; d++) {  // getCornerId (:953-972)
				const double x = samples_[(size_t)e.first].x[(size_t)d];
				if (x > c.box[2 * d + 1] - kContainingEpsilon && x < c.box[2 * d + 1] + kContainingEpsilon) id += 1 << d;
				else if (x < c.box[2 * d] - kContainingEpsilon || x > c.box[2 * d] + kContainingEpsilon) is_corner = false;
			}
			if (is_corner) corner_row[(size_t)id] = e.second;
		}
		*direction = 0;
		double best = 0;
		for (int i = 0; i < N_; i++) {
			double sum = 0;
			for (int k = 0; k < (1 << N_); k++) {
				if (k & (1 << i)) continue;
				const int32_t a = corner_row[(size_t)k], b = corner_row[(size_t)(k | (1 << i))];
				if (a < 0 || b < 0) return Status("could not find corner shape");
				for (int j = 0; j < D_; j++) sum += std::abs(values_[(size_t)a * D_ + j] - values_[(size_t)b * D_ + j]);
			}
			if (sum > best) {
				best = sum;
				*direction = i;
			}
		}
		return Status();
	}

	int nParams() const { return N_; }
	int valueDim() const { return D_; }
	int64_t nSamples() const { return (int64_t)samples_.size(); }
	int64_t nLeaves() const { return (int64_t)leaves().size(); }
	int64_t nFunctionEvaluations() const { return n_fn_calls_; }
	const std::vector<Sample> &samples() const { return samples_; }
	const std::vector<Cell> &cells() const { return cells_; }

	// leaf cells in LCAdaptiveGridCell::getAllLeafCells order (:247-258) = the leaf numbering of the flat model
	std::vector<int32_t> leaves() const
	{
		std::vector<int32_t> out, stack(1, 0);
		while (!stack.empty()) {
			const int32_t c = stack.back();
			stack.pop_back();
			if (cells_[(size_t)c].leaf) { out.push_back(c); continue; }
			stack.push_back(cells_[(size_t)c].child2);
			stack.push_back(cells_[(size_t)c].child1);
		}
		return out;
	}

	// LCAdaptiveGrid::refineCell for the leaf with index `leaf` of the CURRENT leaf numbering, split in cube direction `direction`
	Status refineLeaf(int32_t leaf, int direction)
	{
		const std::vector<int32_t> lv = leaves();
		if (leaf < 0 || (size_t)leaf >= lv.size()) return Status("leaf index out of range");
		return refineCell(lv[(size_t)leaf], direction);
	}

	// The requests of one level-synchronous pass (LCAdaptiveGrid::selectSplits of loco_lc.hpp / refine.select_splits): leaf
	// indices of ONE numbering, in descending order, so that splitting one leaves the indices of those still to come valid.
	Status refineLeaves(const int32_t *leaf, const int32_t *direction, int64_t n)
	{
		const std::vector<int32_t> lv = leaves();
		for (int64_t i = 0; i < n; i++) {
			if (leaf[i] < 0 || (size_t)leaf[i] >= lv.size()) return Status("leaf index out of range");
			if (i && leaf[i] >= leaf[i - 1]) return Status("the requests of a pass must come in descending leaf order");
		}
		for (int64_t i = 0; i < n; i++) LOCO_GRID_RETURN_IF_ERROR(refineCell(lv[(size_t)leaf[i]], direction[i]));
		return Status();
	}

	Status refineCell(int32_t ci, int dir)
	{
		if (ci < 0 || (size_t)ci >= cells_.size() || !cells_[(size_t)ci].leaf) return Status("not a leaf cell");
		if (cells_[(size_t)ci].level < 0) return Status("cannot split border cells");  // LCAdaptiveGridCell.cpp:107-110
		if (dir < 0 || dir >= N_) return Status("split direction out of range");
		if (!fn_) return Status("refineCell needs the function to sample");
		const int N = N_;
		// --- step 1 (:634-656): samples at the midpoints of the cell's edges parallel to dir; evaluate f where none exists
		std::vector<int32_t> new_mid, extra_mid;
		for (const std::vector<double> &mv : midValues(ci, dir))
			if (cellSample(ci, mv) < 0) new_mid.push_back(newSample(mv));
		// --- step 2 (:658-688): refine the basis functions of the samples affecting the cell
		const std::vector<double> cell_box = cells_[(size_t)ci].box;
		double desired = 0.5 * (cell_box[2 * dir + 1] - cell_box[2 * dir]);
		std::vector<int32_t> affecting;
		for (const auto &e : cells_[(size_t)ci].samples) affecting.push_back(e.first);
		if (basis_ == LOCO_CUBIC_BSPLINE) {
			desired = 2 * desired;
			for (int32_t nb : cells_[(size_t)ci].nb)  // getDoubleAdjacentSamples (:781-793)
				for (const auto &e : cells_[(size_t)nb].samples) affecting.push_back(e.first);
		} else if (basis_ != LOCO_LINEAR_BSPLINE)
			return Status("basis type unknown");
		std::sort(affecting.begin(), affecting.end());
		affecting.erase(std::unique(affecting.begin(), affecting.end()), affecting.end());
		for (int32_t s : affecting) LOCO_GRID_RETURN_IF_ERROR(refineBasisFunctions(s, cell_box, dir, desired));
		// --- step 3 (:690-697): split, children take the affecting + new samples they contain and the true value at their centre
		split(ci, dir);
		std::vector<int32_t> candidates = affecting;
		candidates.insert(candidates.end(), new_mid.begin(), new_mid.end());
		setLeafInformation(cells_[(size_t)ci].child1, candidates);
		setLeafInformation(cells_[(size_t)ci].child2, candidates);
		// --- step 4 (:699-709): the (former) neighbours that contain a new sample get it
		const std::vector<int32_t> old_nb = cells_[(size_t)ci].nb;
		for (int32_t nb : old_nb)
			for (int32_t s : new_mid)
				if (containsPoint(nb, samples_[(size_t)s].x)) addSample(nb, s);
		// --- locality repair (:712-822): a basis function of an affecting sample that reaches a cell outside the sample's
		// adjacent (linear) / doubly adjacent (cubic) cells is taken off the sample
		struct Group {  // GroupFunctionInfo keyed by LCBasisFunctionKey (LCSample.h:14-36)
			std::vector<double> c, s, center_sum;
			double weight_sum = 0;
		};
		std::vector<Group> groups;
		auto group_add = [&](Group &gr, const Sample &smp, double w) {  // GroupFunctionInfo::add (LCSample.cpp:9-21)
			if (gr.center_sum.empty()) {
				gr.center_sum.resize((size_t)N);
				for (int d = 0; d < N; d++) gr.center_sum[(size_t)d] = smp.x[(size_t)d] * w;
			} else
				for (int d = 0; d < N; d++) gr.center_sum[(size_t)d] += smp.x[(size_t)d] * w;
			gr.weight_sum += w;
		};
		mark_.assign(cells_.size(), 0);
		epoch_ = 0;
		for (int32_t s : affecting) {
			std::vector<int32_t> allowed;
			if (basis_ == LOCO_LINEAR_BSPLINE) LOCO_GRID_RETURN_IF_ERROR(adjacentCells(samples_[(size_t)s].x, &allowed));
			else LOCO_GRID_RETURN_IF_ERROR(doubleAdjacentCells(samples_[(size_t)s].x, &allowed));
			++epoch_;
			for (int32_t c : allowed) mark_[(size_t)c] = epoch_;
			std::vector<Basis> &bf = samples_[(size_t)s].bf;
			std::vector<Basis> kept;
			for (Basis &f : bf) {
				if (!reachesUnmarkedCell(f)) { kept.push_back(std::move(f)); continue; }
				// grouping of the removed functions (:825-851)
				Group *gr = nullptr;
				for (Group &g2 : groups)
					if (sameKey(g2.c, g2.s, f.c, f.s)) { gr = &g2; break; }
				if (!gr) {
					groups.emplace_back();
					gr = &groups.back();
					gr->c = f.c;
					gr->s = f.s;
				}
				group_add(*gr, samples_[(size_t)s], f.w);
			}
			bf.swap(kept);
		}
		// every sample of the grid hands its function of a grouped key over to the group (:853-863, extractGroupBasisFunctions)
		// (the reference looks every group up in every sample's map; here the groups are indexed by the bucket of their first
		// centre coordinate, buckets twice EPSILON wide, so the keys that can be equal to a function are in three buckets)
		if (!groups.empty()) {
			const double bucket_width = 2 * kEpsilon;
			std::unordered_map<int64_t, std::vector<int32_t>> by_bucket;
			for (size_t k = 0; k < groups.size(); k++) by_bucket[(int64_t)std::floor(groups[k].c[0] / bucket_width)].push_back((int32_t)k);
			std::vector<uint8_t> taken(groups.size());
			std::vector<double> lo((size_t)N, 1e300), hi((size_t)N, -1e300);  // box of the grouped centres: most functions of the grid are outside
			for (const Group &gr : groups)
				for (int d = 0; d < N; d++) {
					lo[(size_t)d] = std::min(lo[(size_t)d], gr.c[(size_t)d] - kEpsilon);
					hi[(size_t)d] = std::max(hi[(size_t)d], gr.c[(size_t)d] + kEpsilon);
				}
			for (Sample &smp : samples_) {
				std::fill(taken.begin(), taken.end(), 0);
				size_t kept = 0;
				for (size_t b = 0; b < smp.bf.size(); b++) {
					bool inside = true;
					for (int d = 0; d < N && inside; d++) inside = smp.bf[b].c[(size_t)d] > lo[(size_t)d] && smp.bf[b].c[(size_t)d] < hi[(size_t)d];
					int32_t hit = -1;
					const int64_t bucket = inside ? (int64_t)std::floor(smp.bf[b].c[0] / bucket_width) : 0;
					for (int64_t probe = bucket - 1; inside && probe <= bucket + 1 && hit < 0; probe++) {
						auto it = by_bucket.find(probe);
						if (it == by_bucket.end()) continue;
						for (int32_t k : it->second)
							if (!taken[(size_t)k] && sameKey(groups[(size_t)k].c, groups[(size_t)k].s, smp.bf[b].c, smp.bf[b].s)) { hit = k; break; }
					}
					if (hit >= 0) {  // one function per sample and group, as a lookup in the sample's map gives
						taken[(size_t)hit] = 1;
						group_add(groups[(size_t)hit], smp, smp.bf[b].w);
						continue;
					}
					if (kept != b) smp.bf[kept] = std::move(smp.bf[b]);
					kept++;
				}
				smp.bf.resize(kept);
			}
		}
		// each group is re-attached, multilinearly weighted, to the corner samples of the intersection of the cells adjacent to
		// the group's weighted sample centre (:865-893)
		for (const Group &gr : groups) {
			std::vector<double> center((size_t)N);
			for (int d = 0; d < N; d++) center[(size_t)d] = gr.center_sum[(size_t)d] / gr.weight_sum;
			std::vector<std::pair<int32_t, double>> rw;
			LOCO_GRID_RETURN_IF_ERROR(replacingSamples(center, &extra_mid, &rw));
			for (const auto &r : rw)
				if (r.second > 0) {
					Basis f;
					f.c = gr.c;
					f.s = gr.s;
					f.w = gr.weight_sum * r.second;
					addBasisFunction(r.first, std::move(f));
				}
		}
		// --- step 5 (:895-911): the samples created by the repair go to both children and to the neighbours containing them
		for (int32_t s : extra_mid) {
			addSample(cells_[(size_t)ci].child1, s);
			addSample(cells_[(size_t)ci].child2, s);
		}
		for (int32_t nb : old_nb)
			for (int32_t s : extra_mid)
				if (containsPoint(nb, samples_[(size_t)s].x)) addSample(nb, s);
		cells_[(size_t)ci].samples.clear();  // adaptiveSample: cell->deleteAllSamples() (:522)
		cells_[(size_t)ci].center_row = -1;
		return Status();
	}

	// plain-array image of the current state for loco_model_from_graph / loco_model_update_from_graph (valid until the next
	// call on this object): tree cells in pre-order, samples in creation order, per-cell sample lists sorted by sample index
	const loco_graph_t &graph()
	{
		Export &x = export_;
		x = Export();
		const int N = N_, D = D_;
		const size_t S = samples_.size();
		x.bf_off.assign(1, 0);
		for (const Sample &s : samples_) {
			x.sample_center.insert(x.sample_center.end(), s.x.begin(), s.x.end());
			x.sample_value.insert(x.sample_value.end(), values_.begin() + (size_t)s.row * D, values_.begin() + ((size_t)s.row + 1) * D);
			x.sample_group.push_back(s.row);
			for (const Basis &f : s.bf) {
				x.bf_center.insert(x.bf_center.end(), f.c.begin(), f.c.end());
				x.bf_support.insert(x.bf_support.end(), f.s.begin(), f.s.end());
				x.bf_weight.push_back(f.w);
			}
			x.bf_off.push_back((int64_t)x.bf_weight.size());
		}
		// per-cell value copies are exported only if some cell holds a copy that is not its sample's own value object (grids
		// taken over from arrays with such copies): a natively built grid hands over none (E x D doubles saved)
		bool with_copies = false;
		for (const Cell &c : cells_)
			for (const auto &e : c.samples) with_copies = with_copies || e.second != samples_[(size_t)e.first].row;
		auto put_cell = [&](const Cell &c, std::vector<double> &ranges, std::vector<double> &center, std::vector<int64_t> &off, std::vector<int32_t> &smp,
		                    std::vector<double> &copies) {
			ranges.insert(ranges.end(), c.box.begin(), c.box.end());
			if (c.center_row >= 0) center.insert(center.end(), values_.begin() + (size_t)c.center_row * D, values_.begin() + ((size_t)c.center_row + 1) * D);
			else center.insert(center.end(), (size_t)D, 0.0);
			std::vector<std::pair<int32_t, int32_t>> sorted = c.samples;
			std::sort(sorted.begin(), sorted.end());
			for (const auto &e : sorted) {
				smp.push_back(e.first);
				if (with_copies) copies.insert(copies.end(), values_.begin() + (size_t)e.second * D, values_.begin() + ((size_t)e.second + 1) * D);
			}
			off.push_back((int64_t)smp.size());
		};
		x.cell_off.assign(1, 0);
		x.border_off.assign(1, 0);
		std::vector<int32_t> order, stack(1, 0), pos(cells_.size(), -1);
		while (!stack.empty()) {
			const int32_t c = stack.back();
			stack.pop_back();
			pos[(size_t)c] = (int32_t)order.size();
			order.push_back(c);
			if (cells_[(size_t)c].leaf) continue;
			stack.push_back(cells_[(size_t)c].child2);
			stack.push_back(cells_[(size_t)c].child1);
		}
		for (int32_t ci : order) {
			const Cell &c = cells_[(size_t)ci];
			put_cell(c, x.cell_ranges, x.cell_center, x.cell_off, x.cell_sample, x.cell_copies);
			x.cell_split_dir.push_back(c.leaf ? -1 : c.split_dir);
			x.cell_split_val.push_back(c.leaf ? 0.0 : c.split_val);
			x.cell_child2.push_back(c.leaf ? -1 : pos[(size_t)c.child2]);
		}
		for (int32_t ci : border_) put_cell(cells_[(size_t)ci], x.border_ranges, x.border_center, x.border_off, x.border_sample, x.border_copies);
		auto have_centers = [&](const std::vector<int32_t> &ids) {
			for (int32_t ci : ids)
				if (cells_[(size_t)ci].leaf && cells_[(size_t)ci].center_row < 0) return false;
			return true;
		};
		loco_graph_t &g = x.g;
		g = loco_graph_t();
		g.n_params = N;
		g.value_dim = D;
		g.basis_type = basis_;
		g.domain = domain_.data();
		g.n_samples = (int64_t)S;
		g.sample_center = x.sample_center.data();
		g.sample_value = x.sample_value.data();
		g.sample_value_group = x.sample_group.data();
		g.bf_offset = x.bf_off.data();
		g.bf_center = x.bf_center.data();
		g.bf_support = x.bf_support.data();
		g.bf_weight = x.bf_weight.data();
		g.n_cells = (int64_t)order.size();
		g.cell_ranges = x.cell_ranges.data();
		g.cell_split_dir = x.cell_split_dir.data();
		g.cell_split_val = x.cell_split_val.data();
		g.cell_child2 = x.cell_child2.data();
		g.cell_center_value = have_centers(order) ? x.cell_center.data() : nullptr;  // a grid taken over without centre values has none to give
		g.cell_sample_offset = x.cell_off.data();
		g.cell_sample = x.cell_sample.data();
		g.cell_sample_value = with_copies ? x.cell_copies.data() : nullptr;
		g.n_border_cells = (int64_t)border_.size();
		g.border_ranges = x.border_ranges.data();
		g.border_center_value = have_centers(border_) ? x.border_center.data() : nullptr;
		g.border_sample_offset = x.border_off.data();
		g.border_sample = x.border_sample.data();
		g.border_sample_value = with_copies ? x.border_copies.data() : nullptr;
		return g;
	}

	// neighbour cells of a cell (introspection for tests)
	const std::vector<int32_t> &neighbors(int32_t cell) const { return cells_[(size_t)cell].nb; }
	const std::vector<int32_t> &borderCells() const { return border_; }

private:
	struct Export {
		loco_graph_t g;
		std::vector<double> sample_center, sample_value, bf_center, bf_support, bf_weight;
		std::vector<int32_t> sample_group;
		std::vector<int64_t> bf_off;
		std::vector<double> cell_ranges, cell_split_val, cell_center, cell_copies;
		std::vector<int32_t> cell_split_dir, cell_child2, cell_sample;
		std::vector<int64_t> cell_off;
		std::vector<double> border_ranges, border_center, border_copies;
		std::vector<int32_t> border_sample;
		std::vector<int64_t> border_off;
	};

	int32_t newRow(const double *v)
	{
		values_.insert(values_.end(), v, v + D_);
		return (int32_t)(values_.size() / (size_t)D_) - 1;
	}
	// parametricShape_->evalShapeInfo(parametricShape_->mapFromStandardHypercube(x)) (LCFunction.cpp:75-87)
	int32_t evalTrueValue(const std::vector<double> &cube)
	{
		std::vector<double> p((size_t)N_), v((size_t)D_, 0.0);
		for (int d = 0; d < N_; d++) {
			const double alpha = (cube[(size_t)d] + 1.) / 2.;
			p[(size_t)d] = domain_[2 * d] * (1 - alpha) + domain_[2 * d + 1] * alpha;
		}
		fn_(p.data(), v.data(), user_);
		n_fn_calls_++;
		return newRow(v.data());
	}
	int32_t newSample(const std::vector<double> &x)
	{
		Sample s;
		s.x = x;
		s.row = evalTrueValue(x);
		samples_.push_back(std::move(s));
		return (int32_t)samples_.size() - 1;
	}

	// ---- cells
	bool containsPoint(int32_t ci, const std::vector<double> &p) const  // cointainsPoint (:168-180)
	{
		const std::vector<double> &r = cells_[(size_t)ci].box;
		for (int d = 0; d < N_; d++)
			if (p[(size_t)d] < r[2 * d] - kContainingEpsilon || p[(size_t)d] > r[2 * d + 1] + kContainingEpsilon) return false;
		return true;
	}
	bool parametersAreInCell(int32_t ci, const std::vector<double> &p) const  // evalPametersAreInCell (:438-454)
	{
		const std::vector<double> &r = cells_[(size_t)ci].box;
		for (int d = 0; d < N_; d++)
			if (p[(size_t)d] < r[2 * d] - kEpsilon || p[(size_t)d] > r[2 * d + 1] + kEpsilon) return false;
		return true;
	}
	bool touches(int32_t a, int32_t b) const  // checkIsNeighbor (:910-925)
	{
		const std::vector<double> &A = cells_[(size_t)a].box, &B = cells_[(size_t)b].box;
		for (int d = 0; d < N_; d++)
			if ((B[2 * d] - A[2 * d + 1]) > kContainingEpsilon || (A[2 * d] - B[2 * d + 1]) > kContainingEpsilon) return false;
		return true;
	}
	void replaceNeighborWithChildren(int32_t self, int32_t old, int32_t c1, int32_t c2)  // (:927-944)
	{
		std::vector<int32_t> &nb = cells_[(size_t)self].nb;
		nb.erase(std::remove(nb.begin(), nb.end(), old), nb.end());
		if (touches(self, c1)) nb.push_back(c1);
		if (touches(self, c2)) nb.push_back(c2);
	}
	int32_t newChild(int32_t parent, const std::vector<double> &box)  // LCAdaptiveGridCell(level, parent, ranges) (:14-35)
	{
		Cell c;
		c.level = cells_[(size_t)parent].level + 1;
		c.parent = parent;
		c.box = box;
		cells_.push_back(std::move(c));
		const int32_t id = (int32_t)cells_.size() - 1;
		for (int32_t nb : cells_[(size_t)parent].nb)
			if (touches(id, nb)) cells_[(size_t)id].nb.push_back(nb);
		return id;
	}
	void split(int32_t ci, int dir)  // LCAdaptiveGridCell::split (:103-137)
	{
		const std::vector<double> box = cells_[(size_t)ci].box;
		const double v = 0.5 * (box[2 * dir] + box[2 * dir + 1]);
		std::vector<double> b1 = box, b2 = box;
		b1[2 * dir + 1] = v;
		b2[2 * dir] = v;
		const int32_t c1 = newChild(ci, b1), c2 = newChild(ci, b2);
		Cell &c = cells_[(size_t)ci];
		c.leaf = false;
		c.split_dir = dir;
		c.split_val = v;
		c.child1 = c1;
		c.child2 = c2;
		const std::vector<int32_t> nbs = c.nb;
		for (int32_t nb : nbs) replaceNeighborWithChildren(nb, ci, c1, c2);
		cells_[(size_t)c1].nb.push_back(c2);
		cells_[(size_t)c2].nb.push_back(c1);
	}
	// neighbour lists of a grid taken over as arrays: the tree exactly as addAllNeighboursTopDown builds them (:139-164, cells
	// visited in pre-order: a cell takes the touching members of its parent's list, then hands its own place in its
	// neighbours' lists to its children), border cells linked to every leaf / border cell they touch (the state
	// bootstrapSample leaves behind, see the file header)
	void rebuildNeighbors()
	{
		for (Cell &c : cells_) c.nb.clear();
		std::vector<int32_t> stack(1, 0);
		while (!stack.empty()) {
			const int32_t ci = stack.back();
			stack.pop_back();
			const int32_t parent = cells_[(size_t)ci].parent;
			if (parent >= 0)
				for (int32_t nb : cells_[(size_t)parent].nb)
					if (touches(ci, nb)) cells_[(size_t)ci].nb.push_back(nb);
			if (cells_[(size_t)ci].leaf) continue;
			const int32_t c1 = cells_[(size_t)ci].child1, c2 = cells_[(size_t)ci].child2;
			const std::vector<int32_t> nbs = cells_[(size_t)ci].nb;
			for (int32_t nb : nbs) replaceNeighborWithChildren(nb, ci, c1, c2);
			cells_[(size_t)c1].nb.push_back(c2);
			cells_[(size_t)c2].nb.push_back(c1);
			stack.push_back(c2);
			stack.push_back(c1);
		}
		linkBorderCells();
	}
	void linkBorderCells()
	{
		const std::vector<int32_t> lv = leaves();
		for (int32_t b : border_) {
			for (int32_t l : lv)
				if (touches(b, l)) {
					cells_[(size_t)b].nb.push_back(l);
					cells_[(size_t)l].nb.push_back(b);
				}
			for (int32_t b2 : border_)
				if (b2 != b && touches(b, b2)) cells_[(size_t)b].nb.push_back(b2);
		}
	}

	// ---- samples of a cell
	int32_t cellSample(int32_t ci, const std::vector<double> &x) const  // LCAdaptiveGridCell::getSample (:754-765)
	{
		for (const auto &e : cells_[(size_t)ci].samples)
			if (distance(samples_[(size_t)e.first].x, x) < kContainingEpsilon) return e.first;
		return -1;
	}
	int32_t gridSample(const std::vector<double> &x) const  // LCAdaptiveGrid::getSample (:459-472)
	{
		for (size_t i = 0; i < samples_.size(); i++)
			if (distance(samples_[i].x, x) < kEpsilon) return (int32_t)i;
		return -1;
	}
	void addSample(int32_t ci, int32_t s)  // LCAdaptiveGridCell::addSample (:414-421): samples_[sample] = copy of its value
	{
		for (auto &e : cells_[(size_t)ci].samples)
			if (e.first == s) { e.second = samples_[(size_t)s].row; return; }
		cells_[(size_t)ci].samples.emplace_back(s, samples_[(size_t)s].row);
	}
	void setLeafInformation(int32_t ci, const std::vector<int32_t> &candidates)  // (:212-231)
	{
		std::vector<double> center((size_t)N_);
		const std::vector<double> box = cells_[(size_t)ci].box;
		for (int d = 0; d < N_; d++) center[(size_t)d] = 0.5 * (box[2 * d + 1] + box[2 * d]);
		const int32_t row = evalTrueValue(center);
		cells_[(size_t)ci].center_row = row;
		for (int32_t s : candidates)
			if (containsPoint(ci, samples_[(size_t)s].x)) addSample(ci, s);
	}
	std::vector<std::vector<double>> midValues(int32_t ci, int dir) const  // getMidValues (:715-751)
	{
		const std::vector<double> &r = cells_[(size_t)ci].box;
		std::vector<std::vector<double>> out;
		const int64_t n = (int64_t)1 << (N_ - 1);
		for (int64_t k = 0; k < n; k++) {
			std::vector<double> x((size_t)N_);
			int64_t num = k;
			for (int d = 0; d < N_; d++) {
				if (d == dir) { x[(size_t)d] = 0.5 * (r[2 * d] + r[2 * d + 1]); continue; }
				const int bit = (int)(num & 1);
				num >>= 1;
				x[(size_t)d] = (1 - bit) * r[2 * d] + bit * r[2 * d + 1];
			}
			out.push_back(std::move(x));
		}
		return out;
	}

	// ---- basis functions
	static double distance(const std::vector<double> &a, const std::vector<double> &b)
	{
		double s = 0;
		for (size_t i = 0; i < a.size(); i++) s += (a[i] - b[i]) * (a[i] - b[i]);
		return std::sqrt(s);
	}
	static bool sameKey(const std::vector<double> &c1, const std::vector<double> &s1, const std::vector<double> &c2, const std::vector<double> &s2)
	{
		// LCBasisFunctionKey::operator== (LCBasisFunction.h:88-92): both Euclidean distances below EPSILON.  A coordinate that
		// differs by EPSILON or more already decides it (the distance is at least any coordinate difference).
		if (!(std::abs(c1[0] - c2[0]) < kEpsilon) || !(std::abs(s1[0] - s2[0]) < kEpsilon)) return false;
		return distance(c1, c2) < kEpsilon && distance(s1, s2) < kEpsilon;
	}
	void addBasisFunction(int32_t s, Basis f)  // LCSample::addBasisFunction (LCSample.cpp:89-103)
	{
		for (Basis &g : samples_[(size_t)s].bf)
			if (sameKey(g.c, g.s, f.c, f.s)) { g.w += f.w; return; }
		samples_[(size_t)s].bf.push_back(std::move(f));
	}
	bool overlaps(const Basis &f, const std::vector<double> &region) const  // checkRegionOverlap (LCBasisFunction.cpp:64-86)
	{
		for (int d = 0; d < N_; d++) {
			const double lo = f.c[(size_t)d] - f.s[(size_t)d], hi = f.c[(size_t)d] + f.s[(size_t)d];
			const double allowed = f.s[(size_t)d] / 100.0;
			if (hi < region[2 * d] + allowed || lo > region[2 * d + 1] - allowed) return false;
		}
		return true;
	}
	// one function -> its dyadic refinement in direction d (LCLinearBSpline::refine :334-360: 1/2, 1, 1/2 of the halved hat;
	// LCCubicBSpline::refine :485-509: 1/8, 1/2, 3/4, 1/2, 1/8); `f` becomes the centre function, the others are appended
	Status refineOne(Basis &f, int d, std::vector<Basis> *out) const
	{
		const double sd = f.s[(size_t)d];
		auto child = [&](double shift, double w) {
			Basis g;
			g.c = f.c;
			g.s = f.s;
			g.c[(size_t)d] += shift;
			g.s[(size_t)d] = 0.5 * sd;
			g.w = w;
			out->push_back(std::move(g));
		};
		if (basis_ == LOCO_LINEAR_BSPLINE) {
			child(-0.5 * sd, f.w * 0.5);
			child(+0.5 * sd, f.w * 0.5);
		} else {
			child(-0.25 * sd, f.w * 1.0 / 2.0);
			child(-0.5 * sd, f.w * 1.0 / 8.0);
			child(+0.25 * sd, f.w * 1.0 / 2.0);
			child(+0.5 * sd, f.w * 1.0 / 8.0);
			f.w *= 6.0 / 8.0;
		}
		f.s[(size_t)d] = 0.5 * sd;
		return Status();
	}
	Status refineBasisFunctions(int32_t s, const std::vector<double> &cell_box, int dir, double desired)  // LCSample.cpp:106-154
	{
		std::vector<Basis> &bf = samples_[(size_t)s].bf;
		std::vector<Basis> kept, fresh;
		for (Basis &f : bf) {
			if (!overlaps(f, cell_box) || (f.s[(size_t)dir] - desired) < 0.00001) { kept.push_back(std::move(f)); continue; }
			if (std::abs(f.s[(size_t)dir] - 2 * desired) > 0.000001) return Status("error, the refinement is not half");
			// refineAllDirections (LCBasisFunction.cpp:209-228): direction by direction, every function made so far is refined
			std::vector<Basis> made;
			for (int d = 0; d < N_; d++) {
				std::vector<Basis> now;
				refineOne(f, d, &now);
				for (Basis &g : made) refineOne(g, d, &now);
				made.insert(made.end(), std::make_move_iterator(now.begin()), std::make_move_iterator(now.end()));
			}
			fresh.insert(fresh.end(), std::make_move_iterator(made.begin()), std::make_move_iterator(made.end()));
			fresh.push_back(std::move(f));
		}
		bf.swap(kept);
		for (Basis &f : fresh) addBasisFunction(s, std::move(f));
		return Status();
	}

	// ---- adjacency queries
	// LCAdaptiveGridCell::getAdjacentLeaves (:310-333), including its early return when a visited leaf does not contain the point
	bool adjacentLeaves(int32_t ci, const std::vector<double> &p, std::vector<int32_t> *out) const
	{
		const Cell &c = cells_[(size_t)ci];
		if (c.leaf) {
			if (!parametersAreInCell(ci, p)) return false;
			if (std::find(out->begin(), out->end(), ci) == out->end()) out->push_back(ci);
			return true;
		}
		if (p[(size_t)c.split_dir] <= c.split_val + kContainingEpsilon)
			if (!adjacentLeaves(c.child1, p, out)) return false;
		if (p[(size_t)c.split_dir] >= c.split_val - kContainingEpsilon)
			if (!adjacentLeaves(c.child2, p, out)) return false;
		return true;
	}
	Status adjacentCells(const std::vector<double> &p, std::vector<int32_t> *out) const  // LCAdaptiveGrid::getAdjacentCells (:141-176)
	{
		const bool inside = adjacentLeaves(0, p, out);
		bool on_border = false;
		for (int32_t b : border_)
			if (parametersAreInCell(b, p)) {
				if (std::find(out->begin(), out->end(), b) == out->end()) out->push_back(b);
				on_border = true;
			}
		if (!on_border && !inside) return Status("Not found on border");
		return Status();
	}
	Status doubleAdjacentCells(const std::vector<double> &p, std::vector<int32_t> *out) const  // getDoubleAdjacentCells (:178-203)
	{
		std::vector<int32_t> adj;
		LOCO_GRID_RETURN_IF_ERROR(adjacentCells(p, &adj));
		*out = adj;
		for (int32_t a : adj) out->insert(out->end(), cells_[(size_t)a].nb.begin(), cells_[(size_t)a].nb.end());
		return Status();
	}
	// does the function overlap a leaf or border cell that is not marked with the current epoch?  (the reference scans all
	// cells of the grid, LCAdaptiveGrid.cpp:763-776; the same predicate evaluated by a range descent of the tree)
	bool reachesUnmarkedCell(const Basis &f) const
	{
		std::vector<int32_t> stack(1, 0);
		while (!stack.empty()) {
			const int32_t ci = stack.back();
			stack.pop_back();
			const Cell &c = cells_[(size_t)ci];
			if (!overlaps(f, c.box)) continue;  // a box that fails the test contains no leaf that passes it
			if (c.leaf) {
				if (mark_[(size_t)ci] != epoch_) return true;
				continue;
			}
			stack.push_back(c.child1);
			stack.push_back(c.child2);
		}
		for (int32_t b : border_)
			if (mark_[(size_t)b] != epoch_ && overlaps(f, cells_[(size_t)b].box)) return true;
		return false;
	}
	// LCAdaptiveGrid::getReplacingSamples (:1002-1122)
	Status replacingSamples(const std::vector<double> &center, std::vector<int32_t> *extra_mid, std::vector<std::pair<int32_t, double>> *weights)
	{
		std::vector<int32_t> adj;
		adjacentCells(center, &adj);  // the reference ignores this error and tests the size only (:1009-1018)
		if (adj.empty()) return Status("adjacent cells is empty");
		std::vector<double> inter = cells_[(size_t)adj[0]].box;
		for (size_t k = 1; k < adj.size(); k++) {  // getIntersection (LCAdaptiveGridCell.cpp:359-378)
			const std::vector<double> &r = cells_[(size_t)adj[k]].box;
			for (int d = 0; d < N_; d++) {
				const double lo = std::max(inter[2 * d], r[2 * d]), hi = std::min(inter[2 * d + 1], r[2 * d + 1]);
				inter[2 * d] = lo;
				inter[2 * d + 1] = hi;
				if (hi < lo - kEpsilon) return Status("empty intersection");
			}
		}
		double sum = 0;
		const int64_t n = (int64_t)1 << N_;
		for (int64_t k = 0; k < n; k++) {
			double w = 1;
			std::vector<double> corner((size_t)N_);
			for (int d = 0; d < N_; d++) {
				double alpha = 1;
				const double area = inter[2 * d + 1] - inter[2 * d];
				if (std::abs(area) > kEpsilon) alpha = (center[(size_t)d] - inter[2 * d]) / area;
				const bool upper = (k >> d) & 1;
				w *= upper ? alpha : (1 - alpha);
				corner[(size_t)d] = upper ? inter[2 * d + 1] : inter[2 * d];
			}
			if (!(w > 0)) continue;
			sum += w;
			int32_t s = gridSample(corner);
			if (s < 0) {
				s = newSample(corner);
				extra_mid->push_back(s);
			}
			weights->emplace_back(s, w);
		}
		if (std::abs(sum - 1) > kEpsilon) return Status("weights do not sum to one: " + std::to_string(sum));
		return Status();
	}

	int N_ = 0, D_ = 1, basis_ = LOCO_LINEAR_BSPLINE;
	loco_value_fn fn_ = nullptr;
	void *user_ = nullptr;
	int64_t n_fn_calls_ = 0;
	std::vector<double> domain_, values_;
	std::vector<Sample> samples_;
	std::vector<Cell> cells_;
	std::vector<int32_t> border_;
	std::vector<uint32_t> mark_;
	uint32_t epoch_ = 0;
	Export export_;
};

} // namespace loco_grid
#endif // LOCO_GRID_HPP
