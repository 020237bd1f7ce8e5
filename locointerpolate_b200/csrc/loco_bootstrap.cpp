// loco_bootstrap.cpp -- direct generator of bootstrap-uniform grids.
//
// Produces the same object graph as the reference's LCAdaptiveGrid::bootstrapSample
// (LCAdaptiveGrid.cpp:256-456) -- samples on the (2^L+1)^N lattice in computeCombinations order
// (LCMathHelper.cpp:4-27, dimension 0 fastest), one unit-weight basis function per sample (:280-296),
// cyclic uniform splits (:234-249), per-leaf sample copies and centre values (:212-231) and, for the
// cubic basis, the ring of border samples (:320-357) and border cells (:359-416) -- but by lattice
// arithmetic instead of the reference's all-pairs containment scans (87 s for 43k samples there).
// The floating-point expressions that place samples, cells and split planes are kept operation for
// operation so that coordinates are bit-identical to the reference's.
#include "loco_bootstrap.h"

#include <algorithm>
#include <cmath>

namespace loco {
namespace {

// position of lattice index j (may be -1 or n+1 for the border ring) in [-1,1] cube coordinates:
// alphaVec = combination / nSamples; min*(1-alpha) + max*alpha with min=-1, max=1   (:270-271, :327-328)
inline double lattice_pos(int j, int n)
{
	double alpha = (double)j / (double)n;
	return (-1.0) * (1.0 - alpha) + (1.0) * alpha;
}

// LCFunction::mapFromStandardHypercube (LCFunction.cpp:75-85)
inline void map_from_cube(const double *cube, const double *domain, int N, double *out)
{
	for (int i = 0; i < N; i++) {
		double alpha = (cube[i] + 1.) / 2.;
		out[i] = domain[2 * i] * (1 - alpha) + domain[2 * i + 1] * alpha;
	}
}

struct Gen {
	int N, n, D;
	bool cubic, eval_corners;
	const double *domain;
	ValueFn fn;
	void *user;
	Graph *g;
	std::vector<int32_t> ext_to_sample;  // extended lattice (n+3)^N -> sample index
	std::vector<double> tmp_params, tmp_val;

	int64_t ipow(int64_t b, int e) const { int64_t r = 1; while (e-- > 0) r *= b; return r; }

	int32_t new_row(const double *cube)
	{
		int32_t row = (int32_t)g->n_rows();
		if (fn) {
			map_from_cube(cube, domain, N, tmp_params.data());
			fn(tmp_params.data(), tmp_val.data(), user);
			g->values.insert(g->values.end(), tmp_val.begin(), tmp_val.end());
		} else {
			g->virtual_rows++;
		}
		return row;
	}

	void add_sample(const double *center, int32_t row)
	{
		double support = (2 / (double)n) * (cubic ? 2.0 : 1.0);  // cellSize (:268) / 2*cellSize (:292)
		for (int i = 0; i < N; i++) {
			g->sample_center.push_back(center[i]);
			g->bf_center.push_back(center[i]);
			g->bf_support.push_back(support);
		}
		g->bf_weight.push_back(1.0);
		g->sample_row.push_back(row);
		g->bf_off.push_back((int64_t)g->bf_weight.size());
	}

	int32_t sample_at(const int *lat) const  // lat[d] in [-1, n+1]
	{
		int64_t id = 0, mul = 1;
		for (int d = 0; d < N; d++) { id += (int64_t)(lat[d] + 1) * mul; mul *= n + 3; }
		return ext_to_sample[(size_t)id];
	}

	// the 2^N lattice corners of a cell whose low corner is lat0 (cointainsPoint, +-1e-7: corners only)
	void add_cell_samples(Graph::CellSet &cs, const int *lat0)
	{
		std::vector<int> lat(N);
		for (int64_t c = 0; c < ((int64_t)1 << N); c++) {
			for (int d = 0; d < N; d++) lat[d] = lat0[d] + (int)((c >> d) & 1);
			int32_t s = sample_at(lat.data());
			if (s < 0) continue;
			cs.ls_sample.push_back(s);
			cs.ls_row.push_back(g->sample_row[s]);  // getHomeophicShapeInfo copies the sample's value
		}
		cs.ls_off.push_back((int64_t)cs.ls_sample.size());
	}

	// bootstrapUniformSplit (:234-249): cells in pre-order
	void split(std::vector<double> &ranges, std::vector<int> &lat0, std::vector<int> &latn, int direction, int levels)
	{
		Graph::CellSet &cs = g->tree;
		int32_t me = (int32_t)cs.split_dir.size();
		cs.ranges.insert(cs.ranges.end(), ranges.begin(), ranges.end());
		cs.split_dir.push_back(-1);
		cs.split_val.push_back(0.0);
		cs.child2.push_back(-1);
		cs.center_row.push_back(-1);
		if (levels <= 0) {
			std::vector<double> center(N);
			for (int i = 0; i < N; i++) center[i] = 0.5 * (ranges[2 * i + 1] + ranges[2 * i]);  // getCenter
			cs.center_row[me] = fn ? new_row(center.data()) : -1;
			add_cell_samples(cs, lat0.data());
			return;
		}
		cs.ls_off.push_back((int64_t)cs.ls_sample.size());
		int next = (direction + 1) % N;
		int sub = next == 0 ? levels - 1 : levels;
		double lo = ranges[2 * direction], hi = ranges[2 * direction + 1];
		double sv = 0.5 * (lo + hi);  // LCAdaptiveGridCell::split (:117)
		cs.split_dir[me] = direction;
		cs.split_val[me] = sv;
		int half = latn[direction] / 2;
		latn[direction] = half;
		ranges[2 * direction + 1] = sv;
		split(ranges, lat0, latn, next, sub);
		cs.child2[me] = (int32_t)cs.split_dir.size();
		ranges[2 * direction + 1] = hi;
		ranges[2 * direction] = sv;
		lat0[direction] += half;
		split(ranges, lat0, latn, next, sub);
		lat0[direction] -= half;
		ranges[2 * direction] = lo;
		latn[direction] = half * 2;
	}

	Status run(int level)
	{
		*g = Graph();
		g->N = N; g->D = D; g->basis = cubic ? CUBIC_BSPLINE : LINEAR_BSPLINE;
		g->domain.assign(domain, domain + 2 * N);
		g->bf_off.push_back(0);
		g->tree.ls_off.push_back(0);
		g->border.ls_off.push_back(0);
		tmp_params.resize(N); tmp_val.resize(D);
		n = 1 << level;
		int64_t ext = ipow(n + 3, N), inner = ipow(n + 1, N);
		if (ext > ((int64_t)1 << 31) || ipow(n, N) > ((int64_t)1 << 28)) return Status("bootstrap grid too large");
		ext_to_sample.assign((size_t)ext, -1);

		std::vector<int> lat(N);
		std::vector<double> center(N);
		// in-domain samples (:263-301)
		for (int64_t i = 0; i < inner; i++) {
			int64_t num = i;
			for (int d = 0; d < N; d++) { lat[d] = (int)(num % (n + 1)); num /= (n + 1); center[d] = lattice_pos(lat[d], n); }
			int32_t row = new_row(center.data());
			int64_t id = 0, mul = 1;
			for (int d = 0; d < N; d++) { id += (int64_t)(lat[d] + 1) * mul; mul *= n + 3; }
			ext_to_sample[(size_t)id] = (int32_t)g->sample_row.size();
			add_sample(center.data(), row);
		}
		// border samples (:320-357): computeCornerCombinations(N, n+3) order
		if (cubic) {
			std::vector<int> clamped(N);
			for (int64_t i = 0; i < ext; i++) {
				int64_t num = i;
				bool corner = false;
				for (int d = 0; d < N; d++) { int v = (int)(num % (n + 3)); num /= (n + 3); lat[d] = v - 1; corner |= (v == 0 || v == n + 2); }
				if (!corner) continue;
				for (int d = 0; d < N; d++) { center[d] = lattice_pos(lat[d], n); clamped[d] = std::min(std::max(lat[d], 0), n); }
				int32_t row = eval_corners ? new_row(center.data()) : g->sample_row[sample_at(clamped.data())];
				ext_to_sample[(size_t)i] = (int32_t)g->sample_row.size();
				add_sample(center.data(), row);
			}
		}
		// tree (:303-318)
		std::vector<double> ranges(2 * N);
		for (int i = 0; i < N; i++) { ranges[2 * i] = -1; ranges[2 * i + 1] = 1; }
		std::vector<int> lat0(N, 0), latn(N, n);
		split(ranges, lat0, latn, 0, level);
		// border cells (:359-416): computeCornerCombinations(N, n+2) order
		if (cubic) {
			int64_t nc = ipow(n + 2, N);
			std::vector<double> cmin(N), cmax(N);
			for (int64_t i = 0; i < nc; i++) {
				int64_t num = i;
				bool corner = false;
				for (int d = 0; d < N; d++) { int v = (int)(num % (n + 2)); num /= (n + 2); lat[d] = v - 1; corner |= (v == 0 || v == n + 1); }
				if (!corner) continue;
				Graph::CellSet &cs = g->border;
				for (int d = 0; d < N; d++) {
					cmin[d] = lattice_pos(lat[d], n);
					cmax[d] = cmin[d] + (1.0 - (-1.0)) / (double)n;  // (:367)
					cs.ranges.push_back(cmin[d]);
					cs.ranges.push_back(cmax[d]);
				}
				cs.split_dir.push_back(-1);
				cs.split_val.push_back(0.0);
				cs.child2.push_back(-1);
				for (int d = 0; d < N; d++) {  // centre clamped into the domain (:391-398)
					double c = 0.5 * (cmax[d] + cmin[d]);
					c = std::max(-1.0, c);
					c = std::min(1.0, c);
					center[d] = c;
				}
				cs.center_row.push_back(fn ? new_row(center.data()) : -1);
				add_cell_samples(cs, lat.data());
			}
		}
		return validate_graph(*g);
	}
};

} // namespace

Status bootstrap_graph(int N, int basis, int level, int D, bool eval_corners, const double *domain,
                       ValueFn fn, void *user, Graph *out)
{
	if (N < 1 || N > 30) return Status("invalid number of parameters");
	if (level < 0) return Status("number of levels must be positive");  // LCAdaptiveGrid.cpp:259-262
	if (level > 20) return Status("bootstrap grid too large");
	if (basis != LINEAR_BSPLINE && basis != CUBIC_BSPLINE) return Status("basis type unknown");
	if (D < 1) return Status("invalid value dimension");
	Gen gen;
	gen.N = N; gen.D = D; gen.cubic = basis == CUBIC_BSPLINE; gen.eval_corners = eval_corners;
	gen.domain = domain; gen.fn = fn; gen.user = user; gen.g = out;
	return gen.run(level);
}

} // namespace loco
