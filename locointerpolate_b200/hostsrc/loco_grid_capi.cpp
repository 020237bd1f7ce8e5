// loco_grid_capi.cpp -- C ABI (include/loco_grid.h) over the header-only host grid (include/loco_grid.hpp).
// Host-only: compiled by g++ into libloco_grid.so, no CUDA; the kernels' library (libloco_b200.so) is not involved.
#include <string>

#include "loco_grid.h"
#include "loco_grid.hpp"

struct loco_grid_handle {
	loco_grid::Grid grid;
	std::string err = "no error";
};

namespace {
thread_local std::string g_ctor_error = "no error";

int fail(loco_grid_t *g, int code, const loco_grid::Status &st)
{
	g->err = st.msg;
	return code;
}
int classify(const std::string &msg)
{
	return msg.find("out of range") != std::string::npos ? LOCO_ERR_RANGE : LOCO_ERR_INVALID;
}
} // namespace

extern "C" {

int loco_grid_bootstrap(int32_t n_params, int32_t basis_type, int32_t level, int32_t value_dim, int32_t eval_corners, const double *domain,
                        loco_value_fn fn, void *user, loco_grid_t **out)
{
	if (!out) return LOCO_ERR_INVALID;
	*out = nullptr;
	loco_grid_t *g = new loco_grid_handle();
	loco_grid::Status st = loco_grid::Grid::bootstrap(n_params, basis_type, level, value_dim, eval_corners != 0, domain, fn, user, &g->grid);
	if (!st.ok) {
		g_ctor_error = st.msg;
		delete g;
		return LOCO_ERR_INVALID;
	}
	*out = g;
	return LOCO_OK;
}

int loco_grid_from_graph(const loco_graph_t *graph, loco_value_fn fn, void *user, loco_grid_t **out)
{
	if (!out || !graph) return LOCO_ERR_INVALID;
	*out = nullptr;
	loco_grid_t *g = new loco_grid_handle();
	loco_grid::Status st = loco_grid::Grid::fromGraph(*graph, fn, user, &g->grid);
	if (!st.ok) {
		g_ctor_error = st.msg;
		delete g;
		return LOCO_ERR_INVALID;
	}
	*out = g;
	return LOCO_OK;
}

void loco_grid_free(loco_grid_t *g) { delete g; }

const char *loco_grid_last_error(const loco_grid_t *g) { return g ? g->err.c_str() : g_ctor_error.c_str(); }

int loco_grid_refine_leaf(loco_grid_t *g, int32_t leaf, int32_t direction)
{
	if (!g) return LOCO_ERR_INVALID;
	loco_grid::Status st = g->grid.refineLeaf(leaf, direction);
	return st.ok ? LOCO_OK : fail(g, classify(st.msg), st);
}

int loco_grid_refine_leaves(loco_grid_t *g, const int32_t *leaf, const int32_t *direction, int64_t n)
{
	if (!g || n < 0 || (n > 0 && (!leaf || !direction))) return LOCO_ERR_INVALID;
	loco_grid::Status st = g->grid.refineLeaves(leaf, direction, n);
	return st.ok ? LOCO_OK : fail(g, classify(st.msg), st);
}

int loco_grid_optimal_split_direction(const loco_grid_t *g, int32_t leaf, int32_t *direction)
{
	if (!g || !direction) return LOCO_ERR_INVALID;
	int d = 0;
	loco_grid::Status st = g->grid.optimalSplitDirection(leaf, &d);
	if (!st.ok) return fail(const_cast<loco_grid_t *>(g), classify(st.msg), st);
	*direction = d;
	return LOCO_OK;
}

int loco_grid_graph(loco_grid_t *g, loco_graph_t *out)
{
	if (!g || !out) return LOCO_ERR_INVALID;
	*out = g->grid.graph();
	return LOCO_OK;
}

int64_t loco_grid_num_leaves(const loco_grid_t *g) { return g ? g->grid.nLeaves() : 0; }
int64_t loco_grid_num_samples(const loco_grid_t *g) { return g ? g->grid.nSamples() : 0; }
int64_t loco_grid_num_function_evaluations(const loco_grid_t *g) { return g ? g->grid.nFunctionEvaluations() : 0; }

int loco_grid_leaf_neighbors(const loco_grid_t *g, int32_t leaf, int32_t *n_leaf, int32_t *n_border)
{
	if (!g || !n_leaf || !n_border) return LOCO_ERR_INVALID;
	const std::vector<int32_t> lv = g->grid.leaves();
	if (leaf < 0 || (size_t)leaf >= lv.size()) return LOCO_ERR_RANGE;
	*n_leaf = *n_border = 0;
	for (int32_t nb : g->grid.neighbors(lv[(size_t)leaf])) (g->grid.cells()[(size_t)nb].level < 0 ? *n_border : *n_leaf)++;
	return LOCO_OK;
}
}
