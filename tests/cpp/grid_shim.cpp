// C shim around include/loco_grid.hpp for tests/test_native_refine.py (ctypes): the native refineCell against the
// compiled reference's.  Test infrastructure only.
#include <cstring>
#include <string>

#include "loco_grid.hpp"

struct Shim {
	loco_grid::Grid grid;
	std::string err;
};

extern "C" {

void *gs_from_graph(const loco_graph_t *g, loco_value_fn fn, void *user)
{
	Shim *s = new Shim();
	loco_grid::Status st = loco_grid::Grid::fromGraph(*g, fn, user, &s->grid);
	if (!st.ok) s->err = st.msg;
	return s;
}
const char *gs_error(void *h) { return ((Shim *)h)->err.c_str(); }
void gs_free(void *h) { delete (Shim *)h; }
int gs_refine(void *h, int leaf, int dir)
{
	Shim *s = (Shim *)h;
	loco_grid::Status st = s->grid.refineLeaf(leaf, dir);
	if (!st.ok) s->err = st.msg;
	return st.ok ? 0 : 1;
}
int gs_refine_pass(void *h, const int32_t *leaf, const int32_t *dir, long n)
{
	Shim *s = (Shim *)h;
	loco_grid::Status st = s->grid.refineLeaves(leaf, dir, n);
	if (!st.ok) s->err = st.msg;
	return st.ok ? 0 : 1;
}
long gs_num_leaves(void *h) { return (long)((Shim *)h)->grid.nLeaves(); }
long gs_num_samples(void *h) { return (long)((Shim *)h)->grid.nSamples(); }
long gs_fn_calls(void *h) { return (long)((Shim *)h)->grid.nFunctionEvaluations(); }
const loco_graph_t *gs_graph(void *h) { return &((Shim *)h)->grid.graph(); }
// neighbour count of the leaf-th leaf: tree leaves / border cells (introspection)
void gs_leaf_neighbors(void *h, int leaf, int *n_leaf, int *n_border)
{
	loco_grid::Grid &g = ((Shim *)h)->grid;
	const int32_t c = g.leaves()[(size_t)leaf];
	*n_leaf = *n_border = 0;
	for (int32_t nb : g.neighbors(c)) (g.cells()[(size_t)nb].level < 0 ? *n_border : *n_leaf)++;
}
}
extern "C" {
void *gs_bootstrap(int N, int basis, int level, int D, int eval_corners, const double *domain, loco_value_fn fn, void *user)
{
	Shim *s = new Shim();
	loco_grid::Status st = loco_grid::Grid::bootstrap(N, basis, level, D, eval_corners != 0, domain, fn, user, &s->grid);
	if (!st.ok) s->err = st.msg;
	return s;
}
int gs_optimal_direction(void *h, int leaf)
{
	Shim *s = (Shim *)h;
	int dir = -1;
	loco_grid::Status st = s->grid.optimalSplitDirection(leaf, &dir);
	if (!st.ok) { s->err = st.msg; return -1; }
	return dir;
}
}
