// oracle/ref_adapter.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" adapter around the UNMODIFIED reference sources, which oracle/Makefile
// compiles where they lie under /root/reference into oracle/_ref/liblocoref.so.  It lets the
// tests and bench.py's CPU-baseline leg
//   * build grids with the reference's own builder  (LCAdaptiveGrid.cpp:16-39, 256-456, 629-914),
//   * wrap them in the reference's LCSampledFunction  (LCSampledFunction.cpp:9-18),
//   * optionally rebuild them the way the protobuf loader does, without protobuf
//     (LCProtoConverter.cpp:356-403: same constructors, addAllNeighboursTopDown, addBorderCells),
//   * dump the object graph as plain arrays in the canonical numbering of SURVEY.md 8(a12)
//     (cells in pre-order with child1 first, leaf index = rank in getAllLeafCells order,
//     sample index = position in samples_),
//   * evaluate query batches through LCSampledFunction::evalShapeInfo (LCSampledFunction.cpp:69-78)
//     and report the containing leaf, the status and the value,
//   * return a leaf's influencing-sample set (LCAdaptiveGridCell.cpp:558-607) as sorted indices.
// Nothing of the reference is copied: this file only CALLS its public API.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <streambuf>
#include <unordered_map>
#include <vector>
#include <algorithm>
#include <sys/wait.h>
#include <unistd.h>

#include "LCError.h"
#include "LCMathHelper.h"
#include "LCFunction.h"
#include "LCFunctionValue.h"
#include "LCRealFunction.h"
#include "LCRealFunctionValue.h"
#include "LCFunctionDeriv.h"
#include "LCRealFunctionDeriv.h"
#include "LCAdaptiveSamplingParams.h"
#include "LCBasisFunction.h"
#include "LCSample.h"
#include "LCAdaptiveGridCell.h"
#include "LCAdaptiveGrid.h"
#include "LCSampledFunction.h"

#include "../include/loco_b200.h"  // only the plain-array graph description loco_graph_t

namespace {

// The reference prints diagnostics on std::cout from inside the hot path (e.g. the
// "parameter ouside cell" text, LCAdaptiveGridCell.cpp:618-622).  Swallow them while we drive it.
struct CoutMute {
	std::streambuf *old;
	std::ostringstream sink;
	CoutMute() { old = std::cout.rdbuf(sink.rdbuf()); }
	~CoutMute() { std::cout.rdbuf(old); }
};

struct RefGrid {
	LCRealFunction *f;
	LCAdaptiveSamplingParams *params;
	LCAdaptiveGrid *grid;
	int basis;
};

struct RefFn {
	LCSampledFunction *fn;
	std::vector<LCAdaptiveGridCell *> cells;   // pre-order, child1 first
	std::vector<LCAdaptiveGridCell *> leaves;  // getAllLeafCells order
	std::unordered_map<LCAdaptiveGridCell *, int> leafIndex;
	std::unordered_map<LCSample *, int> sampleIndex;
	std::vector<LCSample *> samples;
	std::vector<LCAdaptiveGridCell *> border;
	int basis;
};

void preorder(LCAdaptiveGridCell *c, std::vector<LCAdaptiveGridCell *> *out)
{
	out->push_back(c);
	if (!c->isLeaf()) {
		preorder(c->getChild1(), out);
		preorder(c->getChild2(), out);
	}
}

RefFn *wrap(LCSampledFunction *fn)
{
	RefFn *r = new RefFn();
	r->fn = fn;
	preorder(fn->getRoot(), &r->cells);
	fn->getRoot()->getAllLeafCells(&r->leaves);
	for (size_t i = 0; i < r->leaves.size(); i++) r->leafIndex[r->leaves[i]] = (int)i;
	r->samples = fn->getSamples();
	for (size_t i = 0; i < r->samples.size(); i++) r->sampleIndex[r->samples[i]] = (int)i;
	r->border = fn->getBorderCells();
	r->basis = (int)(*r->samples[0]->getBasisFunctions().begin()->second).getType();
	return r;
}

double realVal(LCFunctionValue *v)
{
	LCRealFunctionValue *r = dynamic_cast<LCRealFunctionValue *>(v);
	return r ? r->val : NAN;
}

// The descent of LCAdaptiveGridCell::getCointainingLeaf (LCAdaptiveGridCell.cpp:260-272) is a
// member that also runs the containment test; calling it is the faithful way to get the leaf.
int locate(RefFn *r, const std::vector<double> &cube, int *status)
{
	LCAdaptiveGridCell *cell = nullptr;
	LCError err = r->fn->getRoot()->getCointainingLeaf(cube, &cell);
	*status = err.isOK() ? 0 : 1;
	auto it = r->leafIndex.find(cell);
	return it == r->leafIndex.end() ? -1 : it->second;
}

} // namespace

extern "C" {

// ---- grid construction ---------------------------------------------------------------------
// basis: 0 linear, 1 cubic (LCBasisFunction.h:19).  linearFn: LCRealFunction(N, linear).
void *ref_grid_new(int N, int basis, int linearFn, int maxDepth, double threshold, int bootstrap,
                   int nErrorSamples, int evalCorners)
{
	CoutMute mute;
	RefGrid *g = new RefGrid();
	g->f = new LCRealFunction(N, linearFn != 0);
	g->params = new LCAdaptiveSamplingParams(maxDepth, threshold, bootstrap, nErrorSamples);
	g->basis = basis;
	g->grid = new LCAdaptiveGrid(g->f, (LCBasisFunction::LCBasisFunctionType)basis, g->params, evalCorners != 0);
	return g;
}

int ref_grid_num_leaves(void *gp)
{
	RefGrid *g = (RefGrid *)gp;
	std::vector<LCAdaptiveGridCell *> leafs;
	g->grid->getRoot()->getAllLeafCells(&leafs);
	return (int)leafs.size();
}

// The random refinement used by the reference tests (TestExamples.cpp:225-240, 325-340, 393-408):
// srand(seed); n times { id = rand()%leaves; dir = rand()%N; if level < maxLevel refineCell }.
int ref_grid_refine_random(void *gp, unsigned seed, int n, int maxLevel)
{
	CoutMute mute;
	RefGrid *g = (RefGrid *)gp;
	int N = g->f->getNParams();
	srand(seed);
	int refined = 0;
	std::vector<LCAdaptiveGridCell *> affected;
	for (int i = 0; i < n; i++) {
		std::vector<LCAdaptiveGridCell *> leafs;
		g->grid->getRoot()->getAllLeafCells(&leafs);
		int id = rand() % leafs.size();
		int dir = rand() % N;
		if (leafs[id]->getLevel() < maxLevel) {
			LCError err = g->grid->refineCell(leafs[id], dir, &affected);
			if (!err.isOK()) return -1 - refined;
			refined++;
		}
	}
	return refined;
}

int ref_grid_refine(void *gp, int leaf, int dir)
{
	CoutMute mute;
	RefGrid *g = (RefGrid *)gp;
	std::vector<LCAdaptiveGridCell *> leafs, affected;
	g->grid->getRoot()->getAllLeafCells(&leafs);
	if (leaf < 0 || leaf >= (int)leafs.size()) return -1;
	return g->grid->refineCell(leafs[leaf], dir, &affected).isOK() ? 0 : -2;
}

// LCAdaptiveGridCell::getOptimalSplitDirection (LCAdaptiveGridCell.cpp:977-1041; the ERROR_BASED split policy,
// LCAdaptiveGrid::decideSplitDirection :556-557) for the leaf-th leaf; -1 if the reference reports an error.
int ref_grid_optimal_direction(void *gp, int leaf)
{
	CoutMute mute;
	RefGrid *g = (RefGrid *)gp;
	std::vector<LCAdaptiveGridCell *> leafs;
	g->grid->getRoot()->getAllLeafCells(&leafs);
	if (leaf < 0 || leaf >= (int)leafs.size()) return -1;
	int dir = -1;
	return leafs[leaf]->getOptimalSplitDirection(&dir).isOK() ? dir : -1;
}

// True function the grid samples, in ORIGINAL parameter space (LCRealFunction.cpp:50-77).
void ref_grid_true_values(void *gp, const double *q, long Q, double *out)
{
	RefGrid *g = (RefGrid *)gp;
	int N = g->f->getNParams();
	for (long i = 0; i < Q; i++) {
		std::vector<double> p(q + i * N, q + (i + 1) * N);
		LCFunctionValue *v = nullptr;
		g->f->evalShapeInfo(p, &v);
		out[i] = realVal(v);
		delete v;
	}
}

// LCSampledFunction over the grid, exactly as the reference tests do (TestExamples.cpp:193, 244).
void *ref_fn_from_grid(void *gp)
{
	RefGrid *g = (RefGrid *)gp;
	std::vector<double> ranges;
	g->f->getRanges(&ranges);
	return wrap(new LCSampledFunction(g->grid->getRoot(), ranges, g->grid->getSamples(), g->grid->getBorderCells()));
}

// ---- reload emulation -------------------------------------------------------------------------
// Rebuilds the function from the data a proto file would carry, through the same constructors and
// in the same order as convertToCplusplus (LCProtoConverter.cpp:356-403); protobuf itself cannot be
// linked in this image (SURVEY.md 8c caveat 2).
static LCAdaptiveGridCell *reloadCell(LCAdaptiveGridCell *src, LCAdaptiveGridCell *parent, int level,
                                      std::unordered_map<LCSample *, LCSample *> &map)
{
	std::vector<double> ranges = src->ranges_;
	if (src->isLeaf()) {
		std::unordered_map<LCSample *, LCFunctionValue *> samples;
		for (auto s : src->getPrecomputedSamples())
			samples[map[s.first]] = new LCRealFunctionValue(realVal(s.second));
		LCFunctionValue *center = new LCRealFunctionValue(realVal(src->getCenterShapeInfo()));
		return new LCAdaptiveGridCell(level, parent, ranges, center, samples);
	}
	LCAdaptiveGridCell *cell = new LCAdaptiveGridCell(level, parent, ranges, src->getSplitDirection(), src->getSplitValue());
	LCAdaptiveGridCell *c1 = reloadCell(src->getChild1(), cell, level + 1, map);
	LCAdaptiveGridCell *c2 = reloadCell(src->getChild2(), cell, level + 1, map);
	cell->setChildren(c1, c2);
	return cell;
}

void *ref_fn_reload(void *fp)
{
	CoutMute mute;
	RefFn *r = (RefFn *)fp;
	std::unordered_map<LCSample *, LCSample *> map;
	std::vector<LCSample *> samples;
	for (LCSample *s : r->samples) {
		std::vector<LCBasisFunction *> bfs;
		for (auto kv : s->getBasisFunctions()) {
			LCBasisFunction *bf = kv.second;
			if (bf->getType() == LCBasisFunction::LINEAR_BSPLINE)
				bfs.push_back(new LCLinearBSpline(bf->getCenter(), bf->getSupport(), bf->getWeight()));
			else
				bfs.push_back(new LCCubicBSpline(bf->getCenter(), bf->getSupport(), bf->getWeight()));
		}
		LCSample *ns = new LCSample(s->getCenter(), bfs, new LCRealFunctionValue(realVal(s->getShapeInfo())));
		map[s] = ns;
		samples.push_back(ns);
	}
	LCAdaptiveGridCell *root = reloadCell(r->fn->getRoot(), nullptr, 0, map);
	root->addAllNeighboursTopDown();
	std::vector<LCAdaptiveGridCell *> border;
	for (LCAdaptiveGridCell *b : r->border) border.push_back(reloadCell(b, nullptr, 0, map));
	std::vector<double> ranges;
	r->fn->getRanges(&ranges);
	LCSampledFunction *fn = new LCSampledFunction(root, ranges, samples, border);
	fn->addBorderCells();
	return wrap(fn);
}

// Build REFERENCE objects from a plain-array graph through the loader's code path (the same constructor
// sequence as convertToCplusplus, LCProtoConverter.cpp:356-403).  This is how bench.py hands the
// benchmark model to the reference for the CPU baseline without paying the reference's O(n^2) builder
// (87 s for the 43k-sample config).  Scalar values only (the reference has no other kind).
static LCAdaptiveGridCell *cellFromGraph(const loco_graph_t *g, int64_t *next, LCAdaptiveGridCell *parent, int level,
                                         std::vector<LCSample *> &samples)
{
	int64_t i = (*next)++;
	int N = g->n_params;
	std::vector<double> ranges(g->cell_ranges + i * 2 * N, g->cell_ranges + (i + 1) * 2 * N);
	if (g->cell_split_dir[i] < 0) {
		std::unordered_map<LCSample *, LCFunctionValue *> leafSamples;
		for (int64_t e = g->cell_sample_offset[i]; e < g->cell_sample_offset[i + 1]; e++) {
			int s = g->cell_sample[e];
			leafSamples[samples[s]] = new LCRealFunctionValue(g->cell_sample_value ? g->cell_sample_value[e] : g->sample_value[s]);
		}
		LCFunctionValue *center = new LCRealFunctionValue(g->cell_center_value ? g->cell_center_value[i] : 0.0);
		return new LCAdaptiveGridCell(level, parent, ranges, center, leafSamples);
	}
	LCAdaptiveGridCell *cell = new LCAdaptiveGridCell(level, parent, ranges, g->cell_split_dir[i], g->cell_split_val[i]);
	LCAdaptiveGridCell *c1 = cellFromGraph(g, next, cell, level + 1, samples);
	LCAdaptiveGridCell *c2 = cellFromGraph(g, next, cell, level + 1, samples);
	cell->setChildren(c1, c2);
	return cell;
}

void *ref_fn_from_graph(const loco_graph_t *g)
{
	CoutMute mute;
	if (g->value_dim != 1) return nullptr;
	int N = g->n_params;
	std::vector<LCSample *> samples;
	for (int64_t s = 0; s < g->n_samples; s++) {
		Eigen::VectorXd center(N);
		for (int d = 0; d < N; d++) center(d) = g->sample_center[s * N + d];
		std::vector<LCBasisFunction *> bfs;
		for (int64_t b = g->bf_offset[s]; b < g->bf_offset[s + 1]; b++) {
			Eigen::VectorXd c(N), sp(N);
			for (int d = 0; d < N; d++) { c(d) = g->bf_center[b * N + d]; sp(d) = g->bf_support[b * N + d]; }
			if (g->basis_type == 0) bfs.push_back(new LCLinearBSpline(c, sp, g->bf_weight[b]));
			else bfs.push_back(new LCCubicBSpline(c, sp, g->bf_weight[b]));
		}
		samples.push_back(new LCSample(center, bfs, new LCRealFunctionValue(g->sample_value[s])));
	}
	int64_t next = 0;
	LCAdaptiveGridCell *root = cellFromGraph(g, &next, nullptr, 0, samples);
	root->addAllNeighboursTopDown();
	std::vector<LCAdaptiveGridCell *> border;
	for (int64_t b = 0; b < g->n_border_cells; b++) {
		std::vector<double> ranges(g->border_ranges + b * 2 * N, g->border_ranges + (b + 1) * 2 * N);
		std::unordered_map<LCSample *, LCFunctionValue *> cellSamples;
		for (int64_t e = g->border_sample_offset[b]; e < g->border_sample_offset[b + 1]; e++) {
			int s = g->border_sample[e];
			cellSamples[samples[s]] = new LCRealFunctionValue(g->border_sample_value ? g->border_sample_value[e] : g->sample_value[s]);
		}
		LCFunctionValue *center = new LCRealFunctionValue(g->border_center_value ? g->border_center_value[b] : 0.0);
		border.push_back(new LCAdaptiveGridCell(0, nullptr, ranges, center, cellSamples));
	}
	std::vector<double> ranges(g->domain, g->domain + 2 * N);
	LCSampledFunction *fn = new LCSampledFunction(root, ranges, samples, border);
	fn->addBorderCells();
	return wrap(fn);
}

// ---- graph dump -----------------------------------------------------------------------------------
// sizes[0]=N [1]=#samples [2]=#cells(pre-order) [3]=#leaves [4]=#border cells [5]=#basis functions
// [6]=#(leaf,sample) entries in the tree [7]=#(border cell,sample) entries [8]=basis type
void ref_fn_sizes(void *fp, long *sizes)
{
	RefFn *r = (RefFn *)fp;
	sizes[0] = r->fn->getNParams();
	sizes[1] = (long)r->samples.size();
	sizes[2] = (long)r->cells.size();
	sizes[3] = (long)r->leaves.size();
	sizes[4] = (long)r->border.size();
	long nb = 0;
	for (LCSample *s : r->samples) nb += (long)s->getBasisFunctions().size();
	sizes[5] = nb;
	long nl = 0;
	for (LCAdaptiveGridCell *c : r->cells) if (c->isLeaf()) nl += (long)c->getPrecomputedSamples().size();
	sizes[6] = nl;
	long nbs = 0;
	for (LCAdaptiveGridCell *c : r->border) nbs += (long)c->getPrecomputedSamples().size();
	sizes[7] = nbs;
	sizes[8] = r->basis;
}

void ref_fn_ranges(void *fp, double *ranges)
{
	RefFn *r = (RefFn *)fp;
	std::vector<double> v;
	r->fn->getRanges(&v);
	for (size_t i = 0; i < v.size(); i++) ranges[i] = v[i];
}

// centers [S*N], values [S], valueGroup [S] (samples sharing one LCFunctionValue object get the same
// group id = lowest sample index holding it; LCAdaptiveGrid.cpp:345-355), bfOff [S+1],
// bfCenter/bfSupport [B*N], bfWeight [B], bfType [B]
void ref_fn_dump_samples(void *fp, double *centers, double *values, int *valueGroup, int *bfOff,
                         double *bfCenter, double *bfSupport, double *bfWeight, int *bfType)
{
	RefFn *r = (RefFn *)fp;
	int N = r->fn->getNParams();
	std::unordered_map<LCFunctionValue *, int> group;
	long b = 0;
	for (size_t i = 0; i < r->samples.size(); i++) {
		LCSample *s = r->samples[i];
		for (int d = 0; d < N; d++) centers[i * N + d] = s->getCenter()(d);
		values[i] = realVal(s->getShapeInfo());
		auto g = group.find(s->getShapeInfo());
		if (g == group.end()) { group[s->getShapeInfo()] = (int)i; valueGroup[i] = (int)i; }
		else valueGroup[i] = g->second;
		bfOff[i] = (int)b;
		for (auto kv : s->getBasisFunctions()) {
			LCBasisFunction *bf = kv.second;
			for (int d = 0; d < N; d++) {
				bfCenter[b * N + d] = bf->getCenter()(d);
				bfSupport[b * N + d] = bf->getSupport()(d);
			}
			bfWeight[b] = bf->getWeight();
			bfType[b] = (int)bf->getType();
			b++;
		}
	}
	bfOff[r->samples.size()] = (int)b;
}

static void dumpCellList(RefFn *r, const std::vector<LCAdaptiveGridCell *> &cells, double *ranges, int *isLeaf,
                         int *splitDir, double *splitVal, double *centerVal, int *lsOff, int *lsSample, double *lsValue)
{
	int N = r->fn->getNParams();
	long e = 0;
	for (size_t i = 0; i < cells.size(); i++) {
		LCAdaptiveGridCell *c = cells[i];
		for (int d = 0; d < 2 * N; d++) ranges[i * 2 * N + d] = c->ranges_[d];
		isLeaf[i] = c->isLeaf() ? 1 : 0;
		splitDir[i] = c->isLeaf() ? -1 : c->getSplitDirection();
		splitVal[i] = c->isLeaf() ? 0.0 : c->getSplitValue();
		centerVal[i] = c->isLeaf() ? realVal(c->getCenterShapeInfo()) : 0.0;
		lsOff[i] = (int)e;
		if (c->isLeaf()) {
			for (auto s : c->getPrecomputedSamples()) {
				lsSample[e] = r->sampleIndex[s.first];
				lsValue[e] = realVal(s.second);
				e++;
			}
		}
	}
	lsOff[cells.size()] = (int)e;
}

// tree cells in pre-order: ranges [C*2N], isLeaf/splitDir [C], splitVal/centerVal [C],
// lsOff [C+1] into lsSample/lsValue = the leaf's samples_ map (sample index, per-cell value copy)
void ref_fn_dump_cells(void *fp, double *ranges, int *isLeaf, int *splitDir, double *splitVal, double *centerVal,
                       int *lsOff, int *lsSample, double *lsValue)
{
	RefFn *r = (RefFn *)fp;
	dumpCellList(r, r->cells, ranges, isLeaf, splitDir, splitVal, centerVal, lsOff, lsSample, lsValue);
}

void ref_fn_dump_border(void *fp, double *ranges, double *centerVal, int *lsOff, int *lsSample, double *lsValue)
{
	RefFn *r = (RefFn *)fp;
	std::vector<int> isLeaf(r->border.size() + 1), splitDir(r->border.size() + 1);
	std::vector<double> splitVal(r->border.size() + 1);
	dumpCellList(r, r->border, ranges, isLeaf.data(), splitDir.data(), splitVal.data(), centerVal, lsOff, lsSample, lsValue);
}

// ---- evaluation -------------------------------------------------------------------------------------
// q [Q*N] in ORIGINAL parameter space.  out [Q] (NaN where the reference returned an error),
// leaf [Q] (may be NULL), status [Q] (0 ok, 1 "parameter ouside cell"; may be NULL).
void ref_fn_eval(void *fp, const double *q, long Q, double *out, int *leaf, unsigned char *status)
{
	CoutMute mute;
	RefFn *r = (RefFn *)fp;
	int N = r->fn->getNParams();
	std::vector<double> p(N);
	for (long i = 0; i < Q; i++) {
		for (int d = 0; d < N; d++) p[d] = q[i * N + d];
		if (leaf || status) {
			int st;
			int l = locate(r, r->fn->mapToStandardHypercube(p), &st);
			if (leaf) leaf[i] = l;
			if (status) status[i] = (unsigned char)st;
		}
		LCFunctionValue *v = nullptr;
		LCError err = r->fn->evalShapeInfo(p, &v);
		if (err.isOK()) { out[i] = realVal(v); delete v; }
		else out[i] = NAN;
	}
}

// LCSampledFunction::evalDeriv (LCSampledFunction.cpp:61-67) along cube direction `direction`.
// out [Q] (NaN where the reference returned an error)
void ref_fn_eval_deriv(void *fp, const double *q, long Q, int direction, double *out)
{
	CoutMute mute;
	RefFn *r = (RefFn *)fp;
	int N = r->fn->getNParams();
	std::vector<double> p(N);
	for (long i = 0; i < Q; i++) {
		for (int d = 0; d < N; d++) p[d] = q[i * N + d];
		LCFunctionDeriv *v = nullptr;
		LCError err = r->fn->evalDeriv(p, direction, &v);
		LCRealFunctionDeriv *rv = err.isOK() ? dynamic_cast<LCRealFunctionDeriv *>(v) : nullptr;
		out[i] = rv ? rv->val : NAN;
		delete rv;
	}
}

// Timing helper for the CPU baseline: evaluates only (no leaf/status bookkeeping), returns checksum.
double ref_fn_eval_sum(void *fp, const double *q, long Q)
{
	CoutMute mute;
	RefFn *r = (RefFn *)fp;
	int N = r->fn->getNParams();
	std::vector<double> p(N);
	double acc = 0;
	for (long i = 0; i < Q; i++) {
		for (int d = 0; d < N; d++) p[d] = q[i * N + d];
		LCFunctionValue *v = nullptr;
		LCError err = r->fn->evalShapeInfo(p, &v);
		if (err.isOK()) { acc += realVal(v); delete v; }
	}
	return acc;
}

// Fork-parallel evaluation over `procs` processes (the reference is single-threaded and its cubic
// path mutates LCAdaptiveGridCell::storedMaps, LCAdaptiveGridCell.cpp:541-551, so threads are unsafe;
// children share the built tree copy-on-write).  Results are discarded; returns 0 on success.
int ref_fn_eval_fork(void *fp, const double *q, long Q, int procs)
{
	RefFn *r = (RefFn *)fp;
	int N = r->fn->getNParams();
	std::vector<pid_t> pids;
	for (int k = 0; k < procs; k++) {
		long lo = Q * k / procs, hi = Q * (k + 1) / procs;
		pid_t pid = fork();
		if (pid == 0) {
			volatile double sink = ref_fn_eval_sum(fp, q + lo * N, hi - lo);
			(void)sink;
			_exit(0);
		}
		if (pid < 0) return -1;
		pids.push_back(pid);
	}
	int rc = 0;
	for (pid_t pid : pids) { int st = 0; waitpid(pid, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = -2; }
	return rc;
}

// Weights of the influencing samples at a point given in CUBE coordinates, for the mesh-valued
// restatement: ids/w in the reference's own list order; returns K or <0.
int ref_fn_weights(void *fp, int leafIdx, const double *cube, int *ids, double *w, int cap)
{
	CoutMute mute;
	RefFn *r = (RefFn *)fp;
	int N = r->fn->getNParams();
	LCAdaptiveGridCell *cell = r->leaves[leafIdx];
	std::vector<std::pair<LCSample *, LCFunctionValue *>> infl;
	std::vector<LCFunctionValue *> del;
	if (!cell->getAllInfluencingSamples((LCBasisFunction::LCBasisFunctionType)r->basis, &infl, &del).isOK()) return -1;
	Eigen::VectorXd pos(N);
	for (int d = 0; d < N; d++) pos(d) = cube[d];
	int k = 0;
	for (auto s : infl) {
		if (k >= cap) { k = -2; break; }
		double wt = 0;
		s.first->getInterpolationWeight(pos, &wt);
		ids[k] = r->sampleIndex[s.first];
		w[k] = wt;
		k++;
	}
	for (auto v : del) delete v;
	return k;
}

// Mesh-valued CPU baseline: the reference has no mesh-valued LCFunctionValue (LCFunctionValue.cpp:20-31), so a
// query costs the reference's own locate + influencing-set assembly + weights (getAllInfluencingSamples,
// getInterpolationWeight) followed by LCRealFunctionValue::add's `val += weight * other` applied to each of the D
// components of a row-major table [S x D].  Returns a checksum of the results.
double ref_fn_eval_mesh_sum(void *fp, const double *q, long Q, const double *table, long D)
{
	CoutMute mute;
	RefFn *r = (RefFn *)fp;
	int N = r->fn->getNParams();
	std::vector<double> p(N), res((size_t)D);
	double acc = 0;
	for (long i = 0; i < Q; i++) {
		for (int d = 0; d < N; d++) p[d] = q[i * N + d];
		std::vector<double> cube = r->fn->mapToStandardHypercube(p);
		LCAdaptiveGridCell *cell = nullptr;
		if (!r->fn->getRoot()->getCointainingLeaf(cube, &cell).isOK()) continue;
		std::vector<std::pair<LCSample *, LCFunctionValue *>> infl;
		std::vector<LCFunctionValue *> del;
		if (!cell->getAllInfluencingSamples((LCBasisFunction::LCBasisFunctionType)r->basis, &infl, &del).isOK()) continue;
		Eigen::VectorXd pos(N);
		for (int d = 0; d < N; d++) pos(d) = cube[d];
		std::fill(res.begin(), res.end(), 0.0);
		for (auto s : infl) {
			double wt = 0;
			s.first->getInterpolationWeight(pos, &wt);
			const double *row = table + (long)r->sampleIndex[s.first] * D;
			for (long j = 0; j < D; j++) res[(size_t)j] += wt * row[j];
		}
		for (auto v : del) delete v;
		acc += res[0] + res[(size_t)D - 1];
	}
	return acc;
}

int ref_fn_eval_mesh_fork(void *fp, const double *q, long Q, const double *table, long D, int procs)
{
	RefFn *r = (RefFn *)fp;
	int N = r->fn->getNParams();
	std::vector<pid_t> pids;
	for (int k = 0; k < procs; k++) {
		long lo = Q * k / procs, hi = Q * (k + 1) / procs;
		pid_t pid = fork();
		if (pid == 0) {
			volatile double sink = ref_fn_eval_mesh_sum(fp, q + lo * N, hi - lo, table, D);
			(void)sink;
			_exit(0);
		}
		if (pid < 0) return -1;
		pids.push_back(pid);
	}
	int rc = 0;
	for (pid_t pid : pids) { int st = 0; waitpid(pid, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = -2; }
	return rc;
}

// Sorted sample indices of getAllInfluencingSamples for a leaf; returns K (or -1 if cap too small).
int ref_fn_support(void *fp, int leafIdx, int *ids, int cap)
{
	CoutMute mute;
	RefFn *r = (RefFn *)fp;
	LCAdaptiveGridCell *cell = r->leaves[leafIdx];
	std::vector<std::pair<LCSample *, LCFunctionValue *>> infl;
	std::vector<LCFunctionValue *> del;
	if (!cell->getAllInfluencingSamples((LCBasisFunction::LCBasisFunctionType)r->basis, &infl, &del).isOK()) return -3;
	std::vector<int> v;
	for (auto s : infl) v.push_back(r->sampleIndex[s.first]);
	for (auto d : del) delete d;
	std::sort(v.begin(), v.end());
	if ((int)v.size() > cap) return -1;
	for (size_t i = 0; i < v.size(); i++) ids[i] = v[i];
	return (int)v.size();
}

int ref_fn_num_neighbors(void *fp, int leafIdx)
{
	RefFn *r = (RefFn *)fp;
	return (int)r->leaves[leafIdx]->getNeighbors().size();
}

// The refinement-loop error check for one leaf, CENTER_BASED, with the calls decideSplit makes (LCAdaptiveGrid.cpp:586-596):
// |interp(centre) - f(centre)|.  (LCAdaptiveGridCell::getApproximationError does the same but takes the basis type from
// the cell's own samples, LCAdaptiveGridCell.cpp:1105-1118 — uninitialised, and a crash, for a leaf of a deeply refined tree
// whose own samples have lost all their basis functions to the locality repair; decideSplit uses the grid's basisType_.)
double ref_fn_center_error(void *fp, int leafIdx)
{
	CoutMute mute;
	RefFn *r = (RefFn *)fp;
	LCAdaptiveGridCell *cell = r->leaves[leafIdx];
	std::vector<double> center;
	cell->getCenter(&center);
	LCFunctionValue *approx = nullptr;
	if (!cell->evalShapeInfo(center, (LCBasisFunction::LCBasisFunctionType)r->basis, &approx).isOK()) return NAN;
	Eigen::VectorXd e;
	const bool ok = approx->computeErrorVector(cell->getCenterShapeInfo(), &e).isOK();
	delete approx;
	return ok ? e(0) : NAN;
}

} // extern "C"
