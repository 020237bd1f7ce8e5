// oracle/loco_oracle.cpp -- CPU RESTATEMENT of the reference's evaluation path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library; the product (locointerpolate_b200/) never does and has no CPU path of its own.
//
// Parity status: PINNED.  tests/test_oracle.py checks this restatement against (a) the reference's own
// sources compiled in oracle/_ref/liblocoref.so on the reference tests' topologies (TestExamples.cpp:225-240,
// 457-469) -- leaf indices, support sets, neighbour counts and values -- and (b) the committed golden
// vectors in tests/golden/ that were generated from that library (tests/golden/make_golden.py).
// For mesh-valued functions (D > 1) the reference has no implementation (LCFunctionValue.cpp:20-31 rejects
// anything but LCRealFunctionValue): there the restatement applies LCRealFunctionValue::add component-wise
// and is pinned against D scalar evaluations of the reference on a shared tree.
//
// The structure deliberately follows the reference (pointer tree, per-sample basis lists, per-query
// assembly of the influencing set) rather than the product's flat model, so that the two are independent.
// Each function cites the reference lines it restates.  Deterministic where the reference is not: the
// reference iterates unordered_map<pointer,...> containers, here insertion order is used.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <set>
#include <string>
#include <vector>

#include "../include/loco_b200.h"  // only for the plain-array graph description loco_graph_t and the status codes

namespace {

const double EPSILON = 0.000001;              // LCMathHelper.cpp:84
const double CONTAINING_EPSILON = 0.0000001;  // LCAdaptiveGridCell.h:109

struct Basis {  // LCBasisFunction.h:43-47
	std::vector<double> center, support;
	double weight;
};

struct Sample {  // LCSample.h:60-64
	std::vector<double> center;
	std::vector<double> value;
	std::vector<Basis> basis;
};

struct Cell {  // LCAdaptiveGridCell.h:86-109
	std::vector<double> ranges;
	bool isLeaf = true;
	int level = 0;
	Cell *parent = nullptr, *child1 = nullptr, *child2 = nullptr;
	int splitDirection = -1;
	double splitVal = 0;
	std::list<Cell *> neighbors;
	std::vector<double> centerValue;                          // centerShapeInfo_
	std::vector<std::pair<int, std::vector<double>>> samples; // samples_: sample index -> per-cell copy
	int leafIndex = -1;
	bool hasSample(int s) const { for (auto &p : samples) if (p.first == s) return true; return false; }
	const std::vector<double> *valueOf(int s) const { for (auto &p : samples) if (p.first == s) return &p.second; return nullptr; }
};

struct Function {  // LCSampledFunction.h:29-35
	int N = 0, D = 1, basisType = 0;
	std::vector<double> ranges;
	std::vector<Sample> samples;
	Cell *root = nullptr;
	std::vector<Cell *> borderCells;
	std::vector<Cell *> leaves;  // getAllLeafCells order
	std::vector<Cell *> all;
	~Function() { for (Cell *c : all) delete c; }
};

// LCAdaptiveGridCell::checkIsNeighbor (LCAdaptiveGridCell.cpp:910-925)
bool checkIsNeighbor(const Cell *a, const Cell *b)
{
	int n = (int)a->ranges.size() / 2;
	for (int i = 0; i < n; i++) {
		double minA = a->ranges[2 * i], maxA = a->ranges[2 * i + 1];
		double minB = b->ranges[2 * i], maxB = b->ranges[2 * i + 1];
		if (((minB - maxA) > CONTAINING_EPSILON) || ((minA - maxB) > CONTAINING_EPSILON)) return false;
	}
	return true;
}

// LCAdaptiveGridCell::replaceNeighborWithChildren (:927-945)
void replaceNeighborWithChildren(Cell *self, Cell *neighbor, Cell *c1, Cell *c2)
{
	self->neighbors.remove(neighbor);
	if (checkIsNeighbor(self, c1)) self->neighbors.push_back(c1);
	if (checkIsNeighbor(self, c2)) self->neighbors.push_back(c2);
}

// LCAdaptiveGridCell::addAllNeighboursTopDown (:139-164)
void addAllNeighboursTopDown(Cell *c)
{
	if (c->parent) {
		std::list<Cell *> parentNeighbors = c->parent->neighbors;  // getNeighbors() returns a copy
		for (Cell *nb : parentNeighbors) if (checkIsNeighbor(c, nb)) c->neighbors.push_back(nb);
	}
	if (!c->isLeaf) {
		std::list<Cell *> mine = c->neighbors;  // range-for over the member while others' lists change
		for (Cell *nb : mine) replaceNeighborWithChildren(nb, c, c->child1, c->child2);
		c->child1->neighbors.push_back(c->child2);
		c->child2->neighbors.push_back(c->child1);
		addAllNeighboursTopDown(c->child1);
		addAllNeighboursTopDown(c->child2);
	}
}

// LCAdaptiveGridCell::evalPametersAreInCell (:438-454)
bool paramsInCell(const Cell *c, const double *p)
{
	int n = (int)c->ranges.size() / 2;
	for (int i = 0; i < n; i++)
		if ((p[i] < c->ranges[2 * i] - EPSILON) || (p[i] > c->ranges[2 * i + 1] + EPSILON)) return false;
	return true;
}

// LCAdaptiveGridCell::getAdjacentLeaves (:310-333); false = an LCError was returned (propagates immediately)
bool getAdjacentLeaves(Cell *c, const double *p, std::vector<Cell *> *result)
{
	if (c->isLeaf) {
		bool ok = paramsInCell(c, p);
		if (ok && std::find(result->begin(), result->end(), c) == result->end()) result->push_back(c);
		return ok;
	}
	if (p[c->splitDirection] <= c->splitVal + CONTAINING_EPSILON)
		if (!getAdjacentLeaves(c->child1, p, result)) return false;
	if (p[c->splitDirection] >= c->splitVal - CONTAINING_EPSILON)
		if (!getAdjacentLeaves(c->child2, p, result)) return false;
	return true;
}

// LCSampledFunction::addBorderCells (LCSampledFunction.cpp:24-47)
void addBorderCells(Function *f)
{
	for (Cell *border : f->borderCells) {
		std::vector<Cell *> neighbours;
		for (auto &s : border->samples) getAdjacentLeaves(f->root, f->samples[s.first].center.data(), &neighbours);
		for (Cell *nb : neighbours) {
			nb->neighbors.push_back(border);
			border->neighbors.push_back(nb);
		}
	}
}

// LCAdaptiveGridCell::getAllLeafCells (:247-258)
void getAllLeafCells(Cell *c, std::vector<Cell *> *out)
{
	if (c->isLeaf) out->push_back(c);
	else { getAllLeafCells(c->child1, out); getAllLeafCells(c->child2, out); }
}

// LCFunction::mapToStandardHypercube (LCFunction.cpp:59-73)
void mapToStandardHypercube(const Function *f, const double *params, double *result)
{
	for (int i = 0; i < f->N; i++) {
		double alpha = (params[i] - f->ranges[2 * i]) / (f->ranges[2 * i + 1] - f->ranges[2 * i]);
		result[i] = alpha * 2 - 1;
	}
}

// LCAdaptiveGridCell::getCointainingLeaf (:260-272)
bool getContainingLeaf(Cell *c, const double *p, Cell **leaf)
{
	if (c->isLeaf) { *leaf = c; return paramsInCell(c, p); }
	if (p[c->splitDirection] < c->splitVal) return getContainingLeaf(c->child1, p, leaf);
	return getContainingLeaf(c->child2, p, leaf);
}

// LCLinearBSpline::eval (LCBasisFunction.cpp:273-293)
double evalLinear(const Basis &b, const double *pos, int n)
{
	double value = 1;
	for (int i = 0; i < n; i++) {
		double dimVal = 0;
		double d = std::fabs(pos[i] - b.center[i]) / b.support[i];
		if (d <= 1) dimVal = 1 - d;
		value *= dimVal;
	}
	return value * b.weight;
}

// LCCubicBSpline::eval (LCBasisFunction.cpp:380-419)
double evalCubic(const Basis &b, const double *pos, int n)
{
	double value = 1;
	for (int i = 0; i < n; i++) {
		double dimVal = 0;
		double d = std::fabs(pos[i] - b.center[i]) / b.support[i];
		if (d <= 1) {
			double p = 2 * ((pos[i] - b.center[i]) / b.support[i] + 1);
			int k = (int)std::floor(p);
			double t = 1 - (p - k);
			if (k == 0) dimVal = (1 - t) * (1 - t) * (1 - t) / 6;
			if (k == 1) dimVal = (3 * t * t * t - 6 * t * t + 4) / 6;
			if (k == 2) dimVal = (-3 * t * t * t + 3 * t * t + 3 * t + 1) / 6;
			if (k == 3) dimVal = t * t * t / 6;
		}
		value *= dimVal;
	}
	return value * b.weight;
}

// LCSample::getInterpolationWeight (LCSample.cpp:76-89)
double interpolationWeight(const Function *f, const Sample &s, const double *pos)
{
	double combination = 0;
	for (const Basis &b : s.basis) combination += f->basisType == LOCO_CUBIC_BSPLINE ? evalCubic(b, pos, f->N) : evalLinear(b, pos, f->N);
	return combination;
}

// LCLinearBSpline::evalDeriv (LCBasisFunction.cpp:295-332) -- restated literally: d is the SIGNED distance
double evalDerivLinear(const Basis &b, const double *pos, int direction, int n)
{
	double value = 1;
	for (int i = 0; i < n; i++) {
		double dimVal = 0;
		double d = (pos[i] - b.center[i]) / b.support[i];
		if (i == direction) {
			if (d <= 1) {
				if (d < 0) dimVal = 1.0 / b.support[i];
				if (d > 0) dimVal = -1.0 / b.support[i];
			}
		} else {
			if (d <= 1) dimVal = 1 - d;
		}
		value *= dimVal;
	}
	return value * b.weight;
}

// LCCubicBSpline::evalDeriv (LCBasisFunction.cpp:421-483)
double evalDerivCubic(const Basis &b, const double *pos, int direction, int n)
{
	double value = 1;
	for (int i = 0; i < n; i++) {
		double dimVal = 0;
		double d = std::fabs(pos[i] - b.center[i]) / b.support[i];
		if (d <= 1) {
			double p = 2 * ((pos[i] - b.center[i]) / b.support[i] + 1);
			int k = (int)std::floor(p);
			double t = 1 - (p - k);
			if (i == direction) {
				double dt = -2 / b.support[i];
				if (k == 0) dimVal = -(1 - t) * (1 - t) * dt / 2;
				if (k == 1) dimVal = (3 * t * t - 4 * t) * dt / 2;
				if (k == 2) dimVal = (-3 * t * t + 2 * t + 1) * dt / 2;
				if (k == 3) dimVal = t * t * dt / 2;
			} else {
				if (k == 0) dimVal = (1 - t) * (1 - t) * (1 - t) / 6;
				if (k == 1) dimVal = (3 * t * t * t - 6 * t * t + 4) / 6;
				if (k == 2) dimVal = (-3 * t * t * t + 3 * t * t + 3 * t + 1) / 6;
				if (k == 3) dimVal = t * t * t / 6;
			}
		}
		value *= dimVal;
	}
	return value * b.weight;
}

// LCSample::getDerivInterpolationWeight (LCSample.cpp:92-105)
double derivInterpolationWeight(const Function *f, const Sample &s, const double *pos, int direction)
{
	double combination = 0;
	for (const Basis &b : s.basis)
		combination += f->basisType == LOCO_CUBIC_BSPLINE ? evalDerivCubic(b, pos, direction, f->N) : evalDerivLinear(b, pos, direction, f->N);
	return combination;
}

// LCAdaptiveGridCell::getAllInfluencingSamples (:558-607) with getNeightbourShapeInfoMappingMapping (:507-556;
// only its failure mode matters: the value map is the identity, LCRealFunctionValueMap.cpp:13-25)
int allInfluencingSamples(const Function *f, const Cell *cell, std::vector<std::pair<int, const std::vector<double> *>> *infl)
{
	std::set<int> processed;
	for (auto &s : cell->samples) { infl->push_back({s.first, &s.second}); processed.insert(s.first); }
	if (f->basisType == LOCO_CUBIC_BSPLINE) {
		for (Cell *nb : cell->neighbors) {
			for (auto &ns : nb->samples) {
				if (processed.count(ns.first)) continue;
				bool common = false;
				for (auto &mine : cell->samples) if (nb->hasSample(mine.first)) { common = true; break; }
				if (!common) return LOCO_ST_NO_COMMON_SAMPLE;  // "could not find closest sample"
				processed.insert(ns.first);
				infl->push_back({ns.first, &ns.second});
			}
		}
	}
	return LOCO_ST_OK;
}

// LCAdaptiveGridCell::evalShapeInfo (:611-676) + LCFunctionValue::newFromWeightedSum (LCFunctionValue.cpp:11-32)
// + LCRealFunctionValue::add (LCRealFunctionValue.cpp:51-61) applied to each of the D components.
// scale receives sum_k |w_k| * max_j |v_k[j]| (the magnitude parity tolerances are relative to).
int cellEval(const Function *f, const Cell *cell, const double *pos, double *result, double *scale)
{
	if (!paramsInCell(cell, pos)) return LOCO_ST_OUTSIDE_CELL;
	std::vector<std::pair<int, const std::vector<double> *>> infl;
	int st = allInfluencingSamples(f, cell, &infl);
	if (st != LOCO_ST_OK) return st;
	if (infl.empty()) return LOCO_ST_EMPTY_SUPPORT;
	for (int j = 0; j < f->D; j++) result[j] = 0;
	double sc = 0;
	for (auto &s : infl) {
		double w = interpolationWeight(f, f->samples[s.first], pos);
		double vmax = 0;
		for (int j = 0; j < f->D; j++) { result[j] += w * (*s.second)[j]; vmax = std::max(vmax, std::fabs((*s.second)[j])); }
		sc += std::fabs(w) * vmax;
	}
	if (scale) *scale = sc;
	return LOCO_ST_OK;
}

// LCAdaptiveGridCell::evalDeriv (:678-711) + LCFunctionDeriv::newFromWeightedSum (LCFunctionDeriv.cpp:12-35)
// + LCRealFunctionDeriv::add (LCRealFunctionDeriv.cpp:11-21); scalar functions only, like the reference
int cellEvalDeriv(const Function *f, const Cell *cell, const double *pos, int direction, double *result, double *scale)
{
	if (!paramsInCell(cell, pos)) return LOCO_ST_OUTSIDE_CELL;
	std::vector<std::pair<int, const std::vector<double> *>> infl;
	int st = allInfluencingSamples(f, cell, &infl);
	if (st != LOCO_ST_OK) return st;
	if (infl.empty()) return LOCO_ST_EMPTY_SUPPORT;
	double r = 0, sc = 0;
	for (auto &s : infl) {
		double w = derivInterpolationWeight(f, f->samples[s.first], pos, direction);
		r += w * (*s.second)[0];
		sc += std::fabs(w * (*s.second)[0]);
	}
	*result = r;
	if (scale) *scale = sc;
	return LOCO_ST_OK;
}

Cell *buildCells(Function *f, const loco_graph_t *g, int64_t *next, Cell *parent, int level)
{
	int64_t i = (*next)++;
	int N = f->N, D = f->D;
	Cell *c = new Cell();
	f->all.push_back(c);
	c->ranges.assign(g->cell_ranges + i * 2 * N, g->cell_ranges + (i + 1) * 2 * N);
	c->parent = parent;
	c->level = level;
	if (g->cell_split_dir[i] < 0) {
		c->isLeaf = true;
		if (g->cell_center_value) c->centerValue.assign(g->cell_center_value + i * D, g->cell_center_value + (i + 1) * D);
		for (int64_t e = g->cell_sample_offset[i]; e < g->cell_sample_offset[i + 1]; e++) {
			int s = g->cell_sample[e];
			std::vector<double> v = g->cell_sample_value ? std::vector<double>(g->cell_sample_value + e * D, g->cell_sample_value + (e + 1) * D)
			                                             : f->samples[s].value;
			bool dup = false;
			for (auto &p : c->samples) if (p.first == s) { p.second = v; dup = true; }
			if (!dup) c->samples.push_back({s, v});
		}
	} else {
		c->isLeaf = false;
		c->splitDirection = g->cell_split_dir[i];
		c->splitVal = g->cell_split_val[i];
		c->child1 = buildCells(f, g, next, c, level + 1);
		c->child2 = buildCells(f, g, next, c, level + 1);
	}
	return c;
}

} // namespace

extern "C" {

// convertToCplusplus (LCProtoConverter.cpp:356-403): samples, tree, addAllNeighboursTopDown, border cells,
// ranges, new LCSampledFunction, addBorderCells -- from a plain-array graph instead of a protobuf message.
void *oracle_new(const loco_graph_t *g)
{
	Function *f = new Function();
	f->N = g->n_params; f->D = g->value_dim; f->basisType = g->basis_type;
	int N = f->N, D = f->D;
	f->ranges.assign(g->domain, g->domain + 2 * N);
	f->samples.resize((size_t)g->n_samples);
	for (int64_t s = 0; s < g->n_samples; s++) {
		Sample &sm = f->samples[(size_t)s];
		sm.center.assign(g->sample_center + s * N, g->sample_center + (s + 1) * N);
		sm.value.assign(g->sample_value + s * D, g->sample_value + (s + 1) * D);
		for (int64_t b = g->bf_offset[s]; b < g->bf_offset[s + 1]; b++) {
			Basis bs;
			bs.center.assign(g->bf_center + b * N, g->bf_center + (b + 1) * N);
			bs.support.assign(g->bf_support + b * N, g->bf_support + (b + 1) * N);
			bs.weight = g->bf_weight[b];
			// LCSample::addBasisFunction (LCSample.cpp:133-147) merges functions with equal keys (EPS equality,
			// LCBasisFunction.h:87-91) by adding their weights
			bool merged = false;
			for (Basis &o : sm.basis) {
				double dc = 0, ds = 0;
				for (int i = 0; i < N; i++) { dc += (o.center[i] - bs.center[i]) * (o.center[i] - bs.center[i]); ds += (o.support[i] - bs.support[i]) * (o.support[i] - bs.support[i]); }
				if (std::sqrt(dc) < EPSILON && std::sqrt(ds) < EPSILON) { o.weight += bs.weight; merged = true; break; }
			}
			if (!merged) sm.basis.push_back(bs);
		}
	}
	int64_t next = 0;
	f->root = buildCells(f, g, &next, nullptr, 0);
	addAllNeighboursTopDown(f->root);
	for (int64_t b = 0; b < g->n_border_cells; b++) {
		Cell *c = new Cell();
		f->all.push_back(c);
		c->ranges.assign(g->border_ranges + b * 2 * N, g->border_ranges + (b + 1) * 2 * N);
		if (g->border_center_value) c->centerValue.assign(g->border_center_value + b * D, g->border_center_value + (b + 1) * D);
		for (int64_t e = g->border_sample_offset[b]; e < g->border_sample_offset[b + 1]; e++) {
			int s = g->border_sample[e];
			std::vector<double> v = g->border_sample_value ? std::vector<double>(g->border_sample_value + e * D, g->border_sample_value + (e + 1) * D)
			                                               : f->samples[s].value;
			if (!c->hasSample(s)) c->samples.push_back({s, v});
		}
		f->borderCells.push_back(c);
	}
	getAllLeafCells(f->root, &f->leaves);
	for (size_t i = 0; i < f->leaves.size(); i++) f->leaves[i]->leafIndex = (int)i;
	addBorderCells(f);
	return f;
}

void oracle_free(void *h) { delete (Function *)h; }
int oracle_num_leaves(void *h) { return (int)((Function *)h)->leaves.size(); }

// LCSampledFunction::evalShapeInfo (LCSampledFunction.cpp:69-78) for each query.
// out [Q*D] (NaN where status != 0), leaf [Q], status [Q], scale [Q] (may be NULL)
void oracle_eval(void *h, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, double *scale)
{
	Function *f = (Function *)h;
	std::vector<double> p(f->N);
	for (int64_t i = 0; i < Q; i++) {
		mapToStandardHypercube(f, q + i * f->N, p.data());
		Cell *cell = nullptr;
		bool inside = getContainingLeaf(f->root, p.data(), &cell);
		int st = inside ? cellEval(f, cell, p.data(), out + i * f->D, scale ? scale + i : nullptr) : LOCO_ST_OUTSIDE_CELL;
		if (st != LOCO_ST_OK) {
			for (int j = 0; j < f->D; j++) out[i * f->D + j] = NAN;
			if (scale) scale[i] = 0;
		}
		if (leaf) leaf[i] = cell->leafIndex;
		if (status) status[i] = (uint8_t)st;
	}
}

// LCSampledFunction::evalDeriv (LCSampledFunction.cpp:61-67) for each query, along cube direction `direction`
void oracle_eval_deriv(void *h, const double *q, int64_t Q, int direction, double *out, int32_t *leaf, uint8_t *status, double *scale)
{
	Function *f = (Function *)h;
	std::vector<double> p(f->N);
	for (int64_t i = 0; i < Q; i++) {
		mapToStandardHypercube(f, q + i * f->N, p.data());
		Cell *cell = nullptr;
		bool inside = getContainingLeaf(f->root, p.data(), &cell);
		int st = inside ? cellEvalDeriv(f, cell, p.data(), direction, out + i, scale ? scale + i : nullptr) : LOCO_ST_OUTSIDE_CELL;
		if (st != LOCO_ST_OK) { out[i] = NAN; if (scale) scale[i] = 0; }
		if (leaf) leaf[i] = cell->leafIndex;
		if (status) status[i] = (uint8_t)st;
	}
}

// sorted sample indices of getAllInfluencingSamples; returns K, -1 if cap too small, -2 on LCError
int oracle_support(void *h, int leafIdx, int32_t *ids, int cap)
{
	Function *f = (Function *)h;
	std::vector<std::pair<int, const std::vector<double> *>> infl;
	if (allInfluencingSamples(f, f->leaves[leafIdx], &infl) != LOCO_ST_OK) return -2;
	std::vector<int> v;
	for (auto &s : infl) v.push_back(s.first);
	std::sort(v.begin(), v.end());
	if ((int)v.size() > cap) return -1;
	std::copy(v.begin(), v.end(), ids);
	return (int)v.size();
}

void oracle_neighbors(void *h, int leafIdx, int *nLeaf, int *nBorder)
{
	Function *f = (Function *)h;
	int nl = 0, nb = 0;
	for (Cell *c : f->leaves[leafIdx]->neighbors) (c->leafIndex >= 0 ? nl : nb)++;
	*nLeaf = nl; *nBorder = nb;
}

// The data-parallel part of LCAdaptiveGrid::decideSplit (LCAdaptiveGrid.cpp:578-626):
// err = |approx - truth| (LCRealFunctionValue::computeErrorVector, LCRealFunctionValue.cpp:37-47),
// doSplit = doSplit || !((diff - allowedError).maxCoeff() <= 0), per containing leaf.
void oracle_error_check(void *h, const double *q, const double *truth, int64_t P, double threshold, double *errMax, uint8_t *split, int64_t *nBad)
{
	Function *f = (Function *)h;
	size_t L = f->leaves.size();
	for (size_t l = 0; l < L; l++) { errMax[l] = 0; split[l] = 0; }
	std::vector<double> out((size_t)f->D);
	int64_t bad = 0;
	for (int64_t i = 0; i < P; i++) {
		int32_t leaf; uint8_t st;
		oracle_eval(h, q + i * f->N, 1, out.data(), &leaf, &st, nullptr);
		if (st != LOCO_ST_OK) { bad++; continue; }
		for (int j = 0; j < f->D; j++) {
			double d = std::fabs(out[j] - truth[i * f->D + j]);
			if (!((d - threshold) <= 0)) split[leaf] = 1;
			if (!(d <= errMax[leaf]) && !std::isnan(errMax[leaf])) errMax[leaf] = d;  // a NaN error sticks (it forces the split)
		}
	}
	if (nBad) *nBad = bad;
}

// CENTER_BASED metric: LCAdaptiveGridCell::getApproximationError (:1088-1100) on every leaf
int oracle_center_error(void *h, double threshold, double *errMax, uint8_t *split)
{
	Function *f = (Function *)h;
	std::vector<double> center(f->N), out((size_t)f->D);
	for (size_t l = 0; l < f->leaves.size(); l++) {
		Cell *c = f->leaves[l];
		if ((int)c->centerValue.size() != f->D) return -1;
		for (int i = 0; i < f->N; i++) center[i] = 0.5 * (c->ranges[2 * i + 1] + c->ranges[2 * i]);  // getCenter :204-210
		errMax[l] = 0; split[l] = 0;
		if (cellEval(f, c, center.data(), out.data(), nullptr) != LOCO_ST_OK) { errMax[l] = NAN; split[l] = 1; continue; }
		for (int j = 0; j < f->D; j++) {
			double d = std::fabs(out[j] - c->centerValue[j]);
			if (!((d - threshold) <= 0)) split[l] = 1;
			if (!(d <= errMax[l]) && !std::isnan(errMax[l])) errMax[l] = d;
		}
	}
	return 0;
}

// evaluate on a known cell at CUBE coordinates (what decideSplit does with its jitter points, :611)
void oracle_eval_in_cell(void *h, const double *x, const int32_t *leafIdx, int64_t Q, double *out, uint8_t *status)
{
	Function *f = (Function *)h;
	for (int64_t i = 0; i < Q; i++) {
		int st = cellEval(f, f->leaves[leafIdx[i]], x + i * f->N, out + i * f->D, nullptr);
		if (st != LOCO_ST_OK) for (int j = 0; j < f->D; j++) out[i * f->D + j] = NAN;
		if (status) status[i] = (uint8_t)st;
	}
}

} // extern "C"
