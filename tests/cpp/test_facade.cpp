// tests/cpp/test_facade.cpp -- the reference's own test scenarios (TestExamples/TestExamples.cpp) replayed against
// the C++ facade (include/loco_lc.hpp) that stands in for the reference's class surface on the GPU path.
//
//   protoTest                        TestExamples.cpp:214-310   load a proto, 11x11 sweep, write, reload, sweep again, compare
//   linearPrecisionTestLinear/Cubic  TestExamples.cpp:313-448   the interpolant of f = 1 + sum (i+1) x_i reproduces it (bound 1e-5 there)
//   testFunctionLinearApproximation  TestExamples.cpp:185-211   sine-sum f, grid params (3,0.1,2,0) / (3,0.5,1,0), 21x21 sweep
//   decideSplit                      LCAdaptiveGrid.cpp:578-626 CENTER_BASED and RANDOM_SAMPLES, batched over all leaves
//   LCError behaviour                LCAdaptiveGridCell.cpp:438-454
//   adaptiveSample                   LCAdaptiveGrid.cpp:481-545 the refinement loop, level-synchronous: error check on the GPU,
//                                    refineCell (:629-914) by the host grid of include/loco_grid.hpp, incremental re-flatten
//
// usage:  test_facade <golden .pb> <dump file> [--host-only]
// Unlike the reference's mains (which print and always exit 0) every check counts: exit status = failed checks.
// The sweep values are also written to <dump file> so that tests/test_facade.py can compare them with the oracle.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "loco_lc.hpp"

static int g_failed = 0;
#define CHECK(cond, what)                                                        \
	do {                                                                         \
		if (cond) std::printf("  verified: %s\n", what);                         \
		else { std::printf("  NOT verified: %s\n", what); g_failed++; }          \
	} while (0)

// f = 1 + sum_i (i+1) x_i                       (LCRealFunction.cpp:64-71, linear variant)
class LinearTestFunction : public LCFunction {
public:
	explicit LinearTestFunction(int n) : n_(n) {}
	int getNParams() const override { return n_; }
	double getMinRange(int) const override { return 0.0; }
	double getMaxRange(int) const override { return 1.0; }
	LCError evalShapeInfo(const std::vector<double> &p, LCFunctionValue **result) override
	{
		double v = 1;
		for (int i = 0; i < n_; i++) v += (i + 1) * p[i];
		*result = new LCRealFunctionValue(v);
		return LCError();
	}
private:
	int n_;
};

// f = 1 + sum_i sum_j a_ij sin(3 j x_i + p_ij)   (LCRealFunction.cpp:50-63 with fixed coefficients)
class SineTestFunction : public LCFunction {
public:
	explicit SineTestFunction(int n) : n_(n) {}
	int getNParams() const override { return n_; }
	double getMinRange(int) const override { return 0.0; }
	double getMaxRange(int) const override { return 1.0; }
	double value(const std::vector<double> &p) const
	{
		double v = 1;
		for (int i = 0; i < n_; i++)
			for (int j = 0; j < 10; j++) v += (0.05 + 0.01 * ((7 * i + 3 * j) % 11)) * std::sin(3 * j * p[i] + 0.1 * (i + j));
		return v;
	}
	LCError evalShapeInfo(const std::vector<double> &p, LCFunctionValue **result) override
	{
		*result = new LCRealFunctionValue(value(p));
		return LCError();
	}
private:
	int n_;
};

// a 3-component "mesh": (f, 2f, f*f)
class VectorTestFunction : public SineTestFunction {
public:
	explicit VectorTestFunction(int n) : SineTestFunction(n) {}
	LCError evalShapeInfo(const std::vector<double> &p, LCFunctionValue **result) override
	{
		const double v = value(p);
		*result = new LCVectorFunctionValue(std::vector<double>{v, 2 * v, v * v});
		return LCError();
	}
};

// outputMatrix (TestExamples.cpp:108-142): n x n sweep over the first two parameters
static LCError sweep2d(LCFunction *f, int n, std::vector<double> *vals)
{
	vals->clear();
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++) {
			std::vector<double> p = {f->getMinRange(0) + (f->getMaxRange(0) - f->getMinRange(0)) * i / (n - 1.0),
			                         f->getMinRange(1) + (f->getMaxRange(1) - f->getMinRange(1)) * j / (n - 1.0)};
			LCFunctionValue *v = nullptr;
			LCErrorReturn(f->evalShapeInfo(p, &v));
			vals->push_back(v->data()[0]);
			delete v;
		}
	return LCError();
}

static void protoTest(const std::string &pb, const std::string &dump, bool hostOnly)
{
	std::printf("protoTest (%s)\n", pb.c_str());
	LCSampledFunction *shape = nullptr;
	LCError err = PrecomputedParametricShapeConverter::loadFromProto(pb, false, &shape, hostOnly ? LOCO_HOST_ONLY : 0);
	CHECK(err.isOK(), "loadFromProto");
	if (!err.isOK()) { std::printf("    %s\n", err.internalDescription().c_str()); return; }
	const std::string tmp = dump + ".roundtrip.pb";
	CHECK(PrecomputedParametricShapeConverter::outputToProto(tmp, false, shape).isOK(), "outputToProto (binary)");
	LCSampledFunction *again = nullptr;
	CHECK(PrecomputedParametricShapeConverter::loadFromProto(tmp, false, &again, hostOnly ? LOCO_HOST_ONLY : 0).isOK(), "reload of the written proto");
	CHECK(PrecomputedParametricShapeConverter::outputToProto(tmp + ".txt", true, shape).isOK(), "outputToProto (ascii)");
	LCSampledFunction *fromText = nullptr;
	CHECK(PrecomputedParametricShapeConverter::loadFromProto(tmp + ".txt", true, &fromText, hostOnly ? LOCO_HOST_ONLY : 0).isOK(), "reload of the ascii proto");
	if (again && fromText) {
		bool same = again->getNLeaves() == shape->getNLeaves() && again->getNSamples() == shape->getNSamples() && again->getNBorderCells() == shape->getNBorderCells() &&
		            fromText->getNLeaves() == shape->getNLeaves() && fromText->getNSamples() == shape->getNSamples();
		for (int l = 0; same && l < shape->getNLeaves(); l++) {
			std::vector<int32_t> a, b, c;
			shape->getInfluencingSampleIds(l, &a); again->getInfluencingSampleIds(l, &b); fromText->getInfluencingSampleIds(l, &c);
			same = a == b && a == c;
		}
		CHECK(same, "leaf, sample and border counts and every influencing-sample set survive the round trip");
	}
	if (hostOnly) {
		LCFunctionValue *v = nullptr;
		LCError e = shape->evalShapeInfo(std::vector<double>(shape->getNParams(), 0.5), &v);
		CHECK(!e.isOK() && e.internalDescription().find("no CPU path") != std::string::npos, "evaluation on a host-only model fails loudly (no CPU path)");
	} else if (again) {
		std::vector<double> a, b;
		CHECK(sweep2d(shape, 11, &a).isOK() && sweep2d(again, 11, &b).isOK(), "11x11 evaluation sweeps");
		CHECK(a == b, "sweep of the reloaded function is identical");
		// the same sweep through the batch entry point
		std::vector<double> q, out(a.size());
		for (int i = 0; i < 11; i++) for (int j = 0; j < 11; j++) { q.push_back(i / 10.0 * (shape->getMaxRange(0) - shape->getMinRange(0)) + shape->getMinRange(0)); q.push_back(j / 10.0 * (shape->getMaxRange(1) - shape->getMinRange(1)) + shape->getMinRange(1)); }
		std::vector<uint8_t> st(a.size());
		CHECK(shape->evalShapeInfoBatch(q.data(), (int64_t)a.size(), out.data(), nullptr, st.data()).isOK(), "evalShapeInfoBatch");
		CHECK(out == a, "batch of 121 == 121 batches of one");
		FILE *fp = std::fopen(dump.c_str(), "w");
		if (fp) {
			for (size_t i = 0; i < a.size(); i++) std::fprintf(fp, "%.17g %.17g %.17g\n", q[2 * i], q[2 * i + 1], a[i]);
			std::fclose(fp);
		}
		// LCError behaviour
		LCFunctionValue *v = nullptr;
		LCError e1 = shape->evalShapeInfo(std::vector<double>{shape->getMaxRange(0) + 0.5, shape->getMinRange(1)}, &v);
		CHECK(!e1.isOK() && e1.internalDescription() == "parameter ouside cell", "query outside the domain -> \"parameter ouside cell\"");
		LCError e2 = shape->evalShapeInfo(std::vector<double>{0.5}, &v);
		CHECK(!e2.isOK() && e2.internalDescription() == "invalid number of parameters", "wrong arity -> \"invalid number of parameters\"");
	}
	delete shape; delete again; delete fromText;
}

static void linearPrecisionTest(int N, LCBasisFunction::LCBasisFunctionType basis, int boot)
{
	std::printf("linearPrecisionTest N=%d %s bootstrap %d\n", N, basis == LCBasisFunction::CUBIC_BSPLINE ? "cubic" : "linear", boot);
	LinearTestFunction f(N);
	LCAdaptiveSamplingParams params(0, 0.1, boot, 0);
	LCAdaptiveGrid grid(&f, basis, &params, true);
	CHECK(grid.status().isOK(), "grid construction");
	if (!grid.status().isOK()) return;
	std::vector<double> ranges;
	f.getRanges(&ranges);
	// the reference's call pattern (TestExamples.cpp:196-199)
	LCSampledFunction *approx = new LCSampledFunction(grid.getRoot(), ranges, grid.getSamples(), grid.getBorderCells());
	approx->addBorderCells();
	int n = 11;
	long total = 1;
	for (int i = 0; i < N; i++) total *= n;
	double worst = 0, worstDeriv = 0;
	bool ok = true;
	for (long e = 0; e < total; e++) {
		std::vector<double> p(N);
		long r = e;
		for (int i = 0; i < N; i++) { p[i] = (r % n) / (n - 1.0); r /= n; }
		LCFunctionValue *a = nullptr, *b = nullptr;
		ok = ok && approx->evalShapeInfo(p, &a).isOK() && f.evalShapeInfo(p, &b).isOK();
		if (!ok) break;
		double diff;
		a->computeDifference(b, &diff);
		worst = std::max(worst, diff);
		delete a; delete b;
		bool interior = true;
		for (int i = 0; i < N; i++) interior = interior && p[i] > 0 && p[i] < 1;
		if (basis == LCBasisFunction::CUBIC_BSPLINE && interior)
			for (int d = 0; d < N; d++) {
				LCFunctionDeriv *dv = nullptr;
				ok = ok && approx->evalDeriv(p, d, &dv).isOK();
				if (!ok) break;
				worstDeriv = std::max(worstDeriv, std::fabs(static_cast<LCRealFunctionDeriv *>(dv)->val - (d + 1) / 2.0));
				delete dv;
			}
	}
	CHECK(ok, "all evaluations returned LCError OK");
	std::printf("    max |approx - exact| = %.3e, max derivative error = %.3e\n", worst, worstDeriv);
	CHECK(worst <= 1e-5, "linear precision within the reference's bound 1e-5");
	CHECK(worst <= 1e-12, "linear precision within 1e-12");
	CHECK(worstDeriv <= 1e-11, "cubic evalDeriv of a linear function is its slope (cube coordinates)");
	delete approx;
}

static void functionApproximationTest(double thr, int maxDepth, int boot)
{
	std::printf("testFunctionLinearApproximation (%g, %d, %d) cubic\n", thr, maxDepth, boot);
	SineTestFunction f(2);
	LCAdaptiveSamplingParams params(maxDepth, thr, boot, 0);
	LCAdaptiveGrid grid(&f, LCBasisFunction::CUBIC_BSPLINE, &params, true);
	CHECK(grid.status().isOK(), "grid construction");
	if (!grid.status().isOK()) return;
	std::vector<double> exact, approx;
	CHECK(sweep2d(&f, 21, &exact).isOK() && sweep2d(grid.getSampledFunction(), 21, &approx).isOK(), "21x21 sweeps of the function and of its interpolant");
	double worst = 0;
	bool finite = true;
	for (size_t i = 0; i < exact.size(); i++) { worst = std::max(worst, std::fabs(exact[i] - approx[i])); finite = finite && std::isfinite(approx[i]); }
	std::printf("    max |approx - exact| over the sweep = %.4f\n", worst);
	CHECK(finite && worst < 1.0, "interpolant is finite and tracks the function");
	// decideSplit for every leaf in one batch
	std::vector<uint8_t> split;
	std::vector<double> err;
	CHECK(grid.decideSplitBatch(&split, &err).isOK(), "decideSplitBatch CENTER_BASED");
	bool consistent = (int64_t)split.size() == grid.getSampledFunction()->getNLeaves();
	const bool mayRefine = 2 * boot < maxDepth;
	for (size_t l = 0; consistent && l < split.size(); l++) consistent = split[l] == (mayRefine && !((err[l] - thr) <= 0) ? 1 : 0);
	CHECK(consistent, "split[l] == (level < maxTreeDepth && !(err - threshold <= 0))");
	// the pass selection of the level-synchronous refinement loop: all leaves of a bootstrap grid have one size, so the
	// requests are exactly the flagged leaves, in descending order, each along dimension 0 (square cells: first widest)
	std::vector<LCAdaptiveGrid::SplitRequest> req;
	std::vector<double> err3;
	CHECK(LCAdaptiveGrid::selectSplits(grid.getSampledFunction(), thr, maxDepth, &req, &err3).isOK(), "selectSplits");
	size_t flagged = 0;
	for (size_t l = 0; l < split.size(); l++) flagged += split[l] ? 1 : 0;
	bool sel = req.size() == flagged && err3.size() == err.size();
	for (size_t i = 0; sel && i < req.size(); i++)
		sel = split[(size_t)req[i].leaf] == 1 && req[i].direction == 0 && (i == 0 || req[i].leaf < req[i - 1].leaf);
	for (size_t l = 0; sel && l < err.size(); l++) sel = err3[l] == err[l];
	CHECK(sel, "selectSplits returns the flagged leaves of the largest size class, descending, cyclic direction");
	CHECK(LCAdaptiveGrid::selectSplits(grid.getSampledFunction(), thr, 2 * boot, &req).isOK() && req.empty(), "no requests at the level limit");
	LCAdaptiveSamplingParams rnd(maxDepth + 4, thr, boot, 8, LCAdaptiveSamplingParams::ERROR_BASED, LCAdaptiveSamplingParams::RANDOM_SAMPLES);
	LCAdaptiveGrid grid2(&f, LCBasisFunction::CUBIC_BSPLINE, &rnd, true);
	std::vector<uint8_t> split2;
	std::vector<double> err2;
	srand(7);
	CHECK(grid2.decideSplitBatch(&split2, &err2).isOK(), "decideSplitBatch RANDOM_SAMPLES (8 jitter points per leaf)");
	bool c2 = split2.size() == split.size();
	for (size_t l = 0; c2 && l < split2.size(); l++) c2 = split2[l] == (!((err2[l] - thr) <= 0) ? 1 : 0) && std::isfinite(err2[l]);
	CHECK(c2, "RANDOM_SAMPLES flags follow the per-leaf maximum error");
}

static void vectorValuedTest()
{
	std::printf("vector-valued function (3 components per sample)\n");
	VectorTestFunction f(2);
	SineTestFunction f0(2);
	LCAdaptiveSamplingParams params(0, 0.1, 2, 0);
	LCAdaptiveGrid grid(&f, LCBasisFunction::CUBIC_BSPLINE, &params, true), grid0(&f0, LCBasisFunction::CUBIC_BSPLINE, &params, true);
	CHECK(grid.status().isOK() && grid0.status().isOK(), "grid construction");
	if (!grid.status().isOK() || !grid0.status().isOK()) return;
	bool ok = true;
	double worst = 0;
	for (int i = 0; i <= 10 && ok; i++) {
		std::vector<double> p = {i / 10.0, 0.37};
		LCFunctionValue *v = nullptr, *s = nullptr;
		ok = grid.evalShapeInfo(p, &v).isOK() && grid0.evalShapeInfo(p, &s).isOK() && v->size() == 3;
		if (ok) worst = std::max(worst, std::max(std::fabs(v->data()[0] - s->data()[0]), std::fabs(v->data()[1] - 2 * s->data()[0])));
		delete v; delete s;
	}
	CHECK(ok && worst < 1e-13, "components 0 and 1 equal the scalar interpolant of f and 2f (linearity of the weighted sum)");
}

// The host that owns the grid refines it and hands the new state over as plain arrays (loco_graph_t); the function object
// stays the same.  A 1-D linear grid is small enough to write down by hand: n = 2^level cells over [-1,1], samples on the
// n+1 lattice points with one unit-weight hat each of support 2/n (LCAdaptiveGrid.cpp:268-296), cells in pre-order with
// midpoint splits (:234-249), each leaf listing its two corner samples (:212-231).
struct Grid1D {
	std::vector<double> domain, centers, values, support, weight, ranges, splitVal;
	std::vector<int64_t> bfOff, cellOff;
	std::vector<int32_t> splitDir, child2, cellSample;
	loco_graph_t desc;

	void cell(double lo, double hi, int j0, int nj, const std::vector<int> &leafSample)
	{
		const size_t me = splitDir.size();
		ranges.push_back(lo); ranges.push_back(hi);
		splitDir.push_back(-1); splitVal.push_back(0.0); child2.push_back(-1);
		if (nj == 1) {
			cellSample.push_back(leafSample[j0]); cellSample.push_back(leafSample[j0 + 1]);
			cellOff.push_back((int64_t)cellSample.size());
			return;
		}
		cellOff.push_back((int64_t)cellSample.size());
		const double mid = 0.5 * (lo + hi);
		splitDir[me] = 0; splitVal[me] = mid;
		cell(lo, mid, j0, nj / 2, leafSample);
		child2[me] = (int32_t)splitDir.size();
		cell(mid, hi, j0 + nj / 2, nj / 2, leafSample);
	}

	Grid1D(int level, int basisType, double lo, double hi, double (*f)(double))
	{
		const int n = 1 << level;
		domain = {lo, hi};
		std::vector<int> ids;
		bfOff.push_back(0);
		for (int j = 0; j <= n; j++) {
			const double alpha = (double)j / n, c = -(1.0 - alpha) + alpha;
			centers.push_back(c);
			values.push_back(f(lo * (1 - (c + 1) / 2) + hi * ((c + 1) / 2)));
			support.push_back(2.0 / n); weight.push_back(1.0);
			bfOff.push_back(j + 1);
			ids.push_back(j);
		}
		cellOff.push_back(0);
		cell(-1.0, 1.0, 0, n, ids);
		std::memset(&desc, 0, sizeof desc);
		desc.n_params = 1; desc.value_dim = 1; desc.basis_type = basisType; desc.domain = domain.data();
		desc.n_samples = n + 1; desc.sample_center = centers.data(); desc.sample_value = values.data();
		desc.bf_offset = bfOff.data(); desc.bf_center = centers.data(); desc.bf_support = support.data(); desc.bf_weight = weight.data();
		desc.n_cells = (int64_t)splitDir.size(); desc.cell_ranges = ranges.data(); desc.cell_split_dir = splitDir.data();
		desc.cell_split_val = splitVal.data(); desc.cell_child2 = child2.data();
		desc.cell_sample_offset = cellOff.data(); desc.cell_sample = cellSample.data();
	}
};

static double bumpy(double x) { return std::sin(2.5 * x) + 0.3 * x * x; }

static void updateFromGraphTest(bool hostOnly)
{
	std::printf("updateFromGraph (1-D linear grid, 2 -> 8 cells)\n");
	const double lo = 0.5, hi = 2.5;
	Grid1D coarse(1, LOCO_LINEAR_BSPLINE, lo, hi, bumpy), fine(3, LOCO_LINEAR_BSPLINE, lo, hi, bumpy), other(3, LOCO_CUBIC_BSPLINE, lo, hi, bumpy);
	loco_model_t *h = nullptr;
	CHECK(loco_model_from_graph(&coarse.desc, hostOnly ? LOCO_HOST_ONLY : 0, &h) == LOCO_OK && h, "model from the coarse grid");
	if (!h) return;
	LCSampledFunction f(h);
	CHECK(f.getNLeaves() == 2 && f.getNSamples() == 3, "2 leaves, 3 samples");
	LCError bad = f.updateFromGraph(other.desc);
	CHECK(!bad.isOK() && bad.internalDescription().find("same function") != std::string::npos, "a state with another basis type is refused");
	CHECK(f.getNLeaves() == 2, "and the refused update left the object as it was");
	CHECK(f.updateFromGraph(fine.desc).isOK(), "updateFromGraph to the refined state");
	CHECK(f.getNLeaves() == 8 && f.getNSamples() == 9, "8 leaves, 9 samples after the update");
	CHECK(f.getMinRange(0) == lo && f.getMaxRange(0) == hi, "the domain is kept");
	if (hostOnly) return;
	// the interpolant of the refined state: piecewise linear through f at the 9 lattice points
	const int nq = 97;
	std::vector<double> q(nq), out(nq);
	std::vector<int32_t> leaf(nq);
	std::vector<uint8_t> st(nq);
	for (int i = 0; i < nq; i++) q[i] = lo + (hi - lo) * (i + 0.37) / nq;
	CHECK(f.evalShapeInfoBatch(q.data(), nq, out.data(), leaf.data(), st.data()).isOK(), "evalShapeInfoBatch on the updated object");
	double worst = 0;
	bool okAll = true, leavesOk = true;
	for (int i = 0; i < nq; i++) {
		const double u = (q[i] - lo) / (hi - lo) * 8;
		const int c = (int)u;
		const double a = u - c, x0 = lo + (hi - lo) * c / 8, x1 = lo + (hi - lo) * (c + 1) / 8;
		worst = std::fmax(worst, std::fabs(out[i] - ((1 - a) * bumpy(x0) + a * bumpy(x1))));
		okAll = okAll && st[i] == LOCO_ST_OK;
		leavesOk = leavesOk && leaf[i] == c;  // leaves are numbered in pre-order = left to right in one dimension
	}
	CHECK(okAll && leavesOk, "every query is located in its cell of the refined grid");
	CHECK(worst <= 1e-12, "values equal the piecewise-linear interpolant of the refined state within 1e-12");
}

// LCAdaptiveGrid::adaptiveSample: the loop the reference's constructor runs (LCAdaptiveGrid.cpp:34, :481-545), here in
// level-synchronous passes.  Host-only: the host grid is created and handed to the model, then the error check refuses
// (no CPU path).  On the device: the loop terminates with no leaf flagged below the level limit, every split is accounted
// for, and the interpolant of the refined grid is closer to the function than the bootstrap grid's.
static void adaptiveSampleTest(bool hostOnly, LCBasisFunction::LCBasisFunctionType basis, LCAdaptiveSamplingParams::LCSplitPolicy policy,
                               LCAdaptiveSamplingParams::LCErrorMetric metric)
{
	std::printf("adaptiveSample %s, %s split direction, %s error metric\n", basis == LCBasisFunction::CUBIC_BSPLINE ? "cubic" : "linear",
	            policy == LCAdaptiveSamplingParams::ERROR_BASED ? "ERROR_BASED" : "CYCLIC", metric == LCAdaptiveSamplingParams::CENTER_BASED ? "CENTER_BASED" : "RANDOM_SAMPLES");
	SineTestFunction f(2);
	const double thr = 0.02;
	const int maxDepth = 9, boot = 2;
	LCAdaptiveSamplingParams params(maxDepth, thr, boot, 4, policy, metric);
	LCAdaptiveGrid grid(&f, basis, &params, true, hostOnly ? LOCO_HOST_ONLY : 0);
	CHECK(grid.status().isOK(), "grid construction");
	if (!grid.status().isOK()) return;
	LCSampledFunction *fn = grid.getSampledFunction();
	const int64_t leaves0 = fn->getNLeaves(), samples0 = fn->getNSamples();
	int passes = 0;
	int64_t splits = 0;
	if (hostOnly) {
		LCError e = grid.adaptiveSample(&passes, &splits);
		CHECK(!e.isOK() && e.internalDescription().find("no CPU path") != std::string::npos, "without a device the loop stops at its error check (no CPU path)");
		CHECK(grid.getHostGrid() && grid.getHostGrid()->nLeaves() == leaves0 && grid.getHostGrid()->nSamples() == samples0 &&
		      fn->getNLeaves() == leaves0 && fn->getNSamples() == samples0, "the host grid repeats the bootstrap phase and the model accepted its state");
		// the surgery itself needs no device: split leaf 5 along dimension 1 and hand the new state to the model
		loco_grid::Status st = grid.getHostGrid()->refineLeaf(5, 1);
		CHECK(st.ok && fn->updateFromGraph(grid.getHostGrid()->graph()).isOK() && fn->getNLeaves() == leaves0 + 1 && fn->getNSamples() > samples0,
		      "refineCell on the host grid + updateFromGraph: one more leaf, new mid samples");
		return;
	}
	std::vector<double> exact, before, after;
	CHECK(sweep2d(&f, 41, &exact).isOK() && sweep2d(fn, 41, &before).isOK(), "41x41 sweeps of the function and of the bootstrap interpolant");
	srand(11);
	LCError e = grid.adaptiveSample(&passes, &splits);
	CHECK(e.isOK(), "adaptiveSample");
	if (!e.isOK()) { std::printf("    %s\n", e.internalDescription().c_str()); return; }
	std::printf("    %d passes, %lld splits: %lld -> %lld leaves, %lld -> %lld samples, %lld true-function evaluations by the host grid\n", passes, (long long)splits,
	            (long long)leaves0, (long long)fn->getNLeaves(), (long long)samples0, (long long)fn->getNSamples(), (long long)grid.getHostGrid()->nFunctionEvaluations());
	CHECK(passes > 2 && splits > 0 && fn->getNLeaves() == leaves0 + splits && grid.getHostGrid()->nLeaves() == fn->getNLeaves() &&
	      grid.getHostGrid()->nSamples() == fn->getNSamples(), "every split adds one leaf; model and host grid agree on leaves and samples");
	if (metric == LCAdaptiveSamplingParams::CENTER_BASED) {
		std::vector<LCAdaptiveGrid::SplitRequest> req;
		std::vector<double> err;
		CHECK(LCAdaptiveGrid::selectSplits(fn, thr, maxDepth, &req, &err).isOK() && req.empty(), "after the loop no leaf below the level limit is flagged");
		// leaves that are still above the threshold sit at the level limit
		bool limit = true;
		std::vector<double> box(4);
		for (size_t l = 0; l < err.size(); l++)
			if (!((err[l] - thr) <= 0)) {
				loco_leaf_box(fn->handle(), (int32_t)l, box.data());
				const long level = std::lround(std::log2(2.0 / (box[1] - box[0])) + std::log2(2.0 / (box[3] - box[2])));
				limit = limit && level >= maxDepth;
			}
		CHECK(limit, "leaves whose centre error exceeds the threshold are at maxTreeDepth");
	}
	CHECK(sweep2d(fn, 41, &after).isOK(), "41x41 sweep of the refined interpolant");
	double w0 = 0, w1 = 0, m0 = 0, m1 = 0;
	for (size_t i = 0; i < exact.size() && i < after.size(); i++) {
		w0 = std::max(w0, std::fabs(exact[i] - before[i])); w1 = std::max(w1, std::fabs(exact[i] - after[i]));
		m0 += std::fabs(exact[i] - before[i]) / exact.size(); m1 += std::fabs(exact[i] - after[i]) / exact.size();
	}
	std::printf("    |approx - exact| over the sweep: bootstrap grid max %.4f mean %.4f, refined grid max %.4f mean %.4f\n", w0, m0, w1, m1);
	CHECK(after.size() == exact.size() && m1 < 0.5 * m0, "the refined interpolant is on average at least twice as close to the function as the bootstrap grid's");
}

int main(int argc, char **argv)
{
	if (argc < 3) { std::printf("usage: %s <golden .pb> <dump file> [--host-only]\n", argv[0]); return 2; }
	const bool hostOnly = argc > 3 && !std::strcmp(argv[3], "--host-only");
	protoTest(argv[1], argv[2], hostOnly);
	updateFromGraphTest(hostOnly);
	adaptiveSampleTest(hostOnly, LCBasisFunction::CUBIC_BSPLINE, LCAdaptiveSamplingParams::ERROR_BASED, LCAdaptiveSamplingParams::CENTER_BASED);
	if (!hostOnly) {
		adaptiveSampleTest(false, LCBasisFunction::LINEAR_BSPLINE, LCAdaptiveSamplingParams::CYCLIC, LCAdaptiveSamplingParams::CENTER_BASED);
		adaptiveSampleTest(false, LCBasisFunction::LINEAR_BSPLINE, LCAdaptiveSamplingParams::ERROR_BASED, LCAdaptiveSamplingParams::RANDOM_SAMPLES);
		linearPrecisionTest(1, LCBasisFunction::CUBIC_BSPLINE, 2);
		linearPrecisionTest(3, LCBasisFunction::CUBIC_BSPLINE, 1);
		linearPrecisionTest(2, LCBasisFunction::LINEAR_BSPLINE, 2);
		functionApproximationTest(0.1, 3, 2);
		functionApproximationTest(0.5, 3, 1);
		vectorValuedTest();
	}
	std::printf("%s: %d check(s) failed\n", g_failed ? "FAILED" : "PASSED", g_failed);
	return g_failed;
}
