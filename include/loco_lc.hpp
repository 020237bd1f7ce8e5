// loco_lc.hpp -- C++ facade over the C ABI (loco_b200.h) that keeps the reference's class surface for the
// evaluation hot path, so that code written against mit-gfx/LOCOInterpolate compiles and runs on the GPU path:
//
//   reference class (file)                                          here
//   LCError                (LCUtils/LCError.h:7-20)                  same members, same semantics
//   LCFunctionValue        (LCFunction/LCFunctionValue.h:11-24)      value algebra used by callers (add, error vector)
//   LCRealFunctionValue    (LCRealFunction/LCRealFunctionValue.h)    Y = R, public member `val`
//   LCVectorFunctionValue  (new)                                     Y = R^D: per-vertex mesh vectors (the reference's mesh
//                                                                    type was removed from the public repo)
//   LCFunction             (LCFunction/LCFunction.h:20-49)           abstract f: ranges, hypercube maps, evalShapeInfo
//   LCBasisFunction        (LCAdaptiveSampling/LCBasisFunction.h:19) the type enum only (evaluation lives in the kernels)
//   LCAdaptiveSamplingParams (LCAdaptiveSamplingParams.h:8-39)       same fields and constructor
//   LCAdaptiveGrid         (LCAdaptiveGrid.h:18-75)                  bootstrap construction, the decideSplit error check batched, and
//                                                                    adaptiveSample as level-synchronous passes (error check on the GPU,
//                                                                    refineCell's surgery on the host: loco_grid.hpp)
//   LCSampledFunction      (LCSampledFunction.h:11-36)               evalShapeInfo (batch of one) + evalShapeInfoBatch
//   PrecomputedParametricShapeConverter (LCProtoConverter.h:11-21)   loadFromProto / outputToProto
//
// Conventions kept: LCError by value, results through out-pointers to `new`-ed objects the caller deletes, raw
// non-owning pointers, no exceptions.  Conventions changed: the object graph lives flattened on the GPU, so
// LCAdaptiveGridCell / LCSample are opaque tokens (getRoot()/getSamples()/getBorderCells() exist so that
// `new LCSampledFunction(grid->getRoot(), ranges, grid->getSamples(), grid->getBorderCells())` still compiles);
// Eigen::VectorXd arguments become std::vector<double>.  There is no CPU evaluation path: every evaluating
// call ends in a CUDA kernel launch through libloco_b200.so, and fails with an LCError if no GPU is usable.
//
// Define LOCO_LC_NAMESPACE (e.g. -DLOCO_LC_NAMESPACE=loco_lc) to put the classes in a namespace when linking next
// to the original library.
#ifndef LOCO_LC_HPP
#define LOCO_LC_HPP

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "loco_b200.h"
#include "loco_grid.hpp"

#ifdef LOCO_LC_NAMESPACE
namespace LOCO_LC_NAMESPACE {
#endif

#ifndef LCErrorReturn
#define LCErrorReturn(err) { LCError _lc_e = (err); if (!_lc_e.isOK()) { return _lc_e; } }
#endif

class LCError {
public:
	LCError() : ok_(true), text_("no error") {}
	LCError(std::string internalDescription) : ok_(false), text_(std::move(internalDescription)) {}
	std::string internalDescription() const { return text_; }
	bool isOK() const { return ok_; }

private:
	bool ok_;
	std::string text_;
};

// ---- values -------------------------------------------------------------------------------------------------
class LCFunctionValue {
public:
	virtual ~LCFunctionValue() {}
	virtual int size() const = 0;                 // doubles per value (D)
	virtual const double *data() const = 0;
	virtual double *data() = 0;
	// val += weight * other      (LCRealFunctionValue::add, LCRealFunctionValue.cpp:51-61, per component)
	virtual LCError add(double weight, LCFunctionValue *other)
	{
		if (!other || other->size() != size()) return LCError("incompatible shape types");
		for (int j = 0; j < size(); j++) data()[j] += weight * other->data()[j];
		return LCError();
	}
	// |a - b| per component      (LCRealFunctionValue::computeErrorVector, :37-47)
	virtual LCError computeErrorVector(LCFunctionValue *other, std::vector<double> *result)
	{
		if (!other || other->size() != size()) return LCError("incompatible shape types");
		result->resize(size());
		for (int j = 0; j < size(); j++) (*result)[j] = std::fabs(data()[j] - other->data()[j]);
		return LCError();
	}
	// Euclidean norm of the difference (LCRealFunctionValue::computeDifference, :24-35)
	virtual LCError computeDifference(LCFunctionValue *other, double *result)
	{
		std::vector<double> e;
		LCErrorReturn(computeErrorVector(other, &e));
		double s = 0;
		for (double v : e) s += v * v;
		*result = std::sqrt(s);
		return LCError();
	}
	virtual void log()
	{
		for (int j = 0; j < size(); j++) std::cout << (j ? " " : "val = ") << data()[j];
		std::cout << std::endl;
	}
};

class LCRealFunctionValue : public LCFunctionValue {
public:
	explicit LCRealFunctionValue(double v) : val(v) {}
	int size() const override { return 1; }
	const double *data() const override { return &val; }
	double *data() override { return &val; }
	double val;
};

class LCVectorFunctionValue : public LCFunctionValue {
public:
	explicit LCVectorFunctionValue(int n, double fill = 0.0) : vals(n, fill) {}
	explicit LCVectorFunctionValue(std::vector<double> v) : vals(std::move(v)) {}
	int size() const override { return (int)vals.size(); }
	const double *data() const override { return vals.data(); }
	double *data() override { return vals.data(); }
	std::vector<double> vals;
};

// derivative of a scalar function along one cube direction (LCRealFunctionDeriv.h:12-21)
class LCFunctionDeriv {
public:
	virtual ~LCFunctionDeriv() {}
	virtual void log() = 0;
};
class LCRealFunctionDeriv : public LCFunctionDeriv {
public:
	explicit LCRealFunctionDeriv(double v) : val(v) {}
	void log() override { std::cout << "deriv = " << val << std::endl; }
	double val;
};

// ---- LCFunction ---------------------------------------------------------------------------------------------
class LCFunction {
public:
	virtual ~LCFunction() {}
	virtual int getNParams() const = 0;
	virtual double getMinRange(int iParam) const = 0;
	virtual double getMaxRange(int iParam) const = 0;
	virtual LCError evalShapeInfo(const std::vector<double> &params, LCFunctionValue **result) = 0;
	virtual LCError evalDeriv(const std::vector<double> &, int, LCFunctionDeriv **) { return LCError("evalDeriv is not implemented for this function"); }

	LCError getMidPoint(std::vector<double> *result)
	{
		for (int i = 0; i < getNParams(); i++) {
			const double a = getMinRange(i), b = getMaxRange(i);
			if (b < a) return LCError("ranges are invalid for get MidPoint");
			result->push_back(0.5 * (a + b));
		}
		return LCError();
	}
	LCError validateParameters(const std::vector<double> &parameters) const
	{
		if ((int)parameters.size() != getNParams()) return LCError("number of parameters is incorrect");
		for (size_t i = 0; i < parameters.size(); i++)
			if (parameters[i] < getMinRange((int)i) || parameters[i] > getMaxRange((int)i)) return LCError("ranges are invalid for validateParameters");
		return LCError();
	}
	LCError getRelativeParamVal(int i, double alpha, double *result)
	{
		if (i >= getNParams() || alpha < 0 || alpha > 1) return LCError("parameter out of range");
		*result = getMinRange(i) * (1 - alpha) + getMaxRange(i) * alpha;
		return LCError();
	}
	// LCFunction.cpp:59-73 -- the kernels perform exactly these operations (loco_basis.cuh map_to_cube)
	std::vector<double> mapToStandardHypercube(const std::vector<double> &params)
	{
		std::vector<double> r(params.size());
		for (size_t i = 0; i < params.size(); i++) {
			const double alpha = (params[i] - getMinRange((int)i)) / (getMaxRange((int)i) - getMinRange((int)i));
			r[i] = alpha * 2 - 1;
		}
		return r;
	}
	std::vector<double> mapFromStandardHypercube(const std::vector<double> &unitParams)
	{
		std::vector<double> r(unitParams.size());
		for (size_t i = 0; i < unitParams.size(); i++) {
			const double alpha = (unitParams[i] + 1.) / 2.;
			r[i] = getMinRange((int)i) * (1 - alpha) + getMaxRange((int)i) * alpha;
		}
		return r;
	}
	void getRanges(std::vector<double> *ranges)
	{
		for (int i = 0; i < getNParams(); i++) { ranges->push_back(getMinRange(i)); ranges->push_back(getMaxRange(i)); }
	}
	void log() { std::cout << "nParams " << getNParams() << std::endl; }
};

class LCBasisFunction {
public:
	enum LCBasisFunctionType { LINEAR_BSPLINE = LOCO_LINEAR_BSPLINE, CUBIC_BSPLINE = LOCO_CUBIC_BSPLINE };
};

struct LCAdaptiveSamplingParams {
	enum LCSplitPolicy { ERROR_BASED, CYCLIC };
	enum LCErrorMetric { CENTER_BASED, RANDOM_SAMPLES };
	int maxTreeDepth_;
	double threshold_;
	int bootstrapGridSize_;
	int nErrorSamples_;
	LCSplitPolicy splitPolicy_;
	LCErrorMetric errorMetric_;
	LCAdaptiveSamplingParams(int maxTreeDepth = 0, double allowedError = 0.5, int bootstrap = 0, int errorSamples = 0,
	                         LCSplitPolicy policy = ERROR_BASED, LCErrorMetric errorMetric = CENTER_BASED)
	    : maxTreeDepth_(maxTreeDepth), threshold_(allowedError), bootstrapGridSize_(bootstrap), nErrorSamples_(errorSamples),
	      splitPolicy_(policy), errorMetric_(errorMetric) {}
	void log()
	{
		std::cout << "split policy " << splitPolicy_ << " maxTreeDepth " << maxTreeDepth_ << " threshold " << threshold_ << " bootstrapGridSize_ "
		          << bootstrapGridSize_ << " nErrorSamples " << nErrorSamples_ << " errorMetric " << errorMetric_ << std::endl;
	}
};

// ---- opaque graph tokens ------------------------------------------------------------------------------------
namespace loco_detail {
struct ModelHandle {
	loco_model_t *m = nullptr;
	explicit ModelHandle(loco_model_t *h) : m(h) {}
	~ModelHandle() { if (m) loco_free(m); }
	ModelHandle(const ModelHandle &) = delete;
	ModelHandle &operator=(const ModelHandle &) = delete;
};
inline LCError last_error(const loco_model_t *m) { return LCError(loco_last_error(m)); }
} // namespace loco_detail

// The k-d tree and the samples live flattened in GPU memory; these tokens only carry the model they belong to.
class LCAdaptiveGridCell {
public:
	std::shared_ptr<loco_detail::ModelHandle> model_;
};
class LCSample {
public:
	int index_ = 0;
};

// ---- LCSampledFunction --------------------------------------------------------------------------------------
class LCSampledFunction : public LCFunction {
public:
	// reference signature (LCSampledFunction.cpp:9-18).  `root` must come from LCAdaptiveGrid::getRoot() or
	// LCSampledFunction::getRoot(); ranges/samples/borderCells are implied by the model behind it.
	LCSampledFunction(LCAdaptiveGridCell *root, std::vector<double> ranges, std::vector<LCSample *> samples, std::vector<LCAdaptiveGridCell *> borderCells)
	{
		(void)samples; (void)borderCells;
		if (root) attach(root->model_);
		if (ok() && (int)ranges.size() == 2 * nParams_) ranges_ = ranges;
	}
	explicit LCSampledFunction(loco_model_t *handle) { attach(std::make_shared<loco_detail::ModelHandle>(handle)); }
	explicit LCSampledFunction(std::shared_ptr<loco_detail::ModelHandle> h) { attach(std::move(h)); }

	LCError addBorderCells() { return ok() ? LCError() : LCError("no model attached"); }  // done at flatten time (loco_flatten.cpp)
	int getNParams() const override { return nParams_; }
	double getMinRange(int iParam) const override { return ranges_[2 * iParam]; }
	double getMaxRange(int iParam) const override { return ranges_[2 * iParam + 1]; }

	// LCSampledFunction::evalShapeInfo (LCSampledFunction.cpp:69-78): a batch of one through the same kernels
	LCError evalShapeInfo(const std::vector<double> &params, LCFunctionValue **result) override
	{
		if (!ok()) return LCError("no model attached");
		if ((int)params.size() != nParams_) return LCError(loco_status_string(LOCO_ST_BAD_ARITY));
		std::vector<double> out((size_t)valueDim_);
		uint8_t status = 0;
		if (loco_eval_batch(model_->m, params.data(), 1, out.data(), nullptr, &status)) return loco_detail::last_error(model_->m);
		if (status != LOCO_ST_OK) return LCError(loco_status_string(status));
		if (valueDim_ == 1) *result = new LCRealFunctionValue(out[0]);
		else *result = new LCVectorFunctionValue(std::move(out));
		return LCError();
	}
	// the batched form the reference lacks: params [Q*N] row-major in ORIGINAL parameter space, out [Q*D];
	// leaf [Q] / status [Q] may be null.  status[i] != 0 carries the LCError the reference would have returned
	// for query i (loco_status_string), out is NaN there.
	LCError evalShapeInfoBatch(const double *params, int64_t nQueries, double *out, int32_t *leaf = nullptr, uint8_t *status = nullptr)
	{
		if (!ok()) return LCError("no model attached");
		if (loco_eval_batch(model_->m, params, nQueries, out, leaf, status)) return loco_detail::last_error(model_->m);
		return LCError();
	}
	// device buffers on the model's GPU, asynchronous on `cudaStream`
	LCError evalShapeInfoBatchDevice(const double *params, int64_t nQueries, double *out, int32_t *leaf, uint8_t *status, void *cudaStream)
	{
		if (!ok()) return LCError("no model attached");
		if (loco_eval_batch_device(model_->m, params, nQueries, out, leaf, status, cudaStream)) return loco_detail::last_error(model_->m);
		return LCError();
	}
	// LCSampledFunction::evalDeriv (LCSampledFunction.cpp:61-67): derivative along cube direction `direction`
	LCError evalDeriv(const std::vector<double> &params, int direction, LCFunctionDeriv **result) override
	{
		if (!ok()) return LCError("no model attached");
		if ((int)params.size() != nParams_) return LCError(loco_status_string(LOCO_ST_BAD_ARITY));
		if (valueDim_ != 1) return LCError("Shape info type not specified");
		double out = 0;
		uint8_t status = 0;
		if (loco_eval_deriv_batch(model_->m, params.data(), 1, direction, &out, nullptr, &status)) return loco_detail::last_error(model_->m);
		if (status != LOCO_ST_OK) return LCError(loco_status_string(status));
		*result = new LCRealFunctionDeriv(out);
		return LCError();
	}
	LCError evalDerivBatch(const double *params, int64_t nQueries, int direction, double *out, int32_t *leaf = nullptr, uint8_t *status = nullptr)
	{
		if (!ok()) return LCError("no model attached");
		if (loco_eval_deriv_batch(model_->m, params, nQueries, direction, out, leaf, status)) return loco_detail::last_error(model_->m);
		return LCError();
	}

	// The same object after the host refined its grid (LCAdaptiveGrid::refineCell, LCAdaptiveGrid.cpp:629-914): re-flatten
	// from the grid's new state, described as plain arrays (INTEGRATION.md section B shows how a reference-side binding
	// fills loco_graph_t).  Incremental: leaves outside the split cells' doubly-adjacent sets keep their evaluation cells
	// (loco_model_update_from_graph); getNReusedLeaves() tells how many.
	LCError updateFromGraph(const loco_graph_t &graph)
	{
		if (!ok()) return LCError("no model attached");
		if (loco_model_update_from_graph(model_->m, &graph)) return loco_detail::last_error(model_->m);
		if (loco_model_info(model_->m, &info_)) return loco_detail::last_error(model_->m);
		return LCError();
	}
	int64_t getNReusedLeaves() const { return info_.n_reused_leaves; }

	LCAdaptiveGridCell *getRoot() { return &root_; }
	std::vector<LCSample *> getSamples() { return std::vector<LCSample *>(); }
	std::vector<LCAdaptiveGridCell *> getBorderCells() { return std::vector<LCAdaptiveGridCell *>(); }

	// flat-model introspection (canonical numbering, include/loco_b200.h)
	bool ok() const { return model_ && model_->m; }
	loco_model_t *handle() const { return model_ ? model_->m : nullptr; }
	int getValueDim() const { return valueDim_; }
	LCBasisFunction::LCBasisFunctionType getBasisType() const { return basisType_; }
	int64_t getNLeaves() const { return info_.n_leaves; }
	int64_t getNSamples() const { return info_.n_samples; }
	int64_t getNBorderCells() const { return info_.n_border_cells; }
	// sorted sample indices of LCAdaptiveGridCell::getAllInfluencingSamples for a leaf
	LCError getInfluencingSampleIds(int leaf, std::vector<int32_t> *ids) const
	{
		if (!ok()) return LCError("no model attached");
		int32_t k = 0;
		if (loco_support_set(model_->m, leaf, nullptr, 0, &k)) return loco_detail::last_error(model_->m);
		ids->resize((size_t)k);
		if (k && loco_support_set(model_->m, leaf, ids->data(), k, &k)) return loco_detail::last_error(model_->m);
		return LCError();
	}

private:
	void attach(std::shared_ptr<loco_detail::ModelHandle> h)
	{
		model_ = std::move(h);
		root_.model_ = model_;
		if (!ok() || loco_model_info(model_->m, &info_)) { model_.reset(); root_.model_.reset(); return; }
		nParams_ = info_.n_params;
		valueDim_ = info_.value_dim;
		basisType_ = info_.basis_type == LOCO_CUBIC_BSPLINE ? LCBasisFunction::CUBIC_BSPLINE : LCBasisFunction::LINEAR_BSPLINE;
		ranges_.assign((size_t)2 * nParams_, 0.0);
		loco_model_domain(model_->m, ranges_.data());
	}
	std::shared_ptr<loco_detail::ModelHandle> model_;
	LCAdaptiveGridCell root_;
	loco_info_t info_{};
	int nParams_ = 0, valueDim_ = 1;
	LCBasisFunction::LCBasisFunctionType basisType_ = LCBasisFunction::LINEAR_BSPLINE;
	std::vector<double> ranges_;
};

// ---- proto I/O ----------------------------------------------------------------------------------------------
class PrecomputedParametricShapeConverter {
public:
	// LCProtoConverter.cpp:26-33.  `device`: CUDA device the model is uploaded to (LOCO_HOST_ONLY to inspect / re-save only)
	static LCError loadFromProto(std::string filename, bool useAscii, LCSampledFunction **shape, int device = 0)
	{
		loco_model_t *m = nullptr;
		if (loco_model_from_proto(filename.c_str(), useAscii ? 1 : 0, device, &m)) return LCError(loco_last_error(nullptr));
		*shape = new LCSampledFunction(m);
		return LCError();
	}
	// LCProtoConverter.cpp:14-24
	static LCError outputToProto(std::string filename, bool useAscii, LCSampledFunction *shape)
	{
		if (!shape || !shape->ok()) return LCError("no model attached");
		if (loco_model_save_proto(shape->handle(), filename.c_str(), useAscii ? 1 : 0)) return loco_detail::last_error(shape->handle());
		return LCError();
	}
};

// ---- LCAdaptiveGrid -----------------------------------------------------------------------------------------
// The bootstrap phase of the reference's builder (LCAdaptiveGrid.cpp:16-39 -> bootstrapSample :256-456) and its refinement
// loop (adaptiveSample :481-545) in level-synchronous form: decideSplit's error check (:578-626) for ALL
// leaves in one batch on the GPU, then refineCell (:629-914; host surgery, loco_grid.hpp) for the flagged leaves of the
// largest size class, then an incremental re-flatten (loco_model_update_from_graph).  Unlike the reference's constructor,
// which refines straight away, the constructor here stops after the bootstrap phase (so that the error check of that state
// can be inspected); call adaptiveSample() to run the loop.
class LCAdaptiveGrid {
public:
	LCAdaptiveGrid(LCFunction *parametricShape, LCBasisFunction::LCBasisFunctionType basisType, LCAdaptiveSamplingParams *params,
	               bool evalCorners = false, int device = 0)
	    : params_(params), parametricShape_(parametricShape), basisType_(basisType), evalCorners_(evalCorners)
	{
		err_ = build(evalCorners, device);
		if (!err_.isOK()) std::cout << "error in precomputation: " << err_.internalDescription() << std::endl;  // as LCAdaptiveGrid.cpp:28-38
	}
	LCError status() const { return err_; }

	// LCAdaptiveGrid::evalShapeInfo (LCAdaptiveGrid.cpp:92-99) for points inside the domain
	LCError evalShapeInfo(const std::vector<double> &params, LCFunctionValue **result)
	{
		if (!fn_) return err_;
		return fn_->evalShapeInfo(params, result);
	}
	LCAdaptiveGridCell *getRoot() { return fn_ ? fn_->getRoot() : nullptr; }
	std::vector<LCSample *> getSamples() { return std::vector<LCSample *>(); }
	std::vector<LCAdaptiveGridCell *> getBorderCells() { return std::vector<LCAdaptiveGridCell *>(); }
	LCSampledFunction *getSampledFunction() { return fn_.get(); }

	// decideSplit for every leaf at once.  split[l] = level(l) < maxTreeDepth_ && !(max(err - threshold) <= 0);
	// errMax[l] = largest component error seen in leaf l.  CENTER_BASED compares the interpolant at the leaf
	// centre with the stored centerShapeInfo_; RANDOM_SAMPLES draws nErrorSamples_ jitter points per leaf with
	// getJitterPoint's recipe (rand() as float, :564-575), evaluates the true function on the host (it is the
	// caller's CAD code) and the interpolant on the GPU.
	LCError decideSplitBatch(std::vector<uint8_t> *split, std::vector<double> *errMax)
	{
		LCErrorReturn(errorCheckBatch(split, errMax));
		const int bootLevel = fn_->getNParams() * params_->bootstrapGridSize_;
		if (!grid_ && !(bootLevel < params_->maxTreeDepth_)) split->assign(split->size(), 0);  // doSplit = level < maxLevel_ (:584); after adaptiveSample levels differ per leaf
		return LCError();
	}
	// the error part of decideSplit alone: split[l] = !(max(err - threshold) <= 0), without the level limit
	LCError errorCheckBatch(std::vector<uint8_t> *split, std::vector<double> *errMax)
	{
		if (!fn_) return err_;
		const int64_t L = fn_->getNLeaves();
		split->assign((size_t)L, 0);
		errMax->assign((size_t)L, 0.0);
		loco_model_t *m = fn_->handle();
		if (params_->errorMetric_ == LCAdaptiveSamplingParams::CENTER_BASED) {
			if (loco_center_error(m, params_->threshold_, errMax->data(), split->data())) return loco_detail::last_error(m);
		} else {
			const int N = fn_->getNParams(), D = fn_->getValueDim(), P = params_->nErrorSamples_;
			std::vector<double> box((size_t)2 * N), q, truth;
			q.reserve((size_t)L * P * N);
			truth.reserve((size_t)L * P * D);
			for (int64_t l = 0; l < L; l++) {
				if (loco_leaf_box(m, (int32_t)l, box.data())) return loco_detail::last_error(m);
				for (int i = 0; i < P; i++) {
					std::vector<double> cube((size_t)N);
					for (int d = 0; d < N; d++) {
						const float r = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
						cube[(size_t)d] = box[2 * d] + r * (box[2 * d + 1] - box[2 * d]);
					}
					std::vector<double> p = parametricShape_->mapFromStandardHypercube(cube);
					LCFunctionValue *v = nullptr;
					LCErrorReturn(parametricShape_->evalShapeInfo(p, &v));
					if (!v || v->size() != D) { delete v; return LCError("the size of the error vector and allowed dif are not the same"); }
					q.insert(q.end(), p.begin(), p.end());
					truth.insert(truth.end(), v->data(), v->data() + D);
					delete v;
				}
			}
			if (loco_error_check(m, q.data(), truth.data(), L * P, params_->threshold_, errMax->data(), split->data(), nullptr)) return loco_detail::last_error(m);
		}
		return LCError();
	}

	// One level-synchronous pass of adaptiveSample (LCAdaptiveGrid.cpp:481-545) for ANY sampled function, e.g. one wrapped
	// around a grid the caller refines itself: decideSplit for all leaves in one batch (CENTER_BASED, :594-602), then the
	// flagged leaves of the largest size class below maxLevel (what the reference's size-ordered heap would pop next) with
	// the cyclic direction rule (:553 generalised to N parameters: the first dimension of largest extent), in DESCENDING leaf
	// order so that splitting one of them (leaves are numbered depth-first, child1 first) leaves the others' indices valid.
	// The caller performs refineCell on its own grid for every request and re-flattens.
	struct SplitRequest { int32_t leaf; int direction; };
	static LCError selectSplits(LCSampledFunction *fn, double threshold, int maxLevel, std::vector<SplitRequest> *requests,
	                            std::vector<double> *errMax = nullptr)
	{
		requests->clear();
		if (!fn || !fn->handle()) return LCError("no model attached");
		const int64_t L = fn->getNLeaves();
		std::vector<double> err((size_t)L);
		std::vector<uint8_t> split((size_t)L);
		if (loco_center_error(fn->handle(), threshold, err.data(), split.data())) return loco_detail::last_error(fn->handle());
		LCErrorReturn(selectFromFlags(fn, split, maxLevel, requests));
		if (errMax) *errMax = std::move(err);
		return LCError();
	}
	// the selection alone, from the flags of any error check (CENTER_BASED or RANDOM_SAMPLES)
	static LCError selectFromFlags(LCSampledFunction *fn, const std::vector<uint8_t> &split, int maxLevel, std::vector<SplitRequest> *requests)
	{
		requests->clear();
		if (!fn || !fn->handle()) return LCError("no model attached");
		loco_model_t *m = fn->handle();
		const int64_t L = fn->getNLeaves();
		const int N = fn->getNParams();
		if ((int64_t)split.size() != L) return LCError("one flag per leaf expected");
		std::vector<double> volume((size_t)L), box((size_t)2 * N);
		std::vector<int> level((size_t)L), dir((size_t)L);
		double top = 0;
		for (int64_t l = 0; l < L; l++) {
			if (loco_leaf_box(m, (int32_t)l, box.data())) return loco_detail::last_error(m);
			double vol = 1, lv = 0, widest = -1;
			for (int d = 0; d < N; d++) {
				const double w = box[2 * d + 1] - box[2 * d];
				vol *= w;
				lv += std::log2(2.0 / w);
				if (w > widest) { widest = w; dir[(size_t)l] = d; }
			}
			volume[(size_t)l] = vol;
			level[(size_t)l] = (int)std::lround(lv);
			if (split[(size_t)l] && level[(size_t)l] < maxLevel) top = std::max(top, vol);
		}
		for (int64_t l = L - 1; l >= 0; l--)
			if (split[(size_t)l] && level[(size_t)l] < maxLevel && volume[(size_t)l] >= top * (1 - 1e-12))
				requests->push_back(SplitRequest{(int32_t)l, dir[(size_t)l]});
		return LCError();
	}

	// LCAdaptiveGrid::adaptiveSample (LCAdaptiveGrid.cpp:481-545) as level-synchronous passes: {error check of ALL leaves on the
	// GPU (errorCheckBatch: CENTER_BASED or RANDOM_SAMPLES) -> the flagged leaves of the largest size class below
	// maxTreeDepth_ (what the reference's size-ordered queue pops next) -> refineCell for each of them on the host grid
	// (loco_grid::Grid, the reference's surgery) -> incremental re-flatten of the model} until no leaf is flagged.
	// Split direction: ERROR_BASED = getOptimalSplitDirection (LCAdaptiveGridCell.cpp:977-1041) on the host grid's corner
	// values; CYCLIC = the first dimension of largest extent (decideSplitDirection :553 generalised to N parameters).
	// The host grid is created on the first call by repeating the bootstrap phase (f is evaluated at the lattice again).
	LCError adaptiveSample(int *nPasses = nullptr, int64_t *nSplits = nullptr, int maxPasses = 256)
	{
		if (nPasses) *nPasses = 0;
		if (nSplits) *nSplits = 0;
		if (!fn_) return err_;
		const int N = fn_->getNParams();
		if (!grid_) {
			std::vector<double> ranges;
			parametricShape_->getRanges(&ranges);
			std::unique_ptr<loco_grid::Grid> g(new loco_grid::Grid());
			loco_grid::Status st = loco_grid::Grid::bootstrap(N, (int)basisType_, params_->bootstrapGridSize_, valueDim_, evalCorners_, ranges.data(),
			                                                  &LCAdaptiveGrid::trampoline, this, g.get());
			if (!st.ok) return LCError(st.msg);
			if (!err_.isOK()) return err_;
			LCErrorReturn(fn_->updateFromGraph(g->graph()));
			grid_ = std::move(g);
		}
		for (int pass = 0; pass < maxPasses; pass++) {
			std::vector<uint8_t> split;
			std::vector<double> errMax;
			LCErrorReturn(errorCheckBatch(&split, &errMax));
			std::vector<SplitRequest> req;
			LCErrorReturn(selectFromFlags(fn_.get(), split, params_->maxTreeDepth_, &req));
			if (nPasses) *nPasses = pass + 1;
			if (req.empty()) return LCError();
			std::vector<int32_t> leaf, dir;
			for (const SplitRequest &r : req) {
				int d = r.direction;
				if (params_->splitPolicy_ == LCAdaptiveSamplingParams::ERROR_BASED) {
					loco_grid::Status st = grid_->optimalSplitDirection(r.leaf, &d);
					if (!st.ok) return LCError(st.msg);
				}
				leaf.push_back(r.leaf);
				dir.push_back(d);
			}
			loco_grid::Status st = grid_->refineLeaves(leaf.data(), dir.data(), (int64_t)leaf.size());
			if (!st.ok) return LCError(st.msg);
			if (!err_.isOK()) return err_;  // the function failed inside the surgery
			if (nSplits) *nSplits += (int64_t)leaf.size();
			LCErrorReturn(fn_->updateFromGraph(grid_->graph()));
		}
		return LCError("adaptiveSample: pass limit reached");
	}
	// the host grid behind adaptiveSample (nullptr before the first call)
	loco_grid::Grid *getHostGrid() { return grid_.get(); }

private:
	static void trampoline(const double *p, double *out, void *user)
	{
		LCAdaptiveGrid *g = static_cast<LCAdaptiveGrid *>(user);
		std::vector<double> params(p, p + g->parametricShape_->getNParams());
		LCFunctionValue *v = nullptr;
		LCError e = g->parametricShape_->evalShapeInfo(params, &v);
		if (!e.isOK() || !v || v->size() != g->valueDim_) { if (g->err_.isOK()) g->err_ = e.isOK() ? LCError("undefined type of shape info") : e; delete v; return; }
		for (int j = 0; j < g->valueDim_; j++) out[j] = v->data()[j];
		delete v;
	}
	LCError build(bool evalCorners, int device)
	{
		if (!parametricShape_ || !params_) return LCError("null argument");
		const int N = parametricShape_->getNParams();
		std::vector<double> mid, ranges;
		LCErrorReturn(parametricShape_->getMidPoint(&mid));
		LCFunctionValue *probe = nullptr;
		LCErrorReturn(parametricShape_->evalShapeInfo(mid, &probe));
		if (!probe) return LCError("undefined type of shape info");
		valueDim_ = probe->size();
		delete probe;
		parametricShape_->getRanges(&ranges);
		loco_model_t *m = nullptr;
		if (loco_model_bootstrap(N, (int32_t)basisType_, params_->bootstrapGridSize_, valueDim_, evalCorners ? 1 : 0, ranges.data(), &LCAdaptiveGrid::trampoline,
		                         this, device, &m))
			return LCError(loco_last_error(nullptr));
		if (!err_.isOK()) { loco_free(m); return err_; }
		fn_.reset(new LCSampledFunction(m));
		return LCError();
	}
	LCAdaptiveSamplingParams *params_;
	LCFunction *parametricShape_;
	LCBasisFunction::LCBasisFunctionType basisType_;
	bool evalCorners_ = false;
	int valueDim_ = 1;
	LCError err_;
	std::unique_ptr<LCSampledFunction> fn_;
	std::unique_ptr<loco_grid::Grid> grid_;
};

#ifdef LOCO_LC_NAMESPACE
} // namespace LOCO_LC_NAMESPACE
#endif
#endif // LOCO_LC_HPP
