// loco_kernels.cu -- sm_100a kernels of the LOCO evaluation hot path (direct path + helpers).
//
// Reference semantics restated per kernel; arithmetic that decides the LEAF INDEX is done with explicit
// round-to-nearest intrinsics in the reference's operation order so that cell indices are bit-exact;
// the weight / sum arithmetic may contract to FMA (value parity bar is 1e-12 relative, fp64).
#include "loco_device.h"
#include "loco_basis.cuh"

namespace loco {

// ---------------------------------------------------------------------------------------------------
// fused scalar evaluation, one thread per query            (LCSampledFunction::evalShapeInfo, D == 1)
//   IN_CUBE = false : q is in original parameter space, the leaf is located by descent
//   IN_CUBE = true  : q already holds cube coordinates and leaf_in[] the cell to evaluate in
//                     (LCAdaptiveGridCell::evalShapeInfo called on a known cell, LCAdaptiveGrid.cpp:595,611)
// ---------------------------------------------------------------------------------------------------
template <int NT, bool CUBIC, bool EXACT, bool IN_CUBE>
__global__ void __launch_bounds__(256) eval_scalar_direct_kernel(DeviceModel m, const double *__restrict__ q, int64_t Q,
                                                                  double *__restrict__ out, int32_t *__restrict__ leaf_io,
                                                                  uint8_t *__restrict__ status_out)
{
	const int N = NT > 0 ? NT : m.N;
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= Q) return;
	Coord<NT> x;
	int32_t leaf;
	if (IN_CUBE) {
		for (int d = 0; d < N; d++) x.set(d, q[i * N + d]);
		leaf = leaf_io[i];
	} else {
		map_to_cube<NT>(m, q + i * N, x);
		leaf = descend<NT>(m, x);
	}
	uint8_t st;
	if (IN_CUBE) { st = containment_status<NT>(m, x, leaf); if (st == ST_OK) st = m.leaf_status[leaf]; }
	else st = located_status<NT>(m, x, leaf);
	double r = 0.0;  // LCRealFunctionValue(0), LCFunctionValue.cpp:22
	if (st == ST_OK) {
		int t = m.term_off[leaf];
		const int t1 = m.term_off[leaf + 1];
		int cur = -1;
		int row = 0;
		double comb = 0.0;  // LCSample::getInterpolationWeight accumulator, LCSample.cpp:78
		for (; t < t1; t++) {
			int slot = m.term_slot[t];
			if (slot != cur) {
				if (cur >= 0) r += comb * m.values[(int64_t)row * m.value_stride];  // val += weight * other->val
				comb = 0.0;
				cur = slot;
				row = m.term_row[t];
			}
			comb += term_value<NT, CUBIC, EXACT>(m.term_cs + (int64_t)t * 2 * N, m.term_w[t], x, N);
		}
		if (cur >= 0) r += comb * m.values[(int64_t)row * m.value_stride];
	} else {
		r = __longlong_as_double(0x7ff8000000000000LL);  // the reference returns an LCError and no value
	}
	out[i] = r;
	if (!IN_CUBE && leaf_io) leaf_io[i] = leaf;
	if (status_out) status_out[i] = st;
}

// ---------------------------------------------------------------------------------------------------
// derivative along one cube direction, one thread per query        (LCSampledFunction::evalDeriv,
// LCSampledFunction.cpp:61-67 -> LCAdaptiveGridCell::evalDeriv :678-711 -> LCSample::getDerivInterpolationWeight
// LCSample.cpp:92-105 -> LCFunctionDeriv::newFromWeightedSum / LCRealFunctionDeriv::add)
// ---------------------------------------------------------------------------------------------------
template <int NT, bool CUBIC, bool EXACT, bool UNUSED>
__global__ void __launch_bounds__(256) eval_deriv_direct_kernel(DeviceModel m, const double *__restrict__ q, int64_t Q, int direction,
                                                                 double *__restrict__ out, int32_t *__restrict__ leaf_out,
                                                                 uint8_t *__restrict__ status_out)
{
	const int N = NT > 0 ? NT : m.N;
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= Q) return;
	Coord<NT> x;
	map_to_cube<NT>(m, q + i * N, x);
	const int32_t leaf = descend<NT>(m, x);
	const uint8_t st = located_status<NT>(m, x, leaf);
	double r = __longlong_as_double(0x7ff8000000000000LL);
	if (st == ST_OK) {
		r = 0.0;  // LCRealFunctionDeriv(0)
		for (int t = m.term_off[leaf], t1 = m.term_off[leaf + 1]; t < t1; t++)
			r += term_deriv<NT, CUBIC, EXACT>(m.term_cs + (int64_t)t * 2 * N, m.term_w[t], x, direction, N) * m.values[(int64_t)m.term_row[t] * m.value_stride];
		if (!CUBIC)
			for (int t = m.dterm_off[leaf], t1 = m.dterm_off[leaf + 1]; t < t1; t++)
				r += term_deriv<NT, CUBIC, EXACT>(m.dterm_cs + (int64_t)t * 2 * N, m.dterm_w[t], x, direction, N) * m.values[(int64_t)m.dterm_row[t] * m.value_stride];
	}
	out[i] = r;
	if (leaf_out) leaf_out[i] = leaf;
	if (status_out) status_out[i] = st;
}

// ---------------------------------------------------------------------------------------------------
// locate only                                                (mapToStandardHypercube + getCointainingLeaf)
// ---------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256) locate_kernel(DeviceModel m, const double *__restrict__ q, int64_t Q, double *__restrict__ xs,
                                                      int32_t *__restrict__ leaf_out, uint8_t *__restrict__ status_out)
{
	const int N = NT > 0 ? NT : m.N;
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= Q) return;
	Coord<NT> x;
	map_to_cube<NT>(m, q + i * N, x);
	int32_t leaf = descend<NT>(m, x);
	uint8_t st = located_status<NT>(m, x, leaf);
	if (xs) for (int d = 0; d < N; d++) xs[i * N + d] = x.get(d);
	if (leaf_out) leaf_out[i] = leaf;
	if (status_out) status_out[i] = st;
}

// ---------------------------------------------------------------------------------------------------
// per-sample weights of the containing leaf                 (LCSample::getInterpolationWeight for every
// influencing sample; slot order = sorted support set)       W [Q * max_support]
// ---------------------------------------------------------------------------------------------------
template <int NT, bool CUBIC, bool EXACT, bool IN_CUBE>
__global__ void __launch_bounds__(256) weights_kernel(DeviceModel m, const double *__restrict__ q, int64_t Q, double *__restrict__ W,
                                                       int32_t *__restrict__ leaf_io, uint8_t *__restrict__ status_out)
{
	const int N = NT > 0 ? NT : m.N;
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= Q) return;
	Coord<NT> x;
	int32_t leaf;
	if (IN_CUBE) {
		for (int d = 0; d < N; d++) x.set(d, q[i * N + d]);
		leaf = leaf_io[i];
	} else {
		map_to_cube<NT>(m, q + i * N, x);
		leaf = descend<NT>(m, x);
	}
	uint8_t st;
	if (IN_CUBE) { st = containment_status<NT>(m, x, leaf); if (st == ST_OK) st = m.leaf_status[leaf]; }
	else st = located_status<NT>(m, x, leaf);
	double *w = W + i * (int64_t)m.max_support;
	const int K = m.sup_off[leaf + 1] - m.sup_off[leaf];
	for (int k = 0; k < m.max_support; k++) w[k] = 0.0;
	if (st == ST_OK) {
		int t = m.term_off[leaf];
		const int t1 = m.term_off[leaf + 1];
		int cur = -1;
		double comb = 0.0;
		for (; t < t1; t++) {
			int slot = m.term_slot[t];
			if (slot != cur) {
				if (cur >= 0) w[cur] = comb;
				comb = 0.0;
				cur = slot;
			}
			comb += term_value<NT, CUBIC, EXACT>(m.term_cs + (int64_t)t * 2 * N, m.term_w[t], x, N);
		}
		if (cur >= 0) w[cur] = comb;
	}
	(void)K;
	if (!IN_CUBE && leaf_io) leaf_io[i] = leaf;
	if (status_out) status_out[i] = st;
}

// ---------------------------------------------------------------------------------------------------
// weighted gather-sum of value rows, one CTA per query      (LCFunctionValue::newFromWeightedSum with the
// component-wise generalisation of LCRealFunctionValue::add: r[j] += w * v[j])
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mesh_gather_kernel(DeviceModel m, const double *__restrict__ W, const int32_t *__restrict__ leaf_in,
                                                           const uint8_t *__restrict__ status_in, double *__restrict__ out, int64_t out_stride)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	double *sw = reinterpret_cast<double *>(smem_raw);
	int32_t *srow = reinterpret_cast<int32_t *>(sw + m.max_support);
	const int64_t qi = blockIdx.x;
	const int32_t leaf = leaf_in[qi];
	const bool ok = status_in[qi] == ST_OK;
	const int k0 = m.sup_off[leaf];
	const int K = m.sup_off[leaf + 1] - k0;
	for (int k = threadIdx.x; k < K; k += blockDim.x) {
		sw[k] = W[qi * (int64_t)m.max_support + k];
		srow[k] = m.sup_row[k0 + k];
	}
	__syncthreads();
	double *o = out + qi * out_stride;
	const double nan = __longlong_as_double(0x7ff8000000000000LL);
	// 128-bit path needs 16-byte aligned output rows (value rows always are: value_stride is even for D > 1)
	const bool vec = ((out_stride & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
	const int D2 = vec ? (m.D >> 1) : 0;
	for (int j = threadIdx.x; j < D2; j += blockDim.x) {
		double2 acc = make_double2(0.0, 0.0);
		if (ok) {
			for (int k = 0; k < K; k++) {
				const double2 v = __ldg(reinterpret_cast<const double2 *>(m.values + (int64_t)srow[k] * m.value_stride) + j);
				const double w = sw[k];
				acc.x += w * v.x;
				acc.y += w * v.y;
			}
		} else {
			acc = make_double2(nan, nan);
		}
		reinterpret_cast<double2 *>(o)[j] = acc;
	}
	for (int j = 2 * D2 + threadIdx.x; j < m.D; j += blockDim.x) {
		double acc = 0.0;
		if (ok) for (int k = 0; k < K; k++) acc += sw[k] * m.values[(int64_t)srow[k] * m.value_stride + j];
		else acc = nan;
		o[j] = acc;
	}
}

__global__ void __launch_bounds__(256) checksum_kernel(const double *__restrict__ out, int64_t out_stride, int32_t D, double *__restrict__ checksum)
{
	__shared__ double red[8];
	const double *o = out + (int64_t)blockIdx.x * out_stride;
	double acc = 0.0;
	for (int j = threadIdx.x; j < D; j += blockDim.x) acc += o[j] * (double)(1 + j % 7);
	for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) {
		double s = 0.0;
		for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
		checksum[blockIdx.x] = s;
	}
}

// ---------------------------------------------------------------------------------------------------
// error-check reduction                      (LCRealFunctionValue::computeErrorVector + decideSplit's test)
// err = |approx - truth| per component; per leaf: max err, split |= !((err - threshold) <= 0)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) error_reduce_kernel(DeviceModel m, const double *__restrict__ out, const double *__restrict__ truth,
                                                            const int32_t *__restrict__ leaf_in, const uint8_t *__restrict__ status_in, int64_t P,
                                                            double threshold, unsigned long long *__restrict__ err_bits, uint8_t *__restrict__ split,
                                                            unsigned long long *__restrict__ n_bad)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	// no early return before the warp-wide primitives below: lanes without a point stay in the warp, flagged inactive
	bool active = i < P;
	if (active && status_in[i] != ST_OK) {
		if (n_bad) atomicAdd(n_bad, 1ull);
		active = false;
	}
	const unsigned live = __ballot_sync(0xffffffffu, active);
	if (!active) return;
	const int32_t leaf = leaf_in[i];
	double e = 0.0;
	bool sp = false;
	for (int j = 0; j < m.D; j++) {
		double d = fabs(out[i * m.D + j] - truth[i * m.D + j]);
		sp |= !((d - threshold) <= 0);
		// max over components with a STICKY NaN (a later finite component must not overwrite it): as a bit pattern a
		// positive NaN is larger than any finite value, so it also wins the atomicMax and the all_reduce(MAX) over ranks
		if (isnan(d) || (!isnan(e) && d > e)) e = fabs(d);
	}
	// warp-aggregate points of the same leaf before touching L2
	unsigned long long bits = (unsigned long long)__double_as_longlong(e);
	const unsigned peers = __match_any_sync(live, leaf);
	const int leader = __ffs(peers) - 1;
	unsigned long long mx = bits;
	bool any = sp;
	for (unsigned rest = peers; rest;) {
		int src = __ffs(rest) - 1;
		rest &= rest - 1;
		unsigned long long o = __shfl_sync(peers, bits, src);
		int os = __shfl_sync(peers, (int)sp, src);
		mx = o > mx ? o : mx;
		any |= os != 0;
	}
	if ((int)(threadIdx.x & 31) == leader) {
		atomicMax(err_bits + leaf, mx);
		if (any) split[leaf] = 1;
	}
}

__global__ void fill_values_kernel(double *__restrict__ values, int64_t R, int64_t D, int64_t stride, double seed)
{
	int64_t total = R * D;
	for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
		int64_t row = e / D, j = e - row * D;
		values[row * stride + j] = sin(0.37 * (double)j + 0.61 * (double)row + seed) + 0.001 * (double)(j % 97);
	}
}

// cube coordinates of every leaf centre: 0.5*(hi+lo)   (LCAdaptiveGridCell::getCenter, LCAdaptiveGridCell.cpp:204-210)
__global__ void leaf_centers_kernel(DeviceModel m, double *__restrict__ q_out)
{
	int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= (int64_t)m.L * m.N) return;
	q_out[e] = 0.5 * (m.leaf_hi[e] + m.leaf_lo[e]);
}

__global__ void gather_rows_kernel(DeviceModel m, const int32_t *__restrict__ rows, int64_t n, double *__restrict__ out)
{
	int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n * m.D) return;
	int64_t r = e / m.D, j = e - r * m.D;
	int32_t row = rows[r];
	out[e] = row >= 0 ? m.values[(int64_t)row * m.value_stride + j] : __longlong_as_double(0x7ff8000000000000LL);
}

// jitter points of the RANDOM_SAMPLES metric, generated from (leaf, i, seed) with a counter-based hash
// (getJitterPoint, LCAdaptiveGrid.cpp:564-575: r is a FLOAT in [0,1], pos = lo + r*(hi-lo) in cube
// coordinates), and the reference's test-function family as the truth (LCRealFunction.cpp:50-63) at
// mapFromStandardHypercube(pos) (LCFunction.cpp:75-85).
__device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
	z += 0x9e3779b97f4a7c15ull;
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) jitter_points_kernel(DeviceModel m, int64_t ppl, uint64_t seed, int64_t first, int64_t count,
                                                             const double *__restrict__ amp, const double *__restrict__ phase, int32_t n_terms,
                                                             double *__restrict__ q_out, int32_t *__restrict__ leaf_out, double *__restrict__ truth_out)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const int64_t p = first + i;
	const int32_t leaf = (int32_t)(p / ppl);
	double val = 1.0;
	for (int d = 0; d < m.N; d++) {
		uint64_t h = splitmix64(seed ^ splitmix64((uint64_t)p * (uint64_t)m.N + (uint64_t)d));
		float r = (float)(h >> 40) / 16777215.0f;  // 24 random bits -> float in [0,1], both ends reachable like rand()/RAND_MAX
		double lo = m.leaf_lo[(int64_t)leaf * m.N + d], hi = m.leaf_hi[(int64_t)leaf * m.N + d];
		double pos = __dadd_rn(lo, __dmul_rn((double)r, __dsub_rn(hi, lo)));  // unfused, as the reference's C++ computes it
		q_out[i * m.N + d] = pos;
		if (truth_out) {
			double alpha = (pos + 1.) / 2.;
			double x = m.dom_min[d] * (1 - alpha) + m.dom_max[d] * alpha;
			double fi = 0;
			for (int j = 0; j < n_terms; j++) fi += amp[d * n_terms + j] * sin(3 * j * x + phase[d * n_terms + j]);
			val += fi;
		}
	}
	leaf_out[i] = leaf;
	if (truth_out) truth_out[i] = val;
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
namespace {
inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

template <template <int, bool, bool, bool> class Launcher, bool IN_CUBE, typename... Args>
bool dispatch(const DeviceModel &m, Args... args)
{
	const bool cubic = m.basis == CUBIC_BSPLINE, exact = m.exact_rcp != 0;
#define LOCO_CASE(NT)                                                                               \
	case NT:                                                                                        \
		if (cubic) { if (exact) Launcher<NT, true, true, IN_CUBE>::run(m, args...); else Launcher<NT, true, false, IN_CUBE>::run(m, args...); } \
		else { if (exact) Launcher<NT, false, true, IN_CUBE>::run(m, args...); else Launcher<NT, false, false, IN_CUBE>::run(m, args...); }     \
		return true;
	switch (m.N <= kMaxTemplateN ? m.N : 0) {
		LOCO_CASE(0) LOCO_CASE(1) LOCO_CASE(2) LOCO_CASE(3) LOCO_CASE(4) LOCO_CASE(5) LOCO_CASE(6) LOCO_CASE(7) LOCO_CASE(8)
	}
#undef LOCO_CASE
	return false;
}

template <int NT, bool CUBIC, bool EXACT, bool IN_CUBE>
struct EvalScalarLauncher {
	static void run(const DeviceModel &m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, cudaStream_t st)
	{
		LOCO_KERNEL("eval_scalar_direct", st);
		eval_scalar_direct_kernel<NT, CUBIC, EXACT, IN_CUBE><<<blocks_for(Q, 256), 256, 0, st>>>(m, q, Q, out, leaf, status);
	}
};
template <int NT, bool CUBIC, bool EXACT, bool IN_CUBE>
struct WeightsLauncher {
	static void run(const DeviceModel &m, const double *q, int64_t Q, double *W, int32_t *leaf, uint8_t *status, cudaStream_t st)
	{
		LOCO_KERNEL("weights", st);
		weights_kernel<NT, CUBIC, EXACT, IN_CUBE><<<blocks_for(Q, 256), 256, 0, st>>>(m, q, Q, W, leaf, status);
	}
};
template <int NT, bool CUBIC, bool EXACT, bool UNUSED>
struct EvalDerivLauncher {
	static void run(const DeviceModel &m, const double *q, int64_t Q, int direction, double *out, int32_t *leaf, uint8_t *status, cudaStream_t st)
	{
		LOCO_KERNEL("eval_deriv_direct", st);
		eval_deriv_direct_kernel<NT, CUBIC, EXACT, UNUSED><<<blocks_for(Q, 256), 256, 0, st>>>(m, q, Q, direction, out, leaf, status);
	}
};
} // namespace

int launch_eval_deriv_direct(const DeviceModel &m, const double *q, int64_t Q, int direction, double *out, int32_t *leaf, uint8_t *status, cudaStream_t st)
{
	if (Q <= 0) return 0;
	dispatch<EvalDerivLauncher, false>(m, q, Q, direction, out, leaf, status, st);
	return 1;
}

int launch_eval_scalar_direct(const DeviceModel &m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, cudaStream_t st)
{
	if (Q <= 0) return 0;
	dispatch<EvalScalarLauncher, false>(m, q, Q, out, leaf, status, st);
	return 1;
}
int launch_eval_scalar_in_cube(const DeviceModel &m, const double *x, int64_t Q, double *out, int32_t *leaf, uint8_t *status, cudaStream_t st)
{
	if (Q <= 0) return 0;
	dispatch<EvalScalarLauncher, true>(m, x, Q, out, leaf, status, st);
	return 1;
}

int launch_locate(const DeviceModel &m, const double *q, int64_t Q, double *xs, int32_t *leaf, uint8_t *status, cudaStream_t st)
{
	if (Q <= 0) return 0;
	const unsigned g = blocks_for(Q, 256);
	switch (m.N <= kMaxTemplateN ? m.N : 0) {
	case 1: locate_kernel<1><<<g, 256, 0, st>>>(m, q, Q, xs, leaf, status); break;
	case 2: locate_kernel<2><<<g, 256, 0, st>>>(m, q, Q, xs, leaf, status); break;
	case 3: locate_kernel<3><<<g, 256, 0, st>>>(m, q, Q, xs, leaf, status); break;
	case 4: locate_kernel<4><<<g, 256, 0, st>>>(m, q, Q, xs, leaf, status); break;
	case 5: locate_kernel<5><<<g, 256, 0, st>>>(m, q, Q, xs, leaf, status); break;
	case 6: locate_kernel<6><<<g, 256, 0, st>>>(m, q, Q, xs, leaf, status); break;
	case 7: locate_kernel<7><<<g, 256, 0, st>>>(m, q, Q, xs, leaf, status); break;
	case 8: locate_kernel<8><<<g, 256, 0, st>>>(m, q, Q, xs, leaf, status); break;
	default: locate_kernel<0><<<g, 256, 0, st>>>(m, q, Q, xs, leaf, status); break;
	}
	return 1;
}

int launch_weights(const DeviceModel &m, const double *q, int64_t Q, double *W, int32_t *leaf, uint8_t *status, cudaStream_t st)
{
	if (Q <= 0) return 0;
	dispatch<WeightsLauncher, false>(m, q, Q, W, leaf, status, st);
	return 1;
}
int launch_weights_in_cube(const DeviceModel &m, const double *x, int64_t Q, double *W, int32_t *leaf, uint8_t *status, cudaStream_t st)
{
	if (Q <= 0) return 0;
	dispatch<WeightsLauncher, true>(m, x, Q, W, leaf, status, st);
	return 1;
}

int launch_mesh_gather(const DeviceModel &m, const double *W, const int32_t *leaf, const uint8_t *status, int64_t Q, double *out,
                       int64_t out_stride, cudaStream_t st)
{
	if (Q <= 0) return 0;
	size_t smem = (size_t)m.max_support * (sizeof(double) + sizeof(int32_t)) + 16;
	if (smem > 48 * 1024) cudaFuncSetAttribute(mesh_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	LOCO_KERNEL("mesh_gather", st);
	mesh_gather_kernel<<<(unsigned)Q, 256, smem, st>>>(m, W, leaf, status, out, out_stride);
	return 1;
}

int launch_checksum(const double *out, int64_t out_stride, int32_t D, int64_t Q, double *checksum, cudaStream_t st)
{
	if (Q <= 0) return 0;
	LOCO_KERNEL("checksum", st);
	checksum_kernel<<<(unsigned)Q, 256, 0, st>>>(out, out_stride, D, checksum);
	return 1;
}

int launch_error_reduce(const DeviceModel &m, const double *out, const double *truth, const int32_t *leaf, const uint8_t *status, int64_t P,
                        double threshold, double *err_max, uint8_t *split, unsigned long long *n_bad, cudaStream_t st)
{
	if (P <= 0) return 0;
	LOCO_KERNEL("error_reduce", st);
	error_reduce_kernel<<<blocks_for(P, 256), 256, 0, st>>>(m, out, truth, leaf, status, P, threshold,
	                                                       reinterpret_cast<unsigned long long *>(err_max), split, n_bad);
	return 1;
}

int launch_fill_values(double *values, int64_t R, int64_t D, int64_t stride, double seed, cudaStream_t st)
{
	if (R * D <= 0) return 0;
	fill_values_kernel<<<148 * 8, 256, 0, st>>>(values, R, D, stride, seed);
	return 1;
}

int launch_leaf_centers(const DeviceModel &m, double *q_out, cudaStream_t st)
{
	leaf_centers_kernel<<<blocks_for((int64_t)m.L * m.N, 256), 256, 0, st>>>(m, q_out);
	return 1;
}

int launch_gather_rows(const DeviceModel &m, const int32_t *rows, int64_t n, double *out, cudaStream_t st)
{
	if (n <= 0) return 0;
	gather_rows_kernel<<<blocks_for(n * m.D, 256), 256, 0, st>>>(m, rows, n, out);
	return 1;
}

int launch_jitter_points(const DeviceModel &m, int64_t points_per_leaf, uint64_t seed, int64_t first, int64_t count, const double *amp,
                         const double *phase, int32_t n_terms, double *q_out, int32_t *leaf_out, double *truth_out, cudaStream_t st)
{
	if (count <= 0) return 0;
	LOCO_KERNEL("jitter_points", st);
	jitter_points_kernel<<<blocks_for(count, 256), 256, 0, st>>>(m, points_per_leaf, seed, first, count, amp, phase, n_terms, q_out, leaf_out, truth_out);
	return 1;
}

} // namespace loco
