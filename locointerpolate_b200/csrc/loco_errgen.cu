// loco_errgen.cu -- the RANDOM_SAMPLES error check of the refinement loop in ONE kernel.
//
// LCAdaptiveGrid::decideSplit, RANDOM_SAMPLES branch (LCAdaptiveGrid.cpp:603-619): for every leaf, nErrorSamples_ points from
// getJitterPoint (:564-575), the interpolant evaluated IN that cell (cell->evalShapeInfo), the true function evaluated at
// the same point, err = |approx - truth| (LCRealFunctionValue::computeErrorVector, LCRealFunctionValue.cpp:37-47), split iff
// !(max(err - threshold) <= 0).
//
// The three-kernel form (jitter_points -> eval_tensor_direct -> error_reduce, loco_kernels.cu) moves every point, its truth,
// its leaf and its interpolated value through HBM and spends most of its time in 80 sin() calls per 8-D point.  Here one CTA
// owns a slice of ONE leaf's points: the leaf block (N*F centres, N reciprocals, F^N values) is copied to shared memory once,
// every thread generates its points from the counter-based hash, evaluates truth and interpolant in registers and keeps a
// running maximum; the CTA ends in one atomicMax.  Nothing but the per-leaf results touches HBM.
//
// Truth = the reference's test-function family (LCRealFunction.cpp:50-63) f(x) = 1 + sum_i sum_j a_ij sin(3 j x_i + p_ij).
// The n_terms sines of a dimension share the angle 3 x_i: sin(j t + p) = sin(j t) cos p + cos(j t) sin p with
// (sin j t, cos j t) from ONE sincos(t) by the rotation recurrence; a_ij cos p_ij and a_ij sin p_ij are formed once per CTA.
// Differs from n_terms separate sin() calls by rounding only (a few 1e-16 per term; tests hold the truth to 1e-12).
#include <algorithm>

#include "loco_basis.cuh"
#include "loco_device.h"

namespace loco {

namespace {

__device__ __forceinline__ uint64_t splitmix64_g(uint64_t z)
{
	z += 0x9e3779b97f4a7c15ull;
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}

constexpr int kMaxTruthTerms = 32;

template <int NT, bool CUBIC, bool EXACT>
__global__ void __launch_bounds__(256, 2) error_check_generated_kernel(DeviceModel m, int64_t ppl, int cpl, int64_t span, uint64_t seed,
                                                                     const double *__restrict__ amp, const double *__restrict__ phase, int n_terms,
                                                                     double threshold, unsigned long long *__restrict__ err_bits, uint8_t *__restrict__ split)
{
	extern __shared__ __align__(16) double s_mem[];  // [tl_stride leaf block | NT*n_terms a*cos(p) | NT*n_terms a*sin(p)]
	__shared__ unsigned long long s_max[8];
	__shared__ int s_split;
	const int32_t leaf = (int32_t)(blockIdx.x / cpl);
	const int64_t i0 = (int64_t)(blockIdx.x % cpl) * span, i1 = min(ppl, i0 + span);
	if (m.any_leaf_status && m.leaf_status[leaf] != ST_OK) return;  // the cell reports an error for every point: nothing to compare
	// the leaf's polynomial block where the model has one (every generated point lies inside the leaf box), else the classic block
	const bool poly = m.tl_poly != nullptr;
	const int blk_doubles = poly ? m.tp_stride : m.tl_stride;
	const double *src = poly ? m.tl_poly + (int64_t)leaf * m.tp_stride : m.tl_block + (int64_t)leaf * m.tl_stride;
	double *blk = s_mem;
	double *ac = s_mem + blk_doubles, *as = ac + NT * n_terms;
	for (int e = threadIdx.x; e < blk_doubles; e += blockDim.x) blk[e] = src[e];
	for (int e = threadIdx.x; e < NT * n_terms; e += blockDim.x) {
		double sp, cp;
		sincos(phase[e], &sp, &cp);
		ac[e] = amp[e] * cp;
		as[e] = amp[e] * sp;
	}
	if (threadIdx.x == 0) s_split = 0;
	__syncthreads();
	double lo[NT], w[NT], dlo[NT], dw[NT];
#pragma unroll
	for (int d = 0; d < NT; d++) {
		lo[d] = m.leaf_lo[(int64_t)leaf * NT + d];
		w[d] = __dsub_rn(m.leaf_hi[(int64_t)leaf * NT + d], lo[d]);
		dlo[d] = m.dom_min[d];
		dw[d] = m.dom_max[d];
	}
	unsigned long long best = 0ull;  // bit pattern of the running maximum: all candidates are >= +0, a (positive) NaN is the largest
	bool sp = false;
	for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
		const uint64_t p = (uint64_t)leaf * (uint64_t)ppl + (uint64_t)i;
		Coord<NT> x;
		double truth = 1.0;
#pragma unroll
		for (int d = 0; d < NT; d++) {
			// getJitterPoint: r is a FLOAT in [0,1] (both ends reachable, like rand()/RAND_MAX), pos = lo + r*(hi-lo)
			const uint64_t h = splitmix64_g(seed ^ splitmix64_g(p * (uint64_t)NT + (uint64_t)d));
			const float r = (float)(h >> 40) / 16777215.0f;
			const double pos = __dadd_rn(lo[d], __dmul_rn((double)r, w[d]));  // unfused, as the reference's C++ computes it (hosts can regenerate the points bit for bit)
			x.set(d, pos);
			// mapFromStandardHypercube (LCFunction.cpp:75-85), then the sine sum of this dimension
			const double alpha = (pos + 1.) / 2.;
			const double xo = dlo[d] * (1 - alpha) + dw[d] * alpha;
			double s1, c1;
			sincos(3.0 * xo, &s1, &c1);
			double sj = 0.0, cj = 1.0, fi = 0.0;
			const double *a_c = ac + d * n_terms, *a_s = as + d * n_terms;
			for (int j = 0; j < n_terms; j++) {
				fi = fma(a_c[j], sj, fma(a_s[j], cj, fi));
				const double ns = fma(sj, c1, cj * s1), nc = fma(cj, c1, -sj * s1);
				sj = ns; cj = nc;
			}
			truth += fi;
		}
		const double approx = poly ? tensor_leaf_value<NT, CUBIC, EXACT>(m, nullptr, blk, leaf, x) : tensor_eval_scalar<NT, CUBIC, EXACT>(m, blk, x);
		const double e = fabs(approx - truth);
		sp |= !((e - threshold) <= 0);
		const unsigned long long bits = (unsigned long long)__double_as_longlong(e);
		best = bits > best ? bits : best;
	}
	for (int off = 16; off; off >>= 1) {
		const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
		best = o > best ? o : best;
	}
	const bool any = __any_sync(0xffffffffu, sp);
	if ((threadIdx.x & 31) == 0) {
		s_max[threadIdx.x >> 5] = best;
		if (any) s_split = 1;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long mx = 0ull;
		for (int k = 0; k < (int)(blockDim.x >> 5); k++) mx = s_max[k] > mx ? s_max[k] : mx;
		if (i1 > i0) atomicMax(err_bits + leaf, mx);
		if (s_split) split[leaf] = 1;
	}
}

template <int NT, bool CUBIC, bool EXACT>
void run(const DeviceModel &m, int64_t ppl, uint64_t seed, const double *amp, const double *phase, int n_terms, double threshold, double *err_max,
         uint8_t *split, int sm_count, cudaStream_t st)
{
	// enough CTAs to fill the machine several times over, but never so many that a CTA has less than ~4 points per thread
	const int64_t want = (int64_t)sm_count * 16;
	int cpl = (int)std::max<int64_t>(1, std::min<int64_t>((want + m.L - 1) / m.L, (ppl + 1023) / 1024));
	const int64_t span = (ppl + cpl - 1) / cpl;
	cpl = (int)((ppl + span - 1) / span);
	const size_t smem = ((size_t)std::max(m.tl_stride, m.tp_stride) + 2 * (size_t)NT * n_terms) * sizeof(double);
	auto kern = error_check_generated_kernel<NT, CUBIC, EXACT>;
	if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	LOCO_KERNEL("error_check_generated", st);
	kern<<<(unsigned)((int64_t)m.L * cpl), 256, smem, st>>>(m, ppl, cpl, span, seed, amp, phase, n_terms, threshold,
	                                                         reinterpret_cast<unsigned long long *>(err_max), split);
}

} // namespace

bool error_check_generated_fused_supported(const DeviceModel &m, int n_terms)
{
	return m.D == 1 && m.tensor_F > 0 && m.N >= 1 && m.N <= kMaxTemplateN && n_terms >= 0 && n_terms <= kMaxTruthTerms &&
	       ((size_t)std::max(m.tl_stride, m.tp_stride) + 2 * (size_t)m.N * n_terms) * sizeof(double) <= 160 * 1024;
}

int launch_error_check_generated_fused(const DeviceModel &m, int64_t ppl, uint64_t seed, const double *amp, const double *phase, int n_terms, double threshold,
                                       double *err_max, uint8_t *split, int sm_count, cudaStream_t st)
{
	if (ppl <= 0 || m.L <= 0) return 0;
	const bool cubic = m.basis == CUBIC_BSPLINE, exact = m.exact_rcp != 0;
#define LOCO_GCASE(NT)                                                                                                     \
	case NT:                                                                                                               \
		if (cubic) { if (exact) run<NT, true, true>(m, ppl, seed, amp, phase, n_terms, threshold, err_max, split, sm_count, st);  \
		             else run<NT, true, false>(m, ppl, seed, amp, phase, n_terms, threshold, err_max, split, sm_count, st); }     \
		else { if (exact) run<NT, false, true>(m, ppl, seed, amp, phase, n_terms, threshold, err_max, split, sm_count, st);       \
		       else run<NT, false, false>(m, ppl, seed, amp, phase, n_terms, threshold, err_max, split, sm_count, st); }          \
		break;
	switch (m.N) { LOCO_GCASE(1) LOCO_GCASE(2) LOCO_GCASE(3) LOCO_GCASE(4) LOCO_GCASE(5) LOCO_GCASE(6) LOCO_GCASE(7) LOCO_GCASE(8) default: return 0; }
#undef LOCO_GCASE
	return 1;
}

} // namespace loco
