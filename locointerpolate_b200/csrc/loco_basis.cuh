// loco_basis.cuh -- device building blocks shared by all evaluation kernels.
#pragma once
#include "loco_device.h"

namespace loco {

// Query coordinates in registers.  NT > 0: fully unrolled, dynamic reads (the split dimension during
// descent) become a select chain.  NT == 0: runtime N <= kMaxN, coordinates live in local memory.
template <int NT>
struct Coord {
	double v[NT > 0 ? NT : kMaxN];
	__device__ __forceinline__ void set(int d, double val)
	{
		if (NT > 0) {
#pragma unroll
			for (int i = 0; i < (NT > 0 ? NT : 1); i++) if (i == d) v[i] = val;
		} else v[d] = val;
	}
	__device__ __forceinline__ double get(int d) const
	{
		if (NT > 0) {
			double r = v[0];
#pragma unroll
			for (int i = 1; i < (NT > 0 ? NT : 1); i++) r = (i == d) ? v[i] : r;
			return r;
		}
		return v[d];
	}
};

// LCFunction::mapToStandardHypercube (LCFunction.cpp:59-73): alpha = (p-min)/(max-min); x = alpha*2 - 1.
// Explicit _rn intrinsics: no contraction, IEEE division -- the leaf index depends on every bit of x.
template <int NT>
__device__ __forceinline__ void map_to_cube(const DeviceModel &m, const double *__restrict__ p, Coord<NT> &x)
{
	const int N = NT > 0 ? NT : m.N;
#pragma unroll
	for (int d = 0; d < N; d++) {
		const double num = __dsub_rn(p[d], m.dom_lo[d]);
		const double alpha = ((m.dom_exact_mask >> d) & 1u) ? __dmul_rn(num, m.dom_rcp[d]) : __ddiv_rn(num, m.dom_w[d]);
		x.set(d, __fma_rn(alpha, 2.0, -1.0));  // alpha*2 is exact, so the fused form rounds exactly like alpha*2 - 1
	}
}

// Point location in two halves.
// first half: the locate-grid lookup.  Returns the code of the deepest tree cell known to contain x (leaf-flagged if
// it is a leaf), 0 = start at the root.  Split from the descent so that callers with several queries per thread can
// issue all their lookups before consuming any of them.
template <int NT>
__device__ __forceinline__ uint32_t lut_lookup(const DeviceModel &m, const Coord<NT> &x)
{
	uint32_t c = 0;
	if (m.lut) {
		// locate grid: the grid cell is found with EXACT comparisons against the stored boundaries (the
		// multiplication only proposes a candidate), its entry names the deepest tree cell containing the whole
		// grid cell, and the descent resumes there.  Any inconsistency (NaN, estimate off by more than one
		// cell) falls back to the descent from the root, so the result is always the reference's leaf.
		const int N = NT > 0 ? NT : m.N;
		uint32_t idx = 0;
		int shift = 0;
		bool ok = true;
#pragma unroll
		for (int d = 0; d < N; d++) {
			const int bits = m.lut_bits[d];
			if (bits == 0) continue;
			const int last = (1 << bits) - 1;
			const double xv = x.get(d);
			if ((m.lut_uniform_mask >> d) & 1u) {
				// boundaries are exactly lo + i*h, h and 1/h powers of two: (x - lo) rounds monotonically and the multiplication is
				// exact, so floor() can only be off by one DOWNWARDS-missing, when the rounded difference landed ON boundary i
				// while x itself lies below it -- one exact fma gives that boundary
				int i = (int)fmin(fmax((xv - m.lut_lo[d]) * m.lut_inv[d], 0.0), (double)last);
				const double b0 = fma((double)i, m.lut_h[d], m.lut_lo[d]);
				if (i > 0 && xv < b0) i--;
				ok &= xv == xv;  // a NaN coordinate takes the descent from the root (child2 at every split)
				idx |= (uint32_t)i << shift;
				shift += bits;
				continue;
			}
			const double *B = m.lut_bounds + m.lut_boff[d];
			int i = (int)fmin(fmax((xv - m.lut_lo[d]) * m.lut_inv[d], 0.0), (double)last);
			double b0 = __ldg(B + i), b1 = __ldg(B + i + 1);
			if (i > 0 && xv < b0) { i--; b1 = b0; b0 = __ldg(B + i); }              // rare: the estimate rounded across a boundary
			else if (i < last && !(xv < b1)) { i++; b0 = b1; b1 = __ldg(B + i + 1); }
			ok &= (i == 0 || !(xv < b0)) && (i == last || xv < b1);
			idx |= (uint32_t)i << shift;
			shift += bits;
		}
		if (ok) c = __ldg(m.lut + idx);
	}
	return c;
}

// second half: LCAdaptiveGridCell::getCointainingLeaf (LCAdaptiveGridCell.cpp:260-272) from tree cell `c` down: strict '<'
// sends ties to child2.  One 16-byte read-only load per INNER level; the child codes name the leaf directly (see TreeCell).
template <int NT>
__device__ __forceinline__ int32_t descend_from(const DeviceModel &m, const Coord<NT> &x, uint32_t c)
{
	if (m.n_inner == 0) return 0;  // the root is a leaf
	if (c & kLeafFlag) return (int32_t)(c & (kLeafFlag - 1));
	for (;;) {
		const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(m.cells + c));
		const double split = __hiloint2double((int)raw.y, (int)raw.x);
		const uint32_t next = (x.get((int)(raw.z >> 28)) < split) ? (raw.z & kCodeMask) : raw.w;
		if (next & kLeafFlag) return (int32_t)(next & (kLeafFlag - 1));
		c = next;
	}
}

template <int NT>
__device__ __forceinline__ int32_t descend(const DeviceModel &m, const Coord<NT> &x)
{
	if (m.n_inner == 0) return 0;  // the root is a leaf
	return descend_from<NT>(m, x, lut_lookup<NT>(m, x));
}

// LCAdaptiveGridCell::evalPametersAreInCell (:438-454) against lo-EPSILON / hi+EPSILON precomputed on the
// host with the reference's own double operations.
template <int NT>
__device__ __forceinline__ uint8_t containment_status(const DeviceModel &m, const Coord<NT> &x, int32_t leaf)
{
	const int N = NT > 0 ? NT : m.N;
	bool outside = false;
#pragma unroll
	for (int d = 0; d < N; d++) {
		const double xv = x.get(d);
		outside |= (xv < m.leaf_lo_eps[(int64_t)leaf * N + d]) || (xv > m.leaf_hi_eps[(int64_t)leaf * N + d]);
	}
	return outside ? (uint8_t)ST_OUTSIDE_CELL : (uint8_t)ST_OK;
}

// The same test for a leaf REACHED BY DESCENT in a tree whose leaf boxes are exactly the boxes implied by
// the split planes (DeviceModel::path_consistent, verified at flatten time): the descent already guarantees
// lo_d <= x_d (child2: not x < split) and x_d < hi_d (child1) on every split plane, so the +-EPSILON test
// can only fail on a face of the ROOT box, where leaf bound == root bound bit for bit.  No per-leaf loads.
template <int NT>
__device__ __forceinline__ uint8_t located_status(const DeviceModel &m, const Coord<NT> &x, int32_t leaf)
{
	if (!m.path_consistent) {
		uint8_t st = containment_status<NT>(m, x, leaf);
		if (st == ST_OK) st = m.leaf_status[leaf];
		return st;
	}
	const int N = NT > 0 ? NT : m.N;
	bool outside = false;
#pragma unroll
	for (int d = 0; d < N; d++) {
		const double xv = x.get(d);
		outside |= (xv < m.root_lo_eps[d]) || (xv > m.root_hi_eps[d]);
	}
	uint8_t st = outside ? (uint8_t)ST_OUTSIDE_CELL : (uint8_t)ST_OK;
	if (st == ST_OK && m.any_leaf_status) st = m.leaf_status[leaf];
	return st;
}

__device__ __forceinline__ double linear_1d(double u);
__device__ __forceinline__ double cubic_1d(double u);

// ---- tensor-product leaves -------------------------------------------------------------------------------
// f[d][j] = b((x_d - c_d[j]) / s_d): the N*F one-dimensional factors of a leaf block (see DeviceModel)
// `uniform_knots` (DeviceModel::tensor_uniform, checked at flatten time): the F centres of every dimension are equally
// spaced by half a support, i.e. they are consecutive knots of ONE uniform cubic B-spline.  The four factors are then
// the four polynomial segments of that spline at t = (x - c_1)/h and are formed together from t (14 operations instead
// of four separate evaluations of the piecewise form); they differ from the separate evaluations by rounding only
// (and by < 1e-17 within the +-1e-6 acceptance band outside the leaf).
template <int NT, bool CUBIC, bool EXACT>
__device__ __forceinline__ void tensor_factors(const double *blk, const Coord<NT> &x, double (&f)[NT][CUBIC ? 4 : 2], bool uniform_knots = false)
{
	constexpr int F = CUBIC ? 4 : 2;
	// a NaN coordinate: the reference's `d <= 1` is false, every factor of that dimension is 0 (result 0.0, status OK);
	// the segment polynomials below would propagate the NaN instead, so such a query takes the piecewise form
	bool finite = true;
#pragma unroll
	for (int d = 0; d < NT; d++) finite &= x.get(d) == x.get(d);
	if (CUBIC && uniform_knots && finite) {
#pragma unroll
		for (int d = 0; d < NT; d++) {
			const double r = blk[NT * F + d];
			const double diff = x.get(d) - blk[d * F + 1];
			const double t = 2.0 * (EXACT ? diff * r : diff / r);
			const double omt = 1.0 - t, t2 = t * t, omt2 = omt * omt;
			f[d][0] = omt2 * omt * (1.0 / 6.0);
			f[d][1] = fma(t2, fma(3.0, t, -6.0), 4.0) * (1.0 / 6.0);
			f[d][2 % F] = fma(omt2, fma(3.0, omt, -6.0), 4.0) * (1.0 / 6.0);
			f[d][3 % F] = t2 * t * (1.0 / 6.0);
		}
		return;
	}
#pragma unroll
	for (int d = 0; d < NT; d++) {
		const double r = blk[NT * F + d];
		const double xv = x.get(d);
#pragma unroll
		for (int j = 0; j < F; j++) {
			const double diff = xv - blk[d * F + j];
			const double u = EXACT ? diff * r : diff / r;
			f[d][j] = CUBIC ? cubic_1d(u) : linear_1d(u);
		}
	}
}

// sum over the F^(LEVEL+1) leading entries of v (dimension 0 fastest) of prod_{d<=LEVEL} f[d][j_d] * v[...]
// -- LCFunctionValue::newFromWeightedSum with w_k = prod_d f[d][j_d], evaluated by sum factorisation.
template <int F, typename T = double>
__device__ __forceinline__ T pick_factor(const T (&row)[F], int j)
{
	T r = row[0];
#pragma unroll
	for (int i = 1; i < F; i++) r = (i == j) ? row[i] : r;
	return r;
}
__host__ __device__ constexpr int ipow_c(int b, int e) { return e <= 0 ? 1 : b * ipow_c(b, e - 1); }

template <int NT, int F, int LEVEL>
struct TensorContract {
	static __device__ __forceinline__ double run(const double *v, const double (&f)[NT][F], int stride)
	{
		double s = 0.0;
		if (ipow_c(F, LEVEL + 1) > 64) {
			// outer levels of big tensors stay rolled (code size); the factor is picked by a select chain so
			// that f[][] is never indexed dynamically and stays in registers
#pragma unroll 1
			for (int j = 0; j < F; j++) s += pick_factor<F>(f[LEVEL], j) * TensorContract<NT, F, (LEVEL > 0 ? LEVEL - 1 : 0)>::run(v + j * stride, f, stride / F);
		} else {
#pragma unroll
			for (int j = 0; j < F; j++) s += f[LEVEL][j] * TensorContract<NT, F, (LEVEL > 0 ? LEVEL - 1 : 0)>::run(v + j * stride, f, stride / F);
		}
		return s;
	}
};
template <int NT, int F>
struct TensorContract<NT, F, 0> {
	static __device__ __forceinline__ double run(const double *v, const double (&f)[NT][F], int)
	{
		double s = 0.0;
#pragma unroll
		for (int j = 0; j < F; j += 2) {
			const double2 p = *reinterpret_cast<const double2 *>(v + j);
			s += f[0][j] * p.x;
			s += f[0][j + 1] * p.y;
		}
		return s;
	}
};
// level 1: one F x F slab.  All of its values are requested FIRST (F*F/2 independent 128-bit loads in flight
// together), then contracted from registers: one exposed load latency per slab instead of one per row.
template <int NT, int F>
struct TensorContract<NT, F, 1> {
	static __device__ __forceinline__ double run(const double *v, const double (&f)[NT][F], int)
	{
		double2 t[F * F / 2];
#pragma unroll
		for (int i = 0; i < F * F / 2; i++) t[i] = reinterpret_cast<const double2 *>(v)[i];
		double s = 0.0;
#pragma unroll
		for (int j = 0; j < F; j++) {
			double s0 = 0.0;
#pragma unroll
			for (int i = 0; i < F / 2; i++) {
				s0 += f[0][2 * i] * t[j * (F / 2) + i].x;
				s0 += f[0][2 * i + 1] * t[j * (F / 2) + i].y;
			}
			s += f[1][j] * s0;
		}
		return s;
	}
};
template <int NT, bool CUBIC, bool EXACT>
__device__ __forceinline__ double tensor_eval_scalar(const DeviceModel &m, const double *blk, const Coord<NT> &x)
{
	constexpr int F = CUBIC ? 4 : 2;
	double f[NT][F];
	tensor_factors<NT, CUBIC, EXACT>(blk, x, f, m.tensor_uniform != 0);
	return TensorContract<NT, F, NT - 1>::run(blk + m.tl_voff, f, m.tensor_K / F);
}

// ---- per-leaf polynomial form -------------------------------------------------------------------------
// sum_j coef[j] prod_d t[d]^(j_d) by nested Horner (dimension 0 innermost): F^N - 1 fused multiply-adds and no basis
// evaluation at all, against N*F basis polynomials + (F^N + F^(N-1) + ...) FMAs of the factor/contraction form.
template <int NT, int F, int LEVEL>
struct TensorHorner {
	static __device__ __forceinline__ double run(const double *v, const double (&t)[NT], int stride)
	{
		if (ipow_c(F, LEVEL + 1) > 64) {
			double tl = t[0];
#pragma unroll
			for (int i = 1; i < NT; i++) tl = (i == LEVEL) ? t[i] : tl;
			double s = TensorHorner<NT, F, (LEVEL > 0 ? LEVEL - 1 : 0)>::run(v + (F - 1) * stride, t, stride / F);
#pragma unroll 1
			for (int j = F - 2; j >= 0; j--) s = fma(s, tl, TensorHorner<NT, F, (LEVEL > 0 ? LEVEL - 1 : 0)>::run(v + j * stride, t, stride / F));
			return s;
		} else {
			double s = TensorHorner<NT, F, (LEVEL > 0 ? LEVEL - 1 : 0)>::run(v + (F - 1) * stride, t, stride / F);
#pragma unroll
			for (int j = F - 2; j >= 0; j--) s = fma(s, t[LEVEL], TensorHorner<NT, F, (LEVEL > 0 ? LEVEL - 1 : 0)>::run(v + j * stride, t, stride / F));
			return s;
		}
	}
};
template <int NT, int F>
struct TensorHorner<NT, F, 0> {
	static __device__ __forceinline__ double run(const double *v, const double (&t)[NT], int)
	{
		double c[F];
#pragma unroll
		for (int j = 0; j < F; j += 2) {
			const double2 p = *reinterpret_cast<const double2 *>(v + j);
			c[j] = p.x; c[j + 1] = p.y;
		}
		double s = c[F - 1];
#pragma unroll
		for (int j = F - 2; j >= 0; j--) s = fma(s, t[0], c[j]);
		return s;
	}
};
// level 1: one F x F slab, all of it requested first (independent 128-bit loads in flight together), as in TensorContract
template <int NT, int F>
struct TensorHorner<NT, F, 1> {
	static __device__ __forceinline__ double run(const double *v, const double (&t)[NT], int)
	{
		double2 c[F * F / 2];
#pragma unroll
		for (int i = 0; i < F * F / 2; i++) c[i] = reinterpret_cast<const double2 *>(v)[i];
		double row[F];
#pragma unroll
		for (int j = 0; j < F; j++) {
			double s = c[j * (F / 2) + F / 2 - 1].y;
#pragma unroll
			for (int i = F - 2; i >= 0; i--) s = fma(s, t[0], (i & 1) ? c[j * (F / 2) + i / 2].y : c[j * (F / 2) + i / 2].x);
			row[j] = s;
		}
		double s = row[F - 1];
#pragma unroll
		for (int j = F - 2; j >= 0; j--) s = fma(s, t[1], row[j]);
		return s;
	}
};

// Value of a scalar function on a tensor-product leaf.  pblk != nullptr: the leaf's polynomial block (shared or global
// memory).  A query outside the leaf box (only possible in the +-1e-6 acceptance band around the ROOT box, where the basis
// functions switch pieces -- the linear hat is only C0 there) or with a NaN coordinate takes the factor form on the leaf's
// classic block, i.e. exactly the piecewise definition of LC{Linear,Cubic}BSpline::eval; blk == nullptr: read from global memory.
template <int NT, bool CUBIC, bool EXACT>
__device__ __forceinline__ double tensor_leaf_value(const DeviceModel &m, const double *blk, const double *pblk, int32_t leaf, const Coord<NT> &x)
{
	constexpr int F = CUBIC ? 4 : 2;
	if (pblk) {
		double t[NT];
		bool inside = true;
#pragma unroll
		for (int d = 0; d < NT; d++) {
			const double2 g = *reinterpret_cast<const double2 *>(pblk + 2 * d);
			t[d] = (x.get(d) - g.x) * g.y;
			inside &= fabs(t[d]) <= 1.0;
		}
		if (inside) return TensorHorner<NT, F, NT - 1>::run(pblk + 2 * NT, t, m.tensor_K / F);
		blk = nullptr;  // (a staged classic block is not available next to a staged polynomial block)
	}
	if (!blk) blk = m.tl_block + (int64_t)leaf * m.tl_stride;
	return tensor_eval_scalar<NT, CUBIC, EXACT>(m, blk, x);
}

// One dimension of LCLinearBSpline::eval (LCBasisFunction.cpp:283-288): d = |x-c|/s; d <= 1 ? 1-d : 0.
// u = (x-c)/s is passed in.
__device__ __forceinline__ double linear_1d(double u)
{
	const double d = fabs(u);
	return d <= 1.0 ? 1.0 - d : 0.0;
}

// One dimension of LCCubicBSpline::eval (LCBasisFunction.cpp:390-414).  The reference selects one of four
// pieces with k = floor(2(u+1)), t = 1-(2(u+1)-k); the uniform cubic B-spline is even in u, so with
// v = 2|u| the pieces k=1,2 are (3v^3 - 6v^2 + 4)/6 and the pieces k=0,3 are (2-v)^3/6 -- the same
// polynomials evaluated without the floor and with two branches instead of four.  Differs from the
// reference's expression by rounding only (a few ulp of a value <= 2/3); |u| > 1 and NaN give 0.
__device__ __forceinline__ double cubic_1d(double u)
{
	const double v = 2.0 * fabs(u);
	const double inner = fma(v * v, fma(3.0, v, -6.0), 4.0);
	const double w = 2.0 - v;
	const double outer = w * w * w;
	const double r = v < 1.0 ? inner : (v <= 2.0 ? outer : 0.0);
	return r * (1.0 / 6.0);
}

// LC{Linear,Cubic}BSpline::eval: value = prod_i dimVal_i, result = value * weight_.
// cs = (c_0, s_0, c_1, s_1, ...) with s_i replaced by 1/s_i when EXACT (every support a power of two,
// so the product is bit-identical to the division).
template <int NT, bool CUBIC, bool EXACT>
__device__ __forceinline__ double term_value(const double *__restrict__ cs, double weight, const Coord<NT> &x, int Nrt)
{
	const int N = NT > 0 ? NT : Nrt;
	double value = 1.0;
#pragma unroll
	for (int d = 0; d < N; d++) {
		const double2 c = __ldg(reinterpret_cast<const double2 *>(cs) + d);
		const double diff = x.get(d) - c.x;
		const double u = EXACT ? diff * c.y : diff / c.y;
		value *= CUBIC ? cubic_1d(u) : linear_1d(u);
	}
	return value * weight;
}

// ---- derivatives (cube coordinates) ------------------------------------------------------------------------
// One dimension of LCLinearBSpline::evalDeriv (LCBasisFunction.cpp:295-332), restated LITERALLY: the reference
// compares the SIGNED d = (x-c)/s with 1 (no abs), so for d < 0 a non-direction dimension yields 1-d (> 1) and
// d < -1 is not cut off.  rs = 1/s.
__device__ __forceinline__ double linear_1d_ref_deriv(double d, double rs, bool is_dir)
{
	if (!(d <= 1.0)) return 0.0;
	if (is_dir) return d < 0.0 ? rs : (d > 0.0 ? -rs : 0.0);
	return 1.0 - d;
}
// One dimension of LCCubicBSpline::evalDeriv in the derivative direction (LCBasisFunction.cpp:421-455): the
// reference's own piecewise form, p = 2(u+1), k = floor(p), t = 1-(p-k), dt = -2/s.
__device__ __forceinline__ double cubic_1d_deriv(double u, double rs)
{
	if (!(fabs(u) <= 1.0)) return 0.0;
	const double p = 2.0 * (u + 1.0);
	const double kf = floor(p);
	const double t = 1.0 - (p - kf);
	const double hdt = -rs;  // dt / 2
	const int k = (int)kf;
	double r = 0.0;
	if (k == 0) r = -(1.0 - t) * (1.0 - t) * hdt;
	else if (k == 1) r = (3.0 * t * t - 4.0 * t) * hdt;
	else if (k == 2) r = (-3.0 * t * t + 2.0 * t + 1.0) * hdt;
	else if (k == 3) r = t * t * hdt;
	return r;
}
// LC{Linear,Cubic}BSpline::evalDeriv: product over dimensions with the derivative factor in `dir`, times weight_
template <int NT, bool CUBIC, bool EXACT>
__device__ __forceinline__ double term_deriv(const double *__restrict__ cs, double weight, const Coord<NT> &x, int dir, int Nrt)
{
	const int N = NT > 0 ? NT : Nrt;
	double value = 1.0;
#pragma unroll
	for (int d = 0; d < N; d++) {
		const double2 c = __ldg(reinterpret_cast<const double2 *>(cs) + d);
		const double diff = x.get(d) - c.x;
		const double u = EXACT ? diff * c.y : diff / c.y;
		const double rs = EXACT ? c.y : 1.0 / c.y;
		double f;
		if (CUBIC) f = (d == dir) ? cubic_1d_deriv(u, rs) : cubic_1d(u);
		else f = linear_1d_ref_deriv(u, rs, d == dir);
		value *= f;
	}
	return value * weight;
}

// LCSampledFunction::evalDeriv on a tensor-product leaf: the leaf's terms are the F^N products of one-dimensional
// factors, so the derivative along `dir` is the same contraction with the factors of that dimension replaced by
// LC{Cubic,Linear}BSpline::evalDeriv's one-dimensional derivative (cubic_1d_deriv / the literal linear_1d_ref_deriv, whose
// signed comparison also changes the factors of the OTHER dimensions).  `blk` may be in shared or global memory.
template <int NT, bool CUBIC, bool EXACT>
__device__ __forceinline__ double tensor_eval_scalar_deriv(const DeviceModel &m, const double *blk, const Coord<NT> &x, int dir)
{
	constexpr int F = CUBIC ? 4 : 2;
	double f[NT][F];
#pragma unroll
	for (int d = 0; d < NT; d++) {
		const double r = blk[NT * F + d];
		const double rs = EXACT ? r : 1.0 / r;
		const double xv = x.get(d);
#pragma unroll
		for (int j = 0; j < F; j++) {
			const double diff = xv - blk[d * F + j];
			const double u = EXACT ? diff * r : diff / r;
			if constexpr (CUBIC) f[d][j] = (d == dir) ? cubic_1d_deriv(u, rs) : cubic_1d(u);
			else f[d][j] = linear_1d_ref_deriv(u, rs, d == dir);
		}
	}
	return TensorContract<NT, F, NT - 1>::run(blk + m.tl_voff, f, m.tensor_K / F);
}

// same as term_value with the term's (c, s) pairs in SHARED memory (broadcast reads) and the weight
// already multiplied by the sample value
template <int NT, bool CUBIC, bool EXACT>
__device__ __forceinline__ double term_value_smem(const double *cs, double wv, const Coord<NT> &x, int Nrt)
{
	const int N = NT > 0 ? NT : Nrt;
	double value = 1.0;
#pragma unroll
	for (int d = 0; d < N; d++) {
		const double2 c = reinterpret_cast<const double2 *>(cs)[d];
		const double diff = x.get(d) - c.x;
		const double u = EXACT ? diff * c.y : diff / c.y;
		value *= CUBIC ? cubic_1d(u) : linear_1d(u);
	}
	return value * wv;
}

// ---- optional fp32 mode (north star: 1e-5 relative) ---------------------------------------------------------------------
// Point location stays in fp64 (leaf indices must be bit-exact) and so does the normalised distance u = (x-c)/s; the
// basis polynomials, the products and the weighted sum run in fp32 on an fp32 copy of the leaf's sample values.
__device__ __forceinline__ float linear_1d_f32(float u) { const float d = fabsf(u); return d <= 1.0f ? 1.0f - d : 0.0f; }
__device__ __forceinline__ float cubic_1d_f32(float u)
{
	const float v = 2.0f * fabsf(u);
	const float inner = fmaf(v * v, fmaf(3.0f, v, -6.0f), 4.0f);
	const float w = 2.0f - v;
	const float r = v < 1.0f ? inner : (v <= 2.0f ? w * w * w : 0.0f);
	return r * (1.0f / 6.0f);
}
template <int NT, int F, int LEVEL>
struct TensorContractF32 {
	static __device__ __forceinline__ float run(const float *v, const float (&f)[NT][F], int stride)
	{
		float s = 0.0f;
		if (ipow_c(F, LEVEL + 1) > 64) {
			// outer levels of big tensors stay rolled (code size), as in TensorContract
#pragma unroll 1
			for (int j = 0; j < F; j++) s += pick_factor<F, float>(f[LEVEL], j) * TensorContractF32<NT, F, (LEVEL > 0 ? LEVEL - 1 : 0)>::run(v + j * stride, f, stride / F);
		} else {
#pragma unroll
			for (int j = 0; j < F; j++) s += f[LEVEL][j] * TensorContractF32<NT, F, (LEVEL > 0 ? LEVEL - 1 : 0)>::run(v + j * stride, f, stride / F);
		}
		return s;
	}
};
template <int NT, int F>
struct TensorContractF32<NT, F, 0> {
	static __device__ __forceinline__ float run(const float *v, const float (&f)[NT][F], int)
	{
		float s = 0.0f;
		if (F == 4) {
			const float4 p = *reinterpret_cast<const float4 *>(v);
			s = f[0][0] * p.x + f[0][1] * p.y + f[0][2 % F] * p.z + f[0][3 % F] * p.w;
		} else {
			const float2 p = *reinterpret_cast<const float2 *>(v);
			s = f[0][0] * p.x + f[0][1] * p.y;
		}
		return s;
	}
};
template <int NT, bool CUBIC, bool EXACT>
__device__ __forceinline__ double tensor_eval_scalar_f32(const DeviceModel &m, const double *blk, const float *vals, const Coord<NT> &x)
{
	constexpr int F = CUBIC ? 4 : 2;
	float f[NT][F];
#pragma unroll
	for (int d = 0; d < NT; d++) {
		const double r = blk[NT * F + d];
		const double xv = x.get(d);
#pragma unroll
		for (int j = 0; j < F; j++) {
			const double diff = xv - blk[d * F + j];
			const float u = (float)(EXACT ? diff * r : diff / r);
			f[d][j] = CUBIC ? cubic_1d_f32(u) : linear_1d_f32(u);
		}
	}
	return (double)TensorContractF32<NT, F, NT - 1>::run(vals, f, m.tensor_K / F);
}


} // namespace loco
