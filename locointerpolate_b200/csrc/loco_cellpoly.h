// loco_cellpoly.h -- the interpolant of a scalar function on ONE evaluation cell as a single tensor-product polynomial.
//
// SURVEY.md 8f.4 ("per-leaf tensor-polynomial coefficients when every overlapping basis function is a single polynomial
// piece"), applied per evaluation cell (Flat::ecell_*): a cell that no knot of any listed basis function crosses lies
// inside one polynomial piece of every one of them (LCLinearBSpline::eval, LCBasisFunction.cpp:283-288: two pieces;
// LCCubicBSpline::eval, :384-409: four pieces of the uniform cubic B-spline on knots c + s*{-1,-1/2,0,1/2,1}), so
//     sum_k wv_k * prod_d b((x_d - c_kd) / s_kd)  =  sum_j coef[j] * prod_d t_d^(j_d),    t_d = (x_d - mid_d) / halfwidth_d
// with F^N coefficients (F = 2 linear, 4 cubic).  The functions below are shared by the device kernel that forms the
// coefficients from the current sample values and by the host-side introspection the CPU tests check them through.
// This changes the summation order of the reference (allowed by SURVEY 8f.4 within the 1e-12 bar).
#pragma once

#ifdef __CUDACC__
#define LOCO_HD __host__ __device__
#else
#define LOCO_HD
#endif

namespace loco {

// p[0..F-1]: coefficients in t of  t -> b((mid + t*hw - c) / s)  on the piece that contains t = 0
// (the caller guarantees that no knot of b lies strictly inside (mid - hw, mid + hw)).
template <bool CUBIC>
LOCO_HD inline void cell_piece_1d(double c, double s, double mid, double hw, double *p)
{
	const double u0 = (mid - c) / s, du = hw / s;
	double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;  // the piece as a polynomial in u = (x - c) / s
	if (CUBIC) {
		if (u0 > -1.0 && u0 < 1.0) {
			if (u0 < -0.5) { q0 = 4.0 / 3.0; q1 = 4.0; q2 = 4.0; q3 = 4.0 / 3.0; }        // (4/3)(1+u)^3
			else if (u0 < 0.0) { q0 = 2.0 / 3.0; q2 = -4.0; q3 = -4.0; }                   // 2/3 - 4u^2 - 4u^3
			else if (u0 < 0.5) { q0 = 2.0 / 3.0; q2 = -4.0; q3 = 4.0; }                    // 2/3 - 4u^2 + 4u^3
			else { q0 = 4.0 / 3.0; q1 = -4.0; q2 = 4.0; q3 = -4.0 / 3.0; }                 // (4/3)(1-u)^3
		}
		// u = u0 + du*t:  p_j = sum_{k>=j} q_k C(k,j) u0^(k-j) du^j
		const double u02 = u0 * u0, du2 = du * du;
		p[0] = q0 + u0 * (q1 + u0 * (q2 + u0 * q3));
		p[1] = du * (q1 + 2.0 * q2 * u0 + 3.0 * q3 * u02);
		p[2] = du2 * (q2 + 3.0 * q3 * u0);
		p[3] = du2 * du * q3;
	} else {
		if (u0 > -1.0 && u0 < 1.0) { q0 = 1.0; q1 = u0 < 0.0 ? 1.0 : -1.0; }                // 1 - |u|
		p[0] = q0 + q1 * u0;
		p[1] = q1 * du;
	}
}

// true if a knot of b (centre c, half-support s) lies strictly inside (lo, hi)
template <bool CUBIC>
LOCO_HD inline bool knot_inside(double c, double s, double lo, double hi)
{
	const double tol = 1e-12 * (hi - lo);
	const int n = CUBIC ? 5 : 3;
	for (int j = 0; j < n; j++) {
		const double k = c + s * (CUBIC ? 0.5 * (double)(j - 2) : (double)(j - 1));
		if (k > lo + tol && k < hi - tol) return true;
	}
	return false;
}

// coef[0..F^N) += wv * prod_d p[d][j_d], dimension 0 fastest (the order TensorContract walks)
template <int F>
LOCO_HD inline void cell_poly_accumulate(int N, const double *p /* [N][F] */, double wv, double *coef)
{
	int P = 1;
	for (int d = 0; d < N; d++) P *= F;
	for (int idx = 0; idx < P; idx++) {
		double prod = wv;
		int r = idx;
		for (int d = 0; d < N; d++) { prod *= p[d * F + r % F]; r /= F; }
		coef[idx] += prod;
	}
}

} // namespace loco
