// loco_tensor.cu -- direct (unsorted) evaluation on tensor-product leaves, and the refresh of the
// per-leaf value blocks.  See DeviceModel::tl_block and loco_basis.cuh (tensor_eval_scalar).
//
// One thread per query: locate, then read the leaf's block (N*F centres, N supports, F^N values) straight
// from global memory through L1/L2.  Used for small batches, for models whose blocks fit the L1, and for
// evaluation on a known cell (IN_CUBE: the refinement loop's centre / jitter points).
#include "loco_basis.cuh"
#include "loco_cellpoly.h"
#include "loco_device.h"

namespace loco {

template <int NT, bool CUBIC, bool EXACT, bool IN_CUBE>
__global__ void __launch_bounds__(256) eval_tensor_direct_kernel(DeviceModel m, const double *__restrict__ q, int64_t Q, double *__restrict__ out,
                                                                  int32_t *__restrict__ leaf_io, uint8_t *__restrict__ status_out)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= Q) return;
	Coord<NT> x;
	int32_t leaf;
	uint8_t st;
	if (IN_CUBE) {
#pragma unroll
		for (int d = 0; d < NT; d++) x.set(d, q[i * NT + d]);
		leaf = leaf_io[i];
		st = containment_status<NT>(m, x, leaf);
		if (st == ST_OK) st = m.leaf_status[leaf];
	} else {
		map_to_cube<NT>(m, q + i * NT, x);
		leaf = descend<NT>(m, x);
		st = located_status<NT>(m, x, leaf);
	}
	double r = __longlong_as_double(0x7ff8000000000000LL);
	if (st == ST_OK) r = tensor_leaf_value<NT, CUBIC, EXACT>(m, nullptr, m.tl_poly ? m.tl_poly + (int64_t)leaf * m.tp_stride : nullptr, leaf, x);
	out[i] = r;
	if (!IN_CUBE && leaf_io) leaf_io[i] = leaf;
	if (status_out) status_out[i] = st;
}

// tl_block[leaf][voff + j] = values[tl_row[leaf][j]]   (scalar functions; after every change of the value table)
__global__ void tensor_vals_kernel(DeviceModel m, double *__restrict__ tl_block, float *__restrict__ tl_vals_f32)
{
	int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= (int64_t)m.L * m.tensor_K) return;
	const int64_t leaf = e / m.tensor_K, j = e - leaf * m.tensor_K;
	const double v = m.values[(int64_t)m.tl_row[e] * m.value_stride];
	tl_block[leaf * m.tl_stride + m.tl_voff + j] = v;
	if (tl_vals_f32) tl_vals_f32[e] = (float)v;
}

// tl_poly[leaf] = [mid_d, 1/halfwidth_d per dimension | coefficients of the leaf's single tensor-product polynomial]
// (loco_cellpoly.h: every one-dimensional basis function of the leaf as its polynomial piece in t = (x - mid)/halfwidth, then
//  coef[idx] = sum_k value_k prod_d piece[d][k_d][idx_d]; one CTA per leaf).  Refreshed whenever the value table changes.
template <bool CUBIC>
__global__ void __launch_bounds__(128) tensor_poly_kernel(DeviceModel m, double *__restrict__ tl_poly, int32_t tp_stride)
{
	constexpr int F = CUBIC ? 4 : 2;
	__shared__ double P[kMaxTemplateN * F * F];  // [d][j][i]: coefficient i of basis function j of dimension d
	const int32_t leaf = blockIdx.x;
	const int N = m.N;
	const double *blk = m.tl_block + (int64_t)leaf * m.tl_stride;
	double *pb = tl_poly + (int64_t)leaf * tp_stride;
	for (int e = threadIdx.x; e < N * F; e += blockDim.x) {
		const int d = e / F;
		const double lo = m.leaf_lo[(int64_t)leaf * N + d], hi = m.leaf_hi[(int64_t)leaf * N + d];
		const double mid = 0.5 * (lo + hi), hw = 0.5 * (hi - lo);
		const double r = blk[N * F + d], sup = m.exact_rcp ? 1.0 / r : r;
		cell_piece_1d<CUBIC>(blk[e], sup, mid, hw, P + e * F);
		if (e % F == 0) { pb[2 * d] = mid; pb[2 * d + 1] = 1.0 / hw; }
	}
	__syncthreads();
	const int K = m.tensor_K;
	const double *vals = blk + m.tl_voff;
	for (int idx = threadIdx.x; idx < K; idx += blockDim.x) {
		double acc = 0.0;
		for (int k = 0; k < K; k++) {
			double prod = vals[k];
			int rk = k, ri = idx;
			for (int d = 0; d < N; d++) {
				prod *= P[(d * F + rk % F) * F + ri % F];
				rk /= F; ri /= F;
			}
			acc += prod;
		}
		pb[2 * N + idx] = acc;
	}
}

int launch_update_tensor_poly(const DeviceModel &m, double *tl_poly, int32_t tp_stride, cudaStream_t st)
{
	if (!tl_poly || m.tensor_F <= 0 || m.D != 1 || m.L <= 0) return 0;
	if (m.basis == CUBIC_BSPLINE) tensor_poly_kernel<true><<<(unsigned)m.L, 128, 0, st>>>(m, tl_poly, tp_stride);
	else tensor_poly_kernel<false><<<(unsigned)m.L, 128, 0, st>>>(m, tl_poly, tp_stride);
	return 1;
}

int launch_update_tensor_vals(const DeviceModel &m, double *tl_block, float *tl_vals_f32, cudaStream_t st)
{
	if (m.tensor_F <= 0 || m.D != 1) return 0;
	const int64_t n = (int64_t)m.L * m.tensor_K;
	tensor_vals_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m, tl_block, tl_vals_f32);
	return 1;
}

namespace {
template <int NT, bool IN_CUBE>
void run(const DeviceModel &m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, cudaStream_t st)
{
	const unsigned g = (unsigned)((Q + 255) / 256);
	LOCO_KERNEL("eval_tensor_direct", st);
	const bool cubic = m.basis == CUBIC_BSPLINE, exact = m.exact_rcp != 0;
	if (cubic) {
		if (exact) eval_tensor_direct_kernel<NT, true, true, IN_CUBE><<<g, 256, 0, st>>>(m, q, Q, out, leaf, status);
		else eval_tensor_direct_kernel<NT, true, false, IN_CUBE><<<g, 256, 0, st>>>(m, q, Q, out, leaf, status);
	} else {
		if (exact) eval_tensor_direct_kernel<NT, false, true, IN_CUBE><<<g, 256, 0, st>>>(m, q, Q, out, leaf, status);
		else eval_tensor_direct_kernel<NT, false, false, IN_CUBE><<<g, 256, 0, st>>>(m, q, Q, out, leaf, status);
	}
}
} // namespace

int launch_eval_scalar_tensor_direct(const DeviceModel &m, bool in_cube, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status,
                                     cudaStream_t st)
{
	if (Q <= 0) return 0;
#define LOCO_DCASE(NT) case NT: if (in_cube) run<NT, true>(m, q, Q, out, leaf, status, st); else run<NT, false>(m, q, Q, out, leaf, status, st); break;
	switch (m.N) { LOCO_DCASE(1) LOCO_DCASE(2) LOCO_DCASE(3) LOCO_DCASE(4) LOCO_DCASE(5) LOCO_DCASE(6) LOCO_DCASE(7) LOCO_DCASE(8) default: return 0; }
#undef LOCO_DCASE
	return 1;
}

} // namespace loco
