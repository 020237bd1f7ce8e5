// loco_tensor.cu -- direct (unsorted) evaluation on tensor-product leaves, and the refresh of the
// per-leaf value blocks.  See DeviceModel::tl_block and loco_basis.cuh (tensor_eval_scalar).
//
// One thread per query: locate, then read the leaf's block (N*F centres, N supports, F^N values) straight
// from global memory through L1/L2.  Used for small batches, for models whose blocks fit the L1, and for
// evaluation on a known cell (IN_CUBE: the refinement loop's centre / jitter points).
#include "loco_basis.cuh"
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
	if (st == ST_OK) r = tensor_eval_scalar<NT, CUBIC, EXACT>(m, m.tl_block + (int64_t)leaf * m.tl_stride, x);
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
