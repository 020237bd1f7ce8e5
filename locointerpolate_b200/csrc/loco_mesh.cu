// loco_mesh.cu -- mesh-valued evaluation on tensor-product leaves: the weighted gather-sum
//
//     out[q][:] = sum_k w_k(x_q) * values[row_k(leaf(q))][:]          (LCFunctionValue::newFromWeightedSum with
//                                                                      LCRealFunctionValue::add applied per component)
//
// as a tiled fp64 FMA kernel.  Queries arrive counting-sorted by leaf (loco_tiles.cu), so a CTA works on a
// tile of up to 64 queries of ONE leaf times a chunk of 128 value components:
//   * the leaf's K = F^N value rows are streamed through shared memory in stages of 16 rows by the TMA
//     engine (one cp.async.bulk per row segment, double-buffered, mbarrier completion);
//   * the weights of a stage are formed on the fly from the tile's N*F one-dimensional factors
//     (w_k = prod_d f_d[j_d(k)], LCSample::getInterpolationWeight on a unit-weight tensor grid), so no
//     [Q x K] weight matrix ever goes through HBM;
//   * every thread owns an 8-query x 4-component register tile (32 DFMA per shared-memory k step).
// Work items = (group of 4 component chunks, query tile), ordered chunk-group major and dealt round-robin
// to the CTAs, so that at any time all CTAs read the same 512-component slice of the value table (it stays
// L2 resident: the table is read from HBM about once per batch) and each output element is written once.
// No tensor cores: fp64 DFMA is the roofline here (2*K*D flops per query).
#include <algorithm>

#include "loco_basis.cuh"
#include "loco_device.h"
#include "loco_sort.cuh"
#include "loco_tma.cuh"

namespace loco {

namespace {
constexpr int MT_Q = 64;     // queries per tile
constexpr int MT_D = 128;    // value components per chunk
constexpr int MT_K = 32;     // value rows per stage
constexpr int MT_DG = 4;     // consecutive component chunks handled per work item (factors are computed once per item)
constexpr int MT_THREADS = 256;
}

template <int NT, bool CUBIC, bool EXACT>
__global__ void __launch_bounds__(MT_THREADS, 2) mesh_tensor_tiles_kernel(DeviceModel m, const double *__restrict__ xs, const int32_t *__restrict__ offset,
                                                                           const int32_t *__restrict__ tile_off, double *__restrict__ out,
                                                                           int64_t out_stride, int32_t n_dchunks)
{
	constexpr int F = CUBIC ? 4 : 2;
	constexpr int LOGF = CUBIC ? 2 : 1;
	constexpr int RS = ((NT * 8 + 4 + 31) / 32) * 4;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double *sV = reinterpret_cast<double *>(smem_raw);            // [2][MT_K][MT_D]
	double *sW = sV + 2 * MT_K * MT_D;                            // [2][MT_K][MT_Q]
	double *sF = sW + 2 * MT_K * MT_Q;                            // [NT*F][MT_Q]
	int32_t *sQi = reinterpret_cast<int32_t *>(sF + NT * F * MT_Q);  // [MT_Q]
	__shared__ __align__(8) uint64_t bar[2];
	if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
	__syncthreads();
	uint32_t phases = 0;

	const int32_t L = m.L, K = m.tensor_K;
	const int64_t n_tiles = tile_off[L];
	const int32_t n_dgroups = (n_dchunks + MT_DG - 1) / MT_DG;
	const int64_t total = n_tiles * n_dgroups;
	const int tq = threadIdx.x >> 5, td = threadIdx.x & 31;
	const bool vec_out = ((out_stride & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
	const int n_stages = (K + MT_K - 1) / MT_K;

	for (int64_t item = blockIdx.x; item < total; item += gridDim.x) {
		const int64_t dgroup = item / n_tiles, tile = item - dgroup * n_tiles;
		// leaf of this tile: last leaf whose first tile is <= tile
		int32_t lo = 0, hi = L;
		while (lo < hi) {
			const int32_t mid = (lo + hi) >> 1;
			if (__ldg(tile_off + mid + 1) <= tile) lo = mid + 1; else hi = mid;
		}
		const int32_t leaf = lo;
		const int32_t qbeg = __ldg(offset + leaf) + (int32_t)(tile - __ldg(tile_off + leaf)) * MT_Q;
		const int32_t nq = min(MT_Q, __ldg(offset + leaf + 1) - qbeg);
		const int32_t *rows = m.tl_row + (int64_t)leaf * K;

		// ---- the tile's one-dimensional factors: sF[d*F + j][q] = b((x_q[d] - c_d[j]) / s_d)
		if (threadIdx.x < MT_Q) {
			const int q = threadIdx.x;
			Coord<NT> x;
#pragma unroll
			for (int d = 0; d < NT; d++) x.set(d, 0.0);
			int32_t qi = -1;
			if (q < nq) qi = load_sorted_row<NT>(xs + (int64_t)(qbeg + q) * RS, x, NT);
			sQi[q] = qi;
			const double *blk = m.tl_block + (int64_t)leaf * m.tl_stride;
			double f[NT][F];
			tensor_factors<NT, CUBIC, EXACT>(blk, x, f, m.tensor_uniform != 0);
#pragma unroll
			for (int d = 0; d < NT; d++)
#pragma unroll
				for (int j = 0; j < F; j++) sF[(d * F + j) * MT_Q + q] = q < nq ? f[d][j] : 0.0;
		}
		__syncthreads();

		// ---- flat sequence of steps = (component chunk, row stage); the TMA copies and the weights of step s+1
		//      are produced while step s computes, across chunk boundaries too
		const int32_t dc0 = (int32_t)dgroup * MT_DG;
		const int n_dc = min(MT_DG, n_dchunks - dc0);
		const int n_steps = n_dc * n_stages;
		// issued by warp 0: lane 0 arms the barrier, then every lane starts the copy of one value row segment
		auto issue_step = [&](int s) {
			const int dc = s / n_stages, c = s - dc * n_stages, b = s & 1;
			const int32_t d0 = (dc0 + dc) * MT_D;
			const uint32_t row_bytes = (uint32_t)(((min(MT_D, m.D - d0) + 1) & ~1) * 8);  // rows are padded to an even number of doubles
			const int kc = min(MT_K, K - c * MT_K);
			if (td == 0) mbar_expect_tx(&bar[b], row_bytes * (uint32_t)kc);
			__syncwarp();
			if (td < kc)
				tma_bulk_g2s(sV + ((size_t)b * MT_K + td) * MT_D, m.values + (int64_t)__ldg(rows + c * MT_K + td) * m.value_stride + d0, row_bytes, &bar[b]);
		};
		// weights of a step into sW[b]: w[kk][q] = prod_d sF[d*F + j_d(k)][q], k = c*MT_K + kk, dimension 0 fastest
		auto make_weights = [&](int s) {
			const int c = s % n_stages, b = s & 1;
			const int kc = min(MT_K, K - c * MT_K);
			for (int e = threadIdx.x; e < kc * MT_Q; e += MT_THREADS) {
				const int kk = e / MT_Q, q = e - kk * MT_Q;
				int k = c * MT_K + kk;
				double w = 1.0;
#pragma unroll
				for (int d = 0; d < NT; d++) {
					w *= sF[(d * F + (k & (F - 1))) * MT_Q + q];
					k >>= LOGF;
				}
				sW[((size_t)b * MT_K + kk) * MT_Q + q] = w;
			}
		};
		if (tq == 0) issue_step(0);
		make_weights(0);
		__syncthreads();

		double acc[8][4];
		for (int s = 0; s < n_steps; s++) {
			const int dc = s / n_stages, c = s - dc * n_stages, b = s & 1;
			if (s + 1 < n_steps) {
				if (tq == 0) issue_step(s + 1);  // buffer b^1 was released by the barrier that ended step s-1
				make_weights(s + 1);
			}
			if (c == 0) {
#pragma unroll
				for (int i = 0; i < 8; i++)
#pragma unroll
					for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
			}
			mbar_wait(&bar[b], (phases >> b) & 1u);
			phases ^= 1u << b;
			const int kc = min(MT_K, K - c * MT_K);
			const double *w_base = sW + (size_t)b * MT_K * MT_Q + tq * 8;
			const double *v_base = sV + (size_t)b * MT_K * MT_D + td * 2;
#pragma unroll 4
			for (int kk = 0; kk < kc; kk++) {
				const double2 w01 = *reinterpret_cast<const double2 *>(w_base + kk * MT_Q);
				const double2 w23 = *reinterpret_cast<const double2 *>(w_base + kk * MT_Q + 2);
				const double2 w45 = *reinterpret_cast<const double2 *>(w_base + kk * MT_Q + 4);
				const double2 w67 = *reinterpret_cast<const double2 *>(w_base + kk * MT_Q + 6);
				const double2 va = *reinterpret_cast<const double2 *>(v_base + kk * MT_D);
				const double2 vb = *reinterpret_cast<const double2 *>(v_base + kk * MT_D + 64);
				const double w[8] = {w01.x, w01.y, w23.x, w23.y, w45.x, w45.y, w67.x, w67.y};
#pragma unroll
				for (int i = 0; i < 8; i++) {
					acc[i][0] = fma(w[i], va.x, acc[i][0]);
					acc[i][1] = fma(w[i], va.y, acc[i][1]);
					acc[i][2] = fma(w[i], vb.x, acc[i][2]);
					acc[i][3] = fma(w[i], vb.y, acc[i][3]);
				}
			}
			if (c == n_stages - 1) {
				// ---- epilogue of a chunk: each output element is written exactly once, 16-byte stores where alignment allows
				const int32_t d0 = (dc0 + dc) * MT_D;
				const int32_t nd = min(MT_D, m.D - d0);
#pragma unroll
				for (int i = 0; i < 8; i++) {
					const int32_t qi = sQi[tq * 8 + i];
					if (qi < 0) continue;
					double *o = out + (int64_t)qi * out_stride + d0;
#pragma unroll
					for (int h = 0; h < 2; h++) {
						const int col = td * 2 + 64 * h;
						if (vec_out && col + 1 < nd) {
							__stcs(reinterpret_cast<double2 *>(o + col), make_double2(acc[i][2 * h], acc[i][2 * h + 1]));
						} else {
							if (col < nd) o[col] = acc[i][2 * h];
							if (col + 1 < nd) o[col + 1] = acc[i][2 * h + 1];
						}
					}
				}
			}
			__syncthreads();  // sV[b], sW[b] free again; sW[b^1] of the next step visible; after the last step: sQi / sF free
		}
	}
}

// rows of queries the reference answers with an LCError get NaN in every component
__global__ void __launch_bounds__(256) mesh_nan_rows_kernel(const uint8_t *__restrict__ status, int64_t Q, int32_t D, double *__restrict__ out, int64_t out_stride)
{
	const int64_t qi = blockIdx.x;
	if (status[qi] == ST_OK) return;
	const double nan = __longlong_as_double(0x7ff8000000000000LL);
	for (int j = threadIdx.x; j < D; j += blockDim.x) out[qi * out_stride + j] = nan;
	(void)Q;
}

int launch_mesh_nan_rows(const DeviceModel &m, const uint8_t *status, int64_t Q, double *out, int64_t out_stride, cudaStream_t st)
{
	if (Q <= 0) return 0;
	LOCO_KERNEL("mesh_nan_rows", st);
	mesh_nan_rows_kernel<<<(unsigned)Q, 256, 0, st>>>(status, Q, m.D, out, out_stride);
	return 1;
}

// ring consumer: per-query checksum of results that sit in LEAF ORDER in a ring (row = sorted position - p0); the query
// index comes from the sorted record
__global__ void __launch_bounds__(256) checksum_sorted_kernel(const double *__restrict__ ring, int64_t ring_stride, int32_t D, const double *__restrict__ xs,
                                                               int32_t rs, int64_t p0, double *__restrict__ checksum)
{
	__shared__ double red[8];
	const double *o = ring + (int64_t)blockIdx.x * ring_stride;
	double acc = 0.0;
	for (int j = threadIdx.x; j < D; j += blockDim.x) acc += o[j] * (double)(1 + j % 7);
	for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) {
		double s = 0.0;
		for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
		const int32_t qi = (int32_t)__double_as_longlong(xs[(p0 + blockIdx.x) * rs + rs - 1]);
		checksum[qi] = s;
	}
}
__global__ void checksum_bad_kernel(const uint8_t *__restrict__ status, int64_t Q, double *__restrict__ checksum)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < Q && status[i] != ST_OK) checksum[i] = __longlong_as_double(0x7ff8000000000000LL);
}
int launch_checksum_sorted(const DeviceModel &m, const double *ring, int64_t ring_stride, const SortWork &w, int64_t p0, int64_t n, double *checksum, cudaStream_t st)
{
	if (n <= 0) return 0;
	LOCO_KERNEL("checksum", st);
	checksum_sorted_kernel<<<(unsigned)n, 256, 0, st>>>(ring, ring_stride, m.D, w.xs, sorted_row_doubles(m.N), p0, checksum);
	return 1;
}
int launch_checksum_bad(const uint8_t *status, int64_t Q, double *checksum, cudaStream_t st)
{
	if (Q <= 0) return 0;
	checksum_bad_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, st>>>(status, Q, checksum);
	return 1;
}

size_t mesh_tiles_smem_bytes(int N, int F)
{
	return (size_t)(2 * MT_K * MT_D + 2 * MT_K * MT_Q + N * F * MT_Q) * sizeof(double) + MT_Q * sizeof(int32_t) + 64;
}
int mesh_tile_queries() { return MT_Q; }

namespace {
template <int NT, bool CUBIC, bool EXACT>
void run_mesh(const DeviceModel &m, const SortWork &w, double *out, int64_t out_stride, int n_dchunks, int grid, cudaStream_t st)
{
	auto kern = mesh_tensor_tiles_kernel<NT, CUBIC, EXACT>;
	const size_t smem = mesh_tiles_smem_bytes(NT, CUBIC ? 4 : 2);
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	LOCO_KERNEL("mesh_tensor_tiles", st);
	kern<<<grid, MT_THREADS, smem, st>>>(m, w.xs, w.offset, w.tile_off, out, out_stride, n_dchunks);
}
} // namespace

// queries must already be sorted by leaf with tiles of mesh_tile_queries() (launch_sort_by_leaf)
int launch_mesh_tensor_tiles(const DeviceModel &m, const SortWork &w, const uint8_t *status, int64_t Q, double *out, int64_t out_stride, int sm_count,
                             cudaStream_t st)
{
	if (Q <= 0) return 0;
	const int n_dchunks = (m.D + MT_D - 1) / MT_D;
	const int grid = sm_count * 2;
	const bool cubic = m.basis == CUBIC_BSPLINE, exact = m.exact_rcp != 0;
#define LOCO_MCASE(NT)                                                                                                  \
	case NT:                                                                                                            \
		if (cubic) { if (exact) run_mesh<NT, true, true>(m, w, out, out_stride, n_dchunks, grid, st); else run_mesh<NT, true, false>(m, w, out, out_stride, n_dchunks, grid, st); } \
		else { if (exact) run_mesh<NT, false, true>(m, w, out, out_stride, n_dchunks, grid, st); else run_mesh<NT, false, false>(m, w, out, out_stride, n_dchunks, grid, st); }     \
		break;
	switch (m.N) { LOCO_MCASE(1) LOCO_MCASE(2) LOCO_MCASE(3) LOCO_MCASE(4) LOCO_MCASE(5) LOCO_MCASE(6) LOCO_MCASE(7) LOCO_MCASE(8) default: return 0; }
#undef LOCO_MCASE
	return 1 + launch_mesh_nan_rows(m, status, Q, out, out_stride, st);
}

} // namespace loco
