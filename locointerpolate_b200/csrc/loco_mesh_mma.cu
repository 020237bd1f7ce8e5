// loco_mesh_mma.cu -- mesh-valued evaluation on tensor-product leaves, fp64 tensor-pipe version
//
//     out[q][:] = sum_k w_k(x_q) * values[row_k(leaf(q))][:]          (LCFunctionValue::newFromWeightedSum with
//                                                                      LCRealFunctionValue::add applied per component)
//
// Queries arrive counting-sorted by leaf (loco_tiles.cu).  Work item = (64-query tile of ONE leaf) x (group of
// 128-component chunks of the value rows).  Once the batch is sorted, the weighted gather-sum of a leaf IS a dense fp64
// contraction [Q_leaf x K].[K x D] (K = 4^N = 4,096 rows at the 6-D cubic config: 1.23 GFLOP per query),
// so the inner product runs on the fp64 MMA pipe (mma.sync.m8n8k4.f64, SASS DMMA):
//   * DFMA and DMMA have the same measured peak on B200 (36.4 vs 37.1 TFLOP/s, tools/micro/dmma_rate.cu), but a DFMA
//     register tile needs every operand delivered to every lane that uses it: the 8x4 tile of loco_mesh.cu moves
//     96 B/lane from shared memory per 32 DFMA and is bound by the 128 B/clk shared-memory return path at 56 % of the
//     fp64 peak (profiles/r01a_c4_*: 0.89 shared wavefronts per DFMA).  An MMA fragment holds each operand ONCE per
//     warp: 64 B/lane per 16 DMMA (= 128 DFMA-equivalents).  The arithmetic is still IEEE fp64 FMA.
//
// The CTA is warp-specialised; nothing in the steady state uses __syncthreads:
//   * warps 8-9 (producers, one lane per query of the tile): per item compute the query's N*F one-dimensional factors
//     and its table of low-dimension products; per step wait for a free stage, form the stage's weights
//     w[kk][q] = low[k % LOWK][q] * prod_{d >= LOWD} f_d[digit_d(k)][q] (LCSample::getInterpolationWeight on a
//     unit-weight tensor grid) in shared memory and (producer 0) start the TMA bulk copies of the stage's 16 value-row segments;
//     an mbarrier per stage collects "weights written" + the TMA transaction bytes.
//   * warps 0-7 (consumers): wait for the stage, run 4x4 DMMA tiles per warp (32 queries x 32 components), release
//     the stage with one arrival per warp, and after the last stage of a chunk write each output element once with
//     128-bit streaming stores.
// Items are ordered chunk-group major and dealt round-robin to the CTAs, so that at any time all CTAs read the same
// slice of the value table (it stays L2 resident: the table is read from HBM about once per batch).
#include <algorithm>

#include "loco_basis.cuh"
#include "loco_device.h"
#include "loco_sort.cuh"
#include "loco_tma.cuh"

namespace loco {

namespace {
constexpr int MM_Q = 64;      // queries per tile
constexpr int MM_D = 128;     // value components per chunk
constexpr int MM_K = 16;      // value rows per stage
constexpr int MM_NS = 3;      // stages in flight (2 CTAs of ~98 KB per SM)
constexpr int MM_CONSUMERS = 8;
constexpr int MM_PRODUCERS = MM_Q / 32;  // one producer lane per query of the tile
constexpr int MM_THREADS = (MM_CONSUMERS + MM_PRODUCERS) * 32;
constexpr int MM_LDV = MM_D + 4;  // row strides = 4 (mod 16) doubles: the 4 x 8 fragment loads of a half warp hit 16 distinct 8-byte banks
constexpr int MM_LDW = MM_Q + 4;

__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

struct ItemGeom { int32_t leaf, qbeg, nq, dc0, n_dc; };
// item -> (chunk group, tile) -> leaf and query range; identical arithmetic in the producer and in every consumer warp
__device__ __forceinline__ ItemGeom item_geometry(int64_t item, int64_t n_tiles, int32_t L, int32_t n_dchunks, int dg, const int32_t *__restrict__ offset,
                                                  const int32_t *__restrict__ tile_off)
{
	ItemGeom g;
	const int64_t dgroup = item / n_tiles, tile = item - dgroup * n_tiles;
	int32_t lo = 0, hi = L;  // last leaf whose first tile is <= tile
	while (lo < hi) {
		const int32_t mid = (lo + hi) >> 1;
		if (__ldg(tile_off + mid + 1) <= tile) lo = mid + 1; else hi = mid;
	}
	g.leaf = lo;
	g.qbeg = __ldg(offset + lo) + (int32_t)(tile - __ldg(tile_off + lo)) * MM_Q;
	g.nq = min(MM_Q, __ldg(offset + lo + 1) - g.qbeg);
	g.dc0 = (int32_t)dgroup * dg;
	g.n_dc = min(dg, n_dchunks - g.dc0);
	return g;
}
} // namespace

template <int NT, bool CUBIC, bool EXACT>
__global__ void __launch_bounds__(MM_THREADS, 2) mesh_tensor_mma_kernel(DeviceModel m, const double *__restrict__ xs, const int32_t *__restrict__ offset,
                                                                         const int32_t *__restrict__ tile_off, double *__restrict__ out,
                                                                         int64_t out_stride, int32_t n_dchunks, int32_t dg)
{
	constexpr int F = CUBIC ? 4 : 2;
	constexpr int LOGF = CUBIC ? 2 : 1;
	constexpr int RS = ((NT * 8 + 4 + 31) / 32) * 4;
	// the low LOWD dimensions of the tensor index are tabulated per item (LOWK <= 16 products per query); the remaining
	// "high" product is constant over runs of LOWK consecutive rows
	constexpr int LOWD = NT < (CUBIC ? 2 : 4) ? NT : (CUBIC ? 2 : 4);
	constexpr int LOWK = 1 << (LOGF * LOWD);
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double *sV = reinterpret_cast<double *>(smem_raw);             // [MM_NS][MM_K][MM_LDV]
	double *sW = sV + MM_NS * MM_K * MM_LDV;                       // [MM_NS][MM_K][MM_LDW]
	double *sF = sW + MM_NS * MM_K * MM_LDW;                       // [NT*F][MM_Q]   producer only
	double *sL = sF + NT * F * MM_Q;                               // [LOWK][MM_Q]   producer only
	int32_t *sQi = reinterpret_cast<int32_t *>(sL + LOWK * MM_Q);  // [2][MM_Q]      double-buffered per item: read by the consumers' epilogues
	__shared__ __align__(8) uint64_t full_bar[MM_NS], empty_bar[MM_NS], item_bar[2];
	if (threadIdx.x == 0) {
		for (int i = 0; i < MM_NS; i++) { mbar_init(&full_bar[i], MM_PRODUCERS); mbar_init(&empty_bar[i], MM_CONSUMERS); }
		mbar_init(&item_bar[0], MM_CONSUMERS); mbar_init(&item_bar[1], MM_CONSUMERS);
	}
	__syncthreads();

	const int32_t L = m.L, K = m.tensor_K;
	const int64_t n_tiles = tile_off[L];
	const int32_t n_dgroups = (n_dchunks + dg - 1) / dg;
	const int64_t total = n_tiles * n_dgroups;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int n_stages = (K + MM_K - 1) / MM_K;

	if (warp >= MM_CONSUMERS) {
		// ========================================== producer warps ===========================================
		// lane owns ONE query of the tile (q = 32 * producer index + lane); producer 0 also drives the TMA engine
		const int pw = warp - MM_CONSUMERS, q = pw * 32 + lane;
		uint32_t g = 0;   // running step counter: stage = g % MM_NS, use count = g / MM_NS
		uint32_t it = 0;  // running item counter: item buffer = it & 1
		for (int64_t item = blockIdx.x; item < total; item += gridDim.x, it++) {
			const ItemGeom ig = item_geometry(item, n_tiles, L, n_dchunks, dg, offset, tile_off);
			const int ib = it & 1;
			// the consumers are done with this item buffer (two items ago); a fresh barrier passes the parity-1 wait
			mbar_wait(&item_bar[ib], ((it >> 1) & 1u) ^ 1u);
			const double *blk = m.tl_block + (int64_t)ig.leaf * m.tl_stride;
			// ---- the query's one-dimensional factors f[d][j] = b((x[d] - c_d[j]) / s_d)   (0 for a padding query)
			Coord<NT> x;
#pragma unroll
			for (int d = 0; d < NT; d++) x.set(d, 0.0);
			int32_t qi = -1;
			if (q < ig.nq) qi = load_sorted_row<NT>(xs + (int64_t)(ig.qbeg + q) * RS, x, NT);
			sQi[ib * MM_Q + q] = qi;
			double f[NT][F];
			tensor_factors<NT, CUBIC, EXACT>(blk, x, f);
			// ---- products over the low dimensions, kept in shared memory (column q is private to this lane)
#pragma unroll
			for (int kl = 0; kl < LOWK; kl++) {
				double w = q < ig.nq ? 1.0 : 0.0;
				int k = kl;
#pragma unroll
				for (int d = 0; d < LOWD; d++) { w *= f[d][k & (F - 1)]; k >>= LOGF; }
				sL[kl * MM_Q + q] = w;
			}
#pragma unroll
			for (int d = LOWD; d < NT; d++)
#pragma unroll
				for (int j = 0; j < F; j++) sF[(d * F + j) * MM_Q + q] = f[d][j];
			const int32_t *rows = m.tl_row + (int64_t)ig.leaf * K;
			const int n_steps = ig.n_dc * n_stages;
			// everything a step needs that does not depend on the stage being free is prepared BEFORE waiting for it:
			// the row index of this lane's TMA copy (a global load) and the 16 weights (kept in registers)
			int32_t my_row = (lane < min(MM_K, K)) ? __ldg(rows + lane) : 0;
			for (int s = 0; s < n_steps; s++, g++) {
				const int dc = s / n_stages, c = s - dc * n_stages;
				const int st = g % MM_NS;
				const int kc = min(MM_K, K - c * MM_K);
				// ---- weights of the stage: w[kk] = low[k % LOWK] * high(k / LOWK).  With more than LOWD dimensions LOWK == MM_K, so the
				//      high product is one number per stage (high index = c); otherwise there are no high dimensions at all
				double ph = 1.0;
				if (NT > LOWD) {
					int hh = c;
#pragma unroll
					for (int d = LOWD; d < NT; d++) { ph *= sF[(d * F + (hh & (F - 1))) * MM_Q + q]; hh >>= LOGF; }
				}
				double wreg[MM_K];
#pragma unroll
				for (int kk = 0; kk < MM_K; kk++) wreg[kk] = sL[(kk & (LOWK - 1)) * MM_Q + q] * ph;
				const int32_t row = my_row;
				if (pw == 0 && s + 1 < n_steps) {  // row index of the NEXT step's copy
					const int c1 = (c + 1 == n_stages) ? 0 : c + 1;
					my_row = (lane < min(MM_K, K - c1 * MM_K)) ? __ldg(rows + c1 * MM_K + lane) : 0;
				}
				mbar_wait(&empty_bar[st], ((g / MM_NS) & 1u) ^ 1u);
				double *W_ = sW + (size_t)st * MM_K * MM_LDW + q;
#pragma unroll
				for (int kk = 0; kk < MM_K; kk++)
					if (kk < kc) W_[kk * MM_LDW] = wreg[kk];
				__syncwarp();  // every lane's weights are written before lane 0 arrives on the stage barrier
				if (pw == 0) {
					// ---- the stage's value rows: one TMA bulk copy per row segment
					const int32_t d0 = (ig.dc0 + dc) * MM_D;
					const uint32_t row_bytes = (uint32_t)(((min(MM_D, m.D - d0) + 1) & ~1) * 8);  // rows are padded to an even number of doubles
					if (lane == 0) mbar_expect_tx(&full_bar[st], row_bytes * (uint32_t)kc);
					__syncwarp();
					if (lane < kc)
						tma_bulk_g2s(sV + ((size_t)st * MM_K + lane) * MM_LDV, m.values + (int64_t)row * m.value_stride + d0, row_bytes, &full_bar[st]);
				} else if (lane == 0) {
					mbar_arrive(&full_bar[st]);
				}
			}
		}
		return;
	}

	// =============================================== consumer warps ===============================================
	const int wq = warp >> 2, wd = warp & 3;   // 2 x 4 warps: warp tile = 32 queries x 32 components = 4 x 4 MMA tiles
	const int fr = lane >> 2, fc = lane & 3;   // fragment coordinates: A[fr][fc], B[fc][fr], C[fr][2*fc .. 2*fc+1]
	const bool vec_out = ((out_stride & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
	uint32_t g = 0, it = 0;
	double acc[4][4][2];  // [query block mi][component block ni][2]
	for (int64_t item = blockIdx.x; item < total; item += gridDim.x, it++) {
		const ItemGeom ig = item_geometry(item, n_tiles, L, n_dchunks, dg, offset, tile_off);
		const int ib = it & 1;
		const int n_steps = ig.n_dc * n_stages;
		const int n_mi = min(4, (ig.nq - wq * 32 + 7) >> 3);  // 8-query blocks beyond the tile's last query hold only padding
		for (int s = 0; s < n_steps; s++, g++) {
			const int dc = s / n_stages, c = s - dc * n_stages;
			const int st = g % MM_NS;
			if (c == 0) {
#pragma unroll
				for (int mi = 0; mi < 4; mi++)
#pragma unroll
					for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
			}
			mbar_wait(&full_bar[st], (g / MM_NS) & 1u);
			const int kc = min(MM_K, K - c * MM_K);  // multiple of 4 (K is)
			const double *w_base = sW + (size_t)st * MM_K * MM_LDW + wq * 32 + fr;  // A[fr][fc] = W[k0 + fc][q0 + fr]
			const double *v_base = sV + (size_t)st * MM_K * MM_LDV + wd * 32 + fr;  // B[fc][fr] = V[k0 + fc][d0 + fr]
			if (n_mi > 0) {
#pragma unroll 2
				for (int k0 = 0; k0 < kc; k0 += 4) {
					double a[4], bf[4];
#pragma unroll
					for (int mi = 0; mi < 4; mi++) a[mi] = w_base[(k0 + fc) * MM_LDW + mi * 8];
#pragma unroll
					for (int ni = 0; ni < 4; ni++) bf[ni] = v_base[(k0 + fc) * MM_LDV + ni * 8];
#pragma unroll
					for (int mi = 0; mi < 4; mi++) {
						if (mi >= n_mi) break;
#pragma unroll
						for (int ni = 0; ni < 4; ni++) dmma_m8n8k4(acc[mi][ni], a[mi], bf[ni]);
					}
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty_bar[st]);  // this warp has read everything it needs from the stage
			if (c == n_stages - 1) {
				// ---- epilogue of a chunk: each output element is written exactly once; a fragment row is 2 consecutive doubles
				const int32_t d0 = (ig.dc0 + dc) * MM_D;
				const int32_t nd = min(MM_D, m.D - d0);
#pragma unroll
				for (int mi = 0; mi < 4; mi++) {
					const int32_t qi = sQi[ib * MM_Q + wq * 32 + mi * 8 + fr];
					if (qi < 0) continue;
					double *o = out + (int64_t)qi * out_stride + d0;
#pragma unroll
					for (int ni = 0; ni < 4; ni++) {
						const int col = wd * 32 + ni * 8 + 2 * fc;
						if (vec_out && col + 1 < nd) {
							__stcs(reinterpret_cast<double2 *>(o + col), make_double2(acc[mi][ni][0], acc[mi][ni][1]));
						} else {
							if (col < nd) o[col] = acc[mi][ni][0];
							if (col + 1 < nd) o[col + 1] = acc[mi][ni][1];
						}
					}
				}
			}
		}
		__syncwarp();
		if (lane == 0) mbar_arrive(&item_bar[ib]);  // sQi / factors of this item buffer may be overwritten
	}
}

namespace {
size_t mma_smem_bytes(int N, int F, int lowk)
{
	return (size_t)(MM_NS * MM_K * MM_LDV + MM_NS * MM_K * MM_LDW + N * F * MM_Q + lowk * MM_Q) * sizeof(double) + 2 * MM_Q * sizeof(int32_t) + 64;
}
template <int NT, bool CUBIC, bool EXACT>
void run_mma(const DeviceModel &m, const SortWork &w, double *out, int64_t out_stride, int n_dchunks, int dg, int grid, cudaStream_t st)
{
	auto kern = mesh_tensor_mma_kernel<NT, CUBIC, EXACT>;
	constexpr int LOWD = NT < (CUBIC ? 2 : 4) ? NT : (CUBIC ? 2 : 4);
	const size_t smem = mma_smem_bytes(NT, CUBIC ? 4 : 2, 1 << ((CUBIC ? 2 : 1) * LOWD));
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	LOCO_KERNEL("mesh_tensor_mma", st);
	kern<<<grid, MM_THREADS, smem, st>>>(m, w.xs, w.offset, w.tile_off, out, out_stride, n_dchunks, dg);
}
} // namespace

int mesh_mma_tile_queries() { return MM_Q; }
bool mesh_mma_supported(const DeviceModel &m) { return m.tensor_F > 0 && m.N >= 1 && m.N <= kMaxTemplateN && (m.tensor_K % 4) == 0; }

// queries must already be sorted by leaf with tiles of mesh_mma_tile_queries() (launch_sort_by_leaf)
int launch_mesh_tensor_mma(const DeviceModel &m, const SortWork &w, int64_t Q, double *out, int64_t out_stride, int sm_count, cudaStream_t st)
{
	if (Q <= 0) return 0;
	const int n_dchunks = (m.D + MM_D - 1) / MM_D;
	// chunks per item: enough steps per item to amortise the per-item tables, while the slice of the value table all
	// CTAs share at a time (rows x dg*128 components) stays well inside the L2
	const int n_stages = (m.tensor_K + MM_K - 1) / MM_K;
	int dg = std::max(4, (32 + n_stages - 1) / n_stages);
	const int64_t slice_cap = ((int64_t)48 << 20) / std::max<int64_t>(1, m.R * MM_D * 8);
	dg = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)dg, std::max<int64_t>(slice_cap, 4), (int64_t)n_dchunks}));
	const int grid = sm_count * 2;
	const bool cubic = m.basis == CUBIC_BSPLINE, exact = m.exact_rcp != 0;
#define LOCO_MCASE(NT)                                                                                                  \
	case NT:                                                                                                            \
		if (cubic) { if (exact) run_mma<NT, true, true>(m, w, out, out_stride, n_dchunks, dg, grid, st); else run_mma<NT, true, false>(m, w, out, out_stride, n_dchunks, dg, grid, st); } \
		else { if (exact) run_mma<NT, false, true>(m, w, out, out_stride, n_dchunks, dg, grid, st); else run_mma<NT, false, false>(m, w, out, out_stride, n_dchunks, dg, grid, st); }     \
		break;
	switch (m.N) { LOCO_MCASE(1) LOCO_MCASE(2) LOCO_MCASE(3) LOCO_MCASE(4) LOCO_MCASE(5) LOCO_MCASE(6) LOCO_MCASE(7) LOCO_MCASE(8) default: return 0; }
#undef LOCO_MCASE
	return 1;
}

} // namespace loco
