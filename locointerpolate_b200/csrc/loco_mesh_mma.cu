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
//     and the table of its low-dimension products low[kl][q] (kl < lowK <= 16 rows: ONE stage of value rows); the weight of
//     row k = kl + lowK * h is w = low[kl][q] * high_h[q] with high_h = prod_{d >= LOWD} f_d[digit_d(h)][q]
//     (LCSample::getInterpolationWeight on a unit-weight tensor grid).  low[][] is written ONCE per item; per stage the
//     producers wait for a free stage, write the stage's 64 high factors and (producer 0) start the TMA bulk copies of the
//     stage's value-row segments; an mbarrier per stage collects "factors written" + the TMA transaction bytes.
//     (Round 1 wrote all 16 x 64 weights of every stage: 16 dependent fp64 multiplies per producer lane and stage, queued on
//     the fp64 pipe behind the consumers' DMMAs -- the consumers waited for the producers 38 % of the time.)
//   * warps 0-7 (consumers): wait for the stage, scale their A fragments (low[][] of the item) by the stage's high factor
//     of the fragment's query -- the same product low * high the producers used to form --, run 4x4 DMMA tiles per warp
//     (32 queries x 32 components), release the stage with one arrival per warp, and after the last stage of a chunk write
//     each output element once with 128-bit streaming stores.
// Items are ordered chunk-group major and dealt round-robin to the CTAs, so that at any time all CTAs read the same
// slice of the value table (it stays L2 resident: the table is read from HBM about once per batch).
#include <algorithm>
#include <cstdlib>

#include "loco_basis.cuh"
#include "loco_device.h"
#include "loco_sort.cuh"
#include "loco_tma.cuh"

namespace loco {

namespace {
constexpr int MM_Q = 64;      // queries per tile
constexpr int MM_D = 128;     // value components per chunk
constexpr int MM_K = 16;      // value rows per stage
constexpr int MM_NS_MULTI = 4; // stages in flight when a chunk needs several stages (2 CTAs of ~100 KB per SM)
constexpr int MM_NS_ONE = 4;   // ... when all K <= 16 rows fit one stage: the weights are per item, a stage is value rows only
// !ONE: who applies the stage's high factor to the low-dimension weight products.  true: the (otherwise idle) producers write
// the stage's 16 x 64 weights w = low * high into a per-stage buffer (16 independent multiplies per lane; 3 stages fit);
// false: the consumers scale their A fragments (4 stages, but every one of the 4 warps sharing a query block repeats the
// multiplies: 128 instead of 32 warp-level DMULs per stage on the pipe the DMMAs need -- ncu r02h: 8 x 2.3 % of the samples).
// Measured equal (c4b 40.7 k vs 40.8 k q/s, c4 91.6 k vs 94.1 k): the consumer-side form with its deeper ring is kept.
constexpr bool MM_PSCALE = false;
constexpr bool MM_TMA_STORE = false;  // results through shared memory + TMA bulk stores (measured: no gain over direct 256-bit stores, costs a stage)
constexpr int MM_STG_LD = 40;  // row stride (doubles) of a warp's 8 x 32 result staging block: conflict-free 128-bit stores, 16-byte aligned rows
constexpr int MM_CONSUMERS = 8;
constexpr int MM_PRODUCERS = MM_Q / 32;  // one producer lane per query of the tile
constexpr int MM_THREADS = (MM_CONSUMERS + MM_PRODUCERS) * 32;
constexpr int MM_LDV = MM_D + 4;  // row strides = 4 (mod 16) doubles: the 4 x 8 fragment loads of a half warp hit 16 distinct 8-byte banks
constexpr int MM_LDW = MM_Q + 4;

__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

struct ItemGeom { int32_t leaf, qbeg, nq, dc0, n_dc, row0, K, lowK; };  // row0, K: the leaf's folded row list in DeviceModel::fold_row; lowK: rows per stage (!ONE)
// item -> (chunk group, tile) -> leaf and query range; identical arithmetic in the producer and in every consumer warp
__device__ __forceinline__ ItemGeom item_geometry(int64_t item, int64_t tile0, int64_t n_tiles, int32_t L, int32_t n_dchunks, int dg, const int32_t *__restrict__ offset,
                                                  const int32_t *__restrict__ tile_off, const int32_t *__restrict__ fold_off, const int32_t *__restrict__ tile_leaf)
{
	ItemGeom g;
	const int64_t dgroup = item / n_tiles, tile = tile0 + (item - dgroup * n_tiles);
	const int32_t lo = __ldg(tile_leaf + tile);  // leaf of the tile (SortWork::tile_leaf)
	g.leaf = lo;
	g.qbeg = __ldg(offset + lo) + (int32_t)(tile - __ldg(tile_off + lo)) * MM_Q;
	g.nq = min(MM_Q, __ldg(offset + lo + 1) - g.qbeg);
	g.dc0 = (int32_t)dgroup * dg;
	g.n_dc = min(dg, n_dchunks - g.dc0);
	g.row0 = __ldg(fold_off + lo);
	g.K = __ldg(fold_off + lo + 1) - g.row0;
	g.lowK = MM_K;
	return g;
}
} // namespace

template <int NT, bool CUBIC, bool EXACT, bool ONE>
__global__ void __launch_bounds__(MM_THREADS, 2) mesh_tensor_mma_kernel(DeviceModel m, const double *__restrict__ xs, const int32_t *__restrict__ offset,
                                                                         const int32_t *__restrict__ tile_off, const int32_t *__restrict__ tile_leaf,
                                                                         double *__restrict__ out,
                                                                         int64_t out_stride, int32_t n_dchunks, int32_t dg, int64_t tile0, int64_t n_tiles_arg,
                                                                         int64_t ring_base, int dbg)
{
	constexpr int F = CUBIC ? 4 : 2;
	constexpr int LOGF = CUBIC ? 2 : 1;
	constexpr int RS = ((NT * 8 + 4 + 31) / 32) * 4;
	// the low LOWD dimensions of the tensor index are tabulated per item (LOWK <= 16 products per query); the remaining
	// "high" product is constant over runs of LOWK consecutive rows
	constexpr int LOWD = NT < (CUBIC ? 2 : 4) ? NT : (CUBIC ? 2 : 4);
	constexpr int LOWK = 1 << (LOGF * LOWD);
	constexpr bool PSCALE = MM_PSCALE && !ONE;
	constexpr int MM_NS = ONE ? MM_NS_ONE : (PSCALE ? 3 : MM_NS_MULTI);
	constexpr int MM_NW = PSCALE ? MM_NS : 2;  // weight buffers: per stage (PSCALE), else per item (all rows of a chunk when ONE, the low-dimension products otherwise)
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double *sV = reinterpret_cast<double *>(smem_raw);             // [MM_NS][MM_K][MM_LDV]
	double *sW = sV + MM_NS * MM_K * MM_LDV;                       // [MM_NW][MM_K][MM_LDW]
	double *sF = sW + MM_NW * MM_K * MM_LDW;                       // [NT*F][MM_Q]   producer only
	double *sH = sF + NT * F * MM_Q;                               // [MM_NS][MM_Q]  the stage's high factor of every query (!ONE, !PSCALE)
	double *sWl = sH + MM_NS * MM_Q;                               // [MM_K][MM_Q]   the item's low-dimension products, producer-private (PSCALE)
	int32_t *sQi = reinterpret_cast<int32_t *>(sWl + (PSCALE ? MM_K * MM_Q : 0)); // [2][MM_Q]  double-buffered per item: read by the consumers' epilogues
	double *sStage = reinterpret_cast<double *>(sQi + 2 * MM_Q);   // [MM_CONSUMERS][8][MM_STG_LD]  (ONE only) results on their way to TMA bulk stores
	__shared__ __align__(8) uint64_t full_bar[MM_NS], empty_bar[MM_NS], item_bar[2];
	__shared__ ItemGeom s_item[2];  // the producers locate the item's leaf (a binary search over tile offsets); the consumers read it here
	if (threadIdx.x == 0) {
		for (int i = 0; i < MM_NS; i++) { mbar_init(&full_bar[i], MM_PRODUCERS); mbar_init(&empty_bar[i], MM_CONSUMERS); }
		mbar_init(&item_bar[0], MM_CONSUMERS); mbar_init(&item_bar[1], MM_CONSUMERS);
	}
	__syncthreads();

	// tiles [tile0, tile0 + n_tiles) of the sorted batch; ring_base >= 0: results go to row (sorted position - ring_base) of
	// `out` (a ring the caller consumes in leaf order) instead of row (query index)
	const int32_t L = m.L;
	const int64_t n_tiles = n_tiles_arg >= 0 ? n_tiles_arg : (int64_t)tile_off[L];
	const int32_t n_dgroups = (n_dchunks + dg - 1) / dg;
	const int64_t total = n_tiles * n_dgroups;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	if (warp >= MM_CONSUMERS) {
		// ========================================== producer warps ===========================================
		// lane owns ONE query of the tile (q = 32 * producer index + lane); producer 0 also drives the TMA engine
		const int pw = warp - MM_CONSUMERS, q = pw * 32 + lane;
		const uint64_t keep = l2_policy_evict_last();  // the slice of the value table all CTAs share stays in L2 under the result stream
		uint32_t g = 0;   // running step counter: stage = g % MM_NS, use count = g / MM_NS
		uint32_t it = 0;  // running item counter: item buffer = it & 1
		for (int64_t item = blockIdx.x; item < total; item += gridDim.x, it++) {
			const ItemGeom ig = item_geometry(item, tile0, n_tiles, L, n_dchunks, dg, offset, tile_off, m.fold_off, tile_leaf);
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
			if (pw == 0 && lane == 0) s_item[ib] = ig;  // published by the arrival on the item's first stage barrier
			double f[NT][F];
			tensor_factors<NT, CUBIC, EXACT>(blk, x, f, m.tensor_uniform != 0);
			// ---- folded factors g[d][c] = sum of the factors of the digits in class c (DeviceModel::fold_map), kept in shared
			//      memory (column q is private to this lane) so that they can be indexed by a runtime class index
			const uint8_t *fmap = m.fold_map + (int64_t)ig.leaf * NT * F;
			const uint8_t *fnum = m.fold_n + (int64_t)ig.leaf * NT;
			int nd[NT];
#pragma unroll
			for (int d = 0; d < NT; d++) {
				nd[d] = __ldg(fnum + d);
				double gsum[F];
#pragma unroll
				for (int c = 0; c < F; c++) gsum[c] = 0.0;
#pragma unroll
				for (int j = 0; j < F; j++) {
					const int cls = __ldg(fmap + d * F + j);
#pragma unroll
					for (int c = 0; c < F; c++) gsum[c] += (cls == c) ? f[d][j] : 0.0;
				}
#pragma unroll
				for (int c = 0; c < F; c++) sF[(d * F + c) * MM_Q + q] = q < ig.nq ? gsum[c] : 0.0;
			}
			// ---- products over the low dimensions (mixed radix nd[0..LOWD)): W[kl][q], kl < lowK <= 16, written ONCE per item into
			//      the item's weight buffer (rows lowK .. MM_K-1 are zero: stages are padded to whole MMA k-steps)
			int lowK = 1;
#pragma unroll
			for (int d = 0; d < LOWD; d++) lowK *= nd[d];
			{
				double *W_ = PSCALE ? sWl + q : sW + (size_t)ib * MM_K * MM_LDW + q;
				constexpr int W_LD = PSCALE ? MM_Q : MM_LDW;
				int dig[LOWD];
#pragma unroll
				for (int d = 0; d < LOWD; d++) dig[d] = 0;
				for (int kl = 0; kl < MM_K; kl++) {
					double w = 0.0;
					if (kl < lowK) {
						w = 1.0;
#pragma unroll
						for (int d = 0; d < LOWD; d++) w *= sF[(d * F + dig[d]) * MM_Q + q];
#pragma unroll
						for (int d = 0; d < LOWD; d++) {  // increment the mixed-radix counter
							if (++dig[d] < nd[d]) break;
							dig[d] = 0;
						}
					}
					W_[kl * W_LD] = w;
				}
			}
			const int32_t K = ig.K;
			// ONE: all K <= 16 rows are one stage.  Otherwise a stage is one run of lowK rows sharing their high digits.
			const int rows_per_stage = ONE ? MM_K : lowK;
			const int n_stages = ONE ? 1 : K / lowK;
			if (pw == 0 && lane == 0) s_item[ib].lowK = rows_per_stage;
			const int32_t *rows = m.fold_row + ig.row0;
			const int n_steps = ig.n_dc * n_stages;
			// everything a step needs that does not depend on the stage being free is prepared BEFORE waiting for it:
			// the row index of this lane's TMA copy (a global load) and the 16 weights (kept in registers).
			// Rows are padded to a multiple of 4 per stage with zero weights (the MMA consumes 4 rows at a time); the
			// padding lanes copy the leaf's last row so that 0 * value stays finite.
			const int kc_all = ONE ? min(MM_K, K) : lowK;
			const int kc4 = (kc_all + 3) & ~3;
			auto row_of = [&](int c, int ln) -> int32_t {
				return ln < kc4 ? __ldg(rows + min(c * rows_per_stage + min(ln, kc_all - 1), K - 1)) : 0;
			};
			int32_t my_row = row_of(0, lane);
			int hd[NT > LOWD ? NT - LOWD : 1];
			double ph = 1.0;
			for (int s = 0; s < n_steps; s++, g++) {
				const int dc = s / n_stages, c = s - dc * n_stages;
				const int st = g % MM_NS;
				if (!ONE) {
					// the stage's high factor: mixed-radix digits of stage c over the high dimensions, carried from stage to stage
					if (c == 0) {
#pragma unroll
						for (int d = LOWD; d < NT; d++) hd[d - LOWD] = 0;
					} else {
#pragma unroll
						for (int d = LOWD; d < NT; d++) {
							if (++hd[d - LOWD] < nd[d]) break;
							hd[d - LOWD] = 0;
						}
					}
					ph = 1.0;
#pragma unroll
					for (int d = LOWD; d < NT; d++) ph *= sF[(d * F + hd[d - LOWD]) * MM_Q + q];
				}
				double wreg[PSCALE ? MM_K : 1];
				if constexpr (PSCALE) {
#pragma unroll
					for (int kk = 0; kk < MM_K; kk++) wreg[kk] = sWl[kk * MM_Q + q] * ph;  // rows >= lowK are zero
				}
				const int32_t row = my_row;
				if (pw == 0 && s + 1 < n_steps) my_row = row_of((c + 1 == n_stages) ? 0 : c + 1, lane);  // row index of the NEXT step's copy
				mbar_wait(&empty_bar[st], ((g / MM_NS) & 1u) ^ 1u);
				if constexpr (PSCALE) {
					double *Ws = sW + (size_t)st * MM_K * MM_LDW + q;
#pragma unroll
					for (int kk = 0; kk < MM_K; kk++)
						if (kk < kc4) Ws[kk * MM_LDW] = wreg[kk];
				} else if (!ONE) sH[st * MM_Q + q] = ph;
				__syncwarp();  // every lane's weights / factor are written before lane 0 arrives on the stage barrier
				if (pw == 0) {
					// ---- the stage's value rows: one TMA bulk copy per row segment
					const int32_t d0 = (ig.dc0 + dc) * MM_D;
					const uint32_t row_bytes = (uint32_t)(((min(MM_D, m.D - d0) + 1) & ~1) * 8);  // rows are padded to an even number of doubles
					if (lane == 0) mbar_expect_tx(&full_bar[st], row_bytes * (uint32_t)kc4);
					__syncwarp();
					if (lane < kc4)
					{
						if (dbg & 1) tma_bulk_g2s(sV + ((size_t)st * MM_K + lane) * MM_LDV, m.values + (int64_t)row * m.value_stride + d0, row_bytes, &full_bar[st]);  // EXPERIMENT: no L2 hint
						else tma_bulk_g2s_hint(sV + ((size_t)st * MM_K + lane) * MM_LDV, m.values + (int64_t)row * m.value_stride + d0, row_bytes, &full_bar[st], keep);
					}
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
	const bool wide_out = ((out_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 31) == 0);  // rows 32-byte aligned
	const uint64_t wr_policy = l2_policy_evict_first();
	uint32_t g = 0, it = 0;
	double acc[4][4][2];  // [query block mi][component block ni][2]
	for (int64_t item = blockIdx.x; item < total; item += gridDim.x, it++) {
		const int ib = it & 1;
		// the first stage of the item being full means its geometry (and query indices, weights) are in shared memory
		mbar_wait(&full_bar[g % MM_NS], (g / MM_NS) & 1u);
		const ItemGeom ig = s_item[ib];
		const int32_t K = ig.K;
		const int n_stages = ONE ? 1 : K / ig.lowK;
		const int n_steps = ig.n_dc * n_stages;
		// the tile's 8-query blocks are dealt alternately to the two warp rows (block b = 2*mi + wq): a partially filled tile
		// (the last tile of most leaves) keeps BOTH rows busy instead of leaving warps 4-7 idle below 33 queries; blocks beyond
		// the tile's last query hold only padding and are skipped
		const int n_mi = max(0, min(4, (((ig.nq + 7) >> 3) - wq + 1) >> 1));
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
			const int kc = ((ONE ? min(MM_K, K) : ig.lowK) + 3) & ~3;  // padded to whole MMA k-steps with zero weights
			const double *w_base = sW + (size_t)(PSCALE ? st : ib) * MM_K * MM_LDW + wq * 8 + fr;  // A[fr][fc] = W[k0 + fc][q0 + fr], q0 = (2*mi + wq) * 8
			const double *v_base = sV + (size_t)st * MM_K * MM_LDV + wd * 32 + fr;  // B[fc][fr] = V[k0 + fc][d0 + fr]
			if (n_mi > 0) {
				double hq[4];  // the stage's high factor of this lane's four fragment rows (queries)
#pragma unroll
				for (int mi = 0; mi < 4; mi++) hq[mi] = (ONE || PSCALE) ? 1.0 : sH[st * MM_Q + (2 * mi + wq) * 8 + fr];
#pragma unroll 2
				for (int k0 = 0; k0 < kc; k0 += 4) {
					double a[4], bf[4];
#pragma unroll
					for (int mi = 0; mi < 4; mi++) {
						a[mi] = w_base[(k0 + fc) * MM_LDW + mi * 16];
						if (!ONE && !PSCALE) a[mi] *= hq[mi];  // w = low * high: the product the reference's weight is
					}
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
				if (MM_TMA_STORE && ONE && vec_out && nd == MM_D) {
					// One stage per chunk means an epilogue per stage: hand the results to the TMA engine instead of issuing
					// 16 global stores per thread.  Per 8-query block the warp parks its 8 x 32 fragment block in shared memory
					// and lanes 0-7 each start one 256-byte bulk store of a row; the block is reused once those reads are done.
					double *stg = sStage + warp * (8 * MM_STG_LD);
#pragma unroll
					for (int mi = 0; mi < 4; mi++) {
						if (lane < 8) tma_bulk_wait_read0();
						__syncwarp();
#pragma unroll
						for (int ni = 0; ni < 4; ni++) *reinterpret_cast<double2 *>(stg + fr * MM_STG_LD + ni * 8 + 2 * fc) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
						fence_async_smem();
						__syncwarp();
						if (lane < 8) {
							const int32_t qi = sQi[ib * MM_Q + (2 * mi + wq) * 8 + lane];
							if (qi >= 0) {
								const int64_t orow = ring_base >= 0 ? (int64_t)ig.qbeg + (2 * mi + wq) * 8 + lane - ring_base : (int64_t)qi;
								tma_bulk_s2g(out + orow * out_stride + d0 + wd * 32, stg + lane * MM_STG_LD, 32 * 8);
							}
							tma_bulk_commit();
						}
					}
				} else if (wide_out && nd == MM_D) {
					// 256-bit stores of FULL 128-byte lines: a lane pair (fc, fc^1) swaps one 8x8 fragment row half so that each
					// lane owns 4 consecutive components of one row; the 4 lanes of a row then cover 16 consecutive doubles
					const bool odd = fc & 1;
#pragma unroll
					for (int mi = 0; mi < 4; mi++) {
						const int32_t qi = sQi[ib * MM_Q + (2 * mi + wq) * 8 + fr];
						const int64_t orow = ring_base >= 0 ? (int64_t)ig.qbeg + (2 * mi + wq) * 8 + fr - ring_base : (int64_t)qi;
						double *o = out + orow * out_stride + d0 + wd * 32;
#pragma unroll
						for (int p = 0; p < 2; p++) {
							const int n0 = 2 * p, n1 = 2 * p + 1;
							const double s0 = odd ? acc[mi][n0][0] : acc[mi][n1][0], s1 = odd ? acc[mi][n0][1] : acc[mi][n1][1];
							const double r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
							if (qi < 0) continue;
							// 256-bit stores with an explicit L2 evict_first POLICY (st.global.L2::cache_hint): measured 5 % faster on c3 than
							// the .cs qualifier (17.9 -> 17.0 ms; plain write-back 17.2, evict_unchanged 17.4; LOCO_MESH_DBG 8 / 2 / 6 select them)
							double *po = odd ? o + n1 * 8 + 2 * (fc - 1) : o + n0 * 8 + 2 * fc;
							const double v0 = odd ? r0 : acc[mi][n0][0], v1 = odd ? r1 : acc[mi][n0][1], v2 = odd ? acc[mi][n1][0] : r0, v3 = odd ? acc[mi][n1][1] : r1;
							if ((dbg & 14) == 0) st_policy_f64x4(po, v0, v1, v2, v3, wr_policy);
							else if (dbg & 8) st_stream_f64x4(po, v0, v1, v2, v3);
							else if ((dbg & 6) == 2) st_plain_f64x4(po, v0, v1, v2, v3);
							else st_policy_f64x4(po, v0, v1, v2, v3, l2_policy_no_allocate());
						}
					}
				} else {
#pragma unroll
				for (int mi = 0; mi < 4; mi++) {
					const int32_t qi = sQi[ib * MM_Q + (2 * mi + wq) * 8 + fr];
					if (qi < 0) continue;
					const int64_t orow = ring_base >= 0 ? (int64_t)ig.qbeg + (2 * mi + wq) * 8 + fr - ring_base : (int64_t)qi;
					double *o = out + orow * out_stride + d0;
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
		}
		__syncwarp();
		if (lane == 0) mbar_arrive(&item_bar[ib]);  // sQi / factors of this item buffer may be overwritten
	}
}

namespace {
size_t mma_smem_bytes(int N, int F, int lowk, bool one)
{
	const bool pscale = MM_PSCALE && !one;
	const int ns = one ? MM_NS_ONE : (pscale ? 3 : MM_NS_MULTI), nw = pscale ? ns : 2;
	(void)lowk;
	return (size_t)(ns * MM_K * MM_LDV + nw * MM_K * MM_LDW + N * F * MM_Q + ns * MM_Q + (pscale ? MM_K * MM_Q : 0)) * sizeof(double) + 2 * MM_Q * sizeof(int32_t) + 64 +
	       ((one && MM_TMA_STORE) ? (size_t)MM_CONSUMERS * 8 * MM_STG_LD * sizeof(double) : 0);
}
template <int NT, bool CUBIC, bool EXACT, bool ONE>
void run_mma1(const DeviceModel &m, const SortWork &w, double *out, int64_t out_stride, int n_dchunks, int dg, int grid, int64_t tile0, int64_t n_tiles,
              int64_t ring_base, cudaStream_t st)
{
	auto kern = mesh_tensor_mma_kernel<NT, CUBIC, EXACT, ONE>;
	constexpr int LOWD = NT < (CUBIC ? 2 : 4) ? NT : (CUBIC ? 2 : 4);
	const size_t smem = mma_smem_bytes(NT, CUBIC ? 4 : 2, 1 << ((CUBIC ? 2 : 1) * LOWD), ONE);
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	LOCO_KERNEL("mesh_tensor_mma", st);
	static const int dbg = getenv("LOCO_MESH_DBG") ? atoi(getenv("LOCO_MESH_DBG")) : 0;  // measurement experiments only
	static const int grid_env = getenv("LOCO_MESH_GRID") ? atoi(getenv("LOCO_MESH_GRID")) : 0;
	if (grid_env > 0) grid = grid_env;
	kern<<<grid, MM_THREADS, smem, st>>>(m, w.xs, w.offset, w.tile_off, w.tile_leaf, out, out_stride, n_dchunks, dg, tile0, n_tiles, ring_base, dbg);
}
template <int NT, bool CUBIC, bool EXACT>
void run_mma(const DeviceModel &m, const SortWork &w, double *out, int64_t out_stride, int n_dchunks, int dg, int grid, int64_t tile0, int64_t n_tiles,
             int64_t ring_base, cudaStream_t st)
{
	// K = F^NT <= 16 rows (one stage per chunk) only for small NT; the test is a compile-time constant per instantiation
	constexpr int KT = 1 << ((CUBIC ? 2 : 1) * NT);
	if constexpr (KT <= MM_K) run_mma1<NT, CUBIC, EXACT, true>(m, w, out, out_stride, n_dchunks, dg, grid, tile0, n_tiles, ring_base, st);
	else run_mma1<NT, CUBIC, EXACT, false>(m, w, out, out_stride, n_dchunks, dg, grid, tile0, n_tiles, ring_base, st);
}
} // namespace

int mesh_mma_tile_queries() { return MM_Q; }
bool mesh_mma_supported(const DeviceModel &m) { return m.tensor_F > 0 && m.N >= 1 && m.N <= kMaxTemplateN; }

// queries must already be sorted by leaf with tiles of mesh_mma_tile_queries() (launch_sort_by_leaf)
// tile0 / n_tiles: the range of sorted tiles to evaluate (n_tiles < 0: all of them, read from the device);
// ring_base >= 0: write row (sorted position - ring_base) instead of row (query index)
int launch_mesh_tensor_mma(const DeviceModel &m, const SortWork &w, int64_t Q, double *out, int64_t out_stride, int sm_count, int64_t tile0, int64_t n_tiles,
                           int64_t ring_base, cudaStream_t st)
{
	if (Q <= 0 || n_tiles == 0) return 0;
	if (n_tiles < 0) tile0 = 0;  // the kernel reads the tile count from the device (no host synchronisation)
	const int n_dchunks = (m.D + MM_D - 1) / MM_D;
	// chunks per item: enough steps per item to amortise the per-item tables, while the slice of the value table all
	// CTAs share at a time (rows x dg*128 components) stays well inside the L2
	const int n_stages = (m.tensor_K + MM_K - 1) / MM_K;
	// >= 32 steps per item where the slice allows it (items with few stages need several chunks to amortise their tables),
	// a slice of at most ~24 MB (ncu r02h: the 62 MB slice of the 6-D cubic table missed the L2 on 75 % of its re-reads)
	static const int dg_env = getenv("LOCO_MESH_DG") ? atoi(getenv("LOCO_MESH_DG")) : 0;  // measurement experiments
	int dg = std::max(1, (32 + n_stages - 1) / n_stages);
	const int64_t slice_cap = ((int64_t)24 << 20) / std::max<int64_t>(1, m.R * MM_D * 8);
	dg = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::max(dg, 4), std::max<int64_t>(slice_cap, 1), (int64_t)n_dchunks}));
	if (n_stages < 8) dg = std::max(dg, std::min(4, n_dchunks));  // one-stage chunks: never fewer than 4 chunks per item
	if (dg_env > 0) dg = std::min(dg_env, n_dchunks);
	const int grid = sm_count * 2;
	const bool cubic = m.basis == CUBIC_BSPLINE, exact = m.exact_rcp != 0;
#define LOCO_MCASE(NT)                                                                                                  \
	case NT:                                                                                                            \
		if (cubic) { if (exact) run_mma<NT, true, true>(m, w, out, out_stride, n_dchunks, dg, grid, tile0, n_tiles, ring_base, st); else run_mma<NT, true, false>(m, w, out, out_stride, n_dchunks, dg, grid, tile0, n_tiles, ring_base, st); } \
		else { if (exact) run_mma<NT, false, true>(m, w, out, out_stride, n_dchunks, dg, grid, tile0, n_tiles, ring_base, st); else run_mma<NT, false, false>(m, w, out, out_stride, n_dchunks, dg, grid, tile0, n_tiles, ring_base, st); }     \
		break;
	switch (m.N) { LOCO_MCASE(1) LOCO_MCASE(2) LOCO_MCASE(3) LOCO_MCASE(4) LOCO_MCASE(5) LOCO_MCASE(6) LOCO_MCASE(7) LOCO_MCASE(8) default: return 0; }
#undef LOCO_MCASE
	return 1;
}

} // namespace loco
