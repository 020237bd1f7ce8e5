// loco_mesh_gen.cu -- mesh-valued evaluation on GENERIC leaves (adaptive trees: arbitrary influencing-sample sets and
// basis-function lists per leaf), fp64 MMA tiles
//
//     out[q][:] = sum_k w_k(x_q) * values[row_k(leaf(q))][:]      w_k = LCSample::getInterpolationWeight (LCSample.cpp:76-89)
//
// Same tile machinery as loco_mesh_mma.cu (queries counting-sorted by leaf, 64-query tiles of one leaf x chunks of 128
// components, value rows streamed by TMA bulk copies through an mbarrier ring, 4x4 mma.sync.m8n8k4.f64 tiles per
// consumer warp), but the weights cannot be formed from one-dimensional factors here.  They are computed once per
// batch by `weights_sorted_kernel` -- one thread per sorted query walks its leaf's term list exactly like the scalar
// direct kernel -- into a leaf-major matrix  W[leaf][slot][query of the leaf]  in global memory, and the producer warp
// streams the 16 x 64 weight slice of every stage into shared memory with TMA bulk copies next to the value rows.
#include <algorithm>
#include <cstdlib>

#include "loco_basis.cuh"
#include "loco_device.h"
#include "loco_sort.cuh"
#include "loco_tma.cuh"

namespace loco {

namespace {
constexpr int MG_Q = 64;      // queries per tile
constexpr int MG_D = 128;     // value components per chunk
constexpr int MG_K = 16;      // value rows per stage
constexpr int MG_NS = 3;      // stages in flight
constexpr int MG_CONSUMERS = 8;
constexpr int MG_THREADS = (MG_CONSUMERS + 1) * 32;
constexpr int MG_LDV = MG_D + 4;  // row strides = 4 (mod 16) doubles: conflict-free fragment loads (see loco_mesh_mma.cu)
constexpr int MG_LDW = MG_Q + 4;

__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

struct GenGeom {
	int32_t leaf, qbeg, nq, dc0, n_dc, row0, K, qoff, npad;  // qoff: position of the tile's first query inside its leaf bucket
	int64_t woff;                                            // start of the leaf's weight block
};

// ---- per-leaf offsets of the weight blocks: woff[l] = sum_{l' < l} Kpad(l') * npad(l'),  Kpad = K rounded up to 4 (MMA k-steps),
//      npad = bucket size rounded up to 2 (16-byte rows for the bulk copies)
__global__ void __launch_bounds__(1024) weight_offsets_kernel(const int32_t *__restrict__ offset, const int32_t *__restrict__ sup_off, int32_t L,
                                                               int64_t *__restrict__ woff)
{
	__shared__ long long part[1024];
	const int per = (L + 1023) / 1024;
	const int l0 = min((int)threadIdx.x * per, L), l1 = min(l0 + per, L);
	auto block_size = [&](int l) -> long long {
		const long long K = sup_off[l + 1] - sup_off[l], n = offset[l + 1] - offset[l];
		return ((K + 3) & ~3LL) * ((n + 1) & ~1LL);
	};
	long long mine = 0;
	for (int l = l0; l < l1; l++) mine += block_size(l);
	part[threadIdx.x] = mine;
	__syncthreads();
	for (int s = 1; s < 1024; s <<= 1) {  // inclusive scan (Hillis-Steele; once per batch)
		const long long v = threadIdx.x >= s ? part[threadIdx.x - s] : 0;
		__syncthreads();
		part[threadIdx.x] += v;
		__syncthreads();
	}
	long long run = part[threadIdx.x] - mine;
	for (int l = l0; l < l1; l++) { woff[l] = run; run += block_size(l); }
	if (threadIdx.x == 1023) woff[L] = run;
}

// ---- weights of every sorted query: W[woff[leaf] + slot * npad + (position in the bucket)]
template <int NT, bool CUBIC, bool EXACT>
__global__ void __launch_bounds__(256) weights_sorted_kernel(DeviceModel m, const double *__restrict__ xs, const int32_t *__restrict__ leaf_of,
                                                              const int32_t *__restrict__ offset, const int64_t *__restrict__ woff, double *__restrict__ W)
{
	const int N = NT > 0 ? NT : m.N;
	const int RS = sorted_row_doubles(N);
	const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= __ldg(offset + m.L)) return;
	Coord<NT> x;
	const int32_t qi = load_sorted_row<NT>(xs + p * RS, x, N);
	const int32_t leaf = __ldg(leaf_of + qi);
	const int32_t lofs = __ldg(offset + leaf), n = __ldg(offset + leaf + 1) - lofs;
	const int64_t npad = (n + 1) & ~1;
	double *w = W + __ldg(woff + leaf) + (p - lofs);
	int cur = -1;
	double comb = 0.0;  // LCSample::getInterpolationWeight accumulator (LCSample.cpp:78)
	for (int t = m.term_off[leaf], t1 = m.term_off[leaf + 1]; t < t1; t++) {
		const int slot = m.term_slot[t];
		if (slot != cur) {
			if (cur >= 0) w[cur * npad] = comb;
			comb = 0.0;
			cur = slot;
		}
		comb += term_value<NT, CUBIC, EXACT>(m.term_cs + (int64_t)t * 2 * N, m.term_w[t], x, N);
	}
	if (cur >= 0) w[cur * npad] = comb;
}

__device__ __forceinline__ GenGeom gen_geometry(const DeviceModel &m, int64_t item, int64_t tile0, int64_t n_tiles, int32_t n_dchunks, int dg,
                                                const int32_t *__restrict__ offset, const int32_t *__restrict__ tile_off, const int64_t *__restrict__ woff,
                                                const int32_t *__restrict__ tile_leaf)
{
	GenGeom g;
	const int64_t dgroup = item / n_tiles, tile = tile0 + (item - dgroup * n_tiles);
	const int32_t lo = __ldg(tile_leaf + tile);  // leaf of the tile (SortWork::tile_leaf)
	g.leaf = lo;
	const int32_t lofs = __ldg(offset + lo), n = __ldg(offset + lo + 1) - lofs;
	g.qoff = (int32_t)(tile - __ldg(tile_off + lo)) * MG_Q;
	g.qbeg = lofs + g.qoff;
	g.nq = min(MG_Q, n - g.qoff);
	g.npad = (n + 1) & ~1;
	g.dc0 = (int32_t)dgroup * dg;
	g.n_dc = min(dg, n_dchunks - g.dc0);
	g.row0 = __ldg(m.sup_off + lo);
	g.K = __ldg(m.sup_off + lo + 1) - g.row0;
	g.woff = __ldg(woff + lo);
	return g;
}
} // namespace

__global__ void __launch_bounds__(MG_THREADS, 2) mesh_generic_mma_kernel(DeviceModel m, const double *__restrict__ xs, int32_t rs, const int32_t *__restrict__ offset,
                                                                          const int32_t *__restrict__ tile_off, const int32_t *__restrict__ tile_leaf,
                                                                          const int64_t *__restrict__ woff,
                                                                          const double *__restrict__ W, double *__restrict__ out, int64_t out_stride,
                                                                          int32_t n_dchunks, int32_t dg, int64_t tile0, int64_t n_tiles_arg, int64_t ring_base)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double *sV = reinterpret_cast<double *>(smem_raw);             // [MG_NS][MG_K][MG_LDV]
	double *sW = sV + MG_NS * MG_K * MG_LDV;                       // [MG_NS][MG_K][MG_LDW]
	int32_t *sQi = reinterpret_cast<int32_t *>(sW + MG_NS * MG_K * MG_LDW);  // [2][MG_Q]
	__shared__ __align__(8) uint64_t full_bar[MG_NS], empty_bar[MG_NS], item_bar[2];
	__shared__ GenGeom s_item[2];
	if (threadIdx.x == 0) {
		for (int i = 0; i < MG_NS; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], MG_CONSUMERS); }
		mbar_init(&item_bar[0], MG_CONSUMERS); mbar_init(&item_bar[1], MG_CONSUMERS);
	}
	__syncthreads();
	const int64_t n_tiles = n_tiles_arg >= 0 ? n_tiles_arg : (int64_t)tile_off[m.L];
	const int32_t n_dgroups = (n_dchunks + dg - 1) / dg;
	const int64_t total = n_tiles * n_dgroups;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	if (warp == MG_CONSUMERS) {
		// =========================================== producer warp ===========================================
		const uint64_t keep = l2_policy_evict_last();
		uint32_t g = 0, it = 0;
		for (int64_t item = blockIdx.x; item < total; item += gridDim.x, it++) {
			const GenGeom ig = gen_geometry(m, item, tile0, n_tiles, n_dchunks, dg, offset, tile_off, woff, tile_leaf);
			const int ib = it & 1;
			mbar_wait(&item_bar[ib], ((it >> 1) & 1u) ^ 1u);  // the consumers are done with this item buffer (two items ago)
#pragma unroll
			for (int h = 0; h < MG_Q / 32; h++) {
				const int q = lane + 32 * h;
				sQi[ib * MG_Q + q] = q < ig.nq ? (int32_t)__double_as_longlong(xs[(int64_t)(ig.qbeg + q) * rs + rs - 1]) : -1;
			}
			if (lane == 0) s_item[ib] = ig;
			__syncwarp();
			const int32_t K = ig.K;
			const int n_stages = (K + MG_K - 1) / MG_K;
			const int32_t *rows = m.sup_row + ig.row0;
			const int n_steps = ig.n_dc * n_stages;
			const uint32_t w_bytes = (uint32_t)(((ig.nq + 1) & ~1) * 8);
			for (int s = 0; s < n_steps; s++, g++) {
				const int dc = s / n_stages, c = s - dc * n_stages;
				const int st = g % MG_NS;
				const int kc4 = (min(MG_K, K - c * MG_K) + 3) & ~3;  // rows K..Kpad-1 carry zero weights (the weight blocks are zero-filled)
				const int32_t d0 = (ig.dc0 + dc) * MG_D;
				const uint32_t row_bytes = (uint32_t)(((min(MG_D, m.D - d0) + 1) & ~1) * 8);
				int32_t row = 0;
				if (lane < kc4) row = __ldg(rows + min(c * MG_K + lane, K - 1));  // padding rows copy the last row (finite values)
				mbar_wait(&empty_bar[st], ((g / MG_NS) & 1u) ^ 1u);
				if (lane == 0) mbar_expect_tx(&full_bar[st], (row_bytes + w_bytes) * (uint32_t)kc4);
				__syncwarp();
				if (lane < kc4) {
					tma_bulk_g2s_hint(sV + ((size_t)st * MG_K + lane) * MG_LDV, m.values + (int64_t)row * m.value_stride + d0, row_bytes, &full_bar[st], keep);
				} else if (lane >= 16 && lane - 16 < kc4) {
					const int kk = lane - 16;
					tma_bulk_g2s(sW + ((size_t)st * MG_K + kk) * MG_LDW, W + ig.woff + (int64_t)(c * MG_K + kk) * ig.npad + ig.qoff, w_bytes, &full_bar[st]);
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
	double acc[4][4][2];
	for (int64_t item = blockIdx.x; item < total; item += gridDim.x, it++) {
		const int ib = it & 1;
		mbar_wait(&full_bar[g % MG_NS], (g / MG_NS) & 1u);  // the item's first stage is full: its geometry and query indices are published
		const GenGeom ig = s_item[ib];
		const int32_t K = ig.K;
		const int n_stages = (K + MG_K - 1) / MG_K;
		const int n_steps = ig.n_dc * n_stages;
		const int n_mi = min(4, (ig.nq - wq * 32 + 7) >> 3);  // 8-query blocks beyond the tile's last query hold only padding
		for (int s = 0; s < n_steps; s++, g++) {
			const int dc = s / n_stages, c = s - dc * n_stages;
			const int st = g % MG_NS;
			if (c == 0) {
#pragma unroll
				for (int mi = 0; mi < 4; mi++)
#pragma unroll
					for (int ni = 0; ni < 4; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
			}
			mbar_wait(&full_bar[st], (g / MG_NS) & 1u);
			const int kc = (min(MG_K, K - c * MG_K) + 3) & ~3;
			const double *w_base = sW + (size_t)st * MG_K * MG_LDW + wq * 32 + fr;  // A[fr][fc] = W[k0 + fc][q0 + fr]
			const double *v_base = sV + (size_t)st * MG_K * MG_LDV + wd * 32 + fr;  // B[fc][fr] = V[k0 + fc][d0 + fr]
			if (n_mi > 0) {
#pragma unroll 2
				for (int k0 = 0; k0 < kc; k0 += 4) {
					double a[4], bf[4];
#pragma unroll
					for (int mi = 0; mi < 4; mi++) a[mi] = w_base[(k0 + fc) * MG_LDW + mi * 8];
#pragma unroll
					for (int ni = 0; ni < 4; ni++) bf[ni] = v_base[(k0 + fc) * MG_LDV + ni * 8];
#pragma unroll
					for (int mi = 0; mi < 4; mi++) {
						if (mi >= n_mi) break;
						// rows of padding queries inside a partly filled 8-block read stale shared memory: their results are
						// never stored and MMA rows are independent, so they cannot contaminate real queries
#pragma unroll
						for (int ni = 0; ni < 4; ni++) dmma_m8n8k4(acc[mi][ni], a[mi], bf[ni]);
					}
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty_bar[st]);
			if (c == n_stages - 1) {
				const int32_t d0 = (ig.dc0 + dc) * MG_D;
				const int32_t nd = min(MG_D, m.D - d0);
#pragma unroll
				for (int mi = 0; mi < 4; mi++) {
					const int32_t qi = sQi[ib * MG_Q + wq * 32 + mi * 8 + fr];
					if (qi < 0) continue;
					const int64_t orow = ring_base >= 0 ? (int64_t)ig.qbeg + wq * 32 + mi * 8 + fr - ring_base : (int64_t)qi;
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
		__syncwarp();
		if (lane == 0) mbar_arrive(&item_bar[ib]);
	}
}

int mesh_generic_tile_queries() { return MG_Q; }

// doubles needed for the weight blocks of a batch of Q queries (upper bound: every bucket padded by one, every K by three)
size_t mesh_generic_weight_doubles(const DeviceModel &m, int64_t Q) { return (size_t)(((int64_t)m.max_support + 3) & ~3LL) * (size_t)(Q + m.L + 2); }

// queries must already be sorted by leaf with tiles of mesh_generic_tile_queries(); leaf_of = containing leaf per ORIGINAL query index
int launch_mesh_generic_mma(const DeviceModel &m, const SortWork &w, const int32_t *leaf_of, int64_t Q, int64_t *woff, double *W, double *out,
                            int64_t out_stride, int sm_count, int64_t tile0, int64_t n_tiles, int64_t ring_base, bool weights_ready, cudaStream_t st)
{
	if (Q <= 0 || n_tiles == 0) return 0;
	int launched = 0;
	const bool cubic = m.basis == CUBIC_BSPLINE, exact = m.exact_rcp != 0;
	if (!weights_ready) {
		cudaMemsetAsync(W, 0, mesh_generic_weight_doubles(m, Q) * sizeof(double), st);
		{ LOCO_KERNEL("weight_offsets", st); weight_offsets_kernel<<<1, 1024, 0, st>>>(w.offset, m.sup_off, m.L, woff); }
		const unsigned g = (unsigned)((Q + 255) / 256);
		LOCO_KERNEL("weights_sorted", st);
#define LOCO_WCASE(NT)                                                                                                                      \
	case NT:                                                                                                                                \
		if (cubic) { if (exact) weights_sorted_kernel<NT, true, true><<<g, 256, 0, st>>>(m, w.xs, leaf_of, w.offset, woff, W);              \
		             else weights_sorted_kernel<NT, true, false><<<g, 256, 0, st>>>(m, w.xs, leaf_of, w.offset, woff, W); }                 \
		else { if (exact) weights_sorted_kernel<NT, false, true><<<g, 256, 0, st>>>(m, w.xs, leaf_of, w.offset, woff, W);                   \
		       else weights_sorted_kernel<NT, false, false><<<g, 256, 0, st>>>(m, w.xs, leaf_of, w.offset, woff, W); }                      \
		break;
		switch (m.N <= kMaxTemplateN ? m.N : 0) {
			LOCO_WCASE(0) LOCO_WCASE(1) LOCO_WCASE(2) LOCO_WCASE(3) LOCO_WCASE(4) LOCO_WCASE(5) LOCO_WCASE(6) LOCO_WCASE(7) LOCO_WCASE(8)
		}
#undef LOCO_WCASE
		launched += 2;
	}
	const int n_dchunks = (m.D + MG_D - 1) / MG_D;
	const int n_stages = std::max(1, (m.max_support + MG_K - 1) / MG_K);
	int dg = std::max(4, (32 + n_stages - 1) / n_stages);
	const int64_t slice_cap = ((int64_t)28 << 20) / std::max<int64_t>(1, m.R * MG_D * 8);
	dg = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)dg, std::max<int64_t>(slice_cap, 4), (int64_t)n_dchunks}));
	const size_t smem = (size_t)(MG_NS * MG_K * MG_LDV + MG_NS * MG_K * MG_LDW) * sizeof(double) + 2 * MG_Q * sizeof(int32_t) + 64;
	// per launch: the attribute is per device (one process may hold models on several GPUs)
	cudaFuncSetAttribute(mesh_generic_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
	if (n_tiles < 0) tile0 = 0;
	LOCO_KERNEL("mesh_generic_mma", st);
	mesh_generic_mma_kernel<<<sm_count * 2, MG_THREADS, smem, st>>>(m, w.xs, sorted_row_doubles(m.N), w.offset, w.tile_off, w.tile_leaf, woff, W, out, out_stride, n_dchunks, dg,
	                                                                tile0, n_tiles, ring_base);
	return launched + 1;
}

} // namespace loco
