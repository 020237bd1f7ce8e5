// loco_tiles.cu -- leaf-sorted tile path (the throughput path for scalar functions).
//
// The direct path (loco_kernels.cu) lets every thread walk its own leaf's term list: with 10^4..10^5 leaves
// the lists of one warp are 32 different 4 KB blocks and the kernel is bound by L2 gather latency.
// Here the batch is counting-sorted by containing leaf first, so that a CTA works on up to TILE_Q
// queries of ONE leaf: the leaf's terms are staged once into shared memory by the TMA engine
// (cp.async.bulk + mbarrier) and every thread then evaluates its own query against broadcast shared-memory
// reads -- no divergence, no per-thread gathers.  Results are scattered back to the caller's order, so the
// output does not depend on the (non-deterministic) order inside a leaf bucket.
//
//   1. locate_count   map -> descend -> containment; leaf[], status[]; rank[i] = histogram[leaf]++ (one atomic per query)
//   2. scan           exclusive scan of the histogram -> bucket offsets and tile offsets (one CTA)
//   3. scatter        pos = offset[leaf] + rank: xs[pos] = cube coordinates, idx[pos] = query index
//                     (random 8N-byte WRITES are cheaper than the random reads a gather in step 4 would need:
//                     a missed read pulls a whole 128 B line from HBM, a masked write only its sectors)
//   4. eval_tiles     persistent CTAs over contiguous tile ranges; coalesced xs/idx; leaf data via TMA bulk copies
//
// Reference semantics: identical to the direct path (see loco_kernels.cu); for D == 1 the per-term
// weight and the sample value are pre-multiplied (term_wv = weight_ * val), which only re-associates
// LCRealFunctionValue::add's  val += w * other  (w = sum of the sample's terms).
#include <algorithm>

#include "loco_basis.cuh"
#include "loco_cellpoly.h"
#include "loco_device.h"
#include "loco_sort.cuh"
#include "loco_tma.cuh"

namespace loco {

// ---------------------------------------------------------------------------------------------------
// 1. locate + histogram
// ---------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256) locate_count_kernel(DeviceModel m, const double *__restrict__ q, int64_t Q, int32_t *__restrict__ leaf_out,
                                                            uint8_t *__restrict__ status_out, int32_t *__restrict__ count, int32_t *__restrict__ rank,
                                                            double *__restrict__ out)
{
	const int N = NT > 0 ? NT : m.N;
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= Q) return;
	Coord<NT> x;
	double p[NT > 0 ? NT : kMaxN];
	// the query stream is read exactly once here: keep it out of the way of the tree cells in L1
#pragma unroll
	for (int d = 0; d < N; d++) p[d] = __ldcs(q + i * N + d);
	map_to_cube<NT>(m, p, x);
	const int32_t leaf = descend<NT>(m, x);
	const uint8_t st = located_status<NT>(m, x, leaf);
	__stcs(leaf_out + i, leaf);
	status_out[i] = st;
	if (st == ST_OK) __stcs(rank + i, atomicAdd(count + leaf, 1));
	else if (out) out[i] = __longlong_as_double(0x7ff8000000000000LL);
}

// ---------------------------------------------------------------------------------------------------
// 2. exclusive scan over leaves (single CTA): bucket offsets and tile offsets
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) scan_leaves_kernel(const int32_t *__restrict__ count, int32_t L, int32_t tile_q, int32_t *__restrict__ offset,
                                                            int32_t *__restrict__ tile_off)
{
	__shared__ int2 warp_sums[32];
	__shared__ int2 carry;
	if (threadIdx.x == 0) carry = make_int2(0, 0);
	__syncthreads();
	for (int32_t base = 0; base < L; base += 1024) {
		const int32_t l = base + threadIdx.x;
		const int c = l < L ? count[l] : 0;
		int2 v = make_int2(c, (c + tile_q - 1) / tile_q);
		int2 incl = v;
		for (int off = 1; off < 32; off <<= 1) {
			int a = __shfl_up_sync(0xffffffffu, incl.x, off), b = __shfl_up_sync(0xffffffffu, incl.y, off);
			if ((threadIdx.x & 31) >= off) { incl.x += a; incl.y += b; }
		}
		if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
		__syncthreads();
		if (threadIdx.x < 32) {
			int2 w = warp_sums[threadIdx.x], wi = w;
			for (int off = 1; off < 32; off <<= 1) {
				int a = __shfl_up_sync(0xffffffffu, wi.x, off), b = __shfl_up_sync(0xffffffffu, wi.y, off);
				if (threadIdx.x >= off) { wi.x += a; wi.y += b; }
			}
			warp_sums[threadIdx.x] = make_int2(wi.x - w.x, wi.y - w.y);  // exclusive prefix of the warp totals
		}
		__syncthreads();
		const int2 wbase = warp_sums[threadIdx.x >> 5], cr = carry;
		if (l < L) {
			offset[l] = cr.x + wbase.x + incl.x - v.x;
			tile_off[l] = cr.y + wbase.y + incl.y - v.y;
		}
		__syncthreads();
		if (threadIdx.x == 1023) carry = make_int2(cr.x + wbase.x + incl.x, cr.y + wbase.y + incl.y);
		__syncthreads();
	}
	if (threadIdx.x == 0) { offset[L] = carry.x; tile_off[L] = carry.y; }
}

// leaf of every tile: tiles [tile_off[l], tile_off[l+1]) belong to leaf l
__global__ void tile_leaf_kernel(const int32_t *__restrict__ tile_off, int32_t L, int32_t *__restrict__ tile_leaf)
{
	const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
	if (l >= L) return;
	for (int32_t t = tile_off[l], t1 = tile_off[l + 1]; t < t1; t++) tile_leaf[t] = l;
}

// ---------------------------------------------------------------------------------------------------
// 3. scatter queries (as cube coordinates) and their indices into the leaf buckets
// ---------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256) scatter_kernel(DeviceModel m, const double *__restrict__ q, const int32_t *__restrict__ leaf,
                                                       const uint8_t *__restrict__ status, const int32_t *__restrict__ rank, int64_t Q,
                                                       const int32_t *__restrict__ offset, int32_t *__restrict__ idx, double *__restrict__ xs)
{
	const int N = NT > 0 ? NT : m.N;
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= Q || status[i] != ST_OK) return;
	const int64_t pos = (int64_t)__ldg(offset + __ldcs(leaf + i)) + __ldcs(rank + i);
	double p[NT > 0 ? NT : kMaxN];
#pragma unroll
	for (int d = 0; d < N; d++) p[d] = __ldcs(q + i * N + d);
	Coord<NT> x;
	map_to_cube<NT>(m, p, x);  // same operations as in locate_count: bit-identical coordinates
	store_sorted_row<NT>(xs + pos * sorted_row_doubles(N), x, (int32_t)i, N);
	(void)idx;
}

// ---------------------------------------------------------------------------------------------------
// 4. evaluation over leaf tiles
// ---------------------------------------------------------------------------------------------------
template <int NT, bool CUBIC, bool EXACT, int TILE_Q>
__global__ void __launch_bounds__(TILE_Q) eval_tiles_kernel(DeviceModel m, const double *__restrict__ xs, double *__restrict__ out,
                                                             const int32_t *__restrict__ offset, const int32_t *__restrict__ tile_off,
                                                             const int32_t *__restrict__ idx, int32_t chunk_terms)
{
	const int N = NT > 0 ? NT : m.N;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double *s_cs = reinterpret_cast<double *>(smem_raw);  // [chunk_terms * 2N]
	double *s_wv = s_cs + (size_t)chunk_terms * 2 * N;     // [chunk_terms + 2]
	__shared__ __align__(8) uint64_t bar;
	if (threadIdx.x == 0) mbar_init(&bar, 1);
	__syncthreads();
	uint32_t phase = 0;

	const int32_t L = m.L;
	const int64_t total = tile_off[L];
	const int64_t t_begin = total * blockIdx.x / gridDim.x, t_end = total * (blockIdx.x + 1) / gridDim.x;
	int32_t leaf = -1, leaf_tile0 = 0, leaf_tile1 = 0;  // tiles [leaf_tile0, leaf_tile1) belong to `leaf`
	int32_t staged_leaf = -1;                            // leaf whose (single-chunk) term list sits in shared memory

	for (int64_t tile = t_begin; tile < t_end; tile++) {
		if (tile >= leaf_tile1) {
			// last leaf whose first tile is <= tile (empty leaves have no tiles and are skipped by the upper bound)
			int32_t lo = leaf < 0 ? 0 : leaf + 1, hi = L;
			while (lo < hi) {
				const int32_t mid = (lo + hi) >> 1;
				if (__ldg(tile_off + mid + 1) <= tile) lo = mid + 1; else hi = mid;
			}
			leaf = lo;
			leaf_tile0 = __ldg(tile_off + leaf);
			leaf_tile1 = __ldg(tile_off + leaf + 1);
		}
		const int32_t qbeg = __ldg(offset + leaf) + (int32_t)(tile - leaf_tile0) * TILE_Q;
		const int32_t qend = min(qbeg + TILE_Q, __ldg(offset + leaf + 1));
		const int32_t pos = qbeg + threadIdx.x;
		const bool active = pos < qend;
		int32_t qi = 0;
		Coord<NT> x;
		if (active) {
			qi = load_sorted_row<NT>(xs + (int64_t)pos * sorted_row_doubles(N), x, N);
		} else {
#pragma unroll
			for (int d = 0; d < (NT > 0 ? NT : 1); d++) x.set(d, 0.0);
			if (NT == 0) for (int d = 0; d < N; d++) x.set(d, 0.0);
		}
		const int32_t T0 = __ldg(m.term_off + leaf), T1 = __ldg(m.term_off + leaf + 1);
		double acc = 0.0;
		for (int32_t c0 = T0; c0 < T1; c0 += chunk_terms) {
			const int32_t c1 = min(c0 + chunk_terms, T1);
			const int32_t a0 = c0 & ~1;  // term_wv is copied from an even index so that the source is 16-byte aligned
			if (!(staged_leaf == leaf && T1 - T0 <= chunk_terms)) {
				__syncthreads();  // everyone is done with the previous contents
				if (threadIdx.x == 0) {
					const uint32_t cs_bytes = (uint32_t)(c1 - c0) * 2 * N * 8;
					const uint32_t wv_bytes = (uint32_t)(((c1 + 1) & ~1) - a0) * 8;
					mbar_expect_tx(&bar, cs_bytes + wv_bytes);
					tma_bulk_g2s(s_cs, m.term_cs + (int64_t)c0 * 2 * N, cs_bytes, &bar);
					tma_bulk_g2s(s_wv, m.term_wv + a0, wv_bytes, &bar);
				}
				mbar_wait(&bar, phase);
				phase ^= 1;
				staged_leaf = (T1 - T0 <= chunk_terms) ? leaf : -1;
			}
			const double *wv = s_wv + (c0 - a0);
			const int n = c1 - c0;
#pragma unroll 2
			for (int t = 0; t < n; t++) acc += term_value_smem<NT, CUBIC, EXACT>(s_cs + (size_t)t * 2 * N, wv[t], x, N);
		}
		if (active) out[qi] = acc;
	}
}

// Tensor-product leaves: the leaf's block (centres, supports, F^N sample values in tensor order) is one
// contiguous run of tl_stride doubles -> ONE TMA bulk copy per leaf; each thread then needs N*F
// one-dimensional basis evaluations and a sum-factorised contraction over shared memory.
// position of a tile in the sorted order: its leaf and query range (tiles of one leaf are consecutive)
struct TileCursor {
	int32_t leaf = -1, tile0 = 0, tile1 = 0;
	__device__ __forceinline__ void seek(int64_t tile, const int32_t *__restrict__ tile_off, int32_t L)
	{
		if (tile < tile1) return;
		// last leaf whose first tile is <= tile (leaves without queries own no tiles and are skipped)
		int32_t lo = leaf < 0 ? 0 : leaf + 1, hi = L;
		while (lo < hi) {
			const int32_t mid = (lo + hi) >> 1;
			if (__ldg(tile_off + mid + 1) <= tile) lo = mid + 1; else hi = mid;
		}
		leaf = lo;
		tile0 = __ldg(tile_off + leaf);
		tile1 = __ldg(tile_off + leaf + 1);
	}
};

template <int NT, bool CUBIC, bool EXACT, int TILE_Q>
__global__ void __launch_bounds__(TILE_Q) eval_tensor_tiles_kernel(DeviceModel m, const double *__restrict__ xs, double *__restrict__ out,
                                                                    const int32_t *__restrict__ offset, const int32_t *__restrict__ tile_off,
                                                                    const int32_t *__restrict__ idx)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	// two leaf blocks: the block of the NEXT tile's leaf is fetched by the TMA engine while this tile computes
	double *const s_base = reinterpret_cast<double *>(smem_raw);
#define S_BLK(b) (s_base + (b) * m.tl_stride)
	__shared__ __align__(8) uint64_t bar[2];
	if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
	__syncthreads();
	uint32_t phases = 0;  // bit b = parity to wait for on bar[b]
	const int32_t L = m.L;
	const int64_t total = tile_off[L];
	const int64_t t_begin = total * blockIdx.x / gridDim.x, t_end = total * (blockIdx.x + 1) / gridDim.x;
	if (t_begin >= t_end) return;
	constexpr int RS = ((NT * 8 + 4 + 31) / 32) * 4;
	const uint32_t blk_bytes = (uint32_t)m.tl_stride * 8;

	TileCursor cur;
	// prologue: records and leaf block of the first tile
	cur.seek(t_begin, tile_off, L);
	int buf = 0;
	if (threadIdx.x == 0) {
		mbar_expect_tx(&bar[0], blk_bytes);
		tma_bulk_g2s(S_BLK(0), m.tl_block + (int64_t)cur.leaf * m.tl_stride, blk_bytes, &bar[0]);
	}
	int32_t n_leaf = cur.leaf;
	int32_t n_qbeg = __ldg(offset + cur.leaf) + (int32_t)(t_begin - cur.tile0) * TILE_Q;
	int32_t n_qend = min(n_qbeg + TILE_Q, __ldg(offset + cur.leaf + 1));
	Coord<NT> nx;
	int32_t nqi = 0;
#pragma unroll
	for (int d = 0; d < NT; d++) nx.set(d, 0.0);
	if (n_qbeg + (int)threadIdx.x < n_qend) nqi = load_sorted_row<NT>(xs + (int64_t)(n_qbeg + threadIdx.x) * RS, nx, NT);
	bool need_wait = true;  // the current buffer has a copy in flight

	for (int64_t tile = t_begin; tile < t_end; tile++) {
		const Coord<NT> x = nx;
		const int32_t qi = nqi, leaf = n_leaf;
		const bool active = n_qbeg + (int)threadIdx.x < n_qend;
		// ---- look ahead: next tile's records (and its leaf block if the leaf changes)
		bool next_changes = false;
		if (tile + 1 < t_end) {
			cur.seek(tile + 1, tile_off, L);
			next_changes = cur.leaf != leaf;
			n_leaf = cur.leaf;
			n_qbeg = __ldg(offset + cur.leaf) + (int32_t)(tile + 1 - cur.tile0) * TILE_Q;
			n_qend = min(n_qbeg + TILE_Q, __ldg(offset + cur.leaf + 1));
			if (next_changes && threadIdx.x == 0) {
				// the other buffer was last read two leaves ago; every thread has passed the __syncthreads that
				// followed that leaf, so it is free
				mbar_expect_tx(&bar[buf ^ 1], blk_bytes);
				tma_bulk_g2s(S_BLK(buf ^ 1), m.tl_block + (int64_t)n_leaf * m.tl_stride, blk_bytes, &bar[buf ^ 1]);
			}
#pragma unroll
			for (int d = 0; d < NT; d++) nx.set(d, 0.0);
			nqi = 0;
			if (n_qbeg + (int)threadIdx.x < n_qend) nqi = load_sorted_row<NT>(xs + (int64_t)(n_qbeg + threadIdx.x) * RS, nx, NT);
		}
		// ---- this tile
		if (need_wait) { mbar_wait(&bar[buf], (phases >> buf) & 1u); phases ^= 1u << buf; need_wait = false; }
		const double r = m.tl_poly ? tensor_leaf_value<NT, CUBIC, EXACT>(m, nullptr, m.tl_poly + (int64_t)leaf * m.tp_stride, leaf, x)
		                           : tensor_eval_scalar<NT, CUBIC, EXACT>(m, S_BLK(buf), x);
		if (active) out[qi] = r;
		if (next_changes) {
			__syncthreads();  // all reads of s_blk[buf] are done before it can be refilled (two leaves from now)
			buf ^= 1;
			need_wait = true;
		}
	}
#undef S_BLK
}

// term_wv[t] = term_w[t] * values[term_row[t]]   (scalar functions; refreshed whenever the value table changes)
__global__ void term_wv_kernel(DeviceModel m, double *__restrict__ term_wv, int64_t T)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < T) term_wv[t] = m.term_w[t] * m.values[(int64_t)m.term_row[t] * m.value_stride];
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
namespace {
const int kTileQ = 128;

template <int NT>
void locate_count(const DeviceModel &m, const double *q, int64_t Q, int32_t *leaf, uint8_t *status, const SortWork &w, double *out, cudaStream_t st)
{
	LOCO_KERNEL("locate_count", st);
	locate_count_kernel<NT><<<(unsigned)((Q + 255) / 256), 256, 0, st>>>(m, q, Q, leaf, status, w.count, w.rank, out);
}
template <int NT>
void scatter(const DeviceModel &m, const double *q, int64_t Q, const int32_t *leaf, const uint8_t *status, const SortWork &w, cudaStream_t st)
{
	LOCO_KERNEL("scatter", st);
	scatter_kernel<NT><<<(unsigned)((Q + 255) / 256), 256, 0, st>>>(m, q, leaf, status, w.rank, Q, w.offset, w.idx, w.xs);
}

template <int NT, bool CUBIC, bool EXACT>
void eval_tensor_tiles(const DeviceModel &m, const double *q, double *out, const SortWork &w, int grid, cudaStream_t st)
{
	auto kern = eval_tensor_tiles_kernel<NT, CUBIC, EXACT, kTileQ>;
	const size_t smem = (size_t)m.tl_stride * 8 * 2;
	if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	LOCO_KERNEL("eval_tensor_tiles", st);
	kern<<<grid, kTileQ, smem, st>>>(m, w.xs, out, w.offset, w.tile_off, w.idx);
}

template <int NT, bool CUBIC, bool EXACT>
void eval_tiles(const DeviceModel &m, const double *q, double *out, const SortWork &w, int chunk_terms, size_t smem, int grid, cudaStream_t st)
{
	auto kern = eval_tiles_kernel<NT, CUBIC, EXACT, kTileQ>;
	LOCO_KERNEL("eval_tiles", st);
	kern<<<grid, kTileQ, smem, st>>>(m, w.xs, out, w.offset, w.tile_off, w.idx, chunk_terms);
}
} // namespace

size_t sort_work_bytes(const DeviceModel &m, int64_t Q)
{
	const size_t ints = (size_t)m.L + (size_t)(m.L + 1) * 2 + (size_t)Q + (size_t)(Q / 32 + m.L + 1);
	return ((ints * sizeof(int32_t) + 127) & ~(size_t)127) + (size_t)Q * sorted_row_doubles(m.N) * sizeof(double) + 256;
}

SortWork carve_sort_work(const DeviceModel &m, int64_t Q, void *base)
{
	SortWork w;
	int32_t *p = reinterpret_cast<int32_t *>(base);
	w.count = p; p += m.L;
	w.offset = p; p += m.L + 1;
	w.tile_off = p; p += m.L + 1;
	w.rank = p; p += Q;
	w.tile_leaf = p; p += Q / 32 + m.L + 1;
	w.idx = nullptr;  // the index travels in the last slot of each sorted row
	const size_t ints = (size_t)(p - reinterpret_cast<int32_t *>(base));
	w.xs = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(base) + ((ints * sizeof(int32_t) + 127) & ~(size_t)127));
	return w;
}

// mterm_wv[t] = sum over the merged run of term_w * value   (summed in run order: ascending original term index)
__global__ void mterm_wv_kernel(DeviceModel m, double *__restrict__ mterm_wv)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= m.n_mterms) return;
	double s = 0.0;
	for (int e = m.mrun_off[t], e1 = m.mrun_off[t + 1]; e < e1; e++) {
		const int32_t o = m.mrun_term[e];
		s += m.term_w[o] * m.values[(int64_t)m.term_row[o] * m.value_stride];
	}
	mterm_wv[t] = s;
}
int launch_update_mterm_wv(const DeviceModel &m, double *mterm_wv, cudaStream_t st)
{
	if (!m.mterm_off || m.n_mterms <= 0) return 0;
	mterm_wv_kernel<<<(unsigned)((m.n_mterms + 255) / 256), 256, 0, st>>>(m, mterm_wv);
	return 1;
}

__global__ void eterm_wv_kernel(DeviceModel m, double *__restrict__ eterm_wv)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < m.n_eterms) eterm_wv[t] = m.mterm_wv[m.eterm_src[t]];
}
int launch_update_eterm_wv(const DeviceModel &m, double *eterm_wv, cudaStream_t st)
{
	if (!m.eterm_src || m.n_eterms <= 0) return 0;
	eterm_wv_kernel<<<(unsigned)((m.n_eterms + 255) / 256), 256, 0, st>>>(m, eterm_wv);
	return 1;
}

// coefficients of the polynomial evaluation cells from the current eterm_wv (one thread per cell; runs when values change)
template <bool CUBIC>
__global__ void epoly_kernel(DeviceModel m, const int32_t *__restrict__ epoly_cell, int32_t n_epoly, double *__restrict__ epoly)
{
	constexpr int F = CUBIC ? 4 : 2;
	const int pi = blockIdx.x * blockDim.x + threadIdx.x;
	if (pi >= n_epoly) return;
	const int N = m.N;
	double *blk = epoly + (int64_t)pi * m.epoly_stride;
	double coef[64], p[8 * F];
	int P = 1;
	for (int d = 0; d < N; d++) P *= F;
	for (int j = 0; j < P; j++) coef[j] = 0.0;
	const int32_t cell = epoly_cell[pi];
	for (int t = m.eterm_off[cell], t1 = m.eterm_off[cell + 1]; t < t1; t++) {
		for (int d = 0; d < N; d++) {
			const double c = m.eterm_cs[(int64_t)t * 2 * N + 2 * d], v = m.eterm_cs[(int64_t)t * 2 * N + 2 * d + 1];
			cell_piece_1d<CUBIC>(c, m.exact_rcp ? 1.0 / v : v, blk[2 * d], 1.0 / blk[2 * d + 1], p + d * F);
		}
		cell_poly_accumulate<F>(N, p, m.eterm_wv[t], coef);
	}
	for (int j = 0; j < P; j++) blk[2 * N + j] = coef[j];
}
int launch_update_epoly(const DeviceModel &m, const int32_t *epoly_cell, int32_t n_epoly, double *epoly, cudaStream_t st)
{
	if (!epoly || n_epoly <= 0) return 0;
	if (m.basis == CUBIC_BSPLINE) epoly_kernel<true><<<(n_epoly + 127) / 128, 128, 0, st>>>(m, epoly_cell, n_epoly, epoly);
	else epoly_kernel<false><<<(n_epoly + 127) / 128, 128, 0, st>>>(m, epoly_cell, n_epoly, epoly);
	return 1;
}

int launch_update_term_wv(const DeviceModel &m, double *term_wv, int64_t T, cudaStream_t st)
{
	if (T <= 0) return 0;
	term_wv_kernel<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(m, term_wv, T);
	return 1;
}

// steps 1-3: leaf[], status[], and the queries as sorted records (w.xs) with bucket / tile offsets for tiles of tile_q
int launch_sort_by_leaf(const DeviceModel &m, const double *q, int64_t Q, int32_t *leaf, uint8_t *status, const SortWork &w, int tile_q,
                        double *scalar_out, cudaStream_t st)
{
	double *out = scalar_out;
	cudaMemsetAsync(w.count, 0, (size_t)m.L * sizeof(int32_t), st);
	switch (m.N <= kMaxTemplateN ? m.N : 0) {
	case 1: locate_count<1>(m, q, Q, leaf, status, w, out, st); break;
	case 2: locate_count<2>(m, q, Q, leaf, status, w, out, st); break;
	case 3: locate_count<3>(m, q, Q, leaf, status, w, out, st); break;
	case 4: locate_count<4>(m, q, Q, leaf, status, w, out, st); break;
	case 5: locate_count<5>(m, q, Q, leaf, status, w, out, st); break;
	case 6: locate_count<6>(m, q, Q, leaf, status, w, out, st); break;
	case 7: locate_count<7>(m, q, Q, leaf, status, w, out, st); break;
	case 8: locate_count<8>(m, q, Q, leaf, status, w, out, st); break;
	default: locate_count<0>(m, q, Q, leaf, status, w, out, st); break;
	}
	{ LOCO_KERNEL("scan_leaves", st);
	scan_leaves_kernel<<<1, 1024, 0, st>>>(w.count, m.L, tile_q, w.offset, w.tile_off); }
	tile_leaf_kernel<<<(unsigned)((m.L + 255) / 256), 256, 0, st>>>(w.tile_off, m.L, w.tile_leaf);
	switch (m.N <= kMaxTemplateN ? m.N : 0) {
	case 1: scatter<1>(m, q, Q, leaf, status, w, st); break;
	case 2: scatter<2>(m, q, Q, leaf, status, w, st); break;
	case 3: scatter<3>(m, q, Q, leaf, status, w, st); break;
	case 4: scatter<4>(m, q, Q, leaf, status, w, st); break;
	case 5: scatter<5>(m, q, Q, leaf, status, w, st); break;
	case 6: scatter<6>(m, q, Q, leaf, status, w, st); break;
	case 7: scatter<7>(m, q, Q, leaf, status, w, st); break;
	case 8: scatter<8>(m, q, Q, leaf, status, w, st); break;
	default: scatter<0>(m, q, Q, leaf, status, w, st); break;
	}
	return 3;
}

int launch_eval_scalar_tiles(const DeviceModel &m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, const SortWork &w,
                             int sm_count, cudaStream_t st)
{
	if (Q <= 0) return 0;
	launch_sort_by_leaf(m, q, Q, leaf, status, w, kTileQ, out, st);

	const int per_term = 2 * m.N * 8 + 8;
	int chunk = (40 * 1024) / per_term;
	chunk = std::max(2, std::min(chunk & ~1, (m.max_terms + 1) & ~1));
	const size_t smem = (size_t)chunk * 2 * m.N * 8 + (size_t)(chunk + 2) * 8;
	const int64_t max_tiles = Q / kTileQ + m.L + 1;
	const int grid = (int)std::min<int64_t>(max_tiles, (int64_t)sm_count * 16);
	const bool cubic = m.basis == CUBIC_BSPLINE, exact = m.exact_rcp != 0;
	if (m.tensor_F > 0 && m.N <= kMaxTemplateN && (size_t)m.tl_stride * 16 <= 200 * 1024) {
#define LOCO_XCASE(NT)                                                                                                   \
	case NT:                                                                                                             \
		if (cubic) { if (exact) eval_tensor_tiles<NT, true, true>(m, q, out, w, grid, st); else eval_tensor_tiles<NT, true, false>(m, q, out, w, grid, st); } \
		else { if (exact) eval_tensor_tiles<NT, false, true>(m, q, out, w, grid, st); else eval_tensor_tiles<NT, false, false>(m, q, out, w, grid, st); }     \
		break;
		switch (m.N) { LOCO_XCASE(1) LOCO_XCASE(2) LOCO_XCASE(3) LOCO_XCASE(4) LOCO_XCASE(5) LOCO_XCASE(6) LOCO_XCASE(7) LOCO_XCASE(8) }
#undef LOCO_XCASE
		return 4;
	}
#define LOCO_TCASE(NT)                                                                                                   \
	case NT:                                                                                                             \
		if (cubic) { if (exact) eval_tiles<NT, true, true>(m, q, out, w, chunk, smem, grid, st); else eval_tiles<NT, true, false>(m, q, out, w, chunk, smem, grid, st); } \
		else { if (exact) eval_tiles<NT, false, true>(m, q, out, w, chunk, smem, grid, st); else eval_tiles<NT, false, false>(m, q, out, w, chunk, smem, grid, st); }     \
		break;
	switch (m.N <= kMaxTemplateN ? m.N : 0) {
		LOCO_TCASE(0) LOCO_TCASE(1) LOCO_TCASE(2) LOCO_TCASE(3) LOCO_TCASE(4) LOCO_TCASE(5) LOCO_TCASE(6) LOCO_TCASE(7) LOCO_TCASE(8)
	}
#undef LOCO_TCASE
	return 4;
}

} // namespace loco
