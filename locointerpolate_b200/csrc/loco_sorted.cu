// loco_sorted.cu -- the throughput path for scalar functions: L2-resident chunks, counting sort by leaf, evaluation in
// leaf order.
//
// A big batch is cut into chunks of a few million queries.  Per chunk, four launches:
//   1. count     map -> locate (locate grid, loco_basis.cuh) -> containment; writes leaf[] / status[]; one
//                fire-and-forget atomic per query into the per-leaf histogram.  A one-CTA scan then turns the
//                histogram into bucket cursors and clears it for the next chunk
//   2. scatter   the query (and its leaf) is read again -- from L2: the chunk was just streamed through it --, takes
//                its slot with one atomicAdd on its leaf's cursor and writes a 32-byte-sector record
//                {cube coordinates, query index, leaf}
//   3. eval      thread i evaluates record i.  Records are in leaf order, so the 32 lanes of a warp read the SAME
//                leaf block (or two adjacent ones): every load of leaf data is a warp-wide broadcast served by L1,
//                there is no divergence and no per-thread gather; the result goes to out[query index]
// Why chunks: the records (32 B/query) and the chunk's slice of out[] (8 B/query) stay in the 126 MB L2 between the
// kernels, so the only HBM traffic left is compulsory -- the queries in, and leaf/status/out written back as FULL
// sectors.  Evaluating the whole batch at once instead scatters 8-byte results over 800 MB: every one of them costs a
// 32-byte sector fill + write-back (ncu: 6.4 GB read + 3.2 GB written for 0.8 GB of results, profiles/r01a_c2_*).
//
// Reference semantics are those of the direct kernels (loco_kernels.cu / loco_tensor.cu); the slot order inside a
// leaf bucket depends on atomic arrival order, but every query is evaluated independently, so results do not.
#include <algorithm>

#include "loco_basis.cuh"
#include "loco_device.h"

namespace loco {

namespace {

// record: N cube coordinates, then {int32 query index, int32 leaf} in one 8-byte slot, padded to whole 32-byte sectors
__host__ __device__ constexpr int rec_doubles(int N) { return ((N * 8 + 8 + 31) / 32) * 4; }

__device__ __forceinline__ void st_sector_cg(double *p, double a, double b, double c, double d)
{
	asm volatile("st.global.cg.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld_sector_cg(const double *p, double &a, double &b, double &c, double &d)
{
	asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// number of sort buckets: evaluation cells where the model has them (Flat::ecell_*), leaves otherwise
__host__ __device__ __forceinline__ int32_t sort_buckets(const DeviceModel &m) { return m.ecell_off ? m.n_ecells : m.L; }

// evaluation cell of a located point.  Pure function of (x, leaf): the count and the scatter pass get the same cell.
// The index need not be exact at cell faces — neighbouring cells both list every term that reaches the face.
template <int NT>
__device__ __forceinline__ int32_t eval_cell(const DeviceModel &m, const Coord<NT> &x, int32_t leaf)
{
	if (!m.ecell_off) return leaf;
	const int N = NT > 0 ? NT : m.N;
	int idx = 0, shift = 0;
#pragma unroll
	for (int d = 0; d < (NT > 0 ? NT : kMaxN); d++) {
		if (d >= N) break;
		const int b = m.ecell_bits[(int64_t)leaf * N + d];
		if (b) {
			const double lo = m.leaf_lo[(int64_t)leaf * N + d], hi = m.leaf_hi[(int64_t)leaf * N + d];
			const int g = 1 << b;
			int k = (int)((x.get(d) - lo) * (double)g / (hi - lo));
			k = max(0, min(g - 1, k));
			idx |= k << shift;
			shift += b;
		}
	}
	return __ldg(m.ecell_off + leaf) + idx;
}

// ---------------------------------------------------------------------------------------------------
// 1. locate + histogram
// ---------------------------------------------------------------------------------------------------
constexpr int kCountItems = 2;  // queries per thread: two independent load -> locate -> atomic chains in flight
// AGG: lanes of a warp with the same sort key share one atomic (see bucket_scatter_kernel)
template <int NT, bool AGG>
__global__ void __launch_bounds__(256) sorted_count_kernel(DeviceModel m, const double *__restrict__ q, int64_t Q, int32_t *__restrict__ leaf_out,
                                                            uint8_t *__restrict__ status_out, double *__restrict__ out, int32_t *count)
{
	const int N = NT > 0 ? NT : m.N;
	const int64_t i0 = (int64_t)blockIdx.x * (256 * kCountItems) + threadIdx.x;
	double p[kCountItems][NT > 0 ? NT : kMaxN];
#pragma unroll
	for (int k = 0; k < kCountItems; k++) {
		const int64_t i = i0 + k * 256;
		if (i < Q) {
#pragma unroll
			for (int d = 0; d < N; d++) p[k][d] = q[i * N + d];
		}
	}
#pragma unroll
	for (int k = 0; k < kCountItems; k++) {
		const int64_t i = i0 + k * 256;
		int32_t key = -1 - (int)(threadIdx.x & 31);  // lanes without a valid query match nobody
		if (i < Q) {
			Coord<NT> x;
			map_to_cube<NT>(m, p[k], x);
			const int32_t leaf = descend<NT>(m, x);
			const uint8_t st = located_status<NT>(m, x, leaf);
			leaf_out[i] = leaf;
			status_out[i] = st;
			if (st == ST_OK) key = eval_cell<NT>(m, x, leaf);
			else out[i] = __longlong_as_double(0x7ff8000000000000LL);
		}
		if constexpr (AGG) {
			const unsigned peers = __match_any_sync(0xffffffffu, key);
			if (key >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(count + key, __popc(peers));
		} else {
			if (key >= 0) atomicAdd(count + key, 1);  // result unused: a fire-and-forget RED
		}
	}
}

// exclusive scan of the histogram into bucket cursors; clears the histogram for the next chunk.
// One CTA: the histogram is staged in shared memory with coalesced loads that are all in flight together (walking it
// thread by thread straight from global memory costs one uncoalesced L2 round trip per element: measured 80 us for
// 32,768 leaves), scanned there in conflict-free padded runs, and written back coalesced.
constexpr int kScanThreads = 1024;
constexpr int kScanSeg = 32768;  // leaves per pass
__device__ __forceinline__ int scan_pad(int e) { return e + (e >> 5); }
__global__ void __launch_bounds__(kScanThreads) sorted_scan_kernel(int32_t *count, int32_t *cursor, int32_t L, SkewStats *stats, int32_t stats_cap)
{
	extern __shared__ int s_hist[];  // [kScanSeg + kScanSeg / 32]
	__shared__ int warp_tot[kScanThreads / 32];
	__shared__ unsigned long long s_over;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_over = 0ull;
	long long over = 0;  // queries beyond a fixed-capacity bucket of stats_cap slots (skew statistic)
	int carry = 0;
	for (int base = 0; base < L; base += kScanSeg) {
		const int n = min(kScanSeg, L - base);
		for (int e = threadIdx.x; e < n; e += kScanThreads) s_hist[scan_pad(e)] = count[base + e];
		__syncthreads();
		const int per = (n + kScanThreads - 1) / kScanThreads;
		const int l0 = min((int)threadIdx.x * per, n), l1 = min(l0 + per, n);
		int mine = 0;
		for (int l = l0; l < l1; l++) { const int c = s_hist[scan_pad(l)]; mine += c; over += max(0, c - stats_cap); }
		int incl = mine;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			const int v = __shfl_up_sync(0xffffffffu, incl, off);
			if (lane >= off) incl += v;
		}
		if (lane == 31) warp_tot[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			const int w = warp_tot[lane];
			int wi = w;
#pragma unroll
			for (int off = 1; off < 32; off <<= 1) {
				const int v = __shfl_up_sync(0xffffffffu, wi, off);
				if (lane >= off) wi += v;
			}
			warp_tot[lane] = wi - w;  // exclusive prefix of the warp totals
		}
		__syncthreads();
		int run = carry + warp_tot[warp] + incl - mine;
		for (int l = l0; l < l1; l++) {
			const int c = s_hist[scan_pad(l)];
			s_hist[scan_pad(l)] = run;
			run += c;
		}
		__syncthreads();
		if (threadIdx.x == kScanThreads - 1) warp_tot[0] = run;  // total so far (the last thread owns the last run)
		for (int e = threadIdx.x; e < n; e += kScanThreads) { cursor[base + e] = s_hist[scan_pad(e)]; count[base + e] = 0; }
		__syncthreads();
		carry = warp_tot[0];
		__syncthreads();
	}
	if (threadIdx.x == 0) cursor[L] = carry;
	if (stats) {
		if (over) atomicAdd(&s_over, (unsigned long long)over);
		__syncthreads();
		if (threadIdx.x == 0) { stats->overflow += (long long)s_over; stats->total += carry; }
	}
}

// ---------------------------------------------------------------------------------------------------
// 2. scatter into leaf buckets
// ---------------------------------------------------------------------------------------------------
constexpr int kScatterItems = 2;  // queries per thread: two independent load -> atomic -> store chains in flight
template <int NT, bool AGG>
__global__ void __launch_bounds__(256) sorted_scatter_kernel(DeviceModel m, const double *__restrict__ q, int64_t Q, const int32_t *__restrict__ leaf_in,
                                                              const uint8_t *__restrict__ status_in, int32_t *cursor, double *__restrict__ xs)
{
	const int N = NT > 0 ? NT : m.N;
	constexpr int RSmax = rec_doubles(NT > 0 ? NT : kMaxN);
	const int RS = rec_doubles(N);
	const int64_t i0 = (int64_t)blockIdx.x * (256 * kScatterItems) + threadIdx.x;
	// independent loads first (status, leaf, query), then the one dependent round trip (the cursor atomic)
	uint8_t st[kScatterItems];
	int32_t leaf[kScatterItems];
	double p[kScatterItems][NT > 0 ? NT : kMaxN];
#pragma unroll
	for (int k = 0; k < kScatterItems; k++) {
		const int64_t i = i0 + k * 256;
		st[k] = 0xff;
		if (i < Q) {
			st[k] = status_in[i];
			leaf[k] = __ldcs(leaf_in + i);
#pragma unroll
			for (int d = 0; d < N; d++) p[k][d] = __ldcs(q + i * N + d);
		}
	}
	int64_t pos[kScatterItems];
	if (m.ecell_off) {  // sort key = evaluation cell (needs the cube coordinates before the atomic)
#pragma unroll
		for (int k = 0; k < kScatterItems; k++)
			if (st[k] == ST_OK) {
				Coord<NT> x;
				map_to_cube<NT>(m, p[k], x);
				leaf[k] = eval_cell<NT>(m, x, leaf[k]);
			}
	}
#pragma unroll
	for (int k = 0; k < kScatterItems; k++) {
		if constexpr (AGG) {
			const int lane_id = threadIdx.x & 31;
			const int32_t key = st[k] == ST_OK ? leaf[k] : -1 - lane_id;
			const unsigned peers = __match_any_sync(0xffffffffu, key);
			const int leader = __ffs(peers) - 1;
			int base = 0;
			if (st[k] == ST_OK && lane_id == leader) base = atomicAdd(cursor + leaf[k], __popc(peers));
			base = __shfl_sync(0xffffffffu, base, leader);
			pos[k] = base + __popc(peers & ((1u << lane_id) - 1));
		} else {
			if (st[k] == ST_OK) pos[k] = atomicAdd(cursor + leaf[k], 1);
		}
	}
#pragma unroll
	for (int k = 0; k < kScatterItems; k++) {
		if (st[k] != ST_OK) continue;
		Coord<NT> x;
		map_to_cube<NT>(m, p[k], x);  // the same operations as in the count kernel: bit-identical coordinates
		double v[RSmax];
#pragma unroll
		for (int d = 0; d < RSmax; d++) v[d] = (d < N) ? x.get(d) : 0.0;
		const long long tag = ((long long)(uint32_t)leaf[k] << 32) | (uint32_t)(int32_t)(i0 + k * 256);
		if (NT > 0) v[RSmax - 1] = __longlong_as_double(tag); else v[RS - 1] = __longlong_as_double(tag);
		double *row = xs + pos[k] * RS;
#pragma unroll
		for (int d = 0; d < RSmax; d += 4)
			if (d < RS) st_sector_cg(row + d, v[d], v[d + 1], v[d + 2], v[d + 3]);
	}
}

// ---------------------------------------------------------------------------------------------------
// 3. evaluation in leaf order
// ---------------------------------------------------------------------------------------------------
// Persistent threads: thread t evaluates records t, t+T, t+2T, ...  While record i is being evaluated, record i+T is
// already in registers with the lines of ITS leaf block on their way into L1, and the load of record i+2T is in flight:
// the two dependent memory round trips of a query (record -> leaf block) overlap the arithmetic of its predecessors.
// DERIV (generic leaves, cubic basis): LCSampledFunction::evalDeriv along cube direction `dir` — the same term lists with
// LCCubicBSpline::evalDeriv's factor in that dimension (a term that is identically zero in an evaluation cell has a zero derivative there too).
template <int NT, bool CUBIC, bool EXACT, bool TENSOR, bool FP32, bool DERIV = false>
__global__ void __launch_bounds__(256, (NT > 0 && NT <= 3) ? 3 : 2) sorted_eval_kernel(DeviceModel m, const double *__restrict__ xs, const int32_t *__restrict__ cursor,
                                                           double *__restrict__ out, int dir = -1)
{
	const int N = NT > 0 ? NT : m.N;
	constexpr int RSmax = rec_doubles(NT > 0 ? NT : kMaxN);
	const int RS = rec_doubles(N);
	const int64_t total = __ldg(cursor + sort_buckets(m));  // number of records of this chunk
	const int64_t T = (int64_t)gridDim.x * blockDim.x;
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= total) return;
	// two records ahead: `a` = record i (being evaluated), `b` = record i+T (arrived; its leaf block is being prefetched),
	// and the load of record i+2T is in flight
	double a[RSmax], b[RSmax];
	auto load_rec = [&](double (&v)[RSmax], int64_t r) {
		const double *row = xs + r * RS;
#pragma unroll
		for (int d = 0; d < RSmax; d += 4)
			if (d < RS) ld_sector_cg(row + d, v[d], v[d + 1], v[d + 2], v[d + 3]);
	};
	load_rec(a, i);
	if (i + T < total) load_rec(b, i + T);
	for (; i < total; i += T) {
		Coord<NT> x;
#pragma unroll
		for (int d = 0; d < (NT > 0 ? NT : kMaxN); d++) if (d < N) x.set(d, a[d]);
		const long long tag = __double_as_longlong(NT > 0 ? a[RSmax - 1] : a[RS - 1]);
		const int32_t qi = (int32_t)(uint32_t)tag, leaf = (int32_t)(tag >> 32);
		if (i + T < total) {
			if constexpr (TENSOR) {
				// pull the lines of the NEXT record's leaf block into L1 before starting on this record's arithmetic
				const int32_t nleaf = (int32_t)(__double_as_longlong(b[RSmax - 1]) >> 32);
				const char *blk_bytes = reinterpret_cast<const char *>(m.tl_block + (int64_t)nleaf * m.tl_stride);
				if constexpr (FP32) {
					// centres / supports from the fp64 block, values from the fp32 copy
					const char *val_bytes = reinterpret_cast<const char *>(m.tl_vals_f32 + (int64_t)nleaf * m.tensor_K);
					for (int o = 0; o < m.tl_voff * 8; o += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(blk_bytes + o));
					for (int o = 0; o < m.tensor_K * 4; o += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(val_bytes + o));
				} else if (m.tl_poly) {
					const char *pb = reinterpret_cast<const char *>(m.tl_poly + (int64_t)nleaf * m.tp_stride);
					for (int o = 0; o < m.tp_stride * 8; o += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(pb + o));
				} else
				for (int o = 0; o < m.tl_stride * 8; o += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(blk_bytes + o));
			}
#pragma unroll
			for (int d = 0; d < RSmax; d++) a[d] = b[d];
			if (i + 2 * T < total) load_rec(b, i + 2 * T);
		}
		double r = 0.0;
		if constexpr (TENSOR) {
			if constexpr (FP32) r = tensor_eval_scalar_f32<NT, CUBIC, EXACT>(m, m.tl_block + (int64_t)leaf * m.tl_stride, m.tl_vals_f32 + (int64_t)leaf * m.tensor_K, x);
			else r = tensor_leaf_value<NT, CUBIC, EXACT>(m, nullptr, m.tl_poly ? m.tl_poly + (int64_t)leaf * m.tp_stride : nullptr, leaf, x);
		} else {
			// LCSample::getInterpolationWeight / newFromWeightedSum over the leaf's term list, as in eval_scalar_direct_kernel,
			// with weight_ * val pre-multiplied per term (term_wv)
			auto term = [&](const double *cs, double wv) {
				if constexpr (DERIV) return term_deriv<NT, CUBIC, EXACT>(cs, wv, x, dir, N);
				else return term_value<NT, CUBIC, EXACT>(cs, wv, x, N);
			};
			int32_t pi = -1;
			if constexpr (!DERIV && NT > 0 && ipow_c(CUBIC ? 4 : 2, NT) <= 64) {
				if (m.ecell_off && m.epoly_idx) pi = __ldg(m.epoly_idx + leaf);
			}
			if (pi >= 0) {
				// the cell lies inside one polynomial piece of every function that reaches it (loco_cellpoly.h):
				// powers of the local coordinate per dimension, then the same contraction as a tensor-product leaf
				if constexpr (!DERIV && NT > 0 && ipow_c(CUBIC ? 4 : 2, NT) <= 64) {
					constexpr int F = CUBIC ? 4 : 2;
					const double *blk = m.epoly + (int64_t)pi * m.epoly_stride;
					double f[NT][F];
#pragma unroll
					for (int d = 0; d < NT; d++) {
						const double t = (x.get(d) - __ldg(blk + 2 * d)) * __ldg(blk + 2 * d + 1);
						f[d][0] = 1.0;
						f[d][1] = t;
						if constexpr (CUBIC) { f[d][2] = t * t; f[d][3] = f[d][2] * t; }
					}
					r = TensorContract<NT, F, NT - 1>::run(blk + 2 * NT, f, ipow_c(F, NT) / F);
				}
			} else if (m.ecell_off) {  // `leaf` is the evaluation cell: only the merged terms that reach this part of the leaf
				for (int t = __ldg(m.eterm_off + leaf), t1 = __ldg(m.eterm_off + leaf + 1); t < t1; t++)
					r += term(m.eterm_cs + (int64_t)t * 2 * N, __ldg(m.eterm_wv + t));
			} else if (m.mterm_off) {  // merged: one term per distinct basis function of the leaf
				for (int t = __ldg(m.mterm_off + leaf), t1 = __ldg(m.mterm_off + leaf + 1); t < t1; t++)
					r += term(m.mterm_cs + (int64_t)t * 2 * N, __ldg(m.mterm_wv + t));
			} else {
				for (int t = __ldg(m.term_off + leaf), t1 = __ldg(m.term_off + leaf + 1); t < t1; t++)
					r += term(m.term_cs + (int64_t)t * 2 * N, __ldg(m.term_wv + t));
			}
		}
		out[qi] = r;
	}
}

template <int NT>
void run_count(const DeviceModel &m, const double *q, int64_t Q, int32_t *leaf, uint8_t *status, double *out, const SortedWork &w, cudaStream_t st)
{
	{ LOCO_KERNEL("sorted_count", st);
	const unsigned g = (unsigned)((Q + 256 * kCountItems - 1) / (256 * kCountItems));
	if (w.aggregate) sorted_count_kernel<NT, true><<<g, 256, 0, st>>>(m, q, Q, leaf, status, out, w.count);
	else sorted_count_kernel<NT, false><<<g, 256, 0, st>>>(m, q, Q, leaf, status, out, w.count); }
	LOCO_KERNEL("sorted_scan", st);
	sorted_scan_kernel<<<1, kScanThreads, (kScanSeg + kScanSeg / 32) * sizeof(int), st>>>(w.count, w.cursor, sort_buckets(m), w.stats, w.stats_cap);
}
template <int NT>
void run_scatter(const DeviceModel &m, const double *q, int64_t Q, const int32_t *leaf, const uint8_t *status, const SortedWork &w, cudaStream_t st)
{
	LOCO_KERNEL("sorted_scatter", st);
	const unsigned g = (unsigned)((Q + 256 * kScatterItems - 1) / (256 * kScatterItems));
	if (w.aggregate) sorted_scatter_kernel<NT, true><<<g, 256, 0, st>>>(m, q, Q, leaf, status, w.cursor, w.xs);
	else sorted_scatter_kernel<NT, false><<<g, 256, 0, st>>>(m, q, Q, leaf, status, w.cursor, w.xs);
}
template <int NT, bool CUBIC, bool EXACT>
void run_eval(const DeviceModel &m, int64_t Q, const SortedWork &w, double *out, int sm_count, bool fp32, int deriv_dir, cudaStream_t st)
{
	const unsigned g = (unsigned)std::min<int64_t>((Q + 255) / 256, (int64_t)sm_count * 8);
	LOCO_KERNEL("sorted_eval", st);
	if constexpr (CUBIC) {
		if (deriv_dir >= 0) { sorted_eval_kernel<NT, true, EXACT, false, false, true><<<g, 256, 0, st>>>(m, w.xs, w.cursor, out, deriv_dir); return; }
	}
	if constexpr (NT > 0) {
		if (m.tensor_F > 0 && fp32 && m.tl_vals_f32) { sorted_eval_kernel<NT, CUBIC, EXACT, true, true><<<g, 256, 0, st>>>(m, w.xs, w.cursor, out); return; }
		if (m.tensor_F > 0) { sorted_eval_kernel<NT, CUBIC, EXACT, true, false><<<g, 256, 0, st>>>(m, w.xs, w.cursor, out); return; }
	}
	sorted_eval_kernel<NT, CUBIC, EXACT, false, false><<<g, 256, 0, st>>>(m, w.xs, w.cursor, out);
}

} // namespace

size_t sorted_work_bytes(const DeviceModel &m, int64_t chunk)
{
	const size_t ints = (size_t)sort_buckets(m) * 2 + 1 + 3;
	return ((ints * sizeof(int32_t) + 127) & ~(size_t)127) + (size_t)chunk * rec_doubles(m.N) * sizeof(double) + 256;
}

SortedWork carve_sorted_work(const DeviceModel &m, void *base)
{
	SortedWork w;
	w.stats = nullptr; w.stats_cap = 0x7fffffff; w.aggregate = true;
	int32_t *p = reinterpret_cast<int32_t *>(base);
	w.count = p; p += sort_buckets(m);
	w.cursor = p; p += sort_buckets(m) + 1;
	w.done = reinterpret_cast<unsigned *>(p); p += 3;
	const size_t ints = (size_t)(p - reinterpret_cast<int32_t *>(base));
	w.xs = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(base) + ((ints * sizeof(int32_t) + 127) & ~(size_t)127));
	return w;
}

// the histogram and the arrival counter must be zero before the first chunk; every count kernel leaves them zero again
int launch_sorted_reset(const DeviceModel &m, const SortedWork &w, cudaStream_t st)
{
	cudaMemsetAsync(w.count, 0, ((size_t)sort_buckets(m) * 2 + 1 + 3) * sizeof(int32_t), st);
	return 0;
}

// deriv_dir >= 0: derivatives along that cube direction; callers guarantee a cubic basis and generic (non-tensor) leaves
int launch_eval_scalar_sorted_chunk(const DeviceModel &m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, const SortedWork &w,
                                    int sm_count, bool fp32, cudaStream_t st, int deriv_dir)
{
	// per launch, not once per process: the attribute belongs to the current device's context and one process may hold
	// models on several GPUs
	cudaFuncSetAttribute(sorted_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((kScanSeg + kScanSeg / 32) * sizeof(int)));
	if (Q <= 0) return 0;
	const bool cubic = m.basis == CUBIC_BSPLINE, exact = m.exact_rcp != 0;
#define LOCO_SCASE(NT)                                                                                       \
	case NT:                                                                                                 \
		run_count<NT>(m, q, Q, leaf, status, out, w, st);                                                    \
		run_scatter<NT>(m, q, Q, leaf, status, w, st);                                                       \
		if (cubic) { if (exact) run_eval<NT, true, true>(m, Q, w, out, sm_count, fp32, deriv_dir, st); else run_eval<NT, true, false>(m, Q, w, out, sm_count, fp32, deriv_dir, st); } \
		else { if (exact) run_eval<NT, false, true>(m, Q, w, out, sm_count, fp32, deriv_dir, st); else run_eval<NT, false, false>(m, Q, w, out, sm_count, fp32, deriv_dir, st); }     \
		break;
	switch (m.N <= kMaxTemplateN ? m.N : 0) {
		LOCO_SCASE(0) LOCO_SCASE(1) LOCO_SCASE(2) LOCO_SCASE(3) LOCO_SCASE(4) LOCO_SCASE(5) LOCO_SCASE(6) LOCO_SCASE(7) LOCO_SCASE(8)
	}
#undef LOCO_SCASE
	return 4;
}

} // namespace loco
