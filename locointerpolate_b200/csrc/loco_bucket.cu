// loco_bucket.cu -- single-pass leaf bucketing for scalar functions on tensor-product leaves (path 5).
//
// The exact counting sort of loco_sorted.cu reads every query twice (histogram, then scatter) because a bucket's
// position is only known after the scan.  Here every leaf owns a FIXED-capacity bucket per chunk
// (capacity = mean + 6 sigma of a uniform batch), so one pass suffices:
//   1. bucket_scatter   map -> locate (locate grid) -> containment; leaf[] / status[]; slot = atomicAdd(count[leaf]);
//                       slot < capacity: the 32-byte record {cube coordinates, query index} goes to bucket[leaf][slot];
//                       otherwise the query index is appended to an overflow list
//   2. bucket_eval      one WARP per leaf (persistent): the leaf's block (N*F centres, N reciprocals, F^N values) is copied
//                       into the warp's shared-memory slot once, then the warp walks the bucket 32 records at a time --
//                       coalesced record loads, broadcast shared-memory reads of leaf data, results to out[query index];
//                       the warp clears the leaf's counter for the next chunk
//   3.                  at the end of the same kernel the (normally empty) overflow list goes through the direct per-query evaluation
// A skewed batch only moves queries from step 2 to step 3 (which then finds its leaf data hot in L1); results are
// identical either way because every query is evaluated independently with the same arithmetic.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "loco_basis.cuh"
#include "loco_device.h"
#include "loco_tma.cuh"

namespace loco {

namespace {

__host__ __device__ constexpr int brec_doubles(int N) { return ((N * 8 + 4 + 31) / 32) * 4; }  // N coordinates + int32 index, whole sectors

__device__ __forceinline__ void st_sector_cg(double *p, double a, double b, double c, double d)
{
	asm volatile("st.global.cg.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld_sector_cg(const double *p, double &a, double &b, double &c, double &d)
{
	asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// Every leaf's slot counter sits in its own 128-byte line: the L2 serialises atomics per LINE, and a skewed batch (half of
// the queries in 1 % of the leaves) sends millions of them to the few lines that would hold 32 neighbouring counters each.
constexpr int kCountStride = 32;
constexpr int kItems = 4;  // queries per thread in the scatter pass: independent load -> locate -> atomic -> store chains in flight

// Overflow list (queries whose bucket was full): ONE list-counter atomic per warp for ALL items of its threads -- a skewed
// batch sends half of every warp here, and an append per item (a dependent atomic round trip each) made the scatter pass
// 2.6 x slower on such a batch.  Called by the whole warp, outside divergent code.
template <int ITEMS>
__device__ __forceinline__ void append_overflow(int n_over, const bool (&live)[ITEMS], const int32_t (&slot)[ITEMS], int32_t cap, int64_t i0,
                                                int32_t *__restrict__ ovf, int32_t *n_ovf)
{
	if (!__any_sync(0xffffffffu, n_over > 0)) return;
	const int lane_id = threadIdx.x & 31;
	int incl = n_over;  // inclusive warp prefix sum of the per-thread counts
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		const int v = __shfl_up_sync(0xffffffffu, incl, off);
		if (lane_id >= off) incl += v;
	}
	const int total = __shfl_sync(0xffffffffu, incl, 31);
	int base = 0;
	if (lane_id == 0) base = atomicAdd(n_ovf, total);
	base = __shfl_sync(0xffffffffu, base, 0) + incl - n_over;
#pragma unroll
	for (int k = 0; k < ITEMS; k++)
		if (live[k] && slot[k] >= cap) ovf[base++] = (int32_t)(i0 + k * 256);
}

// AGG: lanes of a warp that ask for a slot in the SAME leaf's bucket share one atomicAdd (structured batches -- lattice
// sweeps, trajectories -- put runs of consecutive queries into one leaf; their same-address atomics would serialise in the L2)
template <int NT, bool AGG>
__global__ void __launch_bounds__(256) bucket_scatter_kernel(DeviceModel m, const double *__restrict__ q, int64_t Q, int32_t *__restrict__ leaf_out,
                                                              uint8_t *__restrict__ status_out, double *__restrict__ out, int32_t *count, int32_t cap,
                                                              double *__restrict__ rec, int32_t *ovf, int32_t *n_ovf, int dbg)
{
	constexpr int RS = brec_doubles(NT);
	const int64_t i0 = (int64_t)blockIdx.x * (256 * kItems) + threadIdx.x;
	// Phases instead of one query after the other: a thread's kItems dependent chains
	// (query -> locate-grid entry -> bucket counter -> record) advance together, one memory round trip per phase.
	Coord<NT> x[kItems];
	uint32_t code[kItems];
	bool live[kItems];
	{
		double p[kItems][NT];
#pragma unroll
		for (int k = 0; k < kItems; k++) {
			const int64_t i = i0 + k * 256;
			live[k] = i < Q;
			if (live[k]) {
#pragma unroll
				for (int d = 0; d < NT; d++) p[k][d] = __ldcs(q + i * NT + d);
			}
		}
#pragma unroll
		for (int k = 0; k < kItems; k++) {
			code[k] = 0;
			if (live[k]) {
				map_to_cube<NT>(m, p[k], x[k]);
				code[k] = lut_lookup<NT>(m, x[k]);
			}
		}
	}
	int32_t leaf[kItems], slot[kItems];
#pragma unroll
	for (int k = 0; k < kItems; k++) {
		if (!live[k]) continue;
		const int64_t i = i0 + k * 256;
		leaf[k] = descend_from<NT>(m, x[k], code[k]);
		const uint8_t st = located_status<NT>(m, x[k], leaf[k]);
		if (!(dbg & 4)) { __stcs(leaf_out + i, leaf[k]); status_out[i] = st; }
		if (st != ST_OK) { out[i] = __longlong_as_double(0x7ff8000000000000LL); live[k] = false; }
	}
#pragma unroll
	for (int k = 0; k < kItems; k++) {
		if constexpr (AGG) {
			const int lane_id = threadIdx.x & 31;
			const int32_t key = live[k] ? leaf[k] : -1 - lane_id;  // lanes without a query match nobody
			const unsigned peers = __match_any_sync(0xffffffffu, key);
			const int leader = __ffs(peers) - 1;
			int base = 0;
			if (dbg & 2) base = (int)((blockIdx.x * 7 + k) % 100);  // EXPERIMENT: no atomic
			else
			if (live[k] && lane_id == leader) base = atomicAdd(count + (int64_t)leaf[k] * kCountStride, __popc(peers));
			base = __shfl_sync(0xffffffffu, base, leader);
			slot[k] = base + __popc(peers & ((1u << lane_id) - 1));
		} else {
			if (live[k]) slot[k] = atomicAdd(count + (int64_t)leaf[k] * kCountStride, 1);
		}
	}
	int n_over = 0;  // this thread's queries whose bucket was full
#pragma unroll
	for (int k = 0; k < kItems; k++) {
		if (!live[k]) continue;
		const int64_t i = i0 + k * 256;
		if (slot[k] < cap) {
			double v[RS];
#pragma unroll
			for (int d = 0; d < RS; d++) v[d] = d < NT ? x[k].get(d) : 0.0;
			v[RS - 1] = __longlong_as_double((long long)(int32_t)i);
			double *row = rec + ((int64_t)leaf[k] * cap + slot[k]) * RS;
			if (dbg & 8) row = rec + ((int64_t)i & 0xffff) * RS;  // EXPERIMENT: records into a small (2 MB) window instead of their buckets
			if (!(dbg & 1)) {
#pragma unroll
			for (int d = 0; d < RS; d += 4) st_sector_cg(row + d, v[d], v[d + 1], v[d + 2], v[d + 3]);
			}
		} else n_over++;
	}
	append_overflow<kItems>(n_over, live, slot, cap, i0, ovf, n_ovf);
}

// The same pass for the COMMON model shape, checked on the host (scatter_fast_ok): every domain width a power of two, a
// locate grid with exactly uniform dyadic boundaries in every dimension, leaf boxes implied by their split planes and no
// per-leaf error status.  ncu of the general kernel (profiles/r02h_c2_*): 273 thread instructions per query, a third of them
// uniform-datapath loads, tests and branches on model flags that are constant for the whole launch; here the flags are
// compile-time facts and the arithmetic is the same, operation for operation (map_to_cube / lut_lookup / located_status).
template <int NT, bool AGG>
__global__ void __launch_bounds__(256) bucket_scatter_fast_kernel(DeviceModel m, const double *__restrict__ q, int64_t Q, int32_t *__restrict__ leaf_out,
                                                                   uint8_t *__restrict__ status_out, double *__restrict__ out, int32_t *count, int32_t cap,
                                                                   double *__restrict__ rec, int32_t *ovf, int32_t *n_ovf)
{
	constexpr int RS = brec_doubles(NT);
	const int64_t i0 = (int64_t)blockIdx.x * (256 * kItems) + threadIdx.x;
	const int lane_id = threadIdx.x & 31;
	int shift[NT], last[NT];
	{
		int sh = 0;
#pragma unroll
		for (int d = 0; d < NT; d++) { shift[d] = sh; last[d] = (1 << m.lut_bits[d]) - 1; sh += m.lut_bits[d]; }
	}
	double x[kItems][NT];
	uint32_t code[kItems];
	bool live[kItems];
	uint8_t st[kItems];
#pragma unroll
	for (int k = 0; k < kItems; k++) {
		const int64_t i = i0 + k * 256;
		live[k] = i < Q;
#pragma unroll
		for (int d = 0; d < NT; d++) x[k][d] = live[k] ? __ldcs(q + i * NT + d) : 0.0;
	}
#pragma unroll
	for (int k = 0; k < kItems; k++) {
		uint32_t idx = 0;
		bool ok = true, outside = false;
#pragma unroll
		for (int d = 0; d < NT; d++) {
			// LCFunction::mapToStandardHypercube with the exact reciprocal of a power-of-two width
			const double xv = __fma_rn(__dmul_rn(__dsub_rn(x[k][d], m.dom_lo[d]), m.dom_rcp[d]), 2.0, -1.0);
			x[k][d] = xv;
			// locate grid, uniform dyadic boundaries (see lut_lookup)
			int i = (int)fmin(fmax((xv - m.lut_lo[d]) * m.lut_inv[d], 0.0), (double)last[d]);
			const double b0 = fma((double)i, m.lut_h[d], m.lut_lo[d]);
			if (i > 0 && xv < b0) i--;
			ok &= xv == xv;
			idx |= (uint32_t)i << shift[d];
			// evalPametersAreInCell against the root box (path_consistent trees)
			outside |= (xv < m.root_lo_eps[d]) || (xv > m.root_hi_eps[d]);
		}
		code[k] = (live[k] && ok) ? __ldg(m.lut + idx) : 0u;
		st[k] = outside ? (uint8_t)ST_OUTSIDE_CELL : (uint8_t)ST_OK;
	}
	int32_t leaf[kItems], slot[kItems];
#pragma unroll
	for (int k = 0; k < kItems; k++) {
		leaf[k] = 0;
		if (!live[k]) continue;
		const int64_t i = i0 + k * 256;
		Coord<NT> xc;
#pragma unroll
		for (int d = 0; d < NT; d++) xc.set(d, x[k][d]);
		leaf[k] = descend_from<NT>(m, xc, code[k]);
		__stcs(leaf_out + i, leaf[k]);
		status_out[i] = st[k];
		if (st[k] != ST_OK) { out[i] = __longlong_as_double(0x7ff8000000000000LL); live[k] = false; }
	}
#pragma unroll
	for (int k = 0; k < kItems; k++) {
		if constexpr (AGG) {
			const int32_t key = live[k] ? leaf[k] : -1 - lane_id;
			const unsigned peers = __match_any_sync(0xffffffffu, key);
			const int leader = __ffs(peers) - 1;
			int base = 0;
			if (live[k] && lane_id == leader) base = atomicAdd(count + (int64_t)leaf[k] * kCountStride, __popc(peers));
			base = __shfl_sync(0xffffffffu, base, leader);
			slot[k] = base + __popc(peers & ((1u << lane_id) - 1));
		} else {
			if (live[k]) slot[k] = atomicAdd(count + (int64_t)leaf[k] * kCountStride, 1);
		}
	}
	int n_over = 0;
#pragma unroll
	for (int k = 0; k < kItems; k++) {
		if (!live[k]) continue;
		const int64_t i = i0 + k * 256;
		if (slot[k] < cap) {
			double v[RS];
#pragma unroll
			for (int d = 0; d < RS; d++) v[d] = d < NT ? x[k][d] : 0.0;
			v[RS - 1] = __longlong_as_double((long long)(int32_t)i);
			double *row = rec + ((int64_t)leaf[k] * cap + slot[k]) * RS;
#pragma unroll
			for (int d = 0; d < RS; d += 4) st_sector_cg(row + d, v[d], v[d + 1], v[d + 2], v[d + 3]);
		} else n_over++;
	}
	append_overflow<kItems>(n_over, live, slot, cap, i0, ovf, n_ovf);
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// one warp per leaf; 8 warps per CTA, each with its own shared-memory copies of its current and (DBL) its next leaf block.
// The three dependent round trips of a leaf (counter -> block + first records -> arithmetic) are software-pipelined ACROSS
// leaves: while leaf l is evaluated, the counter of leaf l' = l + #warps has been read, its block is on its way into the
// other shared-memory slot (cp.async) and its first 32 records are loaded during the last record group of l.
// DERIV (fp64): the derivative along cube direction `dir` instead of the value (LCSampledFunction::evalDeriv).
template <int NT, bool CUBIC, bool EXACT, bool FP32, bool DBL, bool DERIV>
__global__ void __launch_bounds__(256, 3) bucket_eval_kernel(DeviceModel m, const double *__restrict__ rec, int32_t *count, int32_t cap, double *__restrict__ out,
                                                              const double *__restrict__ q, const int32_t *__restrict__ leaf_in, const int32_t *__restrict__ ovf,
                                                              const int32_t *__restrict__ n_ovf, int32_t *n_ovf_next, int dir, SkewStats *stats, int64_t n_chunk, int dbg)
{
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		*n_ovf_next = 0;  // the counter of the NEXT chunk on this workspace (nobody reads it now)
		if (stats) { stats->overflow += *n_ovf; stats->total += n_chunk; }  // one writer per workspace slot (its chunks are serialised on one stream)
	}
	constexpr int RS = brec_doubles(NT);
	extern __shared__ __align__(16) double s_blk[];  // [8][DBL ? 2 : 1][tl_stride]  (+ K/2 doubles per slot holding the fp32 values in fp32 mode)
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int n_warps = gridDim.x * 8;
	// the leaf's polynomial block where the model has one (values only; derivatives and the fp32 mode use the classic block)
	const bool poly = !FP32 && !DERIV && m.tl_poly != nullptr;
	const int slot_doubles = ((FP32 ? m.tl_voff + (m.tensor_K + 1) / 2 : (poly ? m.tp_stride : m.tl_stride)) + 1) & ~1;
	double *slot0 = s_blk + (size_t)warp * (DBL ? 2 : 1) * slot_doubles;
	auto issue_block = [&](int32_t lf, double *dst) {
		const char *src = poly ? reinterpret_cast<const char *>(m.tl_poly + (int64_t)lf * m.tp_stride)
		                       : reinterpret_cast<const char *>(m.tl_block + (int64_t)lf * m.tl_stride);
		const int n16 = (FP32 ? m.tl_voff : (poly ? m.tp_stride : m.tl_stride)) / 2;
		for (int e = lane; e < n16; e += 32) cp_async16(dst + 2 * e, src + 16 * e);
		if (FP32) {
			const char *vs = reinterpret_cast<const char *>(m.tl_vals_f32 + (int64_t)lf * m.tensor_K);
			for (int e = lane; e < m.tensor_K / 4; e += 32) cp_async16(dst + m.tl_voff + 2 * e, vs + 16 * e);
		}
		cp_async_commit();
	};
	double nv[RS];
	auto load_rec = [&](const double *row) {
#pragma unroll
		for (int d = 0; d < RS; d += 4) ld_sector_cg(row + d, nv[d], nv[d + 1], nv[d + 2], nv[d + 3]);
	};
	int32_t leaf = blockIdx.x * 8 + warp, n = 0;
	int buf = 0;
	if (leaf < m.L) {
		n = min(count[(int64_t)leaf * kCountStride], cap);
		issue_block(leaf, slot0);
		load_rec(rec + ((int64_t)leaf * cap + lane) * RS);  // a bucket has >= 32 slots; lanes beyond n read unused memory
	}
	while (leaf < m.L) {
		const int32_t next = leaf + n_warps;
		const bool more = next < m.L;
		int32_t n_next = 0;
		if (more) n_next = count[(int64_t)next * kCountStride];
		if (DBL && more) { issue_block(next, slot0 + (buf ^ 1) * slot_doubles); cp_async_wait<1>(); }
		else cp_async_wait<0>();
		__syncwarp();
		if (lane == 0 && n > 0) count[(int64_t)leaf * kCountStride] = 0;  // ready for the next chunk (the scatter pass of this chunk has finished)
		const double *blk = slot0 + (DBL ? buf : 0) * slot_doubles;
		const double *bucket = rec + (int64_t)leaf * cap * RS;
		const double *nbucket = rec + ((int64_t)next * cap + lane) * RS;
		if (n == 0 && more) load_rec(nbucket);
		// ---- 32 records at a time, the next group's (or the next leaf's first) records prefetched while this group is evaluated
		for (int g0 = 0; g0 < n; g0 += 32) {
			const int s = g0 + lane;
			Coord<NT> x;
#pragma unroll
			for (int d = 0; d < NT; d++) x.set(d, nv[d]);
			const int32_t qi = (int32_t)__double_as_longlong(nv[RS - 1]);
			if (g0 + 32 < n) { if (s + 32 < n) load_rec(bucket + (int64_t)(s + 32) * RS); }
			else if (more) load_rec(nbucket);
			if (s < n) {
				double r;
				if constexpr (DERIV) r = tensor_eval_scalar_deriv<NT, CUBIC, EXACT>(m, blk, x, dir);
				else if constexpr (FP32) r = tensor_eval_scalar_f32<NT, CUBIC, EXACT>(m, blk, reinterpret_cast<const float *>(blk + m.tl_voff), x);
				else if (dbg & 2) r = x.get(0);  // EXPERIMENT: no arithmetic
				else r = poly ? tensor_leaf_value<NT, CUBIC, EXACT>(m, nullptr, blk, leaf, x) : tensor_eval_scalar<NT, CUBIC, EXACT>(m, blk, x);
				if (!(dbg & 1) || r == 12345.678) out[qi] = r;
			}
		}
		__syncwarp();  // every lane is done with this leaf's block before its slot is written again
		if (!DBL && more) issue_block(next, slot0);
		leaf = next;
		n = min(n_next, cap);
		buf ^= 1;
	}
	// ---- the overflow list (queries whose bucket was full; normally empty): one thread each, straight from the global leaf blocks
	const int32_t n_over = *n_ovf;
	for (int32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_over; e += gridDim.x * blockDim.x) {
		const int32_t i = ovf[e];
		Coord<NT> x;
		double p[NT];
#pragma unroll
		for (int d = 0; d < NT; d++) p[d] = q[(int64_t)i * NT + d];
		map_to_cube<NT>(m, p, x);
		const double *gblk = m.tl_block + (int64_t)leaf_in[i] * m.tl_stride;
		if constexpr (DERIV) out[i] = tensor_eval_scalar_deriv<NT, CUBIC, EXACT>(m, gblk, x, dir);
		else out[i] = tensor_leaf_value<NT, CUBIC, EXACT>(m, gblk, poly ? m.tl_poly + (int64_t)leaf_in[i] * m.tp_stride : nullptr, leaf_in[i], x);
	}
}

// ---------------------------------------------------------------------------------------------------------------------
// Evaluation, second form (values in fp64; the kernel above stays for derivatives and the fp32 mode).
// ncu of the kernel above (profiles/r02h_c2_*): 32 % of all stall samples sit on the first use of the record a lane
// prefetched ONE group ahead, the fp64 pipe is 28 % busy -- the warp waits for DRAM.  Here a warp's work item is up to RB
// records of one leaf, and the TMA engine brings the NEXT item's records and leaf block into the warp's other shared-memory
// buffer (two cp.async.bulk on one mbarrier) while the warp evaluates the current one entirely from shared memory: the
// DRAM round trip of an item is hidden behind ~RB evaluations instead of 32.
template <int NT, bool CUBIC, bool EXACT, int RB>
__global__ void __launch_bounds__(256, 3) bucket_eval_bulk_kernel(DeviceModel m, const double *__restrict__ rec, int32_t *count, int32_t cap, double *__restrict__ out,
                                                                   const double *__restrict__ q, const int32_t *__restrict__ leaf_in, const int32_t *__restrict__ ovf,
                                                                   const int32_t *__restrict__ n_ovf, int32_t *n_ovf_next, SkewStats *stats, int64_t n_chunk)
{
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		*n_ovf_next = 0;
		if (stats) { stats->overflow += *n_ovf; stats->total += n_chunk; }
	}
	constexpr int RS = brec_doubles(NT);
	extern __shared__ __align__(128) double s_buf[];  // [8 warps][2 buffers][blk_doubles + RB * RS]
	__shared__ __align__(8) uint64_t bar[8][2];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int n_warps = gridDim.x * 8;
	const bool poly = m.tl_poly != nullptr;
	const int blk_copy = poly ? m.tp_stride : m.tl_stride;  // doubles copied per leaf block (even: a multiple of 16 bytes)
	const int blk_doubles = (blk_copy + 3) & ~3;            // records start 32-byte aligned behind it
	const int buf_doubles = blk_doubles + RB * RS;
	double *my = s_buf + (size_t)warp * 2 * buf_doubles;
	if (lane == 0) { mbar_init(&bar[warp][0], 1); mbar_init(&bar[warp][1], 1); }
	__syncwarp();
	// item = (leaf, first record, number of records <= RB); a leaf with n records is cut into ceil(n / RB) equal items
	struct Item { int32_t leaf, r0, cnt, left; };  // left: records of this leaf after this item
	auto first_item = [&](int32_t leaf, int32_t n) -> Item {
		const int parts = (n + RB - 1) / RB;
		const int len = parts ? (n + parts - 1) / parts : 0;
		return Item{leaf, 0, len, n - len};
	};
	auto issue = [&](const Item &it, int b) {
		if (it.cnt <= 0) return;
		double *dst = my + (size_t)b * buf_doubles;
		if (lane == 0) {
			const uint32_t blk_bytes = (uint32_t)blk_copy * 8, rec_bytes = (uint32_t)it.cnt * RS * 8;
			mbar_expect_tx(&bar[warp][b], blk_bytes + rec_bytes);
			tma_bulk_g2s(dst, poly ? m.tl_poly + (int64_t)it.leaf * m.tp_stride : m.tl_block + (int64_t)it.leaf * m.tl_stride, blk_bytes, &bar[warp][b]);
			tma_bulk_g2s(dst + blk_doubles, rec + ((int64_t)it.leaf * cap + it.r0) * RS, rec_bytes, &bar[warp][b]);
		}
	};
	int32_t leaf = blockIdx.x * 8 + warp;
	Item cur{leaf, 0, 0, 0};
	if (leaf < m.L) {
		const int32_t n = min(count[(int64_t)leaf * kCountStride], cap);
		if (lane == 0 && n > 0) count[(int64_t)leaf * kCountStride] = 0;  // ready for the next chunk (the scatter pass of this chunk has finished)
		cur = first_item(leaf, n);
	}
	// the count of the leaf AFTER the next one is read one item early (its global round trip is hidden as well)
	int32_t n_ahead = 0, leaf_ahead = leaf + n_warps;
	if (leaf_ahead < m.L) n_ahead = count[(int64_t)leaf_ahead * kCountStride];
	uint32_t phase = 0;  // bit b: parity to wait for on bar[warp][b]
	int b = 0;
	issue(cur, 0);
	while (cur.leaf < m.L) {
		// ---- next item: the rest of this leaf, or the next leaf of this warp
		Item nxt;
		if (cur.left > 0) {
			const int len = min(cur.cnt, cur.left);
			nxt = Item{cur.leaf, cur.r0 + cur.cnt, len, cur.left - len};
		} else {
			const int32_t nl = leaf_ahead;
			int32_t n = 0;
			if (nl < m.L) {
				n = min(n_ahead, cap);
				if (lane == 0 && n > 0) count[(int64_t)nl * kCountStride] = 0;
			}
			nxt = first_item(nl, n);
			leaf_ahead = nl + n_warps;
			n_ahead = leaf_ahead < m.L ? count[(int64_t)leaf_ahead * kCountStride] : 0;
		}
		if (nxt.leaf < m.L) issue(nxt, b ^ 1);
		// ---- this item, from shared memory
		if (cur.cnt > 0) {
			mbar_wait(&bar[warp][b], (phase >> b) & 1u);
			phase ^= 1u << b;
			const double *blk = my + (size_t)b * buf_doubles;
			const double *recs = blk + blk_doubles;
			for (int g0 = 0; g0 < cur.cnt; g0 += 32) {
				const int s = g0 + lane;
				if (s < cur.cnt) {
					const double4 *r4 = reinterpret_cast<const double4 *>(recs + (size_t)s * RS);
					Coord<NT> x;
					double v[RS];
#pragma unroll
					for (int d = 0; d < RS; d += 4) { const double4 t = r4[d / 4]; v[d] = t.x; v[d + 1] = t.y; v[d + 2] = t.z; v[d + 3] = t.w; }
#pragma unroll
					for (int d = 0; d < NT; d++) x.set(d, v[d]);
					const int32_t qi = (int32_t)__double_as_longlong(v[RS - 1]);
					out[qi] = poly ? tensor_leaf_value<NT, CUBIC, EXACT>(m, nullptr, blk, cur.leaf, x) : tensor_eval_scalar<NT, CUBIC, EXACT>(m, blk, x);
				}
			}
			__syncwarp();  // every lane is done with this buffer before a later copy lands in it
		}
		cur = nxt;
		b ^= 1;
	}
	// ---- the overflow list (queries whose bucket was full; normally empty): one thread each, straight from the global leaf blocks
	const int32_t n_over = *n_ovf;
	for (int32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_over; e += gridDim.x * blockDim.x) {
		const int32_t i = ovf[e];
		Coord<NT> x;
		double p[NT];
#pragma unroll
		for (int d = 0; d < NT; d++) p[d] = q[(int64_t)i * NT + d];
		map_to_cube<NT>(m, p, x);
		const double *gblk = m.tl_block + (int64_t)leaf_in[i] * m.tl_stride;
		out[i] = tensor_leaf_value<NT, CUBIC, EXACT>(m, gblk, poly ? m.tl_poly + (int64_t)leaf_in[i] * m.tp_stride : nullptr, leaf_in[i], x);
	}
}

// host-side precondition of bucket_scatter_fast_kernel
inline bool scatter_fast_ok(const DeviceModel &m)
{
	if (!m.lut || m.n_inner == 0 || !m.path_consistent || m.any_leaf_status) return false;
	for (int d = 0; d < m.N; d++)
		if (!((m.dom_exact_mask >> d) & 1u) || m.lut_bits[d] <= 0 || !((m.lut_uniform_mask >> d) & 1u)) return false;
	return true;
}

template <int NT, bool CUBIC, bool EXACT>
void run_chunk(const DeviceModel &m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, const BucketWork &w, int sm_count, bool fp32,
               int parity, int deriv_dir, cudaStream_t st)
{
	// overflow counters alternate between the chunks of a workspace: the evaluation of chunk j clears the counter of chunk j+1
	int32_t *n_ovf = w.n_ovf + (parity & 1), *n_ovf_next = w.n_ovf + ((parity + 1) & 1);
	{
		LOCO_KERNEL("bucket_scatter", st);
		const unsigned g = (unsigned)((Q + 256 * kItems - 1) / (256 * kItems));
		static const int dbg = getenv("LOCO_SCATTER_DBG") ? atoi(getenv("LOCO_SCATTER_DBG")) : 0;  // measurement experiments only
		if (w.fast && dbg == 0 && scatter_fast_ok(m)) {
			if (w.aggregate) bucket_scatter_fast_kernel<NT, true><<<g, 256, 0, st>>>(m, q, Q, leaf, status, out, w.count, w.cap, w.rec, w.ovf, n_ovf);
			else bucket_scatter_fast_kernel<NT, false><<<g, 256, 0, st>>>(m, q, Q, leaf, status, out, w.count, w.cap, w.rec, w.ovf, n_ovf);
		} else
		if (w.aggregate) bucket_scatter_kernel<NT, true><<<g, 256, 0, st>>>(m, q, Q, leaf, status, out, w.count, w.cap, w.rec, w.ovf, n_ovf, dbg);
		else bucket_scatter_kernel<NT, false><<<g, 256, 0, st>>>(m, q, Q, leaf, status, out, w.count, w.cap, w.rec, w.ovf, n_ovf, dbg);
	}
	{
		LOCO_KERNEL("bucket_eval", st);
		const bool f32 = fp32 && m.tl_vals_f32 && (m.tensor_K % 4) == 0 && deriv_dir < 0;
		if (w.bulk && !f32 && deriv_dir < 0) {
			constexpr int RB = 96;
			constexpr int RS = brec_doubles(NT);
			const int blk_doubles = ((m.tl_poly ? m.tp_stride : m.tl_stride) + 3) & ~3;
			const size_t smem = (size_t)8 * 2 * (blk_doubles + RB * RS) * sizeof(double);
			if (smem <= 200 * 1024) {
				auto kern = bucket_eval_bulk_kernel<NT, CUBIC, EXACT, RB>;
				cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
				const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(3, (220 * 1024) / (smem + 1024)));
				const int grid = std::min(sm_count * per_sm, (m.L + 7) / 8);
				kern<<<grid, 256, smem, st>>>(m, w.rec, w.count, w.cap, out, q, leaf, w.ovf, n_ovf, n_ovf_next, w.stats, Q);
				return;
			}
		}
		const bool poly = !f32 && deriv_dir < 0 && m.tl_poly != nullptr;
		const int slot_doubles = ((f32 ? m.tl_voff + (m.tensor_K + 1) / 2 : (poly ? m.tp_stride : m.tl_stride)) + 1) & ~1;
		const bool dbl = (size_t)16 * slot_doubles * sizeof(double) <= 72 * 1024;  // two slots per warp while 3 CTAs per SM still fit
		const size_t smem = (size_t)8 * (dbl ? 2 : 1) * slot_doubles * sizeof(double);
		const int grid = std::min(sm_count * 3, (m.L + 7) / 8);
		auto launch = [&](auto kern) {
			if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			static const int edbg = getenv("LOCO_EVAL_DBG") ? atoi(getenv("LOCO_EVAL_DBG")) : 0;  // measurement experiments only
			kern<<<grid, 256, smem, st>>>(m, w.rec, w.count, w.cap, out, q, leaf, w.ovf, n_ovf, n_ovf_next, deriv_dir, w.stats, Q, edbg);
		};
		if (deriv_dir >= 0) {
			if (dbl) launch(bucket_eval_kernel<NT, CUBIC, EXACT, false, true, true>); else launch(bucket_eval_kernel<NT, CUBIC, EXACT, false, false, true>);
			return;
		}
		if (f32) { if (dbl) launch(bucket_eval_kernel<NT, CUBIC, EXACT, true, true, false>); else launch(bucket_eval_kernel<NT, CUBIC, EXACT, true, false, false>); }
		else { if (dbl) launch(bucket_eval_kernel<NT, CUBIC, EXACT, false, true, false>); else launch(bucket_eval_kernel<NT, CUBIC, EXACT, false, false, false>); }
	}
}

} // namespace

// bucket capacity for a chunk of `chunk` queries over L leaves: mean + 6 sigma of a uniform batch, a multiple of 32
int bucket_capacity(const DeviceModel &m, int64_t chunk)
{
	const double mean = (double)chunk / (double)std::max(1, m.L);
	const int cap = (int)std::ceil(mean + 6.0 * std::sqrt(mean) + 16.0);
	return (cap + 31) & ~31;
}
bool bucket_path_supported(const DeviceModel &m) { return m.D == 1 && m.tensor_F > 0 && m.N >= 1 && m.N <= kMaxTemplateN && (size_t)m.tl_stride * 8 * 8 <= 160 * 1024; }

size_t bucket_work_bytes(const DeviceModel &m, int64_t chunk)
{
	const size_t ints = (size_t)m.L * kCountStride + 4 + (size_t)chunk;
	return ((ints * sizeof(int32_t) + 127) & ~(size_t)127) + (size_t)m.L * bucket_capacity(m, chunk) * brec_doubles(m.N) * sizeof(double) + 256;
}
BucketWork carve_bucket_work(const DeviceModel &m, int64_t chunk, void *base)
{
	BucketWork w;
	w.stats = nullptr;
	w.aggregate = true;
	w.bulk = true;
	w.fast = true;
	int32_t *p = reinterpret_cast<int32_t *>(base);
	w.count = p; p += (size_t)m.L * kCountStride;
	w.n_ovf = p; p += 4;
	w.ovf = p; p += chunk;
	const size_t ints = (size_t)(p - reinterpret_cast<int32_t *>(base));
	w.rec = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(base) + ((ints * sizeof(int32_t) + 127) & ~(size_t)127));
	w.cap = bucket_capacity(m, chunk);
	return w;
}
int launch_bucket_reset(const DeviceModel &m, const BucketWork &w, cudaStream_t st)
{
	cudaMemsetAsync(w.count, 0, ((size_t)m.L * kCountStride + 4) * sizeof(int32_t), st);
	return 0;
}

int launch_eval_scalar_bucket_chunk(const DeviceModel &m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, const BucketWork &w,
                                    int sm_count, bool fp32, int parity, cudaStream_t st, int deriv_dir)
{
	if (Q <= 0) return 0;
	const bool cubic = m.basis == CUBIC_BSPLINE, exact = m.exact_rcp != 0;
#define LOCO_BCASE(NT)                                                                                                                   \
	case NT:                                                                                                                             \
		if (cubic) { if (exact) run_chunk<NT, true, true>(m, q, Q, out, leaf, status, w, sm_count, fp32, parity, deriv_dir, st);                            \
		             else run_chunk<NT, true, false>(m, q, Q, out, leaf, status, w, sm_count, fp32, parity, deriv_dir, st); }                               \
		else { if (exact) run_chunk<NT, false, true>(m, q, Q, out, leaf, status, w, sm_count, fp32, parity, deriv_dir, st);                                 \
		       else run_chunk<NT, false, false>(m, q, Q, out, leaf, status, w, sm_count, fp32, parity, deriv_dir, st); }                                    \
		break;
	switch (m.N) { LOCO_BCASE(1) LOCO_BCASE(2) LOCO_BCASE(3) LOCO_BCASE(4) LOCO_BCASE(5) LOCO_BCASE(6) LOCO_BCASE(7) LOCO_BCASE(8) default: return 0; }
#undef LOCO_BCASE
	return 2;
}

} // namespace loco
