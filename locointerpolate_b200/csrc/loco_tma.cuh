// loco_tma.cuh -- mbarrier / TMA bulk-copy helpers (PTX ISA: cp.async.bulk, mbarrier); SASS: UBLKCP, SYNCS
#pragma once
#include <cstdint>

namespace loco {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t done = 0;
	const uint32_t addr = smem_u32(bar);
	while (!done) {
		asm volatile(
		    "{\n\t.reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(done) : "r"(addr), "r"(parity) : "memory");
	}
}
// plain arrival (release semantics at CTA scope: the arriving thread's earlier shared-memory writes are visible to
// threads that observe the phase completion with mbar_wait)
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// global -> shared bulk copy executed by the TMA engine; size and both addresses are multiples of 16 B
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// the same with an L2 eviction-priority hint (createpolicy): value rows that every CTA re-reads are kept with evict_last
// while the result stream passes through the L2 with evict_first stores
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
// 128-bit store that tells the L2 the line will not be read again
__device__ __forceinline__ void st_stream_f64x2(double *p, double a, double b, uint64_t policy)
{
	asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(policy) : "memory");
}
// 256-bit streaming store (sm_100: st.global.v4.f64), 32-byte aligned
__device__ __forceinline__ void st_stream_f64x4(double *p, double a, double b, double c, double d)
{
	asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
// 256-bit store variants for cache-policy experiments: default write-back, and with an explicit L2 policy
__device__ __forceinline__ void st_plain_f64x4(double *p, double a, double b, double c, double d)
{
	asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void st_policy_f64x4(double *p, double a, double b, double c, double d, uint64_t policy)
{
	asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_no_allocate()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
// shared -> global bulk copy executed by the TMA engine (bulk async-group completion); size and both addresses multiples of 16 B
__device__ __forceinline__ void tma_bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void tma_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make this thread's earlier shared-memory writes visible to the async proxy (the TMA engine)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

} // namespace loco
