// ptx_sm100.cuh -- the handful of sm_90+/sm_100a PTX primitives the kernels use:
// mbarrier transaction barriers, TMA bulk copies (1-D and tiled tensor), proxy
// fences.  SASS: UBLKCP (1-D bulk), UTMALDG (tiled tensor), SYNCS (mbarrier).
#pragma once
#include <cuda.h>
#include <cstdint>

namespace firgpu {

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// Make freshly initialised barriers visible to the async (TMA) proxy.
__device__ __forceinline__ void fence_barrier_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Order prior generic-proxy smem accesses before later async-proxy ones.
__device__ __forceinline__ void fence_proxy_async()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
	             : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		"selp.u32 %0, 1, 0, p;\n\t}"
		: "=r"(ok)
		: "r"(bar), "r"(parity)
		: "memory");
	return ok != 0;
}

// try_wait itself blocks for a hardware-defined slice before it reports failure, so a healthy
// wait sees a handful of failures at most.  A bulk copy that faulted never completes its
// transaction count: after MBAR_SPIN_LIMIT failed slices (seconds) the kernel traps, which the
// host sees as a CUDA error (FIR_GPU_ERR_CUDA) instead of a hang.
constexpr uint32_t MBAR_SPIN_LIMIT = 1u << 24;

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	uint32_t spins = 0;
	while (!mbar_try_wait(bar, parity)) {
		if (++spins > MBAR_SPIN_LIMIT) __trap();
	}
}

// 1-D bulk copy global -> shared, completion counted in bytes on `bar`.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
		: "memory");
}

// Tiled 3-D tensor copy global -> shared through a CUtensorMap (box and swizzle
// are in the map).  Out-of-bounds elements arrive as zeros.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint32_t bar)
{
	asm volatile(
		"cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
		"[%0], [%1, {%2, %3, %4}], [%5];"
		::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
		: "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map)
{
	asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

} // namespace firgpu
