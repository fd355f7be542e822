// tma_ptx.cuh -- thin PTX wrappers shared by the streaming kernels: mbarrier, TMA tensor loads,
// system-scope flags, vector shared-memory loads.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace fdtd {

// ---------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Blocks until the phase with this parity has completed.  The suspend-time hint lets the hardware park the warp
// until the barrier flips (or the hint expires) instead of returning after a short system-defined time: without
// it a quarter of all issued instructions of the two-step kernel were try_wait retries of warps that were ahead
// (profiles/r01_ncu_tb2_32x64_exact0.txt), competing for issue slots with the warps they were waiting for.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity), "r"(0x989680)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                            int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                            int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// The same boxes brought into L2 only (no shared-memory slot, no barrier): issued a few stages ahead of the load itself, so that
// the load's latency is an L2 hit and the DRAM latency (and its jitter) is hidden without a deeper shared-memory ring.
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// Spin (bounded) until *flag >= want; the neighbour stores the flag with release.sys after its
// boundary planes have landed in this GPU's memory.  On timeout mark *err and carry on.
__device__ __forceinline__ void wait_flag(const int *flag, int want, int *err)
{
    int v, tries = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= want) break;
        __nanosleep(200);
        if (++tries > 5000000) {  // > 1 s: the neighbour is gone
            atomicExch(err, 1);
            break;
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");  // the ghost planes are read by TMA (async proxy)
}
// The same for the 3 x 3 tiles around (ty, tz) of a per-tile flag array, polled by lanes 0..8 of ONE converged warp
// (indices clamped at the edges of the tile grid); every lane of the warp may rely on the neighbours' data afterwards.
__device__ __forceinline__ void wait_tiles(const int *flags, int ty, int tz, int tiles_y, int tiles_z, int want, int *err)
{
    const int lane = threadIdx.x & 31;
    if (lane < 9) {
        const int y = min(max(ty + lane / 3 - 1, 0), tiles_y - 1), z = min(max(tz + lane % 3 - 1, 0), tiles_z - 1);
        wait_flag(flags + y * tiles_z + z, want, err);
    }
    __syncwarp();
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void raise_flag(int *flag, int value)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
// The same without its own fence, for a thread that has just executed __threadfence_system(): every st.release.sys costs
// another MEMBAR.SYS (~2 us), and a boundary CTA raised three flags behind one fence (7 % of a pass on 128-plane slabs).
__device__ __forceinline__ void raise_flag_fenced(int *flag, int value)
{
    asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
__device__ __forceinline__ float4 lds128(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float2 lds64(const float *p) { return *reinterpret_cast<const float2 *>(p); }


}  // namespace fdtd
