// tma.cuh -- mbarrier + TMA (cp.async.bulk[.tensor]) primitives as inline PTX for sm_100a,
// and the 128-byte-swizzle address helpers shared by the TMA-fed kernels.
#pragma once
#include <cuda.h>  // CUtensorMap
#include <cstdint>

namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// non-blocking test (try_wait may suspend the thread for a system-dependent time limit; test_wait never does)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Waits are bounded by WALL TIME, not by a poll count: under time-slicing with another process, ncu kernel
// replay or a preempted context a legitimate wait can take arbitrarily many polls, and a trap poisons the whole
// CUDA context (the caller's arrays included).  Release builds therefore only give up after MBAR_TIMEOUT_NS
// (20 s -- far beyond any preemption gap, far below a driver watchdog-free hang going unnoticed) and then
// trap so that a protocol bug still surfaces as a failed launch rather than a hung GPU; the clock is read
// once per 4096 polls, so the common path is a bare try_wait loop.
#ifndef DFT_MBAR_TIMEOUT_NS
#define DFT_MBAR_TIMEOUT_NS 20000000000ull
#endif
__device__ __forceinline__ void mbar_timeout_trap(uint32_t parity) {
    printf("[dft_b200] mbarrier wait exceeded %llu ns (block %d,%d thread %d parity %u)\n",
           (unsigned long long)DFT_MBAR_TIMEOUT_NS, blockIdx.x, blockIdx.y, threadIdx.x, parity);
    __trap();
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 4095u) == 0) {
            const uint64_t t = globaltimer_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > DFT_MBAR_TIMEOUT_NS) mbar_timeout_trap(parity);
        }
    }
}

// The same with a sleep between polls, for the producer / scanner threads: they share an SM sub-partition's
// issue slots with two MMA warps, and a thread that polls back to back takes cycles from them.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (ns) __nanosleep(ns);
        if ((++spins & 4095u) == 0) {
            const uint64_t t = globaltimer_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > DFT_MBAR_TIMEOUT_NS) mbar_timeout_trap(parity);
        }
    }
}

// 2-D tiled TMA load global -> shared, completion signalled on `bar` (complete_tx::bytes)
__device__ __forceinline__ void load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 3-D tiled TMA load global -> shared
__device__ __forceinline__ void load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// 1-D bulk copy global -> shared (16-byte aligned address and size)
__device__ __forceinline__ void load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// L2 prefetch of a contiguous range (16-byte aligned address, size a multiple of 16)
__device__ __forceinline__ void prefetch_1d(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes) : "memory");
}

// L2 prefetch of one box of a tiled tensor (no shared-memory destination, no completion signal)
__device__ __forceinline__ void prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1)
                 : "memory");
}

__device__ __forceinline__ void prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Byte offset of double element (row, col) inside a TMA box of 16 doubles (128 B) per row written
// with CU_TENSOR_MAP_SWIZZLE_128B into a 1024-byte aligned buffer: the 16-byte chunk index
// (col/2) is XORed with (row mod 8).
__device__ __forceinline__ uint32_t swz128(uint32_t row, uint32_t col) {
    return row * 128u + ((((col >> 1) ^ row) & 7u) << 4) + ((col & 1u) << 3);
}

}  // namespace tma
