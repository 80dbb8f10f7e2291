// Shared helpers for libmmla_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mmla_b200.h"

void mmla_set_error(const char* fmt, ...);

#define MMLA_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            mmla_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),         \
                           __FILE__, __LINE__);                                            \
            return MMLA_ECUDA;                                                             \
        }                                                                                  \
    } while (0)

#define MMLA_REQUIRE(cond, code, ...)                                                      \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            mmla_set_error(__VA_ARGS__);                                                   \
            return (code);                                                                 \
        }                                                                                  \
    } while (0)

// Called right after every kernel launch: bumps the counter mmla_launch_count() reports and, while a launch
// trace is open (mmla_trace_begin), records a CUDA event on `st` so the kernel's device time can be read back.
void mmla_count_launch(const char* kernel_name, cudaStream_t st);
int mmla_num_sms();   // SM count of the current device (cached per device), <0 on error

// `static MmlaPerDeviceOnce once; if (once.first()) cudaFuncSetAttribute(...)`: function attributes such as
// MaxDynamicSharedMemorySize are per DEVICE, so a per-process flag would leave a second GPU used from the same
// process with the 48 KB default.
struct MmlaPerDeviceOnce {
    bool done[64] = {};
    bool first() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA bulk copy (cp.async.bulk → SASS UBLKCP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global → shared, completion signalled on an mbarrier (TMA engine).
// dst and src must be 16-byte aligned and bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// Orders prior generic-proxy smem accesses before later async-proxy (TMA) accesses.
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// int16 (raw 16 bits, zero-extended) → float without the slow I2F pipe:
// 0x4B000000 | (x ^ 0x8000) is the float 2^23 + (x + 32768); subtract the bias exactly.
__device__ __forceinline__ float s16_bits_to_float(uint32_t u16) {
    return __uint_as_float(0x4B000000u | (u16 ^ 0x8000u)) - 8421376.0f;
}
