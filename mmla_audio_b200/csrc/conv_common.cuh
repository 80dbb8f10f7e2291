// Shared between the fp32 CUDA-core conv kernel (nets.cu) and the tcgen05 TF32 kernel (conv_tc.cu).
#pragma once
#include "common.cuh"

constexpr float kBnEps = 1e-3f;
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_ELU = 2 };

// ---------------------------------------------------------------------------------------------
// implicit-GEMM convolution: Y[M,N] = gather(X)[M,K] * W[K,N] + bias (+ residual)
// ---------------------------------------------------------------------------------------------
struct ConvArgs {
    const void* x;            // NHWC float32, or uint8 when x_is_u8
    const float* w;           // [K][N]
    const float* bias;        // [N]
    const float* pre_scale;   // per input channel BN scale (null: no BN/activation prologue)
    const float* pre_shift;
    const float* res;         // residual rows (null: none)
    float* y;                 // [M][N]
    long long res_row_stride; // floats between residual rows
    long long M;
    int x_is_u8, pre_act;
    int H, W, Cin, Ho, Wo, N, K, kh, kw, stride, pad_t, pad_l;
};

constexpr int kBM = 128, kBK = 16;

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_ELU) return v > 0.f ? v : expm1f(v);
    return v;
}
// TF32 operand paths (conv_tc.cu, conv_slab.cu): the result is rounded to 10 mantissa bits right after, so ELU's
// negative branch is ex2.approx(v * log2 e) - 1 (absolute error ~1e-7) instead of expm1f — expm1f was 60-70 % of
// conv_slab_kernel's executed instructions (profiles/r01/conv_slab_v1_ncu_summary.txt).  Branch-free on purpose: written
// as `v > 0 ? v : __expf(v) - 1` the compiler emitted a divergent branch per element around the MUFU.
__device__ __forceinline__ float apply_act_tc(float v, int act) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v, 0.f) * 1.4426950408889634f));
    const float neg = act == ACT_ELU ? e - 1.f : (act == ACT_RELU ? 0.f : v);      // the value for v <= 0
    return v > 0.f ? v : neg;
}

