// Shared by resblock2d_fused.cu (one CTA per work item) and resblock2d_persist.cu (persistent, warp-specialised):
// kernel arguments and the PTX wrappers of the overlap net's fused res_block conv pair.
#pragma once
#include <stdint.h>

#include "conv_common.cuh"

namespace {

constexpr int kRbBK = 32;                   // K per ring chunk (mmla_tc_arrange_weights)
constexpr int kRbBKh = 64;                  // ... of the fp16-operand mode (mmla_rb_arrange_weights_f16): same bytes per chunk
constexpr int kRbMaxTiles = 4;
constexpr int kRbMaxStages = 40;
constexpr int kRbMaxChunks = 64;           // conv1 + conv2 chunks

struct RbArgs {
    const float* x;
    const float* w1;          // arranged weights (conv_tc.cu layout, one N tile) of the 3x3
    const float* w2;          // ... of the 4x1
    const float* b1;
    const float* b2;
    const float* bn1_scale;
    const float* bn1_shift;
    const float* bn2_scale;
    const float* bn2_shift;
    const float* res;
    float* y;
    const void* xa;           // F16: ELU(BN1(x)) already as fp16 [B,H,W,Cin] (written by the producer of x), or null: the fill converts x
    void* ya;                 // F16: ELU(BN(y)) as fp16 [B,H,W,C] for the NEXT block (its BN1 = ya_scale / ya_shift), or null
    const float* ya_scale;
    const float* ya_shift;
    int y_f16;                // F16 + HPOOL: the row-pooled conv output y = [B, H/2, W, C] is written as fp16 (pool_shortcut_kernel reads it)
    const void* img;          // STEM: the classifier input [B,H,W,3] (uint8 or float32); x is unused
    const float* stem_w;      // STEM: Conv2D(16, 1x1) weights [3][16] and bias [16] (overlap_detector_temp.py:283)
    const float* stem_b;
    int img_is_u8;
    int n_ctas;               // PAIR: real CTAs (the grid is rounded up to whole pairs)
    int u_bufs;               // persistent kernel: u slabs (2 in the fp16-operand form: epilogue 1 (k+1) writes while conv2 (k) reads)
    int mma_hi;               // persistent kernel: the MMA issuer is the CTA's highest-numbered warp (MMLA_PS_MMA_HI)
    int hpool;                // HPOOL: Fp = H + 4, S = 128 T - 4, y = [B, H/2, W, C] = max over row pairs (2i, 2i+1) of the block output
    long long res_row_stride;
    long long img_pixels;     // H * W
    int H, W, Fp;
    unsigned fp_magic;        // floor(2^32 / Fp) + 1
    int total_q;              // W * Fp
    int Cin, lq;              // lq = log2(Cin / 4)
    int nk1, nk;              // ring chunks of conv1 / of both convolutions
    int nmma1_last, nmma2_last;
    int T, S, cpi;            // tiles per CTA, outputs per CTA (128 T - 3), CTAs per image
    int RsX, RsU;             // slab strides in rows
    int stages;
    unsigned ring_off, par_off, bar_off;
    unsigned x1_off, u_off, u_bytes, stg_off;   // persistent kernel: second x slab, u slab (+ size), epilogue-2 staging tiles
    unsigned aoff[kRbMaxChunks * 4];   // per MMA: channel-quad slab + tap row shift, in 16-byte units
    long long* stamps;        // diagnostics: clock64 timeline of CTA `stamp_cta` (null = off)
    int stamp_cta;
};

__device__ __forceinline__ uint32_t rb_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
// try_wait with a suspend-time hint: the warp is parked in hardware instead of spinning through the issue slots of the
// working warps (the persistent kernel has up to twenty warps waiting at a time); traps instead of hanging on a logic error.
__device__ __forceinline__ void rb_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 22); ++i) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity), "r"(100000u)
                     : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
}
// BN + ELU + TF32 rounding of one element: the expression of conv_slab.cu's fill (bit-identical results).
__device__ __forceinline__ uint32_t rb_bn_elu_tf32(float v, float sc, float sh) {
    v = fmaf(v, sc, sh);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v, 0.f) * 1.4426950408889634f));
    v = v > 0.f ? v : e - 1.f;
    return rb_tf32(v);
}
// fp16-operand mode: BN + ELU in fp32 (same expression), then two elements per F2FP into one half2 word.  Operands keep the
// 11 significant bits TF32 keeps; what changes is the exponent range, so the conversion saturates (ELU output is >= -1: only
// the upper bound can be hit) instead of producing infinities, and values below 6e-5 go subnormal (absolute error <= 3e-8).
__device__ __forceinline__ float rb_bn_elu(float v, float sc, float sh) {
    v = fmaf(v, sc, sh);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v, 0.f) * 1.4426950408889634f));
    return v > 0.f ? v : e - 1.f;
}
__device__ __forceinline__ uint32_t rb_pack_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // F2FP.SATFINITE: +-65504 instead of inf
    return r;
}
// uint8 -> float without the quarter-rate I2F pipe: 0x4B000000 | v is the float 2^23 + v; the subtraction is exact.
__device__ __forceinline__ float rb_u8_to_float(unsigned char v) { return __uint_as_float(0x4B000000u | v) - 8388608.0f; }
__device__ __forceinline__ uint64_t rb_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool rb_elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void rb_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// PAIR (cta_group::2): the leader's commit arrives on the barrier at the same offset in BOTH CTAs of the pair.
__device__ __forceinline__ void rb_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
}
__device__ __forceinline__ void rb_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t rb_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void rb_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace
