// tcgen05 stride-1 convolution without an im2col gather: the "slab" kernel of the overlap classifier
// (Conv2D 3x3 and 4x1 of `res_block`, overlap_detector_temp.py:253-280) and of any other stride-1
// k > 1 conv whose Cin is 16 / 32 / 64 / 128.
//
// conv_tc_kernel builds every 128 x 32 im2col chunk with a gather: for a 3x3 conv each input element is
// loaded, batch-normalised, ELU'd (expm1f) and rounded nine times, and the index arithmetic of the gather
// (IMAD / ISETP / LOP3 = 40 % of its instructions, tensor pipe 2.5 % active) is what the kernel spends its
// time on.  Here one CTA owns up to four consecutive 128-pixel tiles of ONE image and loads the pixels
// they need ONCE:
//
//   * The image is addressed by a flat index over a ZERO-PADDED copy whose fast axis has Fp = F + kf - 1
//     entries: p = sp * Fp + fp.  With that numbering output pixel q (same flat index, computed on the
//     unpadded origin) reads padded input pixel q + ks_i * Fp + kf_j for tap (ks_i, kf_j): every filter tap is a
//     UNIFORM row shift.  Outputs with fp >= F are junk rows (kf - 1 of every Fp, 1-2 %) and are not stored.
//     3x3 convs take the image width as the fast axis (contiguous NHWC reads); the 4x1 convs take the
//     HEIGHT, so their four taps are shifts of 0..3 rows and the halo is 3 rows instead of 3 * W.
//   * BN + ELU/ReLU + TF32 rounding are applied once per element while the slab
//     [channel quad][row][16 B] (the UMMA canonical K-major no-swizzle layout: 8-row x 16-byte core
//     matrices, SBO = 128 B, LBO = slab stride) is written; padding rows are written as zeros AFTER the
//     activation, which is what Keras' 'same' padding of the activated tensor means.
//   * A tap's A operand is the slab's descriptor with its start address moved by `shift` rows (16 B each;
//     the no-swizzle address map is linear in the row, so any row is a legal start — the speaker net's
//     resunit_fused_kernel uses the same trick in 1-D).
//   * Weights: the host-arranged [32 x N] K-chunks of conv_tc.cu come through a TMA ring
//     (cp.async.bulk + mbarrier); every chunk feeds the MMAs of all the CTA's tiles (one TMEM accumulator
//     of N columns per tile), so weight traffic per output pixel drops by the tile count.
//   * Epilogue: tcgen05.ld, + bias (+ residual), 64-byte runs per thread straight to the NHWC output.
#include <stdlib.h>
#include <string.h>

#include "conv_common.cuh"

namespace {

constexpr int kSlabBK = 32;                 // K per ring chunk (matches mmla_tc_arrange_weights)
constexpr int kSlabMaxTiles = 4;             // (8 was tried: the 32-MMA elected region issues slower, every layer lost 5-10 %)
constexpr int kSlabMaxStages = 40;
constexpr int kSlabMaxChunks = 64;         // K <= 2048

struct SlabArgs {
    const float* x;
    const float* wg;          // arranged weights (conv_tc.cu layout, one N tile)
    const float* bias;
    const float* pre_scale;
    const float* pre_shift;
    const float* res;
    float* y;
    long long res_row_stride;
    long long img_pixels;     // H * W
    int pre_act;
    int F, S, Fp;             // fast / slow axis lengths, padded fast length
    unsigned fp_magic;        // floor(2^32 / Fp) + 1: p / Fp == umulhi(p, fp_magic) for the p < 2^20 used here
    int padF, padS;           // zero entries before the image on each axis
    int pixF, pixS;           // pixel-index strides of the two axes (NHWC: w -> 1, h -> W)
    int Cin, lq;              // lq = log2(Cin / 4)
    int K, nk;
    int kw;                   // filter width (tap index = ki * kw + kj)
    int shift_h, shift_w;     // row shift per filter row / column
    int tiles, T, cpi;        // 128-row tiles per image, tiles per CTA, CTAs per image
    int Rs;                   // slab stride in rows
    int stages;
    unsigned ring_off, bar_off;
    int nmma_last;            // MMAs (K = 8 each) in the last chunk; every other chunk has four
    unsigned aoff[kSlabMaxChunks * 4];   // per MMA: A-operand offset in 16-byte units = channel-quad slab + tap row shift
    long long* stamps;        // diagnostics: clock64 timeline of CTA `stamp_cta` (null = off)
    int stamp_cta;
};

__device__ __forceinline__ uint32_t sl_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void sl_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    asm volatile("trap;");
}
// BN + activation + TF32 rounding of one element with the activation as a COMPILE-TIME constant: with a run-time
// `act` the compiler wrapped every element's MUFU in its own (uniform) branch — 16 branches per four rows.
// ACT < 0: no BN / activation prologue at all (just the rounding).
template <int V>
struct SlInt {
    static constexpr int value = V;
};
template <int ACT>
__device__ __forceinline__ uint32_t sl_bn_act_tf32(uint32_t raw, float sc, float sh) {
    float v = __uint_as_float(raw);
    if (ACT >= 0) {
        v = fmaf(v, sc, sh);
        if (ACT == ACT_RELU) {
            v = fmaxf(v, 0.f);
        } else if (ACT == ACT_ELU) {
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v, 0.f) * 1.4426950408889634f));
            v = v > 0.f ? v : e - 1.f;
        }
    }
    return sl_tf32(v);
}
__device__ __forceinline__ uint64_t sl_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool sl_elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void sl_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// RES = residual added in the epilogue.  Layers without one run 512 threads at one or two CTAs per SM (the fill and
// the epilogue want warps; <= 64 registers keep two CTAs per SM) or 256 threads at three CTAs per SM (the full-resolution
// layers, which are HBM-heavy and gain most from co-resident CTAs in different phases); layers with a residual keep 256
// threads at two per SM because the residual prefetch holds 40 more registers.
template <int NT, bool RES, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 && !RES ? 3 : 2) conv_slab_kernel(const SlabArgs a) {
    constexpr int kSlabThreads = THREADS;
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(NT >> 3) << 17) |
                                (static_cast<uint32_t>(128 >> 4) << 24);   // D=f32, A=B=tf32, K-major, N, M=128
    constexpr uint32_t kChunkBytes = 8 * NT * 16;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    unsigned char* slab = base;
    unsigned char* ring = base + a.ring_off;
    uint64_t* full = reinterpret_cast<uint64_t*>(base + a.bar_off);      // [stages]
    uint64_t* empty = full + kSlabMaxStages;                             // [stages]
    uint64_t* accum = full + 2 * kSlabMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 2 * kSlabMaxStages + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool stamping = a.stamps != nullptr && static_cast<int>(blockIdx.x) == a.stamp_cta;
    auto stamp = [&](int slot) {
        if (stamping) a.stamps[slot] = clock64();
    };
    if (tid == 0) stamp(0);
    const int img = blockIdx.x / a.cpi;
    const int tile0 = (blockIdx.x - img * a.cpi) * a.T;
    const int Tc = min(a.T, a.tiles - tile0);
    const int q0 = tile0 * 128;
    const uint32_t cols = (Tc * NT <= 32) ? 32u : (Tc * NT <= 64) ? 64u : (Tc * NT <= 128) ? 128u : (Tc * NT <= 256) ? 256u : 512u;

    if (tid == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(accum, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) stamp(1);

    // ---- weight ring: the first `stages` chunks need no free slot, so they are requested before the fill ----
    if (lane == 0) {                          // one chunk per warp at a time: bulk copies of one thread serialise
        const int pre = a.nk < a.stages ? a.nk : a.stages;
        for (int kc = warp; kc < pre; kc += kSlabThreads / 32) {
            mbar_arrive_expect_tx(&full[kc], kChunkBytes);
            tma_bulk_g2s(ring + kc * kChunkBytes, a.wg + static_cast<long long>(kc) * (NT * kSlabBK), kChunkBytes, &full[kc]);
        }
    }

    // ---- slab fill: rows [0, rows) <-> padded flat indices q0 + row; BN + activation + TF32 once per element ----
    const int kh_taps = (a.K / a.Cin) / a.kw;
    const int rows = Tc * 128 + (kh_taps - 1) * a.shift_h + (a.kw - 1) * a.shift_w;
    {
        // A warp instruction covers 8 consecutive rows x 4 channel quads: in the slab that is four contiguous 128-byte
        // runs (one per quad), in global memory 64 contiguous bytes of 8 pixels.  (With lanes running over the channel
        // quads of one or two pixels, every 16-byte piece of a cp.async landed in a different slab row: 32 shared-memory
        // wavefronts per instruction, 34 cycles each measured.)  A warp keeps its quad group, so BN parameters load once.
        const int lqg = a.lq - 2;                             // log2(quad groups of 4)
        const int c4 = ((warp & ((1 << lqg) - 1)) << 2) + (lane >> 3);
        const int rpp = ((kSlabThreads / 32) >> lqg) * 8;     // rows per pass of the whole CTA
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool bn = a.pre_scale != nullptr;
        if (bn) {
            sc = __ldg(reinterpret_cast<const float4*>(a.pre_scale) + c4);
            sh = __ldg(reinterpret_cast<const float4*>(a.pre_shift) + c4);
        }
        const float* ximg = a.x + static_cast<long long>(img) * a.img_pixels * a.Cin + 4 * c4;
        unsigned char* dst0 = slab + static_cast<size_t>(c4) * a.Rs * 16;
        // Pass 1: every element goes global -> slab by a 16-byte cp.async (no register staging: all of the thread's
        // loads are in flight at once — the register-staged version paid one memory round trip per batch of four,
        // 12-20 k cycles per CTA); padding rows are written as zeros.  Pass 2, after cp.async.wait_all: the same
        // thread activates and rounds its own elements in place.
        // (both loops are unrolled by four independent rows: with 2-4 warps per scheduler the fill is bound by the
        //  dependent-issue latency of each warp's instruction chain, not by memory — clock64 stamps: 1 k cycles of
        //  cp.async wait against 10-30 k cycles of address arithmetic and activation at one row per iteration)
        unsigned long long okmask = 0ull;                     // bit i: row r0 + i * rpp holds image data
        const int r0 = (warp >> lqg) * 8 + (lane & 7);
        auto copy_batch = [&](int it) {                       // rows r0 + (it .. it+3) * rpp: one cp.async group
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + (it + u) * rpp;
                if (r < rows) {                               // (predication, no divergent paths: padding = zero-fill copy)
                    const int p = q0 + r;
                    const int sp = static_cast<int>(__umulhi(static_cast<unsigned>(p), a.fp_magic));
                    const int sI = sp - a.padS, f = p - sp * a.Fp - a.padF;
                    const bool ok = sI >= 0 && sI < a.S && f >= 0 && f < a.F;
                    okmask |= static_cast<unsigned long long>(ok) << (it + u);
                    const float* src = ximg + (ok ? static_cast<long long>(sI * a.pixS + f * a.pixF) * a.Cin : 0ll);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst0 + static_cast<size_t>(r) * 16)),
                                 "l"(src), "r"(ok ? 16 : 0)
                                 : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto xform_batch = [&](int it, auto act_c) {
            constexpr int ACT = decltype(act_c)::value;
            uint4 raw[4];
            const unsigned m4 = static_cast<unsigned>(okmask >> it) & 15u;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + (it + u) * rpp < rows) raw[u] = *reinterpret_cast<const uint4*>(dst0 + static_cast<size_t>(r0 + (it + u) * rpp) * 16);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (r0 + (it + u) * rpp < rows) {
                    const uint32_t keep = ((m4 >> u) & 1u) ? 0xFFFFFFFFu : 0u;         // padding rows stay zero
                    *reinterpret_cast<uint4*>(dst0 + static_cast<size_t>(r0 + (it + u) * rpp) * 16) =
                        make_uint4(sl_bn_act_tf32<ACT>(raw[u].x, sc.x, sh.x) & keep, sl_bn_act_tf32<ACT>(raw[u].y, sc.y, sh.y) & keep,
                                   sl_bn_act_tf32<ACT>(raw[u].z, sc.z, sh.z) & keep, sl_bn_act_tf32<ACT>(raw[u].w, sc.w, sh.w) & keep);
                }
            }
        };
        // Software pipeline over batches of four rows: batch i is activated while batches i+1 .. i+kAhead are still
        // landing, so a warp that the memory system holds back in its copies leaves the issue slots to warps that are
        // activating (copy-everything-then-activate-everything had every warp in the same phase at the same time).
        constexpr int kAhead = 3;
        auto fill = [&](auto act_c) {
            int it = 0;
            for (; r0 + it * rpp < rows; it += 4) {
                copy_batch(it);
                if (it >= 4 * kAhead) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(kAhead) : "memory");
                    xform_batch(it - 4 * kAhead, act_c);
                }
            }
            if (tid == 0) stamp(9);
            asm volatile("cp.async.wait_all;" ::: "memory");
            if (tid == 0) stamp(10);
            for (int jt = it >= 4 * kAhead ? it - 4 * kAhead : 0; jt < it; jt += 4) xform_batch(jt, act_c);
        };
        if (!bn) fill(SlInt<-1>{});
        else if (a.pre_act == ACT_ELU) fill(SlInt<ACT_ELU>{});
        else if (a.pre_act == ACT_RELU) fill(SlInt<ACT_RELU>{});
        else fill(SlInt<ACT_NONE>{});
    }
    if (tid == 0) stamp(2);
    fence_proxy_async_smem();                // generic-proxy slab writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) stamp(3);

    if (warp == 1) {
        // ================= weight producer: remaining chunks =================
        int stg = 0;
        uint32_t ph = 0u;                     // ring position kept in counters (a runtime % and / per chunk are ~50 instructions)
        for (int kc = a.stages; kc < a.nk; ++kc) {
            sl_wait(&empty[stg], ph);
            if (lane == 0) {
                mbar_arrive_expect_tx(&full[stg], kChunkBytes);
                tma_bulk_g2s(ring + stg * kChunkBytes, a.wg + static_cast<long long>(kc) * (NT * kSlabBK), kChunkBytes, &full[stg]);
            }
            __syncwarp();
            if (++stg == a.stages) { stg = 0; ph ^= 1u; }
        }
    } else if (warp == 0) {
        // ================= MMA issuer: warp-uniform loop, one elected lane issues =================
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t dA = sl_desc(smem_u32(slab), static_cast<uint32_t>(a.Rs) * 16u, 128u);
        const uint64_t dB = sl_desc(smem_u32(ring), NT * 16, 128);
        int stg = 0;
        uint32_t ph = 0u;
        for (int kc = 0; kc < a.nk; ++kc) {
            sl_wait(&full[stg], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (kc == 0 && lane == 0) stamp(4);
            const uint64_t bd0 = dB + static_cast<uint64_t>(stg * (kChunkBytes / 16));
            // per-MMA operand offsets come from the kernel parameters (host-computed table, constant bank -> uniform
            // registers): computing tap = k / Cin / kw here cost ~2 k cycles of dependent scalar code per chunk, which
            // — not the MMAs — set the pace of the issue loop
            const uint32_t aoff[4] = {a.aoff[kc * 4], a.aoff[kc * 4 + 1], a.aoff[kc * 4 + 2], a.aoff[kc * 4 + 3]};
            const int nmma = kc == a.nk - 1 ? a.nmma_last : 4;
            // descriptor arithmetic stays in warp-uniform code (uniform registers feed UTCHMMA directly; computed per
            // lane inside the elected region every MMA paid an R2UR waterfall) and the whole chunk — up to 16 MMAs —
            // goes out from ONE elected region: the tensor pipe queues only ~4 MMAs, so whatever the issuing warp
            // does between two MMAs is a bubble
            // (only the low word of a descriptor moves: start address in 16-byte units; LBO / SBO / version stay)
            const uint32_t alo = static_cast<uint32_t>(dA), ahi = static_cast<uint32_t>(dA >> 32);
            const uint32_t blo = static_cast<uint32_t>(bd0), bhi = static_cast<uint32_t>(bd0 >> 32);
            if (sl_elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                    for (int t = 0; t < kSlabMaxTiles; ++t) {
                        if (kk < nmma && t < Tc) {
                            const uint32_t acc = (kc | kk) != 0 ? 1u : 0u;
                            asm volatile(
                                "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\n"
                                "mov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
                                "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n}\n" ::"r"(tmem + static_cast<uint32_t>(t * NT)),
                                "r"(alo + aoff[kk] + static_cast<uint32_t>(t * 128)), "r"(ahi), "r"(blo + static_cast<uint32_t>(kk * 2 * NT)),
                                "r"(bhi), "r"(kIdesc), "r"(acc)
                                : "memory");
                        }
                    }
                }
                sl_commit(&empty[stg]);
                if (kc == a.nk - 1) sl_commit(accum);
            }
            if (++stg == a.stages) { stg = 0; ph ^= 1u; }
            __syncwarp();
        }
        if (lane == 0) stamp(5);
    }

    // ---- epilogue ----
    {
        // Every warp owns 32 accumulator lanes (warp % 4) and every kGroups-th (tile, 32-column chunk) unit (warp / 4).
        // A unit goes TMEM -> registers -> a warp-private [32 rows][32 + 4 floats] staging tile in the dead slab ->
        // global memory with eight lanes per row: each store instruction writes four whole 128-byte lines (the
        // direct path below writes 16 bytes into each of 32 lines), the residual is read the same way.
        constexpr int kChunks = NT / 32;
        constexpr int kGroups = kSlabThreads / 128;
        const int quarter = warp & 3, half = warp >> 2;
        float* stg = reinterpret_cast<float*>(slab) + warp * (32 * 36);
        const int seg = lane & 7, rsub = lane >> 3;
        const int out_pixels = a.S * a.Fp;
        // per unit and lane: the eight output rows (4i + rsub), their pixel offsets and residual values; the residual
        // of a unit is requested before its accumulator is touched (for the first unit: before the MMAs are even
        // waited for), so its memory round trip overlaps the tensor work instead of serialising row by row
        int pixoff[8];                          // pixel index inside the image, -1: junk row
        const long long imgbase = static_cast<long long>(img) * a.img_pixels;
        float4 rr[RES ? 8 : 1];
        auto prefetch = [&](int u) {
            const int t = u / kChunks, col0 = (u - t * kChunks) * 32;
            const int qb = q0 + t * 128 + quarter * 32 + rsub;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int q = qb + 4 * i;
                const int so = static_cast<int>(__umulhi(static_cast<unsigned>(q), a.fp_magic)), fo = q - so * a.Fp;
                const bool valid = q < out_pixels && fo < a.F;
                pixoff[i] = valid ? so * a.pixS + fo * a.pixF : -1;
                if (RES) rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (RES && valid) rr[i] = __ldg(reinterpret_cast<const float4*>(a.res + (imgbase + pixoff[i]) * a.res_row_stride + col0) + seg);
            }
        };
        const int nunits = Tc * kChunks;
        if (half < nunits) prefetch(half);
        // ONE warp polls the accumulator mbarrier; everybody else sleeps in a hardware barrier.  (With every warp
        // polling, try_wait + nanosleep + branch were a quarter of the kernel's executed instructions —
        // profiles/r01/conv_slab_v15_fullres_ncu_summary.txt — taken from the issue slots of co-resident CTAs.)
        if (warp == 0) sl_wait(accum, 0u);
        asm volatile("bar.sync 1, %0;" ::"n"(kSlabThreads) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 64) stamp(6);
        for (int u = half; u < nunits; u += kGroups) {
            const int t = u / kChunks, col0 = (u - t * kChunks) * 32;
            uint32_t r[32];
            const uint32_t taddr = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(t * NT + col0);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j)       // lane = row; row stride 144 B: a quarter-warp's STS.128 covers 8 distinct 16-byte slots
                *reinterpret_cast<uint4*>(stg + lane * 36 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            __syncwarp();
            const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + col0) + seg);
#pragma unroll
            for (int i = 0; i < 8; ++i) {     // rows 4i .. 4i+3 of the warp's 32, eight lanes (128 B) per row
                if (pixoff[i] >= 0) {
                    const float4 v = *reinterpret_cast<const float4*>(stg + (4 * i + rsub) * 36 + 4 * seg);
                    *(reinterpret_cast<float4*>(a.y + (imgbase + pixoff[i]) * NT + col0) + seg) =
                        RES ? make_float4(v.x + bv.x + rr[i].x, v.y + bv.y + rr[i].y, v.z + bv.z + rr[i].z, v.w + bv.w + rr[i].w)
                            : make_float4(v.x + bv.x, v.y + bv.y, v.z + bv.z, v.w + bv.w);
                }
            }
            __syncwarp();                     // the staging tile is rewritten by the next unit
            if (u + kGroups < nunits) prefetch(u + kGroups);
        }
    }
    if (tid == 64) stamp(7);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) stamp(8);
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
    }
}

int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

bool slab_enabled() {
    const char* e = getenv("MMLA_CONV_SLAB");     // read per launch: tests flip it between calls
    return !(e && e[0] == '0');
}

long long* g_slab_stamps = nullptr;       // mmla_debug_conv_slab_stamps: 64 rows (launch ordinal) x 16 slots
int g_slab_stamp_cta = 0, g_slab_stamp_row = 0;

template <int NT, bool RES, int THREADS>
int launch_slab(const SlabArgs& s, long long images, size_t smem, cudaStream_t st) {
    static size_t attr[64] = {};                                  // per device: function attributes are per device
    int dev = 0;
    MMLA_CUDA_CHECK(cudaGetDevice(&dev));
    MMLA_REQUIRE(dev >= 0 && dev < 64, MMLA_EUNSUP, "conv_slab: device ordinal %d out of range", dev);
    if (smem > attr[dev]) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(conv_slab_kernel<NT, RES, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem)));
        attr[dev] = smem;
    }
    conv_slab_kernel<NT, RES, THREADS><<<static_cast<unsigned>(images * s.cpi), THREADS, smem, st>>>(s);
    mmla_count_launch("conv_slab_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}

}  // namespace

// Whether conv_slab_kernel takes this layer (otherwise conv_tc_kernel's gather does).
bool mmla_conv_slab_eligible(const ConvArgs& a) {
    if (!slab_enabled()) return false;
    if (a.stride != 1 || a.x_is_u8 || a.kh * a.kw <= 1) return false;
    if (a.Cin != 16 && a.Cin != 32 && a.Cin != 64 && a.Cin != 128) return false;
    if (a.N != 32 && a.N != 64 && a.N != 128) return false;
    if (a.Ho != a.H || a.Wo != a.W || a.K != a.kh * a.kw * a.Cin) return false;
    if (a.H <= 1) return false;                    // the 1-D speaker net has its own fused kernels
    if (a.res && (a.res_row_stride % 4) != 0) return false;
    return true;
}

int mmla_launch_conv_slab(const ConvArgs& a, const float* wg, cudaStream_t st) {
    SlabArgs s;
    memset(&s, 0, sizeof(s));
    s.x = static_cast<const float*>(a.x); s.wg = wg; s.bias = a.bias;
    s.pre_scale = a.pre_scale; s.pre_shift = a.pre_shift; s.pre_act = a.pre_act;
    s.res = a.res; s.res_row_stride = a.res_row_stride; s.y = a.y;
    s.img_pixels = static_cast<long long>(a.H) * a.W;
    const long long images = a.M / s.img_pixels;
    const bool fast_h = a.kw == 1;                 // k x 1 filters: the taps run along the fast axis
    if (fast_h) {
        s.F = a.H; s.S = a.W; s.Fp = a.H + a.kh - 1; s.padF = a.pad_t; s.padS = 0;
        s.pixF = a.W; s.pixS = 1; s.shift_h = 1; s.shift_w = 0;
    } else {
        s.F = a.W; s.S = a.H; s.Fp = a.W + a.kw - 1; s.padF = a.pad_l; s.padS = a.pad_t;
        s.pixF = 1; s.pixS = a.W; s.shift_h = s.Fp; s.shift_w = 1;
    }
    s.fp_magic = static_cast<unsigned>((1ULL << 32) / static_cast<unsigned>(s.Fp)) + 1u;
    s.Cin = a.Cin; s.lq = ilog2(a.Cin / 4);
    s.K = a.K; s.nk = (a.K + kSlabBK - 1) / kSlabBK; s.kw = a.kw;
    MMLA_REQUIRE(s.nk <= kSlabMaxChunks, MMLA_EUNSUP, "conv_slab: K = %d is too large", a.K);
    const int halo = (a.kh - 1) * s.shift_h + (a.kw - 1) * s.shift_w;
    s.tiles = static_cast<int>((static_cast<long long>(s.S) * s.Fp + 127) / 128);
    const size_t chunk = static_cast<size_t>(8) * a.N * 16;      // one [32 x N] K-chunk of weights
    auto slab_rows = [&](int T) {
        int r = T * 128 + halo;
        if (a.Cin == 16) { while ((r & 7) != 2) ++r; } else if ((r & 1) == 0) ++r;     // conflict-free fill stores
        return r;
    };
    int threads = a.res ? 256 : 512;
    size_t kStagingBytes = static_cast<size_t>(threads / 32) * 32 * 36 * 4;   // epilogue staging tiles (one per warp) live in the dead slab
    constexpr size_t kBarBytes = 1024 + 128;                    // mbarriers + alignment slack
    auto slab_bytes = [&](int T) {
        const size_t b = static_cast<size_t>(a.Cin / 4) * slab_rows(T) * 16;
        return b < kStagingBytes ? kStagingBytes : b;
    };
    // Tiles per CTA, ring depth and CTAs per SM come from a small cost model fitted to the clock64 timelines of
    // scripts/prof_conv_slab.py: a weight-ring refill takes ~3000 cycles from the commit that frees the slot (with 2-3
    // slots the issue loop waited on weights for 2/3 of its time), a chunk holds T * 4 MMAs of max(N / 2, 45) cycles,
    // fill and epilogue pay a memory round trip each, and co-resident CTAs overlap each other's phases.
    auto bytes = [&](int T, int stages) { return slab_bytes(T) + stages * chunk + kBarBytes; };
    int tmax = kSlabMaxTiles;
    if (tmax > 512 / a.N) tmax = 512 / a.N;
    if (tmax > s.tiles) tmax = s.tiles;
    int force_t = 0, force_kb = 0, force_stages = 0;
    if (const char* e = getenv("MMLA_CONV_SLAB_TILES")) force_t = atoi(e);
    if (const char* e = getenv("MMLA_CONV_SLAB_KB")) force_kb = atoi(e);
    if (const char* e = getenv("MMLA_CONV_SLAB_STAGES")) force_stages = atoi(e);
    const int min2 = s.nk < 2 ? s.nk : 2;
    int T = 0;
    double best = 0.0;
    int best_bi = 1;
    for (int pass = 0; pass < 2 && !T; ++pass) {       // pass 1: a forced tile count / budget that fits nowhere is ignored
        if (pass == 1) force_t = force_kb = 0;
        // three (256 threads, no residual), two or one CTA per SM
        const int budgets_kb[3] = {75, 113, 226};
        double overlap[3] = {4.5, 3.2, 1.0};             // measured: co-resident CTAs in different phases are what pays (sweep_conv_slab_v9)
        if (const char* e = getenv("MMLA_CONV_SLAB_OVL3")) overlap[0] = atof(e);
        if (const char* e = getenv("MMLA_CONV_SLAB_OVL2")) overlap[1] = atof(e);
        for (int bi = a.res ? 1 : 0; bi < 3; ++bi) {
            const int nthr = a.res || bi == 0 ? 256 : 512;
            kStagingBytes = static_cast<size_t>(nthr / 32) * 32 * 36 * 4;
            const size_t budget = static_cast<size_t>(force_kb >= 16 && force_kb <= 226 ? force_kb : budgets_kb[bi]) * 1024;
            for (int t = 1; t <= tmax; ++t) {
                if (force_t >= 1 && force_t <= tmax && t != force_t) continue;
                if (bytes(t, min2) > budget) continue;
                int stages = static_cast<int>((budget - slab_bytes(t) - kBarBytes) / chunk);
                if (stages > s.nk) stages = s.nk;
                if (stages > kSlabMaxStages) stages = kSlabMaxStages;
                if (force_stages >= 1 && force_stages <= stages) stages = force_stages;
                const double per_chunk = t * 4.0 * (a.N / 2 > 45 ? a.N / 2 : 45);
                const double refill = (3000.0 + per_chunk) / stages;
                const double mma = s.nk * (per_chunk > refill ? per_chunk : refill);
                const double fill = 5000.0 + slab_rows(t) * (a.Cin / 4) / static_cast<double>(nthr) * 40.0;
                const double epi = 3000.0 + 700.0 * t * (a.N / 32);
                // useful tiles: the last CTA of an image may be partly empty
                const int cpi = (s.tiles + t - 1) / t;
                const double cost = (fill + mma + epi) * cpi / overlap[bi] / s.tiles;
                if (!T || cost < best) {
                    T = t; best = cost; s.stages = stages; best_bi = bi;
                }
            }
        }
    }
    MMLA_REQUIRE(T > 0, MMLA_EUNSUP, "conv_slab: layer does not fit in shared memory (Cin %d, halo %d rows)", a.Cin, halo);
    threads = a.res || best_bi == 0 ? 256 : 512;
    kStagingBytes = static_cast<size_t>(threads / 32) * 32 * 36 * 4;
    const size_t ring = s.stages * chunk;
    s.T = T;
    s.nmma_last = 0;
    for (int kc = 0; kc < s.nk; ++kc)
        for (int kk = 0; kk < 4; ++kk) {
            const int k = kc * kSlabBK + kk * 8;
            s.aoff[kc * 4 + kk] = 0;
            if (k < a.K) {
                const int tap = k / a.Cin, c0 = k % a.Cin;
                const int ki = tap / a.kw, kj = tap % a.kw;
                s.aoff[kc * 4 + kk] = static_cast<unsigned>((c0 >> 2) * slab_rows(T) + ki * s.shift_h + kj * s.shift_w);
                if (kc == s.nk - 1) s.nmma_last = kk + 1;
            }
        }
    s.cpi = (s.tiles + T - 1) / T;
    s.Rs = slab_rows(T);
    s.ring_off = static_cast<unsigned>((slab_bytes(T) + 127) / 128 * 128);
    s.bar_off = s.ring_off + static_cast<unsigned>(ring);
    const size_t smem = s.bar_off + kBarBytes;
    if (getenv("MMLA_CONV_SLAB_VERBOSE"))
        fprintf(stderr, "conv_slab: %dx%d Cin %d N %d k %dx%d: %d tiles/image, T %d, %d CTAs/image, ring %d of %d chunks, %zu KB smem, %d threads\n",
                a.H, a.W, a.Cin, a.N, a.kh, a.kw, s.tiles, s.T, s.cpi, s.stages, s.nk, smem / 1024, threads);
    if (g_slab_stamps && g_slab_stamp_row < 64) {
        s.stamps = g_slab_stamps + 16 * g_slab_stamp_row++;
        s.stamp_cta = static_cast<int>((static_cast<long long>(g_slab_stamp_cta) % images) * s.cpi + s.cpi / 2);   // a mid-image CTA
    }
    MMLA_REQUIRE(static_cast<long long>(s.tiles) * 128 + halo < (1LL << 20), MMLA_EUNSUP, "conv_slab: image too large");
    MMLA_REQUIRE(images * s.cpi < (1LL << 31) && images * s.img_pixels * (a.Cin > a.N ? a.Cin : a.N) < (1LL << 40), MMLA_EUNSUP,
                 "conv_slab: batch too large");
    if (a.res) {
        switch (a.N) {
            case 32: return launch_slab<32, true, 256>(s, images, smem, st);
            case 64: return launch_slab<64, true, 256>(s, images, smem, st);
            default: return launch_slab<128, true, 256>(s, images, smem, st);
        }
    }
    if (threads == 256) {
        switch (a.N) {
            case 32: return launch_slab<32, false, 256>(s, images, smem, st);
            case 64: return launch_slab<64, false, 256>(s, images, smem, st);
            default: return launch_slab<128, false, 256>(s, images, smem, st);
        }
    }
    switch (a.N) {
        case 32: return launch_slab<32, false, 512>(s, images, smem, st);
        case 64: return launch_slab<64, false, 512>(s, images, smem, st);
        default: return launch_slab<128, false, 512>(s, images, smem, st);
    }
}

extern "C" __attribute__((visibility("default"))) void mmla_debug_conv_slab_stamps(long long* dev_stamps, int32_t image) {
    g_slab_stamps = dev_stamps;
    g_slab_stamp_cta = image;
    g_slab_stamp_row = 0;
}
