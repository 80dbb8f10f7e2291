// Persistent, warp-specialised version of resblock2d_fused_kernel for the C = 32 blocks of the overlap classifier
// (overlap_detector_temp.py:253-280).  Default: block 1 (128 x 151 x 16 -> 32, with the stem); MMLA_NET_PERSIST=2 also takes
// blocks 2-3 (64 x 76 x 32 -> 32), where it measured no faster than resblock2d_fused_kernel (see mmla_try_launch_...).
//
// In resblock2d_fused_kernel a CTA runs fill -> conv1 -> epilogue 1 -> conv2 -> epilogue 2 back to back and the tensor pipe
// only stays busy as far as two or three co-resident CTAs happen to be in different phases: 53 % of the issue slots of the
// N = 32 MMAs (45 cycles each, the operand-fetch floor) on block 1.  Here ONE CTA per SM walks over work items
// (image, chunk of 128 T - 3 outputs) and every phase has its own warps, so the MMA warp always has an item to work on:
//
//   warp 0        MMA issuer, software-pipelined by one item: conv1(k) | conv2(k-1) | conv1(k+1) | conv2(k) ...
//   warp 1        loads both convolutions' weights ONCE (36 / 52 KB stay resident: no weight ring)
//   warps 2-13    fill: x slab of item k+1 / k+2 (cp.async + in-place BN1 / ELU / TF32, or the stem from the image bytes)
//   warps 14-17   epilogue 1: accumulator 1 -> + b1 -> BN2 -> ELU -> TF32 -> u slab
//   warps 18-21   epilogue 2: accumulator 2 -> staging -> + b2 (+ residual | row max) -> NHWC
//
// Double buffered: the x slab and accumulator 1 (conv1 of item k+1 runs while epilogue 1 reads item k); single: the u slab and
// accumulator 2 (conv2(k-1) is issued right behind conv1(k), so it has finished long before epilogue 1(k) wants the slab back).
// Ten mbarriers carry the hand-offs; a producer is never more than one phase ahead of its consumer (comments at the waits).
// The arithmetic per element is resblock2d_fused_kernel's, so the results are bit-identical (tests/test_resblock2d_gpu.py).
#include <stdlib.h>
#include <string.h>

#include "resblock2d_common.cuh"

namespace {

constexpr int kPsFillWarps = 12;
constexpr int kPsE1Warp0 = 2 + kPsFillWarps;        // four epilogue-1 warps: consecutive, so warp & 3 covers the four TMEM lane quarters
constexpr int kPsE2Warp0 = kPsE1Warp0 + 4;
constexpr int kPsThreads = (kPsE2Warp0 + 4) * 32;
constexpr int NT = 32;

struct PsBars {
    uint64_t wfull, xfull[2], xempty[2], a1full[2], a1empty[2], ufull[2], uempty[2], a2full, a2empty;
    uint32_t tmem;
};

struct PsItem {
    int img, Qc, nq, Tc;
};
__device__ __forceinline__ PsItem ps_item(const RbArgs& a, int k) {
    const int it = blockIdx.x + k * gridDim.x;
    PsItem r;
    r.img = it / a.cpi;
    r.Qc = (it - r.img * a.cpi) * a.S;
    r.nq = min(a.S, a.total_q - r.Qc);
    r.Tc = (r.nq + 3 + 127) >> 7;
    return r;
}
__device__ __forceinline__ void ps_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// F16: the fp16-operand form (resblock2d_fused.cu; here for the stem block only): slabs of 8 channels per 16-byte row, K = 16
// per MMA, fp16 weight chunks.
template <bool RES, bool STEM, bool HPOOL, bool F16 = false>
__global__ void __launch_bounds__(kPsThreads, 1) resblock2d_persist_kernel(const RbArgs a) {
    // (fp16 operands: the stem block, or a block whose producer has written the activated fp16 operand `xa`)
    constexpr uint32_t kIdesc = (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | (static_cast<uint32_t>(NT >> 3) << 17) |
                                (static_cast<uint32_t>(128 >> 4) << 24);   // D=f32, A=B=tf32 (or f16), K-major, N, M=128
    constexpr uint32_t kChunkBytes = 8 * NT * 16;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    unsigned char* xs0 = base;                                             // x slabs: base, base + x1_off
    unsigned char* us = base + a.u_off;
    unsigned char* wts = base + a.ring_off;                                // conv1's chunks, then conv2's
    float* stg_all = reinterpret_cast<float*>(base + a.stg_off);           // epilogue-2 staging tiles, one per warp
    float* par = reinterpret_cast<float*>(base + a.par_off);               // b1 | bn2 scale | bn2 shift
    PsBars& bar = *reinterpret_cast<PsBars*>(base + a.bar_off);

    // `warp` is the ROLE index below.  mma_hi: hardware warps 0-19 take roles 2-21 (fill, epilogues) and the two highest warps
    // the weight loader and the MMA issuer, so that the issuer is the warp its scheduler's arbiter prefers (highest id first).
    const int hw = static_cast<int>(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int nw = kPsThreads / 32;
    const int warp = a.mma_hi ? (hw < nw - 2 ? hw + 2 : nw - 1 - hw) : hw;
    const int tid = warp * 32 + lane;                                      // role-relative thread index
    const int n_items = a.n_ctas;                                          // work items of the launch
    const int my = static_cast<int>(blockIdx.x) < n_items ? (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
    const int T = a.T;
    const uint32_t cols = 3 * T * NT <= 256 ? 256u : 512u;
    // diagnostics: clock64 stamps of items 4..7 of CTA `stamp_cta` (16 slots per item): 0/1 fill, 2/3 conv1 issue, 4/5 conv2 issue,
    // 6/7/8 epilogue 1 (accumulator ready, slab free, done), 9/10/11 epilogue 2 (accumulator ready, drained, stored)
    auto stamp = [&](int k, int slot) {
        if (a.stamps && static_cast<int>(blockIdx.x) == a.stamp_cta && k >= 4 && k < 8) a.stamps[(k - 4) * 16 + slot] = clock64();
    };

    if (tid == 0) {
        mbar_init(&bar.wfull, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar.xfull[i], kPsFillWarps);
            mbar_init(&bar.xempty[i], 1);
            mbar_init(&bar.a1full[i], 1);
            mbar_init(&bar.a1empty[i], 4);
            mbar_init(&bar.ufull[i], 4);
            mbar_init(&bar.uempty[i], 1);
        }
        mbar_init(&bar.a2full, 1);
        mbar_init(&bar.a2empty, 4);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bar.tmem)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 3 * NT; i += kPsThreads) par[i] = __ldg((i < NT ? a.b1 : i < 2 * NT ? a.bn2_scale : a.bn2_shift) + (i % NT));
    // the u slab's rows past the last tile only feed outputs that are never stored, but they must be finite from the start
    for (int i = tid; i < static_cast<int>(a.u_bytes / 16) * a.u_bufs; i += kPsThreads) reinterpret_cast<uint4*>(us)[i] = make_uint4(0u, 0u, 0u, 0u);
    // u slab of item j: buffer j & 1 when there are two (its barriers then run one phase per two items), else the only one
    const bool u2 = a.u_bufs == 2;
    fence_proxy_async_smem();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = bar.tmem;

    if (warp == 0) {
        // ================= MMA issuer =================
        const uint64_t dW = rb_desc(smem_u32(wts), NT * 16, 128);
        auto conv = [&](unsigned char* slab, uint32_t rs16, int kc0, int kc1, int nmma_last, uint32_t dcol, int Tc) {
            const uint64_t dA = rb_desc(smem_u32(slab), rs16, 128u);
            const uint32_t alo = static_cast<uint32_t>(dA), ahi = static_cast<uint32_t>(dA >> 32);
            for (int kc = kc0; kc < kc1; ++kc) {
                const uint64_t bd0 = dW + static_cast<uint64_t>(kc * (kChunkBytes / 16));
                const uint32_t blo = static_cast<uint32_t>(bd0), bhi = static_cast<uint32_t>(bd0 >> 32);
                const uint32_t aoff[4] = {a.aoff[kc * 4], a.aoff[kc * 4 + 1], a.aoff[kc * 4 + 2], a.aoff[kc * 4 + 3]};
                const int nmma = kc == kc1 - 1 ? nmma_last : 4;
                if (rb_elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                        for (int t = 0; t < kRbMaxTiles; ++t) {
                            if (kk < nmma && t < Tc) {
                                const uint32_t acc = (kc != kc0 || kk != 0) ? 1u : 0u;
                                if constexpr (F16)
                                    asm volatile(
                                        "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\n"
                                        "mov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
                                        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tmem + dcol + static_cast<uint32_t>(t * NT)),
                                        "r"(alo + aoff[kk] + static_cast<uint32_t>(t * 128)), "r"(ahi), "r"(blo + static_cast<uint32_t>(kk * 2 * NT)),
                                        "r"(bhi), "r"(kIdesc), "r"(acc)
                                        : "memory");
                                else
                                asm volatile(
                                    "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\n"
                                    "mov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
                                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n}\n" ::"r"(tmem + dcol + static_cast<uint32_t>(t * NT)),
                                    "r"(alo + aoff[kk] + static_cast<uint32_t>(t * 128)), "r"(ahi), "r"(blo + static_cast<uint32_t>(kk * 2 * NT)),
                                    "r"(bhi), "r"(kIdesc), "r"(acc)
                                    : "memory");
                            }
                        }
                    }
                }
                __syncwarp();
            }
        };
        auto conv2 = [&](int j) {
            const PsItem it = ps_item(a, j);
            const int ub = u2 ? (j & 1) : 0;
            rb_wait(&bar.ufull[ub], static_cast<uint32_t>((u2 ? j >> 1 : j) & 1));   // epilogue 1 (j) wrote the u slab
            if (j >= 1) rb_wait(&bar.a2empty, static_cast<uint32_t>((j - 1) & 1));   // epilogue 2 (j-1) drained accumulator 2
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) stamp(j, 4);
            conv(us + ub * a.u_bytes, static_cast<uint32_t>(a.RsU) * 16u, a.nk1, a.nk, a.nmma2_last, static_cast<uint32_t>(2 * T * NT), it.Tc);
            if (lane == 0) stamp(j, 5);
            if (rb_elect_one()) {
                rb_commit(&bar.uempty[ub]);
                rb_commit(&bar.a2full);
            }
            __syncwarp();
        };
        rb_wait(&bar.wfull, 0u);
        for (int k = 0; k < my; ++k) {
            const PsItem it = ps_item(a, k);
            const int b = k & 1;
            rb_wait(&bar.xfull[b], static_cast<uint32_t>((k >> 1) & 1));             // fill (k) done
            if (k >= 2) rb_wait(&bar.a1empty[b], static_cast<uint32_t>(((k >> 1) - 1) & 1));   // epilogue 1 (k-2) drained accumulator 1[b]
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) stamp(k, 2);
            conv(xs0 + b * a.x1_off, static_cast<uint32_t>(a.RsX) * 16u, 0, a.nk1, a.nmma1_last, static_cast<uint32_t>(b * T * NT), it.Tc);
            if (lane == 0) stamp(k, 3);
            if (rb_elect_one()) {
                rb_commit(&bar.xempty[b]);
                rb_commit(&bar.a1full[b]);
            }
            __syncwarp();
            if (k >= 1) conv2(k - 1);
        }
        if (my >= 1) conv2(my - 1);
    } else if (warp == 1) {
        // ================= weights: once =================
        if (lane == 0) {
            const uint32_t b1 = static_cast<uint32_t>(a.nk1) * kChunkBytes, b2 = static_cast<uint32_t>(a.nk - a.nk1) * kChunkBytes;
            mbar_arrive_expect_tx(&bar.wfull, b1 + b2);
            tma_bulk_g2s(wts, a.w1, b1, &bar.wfull);
            tma_bulk_g2s(wts + b1, a.w2, b2, &bar.wfull);
        }
        __syncwarp();
    } else if (warp < 2 + kPsFillWarps) {
        // ================= fill: x slab of item k into buffer k & 1 =================
        const int fw = warp - 2;
        for (int k = 0; k < my; ++k) {
            const PsItem it = ps_item(a, k);
            const int b = k & 1;
            if (k >= 2) rb_wait(&bar.xempty[b], static_cast<uint32_t>(((k >> 1) - 1) & 1));   // conv1 (k-2) has read the buffer
            if (tid == 64) stamp(k, 0);
            unsigned char* slab = xs0 + b * a.x1_off;
            const int rows = it.Tc * 128 + 2 * a.Fp + 2;
            const int Qc = it.Qc;
            if constexpr (STEM) {
                const int c4 = lane >> 3;                             // Cin = 16: four quads, eight rows per warp instruction
                const float4 sc = __ldg(reinterpret_cast<const float4*>(a.bn1_scale) + c4);
                const float4 sh = __ldg(reinterpret_cast<const float4*>(a.bn1_shift) + c4);
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.stem_w) + c4);
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.stem_w + 16) + c4);
                const float4 w2 = __ldg(reinterpret_cast<const float4*>(a.stem_w + 32) + c4);
                const float4 sb = __ldg(reinterpret_cast<const float4*>(a.stem_b) + c4);
                const unsigned char* img8 = static_cast<const unsigned char*>(a.img) + static_cast<long long>(it.img) * a.img_pixels * 3;
                const float* imgf = static_cast<const float*>(a.img) + static_cast<long long>(it.img) * a.img_pixels * 3;
                unsigned char* dst0 = slab + static_cast<size_t>(c4) * a.RsX * 16;
                for (int rb = fw * 8; rb < rows; rb += 4 * kPsFillWarps * 8) {      // warp-uniform trip count (the F16 form shuffles)
                    const int r0 = rb + (lane & 7);
                    float c[4][3];
                    bool ok[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = r0 + u * kPsFillWarps * 8;
                        const int p = Qc - 1 + r;
                        const int wp = static_cast<int>(__umulhi(static_cast<unsigned>(p < 0 ? 0 : p), a.fp_magic));
                        const int w = wp - 1, h = p - wp * a.Fp - 1;
                        ok[u] = r < rows && p >= 0 && w >= 0 && w < a.W && h >= 0 && h < a.H;
                        const long long px = ok[u] ? static_cast<long long>(h * a.W + w) * 3 : 0ll;
                        if (a.img_is_u8) {
                            c[u][0] = rb_u8_to_float(img8[px]); c[u][1] = rb_u8_to_float(img8[px + 1]); c[u][2] = rb_u8_to_float(img8[px + 2]);
                        } else {
                            c[u][0] = imgf[px]; c[u][1] = imgf[px + 1]; c[u][2] = imgf[px + 2];
                        }
                    }
                    if constexpr (F16) {
                        // resblock2d_fused.cu's fp16 stem fill: half2 pairs per lane, the two quads of an octet joined by one
                        // shuffle pair per row (lanes l, l ^ 8); the even quad's lane stores rows u = 0, 2, the odd one's u = 1, 3
                        uint32_t pk[4][2];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t keep = ok[u] ? 0xFFFFFFFFu : 0u;
                            const float c0 = c[u][0], c1 = c[u][1], c2 = c[u][2];
                            pk[u][0] = rb_pack_h2(rb_bn_elu(fmaf(c2, w2.x, fmaf(c1, w1.x, fmaf(c0, w0.x, sb.x))), sc.x, sh.x),
                                                  rb_bn_elu(fmaf(c2, w2.y, fmaf(c1, w1.y, fmaf(c0, w0.y, sb.y))), sc.y, sh.y)) & keep;
                            pk[u][1] = rb_pack_h2(rb_bn_elu(fmaf(c2, w2.z, fmaf(c1, w1.z, fmaf(c0, w0.z, sb.z))), sc.z, sh.z),
                                                  rb_bn_elu(fmaf(c2, w2.w, fmaf(c1, w1.w, fmaf(c0, w0.w, sb.w))), sc.w, sh.w)) & keep;
                        }
                        const bool odd = (c4 & 1) != 0;
                        unsigned char* dsth = slab + static_cast<size_t>(c4 >> 1) * a.RsX * 16;
#pragma unroll
                        for (int u = 0; u < 4; u += 2) {
                            const uint32_t g0 = __shfl_xor_sync(0xffffffffu, odd ? pk[u][0] : pk[u + 1][0], 8);
                            const uint32_t g1 = __shfl_xor_sync(0xffffffffu, odd ? pk[u][1] : pk[u + 1][1], 8);
                            const int r = r0 + (u + (odd ? 1 : 0)) * kPsFillWarps * 8;
                            if (r < rows)
                                *reinterpret_cast<uint4*>(dsth + static_cast<size_t>(r) * 16) =
                                    odd ? make_uint4(g0, g1, pk[u + 1][0], pk[u + 1][1]) : make_uint4(pk[u][0], pk[u][1], g0, g1);
                        }
                        continue;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = r0 + u * kPsFillWarps * 8;
                        if (r < rows) {
                            const uint32_t keep = ok[u] ? 0xFFFFFFFFu : 0u;
                            const float c0 = c[u][0], c1 = c[u][1], c2 = c[u][2];
                            *reinterpret_cast<uint4*>(dst0 + static_cast<size_t>(r) * 16) = make_uint4(
                                rb_bn_elu_tf32(fmaf(c2, w2.x, fmaf(c1, w1.x, fmaf(c0, w0.x, sb.x))), sc.x, sh.x) & keep,
                                rb_bn_elu_tf32(fmaf(c2, w2.y, fmaf(c1, w1.y, fmaf(c0, w0.y, sb.y))), sc.y, sh.y) & keep,
                                rb_bn_elu_tf32(fmaf(c2, w2.z, fmaf(c1, w1.z, fmaf(c0, w0.z, sb.z))), sc.z, sh.z) & keep,
                                rb_bn_elu_tf32(fmaf(c2, w2.w, fmaf(c1, w1.w, fmaf(c0, w0.w, sb.w))), sc.w, sh.w) & keep);
                        }
                    }
                }
            } else if constexpr (F16) {
                // the producer of x has written ELU(BN1(x)) as fp16: a plain asynchronous gather, one channel octet of one row per
                // copy (zero-fill form for the padding rows), every copy of the item in flight at once (resblock2d_fused.cu)
                const int no = a.Cin >> 3;                            // channel octets per row: 2, 4
                const int osh = no >= 4 ? 2 : 1;
                const int rpi = 32 >> osh, ngrp = no >> osh;
                const int c8 = ((fw % ngrp) << osh) + (lane >> (5 - osh));
                const int rpp = (kPsFillWarps / ngrp) * rpi;
                const uint16_t* ximg = static_cast<const uint16_t*>(a.xa) + static_cast<long long>(it.img) * a.img_pixels * a.Cin + 8 * c8;
                unsigned char* dsth = slab + static_cast<size_t>(c8) * a.RsX * 16;
                for (int r = (fw / ngrp) * rpi + (lane & (rpi - 1)); r < rows; r += rpp) {
                    const int p = Qc - 1 + r;
                    const int wp = static_cast<int>(__umulhi(static_cast<unsigned>(p < 0 ? 0 : p), a.fp_magic));
                    const int w = wp - 1, h = p - wp * a.Fp - 1;
                    const bool ok = p >= 0 && w >= 0 && w < a.W && h >= 0 && h < a.H;
                    const uint16_t* src = ximg + (ok ? static_cast<long long>(h * a.W + w) * a.Cin : 0ll);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dsth + static_cast<size_t>(r) * 16)), "l"(src),
                                 "r"(ok ? 16 : 0)
                                 : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_all;" ::: "memory");
            } else {
                const int lqg = a.lq - 2;                             // log2(quad groups of 4)
                const int c4 = ((fw & ((1 << lqg) - 1)) << 2) + (lane >> 3);
                const int rpp = (kPsFillWarps >> lqg) * 8;            // rows per pass of the eight fill warps
                const float4 sc = __ldg(reinterpret_cast<const float4*>(a.bn1_scale) + c4);
                const float4 sh = __ldg(reinterpret_cast<const float4*>(a.bn1_shift) + c4);
                const float* ximg = a.x + static_cast<long long>(it.img) * a.img_pixels * a.Cin + 4 * c4;
                unsigned char* dst0 = slab + static_cast<size_t>(c4) * a.RsX * 16;
                unsigned long long okmask = 0ull;                     // bit i: row r0 + i * rpp holds image data
                const int r0 = (fw >> lqg) * 8 + (lane & 7);
                auto copy_batch = [&](int i0) {                       // rows r0 + (i0 .. i0+3) * rpp: one cp.async group
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = r0 + (i0 + u) * rpp;
                        if (r < rows) {
                            const int p = Qc - 1 + r;
                            const int wp = static_cast<int>(__umulhi(static_cast<unsigned>(p < 0 ? 0 : p), a.fp_magic));
                            const int w = wp - 1, h = p - wp * a.Fp - 1;
                            const bool ok = p >= 0 && w >= 0 && w < a.W && h >= 0 && h < a.H;
                            okmask |= static_cast<unsigned long long>(ok) << (i0 + u);
                            const float* src = ximg + (ok ? static_cast<long long>(h * a.W + w) * a.Cin : 0ll);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst0 + static_cast<size_t>(r) * 16)),
                                         "l"(src), "r"(ok ? 16 : 0)
                                         : "memory");
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                auto xform_batch = [&](int i0) {
                    uint4 raw[4];
                    const unsigned m4 = static_cast<unsigned>(okmask >> i0) & 15u;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (r0 + (i0 + u) * rpp < rows) raw[u] = *reinterpret_cast<const uint4*>(dst0 + static_cast<size_t>(r0 + (i0 + u) * rpp) * 16);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (r0 + (i0 + u) * rpp < rows) {
                            const uint32_t keep = ((m4 >> u) & 1u) ? 0xFFFFFFFFu : 0u;         // padding rows stay zero
                            *reinterpret_cast<uint4*>(dst0 + static_cast<size_t>(r0 + (i0 + u) * rpp) * 16) = make_uint4(
                                rb_bn_elu_tf32(__uint_as_float(raw[u].x), sc.x, sh.x) & keep, rb_bn_elu_tf32(__uint_as_float(raw[u].y), sc.y, sh.y) & keep,
                                rb_bn_elu_tf32(__uint_as_float(raw[u].z), sc.z, sh.z) & keep, rb_bn_elu_tf32(__uint_as_float(raw[u].w), sc.w, sh.w) & keep);
                        }
                    }
                };
                constexpr int kAhead = 3;
                int i0 = 0;
                for (; r0 + i0 * rpp < rows; i0 += 4) {
                    copy_batch(i0);
                    if (i0 >= 4 * kAhead) {
                        asm volatile("cp.async.wait_group %0;" ::"n"(kAhead) : "memory");
                        xform_batch(i0 - 4 * kAhead);
                    }
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                for (int j0 = i0 >= 4 * kAhead ? i0 - 4 * kAhead : 0; j0 < i0; j0 += 4) xform_batch(j0);
            }
            fence_proxy_async_smem();            // generic-proxy slab writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) ps_arrive(&bar.xfull[b]);
            if (tid == 64) stamp(k, 1);
        }
    } else if (warp < kPsE2Warp0) {
        // ================= epilogue 1: accumulator 1[k & 1] -> + b1 -> BN2 -> ELU -> TF32 -> u slab =================
        const int quarter = hw & 3;              // TMEM lane quarter of the HARDWARE warp
        const int te = tid - kPsE1Warp0 * 32;     // 0..127 within the role
        const float4* par4 = reinterpret_cast<const float4*>(par);
        for (int k = 0; k < my; ++k) {
            const PsItem it = ps_item(a, k);
            const int b = k & 1;
            rb_wait(&bar.a1full[b], static_cast<uint32_t>((k >> 1) & 1));            // conv1 (k) complete
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (te == 0) stamp(k, 6);
            for (int t = 0; t < it.Tc; ++t) {
                const int i = t * 128 + quarter * 32 + lane;      // u slab row = padded index Qc + i; conv1's flat output j = Qc - 1 + i
                const int j = it.Qc - 1 + i;
                bool valid = j >= 0 && j < a.total_q;
                if (valid) {
                    const int wj = static_cast<int>(__umulhi(static_cast<unsigned>(j), a.fp_magic));
                    valid = j - wj * a.Fp < a.H;
                }
                const uint32_t keep = valid ? 0xFFFFFFFFu : 0u;   // junk rows ARE conv2's zero padding
                uint32_t r[32];
                rb_tmem_ld32(tmem + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>((b * T + t) * NT), r);
                uint4 o[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 bb = par4[g], sc = par4[(NT >> 2) + g], sh = par4[(NT >> 1) + g];
                    if constexpr (F16) {              // o[g >> 1] = one channel octet (quads g, g + 1) as eight halves
                        const uint32_t h0 = rb_pack_h2(rb_bn_elu(__uint_as_float(r[4 * g]) + bb.x, sc.x, sh.x),
                                                       rb_bn_elu(__uint_as_float(r[4 * g + 1]) + bb.y, sc.y, sh.y)) & keep;
                        const uint32_t h1 = rb_pack_h2(rb_bn_elu(__uint_as_float(r[4 * g + 2]) + bb.z, sc.z, sh.z),
                                                       rb_bn_elu(__uint_as_float(r[4 * g + 3]) + bb.w, sc.w, sh.w)) & keep;
                        if (g & 1) { o[g >> 1].z = h0; o[g >> 1].w = h1; } else { o[g >> 1].x = h0; o[g >> 1].y = h1; }
                    } else
                    o[g] = make_uint4(rb_bn_elu_tf32(__uint_as_float(r[4 * g]) + bb.x, sc.x, sh.x) & keep,
                                      rb_bn_elu_tf32(__uint_as_float(r[4 * g + 1]) + bb.y, sc.y, sh.y) & keep,
                                      rb_bn_elu_tf32(__uint_as_float(r[4 * g + 2]) + bb.z, sc.z, sh.z) & keep,
                                      rb_bn_elu_tf32(__uint_as_float(r[4 * g + 3]) + bb.w, sc.w, sh.w) & keep);
                }
                // the slab is free once conv2 (k-1) has completed (it was issued right behind conv1 (k)); with two slabs: conv2 (k-2)
                if (t == 0 && !u2 && k >= 1) rb_wait(&bar.uempty[0], static_cast<uint32_t>((k - 1) & 1));
                if (t == 0 && u2 && k >= 2) rb_wait(&bar.uempty[k & 1], static_cast<uint32_t>(((k >> 1) - 1) & 1));
                if (t == 0 && te == 0) stamp(k, 7);
                unsigned char* dst = us + (u2 ? (k & 1) * a.u_bytes : 0u) + static_cast<size_t>(i) * 16;
#pragma unroll
                for (int g = 0; g < (F16 ? 4 : 8); ++g) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(g) * a.RsU * 16) = o[g];
            }
            // rows 128 Tc .. + 2: zero (they only feed outputs that are never stored)
            if (te < 3 * (NT / (F16 ? 8 : 4)))
                *reinterpret_cast<uint4*>(us + (u2 ? (k & 1) * a.u_bytes : 0u) + (static_cast<size_t>(te / 3) * a.RsU + it.Tc * 128 + te % 3) * 16) =
                    make_uint4(0u, 0u, 0u, 0u);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                ps_arrive(&bar.a1empty[b]);
                ps_arrive(&bar.ufull[u2 ? (k & 1) : 0]);
            }
            if (te == 0) stamp(k, 8);
        }
    } else {
        // ================= epilogue 2: accumulator 2 -> staging -> + b2 (+ residual | row max) -> NHWC =================
        const int quarter = hw & 3;              // TMEM lane quarter of the HARDWARE warp
        float* stg = stg_all + (warp - kPsE2Warp0) * (32 * 36);
        const int seg = lane & 7, rsub = lane >> 3;
        constexpr int kRowsPerLane = HPOOL ? 4 : 8;
        const float4 bv = __ldg(reinterpret_cast<const float4*>(a.b2) + seg);
        float4 ysc = make_float4(0.f, 0.f, 0.f, 0.f), ysh = ysc;      // F16: the next block's BN1 (its activated fp16 operand is written here)
        if (F16 && !HPOOL && a.ya) {
            ysc = __ldg(reinterpret_cast<const float4*>(a.ya_scale) + seg);
            ysh = __ldg(reinterpret_cast<const float4*>(a.ya_shift) + seg);
        }
        for (int k = 0; k < my; ++k) {
            const PsItem it = ps_item(a, k);
            const long long imgbase = static_cast<long long>(it.img) * (HPOOL ? a.img_pixels / 2 : a.img_pixels);
            int pixoff[kRowsPerLane];               // pixel index inside the (half-pooled) image, -1: junk row
            float4 rr[RES ? 8 : 1];
            auto prefetch = [&](int t) {
                if constexpr (HPOOL) {
                    const int rb = t * 128 + quarter * 32 + 2 * rsub;     // rows (rb + 8 i, rb + 8 i + 1): even chunk start, even pitch
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = rb + 8 * i, q = it.Qc + row;
                        const int wq = static_cast<int>(__umulhi(static_cast<unsigned>(q), a.fp_magic)), hq = q - wq * a.Fp;
                        pixoff[i] = (row < it.nq && hq < a.H) ? (hq >> 1) * a.W + wq : -1;
                    }
                } else {
                    const int rb = t * 128 + quarter * 32 + rsub;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = rb + 4 * i, q = it.Qc + row;
                        const int wq = static_cast<int>(__umulhi(static_cast<unsigned>(q), a.fp_magic)), hq = q - wq * a.Fp;
                        const bool valid = row < it.nq && hq < a.H;
                        pixoff[i] = valid ? hq * a.W + wq : -1;
                        if (RES) rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (RES && valid) rr[i] = __ldg(reinterpret_cast<const float4*>(a.res + (imgbase + pixoff[i]) * a.res_row_stride) + seg);
                    }
                }
            };
            if constexpr (RES) {
                // pull the residual rows of the NEXT item into L2 now (the block input is larger than L2: from DRAM every tile's
                // residual load cost the whole memory latency, four times per item, with nothing to hide it behind — e2 20 k cycles)
                auto l2_prefetch = [&](int kk) {
                    const PsItem nx = ps_item(a, kk);
                    const long long nbase = static_cast<long long>(nx.img) * a.img_pixels;
                    for (int r = (warp - kPsE2Warp0) * 32 + lane; r < nx.nq; r += 128) {
                        const int q = nx.Qc + r;
                        const int wq = static_cast<int>(__umulhi(static_cast<unsigned>(q), a.fp_magic)), hq = q - wq * a.Fp;
                        if (hq < a.H)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + (nbase + hq * a.W + wq) * a.res_row_stride));
                    }
                };
                if (k == 0) l2_prefetch(0);
                if (k + 1 < my) l2_prefetch(k + 1);
            }
            prefetch(0);                            // the residual's round trip overlaps conv2
            rb_wait(&bar.a2full, static_cast<uint32_t>(k & 1));                      // conv2 (k) complete
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == kPsE2Warp0 * 32) stamp(k, 9);
            for (int t = 0; t < it.Tc; ++t) {
                uint32_t r[32];
                rb_tmem_ld32(tmem + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>((2 * T + t) * NT), r);
                if (t == it.Tc - 1) {               // accumulator 2 is drained: conv2 (k+1) may overwrite it
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) ps_arrive(&bar.a2empty);
                    if (tid == kPsE2Warp0 * 32) stamp(k, 10);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)       // lane = row; row stride 144 B: a quarter-warp's STS.128 covers 8 distinct 16-byte slots
                    *reinterpret_cast<uint4*>(stg + lane * 36 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                __syncwarp();
                if constexpr (HPOOL) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { // row pairs (8i + 2 rsub, + 1) of the warp's 32, eight lanes (128 B) per pooled row
                        if (pixoff[i] >= 0) {
                            const float4 v0 = *reinterpret_cast<const float4*>(stg + (8 * i + 2 * rsub) * 36 + 4 * seg);
                            const float4 v1 = *reinterpret_cast<const float4*>(stg + (8 * i + 2 * rsub + 1) * 36 + 4 * seg);
                            const float4 o = make_float4(fmaxf(v0.x + bv.x, v1.x + bv.x), fmaxf(v0.y + bv.y, v1.y + bv.y),
                                                         fmaxf(v0.z + bv.z, v1.z + bv.z), fmaxf(v0.w + bv.w, v1.w + bv.w));
                            if (F16 && a.y_f16)
                                *(reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(a.y) + (imgbase + pixoff[i]) * NT) + seg) =
                                    make_uint2(rb_pack_h2(o.x, o.y), rb_pack_h2(o.z, o.w));
                            else
                                *(reinterpret_cast<float4*>(a.y + (imgbase + pixoff[i]) * NT) + seg) = o;
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { // rows 4i .. 4i+3 of the warp's 32, eight lanes (128 B) per row
                        if (pixoff[i] >= 0) {
                            const float4 v = *reinterpret_cast<const float4*>(stg + (4 * i + rsub) * 36 + 4 * seg);
                            const float4 o = RES ? make_float4(v.x + bv.x + rr[i].x, v.y + bv.y + rr[i].y, v.z + bv.z + rr[i].z, v.w + bv.w + rr[i].w)
                                                 : make_float4(v.x + bv.x, v.y + bv.y, v.z + bv.z, v.w + bv.w);
                            *(reinterpret_cast<float4*>(a.y + (imgbase + pixoff[i]) * NT) + seg) = o;
                            if (F16 && a.ya)
                                *(reinterpret_cast<uint2*>(static_cast<uint16_t*>(a.ya) + (imgbase + pixoff[i]) * NT) + seg) =
                                    make_uint2(rb_pack_h2(rb_bn_elu(o.x, ysc.x, ysh.x), rb_bn_elu(o.y, ysc.y, ysh.y)),
                                               rb_pack_h2(rb_bn_elu(o.z, ysc.z, ysh.z), rb_bn_elu(o.w, ysc.w, ysh.w)));
                        }
                    }
                }
                __syncwarp();                     // the staging tile is rewritten by the next unit
                if (t + 1 < it.Tc) prefetch(t + 1);
            }
            if (tid == kPsE2Warp0 * 32) stamp(k, 11);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
}

long long* g_ps_stamps = nullptr;         // mmla_debug_resblock2d_persist_stamps: 4 launches x 4 items x 16 slots
int g_ps_stamp_row = 0;

template <bool RES, bool STEM, bool HPOOL, bool F16 = false>
int launch_ps(const RbArgs& s, unsigned grid, size_t smem, cudaStream_t st) {
    static size_t attr[64] = {};                                  // per device: function attributes are per device
    int dev = 0;
    MMLA_CUDA_CHECK(cudaGetDevice(&dev));
    MMLA_REQUIRE(dev >= 0 && dev < 64, MMLA_EUNSUP, "resblock2d: device ordinal %d out of range", dev);
    auto kern = resblock2d_persist_kernel<RES, STEM, HPOOL, F16>;
    if (smem > attr[dev]) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        attr[dev] = smem;
    }
    kern<<<grid, kPsThreads, smem, st>>>(s);
    mmla_count_launch(F16 ? (STEM ? "stem_resblock2d_persist_f16_kernel" : "resblock2d_persist_f16_kernel")
                          : STEM ? "stem_resblock2d_persist_kernel" : "resblock2d_persist_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}

}  // namespace

// The persistent kernel takes the C = 32 blocks (Cin 16 or 32) whose slabs and weights fit one SM's shared memory
// (MMLA_NET_PERSIST=0: resblock2d_fused_kernel everywhere).  Returns 1 if it launched, 0 if the caller should use the
// one-CTA-per-item kernel, < 0 on error (MMLA_* code negated).
int mmla_try_launch_resblock2d_persist(const float* x, float* y, long long B, int H, int W, int Cin, int C, const float* bn1_scale,
                                       const float* bn1_shift, const float* w1, const float* b1, const float* bn2_scale,
                                       const float* bn2_shift, const float* w2, const float* b2, const float* res,
                                       long long res_row_stride, cudaStream_t st, const void* img, int img_is_u8,
                                       const float* stem_w, const float* stem_b, int hpool, const void* w1_h, const void* w2_h, int y_f16,
                                       const void* xa, void* ya, const float* ya_scale, const float* ya_shift) {
    // w1_h / w2_h: fp16 weight chunks (mmla_rb_arrange_weights_f16) => the fp16-operand form: the stem block, or a block whose
    // producer has written the activated fp16 operand xa (the fill is then a plain asynchronous copy)
    const bool f16 = w1_h && w2_h;
    if (f16 && !img && !xa) return 0;
    if (f16) { w1 = static_cast<const float*>(w1_h); w2 = static_cast<const float*>(w2_h); }
    if (!f16 && (xa || ya)) return 0;
    const char* e = getenv("MMLA_NET_PERSIST");             // 0: never, 2: every C = 32 block, default: Cin = 16 (block 1) only
    if (e && e[0] == '0') return 0;
    // Measured (512 clips): block 1 (128 x 151, 16 -> 32) 0.866 -> 0.757 ms; blocks 2, 3 (64 x 76, 32 -> 32) 0.412 -> 0.417 ms — with
    // every role running at once the N = 32 MMAs take 83 instead of 45 cycles there (the SM's shared-memory bandwidth is shared
    // by 520 KB of operand fetches and 250 KB of fill / epilogue traffic per item), which is what the co-resident CTAs of the
    // one-CTA-per-item kernel already reached.
    // fp16 operands, blocks 2, 3 (MMLA_NET_PERSIST=2; needs the activated fp16 operand xa): 0.299 -> 0.477 ms — the four epilogue-2
    // warps need 17.5 k cycles per item for the residual add + the next block's fp16 operand (3.5-4 k per tile and warp, latency-
    // bound chains; the one-CTA-per-item kernel puts eight warps on it and hides the rest behind co-resident CTAs), and every other
    // role waits for them (profiles/r02/experiment_notes.txt).
    if (Cin != 16 && !(e && e[0] == '2')) return 0;
    if (C != 32 || (Cin != 16 && Cin != 32) || (img && Cin != 16) || (hpool && (res || (H & 1))) || (img && res)) return 0;
    if (res && res_row_stride != C) return 0;
    RbArgs s;
    memset(&s, 0, sizeof(s));
    s.x = x; s.y = y; s.w1 = w1; s.w2 = w2; s.b1 = b1; s.b2 = b2;
    s.bn1_scale = bn1_scale; s.bn1_shift = bn1_shift; s.bn2_scale = bn2_scale; s.bn2_shift = bn2_shift;
    s.res = res; s.res_row_stride = res_row_stride;
    s.img = img; s.img_is_u8 = img_is_u8; s.stem_w = stem_w; s.stem_b = stem_b;
    s.xa = img ? nullptr : xa; s.ya = ya; s.ya_scale = ya_scale; s.ya_shift = ya_shift;
    if (ya && (hpool || !ya_scale || !ya_shift)) return 0;
    s.img_pixels = static_cast<long long>(H) * W;
    s.hpool = hpool;
    s.y_f16 = f16 && hpool ? y_f16 : 0;
    if (y_f16 && !s.y_f16) return 0;
    const int drop = hpool ? 4 : 3;
    s.H = H; s.W = W; s.Fp = H + drop;
    s.fp_magic = static_cast<unsigned>((1ULL << 32) / static_cast<unsigned>(s.Fp)) + 1u;
    s.total_q = W * s.Fp;
    s.Cin = Cin;
    s.lq = Cin == 16 ? 2 : 3;
    const int K1 = 9 * Cin, K2 = 4 * C;
    const int BK = f16 ? kRbBKh : kRbBK, cpr = f16 ? 8 : 4;      // K per weight chunk, channels per 16-byte slab row
    s.nk1 = (K1 + BK - 1) / BK;
    s.nk = s.nk1 + K2 / BK;
    const size_t chunk = static_cast<size_t>(8) * C * 16;
    auto rows_x = [&](int T) {
        int r = T * 128 + 2 * s.Fp + 2;
        if (f16) return r | 1;
        if (Cin == 16) { while ((r & 7) != 2) ++r; } else if ((r & 1) == 0) ++r;     // conflict-free fill stores (conv_slab.cu)
        return r;
    };
    auto rows_u = [&](int T) { return (T * 128 + 3) | 1; };
    int force_t = 0;
    if (const char* f = getenv("MMLA_RB_TILES")) force_t = atoi(f);
    int T = 0;
    size_t xb = 0, ub = 0;
    for (int t = kRbMaxTiles; t >= 1; --t) {
        if (force_t >= 1 && force_t <= kRbMaxTiles && t != force_t) continue;
        xb = (static_cast<size_t>(Cin / cpr) * rows_x(t) * 16 + 127) / 128 * 128;
        ub = (static_cast<size_t>(C / cpr) * rows_u(t) * 16 + 127) / 128 * 128;
        const size_t total = 2 * xb + (f16 ? 2 : 1) * ub + s.nk * chunk + 4 * 32 * 36 * 4 + 512 + 256 + 256;
        if (total <= 226 * 1024) { T = t; break; }
    }
    if (!T) return 0;
    {
        const int need = (s.total_q + 3 + 127) / 128;             // a whole image in one item
        if (T > need) {
            T = need;
            xb = (static_cast<size_t>(Cin / cpr) * rows_x(T) * 16 + 127) / 128 * 128;
            ub = (static_cast<size_t>(C / cpr) * rows_u(T) * 16 + 127) / 128 * 128;
        }
    }
    s.T = T;
    s.S = T * 128 - drop;
    s.cpi = (s.total_q + s.S - 1) / s.S;
    s.RsX = rows_x(T);
    s.RsU = rows_u(T);
    for (int kc = 0; kc < s.nk; ++kc)
        for (int kk = 0; kk < 4; ++kk) {
            s.aoff[kc * 4 + kk] = 0;
            if (kc < s.nk1) {
                const int k = kc * BK + kk * (BK / 4);
                if (k < K1) {
                    const int tap = k / Cin, c0 = k % Cin;
                    const int dh = tap / 3, dw = tap % 3;         // Keras HWIO: tap = kernel row * 3 + kernel column
                    s.aoff[kc * 4 + kk] = static_cast<unsigned>((c0 / cpr) * s.RsX + dw * s.Fp + dh);
                    if (kc == s.nk1 - 1) s.nmma1_last = kk + 1;
                }
            } else {
                const int k = (kc - s.nk1) * BK + kk * (BK / 4);
                const int dh = k / C, c0 = k % C;
                s.aoff[kc * 4 + kk] = static_cast<unsigned>((c0 / cpr) * s.RsU + dh);
                if (kc == s.nk - 1) s.nmma2_last = kk + 1;
            }
        }
    s.x1_off = static_cast<unsigned>(xb);
    s.u_off = static_cast<unsigned>(2 * xb);
    s.u_bytes = static_cast<unsigned>(ub);
    s.u_bufs = f16 ? 2 : 1;                                      // half-size slabs: room for a second u slab
    s.ring_off = s.u_off + static_cast<unsigned>(ub) * s.u_bufs;
    s.stg_off = s.ring_off + static_cast<unsigned>(s.nk * chunk);
    s.par_off = s.stg_off + 4 * 32 * 36 * 4;
    s.bar_off = s.par_off + 512;
    const size_t smem = s.bar_off + 256 + 128;
    const long long items = B * s.cpi;
    if (items >= (1LL << 31) - 2 || B * s.img_pixels * 32 >= (1LL << 40)) return 0;
    s.n_ctas = static_cast<int>(items);
    {
        const char* mh = getenv("MMLA_PS_MMA_HI");
        s.mma_hi = (mh && mh[0] == '1') ? 1 : 0;
    }
    const int sms = mmla_num_sms();
    if (sms <= 0) return 0;
    const unsigned grid = static_cast<unsigned>(items < sms ? items : sms);
    if (getenv("MMLA_RB_VERBOSE"))
        fprintf(stderr, "resblock2d (persistent): %dx%d Cin %d C %d: %d outputs/image, T %d, %lld items on %u CTAs, %zu KB smem\n", H, W, Cin, C,
                s.total_q, s.T, items, grid, smem / 1024);
    if (g_ps_stamps && g_ps_stamp_row < 4) {
        s.stamps = g_ps_stamps + 64 * g_ps_stamp_row++;
        s.stamp_cta = static_cast<int>(grid / 2);
    }
    int rc;
    if (img && f16) rc = hpool ? launch_ps<false, true, true, true>(s, grid, smem, st) : launch_ps<false, true, false, true>(s, grid, smem, st);
    else if (f16 && res) rc = launch_ps<true, false, false, true>(s, grid, smem, st);
    else if (f16) rc = hpool ? launch_ps<false, false, true, true>(s, grid, smem, st) : launch_ps<false, false, false, true>(s, grid, smem, st);
    else if (img) rc = hpool ? launch_ps<false, true, true>(s, grid, smem, st) : launch_ps<false, true, false>(s, grid, smem, st);
    else if (res) rc = launch_ps<true, false, false>(s, grid, smem, st);
    else rc = hpool ? launch_ps<false, false, true>(s, grid, smem, st) : launch_ps<false, false, false>(s, grid, smem, st);
    return rc == MMLA_OK ? 1 : -rc;
}

extern "C" __attribute__((visibility("default"))) void mmla_debug_resblock2d_persist_stamps(long long* dev_stamps) {
    g_ps_stamps = dev_stamps;
    g_ps_stamp_row = 0;
}
