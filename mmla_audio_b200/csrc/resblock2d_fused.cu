// One launch per `res_block` conv pair of the overlap classifier (overlap_detector_temp.py:253-280):
//
//     u = Conv2D(C, 3, 'same')(ELU(BN1(x)))          v = Conv2D(C, (4, 1), 'same')(ELU(BN2(u)))  (+ x for a plain block)
//
// conv_slab_kernel ran the two convolutions as two launches with u (fp32, up to 1.27 GB per 512 clips) written to HBM by
// the first and gathered + activated again by the second: 6.7 GB of the net's 12.6 GB of DRAM traffic per 512 clips was u.
// Here u never leaves the SM.  Both convolutions use the same flat pixel numbering with the image HEIGHT as the fast axis
// and Fp = H + 3 entries per column (Keras 'same' for k = 4: one row before, two after; the 3x3 needs 1 + 1 <= 3):
//
//     q = w * Fp + h                                  flat index of an output pixel (h >= H: junk rows, 2 %, never stored)
//     conv1:  u(j) = b1 + sum_{dh,dw} W1[dh][dw] . xpad[j + dw * Fp + dh]       xpad[(w + 1) * Fp + h + 1] = ELU(BN1(x[h, w]))
//     conv2:  v(q) = b2 + sum_{dh}    W2[dh]     . upad[q + dh]                 upad[j + 1] = valid(j) ? ELU(BN2(u(j))) : 0
//
// i.e. every filter tap of either convolution is a UNIFORM row shift of one operand slab (csrc/conv_slab.cu), and conv1's
// accumulator row j is exactly conv2's padded operand row j + 1 with the junk rows replaced by the zero padding.  A CTA owns
// S = 128 T - 3 consecutive outputs q of one image (T <= 4 tiles of 128 rows):
//   fill      x rows [Q - 1, Q - 1 + 128 T + 2 Fp + 2) by 16-byte cp.async, BN1 + ELU + TF32 rounding in place (as conv_slab)
//   conv1     T accumulators of C columns in TMEM, weights through a TMA ring ([32 x C] K-chunks, conv_tc.cu arrangement)
//   epilogue1 tcgen05.ld -> + b1 -> BN2 -> ELU -> TF32 -> the u slab [channel quad][row][16 B], which OVERLAYS the dead x slab
//   conv2     the same TMEM columns and the same weight ring (its chunks follow conv1's; they are prefetched into the slots
//             conv1 frees while epilogue 1 runs)
//   epilogue2 tcgen05.ld -> staging tiles in the dead slab -> + b2 (+ residual) -> whole-line NHWC stores
// Arithmetic per element is the same sequence of operations as the two conv_slab launches (same TF32 operands, same K order,
// same fp32 epilogue expressions), so the results are bit-identical to them (tests/test_resblock2d_gpu.py).
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "resblock2d_common.cuh"

namespace {

// STEM: the block's input is the net's stem Conv2D(16, 1x1) of the 3-channel classifier image, computed in the fill from the
// image bytes (the expression of stem1x1_kernel, csrc/nets.cu): the [B,128,151,16] stem tensor is never written or read.
// HPOOL (pooled blocks): the MaxPool2D(2)'s maximum over the two ROWS of each window is taken in epilogue 2 — with an even
// column pitch (Fp = H + 4) and an even chunk length the partners are adjacent accumulator rows of one warp's staging tile —
// so the kernel writes [B, H/2, W, C], half of the block output, and pool_shortcut_kernel reads half as much.
// PAIR (C >= 64): two CTAs of a cluster run every MMA together (`tcgen05.mma.cta_group::2`, M = 256: 128 rows of each CTA's
// slab) and each holds HALF of every weight chunk (N / 2 columns), so the same ring bytes keep twice as many chunks in flight
// and the weight stream out of L2 halves — the C >= 64 blocks wait on exactly that (a ring slot is re-requested when its
// MMAs complete and lands ~3 000 cycles later).  OPT-IN, see mmla_rb_pair_wanted: it measured slower.  The leader (rank 0) issues; the peer's warp 0 relays "my half has landed" to
// the leader's `pfull` barriers; `empty` / `accum` are arrived in both CTAs by multicast commits; cluster barriers replace the
// CTA barriers where the leader's MMAs read the peer's slab.
// F16 (precision mode "fp16" of the overlap net): both convolutions' operands are fp16 (activations converted in the fill /
// in epilogue 1, weights arranged by mmla_rb_arrange_weights_f16), accumulation stays fp32 in TMEM (`kind::f16`).  An operand
// slab row is then 8 channels per 16 bytes ([channel octet][row][16 B]) and one MMA covers K = 16: half the operand bytes
// through shared memory, half the MMAs and half the ring chunks per output of the TF32 form, which is what bounds these
// kernels.  Same 11 significant bits per operand as TF32; NOT bit-identical to it (round-to-nearest-even
// instead of round-half-up, K = 16 summation groups).
template <int NT, bool RES, int THREADS, bool STEM = false, bool HPOOL = false, bool PAIR = false, bool F16 = false>
__global__ void __launch_bounds__(THREADS, THREADS == 256 && !RES ? 3 : (THREADS == 512 && RES ? 1 : 2)) resblock2d_fused_kernel(const RbArgs a) {
    static_assert(!(PAIR && F16), "the CTA-pair mode exists for TF32 operands only");
    constexpr uint32_t kIdesc = (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | (static_cast<uint32_t>(NT >> 3) << 17) |
                                (static_cast<uint32_t>((PAIR ? 256 : 128) >> 4) << 24);   // D=f32, A=B=tf32 (or f16), K-major, N, M
    constexpr uint32_t kChunkBytes = 8 * NT * 16 / (PAIR ? 2 : 1);
    constexpr int kNB = PAIR ? NT / 2 : NT;                               // weight columns this CTA holds
    constexpr int kWarps = THREADS / 32;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    unsigned char* slab = base;
    unsigned char* ring = base + a.ring_off;
    float* par = reinterpret_cast<float*>(base + a.par_off);              // b1 | bn2 scale | bn2 shift, NT floats each
    uint64_t* full = reinterpret_cast<uint64_t*>(base + a.bar_off);      // [stages]
    uint64_t* empty = full + kRbMaxStages;                               // [stages]
    uint64_t* pfull = full + 2 * kRbMaxStages;                           // [stages] PAIR, leader: the peer's half has landed
    uint64_t* accum = full + 3 * kRbMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 3 * kRbMaxStages + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool stamping = a.stamps != nullptr && static_cast<int>(blockIdx.x) == a.stamp_cta;
    auto stamp = [&](int slot) {
        if (stamping) a.stamps[slot] = clock64();
    };
    if (tid == 0) stamp(0);
    const uint32_t rank = PAIR ? rb_cluster_rank() : 0u;
    const bool dummy = PAIR && static_cast<int>(blockIdx.x) >= a.n_ctas;  // pads the grid to whole pairs: runs the protocol, no data
    const int img = dummy ? 0 : blockIdx.x / a.cpi;
    const int Qc = dummy ? 0 : (blockIdx.x - img * a.cpi) * a.S;          // first output of this CTA
    const int nq = dummy ? 0 : min(a.S, a.total_q - Qc);                  // outputs of this CTA
    const int Tc = dummy ? 0 : (nq + 3 + 127) >> 7;                       // tiles of either convolution
    const int Tm = PAIR ? a.T : Tc;                                       // tiles the MMAs run over (a pair: the same in both CTAs)
    const uint32_t cols = (Tm * NT <= 32) ? 32u : (Tm * NT <= 64) ? 64u : (Tm * NT <= 128) ? 128u : (Tm * NT <= 256) ? 256u : 512u;

    if (tid == 0) {
        for (int i = 0; i < a.stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
            if (PAIR) mbar_init(&pfull[i], 1);
        }
        mbar_init(accum, 1);
        mbar_fence_init();
    }
    // The barrier initialisation is all the weight prefetch and the fill need; the TMEM allocation (which can wait for a
    // co-resident CTA's columns) and the epilogue parameters land behind the fill, before the barrier that ends it.
    float pv[(3 * NT + THREADS - 1) / THREADS];
#pragma unroll
    for (int i = 0; i < (3 * NT + THREADS - 1) / THREADS; ++i) {
        const int e = tid + i * THREADS;
        pv[i] = 0.f;
        if (e < 3 * NT) pv[i] = __ldg((e < NT ? a.b1 : e < 2 * NT ? a.bn2_scale : a.bn2_shift) + (e % NT));
    }
    __syncthreads();
    if (tid == 0) stamp(1);

    auto wsrc = [&](int kc) {                 // PAIR: the pair arrangement keeps this CTA's half of a chunk contiguous
        return (kc < a.nk1 ? a.w1 + static_cast<long long>(kc) * (NT * kRbBK) : a.w2 + static_cast<long long>(kc - a.nk1) * (NT * kRbBK)) +
               rank * (kNB * kRbBK);
    };
    // ---- weight ring: the first `stages` chunks need no free slot, so they are requested before the fill ----
    if (lane == 0) {                          // one chunk per warp at a time: bulk copies of one thread serialise
        const int pre = a.nk < a.stages ? a.nk : a.stages;
        for (int kc = warp; kc < pre; kc += kWarps) {
            mbar_arrive_expect_tx(&full[kc], kChunkBytes);
            tma_bulk_g2s(ring + kc * kChunkBytes, wsrc(kc), kChunkBytes, &full[kc]);
        }
    }

    if constexpr (STEM) {
        // ---- x slab fill from the image: 3 values per pixel -> 16 stem channels -> BN1 + ELU + TF32, one pass, no staging ----
        const int rows = dummy ? 0 : Tc * 128 + 2 * a.Fp + 2;
        const int c4 = lane >> 3;                             // Cin = 16: four quads, eight rows per warp instruction
        const float4 sc = __ldg(reinterpret_cast<const float4*>(a.bn1_scale) + c4);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(a.bn1_shift) + c4);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.stem_w) + c4);
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.stem_w + 16) + c4);
        const float4 w2 = __ldg(reinterpret_cast<const float4*>(a.stem_w + 32) + c4);
        const float4 sb = __ldg(reinterpret_cast<const float4*>(a.stem_b) + c4);
        const unsigned char* img8 = static_cast<const unsigned char*>(a.img) + static_cast<long long>(img) * a.img_pixels * 3;
        const float* imgf = static_cast<const float*>(a.img) + static_cast<long long>(img) * a.img_pixels * 3;
        unsigned char* dst0 = slab + static_cast<size_t>(c4) * a.RsX * 16;
        // four rows per iteration: all twelve pixel loads first, then the arithmetic (one row at a time the loop waited a
        // memory round trip per row: 13.5 k cycles per CTA against 9.4 k for the cp.async fill of the materialised tensor)
        for (int rb = warp * 8; rb < rows; rb += 4 * kWarps * 8) {      // warp-uniform trip count (the F16 form shuffles)
            const int r0 = rb + (lane & 7);
            float c[4][3];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u * kWarps * 8;
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
                // this lane's four channels of four rows as half2 pairs; lanes l and l ^ 8 hold the two quads of one octet:
                // the even quad's lane assembles rows u = 0, 2, the odd quad's lane rows u = 1, 3 (one shuffle pair per row)
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
                    const int r = r0 + (u + (odd ? 1 : 0)) * kWarps * 8;
                    if (r < rows)
                        *reinterpret_cast<uint4*>(dsth + static_cast<size_t>(r) * 16) =
                            odd ? make_uint4(g0, g1, pk[u + 1][0], pk[u + 1][1]) : make_uint4(pk[u][0], pk[u][1], g0, g1);
                }
                continue;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u * kWarps * 8;
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
    } else if (F16 && a.xa != nullptr) {
        // ---- x slab fill, fp16 operands, the producer of x has already written ELU(BN1(x)) as fp16: the fill is a plain
        // asynchronous gather, 16 bytes = one channel octet of one row per copy (zero-fill form for the padding rows), every
        // copy of the CTA in flight at once and nothing to transform ----
        const int rows = Tc * 128 + 2 * a.Fp + 2;
        const int no = a.Cin >> 3;                            // channel octets per row: 2, 4, 8, 16
        const int osh = no >= 4 ? 2 : 1;                      // a warp instruction covers 4 octets x 8 rows (2 x 16 for Cin = 16)
        const int rpi = 32 >> osh, ngrp = no >> osh;
        const int c8 = ((warp % ngrp) << osh) + (lane >> (5 - osh));
        const int rpp = (kWarps / ngrp) * rpi;
        const uint16_t* ximg = static_cast<const uint16_t*>(a.xa) + static_cast<long long>(img) * a.img_pixels * a.Cin + 8 * c8;
        unsigned char* dsth = slab + static_cast<size_t>(c8) * a.RsX * 16;
        for (int r = (warp / ngrp) * rpi + (lane & (rpi - 1)); r < rows; r += rpp) {
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
    } else if constexpr (F16) {
        // ---- x slab fill, fp16 operands: 16-byte loads of one channel quad per lane (the cp.async form's lane map: 8 rows x 4
        // quads per warp instruction), eight rows in flight per lane, BN1 + ELU in fp32, half2 packing, and the two quads of an
        // octet joined by one shuffle pair per row (lanes l, l ^ 8) before the 16-byte store ----
        const int rows = Tc * 128 + 2 * a.Fp + 2;
        const int lqg = a.lq - 2;
        const int c4 = ((warp & ((1 << lqg) - 1)) << 2) + (lane >> 3);
        const int rpp = (kWarps >> lqg) * 8;
        const float4 sc = __ldg(reinterpret_cast<const float4*>(a.bn1_scale) + c4);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(a.bn1_shift) + c4);
        const float* ximg = a.x + static_cast<long long>(img) * a.img_pixels * a.Cin + 4 * c4;
        unsigned char* dsth = slab + static_cast<size_t>(c4 >> 1) * a.RsX * 16;
        const bool odd = (c4 & 1) != 0;
        for (int rb = (warp >> lqg) * 8; rb < rows; rb += 8 * rpp) {    // warp-uniform trip count
            const int r0 = rb + (lane & 7);
            float4 raw[8];
            unsigned okm = 0u;                                          // bit u: row u holds image data (padding rows stay zero)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + u * rpp;
                const int p = Qc - 1 + r;
                const int wp = static_cast<int>(__umulhi(static_cast<unsigned>(p < 0 ? 0 : p), a.fp_magic));
                const int w = wp - 1, h = p - wp * a.Fp - 1;
                const bool ok = r < rows && p >= 0 && w >= 0 && w < a.W && h >= 0 && h < a.H;
                raw[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) raw[u] = __ldg(reinterpret_cast<const float4*>(ximg + static_cast<long long>(h * a.W + w) * a.Cin));
                okm |= static_cast<unsigned>(ok) << u;
            }
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                uint32_t pk[2][2];
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const uint32_t keep = ((okm >> (u + v)) & 1u) ? 0xFFFFFFFFu : 0u;
                    pk[v][0] = rb_pack_h2(rb_bn_elu(raw[u + v].x, sc.x, sh.x), rb_bn_elu(raw[u + v].y, sc.y, sh.y)) & keep;
                    pk[v][1] = rb_pack_h2(rb_bn_elu(raw[u + v].z, sc.z, sh.z), rb_bn_elu(raw[u + v].w, sc.w, sh.w)) & keep;
                }
                const uint32_t g0 = __shfl_xor_sync(0xffffffffu, odd ? pk[0][0] : pk[1][0], 8);
                const uint32_t g1 = __shfl_xor_sync(0xffffffffu, odd ? pk[0][1] : pk[1][1], 8);
                const int r = r0 + (u + (odd ? 1 : 0)) * rpp;
                if (r < rows)
                    *reinterpret_cast<uint4*>(dsth + static_cast<size_t>(r) * 16) =
                        odd ? make_uint4(g0, g1, pk[1][0], pk[1][1]) : make_uint4(pk[0][0], pk[0][1], g0, g1);
            }
        }
    } else
    // ---- x slab fill: row r <-> padded flat index Qc - 1 + r; BN1 + ELU + TF32 once per element (conv_slab.cu's fill) ----
    {
        const int rows = dummy ? 0 : Tc * 128 + 2 * a.Fp + 2;
        const int lqg = a.lq - 2;                             // log2(quad groups of 4)
        const int c4 = ((warp & ((1 << lqg) - 1)) << 2) + (lane >> 3);
        const int rpp = (kWarps >> lqg) * 8;                  // rows per pass of the whole CTA
        const float4 sc = __ldg(reinterpret_cast<const float4*>(a.bn1_scale) + c4);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(a.bn1_shift) + c4);
        const float* ximg = a.x + static_cast<long long>(img) * a.img_pixels * a.Cin + 4 * c4;
        unsigned char* dst0 = slab + static_cast<size_t>(c4) * a.RsX * 16;
        unsigned long long okmask = 0ull;                     // bit i: row r0 + i * rpp holds image data
        const int r0 = (warp >> lqg) * 8 + (lane & 7);
        auto copy_batch = [&](int it) {                       // rows r0 + (it .. it+3) * rpp: one cp.async group
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + (it + u) * rpp;
                if (r < rows) {
                    const int p = Qc - 1 + r;
                    const int wp = static_cast<int>(__umulhi(static_cast<unsigned>(p < 0 ? 0 : p), a.fp_magic));
                    const int w = wp - 1, h = p - wp * a.Fp - 1;
                    const bool ok = p >= 0 && w >= 0 && w < a.W && h >= 0 && h < a.H;
                    okmask |= static_cast<unsigned long long>(ok) << (it + u);
                    const float* src = ximg + (ok ? static_cast<long long>(h * a.W + w) * a.Cin : 0ll);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst0 + static_cast<size_t>(r) * 16)),
                                 "l"(src), "r"(ok ? 16 : 0)
                                 : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto xform_batch = [&](int it) {
            uint4 raw[4];
            const unsigned m4 = static_cast<unsigned>(okmask >> it) & 15u;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (r0 + (it + u) * rpp < rows) raw[u] = *reinterpret_cast<const uint4*>(dst0 + static_cast<size_t>(r0 + (it + u) * rpp) * 16);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (r0 + (it + u) * rpp < rows) {
                    const uint32_t keep = ((m4 >> u) & 1u) ? 0xFFFFFFFFu : 0u;         // padding rows stay zero
                    *reinterpret_cast<uint4*>(dst0 + static_cast<size_t>(r0 + (it + u) * rpp) * 16) = make_uint4(
                        rb_bn_elu_tf32(__uint_as_float(raw[u].x), sc.x, sh.x) & keep, rb_bn_elu_tf32(__uint_as_float(raw[u].y), sc.y, sh.y) & keep,
                        rb_bn_elu_tf32(__uint_as_float(raw[u].z), sc.z, sh.z) & keep, rb_bn_elu_tf32(__uint_as_float(raw[u].w), sc.w, sh.w) & keep);
                }
            }
        };
        constexpr int kAhead = 3;
        int it = 0;
        for (; r0 + it * rpp < rows; it += 4) {
            copy_batch(it);
            if (it >= 4 * kAhead) {
                asm volatile("cp.async.wait_group %0;" ::"n"(kAhead) : "memory");
                xform_batch(it - 4 * kAhead);
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        for (int jt = it >= 4 * kAhead ? it - 4 * kAhead : 0; jt < it; jt += 4) xform_batch(jt);
    }
    if (tid == 0) stamp(2);
    if (warp == 0) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
#pragma unroll
    for (int i = 0; i < (3 * NT + THREADS - 1) / THREADS; ++i)
        if (tid + i * THREADS < 3 * NT) par[tid + i * THREADS] = pv[i];
    fence_proxy_async_smem();                // generic-proxy slab writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if constexpr (PAIR) rb_cluster_sync();   // the leader's MMAs read both slabs; both CTAs' barriers are initialised
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) stamp(3);

    // Ring state.  Weight producers: warps 1 .. kProducers, warp w owns chunks kc = stages + (w - 1) + i * kProducers (one
    // thread gets one bulk copy through every ~420 cycles whatever its size, so a 16 KB chunk per 4 x 64-cycle MMAs needs
    // more than one producer: scripts/microbench/tma_stream.cu).  MMA issuer: warp 0.
    // (never more producers than ring slots: a producer two phases ahead of a slot's `empty` barrier would see the 1-bit
    //  parity of the phase before last and overwrite a chunk that has not been consumed)
    const int kProducers = a.stages < 3 ? a.stages : 3;
    int p_kc = a.stages + (warp - 1);         // next chunk this producer warp requests
    int m_stg = 0;
    uint32_t m_ph = 0u;
    auto produce = [&](int kc_end) {          // this warp's chunks below kc_end
        for (; p_kc < kc_end; p_kc += kProducers) {
            const int use = p_kc / a.stages, stg = p_kc - use * a.stages;
            rb_wait(&empty[stg], static_cast<uint32_t>((use - 1) & 1));
            if (lane == 0) {
                mbar_arrive_expect_tx(&full[stg], kChunkBytes);
                tma_bulk_g2s(ring + stg * kChunkBytes, wsrc(p_kc), kChunkBytes, &full[stg]);
            }
            __syncwarp();
        }
    };
    auto issue = [&](int kc0, int kc1, int nmma_last, uint32_t rs16) {   // chunks [kc0, kc1) of one convolution
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t dA = rb_desc(smem_u32(slab), rs16, 128u);
        const uint64_t dB = rb_desc(smem_u32(ring), kNB * 16, 128);
        for (int kc = kc0; kc < kc1; ++kc) {
            rb_wait(&full[m_stg], m_ph);
            if constexpr (PAIR) rb_wait(&pfull[m_stg], m_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t bd0 = dB + static_cast<uint64_t>(m_stg * (kChunkBytes / 16));
            const uint32_t aoff[4] = {a.aoff[kc * 4], a.aoff[kc * 4 + 1], a.aoff[kc * 4 + 2], a.aoff[kc * 4 + 3]};
            const int nmma = kc == kc1 - 1 ? nmma_last : 4;
            const uint32_t alo = static_cast<uint32_t>(dA), ahi = static_cast<uint32_t>(dA >> 32);
            const uint32_t blo = static_cast<uint32_t>(bd0), bhi = static_cast<uint32_t>(bd0 >> 32);
            if (rb_elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                    for (int t = 0; t < kRbMaxTiles; ++t) {
                        if (kk < nmma && t < Tm) {
                            const uint32_t acc = (kc != kc0 || kk != 0) ? 1u : 0u;
                            if constexpr (PAIR) {
                                asm volatile(
                                    "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\n"
                                    "mov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
                                    "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n}\n" ::"r"(tmem + static_cast<uint32_t>(t * NT)),
                                    "r"(alo + aoff[kk] + static_cast<uint32_t>(t * 128)), "r"(ahi), "r"(blo + static_cast<uint32_t>(kk * 2 * kNB)),
                                    "r"(bhi), "r"(kIdesc), "r"(acc)
                                    : "memory");
                            } else if constexpr (F16) {
                                asm volatile(
                                    "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\n"
                                    "mov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
                                    "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tmem + static_cast<uint32_t>(t * NT)),
                                    "r"(alo + aoff[kk] + static_cast<uint32_t>(t * 128)), "r"(ahi), "r"(blo + static_cast<uint32_t>(kk * 2 * kNB)),
                                    "r"(bhi), "r"(kIdesc), "r"(acc)
                                    : "memory");
                            } else {
                                asm volatile(
                                    "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\n"
                                    "mov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
                                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n}\n" ::"r"(tmem + static_cast<uint32_t>(t * NT)),
                                    "r"(alo + aoff[kk] + static_cast<uint32_t>(t * 128)), "r"(ahi), "r"(blo + static_cast<uint32_t>(kk * 2 * kNB)),
                                    "r"(bhi), "r"(kIdesc), "r"(acc)
                                    : "memory");
                            }
                        }
                    }
                }
                if constexpr (PAIR) {
                    rb_commit_pair(&empty[m_stg]);
                    if (kc == kc1 - 1) rb_commit_pair(accum);
                } else {
                    rb_commit(&empty[m_stg]);
                    if (kc == kc1 - 1) rb_commit(accum);
                }
            }
            if (++m_stg == a.stages) { m_stg = 0; m_ph ^= 1u; }
            __syncwarp();
        }
    };
    // PAIR, peer CTA: warp 0 tells the leader when this CTA's half of a chunk has landed
    auto relay = [&](int kc0, int kc1) {
        uint32_t remote0;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote0) : "r"(smem_u32(pfull)), "r"(0));
        for (int kc = kc0; kc < kc1; ++kc) {
            rb_wait(&full[m_stg], m_ph);
            if (lane == 0)
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote0 + 8u * m_stg) : "memory");
            if (++m_stg == a.stages) { m_stg = 0; m_ph ^= 1u; }
            __syncwarp();
        }
    };

    // ================= conv1 =================
    if (warp >= 1 && warp <= kProducers) {
        // conv2's chunks go into the slots conv1's MMAs free; the rest waits for conv2 (after epilogue 1: these warps take part)
        const int lim = a.nk1 + a.stages;
        produce(a.nk < lim ? a.nk : lim);
    } else if (warp == 0) {
        if (rank == 0) issue(0, a.nk1, a.nmma1_last, static_cast<uint32_t>(a.RsX) * 16u);
        else relay(0, a.nk1);
        if (lane == 0) stamp(4);
    }

    constexpr int kChunks = NT / 32;
    constexpr int kGroups = THREADS / 128;
    const int quarter = warp & 3, half = warp >> 2;
    const int nunits = Tc * kChunks;

    // ================= epilogue 1: accumulator -> + b1 -> BN2 -> ELU -> TF32 -> u slab (over the dead x slab) =================
    {
        if (warp == 0) rb_wait(accum, 0u);    // one polling warp, everybody else sleeps in the hardware barrier
        asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 64) stamp(5);
        const float4* par4 = reinterpret_cast<const float4*>(par);
        for (int u = half; u < nunits; u += kGroups) {
            const int t = u / kChunks, col0 = (u - t * kChunks) * 32;
            const int i = t * 128 + quarter * 32 + lane;      // u slab row = padded index Qc + i; conv1's flat output j = Qc - 1 + i
            const int j = Qc - 1 + i;
            bool valid = j >= 0 && j < a.total_q;
            if (valid) {
                const int wj = static_cast<int>(__umulhi(static_cast<unsigned>(j), a.fp_magic));
                valid = j - wj * a.Fp < a.H;
            }
            const uint32_t keep = valid ? 0xFFFFFFFFu : 0u;   // junk rows ARE conv2's zero padding
            uint32_t r[32];
            rb_tmem_ld32(tmem + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(t * NT + col0), r);
            if constexpr (F16) {
                unsigned char* dsth = slab + (static_cast<size_t>(col0 >> 3) * a.RsU + i) * 16;
#pragma unroll
                for (int g = 0; g < 4; ++g) {                 // one channel octet per 16-byte store
                    uint32_t h[4];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float4 bb = par4[(col0 >> 2) + 2 * g + e], sc = par4[(NT >> 2) + (col0 >> 2) + 2 * g + e],
                                     sh = par4[(NT >> 1) + (col0 >> 2) + 2 * g + e];
                        h[2 * e] = rb_pack_h2(rb_bn_elu(__uint_as_float(r[8 * g + 4 * e]) + bb.x, sc.x, sh.x),
                                              rb_bn_elu(__uint_as_float(r[8 * g + 4 * e + 1]) + bb.y, sc.y, sh.y)) & keep;
                        h[2 * e + 1] = rb_pack_h2(rb_bn_elu(__uint_as_float(r[8 * g + 4 * e + 2]) + bb.z, sc.z, sh.z),
                                                  rb_bn_elu(__uint_as_float(r[8 * g + 4 * e + 3]) + bb.w, sc.w, sh.w)) & keep;
                    }
                    *reinterpret_cast<uint4*>(dsth + static_cast<size_t>(g) * a.RsU * 16) = make_uint4(h[0], h[1], h[2], h[3]);
                }
                continue;
            }
            unsigned char* dst = slab + (static_cast<size_t>(col0 >> 2) * a.RsU + i) * 16;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float4 bb = par4[(col0 >> 2) + g], sc = par4[(NT >> 2) + (col0 >> 2) + g], sh = par4[(NT >> 1) + (col0 >> 2) + g];
                *reinterpret_cast<uint4*>(dst + static_cast<size_t>(g) * a.RsU * 16) =
                    make_uint4(rb_bn_elu_tf32(__uint_as_float(r[4 * g]) + bb.x, sc.x, sh.x) & keep,
                               rb_bn_elu_tf32(__uint_as_float(r[4 * g + 1]) + bb.y, sc.y, sh.y) & keep,
                               rb_bn_elu_tf32(__uint_as_float(r[4 * g + 2]) + bb.z, sc.z, sh.z) & keep,
                               rb_bn_elu_tf32(__uint_as_float(r[4 * g + 3]) + bb.w, sc.w, sh.w) & keep);
            }
        }
        // rows 128 Tc .. + 2 only feed outputs that are never stored, but they must be finite
        for (int i = tid; i < 3 * (NT / (F16 ? 8 : 4)); i += THREADS)
            *reinterpret_cast<uint4*>(slab + (static_cast<size_t>(i / 3) * a.RsU + Tc * 128 + i % 3) * 16) = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        fence_proxy_async_smem();
        if constexpr (PAIR) rb_cluster_sync();   // conv2's MMAs read both CTAs' u slabs and overwrite both accumulators
        else __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) stamp(6);
    }

    // ================= conv2 =================
    if (warp >= 1 && warp <= kProducers) {
        produce(a.nk);
    } else if (warp == 0) {
        if (rank == 0) issue(a.nk1, a.nk, a.nmma2_last, static_cast<uint32_t>(a.RsU) * 16u);
        else relay(a.nk1, a.nk);
        if (lane == 0) stamp(7);
    }

    // ================= epilogue 2: + b2 (+ residual) -> NHWC (conv_slab.cu's staged epilogue) =================
    {
        float* stg = reinterpret_cast<float*>(slab) + warp * (32 * 36);
        const int seg = lane & 7, rsub = lane >> 3;
        constexpr int kRowsPerLane = HPOOL ? 4 : 8;
        int pixoff[kRowsPerLane];               // pixel index inside the (half-pooled) image, -1: junk row
        const long long imgbase = static_cast<long long>(img) * (HPOOL ? a.img_pixels / 2 : a.img_pixels);
        float4 rr[RES ? 8 : 1];
        auto prefetch = [&](int u) {
            const int t = u / kChunks, col0 = (u - t * kChunks) * 32;
            if constexpr (HPOOL) {
                const int rb = t * 128 + quarter * 32 + 2 * rsub;     // rows (rb + 8 i, rb + 8 i + 1): even chunk start, even pitch
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = rb + 8 * i, q = Qc + row;
                    const int wq = static_cast<int>(__umulhi(static_cast<unsigned>(q), a.fp_magic)), hq = q - wq * a.Fp;
                    pixoff[i] = (row < nq && hq < a.H) ? (hq >> 1) * a.W + wq : -1;
                }
            } else {
                const int rb = t * 128 + quarter * 32 + rsub;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = rb + 4 * i, q = Qc + row;
                    const int wq = static_cast<int>(__umulhi(static_cast<unsigned>(q), a.fp_magic)), hq = q - wq * a.Fp;
                    const bool valid = row < nq && hq < a.H;
                    pixoff[i] = valid ? hq * a.W + wq : -1;
                    if (RES) rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (RES && valid) rr[i] = __ldg(reinterpret_cast<const float4*>(a.res + (imgbase + pixoff[i]) * a.res_row_stride + col0) + seg);
                }
            }
        };
        if (half < nunits) prefetch(half);
        if (warp == 0) rb_wait(accum, 1u);
        asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 64) stamp(8);
        for (int u = half; u < nunits; u += kGroups) {
            const int t = u / kChunks, col0 = (u - t * kChunks) * 32;
            uint32_t r[32];
            rb_tmem_ld32(tmem + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(t * NT + col0), r);
#pragma unroll
            for (int j = 0; j < 8; ++j)       // lane = row; row stride 144 B: a quarter-warp's STS.128 covers 8 distinct 16-byte slots
                *reinterpret_cast<uint4*>(stg + lane * 36 + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            __syncwarp();
            const float4 bv = __ldg(reinterpret_cast<const float4*>(a.b2 + col0) + seg);
            float4 ysc = make_float4(0.f, 0.f, 0.f, 0.f), ysh = ysc;
            if (F16 && !HPOOL && a.ya) {
                ysc = __ldg(reinterpret_cast<const float4*>(a.ya_scale + col0) + seg);
                ysh = __ldg(reinterpret_cast<const float4*>(a.ya_shift + col0) + seg);
            }
            if constexpr (HPOOL) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { // row pairs (8i + 2 rsub, + 1) of the warp's 32, eight lanes (128 B) per pooled row
                    if (pixoff[i] >= 0) {
                        const float4 v0 = *reinterpret_cast<const float4*>(stg + (8 * i + 2 * rsub) * 36 + 4 * seg);
                        const float4 v1 = *reinterpret_cast<const float4*>(stg + (8 * i + 2 * rsub + 1) * 36 + 4 * seg);
                        const float4 o = make_float4(fmaxf(v0.x + bv.x, v1.x + bv.x), fmaxf(v0.y + bv.y, v1.y + bv.y),
                                                     fmaxf(v0.z + bv.z, v1.z + bv.z), fmaxf(v0.w + bv.w, v1.w + bv.w));
                        if (F16 && a.y_f16)           // the conv branch's output, already 11-bit operands deep: half the bytes to the pooling kernel
                            *(reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(a.y) + (imgbase + pixoff[i]) * NT + col0) + seg) =
                                make_uint2(rb_pack_h2(o.x, o.y), rb_pack_h2(o.z, o.w));
                        else
                            *(reinterpret_cast<float4*>(a.y + (imgbase + pixoff[i]) * NT + col0) + seg) = o;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) { // rows 4i .. 4i+3 of the warp's 32, eight lanes (128 B) per row
                    if (pixoff[i] >= 0) {
                        const float4 v = *reinterpret_cast<const float4*>(stg + (4 * i + rsub) * 36 + 4 * seg);
                        const float4 o = RES ? make_float4(v.x + bv.x + rr[i].x, v.y + bv.y + rr[i].y, v.z + bv.z + rr[i].z, v.w + bv.w + rr[i].w)
                                             : make_float4(v.x + bv.x, v.y + bv.y, v.z + bv.z, v.w + bv.w);
                        *(reinterpret_cast<float4*>(a.y + (imgbase + pixoff[i]) * NT + col0) + seg) = o;
                        if (F16 && a.ya)          // the next block's operand: its BN1 + ELU applied here, once per element, as fp16
                            *(reinterpret_cast<uint2*>(static_cast<uint16_t*>(a.ya) + (imgbase + pixoff[i]) * NT + col0) + seg) =
                                make_uint2(rb_pack_h2(rb_bn_elu(o.x, ysc.x, ysh.x), rb_bn_elu(o.y, ysc.y, ysh.y)),
                                           rb_pack_h2(rb_bn_elu(o.z, ysc.z, ysh.z), rb_bn_elu(o.w, ysc.w, ysh.w)));
                    }
                }
            }
            __syncwarp();                     // the staging tile is rewritten by the next unit
            if (u + kGroups < nunits) prefetch(u + kGroups);
        }
    }
    if (tid == 64) stamp(9);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if constexpr (PAIR) rb_cluster_sync();       // neither CTA leaves while the pair's MMAs / commits / relays could still target it
    else __syncthreads();
    if (tid == 0) stamp(10);
    if (warp == 0) {
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
    }
}

int rb_ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

long long* g_rb_stamps = nullptr;         // mmla_debug_resblock2d_stamps: 16 rows (launch ordinal) x 16 slots
int g_rb_stamp_cta = 0, g_rb_stamp_row = 0;

template <int NT, bool RES, int THREADS, bool STEM = false, bool HPOOL = false, bool PAIR = false, bool F16 = false>
int launch_rb(const RbArgs& s, long long images, size_t smem, cudaStream_t st) {
    static size_t attr[64] = {};                                  // per device: function attributes are per device
    int dev = 0;
    MMLA_CUDA_CHECK(cudaGetDevice(&dev));
    MMLA_REQUIRE(dev >= 0 && dev < 64, MMLA_EUNSUP, "resblock2d: device ordinal %d out of range", dev);
    auto kern = resblock2d_fused_kernel<NT, RES, THREADS, STEM, HPOOL, PAIR, F16>;
    if (smem > attr[dev]) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        attr[dev] = smem;
    }
    const long long ctas = images * s.cpi;
    if constexpr (PAIR) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(static_cast<unsigned>((ctas + 1) / 2 * 2), 1, 1);   // whole pairs; the odd one out runs the protocol only
        cfg.blockDim = dim3(THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        MMLA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, s));
    } else {
        kern<<<static_cast<unsigned>(ctas), THREADS, smem, st>>>(s);
    }
    mmla_count_launch(F16 ? (STEM ? "stem_resblock2d_f16_kernel" : "resblock2d_f16_kernel")
                          : (STEM ? "stem_resblock2d_fused_kernel" : "resblock2d_fused_kernel"), st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}

// one CTA per MMA (no pairs): residual / stem / row-pooled / plain variants at 256 or 512 threads
template <bool F16>
int dispatch_rb(const RbArgs& s, long long B, size_t smem, cudaStream_t st, int C, bool res, bool img, bool hpool, int thr) {
    if (res) {
        if (thr == 512) {
            switch (C) {
                case 32: return launch_rb<32, true, 512, false, false, false, F16>(s, B, smem, st);
                case 64: return launch_rb<64, true, 512, false, false, false, F16>(s, B, smem, st);
                default: return launch_rb<128, true, 512, false, false, false, F16>(s, B, smem, st);
            }
        }
        switch (C) {
            case 32: return launch_rb<32, true, 256, false, false, false, F16>(s, B, smem, st);
            case 64: return launch_rb<64, true, 256, false, false, false, F16>(s, B, smem, st);
            default: return launch_rb<128, true, 256, false, false, false, F16>(s, B, smem, st);
        }
    }
    if (img) return hpool ? launch_rb<32, false, 256, true, true, false, F16>(s, B, smem, st)
                          : launch_rb<32, false, 256, true, false, false, F16>(s, B, smem, st);
    if (hpool) {
        if (thr == 256) {
            switch (C) {
                case 32: return launch_rb<32, false, 256, false, true, false, F16>(s, B, smem, st);
                case 64: return launch_rb<64, false, 256, false, true, false, F16>(s, B, smem, st);
                default: return launch_rb<128, false, 256, false, true, false, F16>(s, B, smem, st);
            }
        }
        switch (C) {
            case 32: return launch_rb<32, false, 512, false, true, false, F16>(s, B, smem, st);
            case 64: return launch_rb<64, false, 512, false, true, false, F16>(s, B, smem, st);
            default: return launch_rb<128, false, 512, false, true, false, F16>(s, B, smem, st);
        }
    }
    if (thr == 256) {
        switch (C) {
            case 32: return launch_rb<32, false, 256, false, false, false, F16>(s, B, smem, st);
            case 64: return launch_rb<64, false, 256, false, false, false, F16>(s, B, smem, st);
            default: return launch_rb<128, false, 256, false, false, false, F16>(s, B, smem, st);
        }
    }
    switch (C) {
        case 32: return launch_rb<32, false, 512, false, false, false, F16>(s, B, smem, st);
        case 64: return launch_rb<64, false, 512, false, false, false, F16>(s, B, smem, st);
        default: return launch_rb<128, false, 512, false, false, false, F16>(s, B, smem, st);
    }
}

}  // namespace

// Whether resblock2d_fused_kernel takes a res_block's conv pair: 3x3 'same' (Cin 16 / 32 / 64 / 128 -> C) then (4, 1) 'same'
// (C -> C), both stride 1, ELU prologues.  MMLA_NET_FUSE_BLOCKS=0 keeps the two conv_slab launches.
bool mmla_resblock2d_eligible(int H, int W, int Cin, int C, int kh1, int kw1, int kh2, int kw2, int act) {
    const char* e = getenv("MMLA_NET_FUSE_BLOCKS");     // read per launch: tests flip it between calls
    if (e && e[0] == '0') return false;
    if (kh1 != 3 || kw1 != 3 || kh2 != 4 || kw2 != 1 || act != ACT_ELU) return false;
    if (Cin != 16 && Cin != 32 && Cin != 64 && Cin != 128) return false;
    if (C != 32 && C != 64 && C != 128) return false;
    if (H < 2 || W < 1) return false;
    if (static_cast<long long>(W + 2) * (H + 4) + 512 + 8 >= (1LL << 20)) return false;
    return true;
}

// Weight arrangement of the PAIR mode: per [32 x N] K-chunk the two column halves one after the other, each in conv_tc.cu's
// layout for N / 2 columns — out[kchunk][half][slab 0..7][n 0..N/2-1][4] = tf32(W[kchunk*32 + slab*4 + j][half*N/2 + n]) — so
// that a CTA's half of a chunk is one contiguous bulk copy.  Same size as mmla_tc_arrange_weights' output; K % 32 == 0.
void mmla_rb_arrange_weights_pair(const float* w, int K, int N, float* out) {
    const int nk = K / kRbBK, nh = N / 2;
    for (int kc = 0; kc < nk; ++kc)
        for (int h = 0; h < 2; ++h)
            for (int slab = 0; slab < 8; ++slab)
                for (int n = 0; n < nh; ++n)
                    for (int j = 0; j < 4; ++j) {
                        float v = w[static_cast<long long>(kc * kRbBK + slab * 4 + j) * N + h * nh + n];
                        uint32_t u;
                        memcpy(&u, &v, 4);
                        if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & ~0x1FFFu;
                        memcpy(&v, &u, 4);
                        out[((((static_cast<long long>(kc) * 2 + h) * 8 + slab) * nh + n) * 4) + j] = v;
                    }
}
// Weight arrangement of the fp16-operand mode: [64 x N] K-chunks, each [k octet 0..7][n 0..N-1][8 halves] =
// half(W[kchunk*64 + octet*8 + j][n]) (UMMA K-major no-swizzle core matrices, the fp16 twin of mmla_tc_arrange_weights'
// [k quad][n][4 tf32]); rows k >= K are zero.  Round-to-nearest-even, saturating at the largest finite half.  `out` holds
// ceil(K / 64) * 64 * N halves = as many BYTES per chunk as the TF32 arrangement.
long long mmla_rb_f16_arranged_halves(int K, int N) { return static_cast<long long>((K + kRbBKh - 1) / kRbBKh) * kRbBKh * N; }
static uint16_t rb_float_to_half_rn(float f) {
    if (f > 65504.f) f = 65504.f;
    if (f < -65504.f) f = -65504.f;
    const __half h = __float2half_rn(f);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
}
void mmla_rb_arrange_weights_f16(const float* w, int K, int N, uint16_t* out) {
    const int nk = (K + kRbBKh - 1) / kRbBKh;
    for (int kc = 0; kc < nk; ++kc)
        for (int oct = 0; oct < 8; ++oct)
            for (int n = 0; n < N; ++n)
                for (int j = 0; j < 8; ++j) {
                    const int k = kc * kRbBKh + oct * 8 + j;
                    out[((static_cast<long long>(kc) * 8 + oct) * N + n) * 8 + j] =
                        k < K ? rb_float_to_half_rn(w[static_cast<long long>(k) * N + n]) : static_cast<uint16_t>(0);
                }
}
bool mmla_rb_pair_wanted(int Cin, int C) {
    // Opt-in (MMLA_RB_PAIR=1): measured SLOWER than one CTA per MMA on every C >= 64 block of the overlap net (0.68 -> 0.78,
    // 0.29 -> 0.35, 0.56 -> 0.70, 0.29 -> 0.34 ms per 512 clips, profiles/r02/experiment_notes.txt): the N <= 128, K = 8 MMAs are
    // too small to amortise the pair's per-instruction hand-shake (162 vs 110 cycles per MMA at N = 64) and the three cluster
    // barriers keep the two CTAs in lock-step.  Kept as a tested alternative (bit-identical results).
    const char* e = getenv("MMLA_RB_PAIR");
    return e && e[0] == '1' && C >= 64 && (9 * Cin) % kRbBK == 0;
}

// resblock2d_persist.cu
int mmla_try_launch_resblock2d_persist(const float* x, float* y, long long B, int H, int W, int Cin, int C, const float* bn1_scale,
                                       const float* bn1_shift, const float* w1, const float* b1, const float* bn2_scale,
                                       const float* bn2_shift, const float* w2, const float* b2, const float* res,
                                       long long res_row_stride, cudaStream_t st, const void* img, int img_is_u8,
                                       const float* stem_w, const float* stem_b, int hpool, const void* w1_h, const void* w2_h, int y_f16,
                                       const void* xa, void* ya, const float* ya_scale, const float* ya_shift);

// img != null: stem mode — x is ignored, the block input is Conv2D(16, 1x1)(img) computed in the fill (Cin must be 16, no
// residual, 256-thread configuration).
int mmla_launch_resblock2d_fused(const float* x, float* y, long long B, int H, int W, int Cin, int C, const float* bn1_scale,
                                 const float* bn1_shift, const float* w1, const float* b1, const float* bn2_scale,
                                 const float* bn2_shift, const float* w2, const float* b2, const float* res,
                                 long long res_row_stride, cudaStream_t st, const void* img, int img_is_u8, const float* stem_w,
                                 const float* stem_b, int hpool, const float* w1_pair, const float* w2_pair, const void* w1_h,
                                 const void* w2_h, const void* xa, void* ya, const float* ya_scale, const float* ya_shift, int y_f16) {
    if (B <= 0) return MMLA_OK;
    // w1_h / w2_h: the same weights as fp16 chunks (mmla_rb_arrange_weights_f16); when given the block runs with fp16 operands
    const bool f16 = w1_h && w2_h;
    {   // the C = 32 blocks run on the persistent, warp-specialised kernel where their slabs fit (resblock2d_persist.cu; fp16
        // operands: the stem block only)
        const int pr = mmla_try_launch_resblock2d_persist(x, y, B, H, W, Cin, C, bn1_scale, bn1_shift, w1, b1, bn2_scale, bn2_shift, w2, b2,
                                                          res, res_row_stride, st, img, img_is_u8, stem_w, stem_b, hpool,
                                                          f16 ? w1_h : nullptr, f16 ? w2_h : nullptr, y_f16, f16 ? xa : nullptr,
                                                          f16 ? ya : nullptr, ya_scale, ya_shift);
        if (pr < 0) return -pr;
        if (pr > 0) return MMLA_OK;
    }
    if (f16) {
        w1 = static_cast<const float*>(w1_h);      // chunk c starts c * 128 * C bytes in, in either arrangement
        w2 = static_cast<const float*>(w2_h);
    }
    // w1_pair / w2_pair: the same weights in the PAIR arrangement (mmla_rb_arrange_weights_pair); when given (and wanted) the
    // block runs on CTA pairs
    const bool pair = !f16 && w1_pair && w2_pair && !img && mmla_rb_pair_wanted(Cin, C);
    if (pair) { w1 = w1_pair; w2 = w2_pair; }
    MMLA_REQUIRE(!hpool || (!res && H % 2 == 0), MMLA_EUNSUP, "resblock2d: row-pooled output needs an even height and no residual");
    MMLA_REQUIRE(!img || (Cin == 16 && C == 32 && !res && stem_w && stem_b), MMLA_EUNSUP, "resblock2d: stem mode needs Cin 16, C 32, no residual");
    MMLA_REQUIRE(!res || res_row_stride % 4 == 0, MMLA_EUNSUP, "resblock2d: residual row stride must be a multiple of 4 floats");
    RbArgs s;
    memset(&s, 0, sizeof(s));
    s.x = x; s.y = y; s.w1 = w1; s.w2 = w2; s.b1 = b1; s.b2 = b2;
    s.bn1_scale = bn1_scale; s.bn1_shift = bn1_shift; s.bn2_scale = bn2_scale; s.bn2_shift = bn2_shift;
    s.res = res; s.res_row_stride = res_row_stride;
    s.img = img; s.img_is_u8 = img_is_u8; s.stem_w = stem_w; s.stem_b = stem_b;
    MMLA_REQUIRE(f16 || (!xa && !ya), MMLA_EINVAL, "resblock2d: fp16 activations need the fp16-operand mode");
    MMLA_REQUIRE(!ya || (!hpool && ya_scale && ya_shift), MMLA_EINVAL, "resblock2d: fp16 output needs the next block's BN and a full-resolution output");
    MMLA_REQUIRE(!y_f16 || (f16 && hpool), MMLA_EINVAL, "resblock2d: an fp16 output exists for the row-pooled fp16-operand form only");
    s.xa = img ? nullptr : xa; s.ya = ya; s.ya_scale = ya_scale; s.ya_shift = ya_shift; s.y_f16 = y_f16;
    s.img_pixels = static_cast<long long>(H) * W;
    s.hpool = hpool;
    const int drop = hpool ? 4 : 3;                              // outputs a CTA gives up to the 4x1 halo (even for HPOOL)
    s.H = H; s.W = W; s.Fp = H + drop;
    s.fp_magic = static_cast<unsigned>((1ULL << 32) / static_cast<unsigned>(s.Fp)) + 1u;
    s.total_q = W * s.Fp;
    s.Cin = Cin; s.lq = rb_ilog2(Cin / 4);
    const int K1 = 9 * Cin, K2 = 4 * C;
    const int BK = f16 ? kRbBKh : kRbBK;                         // K per ring chunk (four MMAs), channels per 16-byte slab row
    const int cpr = f16 ? 8 : 4;
    s.nk1 = (K1 + BK - 1) / BK;
    s.nk = s.nk1 + K2 / BK;
    MMLA_REQUIRE(s.nk <= kRbMaxChunks, MMLA_EUNSUP, "resblock2d: K = %d + %d is too large", K1, K2);
    const size_t chunk = static_cast<size_t>(8) * C * 16 / (pair ? 2 : 1);   // one [32 x C] K-chunk of weights (PAIR: this CTA's half)
    constexpr size_t kBarBytes = 1024 + 128;                    // mbarriers + alignment slack
    const size_t par_bytes = static_cast<size_t>(3) * C * 4;
    auto rows_x = [&](int T) {
        int r = T * 128 + 2 * s.Fp + 2;
        if (f16) return r | 1;
        if (Cin == 16) { while ((r & 7) != 2) ++r; } else if ((r & 1) == 0) ++r;     // conflict-free fill stores (conv_slab.cu)
        return r;
    };
    auto rows_u = [&](int T) { return (T * 128 + 3) | 1; };
    auto slab_bytes = [&](int T, int nthr) {
        const size_t bx = static_cast<size_t>(Cin / cpr) * rows_x(T) * 16, bu = static_cast<size_t>(C / cpr) * rows_u(T) * 16;
        const size_t stg = static_cast<size_t>(nthr / 32) * 32 * 36 * 4;               // epilogue-2 staging tiles, one per warp
        size_t b = bx > bu ? bx : bu;
        return b > stg ? b : stg;
    };
    // Tiles per CTA, ring depth and CTAs per SM: the cost model of conv_slab.cu (fitted to its clock64 timelines) with both
    // convolutions' phases in one CTA; co-resident CTAs in different phases are what hides the fill and the epilogues.
    int force_t = 0, force_kb = 0, force_stages = 0;
    if (const char* e = getenv("MMLA_RB_TILES")) force_t = atoi(e);
    if (const char* e = getenv("MMLA_RB_KB")) force_kb = atoi(e);
    if (const char* e = getenv("MMLA_RB_STAGES")) force_stages = atoi(e);
    // Measured on B200 for the overlap classifier's own shapes (scripts/sweep_resblock2d.sh, profiles/r02/sweep_resblock2d_v2.txt)
    // where the model below picks a slower configuration: the full-resolution block prefers four tiles at three CTAs per SM
    // even with a two-slot ring (0.89 vs 1.02 ms per 512 clips), the 16 x 19 blocks prefer two tiles on one CTA per SM (half
    // the weight stream from L2 per output: 0.30 vs 0.35 ms).
    // fp16 operands (profiles/r02/sweep_resblock2d_f16_v2.txt): the slabs are half as large, so the full-resolution block keeps
    // three CTAs per SM at 56 KB (0.74 vs 0.76 ms) and the 16 x 19 blocks run two CTAs per SM (0.143 / 0.135 vs 0.182 / 0.172 ms).
    if (!force_t && !force_kb) {
        if (H == 128 && Cin == 16 && C == 32) { force_t = 4; force_kb = f16 ? 56 : 75; }
        if (H == 16 && Cin == 128 && C == 128) { force_t = 2; force_kb = f16 ? 113 : 226; }
    }
    int tmax = kRbMaxTiles;
    if (tmax > 512 / C) tmax = 512 / C;
    {
        const int need = (s.total_q + 3 + 127) / 128;             // a whole image in one CTA
        if (tmax > need) tmax = need;
    }
    int T = 0, best_thr = 256;
    double best = 0.0;
    for (int pass = 0; pass < 2 && !T; ++pass) {       // pass 1: a forced tile count / budget that fits nowhere is ignored
        if (pass == 1) force_t = force_kb = 0;
        const int budgets_kb[3] = {75, 113, 226};
        const double overlap[3] = {4.5, 3.2, 1.0};
        for (int bi = res ? 1 : 0; bi < 3; ++bi) {
            // (a residual's prefetch holds 40 more registers: 256 threads at two CTAs per SM, 512 when the CTA has the SM alone)
            const int nthr = img || bi == 0 || (res && bi == 1) ? 256 : 512;
            const int ctas = 3 - bi;
            if (force_kb >= 16 && force_kb <= 226 && force_kb > budgets_kb[bi]) continue;   // a forced budget that large means fewer CTAs per SM
            const size_t budget = static_cast<size_t>(force_kb >= 16 && force_kb <= 226 ? force_kb : budgets_kb[bi]) * 1024;
            for (int t = 1; t <= tmax; ++t) {
                if (force_t >= 1 && force_t <= tmax && t != force_t) continue;
                int cols = 32;
                while (cols < t * C) cols *= 2;
                if (cols * ctas > 512) continue;                  // the CTAs of an SM share 512 TMEM columns
                const size_t fixed = slab_bytes(t, nthr) + par_bytes + kBarBytes + 256;
                if (fixed + 2 * chunk > budget) continue;
                int stages = static_cast<int>((budget - fixed) / chunk);
                if (stages > s.nk) stages = s.nk;
                if (stages > kRbMaxStages) stages = kRbMaxStages;
                if (force_stages >= 1 && force_stages <= stages) stages = force_stages;
                const double per_chunk = t * 4.0 * (C / 2 > 45 ? C / 2 : 45);
                const double refill = (3000.0 + per_chunk) / stages;
                const double mma = s.nk * (per_chunk > refill ? per_chunk : refill);
                const double fill = 5000.0 + rows_x(t) * (Cin / 4) / static_cast<double>(nthr) * 40.0;
                const double epi = 2.0 * (3000.0 + 700.0 * t * (C / 32));
                const int outs = t * 128 - drop;
                const int cpi = (s.total_q + outs - 1) / outs;
                const double cost = (fill + mma + epi) * cpi / overlap[bi] / s.total_q;
                if (!T || cost < best) {
                    T = t; best = cost; s.stages = stages; best_thr = nthr;
                }
            }
        }
    }
    MMLA_REQUIRE(T > 0, MMLA_EUNSUP, "resblock2d: block does not fit in shared memory (Cin %d, C %d, H %d)", Cin, C, H);
    s.T = T;
    s.S = T * 128 - drop;
    s.cpi = (s.total_q + s.S - 1) / s.S;
    s.RsX = rows_x(T);
    s.RsU = rows_u(T);
    s.nmma1_last = s.nmma2_last = 0;
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
    const size_t sb = slab_bytes(T, best_thr);
    s.ring_off = static_cast<unsigned>((sb + 127) / 128 * 128);
    s.par_off = s.ring_off + static_cast<unsigned>(s.stages * chunk);
    s.bar_off = static_cast<unsigned>((s.par_off + par_bytes + 127) / 128 * 128);
    const size_t smem = s.bar_off + kBarBytes;
    if (getenv("MMLA_RB_VERBOSE"))
        fprintf(stderr, "resblock2d: %dx%d Cin %d C %d: %d outputs/image, T %d, %d CTAs/image, ring %d of %d chunks, %zu KB smem, %d threads%s\n",
                H, W, Cin, C, s.total_q, s.T, s.cpi, s.stages, s.nk, smem / 1024, best_thr, pair ? ", CTA pairs" : f16 ? ", fp16 operands" : "");
    if (g_rb_stamps && g_rb_stamp_row < 16) {
        s.stamps = g_rb_stamps + 16 * g_rb_stamp_row++;
        s.stamp_cta = static_cast<int>((static_cast<long long>(g_rb_stamp_cta) % B) * s.cpi + s.cpi / 2);   // a mid-image CTA
    }
    MMLA_REQUIRE(B * s.cpi < (1LL << 31) - 2 && B * s.img_pixels * (Cin > C ? Cin : C) < (1LL << 40), MMLA_EUNSUP,
                 "resblock2d: batch too large");
    s.n_ctas = static_cast<int>(B * s.cpi);
    if (pair) {
        if (res) return C == 64 ? launch_rb<64, true, 256, false, false, true>(s, B, smem, st)
                                : launch_rb<128, true, 256, false, false, true>(s, B, smem, st);
        if (hpool) {
            if (best_thr == 256)
                return C == 64 ? launch_rb<64, false, 256, false, true, true>(s, B, smem, st)
                               : launch_rb<128, false, 256, false, true, true>(s, B, smem, st);
            return C == 64 ? launch_rb<64, false, 512, false, true, true>(s, B, smem, st)
                           : launch_rb<128, false, 512, false, true, true>(s, B, smem, st);
        }
        if (best_thr == 256)
            return C == 64 ? launch_rb<64, false, 256, false, false, true>(s, B, smem, st)
                           : launch_rb<128, false, 256, false, false, true>(s, B, smem, st);
        return C == 64 ? launch_rb<64, false, 512, false, false, true>(s, B, smem, st)
                       : launch_rb<128, false, 512, false, false, true>(s, B, smem, st);
    }
    return f16 ? dispatch_rb<true>(s, B, smem, st, C, res != nullptr, img != nullptr, hpool != 0, best_thr)
               : dispatch_rb<false>(s, B, smem, st, C, res != nullptr, img != nullptr, hpool != 0, best_thr);
}

extern "C" __attribute__((visibility("default"))) void mmla_debug_resblock2d_stamps(long long* dev_stamps, int32_t image) {
    g_rb_stamps = dev_stamps;
    g_rb_stamp_cta = image;
    g_rb_stamp_row = 0;
}
