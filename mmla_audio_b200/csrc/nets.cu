// Classifier forward passes for sm_100a (fp32 path, bit-faithful layer semantics):
//   overlap net  — Conv2D 1x1 -> 9 pre-activation residual blocks (BN-ELU-Conv3x3-BN-ELU-Conv(4,1),
//                  three with stride-2 1x1 shortcut + MaxPool2x2 'same') -> mean over mel axis ->
//                  BiLSTM(256) -> LeakyReLU(0.3) -> Dense(2) softmax
//                  (OverlapDetection/scripts/overlap_detector_temp.py:253-303)
//   speaker net  — Conv1D k4 -> 9 res units (MaxPool1D + stride-2 1x1 shortcut on three) ->
//                  BN-ReLU-AvgPool1D(4) -> BiLSTM(256) -> Dense(n) softmax | sigmoid
//                  (SpeakerIdentification/scripts/speaker_identification.py:168-218,401-410)
// replacing model.predict(x) at OverlapDetection/scripts/record_on_pc.py:159 and
// SpeakerIdentification/scripts/record_on_pc.py:136.
//
// Every convolution / dense projection / LSTM matmul runs through ONE implicit-GEMM kernel
// (NHWC activations, TF HWIO weights = row-major [K][N]) whose A-operand gather applies the
// preceding BatchNorm + activation on the fly and whose epilogue adds bias and the residual.
// Keras 'same' padding (extra element at the end) is handled by explicit pad_top/pad_left.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"
#include <cuda_fp16.h>

#include "conv_common.cuh"

namespace {

// (ConvArgs, activation ids and apply_act live in conv_common.cuh)

template <int TN>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const ConvArgs a) {
    constexpr int BN = 16 * TN;
    constexpr int B_LOADS = (kBK * BN / 4 + 255) / 256;          // float4 per thread for the W tile
    __shared__ __align__(16) float As[2][kBK][kBM];
    __shared__ __align__(16) float Bs[2][kBK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = static_cast<long long>(blockIdx.x) * kBM;
    const int n0 = blockIdx.y * BN;

    // A-gather bookkeeping: this thread always fills row `arow` of the tile
    const int arow = tid & 127, ahalf = tid >> 7;
    const long long am = m0 + arow;
    const bool avalid = am < a.M;
    int hi0 = 0, wi0 = 0;
    long long xbase = 0;
    if (avalid) {
        const int hw = a.Ho * a.Wo;
        const long long b = am / hw;
        const int r = static_cast<int>(am - b * hw);
        const int ho = r / a.Wo, wo = r - ho * a.Wo;
        hi0 = ho * a.stride - a.pad_t;
        wi0 = wo * a.stride - a.pad_l;
        xbase = b * a.H * a.W * a.Cin;
    }
    const bool fast = (a.Cin % kBK == 0) && !a.x_is_u8;
    const float* xf = static_cast<const float*>(a.x);
    const unsigned char* xu = static_cast<const unsigned char*>(a.x);

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float ra[8];
    float4 rb[B_LOADS];

    auto load_tiles = [&](int k0) {
        if (fast) {
            const int tap = k0 / a.Cin, c0 = k0 - tap * a.Cin;
            const int ki = tap / a.kw, kj = tap - ki * a.kw;
            const int hi = hi0 + ki, wi = wi0 + kj;
            const bool inb = avalid && hi >= 0 && hi < a.H && wi >= 0 && wi < a.W;
            const float* src = xf + xbase + (static_cast<long long>(hi) * a.W + wi) * a.Cin + c0 + ahalf * 8;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (inb) {
                    v = *reinterpret_cast<const float4*>(src + 4 * q);
                    if (a.pre_scale) {
                        const int c = c0 + ahalf * 8 + 4 * q;
                        const float4 sc = *reinterpret_cast<const float4*>(a.pre_scale + c);
                        const float4 sh = *reinterpret_cast<const float4*>(a.pre_shift + c);
                        v.x = apply_act(fmaf(v.x, sc.x, sh.x), a.pre_act);
                        v.y = apply_act(fmaf(v.y, sc.y, sh.y), a.pre_act);
                        v.z = apply_act(fmaf(v.z, sc.z, sh.z), a.pre_act);
                        v.w = apply_act(fmaf(v.w, sc.w, sh.w), a.pre_act);
                    }
                }
                ra[4 * q + 0] = v.x; ra[4 * q + 1] = v.y; ra[4 * q + 2] = v.z; ra[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = k0 + ahalf * 8 + e;
                float v = 0.f;
                if (avalid && k < a.K) {
                    const int tap = k / a.Cin, c = k - tap * a.Cin;
                    const int ki = tap / a.kw, kj = tap - ki * a.kw;
                    const int hi = hi0 + ki, wi = wi0 + kj;
                    if (hi >= 0 && hi < a.H && wi >= 0 && wi < a.W) {
                        const long long idx = xbase + (static_cast<long long>(hi) * a.W + wi) * a.Cin + c;
                        v = a.x_is_u8 ? static_cast<float>(xu[idx]) : xf[idx];
                        if (a.pre_scale) v = apply_act(fmaf(v, a.pre_scale[c], a.pre_shift[c]), a.pre_act);
                    }
                }
                ra[e] = v;
            }
        }
#pragma unroll
        for (int l = 0; l < B_LOADS; ++l) {
            const int idx = tid + 256 * l;                          // float4 index in [kBK][BN/4]
            const int kr = idx / (BN / 4), c4 = idx - kr * (BN / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kr < kBK) {
                const int k = k0 + kr, n = n0 + 4 * c4;
                if (k < a.K) {
                    const float* src = a.w + static_cast<long long>(k) * a.N + n;
                    if (n + 3 < a.N && (a.N & 3) == 0) {
                        v = *reinterpret_cast<const float4*>(src);
                    } else {
                        if (n + 0 < a.N) v.x = src[0];
                        if (n + 1 < a.N) v.y = src[1];
                        if (n + 2 < a.N) v.z = src[2];
                        if (n + 3 < a.N) v.w = src[3];
                    }
                }
            }
            rb[l] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 8; ++e) As[buf][ahalf * 8 + e][arow] = ra[e];
#pragma unroll
        for (int l = 0; l < B_LOADS; ++l) {
            const int idx = tid + 256 * l;
            const int kr = idx / (BN / 4), c4 = idx - kr * (BN / 4);
            if (kr < kBK) *reinterpret_cast<float4*>(&Bs[buf][kr][4 * c4]) = rb[l];
        }
    };

    const int nk = (a.K + kBK - 1) / kBK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kc = 0; kc < nk; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nk) load_tiles((kc + 1) * kBK);
#pragma unroll
        for (int k = 0; k < kBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = Bs[buf][k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kc + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + ty * 8 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n < a.N) {
                float v = acc[i][j] + a.bias[n];
                if (a.res) v += a.res[m * a.res_row_stride + n];
                a.y[m * a.N + n] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// small layer kernels
// ---------------------------------------------------------------------------------------------
// MaxPool 'same' with window (ph, pw), stride = window, NHWC; out-of-range taps are -inf.
// One thread per (output pixel, channel quad): 128-bit loads/stores, 32-bit index math.
__global__ void __launch_bounds__(256) maxpool_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H,
                                                      int W, int C, int ph, int pw, int Ho, int Wo) {
    const int C4 = C >> 2;
    const int total = B * Ho * Wo * C4;                              // < 2^31 (launcher checks)
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int c4 = e % C4;
        int r = e / C4;
        const int wo = r % Wo;
        r /= Wo;
        const int ho = r % Ho;
        const int b = r / Ho;
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int i = 0; i < ph; ++i)
            for (int j = 0; j < pw; ++j) {
                const int hi = ho * ph + i, wi = wo * pw + j;
                if (hi < H && wi < W) {
                    const float4 v = *reinterpret_cast<const float4*>(x + (static_cast<long long>((b * H + hi) * W + wi) * C + 4 * c4));
                    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
                }
            }
        *reinterpret_cast<float4*>(y + static_cast<long long>(e) * 4) = m;
    }
}

// speaker stem input [rows][39] -> [rows][40] (zero column) so the stem conv takes the vectorised
// Cin % 4 == 0 gather path of conv_tc_kernel
__global__ void __launch_bounds__(256) pad_channels_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows,
                                                           int cin, int cpad) {
    const int quads = cpad >> 2;
    const long long total = rows * quads;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long r = e / quads;
        const int c = 4 * static_cast<int>(e - r * quads);
        const float* src = x + r * cin + c;
        float4 v;
        v.x = c + 0 < cin ? src[0] : 0.f;
        v.y = c + 1 < cin ? src[1] : 0.f;
        v.z = c + 2 < cin ? src[2] : 0.f;
        v.w = c + 3 < cin ? src[3] : 0.f;
        *reinterpret_cast<float4*>(y + e * 4) = v;
    }
}

// overlap: Lambda(mean over axis 1 = H): [B,H,W,C] -> [B,W,C]
// Tail of a pooled res_block of the overlap net (overlap_detector_temp.py:262-270) in one pass:
//   y = MaxPool2x2/2 'same'(z) + Conv2D(N, 1x1, stride 2)(x) + bias
// fp16-operand mode: the expressions of resblock2d_common.cuh's rb_bn_elu / rb_pack_h2 (BN + ELU in fp32, saturating half2 pack)
__device__ __forceinline__ float act_bn_elu(float v, float sc, float sh) {
    v = fmaf(v, sc, sh);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v, 0.f) * 1.4426950408889634f));
    return v > 0.f ? v : e - 1.f;
}
__device__ __forceinline__ uint32_t act_pack_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // F2FP.SATFINITE: +-65504 instead of inf
    return r;
}
// z = the block's second conv at full resolution [B,H,W,N], x = the block input [B,H,W,Cin].  One thread owns four output
// channels of one pooled pixel; the [Cin][N] shortcut weights sit in shared memory; the products are exact fp32 FMAs
// (K = Cin <= 64 is too short for the tensor-core path to pay: the separate maxpool_kernel + im2col conv_tc_kernel pair
// took 0.88 ms per 512 clips, this is bound by reading z once).
// img != null (first block, stem folded into the conv-pair kernel): x is never materialised — the block input at the pooled
// pixel is the stem Conv2D(16, 1x1) of the 3-channel image, recomputed here with stem1x1_kernel's expression (Cin = 16).
#ifndef MMLA_POOL_MIN_CTAS
#define MMLA_POOL_MIN_CTAS 3
#endif
template <bool Z16>
__global__ void __launch_bounds__(256, MMLA_POOL_MIN_CTAS) pool_shortcut_kernel(const float* __restrict__ x, const float* __restrict__ z,
                                                            const float* __restrict__ ws, const float* __restrict__ bs,
                                                            float* __restrict__ y, long long B, int H, int W, int Cin, int N,
                                                            const void* __restrict__ img, int img_is_u8,
                                                            const float* __restrict__ stem_w, const float* __restrict__ stem_b,
                                                            int z_half, unsigned short* __restrict__ ya = nullptr,
                                                            const float* __restrict__ ya_scale = nullptr,
                                                            const float* __restrict__ ya_shift = nullptr) {
    // ya != null (fp16-operand mode): also writes ELU(BN(y)) as fp16 [B,Ho,Wo,N] with the NEXT block's BN1 (ya_scale / ya_shift),
    // i.e. the next conv-pair kernel's operand, so that its fill is a plain asynchronous copy (resblock2d_fused.cu)
    extern __shared__ float4 wsm4[];                  // [Cin][N / 4] (+ stem mode: [4][16] = w0 | w1 | w2 | bias)
    for (int i = threadIdx.x; i < Cin * N / 4; i += blockDim.x) wsm4[i] = reinterpret_cast<const float4*>(ws)[i];
    float* stem_s = reinterpret_cast<float*>(wsm4 + Cin * N / 4);
    if (img)
        for (int i = threadIdx.x; i < 64; i += blockDim.x) stem_s[i] = i < 48 ? stem_w[i] : stem_b[i - 48];
    __syncthreads();
    // One thread owns four output channels of kPix consecutive pooled pixels of a row: every shortcut weight read from shared
    // memory feeds kPix FMAs (with one pixel per thread the kernel was bound by its LDS.128 stream, one per four FMAs:
    // ~0.14 ms of each launch's 0.22-0.37 ms per 512 clips).  The FMA order per output is unchanged (c ascending).
    constexpr int kPix = 4;
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, Q = N / 4, G = (Wo + kPix - 1) / kPix;
    const long long total = B * Ho * G * Q;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int q = static_cast<int>(idx % Q);
        const long long pg = idx / Q;
        const int g = static_cast<int>(pg % G);
        const long long t = pg / G;
        const int ho = static_cast<int>(t % Ho);
        const long long b = t / Ho;
        const int h0 = 2 * ho;
        float4 m[kPix], acc[kPix];
        float c0[kPix], c1[kPix], c2[kPix];
        const float* xb[kPix];
        const float4 bias = *reinterpret_cast<const float4*>(bs + 4 * q);
#pragma unroll
        for (int u = 0; u < kPix; ++u) {
            const int wo = min(kPix * g + u, Wo - 1);  // the tail group recomputes the row's last pixel (not stored)
            const int w0 = 2 * wo;
            const long long pin = (b * H + h0) * W + w0;
            // z_half: z = [B, H/2, W, N] already holds the maximum over each window's two rows (resblock2d_fused.cu, HPOOL);
            // Z16: ... as fp16 (fp16-operand mode: the conv branch's output in half the bytes)
            const long long zoff = (z_half ? ((b * Ho + ho) * W + w0) : pin) * N + 4 * q;
            auto ldz = [&](long long d) {               // four channels at element offset d from this pixel
                if constexpr (Z16) {
                    const uint2 r = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(z) + zoff + d);
                    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
                    return make_float4(lo.x, lo.y, hi.x, hi.y);
                } else {
                    return *reinterpret_cast<const float4*>(z + zoff + d);
                }
            };
            m[u] = ldz(0);
            auto mx = [&](long long d) {
                const float4 v = ldz(d);
                m[u].x = fmaxf(m[u].x, v.x); m[u].y = fmaxf(m[u].y, v.y); m[u].z = fmaxf(m[u].z, v.z); m[u].w = fmaxf(m[u].w, v.w);
            };
            if (w0 + 1 < W) mx(N);                     // 'same' pooling: the missing right / bottom neighbours are -inf
            if (!z_half && h0 + 1 < H) {
                mx(static_cast<long long>(W) * N);
                if (w0 + 1 < W) mx(static_cast<long long>(W) * N + N);
            }
            xb[u] = x + pin * Cin;
            acc[u] = bias;
            c0[u] = c1[u] = c2[u] = 0.f;
            if (img) {
                if (img_is_u8) {
                    const unsigned char* px = static_cast<const unsigned char*>(img) + pin * 3;
                    c0[u] = static_cast<float>(px[0]); c1[u] = static_cast<float>(px[1]); c2[u] = static_cast<float>(px[2]);
                } else {
                    const float* px = static_cast<const float*>(img) + pin * 3;
                    c0[u] = px[0]; c1[u] = px[1]; c2[u] = px[2];
                }
            }
        }
        for (int c = 0; c < Cin; c += 4) {
            const float4 w0v = wsm4[(c + 0) * Q + q], w1v = wsm4[(c + 1) * Q + q], w2v = wsm4[(c + 2) * Q + q], w3v = wsm4[(c + 3) * Q + q];
#pragma unroll
            for (int u = 0; u < kPix; ++u) {
                float4 xv;
                if (img) {
                    xv.x = fmaf(c2[u], stem_s[32 + c], fmaf(c1[u], stem_s[16 + c], fmaf(c0[u], stem_s[c], stem_s[48 + c])));
                    xv.y = fmaf(c2[u], stem_s[33 + c], fmaf(c1[u], stem_s[17 + c], fmaf(c0[u], stem_s[c + 1], stem_s[49 + c])));
                    xv.z = fmaf(c2[u], stem_s[34 + c], fmaf(c1[u], stem_s[18 + c], fmaf(c0[u], stem_s[c + 2], stem_s[50 + c])));
                    xv.w = fmaf(c2[u], stem_s[35 + c], fmaf(c1[u], stem_s[19 + c], fmaf(c0[u], stem_s[c + 3], stem_s[51 + c])));
                } else {
                    xv = *reinterpret_cast<const float4*>(xb[u] + c);
                }
                float4& a = acc[u];
                a.x = fmaf(xv.x, w0v.x, a.x); a.y = fmaf(xv.x, w0v.y, a.y); a.z = fmaf(xv.x, w0v.z, a.z); a.w = fmaf(xv.x, w0v.w, a.w);
                a.x = fmaf(xv.y, w1v.x, a.x); a.y = fmaf(xv.y, w1v.y, a.y); a.z = fmaf(xv.y, w1v.z, a.z); a.w = fmaf(xv.y, w1v.w, a.w);
                a.x = fmaf(xv.z, w2v.x, a.x); a.y = fmaf(xv.z, w2v.y, a.y); a.z = fmaf(xv.z, w2v.z, a.z); a.w = fmaf(xv.z, w2v.w, a.w);
                a.x = fmaf(xv.w, w3v.x, a.x); a.y = fmaf(xv.w, w3v.y, a.y); a.z = fmaf(xv.w, w3v.z, a.z); a.w = fmaf(xv.w, w3v.w, a.w);
            }
        }
        float4 ysc = make_float4(0.f, 0.f, 0.f, 0.f), ysh = ysc;
        if (ya) {
            ysc = *reinterpret_cast<const float4*>(ya_scale + 4 * q);
            ysh = *reinterpret_cast<const float4*>(ya_shift + 4 * q);
        }
#pragma unroll
        for (int u = 0; u < kPix; ++u) {
            const int wo = kPix * g + u;
            if (wo < Wo) {
                const float4 o = make_float4(m[u].x + acc[u].x, m[u].y + acc[u].y, m[u].z + acc[u].z, m[u].w + acc[u].w);
                *reinterpret_cast<float4*>(y + ((b * Ho + ho) * Wo + wo) * N + 4 * q) = o;
                if (ya)
                    *reinterpret_cast<uint2*>(ya + ((b * Ho + ho) * Wo + wo) * N + 4 * q) =
                        make_uint2(act_pack_h2(act_bn_elu(o.x, ysc.x, ysh.x), act_bn_elu(o.y, ysc.y, ysh.y)),
                                   act_pack_h2(act_bn_elu(o.z, ysc.z, ysh.z), act_bn_elu(o.w, ysc.w, ysh.w)));
            }
        }
    }
}

// Overlap-net stem Conv2D(16, 1x1) on the 3-channel classifier input (overlap_detector_temp.py:283): a per-pixel
// 3 -> 16 affine map, i.e. a streaming kernel (3 bytes in, 64 bytes out per pixel).  Four lanes share a pixel and
// write one float4 each, so a warp store covers 512 contiguous bytes.  (Through the im2col tensor-core kernel the
// layer padded K = 3 to 32 and took 0.81 ms per 512 clips; this is bound by the 633 MB it writes.)
__global__ void __launch_bounds__(256) stem1x1_kernel(const void* __restrict__ x, int x_is_u8, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ y, long long pixels) {
    const int quad = threadIdx.x & 3;
    const float4 w0 = *reinterpret_cast<const float4*>(w + 4 * quad);          // w[k][16], k = 0..2
    const float4 w1 = *reinterpret_cast<const float4*>(w + 16 + 4 * quad);
    const float4 w2 = *reinterpret_cast<const float4*>(w + 32 + 4 * quad);
    const float4 b = *reinterpret_cast<const float4*>(bias + 4 * quad);
    const long long stride = static_cast<long long>(gridDim.x) * (blockDim.x >> 2);
    for (long long p = static_cast<long long>(blockIdx.x) * (blockDim.x >> 2) + (threadIdx.x >> 2); p < pixels; p += stride) {
        float c0, c1, c2;
        if (x_is_u8) {
            const unsigned char* px = static_cast<const unsigned char*>(x) + 3 * p;
            c0 = static_cast<float>(px[0]); c1 = static_cast<float>(px[1]); c2 = static_cast<float>(px[2]);
        } else {
            const float* px = static_cast<const float*>(x) + 3 * p;
            c0 = px[0]; c1 = px[1]; c2 = px[2];
        }
        float4 o;
        o.x = fmaf(c2, w2.x, fmaf(c1, w1.x, fmaf(c0, w0.x, b.x)));
        o.y = fmaf(c2, w2.y, fmaf(c1, w1.y, fmaf(c0, w0.y, b.y)));
        o.z = fmaf(c2, w2.z, fmaf(c1, w1.z, fmaf(c0, w0.z, b.z)));
        o.w = fmaf(c2, w2.w, fmaf(c1, w1.w, fmaf(c0, w0.w, b.w)));
        *reinterpret_cast<float4*>(y + p * 16 + 4 * quad) = o;
    }
}

__global__ void __launch_bounds__(256) mean_h_kernel(const float* __restrict__ x, float* __restrict__ y, long long B,
                                                     int H, int W, int C) {
    const long long total = B * W * C;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int c = static_cast<int>(e % C);
        const long long r = e / C;
        const int w = static_cast<int>(r % W);
        const long long b = r / W;
        float s = 0.f;
        for (int h = 0; h < H; ++h) s += x[((b * H + h) * W + w) * C + c];
        y[e] = s / static_cast<float>(H);
    }
}

// speaker: BN -> ReLU -> AveragePooling1D(4): [B,T,C] -> [B,T/4,C]
__global__ void __launch_bounds__(256) bn_relu_avgpool4_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               const float* __restrict__ sc, const float* __restrict__ sh,
                                                               long long B, int T, int C) {
    const int To = T / 4;
    const long long total = B * To * C;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int c = static_cast<int>(e % C);
        const long long r = e / C;
        const int to = static_cast<int>(r % To);
        const long long b = r / To;
        float s = 0.f;
        for (int i = 0; i < 4; ++i) s += fmaxf(fmaf(x[((b * T + 4 * to + i)) * C + c], sc[c], sh[c]), 0.f);
        y[e] = s * 0.25f;
    }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// LSTM pointwise step: z [B,4u] (i,f,c,o pre-activations) -> c, h  (Keras gate order)
__global__ void __launch_bounds__(256) lstm_gates_kernel(const float* __restrict__ z, long long z_row_stride,
                                                         float* __restrict__ c, float* __restrict__ h, long long B,
                                                         int u, int first_step) {
    const long long total = B * u;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long b = e / u;
        const int j = static_cast<int>(e - b * u);
        const float* zr = z + b * z_row_stride;
        const float ig = sigmoidf_(zr[j]);
        const float fg = sigmoidf_(zr[u + j]);
        const float gg = tanhf(zr[2 * u + j]);
        const float og = sigmoidf_(zr[3 * u + j]);
        const float cprev = first_step ? 0.f : c[e];
        const float cn = fg * cprev + ig * gg;
        c[e] = cn;
        h[e] = og * tanhf(cn);
    }
}

// [B,256] forward h | [B,256] backward h -> [B,512], the Bidirectional layer's concatenated output
__global__ void __launch_bounds__(256) embed_concat_kernel(const float* __restrict__ hf, const float* __restrict__ hb, long long B,
                                                           float* __restrict__ out) {
    const long long total = B * 512;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = e >> 9;
        const int i = static_cast<int>(e & 511);
        out[e] = i < 256 ? hf[b * 256 + i] : hb[b * 256 + i - 256];
    }
}

// head: z [B,512] (fwd|bwd) -> optional LeakyReLU -> Dense -> softmax|sigmoid -> prob, argmax
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ hf, const float* __restrict__ hb,
                                                   const float* __restrict__ wk, const float* __restrict__ wb,
                                                   float leaky, int use_leaky, int n_classes, int sigmoid_head,
                                                   long long B, float* __restrict__ prob, int* __restrict__ labels) {
    extern __shared__ float zs[];                                   // [warps][512]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* z = zs + warp * 512;
    for (long long b = static_cast<long long>(blockIdx.x) * nw + warp; b < B; b += static_cast<long long>(gridDim.x) * nw) {
        for (int i = lane; i < 512; i += 32) {
            float v = i < 256 ? hf[b * 256 + i] : hb[b * 256 + i - 256];
            if (use_leaky) v = v > 0.f ? v : leaky * v;
            z[i] = v;
        }
        __syncwarp();
        float* pr = prob + b * n_classes;
        float vmax = -INFINITY;
        if (n_classes <= 32) {
            // few classes (the transferred speaker heads, the 2-way overlap head): the 32 lanes split K instead of
            // 10 lanes walking 512 dependent FMAs each; lane n ends up with logit n
            float zr[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) zr[k] = z[lane + 32 * k];
            float mine = 0.f;
            for (int n = 0; n < n_classes; ++n) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 16; ++k) acc = fmaf(zr[k], __ldg(wk + static_cast<long long>(lane + 32 * k) * n_classes + n), acc);
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == n) mine = acc + wb[n];
            }
            if (lane < n_classes) {
                pr[lane] = mine;                                    // logits for now
                vmax = mine;
            }
        } else
        for (int n = lane; n < n_classes; n += 32) {
            float acc = 0.f;
            for (int i = 0; i < 512; ++i) acc = fmaf(z[i], wk[static_cast<long long>(i) * n_classes + n], acc);
            acc += wb[n];
            pr[n] = acc;                                            // logits for now
            vmax = fmaxf(vmax, acc);
        }
        for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        float denom = 0.f;
        if (!sigmoid_head) {
            for (int n = lane; n < n_classes; n += 32) denom += expf(pr[n] - vmax);
            for (int o = 16; o > 0; o >>= 1) denom += __shfl_xor_sync(0xffffffffu, denom, o);
        }
        float best = -INFINITY;
        int besti = 0x7fffffff;
        for (int n = lane; n < n_classes; n += 32) {
            const float l = pr[n];
            const float pv = sigmoid_head ? sigmoidf_(l) : expf(l - vmax) / denom;
            pr[n] = pv;
            if (pv > best) { best = pv; besti = n; }               // first maximum within the lane
        }
        for (int o = 16; o > 0; o >>= 1) {                          // np.argmax: first max wins
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (labels && lane == 0) labels[b] = besti;
        __syncwarp();
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// network description + executor
// ---------------------------------------------------------------------------------------------
constexpr long long kOvHalfFloats = 64LL * 76 * 32 / 2;   // floats holding one fp16 [64,76,32] activation (the largest conv-pair operand after block 1)
struct ConvW {
    const float* k = nullptr;
    const float* k_tc = nullptr;          // same weights arranged + TF32-rounded for conv_tc.cu (or null)
    const float* k_tc2 = nullptr;         // ... in the CTA-pair arrangement of resblock2d_fused.cu (overlap net, cout >= 64)
    const float* k_h = nullptr;           // ... as fp16 chunks for resblock2d_fused.cu's fp16-operand mode (overlap net's conv pairs)
    const float* b = nullptr;
    int kh = 1, kw = 1, cin = 0, cout = 0, stride = 1;
};
struct BnW {
    const float* scale = nullptr;
    const float* shift = nullptr;
};
struct BlockW {
    typedef const BnW* BnPtr;
    BnW bn1, bn2;
    ConvW conv1, conv2, shortcut;
    bool pool = false;
};
struct MmlaNet {
    int kind = 0, n_classes = 0, head = 0;
    int precision = MMLA_PRECISION_FP32;  // MMLA_PRECISION_*
    float* dev = nullptr;                 // device blob (weights + folded BN)
    ConvW stem;
    ConvW stem_pad;                       // speaker stem with Cin padded 39 -> 40 (tensor-core path only)
    std::vector<BlockW> blocks;
    BnW final_bn;
    ConvW lstm_in[2];                     // [feat,1024] projection (bias = LSTM bias)
    ConvW lstm_rec[2];                    // [256,1024] recurrent (bias = zeros)
    const float* lstm_rec_fused[2] = {nullptr, nullptr};   // chunk stream for lstm_fused_kernel
    const float* lstm_rec_f16[2] = {nullptr, nullptr};     // ... as fp16 chunks (MMLA_PRECISION_F16)
    const float* lstm_in_fused[2] = {nullptr, nullptr};    // chunk stream for xproj_fused_kernel
    const float* dense_k = nullptr;
    const float* dense_b = nullptr;
    int in_h = 0, in_w = 0, in_c = 0;     // per-clip input geometry
    int seq_len = 0, feat = 0;
    long long per_clip_floats = 0;        // workspace floats per clip
    int micro = 0;                        // clips per micro-batch
    bool fuse_stages = true;              // speaker, TF32: one launch per ResNet stage (MMLA_NET_FUSE_STAGES=0: per unit)
    bool fuse_stem = true;                // ... and the stem inside the first stage when the input is MFCC-13 rows (MMLA_NET_FUSE_STEM=0: own launch)
};

// conv_tc.cu
int mmla_tc_ntile(int n);
long long mmla_tc_arranged_floats(int K, int N);
void mmla_tc_arrange_weights(const float* w, int K, int N, float* out);
int mmla_launch_conv_tc(const ConvArgs& a, const float* wg, cudaStream_t st);
int mmla_launch_stem_fused(const float* x, const float* wg, const float* bias, float* y, long long B, cudaStream_t st);
int mmla_launch_stem_from_cepstra(const float* cep, long long cep_clip_stride, int n_frames, const float* wg, const float* bias,
                                  float* y, long long B, cudaStream_t st);
// resunit_fused.cu
int mmla_launch_resstage_fused(const float* x, float* y, long long B, int T, int Cin, int C, const float* const (*p)[8],
                               const float* ws, const float* bs, const float* fin_scale, const float* fin_shift,
                               float* pooled, cudaStream_t st);
int mmla_launch_resstage_stem_fused(const float* cepstra, long long cep_clip_stride, int n_frames, const float* stem_w,
                                    const float* stem_b, float* y, long long B, const float* const (*p)[8], const float* ws,
                                    const float* bs, cudaStream_t st);
int mmla_launch_resunit_fused(const float* x, float* y, long long B, int T, int Cin, int C, const float* bn1_scale,
                              const float* bn1_shift, const float* w1, const float* b1, const float* bn2_scale,
                              const float* bn2_shift, const float* w2, const float* b2, const float* ws, const float* bs,
                              cudaStream_t st);
// resblock2d_fused.cu (overlap net: both convolutions of a res_block in one launch)
bool mmla_resblock2d_eligible(int H, int W, int Cin, int C, int kh1, int kw1, int kh2, int kw2, int act);
int mmla_launch_resblock2d_fused(const float* x, float* y, long long B, int H, int W, int Cin, int C, const float* bn1_scale,
                                 const float* bn1_shift, const float* w1, const float* b1, const float* bn2_scale,
                                 const float* bn2_shift, const float* w2, const float* b2, const float* res,
                                 long long res_row_stride, cudaStream_t st, const void* img = nullptr, int img_is_u8 = 0,
                                 const float* stem_w = nullptr, const float* stem_b = nullptr, int hpool = 0,
                                 const float* w1_pair = nullptr, const float* w2_pair = nullptr, const void* w1_h = nullptr,
                                 const void* w2_h = nullptr, const void* xa = nullptr, void* ya = nullptr,
                                 const float* ya_scale = nullptr, const float* ya_shift = nullptr, int y_f16 = 0);
void mmla_rb_arrange_weights_pair(const float* w, int K, int N, float* out);
long long mmla_rb_f16_arranged_halves(int K, int N);
void mmla_rb_arrange_weights_f16(const float* w, int K, int N, uint16_t* out);
bool mmla_rb_pair_wanted(int Cin, int C);
// lstm_fused.cu
long long mmla_xproj_arranged_floats();
void mmla_xproj_arrange_weights(const float* W, float* out);
int mmla_launch_xproj_fused(const float* seq, const float* w_f, const float* w_b, const float* b_f, const float* b_b,
                            float* xp_f, float* xp_b, long long B, int T, cudaStream_t st);
long long mmla_lstm_arranged_floats();
void mmla_lstm_arrange_weights(const float* U, float* out);
long long mmla_lstm_arranged_halves();
void mmla_lstm_arrange_weights_f16(const float* U, uint16_t* out);
int mmla_launch_lstm_fused(const float* xp_f, const float* xp_b, const float* wr_f, const float* wr_b, float* h_f,
                           float* h_b, float* scratch_f, float* scratch_b, long long B, int T, cudaStream_t st, int f16 = 0);

namespace {

int same_out(int n, int s) { return (n + s - 1) / s; }
int same_pad_before(int n, int k, int s) {
    const int out = same_out(n, s);
    int total = (out - 1) * s + k - n;
    if (total < 0) total = 0;
    return total / 2;
}

int launch_conv(const ConvW& c, const void* x, int x_is_u8, long long B, int H, int W, const BnW* pre, int pre_act,
                const float* res, long long res_row_stride, float* y, cudaStream_t st, bool tc = false) {
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.x_is_u8 = x_is_u8; a.w = c.k; a.bias = c.b;
    a.pre_scale = pre ? pre->scale : nullptr;
    a.pre_shift = pre ? pre->shift : nullptr;
    a.pre_act = pre_act;
    a.res = res; a.res_row_stride = res_row_stride; a.y = y;
    a.H = H; a.W = W; a.Cin = c.cin;
    a.Ho = same_out(H, c.stride); a.Wo = same_out(W, c.stride);
    a.N = c.cout; a.K = c.kh * c.kw * c.cin;
    a.kh = c.kh; a.kw = c.kw; a.stride = c.stride;
    a.pad_t = same_pad_before(H, c.kh, c.stride);
    a.pad_l = same_pad_before(W, c.kw, c.stride);
    a.M = B * a.Ho * a.Wo;
    if (a.M == 0) return MMLA_OK;
    if (tc && c.k_tc) return mmla_launch_conv_tc(a, c.k_tc, st);
    const unsigned gx = static_cast<unsigned>((a.M + kBM - 1) / kBM);
    if (c.cout <= 16) {
        conv_igemm_kernel<1><<<dim3(gx, (c.cout + 15) / 16), 256, 0, st>>>(a);
        mmla_count_launch("conv_igemm_kernel", st);
    } else if (c.cout <= 32) {
        conv_igemm_kernel<2><<<dim3(gx, (c.cout + 31) / 32), 256, 0, st>>>(a);
        mmla_count_launch("conv_igemm_kernel", st);
    } else if (c.cout <= 64) {
        conv_igemm_kernel<4><<<dim3(gx, (c.cout + 63) / 64), 256, 0, st>>>(a);
        mmla_count_launch("conv_igemm_kernel", st);
    } else {
        conv_igemm_kernel<8><<<dim3(gx, (c.cout + 127) / 128), 256, 0, st>>>(a);
        mmla_count_launch("conv_igemm_kernel", st);
    }
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}

unsigned ew_grid(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = 32LL * (mmla_num_sms() > 0 ? mmla_num_sms() : 148);
    return static_cast<unsigned>(g < 1 ? 1 : (g > cap ? cap : g));
}

// host-side blob reader (spec traversal order, see mmla_audio_b200/models.py::pack_weights)
struct BlobReader {
    const float* p;
    long long n, pos = 0;
    bool ok = true;
    const float* take(long long cnt) {
        if (pos + cnt > n) { ok = false; return p; }
        const float* r = p + pos;
        pos += cnt;
        return r;
    }
};

}  // namespace

#define EXPORT extern "C" __attribute__((visibility("default")))

EXPORT int mmla_net_create(int32_t kind, int32_t n_classes, int32_t head, const float* w_host, int64_t n_weights,
                           MmlaNet** out_net) {
    MMLA_REQUIRE(out_net && w_host, MMLA_EINVAL, "net_create: null argument");
    MMLA_REQUIRE(kind == MMLA_NET_OVERLAP || kind == MMLA_NET_SPEAKER, MMLA_EINVAL, "net_create: bad kind %d", kind);
    MMLA_REQUIRE(n_classes >= 1 && n_classes <= 4096, MMLA_EINVAL, "net_create: bad n_classes %d", n_classes);
    MMLA_REQUIRE(head == MMLA_HEAD_SOFTMAX || head == MMLA_HEAD_SIGMOID, MMLA_EINVAL, "net_create: bad head %d", head);
    MMLA_REQUIRE(mmla_num_sms() > 0, MMLA_ECUDA, "net_create: no CUDA device");
    const bool ov = kind == MMLA_NET_OVERLAP;
    // host staging blob: [weights as given][folded BN scale/shift][zero recurrent bias]
    std::vector<float> stage(w_host, w_host + n_weights);
    struct Fix { const float** dst; long long off; };
    std::vector<Fix> fixes;
    MmlaNet* net = new MmlaNet;
    net->kind = kind; net->n_classes = n_classes; net->head = head;
    BlobReader rd{w_host, n_weights};
    auto take_ptr = [&](const float** dst, long long cnt) {
        const long long off = rd.pos;
        rd.take(cnt);
        fixes.push_back({dst, off});
    };
    auto add_tc = [&](ConvW& c, const float* wsrc) {     // tensor-core copy of the weights (if eligible)
        const int K = c.kh * c.kw * c.cin;
        if (!rd.ok || mmla_tc_ntile(c.cout) == 0) return;
        while (stage.size() % 4) stage.push_back(0.f);
        const long long off = static_cast<long long>(stage.size());
        stage.resize(stage.size() + mmla_tc_arranged_floats(K, c.cout));
        mmla_tc_arrange_weights(wsrc, K, c.cout, stage.data() + off);
        fixes.push_back({&c.k_tc, off});
        if (ov && c.cout >= 64 && c.cout <= 128 && K % 32 == 0 && c.kh * c.kw > 1) {
            const long long off2 = static_cast<long long>(stage.size());
            stage.resize(stage.size() + mmla_tc_arranged_floats(K, c.cout));
            mmla_rb_arrange_weights_pair(wsrc, K, c.cout, stage.data() + off2);
            fixes.push_back({&c.k_tc2, off2});
        }
        if (ov && c.cout >= 32 && c.cout <= 128 && c.cin % 16 == 0 && c.kh * c.kw > 1) {
            // fp16 chunks of the conv pairs (MMLA_PRECISION_F16): the halves ride in the float blob, two per element
            const long long halves = mmla_rb_f16_arranged_halves(K, c.cout);
            std::vector<uint16_t> hbuf(static_cast<size_t>(halves));
            mmla_rb_arrange_weights_f16(wsrc, K, c.cout, hbuf.data());
            while (stage.size() % 4) stage.push_back(0.f);
            const long long off3 = static_cast<long long>(stage.size());
            stage.resize(stage.size() + static_cast<size_t>((halves + 1) / 2));
            memcpy(stage.data() + off3, hbuf.data(), static_cast<size_t>(halves) * 2);
            fixes.push_back({&c.k_h, off3});
        }
    };
    auto take_conv = [&](ConvW& c, int kh, int kw, int cin, int cout, int stride) {
        c.kh = kh; c.kw = kw; c.cin = cin; c.cout = cout; c.stride = stride;
        const float* wsrc = rd.p + rd.pos;
        take_ptr(&c.k, static_cast<long long>(kh) * kw * cin * cout);
        take_ptr(&c.b, cout);
        add_tc(c, wsrc);
    };
    auto take_bn = [&](BnW& bn, int ch) {
        const float* g = rd.take(ch);
        const float* be = rd.take(ch);
        const float* mu = rd.take(ch);
        const float* var = rd.take(ch);
        if (!rd.ok) return;
        while (stage.size() % 4) stage.push_back(0.f);             // float4-aligned scale/shift
        const long long off = static_cast<long long>(stage.size());
        for (int i = 0; i < ch; ++i) stage.push_back(static_cast<float>(g[i] / sqrt(static_cast<double>(var[i]) + kBnEps)));
        for (int i = 0; i < ch; ++i) {
            const double sc = g[i] / sqrt(static_cast<double>(var[i]) + kBnEps);
            stage.push_back(static_cast<float>(be[i] - mu[i] * sc));
        }
        fixes.push_back({&bn.scale, off});
        fixes.push_back({&bn.shift, off + ch});
    };
    const int chans[3] = {32, 64, 128};
    if (ov) {
        net->in_h = 128; net->in_w = 151; net->in_c = 3;
        take_conv(net->stem, 1, 1, 3, 16, 1);
    } else {
        net->in_h = 1; net->in_w = 256; net->in_c = 39;
        const float* stem_src = rd.p + rd.pos;
        take_conv(net->stem, 1, 4, 39, 32, 1);
        if (rd.ok) {                      // zero-padded copy [4][40][32] for the vectorised gather
            std::vector<float> wpad(4 * 40 * 32, 0.f);
            for (int k = 0; k < 4; ++k)
                for (int c = 0; c < 39; ++c) memcpy(&wpad[(k * 40 + c) * 32], stem_src + (k * 39 + c) * 32, 32 * sizeof(float));
            net->stem_pad = net->stem;
            net->stem_pad.cin = 40;
            net->stem_pad.k = nullptr;
            while (stage.size() % 4) stage.push_back(0.f);
            const long long off = static_cast<long long>(stage.size());
            stage.resize(stage.size() + mmla_tc_arranged_floats(160, 32));
            mmla_tc_arrange_weights(wpad.data(), 160, 32, stage.data() + off);
            fixes.push_back({&net->stem_pad.k_tc, off});
            fixes.push_back({&net->stem_pad.b, static_cast<long long>((stem_src + 4 * 39 * 32) - w_host)});
        }
    }
    int cin = ov ? 16 : 32;
    net->blocks.resize(9);
    for (int s = 0; s < 3; ++s)
        for (int r = 0; r < 3; ++r) {
            BlockW& b = net->blocks[3 * s + r];
            const int cout = chans[s];
            b.pool = (r == 0);
            take_bn(b.bn1, cin);
            take_conv(b.conv1, ov ? 3 : 1, 3, cin, cout, 1);
            take_bn(b.bn2, cout);
            take_conv(b.conv2, ov ? 4 : 1, ov ? 1 : 3, cout, cout, 1);
            if (b.pool) take_conv(b.shortcut, 1, 1, cin, cout, 2);
            cin = cout;
        }
    if (!ov) take_bn(net->final_bn, 128);
    net->feat = 128;
    net->seq_len = ov ? 19 : 8;
    long long zero_bias_off = -1;
    for (int d = 0; d < 2; ++d) {
        ConvW& pi = net->lstm_in[d];
        ConvW& pr = net->lstm_rec[d];
        pi.kh = pi.kw = 1; pi.cin = 128; pi.cout = 1024; pi.stride = 1;
        pr.kh = pr.kw = 1; pr.cin = 256; pr.cout = 1024; pr.stride = 1;
        const float* wi_src = rd.p + rd.pos;
        take_ptr(&pi.k, 128LL * 1024);
        const float* wr_src = rd.p + rd.pos;
        take_ptr(&pr.k, 256LL * 1024);
        take_ptr(&pi.b, 1024);
        add_tc(pi, wi_src);
        add_tc(pr, wr_src);
        if (rd.ok) {
            while (stage.size() % 4) stage.push_back(0.f);
            const long long off = static_cast<long long>(stage.size());
            stage.resize(stage.size() + mmla_lstm_arranged_floats());
            mmla_lstm_arrange_weights(wr_src, stage.data() + off);
            fixes.push_back({&net->lstm_rec_fused[d], off});
            {
                const long long halves = mmla_lstm_arranged_halves();
                std::vector<uint16_t> hbuf(static_cast<size_t>(halves));
                mmla_lstm_arrange_weights_f16(wr_src, hbuf.data());
                const long long off_h = static_cast<long long>(stage.size());       // still 16-byte aligned: the chunk stream above is
                stage.resize(stage.size() + static_cast<size_t>(halves / 2));
                memcpy(stage.data() + off_h, hbuf.data(), static_cast<size_t>(halves) * 2);
                fixes.push_back({&net->lstm_rec_f16[d], off_h});
            }
            const long long off_in = static_cast<long long>(stage.size());
            stage.resize(stage.size() + mmla_xproj_arranged_floats());
            mmla_xproj_arrange_weights(wi_src, stage.data() + off_in);
            fixes.push_back({&net->lstm_in_fused[d], off_in});
        }
        if (zero_bias_off < 0) {
            while (stage.size() % 4) stage.push_back(0.f);
            zero_bias_off = static_cast<long long>(stage.size());
            stage.insert(stage.end(), 1024, 0.f);
        }
        fixes.push_back({&pr.b, zero_bias_off});
    }
    take_ptr(&net->dense_k, 512LL * n_classes);
    take_ptr(&net->dense_b, n_classes);
    if (!rd.ok || rd.pos != n_weights) {
        mmla_set_error("net_create: weight blob has %lld floats, the %s net with %d classes needs %lld",
                       static_cast<long long>(n_weights), ov ? "overlap" : "speaker", n_classes, rd.pos);
        delete net;
        return MMLA_EINVAL;
    }
    while (stage.size() % 4) stage.push_back(0.f);
    cudaError_t e = cudaMalloc(&net->dev, stage.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(net->dev, stage.data(), stage.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        mmla_set_error("net_create: device upload failed: %s", cudaGetErrorString(e));
        if (net->dev) cudaFree(net->dev);
        delete net;
        return MMLA_ECUDA;
    }
    for (const Fix& f : fixes) *f.dst = net->dev + f.off;

    // workspace plan (floats per clip): three rotating activation buffers + LSTM scratch
    const long long act = ov ? 128LL * 151 * 32 : 256LL * 32;
    const long long T = net->seq_len;
    // (the LSTM buffers of the tensor-core path are laid out in whole 128-clip tiles: mmla_net_workspace_bytes rounds up)
    // (overlap: + two fp16 [64,76,32] buffers = the conv pairs' activated operands of the fp16-operand mode)
    net->per_clip_floats = 3 * act + T * 128 + 2 * T * 1024 + 1024 + 2 * 256 + 6 * 256 + (ov ? 2 * kOvHalfFloats : 256 * 40);
    net->micro = ov ? 512 : 4096;         // measured on B200: larger micro-batches win (overlap, 512 clips: 24.8 ms at 128, 21.4 ms at 512)
    if (const char* e = getenv("MMLA_NET_MICRO")) {
        const int v = atoi(e);
        if (v > 0) net->micro = v;
    }
    if (const char* e = getenv("MMLA_NET_FUSE_STAGES")) net->fuse_stages = atoi(e) != 0;
    if (const char* e = getenv("MMLA_NET_FUSE_STEM")) net->fuse_stem = atoi(e) != 0;
    *out_net = net;
    return MMLA_OK;
}

EXPORT void mmla_net_destroy(MmlaNet* net) {
    if (!net) return;
    if (net->dev) cudaFree(net->dev);
    delete net;
}

EXPORT int64_t mmla_net_workspace_bytes(const MmlaNet* net, int64_t batch) {
    if (!net || batch < 0) return -1;
    const long long mb = batch < net->micro ? batch : net->micro;
    const long long mbp = ((mb < 1 ? 1 : mb) + 127) / 128 * 128;
    return mbp * net->per_clip_floats * static_cast<long long>(sizeof(float)) + 256;
}

static int forward_impl(MmlaNet* net, const void* x, int32_t x_is_u8, int32_t cep_frames, int64_t cep_clip_stride, int64_t batch,
                        void* workspace, int64_t workspace_bytes, float* prob, int32_t* labels, void* stream,
                        float* embed = nullptr);

EXPORT int mmla_net_forward(MmlaNet* net, const void* x, int32_t x_is_u8, int64_t batch, void* workspace,
                            int64_t workspace_bytes, float* prob, int32_t* labels, void* stream) {
    return forward_impl(net, x, x_is_u8, 0, 0, batch, workspace, workspace_bytes, prob, labels, stream);
}

EXPORT int mmla_net_forward_cepstra(MmlaNet* net, const float* cepstra, int64_t cep_clip_stride, int32_t n_frames, int64_t batch,
                                    void* workspace, int64_t workspace_bytes, float* prob, int32_t* labels, void* stream) {
    MMLA_REQUIRE(net && net->kind == MMLA_NET_SPEAKER && (net->precision == MMLA_PRECISION_TF32 || net->precision == MMLA_PRECISION_F16) &&
                     net->stem_pad.k_tc,
                 MMLA_EUNSUP, "net_forward_cepstra: needs the speaker net in a tensor-core mode (TF32 / F16)");
    MMLA_REQUIRE(n_frames >= 1 && n_frames <= 256 && cep_clip_stride >= static_cast<int64_t>(n_frames) * 16, MMLA_EINVAL,
                 "net_forward_cepstra: bad cepstra geometry (n_frames %d, clip stride %lld)", n_frames,
                 static_cast<long long>(cep_clip_stride));
    return forward_impl(net, cepstra, 3, n_frames, cep_clip_stride, batch, workspace, workspace_bytes, prob, labels, stream);
}

EXPORT int mmla_net_embed(MmlaNet* net, const void* x, int32_t x_is_u8, int64_t batch, void* workspace, int64_t workspace_bytes,
                          float* embed, void* stream) {
    MMLA_REQUIRE(embed != nullptr, MMLA_EINVAL, "net_embed: null output");
    return forward_impl(net, x, x_is_u8, 0, 0, batch, workspace, workspace_bytes, nullptr, nullptr, stream, embed);
}

static int forward_impl(MmlaNet* net, const void* x, int32_t x_is_u8, int32_t cep_frames, int64_t cep_clip_stride, int64_t batch,
                        void* workspace, int64_t workspace_bytes, float* prob, int32_t* labels, void* stream, float* embed) {
    MMLA_REQUIRE(net && x && (prob || embed) && workspace, MMLA_EINVAL, "net_forward: null argument");
    MMLA_REQUIRE(batch >= 0, MMLA_EINVAL, "net_forward: negative batch");
    const bool from_cep = x_is_u8 == 3;       // MFCC-13 rows; delta / delta-delta / padding happen inside the stem kernel
    if (from_cep) x_is_u8 = 0;
    MMLA_REQUIRE(x_is_u8 >= 0 && x_is_u8 <= 2, MMLA_EINVAL, "net_forward: bad input kind %d", x_is_u8);
    MMLA_REQUIRE(x_is_u8 != 1 || net->kind == MMLA_NET_OVERLAP, MMLA_EINVAL, "net_forward: uint8 input is for the overlap net");
    MMLA_REQUIRE(x_is_u8 != 2 || net->kind == MMLA_NET_SPEAKER, MMLA_EINVAL, "net_forward: 40-channel input is for the speaker net");
    const bool pad40 = x_is_u8 == 2;          // speaker features already laid out [B,256,40] with channel 39 = 0
    if (pad40) x_is_u8 = 0;
    MMLA_REQUIRE(workspace_bytes >= mmla_net_workspace_bytes(net, batch), MMLA_EINVAL,
                 "net_forward: workspace too small (%lld < %lld bytes)", static_cast<long long>(workspace_bytes),
                 static_cast<long long>(mmla_net_workspace_bytes(net, batch)));
    MMLA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, MMLA_EINVAL, "net_forward: workspace must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool ov = net->kind == MMLA_NET_OVERLAP;
    const bool tc = net->precision == MMLA_PRECISION_TF32 || net->precision == MMLA_PRECISION_F16;
    const bool f16 = net->precision == MMLA_PRECISION_F16;     // overlap net: the conv pairs take fp16 operands, the rest is TF32
    const long long in_elems = from_cep ? cep_clip_stride : static_cast<long long>(net->in_h) * net->in_w * (pad40 ? 40 : net->in_c);
    const long long act = ov ? 128LL * 151 * 32 : 256LL * 32;
    const int T = net->seq_len;

    for (long long b0 = 0; b0 < batch; b0 += net->micro) {
        const long long B = (batch - b0 < net->micro) ? batch - b0 : net->micro;
        float* ws = static_cast<float*>(workspace);
        float* buf[3] = {ws, ws + B * act, ws + 2 * B * act};
        float* seq = ws + 3 * B * act;                    // [B,T,128]
        const long long Bp = (B + 127) / 128 * 128;       // the tensor-core LSTM path works in whole 128-clip tiles
        float* xp[2] = {seq + B * T * 128, seq + B * T * 128 + Bp * T * 1024};  // per direction: [B,T,1024], or row-tiled
        float* z = xp[1] + Bp * T * 1024;                 // [B,1024] gate pre-activations
        float* hdir[2] = {z + B * 1024, z + B * 1024 + B * 256};   // final h of the fwd / bwd layer
        float* cst = hdir[1] + B * 256;                   // [B,256] cell state; fused kernel: 2 x 3 row-tiled [Bp,256] blocks
        float* xpad = cst + 6 * Bp * 256;                 // speaker: [B,256,40] channel-padded input
        // overlap, fp16-operand mode: ELU(BN1(x)) of the current block input as fp16 (written by the kernel that produced x)
        unsigned short* hact[2] = {reinterpret_cast<unsigned short*>(xpad), reinterpret_cast<unsigned short*>(xpad + Bp * kOvHalfFloats)};
        int hcur = 0;
        bool have_xa = false;                             // hact[hcur] holds the activated copy of buf[cur]

        const void* xin = x_is_u8 ? static_cast<const void*>(static_cast<const unsigned char*>(x) + b0 * in_elems)
                                  : static_cast<const void*>(static_cast<const float*>(x) + b0 * in_elems);
        int H = net->in_h, W = net->in_w;
        int cur = 0;
        int rc;
        bool seq_done = false;
        bool stem_folded = false;
        const int act_kind = ov ? ACT_ELU : ACT_RELU;
        auto pair_fusable = [&](const BlockW& k, int h, int w) {
            return k.conv1.k_tc && k.conv2.k_tc && k.conv1.stride == 1 && k.conv2.stride == 1 && k.conv2.cin == k.conv1.cout &&
                   k.conv2.cout == k.conv1.cout && k.bn1.scale && k.bn2.scale &&
                   mmla_resblock2d_eligible(h, w, k.conv1.cin, k.conv1.cout, k.conv1.kh, k.conv1.kw, k.conv2.kh, k.conv2.kw, act_kind);
        };
        // fp16-operand mode: the BN1 of block bi + 1 when that block will run on the fp16 conv-pair kernel from a [h,w,c] input
        // that fits the fp16 buffers (its producer then also writes the activated fp16 operand), else null
        auto next_bn = [&](size_t bi, int c, int h, int w) -> const BlockW::BnPtr {
            const char* e = getenv("MMLA_NET_F16_ACT");       // "0": every fill converts from the fp32 tensor
            if (!f16 || (e && e[0] == '0') || bi + 1 >= net->blocks.size()) return nullptr;
            const BlockW& k = net->blocks[bi + 1];
            if (!k.conv1.k_h || !k.conv2.k_h || k.conv1.cin != c || !pair_fusable(k, h, w)) return nullptr;
            if (static_cast<long long>(h) * w * c > 2 * kOvHalfFloats) return nullptr;
            return &k.bn1;
        };
        auto pool_fusable = [&](const BlockW& k) {
            const char* fp = getenv("MMLA_NET_FUSE_POOL");
            return k.pool && !(fp && fp[0] == '0') && k.shortcut.kh == 1 && k.shortcut.kw == 1 && k.shortcut.stride == 2 &&
                   k.shortcut.cin % 4 == 0 && k.conv2.cout % 4 == 0 &&
                   static_cast<size_t>(k.shortcut.cin) * k.conv2.cout * sizeof(float) <= 48 * 1024;
        };
        // label pipeline: the stem runs inside the first stage's kernel (resstage_fused.cu, STEM variant)
        auto unit_tc = [&](const BlockW& k) { return k.conv1.k_tc && k.conv2.k_tc && (!k.pool || k.shortcut.k_tc); };
        const bool stem_in_stage = from_cep && tc && !ov && net->fuse_stages && net->fuse_stem && net->stem_pad.k_tc &&
                                   net->blocks.size() >= 3 && net->blocks[0].pool && !net->blocks[1].pool && !net->blocks[2].pool &&
                                   unit_tc(net->blocks[0]) && unit_tc(net->blocks[1]) && unit_tc(net->blocks[2]) &&
                                   net->blocks[0].conv1.cin == 32 && net->blocks[0].conv1.cout == 32 &&
                                   net->blocks[1].conv1.cin == 32 && net->blocks[2].conv1.cin == 32 && W == 256;
        if (stem_in_stage) {
            rc = 0;
        } else if (from_cep) {
            rc = mmla_launch_stem_from_cepstra(static_cast<const float*>(xin), cep_clip_stride, cep_frames, net->stem_pad.k_tc,
                                               net->stem_pad.b, buf[cur], B, st);                     // stem_fused.cu
        } else if (tc && !ov && net->stem_pad.k_tc) {
            const float* x40 = static_cast<const float*>(xin);
            if (!pad40) {
                pad_channels_kernel<<<ew_grid(B * 256 * 10), 256, 0, st>>>(static_cast<const float*>(xin), xpad, B * 256, 39, 40);
                mmla_count_launch("pad_channels_kernel", st);
                MMLA_CUDA_CHECK(cudaGetLastError());
                x40 = xpad;
            }
            rc = mmla_launch_stem_fused(x40, net->stem_pad.k_tc, net->stem_pad.b, buf[cur], B, st);   // stem_fused.cu
        } else if (pad40) {
            mmla_set_error("net_forward: 40-channel input needs the TF32 tensor-core mode");
            return MMLA_EUNSUP;
        } else if (ov && tc && net->stem.kh == 1 && net->stem.kw == 1 && net->stem.cin == 3 && net->stem.cout == 16) {
            const long long pixels = B * H * W;
            // first block eligible: the stem runs inside its conv-pair kernel and inside its pooling kernel (resblock2d_fused.cu,
            // STEM mode) and the [B,128,151,16] stem tensor is never materialised (MMLA_NET_FUSE_STEM2D=0: own launch)
            {
                const char* e = getenv("MMLA_NET_FUSE_STEM2D");
                stem_folded = !(e && e[0] == '0') && !net->blocks.empty() && net->blocks[0].conv1.cin == 16 &&
                              net->blocks[0].conv1.cout == 32 && pair_fusable(net->blocks[0], H, W) && pool_fusable(net->blocks[0]);
            }
            if (pixels > 0 && !stem_folded) {
                stem1x1_kernel<<<ew_grid(pixels * 4), 256, 0, st>>>(xin, x_is_u8, net->stem.k, net->stem.b, buf[cur], pixels);
                mmla_count_launch("stem1x1_kernel", st);
                MMLA_CUDA_CHECK(cudaGetLastError());
            }
            rc = 0;
        } else {
            rc = launch_conv(net->stem, xin, x_is_u8, B, H, W, nullptr, ACT_NONE, nullptr, 0, buf[cur], st, tc);
        }
        if (rc) return rc;
        for (size_t bi = 0; bi < net->blocks.size(); ++bi) {
            const BlockW& blk = net->blocks[bi];
            float* X = buf[cur];
            float* A = buf[(cur + 1) % 3];
            float* Bf = buf[(cur + 2) % 3];
            const bool xa_in = have_xa;                   // hact[hcur] = the activated fp16 copy of X (fp16-operand mode)
            have_xa = false;                              // set again by the kernel that produces the next block's input
            auto fusable = [&](const BlockW& k) { return k.conv1.k_tc && k.conv2.k_tc && (!k.pool || k.shortcut.k_tc); };
            if (tc && !ov && net->fuse_stages && blk.pool && bi + 2 < net->blocks.size() && fusable(blk) &&
                fusable(net->blocks[bi + 1]) && fusable(net->blocks[bi + 2]) && !net->blocks[bi + 1].pool &&
                !net->blocks[bi + 2].pool && net->blocks[bi + 1].conv1.cin == blk.conv1.cout &&
                net->blocks[bi + 2].conv1.cin == blk.conv1.cout) {
                // a whole ResNet stage (pooled unit + two plain units) in ONE launch: the activations between the
                // units stay in shared memory (resunit_fused.cu, stage mode)
                const int Wo = same_out(W, 2);
                const bool with_stem = stem_in_stage && bi == 0;
                const float* prm[3][8];
                for (int u = 0; u < 3; ++u) {
                    const BlockW& k = net->blocks[bi + u];
                    prm[u][0] = k.bn1.scale; prm[u][1] = k.bn1.shift; prm[u][2] = k.conv1.k_tc; prm[u][3] = k.conv1.b;
                    prm[u][4] = k.bn2.scale; prm[u][5] = k.bn2.shift; prm[u][6] = k.conv2.k_tc; prm[u][7] = k.conv2.b;
                }
                // the last stage also applies the net's tail (BN -> ReLU -> AveragePooling1D(4)) and writes `seq` directly
                const bool tail = bi + 3 == net->blocks.size() && Wo % 4 == 0 && blk.conv1.cout == 128;
                if (with_stem)
                    rc = mmla_launch_resstage_stem_fused(static_cast<const float*>(xin), cep_clip_stride, cep_frames,
                                                         net->stem_pad.k_tc, net->stem_pad.b, A, B, prm, blk.shortcut.k_tc,
                                                         blk.shortcut.b, st);
                else
                    rc = mmla_launch_resstage_fused(X, A, B, Wo, blk.conv1.cin, blk.conv1.cout, prm, blk.shortcut.k_tc,
                                                    blk.shortcut.b, tail ? net->final_bn.scale : nullptr,
                                                    tail ? net->final_bn.shift : nullptr, tail ? seq : nullptr, st);
                if (rc) return rc;
                seq_done = tail;
                W = Wo;
                cur = (cur + 1) % 3;
                bi += 2;
            } else if (tc && !ov && blk.conv1.k_tc && blk.conv2.k_tc && (!blk.pool || blk.shortcut.k_tc)) {
                // speaker res_unit, tensor-core mode: ONE fused kernel per unit (resunit_fused.cu) —
                // pooled units max-pool on load and fold the stride-2 shortcut into the accumulator
                const int Wo = blk.pool ? same_out(W, 2) : W;
                if ((rc = mmla_launch_resunit_fused(X, A, B, Wo, blk.conv1.cin, blk.conv1.cout, blk.bn1.scale, blk.bn1.shift,
                                                    blk.conv1.k_tc, blk.conv1.b, blk.bn2.scale, blk.bn2.shift,
                                                    blk.conv2.k_tc, blk.conv2.b, blk.pool ? blk.shortcut.k_tc : nullptr,
                                                    blk.pool ? blk.shortcut.b : nullptr, st)))
                    return rc;
                W = Wo;
                cur = (cur + 1) % 3;
            } else if (!blk.pool) {
                // out = conv2(act(bn2(conv1(act(bn1(x)))))) + x
                if (ov && tc && pair_fusable(blk, H, W)) {
                    // both convolutions in one launch, the intermediate stays in shared memory (resblock2d_fused.cu)
                    const bool h16 = f16 && blk.conv1.k_h && blk.conv2.k_h;
                    const BnW* nbn = next_bn(bi, blk.conv2.cout, H, W);      // fp16 mode: this block also writes the next one's operand
                    if ((rc = mmla_launch_resblock2d_fused(X, Bf, B, H, W, blk.conv1.cin, blk.conv1.cout, blk.bn1.scale, blk.bn1.shift,
                                                           blk.conv1.k_tc, blk.conv1.b, blk.bn2.scale, blk.bn2.shift, blk.conv2.k_tc,
                                                           blk.conv2.b, X, blk.conv2.cout, st, nullptr, 0, nullptr, nullptr, 0,
                                                           blk.conv1.k_tc2, blk.conv2.k_tc2, h16 ? blk.conv1.k_h : nullptr,
                                                           h16 ? blk.conv2.k_h : nullptr, h16 && xa_in ? hact[hcur] : nullptr,
                                                           h16 && nbn ? hact[hcur ^ 1] : nullptr, nbn ? nbn->scale : nullptr,
                                                           nbn ? nbn->shift : nullptr)))
                        return rc;
                    have_xa = h16 && nbn;
                    hcur ^= 1;
                } else {
                    if ((rc = launch_conv(blk.conv1, X, 0, B, H, W, &blk.bn1, act_kind, nullptr, 0, A, st, tc))) return rc;
                    if ((rc = launch_conv(blk.conv2, A, 0, B, H, W, &blk.bn2, act_kind, X, blk.conv2.cout, Bf, st, tc))) return rc;
                }
                cur = (cur + 2) % 3;
            } else if (ov) {
                // full-resolution convs, MaxPool2x2 'same', then shortcut conv (stride 2) + pooled
                const bool fold = stem_folded && bi == 0;
                // the maximum over each pooling window's two rows is taken in the conv-pair kernel's epilogue (HPOOL): Bf is then
                // [B, H/2, W, C] and the pooling kernel reads half as much (MMLA_NET_FUSE_HPOOL=0: full-resolution Bf)
                bool hpool = false, z16 = false;
                if (tc && pair_fusable(blk, H, W) && pool_fusable(blk) && H % 2 == 0) {
                    const char* e = getenv("MMLA_NET_FUSE_HPOOL");
                    hpool = !(e && e[0] == '0');
                }
                if (tc && pair_fusable(blk, H, W)) {
                    const bool h16 = f16 && blk.conv1.k_h && blk.conv2.k_h;
                    {   // fp16 mode, OPT-IN (MMLA_NET_F16_Z=1): the row-pooled conv output goes to the pooling kernel as fp16.  Measured
                        // 3.488 -> 3.470 ms per 512 clips only (pool_shortcut_kernel 0.645 -> 0.625 ms: it is not bound by reading z),
                        // which does not pay for one more 11-bit rounding on the conv branch: off by default.
                        const char* e = getenv("MMLA_NET_F16_Z");
                        z16 = h16 && hpool && pool_fusable(blk) && e && e[0] == '1';
                    }
                    if ((rc = mmla_launch_resblock2d_fused(X, Bf, B, H, W, blk.conv1.cin, blk.conv1.cout, blk.bn1.scale, blk.bn1.shift,
                                                           blk.conv1.k_tc, blk.conv1.b, blk.bn2.scale, blk.bn2.shift, blk.conv2.k_tc,
                                                           blk.conv2.b, nullptr, 0, st, fold ? xin : nullptr, x_is_u8,
                                                           fold ? net->stem.k : nullptr, fold ? net->stem.b : nullptr, hpool ? 1 : 0,
                                                           blk.conv1.k_tc2, blk.conv2.k_tc2, h16 ? blk.conv1.k_h : nullptr,
                                                           h16 ? blk.conv2.k_h : nullptr, h16 && xa_in ? hact[hcur] : nullptr, nullptr,
                                                           nullptr, nullptr, z16 ? 1 : 0)))
                        return rc;
                } else {
                    if ((rc = launch_conv(blk.conv1, X, 0, B, H, W, &blk.bn1, act_kind, nullptr, 0, A, st, tc))) return rc;
                    if ((rc = launch_conv(blk.conv2, A, 0, B, H, W, &blk.bn2, act_kind, nullptr, 0, Bf, st, tc))) return rc;
                }
                const int Ho = same_out(H, 2), Wo = same_out(W, 2), C = blk.conv2.cout;
                const size_t wbytes = static_cast<size_t>(blk.shortcut.cin) * C * sizeof(float) + 256;
                if (tc && pool_fusable(blk)) {
                    // MaxPool + stride-2 shortcut + add in one pass (reads X and Bf, writes A)
                    const BnW* nbn = f16 ? next_bn(bi, C, Ho, Wo) : nullptr;
                    if (z16)
                        pool_shortcut_kernel<true><<<ew_grid(B * Ho * ((Wo + 3) / 4) * C / 4), 256, wbytes, st>>>(
                            X, Bf, blk.shortcut.k, blk.shortcut.b, A, B, H, W, blk.shortcut.cin, C, fold ? xin : nullptr, x_is_u8,
                            fold ? net->stem.k : nullptr, fold ? net->stem.b : nullptr, 1, nbn ? hact[hcur ^ 1] : nullptr,
                            nbn ? nbn->scale : nullptr, nbn ? nbn->shift : nullptr);
                    else
                        pool_shortcut_kernel<false><<<ew_grid(B * Ho * ((Wo + 3) / 4) * C / 4), 256, wbytes, st>>>(
                            X, Bf, blk.shortcut.k, blk.shortcut.b, A, B, H, W, blk.shortcut.cin, C, fold ? xin : nullptr, x_is_u8,
                            fold ? net->stem.k : nullptr, fold ? net->stem.b : nullptr, hpool ? 1 : 0, nbn ? hact[hcur ^ 1] : nullptr,
                            nbn ? nbn->scale : nullptr, nbn ? nbn->shift : nullptr);
                    mmla_count_launch("pool_shortcut_kernel", st);
                    MMLA_CUDA_CHECK(cudaGetLastError());
                    have_xa = nbn != nullptr;
                    hcur ^= 1;
                    H = Ho; W = Wo;
                    cur = (cur + 1) % 3;
                    continue;
                }
                maxpool_kernel<<<ew_grid(B * Ho * Wo * C / 4), 256, 0, st>>>(Bf, A, static_cast<int>(B), H, W, C, 2, 2, Ho, Wo);
                mmla_count_launch("maxpool_kernel", st);
                MMLA_CUDA_CHECK(cudaGetLastError());
                if ((rc = launch_conv(blk.shortcut, X, 0, B, H, W, nullptr, ACT_NONE, A, C, Bf, st, tc))) return rc;
                H = Ho; W = Wo;
                cur = (cur + 2) % 3;
            } else {
                // speaker: x' = MaxPool1D(x); res = conv_k1_s2(x); out = conv2(..conv1(..x'..)) + res
                const int Wo = same_out(W, 2), Cin = blk.conv1.cin;
                maxpool_kernel<<<ew_grid(B * Wo * Cin / 4), 256, 0, st>>>(X, A, static_cast<int>(B), 1, W, Cin, 1, 2, 1, Wo);
                mmla_count_launch("maxpool_kernel", st);
                MMLA_CUDA_CHECK(cudaGetLastError());
                if ((rc = launch_conv(blk.conv1, A, 0, B, 1, Wo, &blk.bn1, act_kind, nullptr, 0, Bf, st, tc))) return rc;
                if ((rc = launch_conv(blk.shortcut, X, 0, B, 1, W, nullptr, ACT_NONE, nullptr, 0, A, st, tc))) return rc;
                // conv2 reads Bf, adds A (shortcut), writes X's buffer (X is dead now)
                if ((rc = launch_conv(blk.conv2, Bf, 0, B, 1, Wo, &blk.bn2, act_kind, A, blk.conv2.cout, X, st, tc))) return rc;
                W = Wo;
            }
        }
        // sequence features [B,T,128]
        if (seq_done) {
            // written by the last fused stage
        } else if (ov) {
            mean_h_kernel<<<ew_grid(B * W * 128), 256, 0, st>>>(buf[cur], seq, B, H, W, 128);
            mmla_count_launch("mean_h_kernel", st);
        } else {
            bn_relu_avgpool4_kernel<<<ew_grid(B * (W / 4) * 128), 256, 0, st>>>(buf[cur], seq, net->final_bn.scale,
                                                                               net->final_bn.shift, B, W, 128);
            mmla_count_launch("bn_relu_avgpool4_kernel", st);
        }
        MMLA_CUDA_CHECK(cudaGetLastError());
        // BiLSTM(256): input projections for all steps, then the recurrence
        // tensor-core mode: xproj_fused.cu -> lstm_fused.cu, which share the row-tiled xp layout
        const bool fused_lstm = tc && net->lstm_in_fused[0] && net->lstm_in_fused[1] && net->lstm_rec_fused[0] && net->lstm_rec_fused[1];
        if (fused_lstm) {
            // both directions in one launch (xproj_fused.cu)
            if ((rc = mmla_launch_xproj_fused(seq, net->lstm_in_fused[0], net->lstm_in_fused[1], net->lstm_in[0].b,
                                              net->lstm_in[1].b, xp[0], xp[1], B, T, st)))
                return rc;
        } else {
            for (int d = 0; d < 2; ++d)
                if ((rc = launch_conv(net->lstm_in[d], seq, 0, B * T, 1, 1, nullptr, ACT_NONE, nullptr, 0, xp[d], st, tc))) return rc;
        }
        if (fused_lstm) {
            // one persistent launch: both directions, all time steps (lstm_fused.cu)
            const bool l16 = f16 && net->lstm_rec_f16[0] && net->lstm_rec_f16[1];     // fp16 mode: h and U as halves
            if ((rc = mmla_launch_lstm_fused(xp[0], xp[1], l16 ? net->lstm_rec_f16[0] : net->lstm_rec_fused[0],
                                             l16 ? net->lstm_rec_f16[1] : net->lstm_rec_fused[1], hdir[0], hdir[1], cst,
                                             cst + 3 * Bp * 256, B, T, st, l16 ? 1 : 0)))
                return rc;
        }
        for (int d = 0; d < 2 && !fused_lstm; ++d) {
            float* h = hdir[d];                            // updated in place: the recurrent GEMM of a
            for (int s = 0; s < T; ++s) {                  // step finishes before its gate kernel writes h
                const int t = d == 0 ? s : T - 1 - s;     // backward layer walks t = T-1 .. 0
                const float* xpt = xp[d] + static_cast<long long>(t) * 1024;
                if (s == 0) {                              // h0 = 0: pre-activations are xp[:, t, :]
                    lstm_gates_kernel<<<ew_grid(B * 256), 256, 0, st>>>(xpt, static_cast<long long>(T) * 1024, cst, h, B, 256, 1);
                    mmla_count_launch("lstm_gates_kernel", st);
                } else {
                    if ((rc = launch_conv(net->lstm_rec[d], h, 0, B, 1, 1, nullptr, ACT_NONE, xpt,
                                          static_cast<long long>(T) * 1024, z, st, tc)))
                        return rc;
                    lstm_gates_kernel<<<ew_grid(B * 256), 256, 0, st>>>(z, 1024, cst, h, B, 256, 0);
                    mmla_count_launch("lstm_gates_kernel", st);
                }
                MMLA_CUDA_CHECK(cudaGetLastError());
            }
        }
        if (embed) {
            // the trunk's output = layers[-2] of the Keras model: [fwd h | bwd h] (speaker_identification.py:403)
            embed_concat_kernel<<<ew_grid(B * 512), 256, 0, st>>>(hdir[0], hdir[1], B, embed + b0 * 512);
            mmla_count_launch("embed_concat_kernel", st);
            MMLA_CUDA_CHECK(cudaGetLastError());
            if (!prob) continue;
        }
        const int warps = 8;
        long long hgrid = (B + warps - 1) / warps;
        if (hgrid > 8LL * mmla_num_sms()) hgrid = 8LL * mmla_num_sms();
        head_kernel<<<static_cast<unsigned>(hgrid), warps * 32, warps * 512 * sizeof(float), st>>>(
            hdir[0], hdir[1], net->dense_k, net->dense_b, 0.3f, ov ? 1 : 0, net->n_classes, net->head == MMLA_HEAD_SIGMOID, B,
            prob + b0 * net->n_classes, labels ? labels + b0 : nullptr);
        mmla_count_launch("head_kernel", st);
        MMLA_CUDA_CHECK(cudaGetLastError());
    }
    return MMLA_OK;
}

// conv_slab.cu
bool mmla_conv_slab_eligible(const ConvArgs& a);
int mmla_launch_conv_slab(const ConvArgs& a, const float* wg, cudaStream_t st);

EXPORT int mmla_debug_conv2d(const float* x, const float* w_host, const float* bias, const float* pre_scale, const float* pre_shift,
                             int32_t pre_act, const float* res, float* y, int64_t B, int32_t H, int32_t W, int32_t Cin, int32_t N,
                             int32_t kh, int32_t kw, int32_t kernel, void* stream) {
    MMLA_REQUIRE(x && w_host && bias && y && B >= 0 && H > 0 && W > 0 && Cin > 0 && N > 0 && kh > 0 && kw > 0, MMLA_EINVAL,
                 "debug_conv2d: bad argument");
    MMLA_REQUIRE(kernel >= 0 && kernel <= 2, MMLA_EINVAL, "debug_conv2d: unknown kernel %d", kernel);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int K = kh * kw * Cin;
    const bool tc = kernel != 0;
    MMLA_REQUIRE(!tc || mmla_tc_ntile(N) != 0, MMLA_EUNSUP, "debug_conv2d: N=%d is not eligible for the tensor-core kernels", N);
    std::vector<float> host(tc ? mmla_tc_arranged_floats(K, N) : static_cast<long long>(K) * N);
    if (tc) mmla_tc_arrange_weights(w_host, K, N, host.data());
    else memcpy(host.data(), w_host, host.size() * sizeof(float));
    float* wdev = nullptr;
    MMLA_CUDA_CHECK(cudaMalloc(&wdev, host.size() * sizeof(float)));
    int rc = MMLA_OK;
    if (cudaMemcpyAsync(wdev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = MMLA_ECUDA;
    if (rc == MMLA_OK) {
        ConvW c;
        c.kh = kh; c.kw = kw; c.cin = Cin; c.cout = N; c.stride = 1;
        c.k = tc ? nullptr : wdev; c.b = bias; c.k_tc = tc ? wdev : nullptr;
        BnW bn;
        bn.scale = pre_scale; bn.shift = pre_shift;
        if (kernel == 2) {
            ConvArgs a;
            memset(&a, 0, sizeof(a));
            a.x = x; a.bias = bias; a.pre_scale = pre_scale; a.pre_shift = pre_shift; a.pre_act = pre_act;
            a.res = res; a.res_row_stride = N; a.y = y; a.H = H; a.W = W; a.Cin = Cin; a.Ho = H; a.Wo = W; a.N = N; a.K = K;
            a.kh = kh; a.kw = kw; a.stride = 1; a.pad_t = same_pad_before(H, kh, 1); a.pad_l = same_pad_before(W, kw, 1);
            a.M = B * H * W;
            if (!mmla_conv_slab_eligible(a)) {
                mmla_set_error("debug_conv2d: layer is not eligible for conv_slab_kernel");
                rc = MMLA_EUNSUP;
            } else if (a.M > 0) {
                rc = mmla_launch_conv_slab(a, wdev, st);
            }
        } else {
            // kernel 1 must be the gather kernel even where the slab kernel is eligible
            const char* old = getenv("MMLA_CONV_SLAB");
            const std::string keep = old ? old : "";
            if (kernel == 1) setenv("MMLA_CONV_SLAB", "0", 1);
            rc = launch_conv(c, x, 0, B, H, W, pre_scale ? &bn : nullptr, pre_act, res, N, y, st, tc);
            if (kernel == 1) { if (old) setenv("MMLA_CONV_SLAB", keep.c_str(), 1); else unsetenv("MMLA_CONV_SLAB"); }
        }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == MMLA_OK) {
        mmla_set_error("debug_conv2d: %s", cudaGetErrorString(cudaGetLastError()));
        rc = MMLA_ECUDA;
    }
    cudaFree(wdev);
    return rc;
}

EXPORT int mmla_debug_resblock2d(const float* x, const float* w1_host, const float* b1, const float* bn1_scale, const float* bn1_shift,
                                 const float* w2_host, const float* b2, const float* bn2_scale, const float* bn2_shift, const float* res,
                                 float* y, int64_t B, int32_t H, int32_t W, int32_t Cin, int32_t C, int32_t hpool, void* stream) {
    MMLA_REQUIRE(x && w1_host && b1 && bn1_scale && bn1_shift && w2_host && b2 && bn2_scale && bn2_shift && y && B >= 0 && H > 0 &&
                     W > 0 && Cin > 0 && C > 0,
                 MMLA_EINVAL, "debug_resblock2d: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MMLA_REQUIRE(mmla_resblock2d_eligible(H, W, Cin, C, 3, 3, 4, 1, ACT_ELU), MMLA_EUNSUP,
                 "debug_resblock2d: block is not eligible for resblock2d_fused_kernel");
    const int K1 = 9 * Cin, K2 = 4 * C;
    const long long n1 = mmla_tc_arranged_floats(K1, C), n2 = mmla_tc_arranged_floats(K2, C);
    const bool pair = mmla_rb_pair_wanted(Cin, C);
    const char* e16 = getenv("MMLA_RB_F16");                    // "1": the fp16-operand form of the kernel (MMLA_PRECISION_F16)
    const bool f16 = e16 && e16[0] == '1';
    const long long h1 = mmla_rb_f16_arranged_halves(K1, C), h2 = mmla_rb_f16_arranged_halves(K2, C);
    const long long f16_floats = f16 ? (h1 + h2 + 1) / 2 + 8 : 0;
    std::vector<float> host((n1 + n2) * (pair ? 2 : 1) + f16_floats);
    mmla_tc_arrange_weights(w1_host, K1, C, host.data());
    mmla_tc_arrange_weights(w2_host, K2, C, host.data() + n1);
    if (pair) {
        mmla_rb_arrange_weights_pair(w1_host, K1, C, host.data() + n1 + n2);
        mmla_rb_arrange_weights_pair(w2_host, K2, C, host.data() + 2 * n1 + n2);
    }
    long long hoff = ((n1 + n2) * (pair ? 2 : 1) + 3) / 4 * 4;   // 16-byte aligned start of the fp16 chunks (in floats)
    if (f16) {
        MMLA_REQUIRE(h1 % 8 == 0, MMLA_EUNSUP, "debug_resblock2d: odd fp16 chunk size");
        uint16_t* hp = reinterpret_cast<uint16_t*>(host.data() + hoff);
        mmla_rb_arrange_weights_f16(w1_host, K1, C, hp);
        mmla_rb_arrange_weights_f16(w2_host, K2, C, hp + h1);
    }
    float* wdev = nullptr;
    MMLA_CUDA_CHECK(cudaMalloc(&wdev, host.size() * sizeof(float)));
    int rc = MMLA_OK;
    if (cudaMemcpyAsync(wdev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = MMLA_ECUDA;
    if (rc == MMLA_OK)
        rc = mmla_launch_resblock2d_fused(x, y, B, H, W, Cin, C, bn1_scale, bn1_shift, wdev, b1, bn2_scale, bn2_shift, wdev + n1, b2, res, C, st,
                                          nullptr, 0, nullptr, nullptr, hpool, pair ? wdev + n1 + n2 : nullptr,
                                          pair ? wdev + 2 * n1 + n2 : nullptr,
                                          f16 ? static_cast<const void*>(reinterpret_cast<const uint16_t*>(wdev + hoff)) : nullptr,
                                          f16 ? static_cast<const void*>(reinterpret_cast<const uint16_t*>(wdev + hoff) + h1) : nullptr);
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == MMLA_OK) {
        mmla_set_error("debug_resblock2d: %s", cudaGetErrorString(cudaGetLastError()));
        rc = MMLA_ECUDA;
    }
    cudaFree(wdev);
    return rc;
}

EXPORT int mmla_net_set_precision(MmlaNet* net, int32_t mode) {
    MMLA_REQUIRE(net != nullptr, MMLA_EINVAL, "net_set_precision: null net");
    MMLA_REQUIRE(mode == MMLA_PRECISION_FP32 || mode == MMLA_PRECISION_TF32 || mode == MMLA_PRECISION_F16, MMLA_EINVAL,
                 "net_set_precision: unknown mode %d", mode);

    net->precision = mode;
    return MMLA_OK;
}
