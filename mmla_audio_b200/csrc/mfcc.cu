// Fused python_speech_features-style MFCC (+ delta, delta-delta, pad-to-N) for sm_100a.
//
// One persistent CTA (8 warps) walks "units" = (clip, tile of <=256 output frames).  Per unit it
// streams the clip's int16 PCM through shared memory in 16-frame batches with the TMA bulk-copy
// engine (cp.async.bulk + mbarrier, double buffered), and for every batch runs
//   P0  int16 -> float + pre-emphasis                      (psf.sigproc.preemphasis)
//   P1  512-point FFT of TWO real frames per warp packed as one complex transform
//       (radix-8 x 8 x 8, registers + two conflict-free shared-memory exchanges),
//       power spectrum 1/512 |X|^2 and frame energy     (psf.sigproc.framesig / powspec)
//   P2  sparse triangular mel filterbank + log            (psf.base.get_filterbanks / fbank)
//   P3  DCT-II (ortho) + lifter + c0 := ln(energy)        (psf.base.mfcc)
// into a per-tile cepstra buffer in shared memory, then an epilogue computes the reference's
// delta(feat,2) twice (SpeakerIdentification/scripts/speaker_identification.py:141-151), pads
// to `pad_frames` rows (:391-395) and writes coalesced rows.  Frames are never materialised in
// HBM: traffic = int16 PCM in + feature rows out.
//
// Replaces: mfcc(sig, rate, winlen=0.025, winstep=0.01, nfft=512) at
//   SpeakerIdentification/scripts/speaker_identification.py:89,285,341,386
//   SpeakerIdentification/scripts/speaker_identification_post_processing.py:256
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kBatch = 16;            // frames per batch = 2 per warp
constexpr int kNfft = 512;
constexpr int kBins = 257;
constexpr int kE1Stride = 72;         // complex stride of exchange-1 rows  [k0][m]
constexpr int kE2Stride = 66;         // complex stride of exchange-2 rows  [n0][q]
constexpr int kScratchC = 8 * kE1Stride;   // complex elements of FFT scratch per warp
constexpr int kPsStride = 258;        // float stride of one frame's power spectrum
constexpr int kMaxFilt = 64;
constexpr int kMaxCep = 16;          // smem row stride of cepstra; numcep itself is limited to 14
constexpr int kMaxNnz = 896;          // zero-padded filter halves
constexpr int kTileOut = 256;         // output frames per tile
constexpr int kHalo = 4;              // delta-delta context
constexpr int kTileFrames = kTileOut + 2 * kHalo;
constexpr int kMaxBatchSamples = 2816;     // (kBatch-1)*step + frame_len, padded
constexpr int kPcmBufSamples = kMaxBatchSamples + 16;
constexpr int kYpre = kPcmBufSamples;      // pre-emphasised floats, same indexing as the PCM buffer
constexpr float kEps = 2.220446049250313e-16f;

struct MfccTables {
    int4 fmeta[kMaxFilt];             // start bin, split point a, iterations per half, weight offset
    int warp_cnt[kWarps];
    unsigned char warp_list[kWarps][kMaxFilt];
    float fbw[kMaxNnz];               // per filter: [half 0 | half 1], each `iters` long, zero padded
    float dct[kMaxFilt][2][8];        // [filter][half][7 cepstra (+pad)]: half 0 = c0..c6, half 1 = c7..c13
    float window[kNfft];
};

struct MfccKernelParams {
    const int16_t* pcm;
    const int64_t* clip_off;          // device, or null (uniform)
    const int32_t* clip_len_arr;      // device, or null (uniform)
    const int2* units;                // device (clip, tile), or null (uniform arithmetic)
    const MfccTables* tables;
    float* out;
    long long n_units;
    long long clip_stride;
    long long out_clip_stride;
    int out_row_stride;               // floats between output rows (>= dim; extra columns are written as zeros)
    int tiles_per_clip;
    int clip_len;
    int frame_len, frame_step, nfilt, numcep;
    int append_energy, with_deltas, pad_frames, windowed;
    float preemph;
};

struct Smem {
    float2 scratch[kWarps][kScratchC];            // FFT exchanges; reused as delta tile
    float pspec[kBatch][kPsStride];
    alignas(16) float ypre[kYpre];                // ypre[i] = y at PCM-buffer sample i
    float feat[kTileFrames * kMaxCep];            // stride = numcep
    float cpart[kWarps][kBatch][kMaxCep];         // per-warp partial cepstra of the batch
    float energy[kBatch];
    alignas(16) int16_t pcm[2][kPcmBufSamples];
    alignas(8) uint64_t full_bar[2];
    MfccTables tab;
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // * (-i)

// 8-point DFT, natural order in and out (radix-2 DIF, bit reversal resolved by naming).
__device__ __forceinline__ void fft8(float2 (&a)[8]) {
    const float r = 0.70710678118654752440f;
    float2 b0 = cadd(a[0], a[4]), b4 = csub(a[0], a[4]);
    float2 b1 = cadd(a[1], a[5]), t5 = csub(a[1], a[5]);
    float2 b2 = cadd(a[2], a[6]), b6 = mul_mi(csub(a[2], a[6]));
    float2 b3 = cadd(a[3], a[7]), t7 = csub(a[3], a[7]);
    float2 b5 = make_float2((t5.x + t5.y) * r, (t5.y - t5.x) * r);     // * W8^1
    float2 b7 = make_float2((t7.y - t7.x) * r, -(t7.x + t7.y) * r);    // * W8^3
    float2 c0 = cadd(b0, b2), c2 = csub(b0, b2);
    float2 c1 = cadd(b1, b3), c3 = mul_mi(csub(b1, b3));
    float2 c4 = cadd(b4, b6), c6 = csub(b4, b6);
    float2 c5 = cadd(b5, b7), c7 = mul_mi(csub(b5, b7));
    a[0] = cadd(c0, c1); a[4] = csub(c0, c1);
    a[2] = cadd(c2, c3); a[6] = csub(c2, c3);
    a[1] = cadd(c4, c5); a[5] = csub(c4, c5);
    a[3] = cadd(c6, c7); a[7] = csub(c6, c7);
}

__device__ __forceinline__ int psf_frames(int len, int frame_len, int step) {
    return len <= frame_len ? 1 : 1 + (len - frame_len + step - 1) / step;
}

struct Unit {
    long long clip_off;   // first sample of the clip in pcm[]
    long long clip;
    int len;              // samples in the clip
    int T;                // psf frame count
    int n_real;           // real rows to emit (T, or min(T, pad_frames))
    int lo, hi;           // output frame range of this tile
    int c0, c1;           // frame range to compute (with delta halo)
    bool last_tile;
};

__device__ __forceinline__ Unit decode_unit(const MfccKernelParams& p, long long u) {
    Unit un;
    int tile;
    if (p.units) {
        int2 ct = p.units[u];
        un.clip = ct.x;
        tile = ct.y;
    } else {
        un.clip = u / p.tiles_per_clip;
        tile = static_cast<int>(u - un.clip * p.tiles_per_clip);
    }
    un.clip_off = p.clip_off ? p.clip_off[un.clip] : un.clip * p.clip_stride;
    un.len = p.clip_len_arr ? p.clip_len_arr[un.clip] : p.clip_len;
    un.T = psf_frames(un.len, p.frame_len, p.frame_step);
    un.n_real = p.pad_frames > 0 ? min(un.T, p.pad_frames) : un.T;
    un.lo = tile * kTileOut;
    un.hi = min(un.lo + kTileOut, un.n_real);
    const int h = p.with_deltas ? kHalo : 0;
    un.c0 = max(un.lo - h, 0);
    un.c1 = min(un.hi + h, un.T);
    un.last_tile = (un.hi >= un.n_real);
    return un;
}

// Issue the TMA bulk load of the PCM a batch needs.  Returns nothing; all threads can recompute
// the smem placement with batch_pcm_base().
__device__ __forceinline__ long long batch_pcm_base(const Unit& un, int frame0, int step) {
    long long first = un.clip_off + max(frame0 * step - 1, 0);
    return first & ~7LL;                                   // 16-byte aligned sample index
}

__device__ __forceinline__ void issue_batch_load(const MfccKernelParams& p, Smem& s, const Unit& un,
                                                 int frame0, int buf) {
    const long long a0 = batch_pcm_base(un, frame0, p.frame_step);
    long long need_end = un.clip_off + min(static_cast<long long>(un.len),
                                           static_cast<long long>(frame0 + kBatch - 1) * p.frame_step +
                                               p.frame_len);
    long long a1 = (need_end + 7) & ~7LL;
    uint32_t bytes = a1 > a0 ? static_cast<uint32_t>((a1 - a0) * 2) : 0u;
    if (bytes == 0) bytes = 16;                            // degenerate (empty clip): harmless 16 B
    fence_proxy_async_smem();                              // prior generic reads of this buffer
    mbar_arrive_expect_tx(&s.full_bar[buf], bytes);
    tma_bulk_g2s(&s.pcm[buf][0], p.pcm + a0, bytes, &s.full_bar[buf]);
}

__global__ void __launch_bounds__(kThreads, 2) mfcc_fused_kernel(const __grid_constant__ MfccKernelParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    // ---- one-time setup: tables to smem, barriers, per-lane twiddles ----------------------------
    {
        const int4* src = reinterpret_cast<const int4*>(p.tables);
        int4* dst = reinterpret_cast<int4*>(&s.tab);
        for (int i = tid; i < static_cast<int>(sizeof(MfccTables) / 16); i += kThreads) dst[i] = src[i];
    }
    for (int i = tid; i < kBatch * kPsStride; i += kThreads) (&s.pspec[0][0])[i] = 0.f;   // finite everywhere
    if (tid == 0) {
        mbar_init(&s.full_bar[0], 1);
        mbar_init(&s.full_bar[1], 1);
        mbar_fence_init();
    }
    // stage-1 twiddles W512^(m*k0), m = lane (+32); stage-2 twiddles W64^(n0*k1), n0 = lane&7
    float2 tw1a[7], tw1b[7], tw2[7];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        float sn, cs;
        sincospif(-2.0f * static_cast<float>((lane * k) & 511) / 512.0f, &sn, &cs);
        tw1a[k - 1] = make_float2(cs, sn);
        sincospif(-2.0f * static_cast<float>(((lane + 32) * k) & 511) / 512.0f, &sn, &cs);
        tw1b[k - 1] = make_float2(cs, sn);
        sincospif(-2.0f * static_cast<float>(((lane & 7) * k) & 63) / 64.0f, &sn, &cs);
        tw2[k - 1] = make_float2(cs, sn);
    }
    __syncthreads();

    const int ncep = p.numcep;
    const int dim = p.with_deltas ? 3 * ncep : ncep;
    const int step = p.frame_step;
    const int flen = p.frame_len;
    uint32_t phase[2] = {0u, 0u};

    for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const Unit un = decode_unit(p, u);
        const int nframes = un.c1 - un.c0;
        const int nb = (nframes + kBatch - 1) / kBatch;
        if (tid == 0 && nb > 0) issue_batch_load(p, s, un, un.c0, 0);

        for (int b = 0; b < nb; ++b) {
            const int buf = b & 1;
            const int f0 = un.c0 + b * kBatch;             // first frame of the batch
            // prefetch the next batch of this unit into the other buffer (its last readers
            // finished before the S0 barrier of the previous batch)
            if (tid == 0 && b + 1 < nb) issue_batch_load(p, s, un, f0 + kBatch, buf ^ 1);
            mbar_wait(&s.full_bar[buf], phase[buf]);
            phase[buf] ^= 1u;

            // ---- P0: int16 -> float + pre-emphasis, 8 samples per thread -------------------------
            // ypre uses the PCM buffer's own (16-byte aligned) indexing, so loads and stores are
            // 128-bit; `yoff` is where the batch's first sample sits.
            const long long a0 = batch_pcm_base(un, f0, step);
            const int shift = static_cast<int>(un.clip_off - a0);       // buffer index of clip sample 0
            const int yoff = shift + f0 * step;                         // buffer index of y[f0*step]
            {
                const uint4* px4 = reinterpret_cast<const uint4*>(&s.pcm[buf][0]);
                const unsigned short* px = reinterpret_cast<const unsigned short*>(&s.pcm[buf][0]);
                const int ngroups = (yoff + (kBatch - 1) * step + flen + 7) >> 3;
                const float pre = p.preemph;
                for (int g = tid; g < ngroups; g += kThreads) {
                    const uint4 w = px4[g];
                    const uint32_t prev16 = g > 0 ? px[8 * g - 1] : 0u;
                    const int s0 = 8 * g - shift;                       // clip sample index of element 0
                    float x[9];
                    x[0] = s16_bits_to_float(prev16);
                    x[1] = s16_bits_to_float(w.x & 0xffffu); x[2] = s16_bits_to_float(w.x >> 16);
                    x[3] = s16_bits_to_float(w.y & 0xffffu); x[4] = s16_bits_to_float(w.y >> 16);
                    x[5] = s16_bits_to_float(w.z & 0xffffu); x[6] = s16_bits_to_float(w.z >> 16);
                    x[7] = s16_bits_to_float(w.w & 0xffffu); x[8] = s16_bits_to_float(w.w >> 16);
                    float y[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int si = s0 + j;
                        const float xp = si > 0 ? x[j] : 0.f;           // y[0] = x[0]
                        y[j] = (si >= 0 && si < un.len) ? fmaf(-pre, xp, x[j + 1]) : 0.f;   // zero padded tail
                    }
                    float4* dst = reinterpret_cast<float4*>(&s.ypre[8 * g]);
                    dst[0] = make_float4(y[0], y[1], y[2], y[3]);
                    dst[1] = make_float4(y[4], y[5], y[6], y[7]);
                }
            }
            __syncthreads();                                            // S0

            // ---- P1: packed two-frame 512-point FFT per warp ------------------------------------
            {
                const int fa = f0 + 2 * warp;                           // frame A (B = A+1)
                if (fa < un.c1) {
                    const float* ya = &s.ypre[yoff + (2 * warp) * step];
                    const float* yb = ya + step;
                    float2* sc = &s.scratch[warp][0];
                    float2 v0[8], v1[8];
                    // stage 1: butterflies m = lane, lane+32 over n2 (stride 64)
                    if (flen == 400 && !p.windowed) {
                        // the reference configuration: samples 0..383 always present, 384..399 only
                        // for m < 16, everything above is the FFT's zero padding
#pragma unroll
                        for (int n2 = 0; n2 < 6; ++n2) {
                            v0[n2] = make_float2(ya[64 * n2 + lane], yb[64 * n2 + lane]);
                            v1[n2] = make_float2(ya[64 * n2 + 32 + lane], yb[64 * n2 + 32 + lane]);
                        }
                        v0[6] = lane < 16 ? make_float2(ya[384 + lane], yb[384 + lane]) : make_float2(0.f, 0.f);
                        v0[7] = v1[6] = v1[7] = make_float2(0.f, 0.f);
                    } else {
#pragma unroll
                        for (int n2 = 0; n2 < 8; ++n2) {
                            const int n_a = 64 * n2 + lane;
                            const int n_b = n_a + 32;
                            float2 za = make_float2(0.f, 0.f), zb = make_float2(0.f, 0.f);
                            if (n_a < flen) {
                                const float wv = p.windowed ? s.tab.window[n_a] : 1.f;
                                za = make_float2(ya[n_a] * wv, yb[n_a] * wv);
                            }
                            if (n_b < flen) {
                                const float wv = p.windowed ? s.tab.window[n_b] : 1.f;
                                zb = make_float2(ya[n_b] * wv, yb[n_b] * wv);
                            }
                            v0[n2] = za;
                            v1[n2] = zb;
                        }
                    }
                    fft8(v0);
                    fft8(v1);
                    sc[lane] = v0[0];
                    sc[lane + 32] = v1[0];
#pragma unroll
                    for (int k = 1; k < 8; ++k) {
                        sc[k * kE1Stride + lane] = cmul(v0[k], tw1a[k - 1]);
                        sc[k * kE1Stride + lane + 32] = cmul(v1[k], tw1b[k - 1]);
                    }
                    __syncwarp();
                    // stage 2: tasks t = lane, lane+32 -> (k0 = t>>3, n0 = t&7) over n1
                    {
                        const int n0 = lane & 7;
                        const int k0a = lane >> 3, k0b = k0a + 4;
#pragma unroll
                        for (int n1 = 0; n1 < 8; ++n1) {
                            v0[n1] = sc[k0a * kE1Stride + 8 * n1 + n0];
                            v1[n1] = sc[k0b * kE1Stride + 8 * n1 + n0];
                        }
                        __syncwarp();
                        fft8(v0);
                        fft8(v1);
                        sc[n0 * kE2Stride + k0a] = v0[0];
                        sc[n0 * kE2Stride + k0b] = v1[0];
#pragma unroll
                        for (int k = 1; k < 8; ++k) {
                            sc[n0 * kE2Stride + k0a + 8 * k] = cmul(v0[k], tw2[k - 1]);
                            sc[n0 * kE2Stride + k0b + 8 * k] = cmul(v1[k], tw2[k - 1]);
                        }
                    }
                    __syncwarp();
                    // stage 3: tasks q = lane (P) and q' = 64-lane (Q); lane 0 takes q = 0 and 32
                    const int qp = lane;
                    const int qq = lane == 0 ? 32 : 64 - lane;
#pragma unroll
                    for (int n0 = 0; n0 < 8; ++n0) {
                        v0[n0] = sc[n0 * kE2Stride + qp];
                        v1[n0] = sc[n0 * kE2Stride + qq];
                    }
                    fft8(v0);                                           // v0[k2] = Z[qp + 64 k2]
                    fft8(v1);                                           // v1[k2] = Z[qq + 64 k2]
                    // split the packed transform: A = (Z[k]+conj Z[512-k])/2, B = (Z[k]-conj Z[512-k])/2i
                    const float sc_p = 0.25f / static_cast<float>(kNfft);
                    float* psa = &s.pspec[2 * warp][0];
                    float* psb = &s.pspec[2 * warp + 1][0];
                    float ea = 0.f, eb = 0.f;
                    const bool l0 = (lane == 0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        {   // bin k = qp + 64 j ; partner 512-k lives in Q[7-j] (lane 0: P[(8-j)&7])
                            const float2 z = v0[j];
                            const float2 other = v1[7 - j];
                            const float2 self = v0[(8 - j) & 7];
                            const float2 zc = make_float2(l0 ? self.x : other.x, l0 ? self.y : other.y);
                            const float ar = z.x + zc.x, ai = z.y - zc.y;
                            const float br = z.y + zc.y, bi = z.x - zc.x;
                            const float pa = (ar * ar + ai * ai) * sc_p;
                            const float pb = (br * br + bi * bi) * sc_p;
                            psa[qp + 64 * j] = pa;
                            psb[qp + 64 * j] = pb;
                            ea += pa;
                            eb += pb;
                        }
                        {   // bin k = qq + 64 j ; partner in P[7-j] (lane 0: Q[7-j])
                            const float2 z = v1[j];
                            const float2 other = v0[7 - j];
                            const float2 self = v1[7 - j];
                            const float2 zc = make_float2(l0 ? self.x : other.x, l0 ? self.y : other.y);
                            const float ar = z.x + zc.x, ai = z.y - zc.y;
                            const float br = z.y + zc.y, bi = z.x - zc.x;
                            const float pa = (ar * ar + ai * ai) * sc_p;
                            const float pb = (br * br + bi * bi) * sc_p;
                            psa[qq + 64 * j] = pa;
                            psb[qq + 64 * j] = pb;
                            ea += pa;
                            eb += pb;
                        }
                    }
                    if (l0) {                                           // Nyquist bin 256 = P[4], self-paired
                        const float2 z = v0[4];
                        const float pa = (4.f * z.x * z.x) * sc_p;
                        const float pb = (4.f * z.y * z.y) * sc_p;
                        psa[256] = pa;
                        psb[256] = pb;
                        ea += pa;
                        eb += pb;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        ea += __shfl_xor_sync(0xffffffffu, ea, o);
                        eb += __shfl_xor_sync(0xffffffffu, eb, o);
                    }
                    if (l0) {
                        s.energy[2 * warp] = ea;
                        s.energy[2 * warp + 1] = eb;
                    }
                }
            }
            __syncthreads();                                            // S1

            // ---- P2: mel filterbank + log + this warp's share of the DCT ------------------------
            // lane = (frame f, half h): the two halves split each filter's bins, then split the
            // cepstra (h=0: c0..c6, h=1: c7..c13) of the DCT-II partial sums.
            {
                const int f = lane & 15;
                const int h = lane >> 4;
                const float* ps = &s.pspec[f][0];
                const int cnt = s.tab.warp_cnt[warp];
                float cp[7];
#pragma unroll
                for (int c = 0; c < 7; ++c) cp[c] = 0.f;
                for (int li = 0; li < cnt; ++li) {
                    const int j = s.tab.warp_list[warp][li];
                    const int4 m = s.tab.fmeta[j];                      // start, a, iters, woff
                    const float* w = &s.tab.fbw[m.w + h * m.z];
                    const float* x = ps + m.x + h * m.y;
                    float acc0 = 0.f, acc1 = 0.f;
                    for (int i = 0; i < m.z; i += 4) {                  // zero-padded: no bounds tests
                        const float4 wv = *reinterpret_cast<const float4*>(w + i);
                        acc0 = fmaf(wv.x, x[i], acc0);
                        acc1 = fmaf(wv.y, x[i + 1], acc1);
                        acc0 = fmaf(wv.z, x[i + 2], acc0);
                        acc1 = fmaf(wv.w, x[i + 3], acc1);
                    }
                    float acc = acc0 + acc1;
                    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                    const float lm = __logf(acc == 0.f ? kEps : acc);
                    const float4 d0 = *reinterpret_cast<const float4*>(&s.tab.dct[j][h][0]);
                    const float4 d1 = *reinterpret_cast<const float4*>(&s.tab.dct[j][h][4]);
                    cp[0] = fmaf(lm, d0.x, cp[0]); cp[1] = fmaf(lm, d0.y, cp[1]);
                    cp[2] = fmaf(lm, d0.z, cp[2]); cp[3] = fmaf(lm, d0.w, cp[3]);
                    cp[4] = fmaf(lm, d1.x, cp[4]); cp[5] = fmaf(lm, d1.y, cp[5]);
                    cp[6] = fmaf(lm, d1.z, cp[6]);
                }
                float* dst = &s.cpart[warp][f][7 * h];
#pragma unroll
                for (int c = 0; c < 7; ++c) dst[c] = cp[c];
            }
            __syncthreads();                                            // S2

            // ---- P3: sum the 8 warps' partial cepstra; c0 := ln(energy) -------------------------
            {
                const int f = tid >> 4, c = tid & 15;
                const int fr = f0 + f;
                if (c < ncep && fr < un.c1) {
                    float v = 0.f;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) v += s.cpart[w][f][c];
                    if (c == 0 && p.append_energy) {
                        const float e = s.energy[f];
                        v = logf(e == 0.f ? kEps : e);
                    }
                    s.feat[(fr - un.c0) * ncep + c] = v;
                }
            }
            // no barrier needed here: the next batch's P0 only writes ypre (last read before S1), and
            // cpart / energy are next written after the next S0 / S1
        }
        __syncthreads();

        // ---- epilogue: (delta, delta-delta), padding rows, coalesced store ---------------------
        float* out_clip = p.out + un.clip * p.out_clip_stride;
        const int rs = p.out_row_stride;
        if (!p.with_deltas) {
            const int n = (un.hi - un.lo) * rs;
            const float* src = &s.feat[(un.lo - un.c0) * ncep];
            float* dst = out_clip + static_cast<long long>(un.lo) * rs;
            for (int e = tid; e < n; e += kThreads) {
                const int r = e / rs, col = e - r * rs;
                dst[e] = col < ncep ? src[r * ncep + col] : 0.f;
            }
        } else {
            float* dt = reinterpret_cast<float*>(&s.scratch[0][0]);     // delta tile, same indexing as feat
            const int Tm1 = un.T - 1;
            const int d0 = max(un.lo - 2, 0), d1 = min(un.hi + 2, un.T);
            for (int e = tid; e < (d1 - d0) * ncep; e += kThreads) {
                const int t = d0 + e / ncep, c = e % ncep;
                const float xm2 = s.feat[(max(t - 2, 0) - un.c0) * ncep + c];
                const float xm1 = s.feat[(max(t - 1, 0) - un.c0) * ncep + c];
                const float xp1 = s.feat[(min(t + 1, Tm1) - un.c0) * ncep + c];
                const float xp2 = s.feat[(min(t + 2, Tm1) - un.c0) * ncep + c];
                dt[(t - un.c0) * ncep + c] = (-2.f * xm2 - xm1 + xp1 + 2.f * xp2) / 10.f;
            }
            __syncthreads();
            const int n = (un.hi - un.lo) * rs;
            float* dst = out_clip + static_cast<long long>(un.lo) * rs;
            for (int e = tid; e < n; e += kThreads) {
                const int t = un.lo + e / rs, col = e % rs;
                float v;
                if (col >= dim) {
                    v = 0.f;
                } else if (col < ncep) {
                    v = s.feat[(t - un.c0) * ncep + col];
                } else if (col < 2 * ncep) {
                    v = dt[(t - un.c0) * ncep + col - ncep];
                } else {
                    const int c = col - 2 * ncep;
                    const float xm2 = dt[(max(t - 2, 0) - un.c0) * ncep + c];
                    const float xm1 = dt[(max(t - 1, 0) - un.c0) * ncep + c];
                    const float xp1 = dt[(min(t + 1, Tm1) - un.c0) * ncep + c];
                    const float xp2 = dt[(min(t + 2, Tm1) - un.c0) * ncep + c];
                    v = (-2.f * xm2 - xm1 + xp1 + 2.f * xp2) / 10.f;
                }
                dst[e] = v;
            }
        }
        if (un.last_tile && p.pad_frames > un.n_real) {
            const int n = (p.pad_frames - un.n_real) * p.out_row_stride;
            float* dst = out_clip + static_cast<long long>(un.n_real) * p.out_row_stride;
            for (int e = tid; e < n; e += kThreads) dst[e] = 0.f;
        }
        __syncthreads();                                                // feat / scratch reuse
    }
}

// -------------------------------------------------------------------------------------------------
// host side: tables (float64 maths, as python_speech_features does) and launch
// -------------------------------------------------------------------------------------------------
struct TableKey {
    int sr, flen, nfft, nfilt, ncep, lifter, window;
    float lo, hi;
    int dev;
    bool operator<(const TableKey& o) const { return memcmp(this, &o, sizeof(TableKey)) < 0; }
};

std::mutex g_tab_mu;
std::map<TableKey, MfccTables*> g_tab_cache;

int build_tables(const MmlaMfccParams& p, MfccTables& t) {
    memset(&t, 0, sizeof(t));
    const int nb = p.nfft / 2 + 1;
    // psf.base.get_filterbanks (HTK mel, floor((nfft+1)*hz/sr) edges)
    const double highfreq = p.highfreq > 0 ? p.highfreq : p.samplerate / 2.0;
    const double lowmel = 2595.0 * log10(1.0 + p.lowfreq / 700.0);
    const double highmel = 2595.0 * log10(1.0 + highfreq / 700.0);
    std::vector<double> bin(p.nfilt + 2);
    for (int i = 0; i < p.nfilt + 2; ++i) {
        // numpy.linspace: start + i*step with step = (stop-start)/(n-1); last point exact
        double mel = (i == p.nfilt + 1) ? highmel : lowmel + i * ((highmel - lowmel) / (p.nfilt + 1));
        double hz = 700.0 * (pow(10.0, mel / 2595.0) - 1.0);
        bin[i] = floor((p.nfft + 1) * hz / p.samplerate);
    }
    std::vector<double> row(nb);
    int nnz_total = 0;
    std::vector<int> cost(p.nfilt);
    for (int j = 0; j < p.nfilt; ++j) {
        std::fill(row.begin(), row.end(), 0.0);
        for (int i = static_cast<int>(bin[j]); i < static_cast<int>(bin[j + 1]); ++i)
            if (i >= 0 && i < nb) row[i] = (i - bin[j]) / (bin[j + 1] - bin[j]);
        for (int i = static_cast<int>(bin[j + 1]); i < static_cast<int>(bin[j + 2]); ++i)
            if (i >= 0 && i < nb) row[i] = (bin[j + 2] - i) / (bin[j + 2] - bin[j + 1]);
        int first = -1, last = -1;
        for (int i = 0; i < nb; ++i)
            if (row[i] != 0.0) {
                if (first < 0) first = i;
                last = i;
            }
        const int len = first < 0 ? 0 : last - first + 1;
        int a = len >= 2 ? ((len / 2) | 1) : len;          // odd split point (bank-friendly)
        if (a > len) a = len;
        int iters = (len - a > a ? len - a : a);
        iters = (iters + 3) & ~3;                          // float4 weight loads, no bounds tests
        if (nnz_total + 2 * iters > kMaxNnz) {
            mmla_set_error("mfcc: filterbank needs more than %d padded weights", kMaxNnz);
            return MMLA_EUNSUP;
        }
        const int start = first < 0 ? 0 : first;
        t.fmeta[j] = make_int4(start, a, iters, nnz_total);
        for (int i = 0; i < a; ++i) t.fbw[nnz_total + i] = static_cast<float>(row[start + i]);
        for (int i = a; i < len; ++i) t.fbw[nnz_total + iters + (i - a)] = static_cast<float>(row[start + i]);
        nnz_total += 2 * iters;
        cost[j] = (iters / 4) * 6 + 28;                    // loop trips + fixed per-filter work
    }
    // longest-processing-time partition of filters over the 8 warps
    std::vector<int> order(p.nfilt);
    for (int j = 0; j < p.nfilt; ++j) order[j] = j;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
    int load[kWarps] = {0};
    for (int j : order) {
        int w = 0;
        for (int k = 1; k < kWarps; ++k)
            if (load[k] < load[w]) w = k;
        t.warp_list[w][t.warp_cnt[w]++] = static_cast<unsigned char>(j);
        load[w] += cost[j];
    }
    // DCT-II ortho (scipy.fftpack.dct norm='ortho') with the psf lifter folded in, laid out per
    // filter as [half][7]: half 0 holds c0..c6, half 1 holds c7..c13
    const double PI = 3.14159265358979323846;
    for (int c = 0; c < p.numcep; ++c) {
        double scale = c == 0 ? sqrt(1.0 / p.nfilt) : sqrt(2.0 / p.nfilt);
        double lift = p.ceplifter > 0 ? 1.0 + (p.ceplifter / 2.0) * sin(PI * c / p.ceplifter) : 1.0;
        for (int j = 0; j < p.nfilt; ++j)
            t.dct[j][c / 7][c % 7] = static_cast<float>(lift * scale * cos(PI * c * (2 * j + 1) / (2.0 * p.nfilt)));
    }
    for (int n = 0; n < kNfft; ++n) {
        double w = 1.0;
        if (n >= p.frame_len) w = 0.0;
        else if (p.window == MMLA_WINDOW_HANN) w = 0.5 - 0.5 * cos(2.0 * PI * n / (p.frame_len - 1));
        else if (p.window == MMLA_WINDOW_HAMMING) w = 0.54 - 0.46 * cos(2.0 * PI * n / (p.frame_len - 1));
        t.window[n] = static_cast<float>(w);
    }
    return MMLA_OK;
}

int get_tables(const MmlaMfccParams& p, const MfccTables** out) {
    TableKey key;
    memset(&key, 0, sizeof(key));
    int dev = 0;
    MMLA_CUDA_CHECK(cudaGetDevice(&dev));
    key.sr = p.samplerate; key.flen = p.frame_len; key.nfft = p.nfft; key.nfilt = p.nfilt;
    key.ncep = p.numcep; key.lifter = p.ceplifter; key.window = p.window;
    key.lo = p.lowfreq; key.hi = p.highfreq; key.dev = dev;
    std::lock_guard<std::mutex> g(g_tab_mu);
    auto it = g_tab_cache.find(key);
    if (it != g_tab_cache.end()) {
        *out = it->second;
        return MMLA_OK;
    }
    MfccTables* host = new MfccTables;
    int rc = build_tables(p, *host);
    if (rc != MMLA_OK) {
        delete host;
        return rc;
    }
    MfccTables* devp = nullptr;
    cudaError_t e = cudaMalloc(&devp, sizeof(MfccTables));
    if (e == cudaSuccess) e = cudaMemcpy(devp, host, sizeof(MfccTables), cudaMemcpyHostToDevice);
    delete host;
    if (e != cudaSuccess) {
        mmla_set_error("mfcc tables upload failed: %s", cudaGetErrorString(e));
        return MMLA_ECUDA;
    }
    g_tab_cache[key] = devp;
    *out = devp;
    return MMLA_OK;
}

}  // namespace

// csrc/mfcc_tc.cu: tensor-core path for the reference parameterisation
int mmla_mfcc_tc_try(const int16_t* pcm, int64_t pcm_total, const int64_t* clip_off_host, const int32_t* clip_len_host,
                     int64_t n_clips, int32_t clip_len, int64_t clip_stride, const MmlaMfccParams& p, float* out,
                     int64_t out_clip_stride, int32_t out_row_stride, cudaStream_t st, float* dbg, long long* prof, int* handled);
static float* g_tc_dump = nullptr;
static long long* g_tc_prof = nullptr;
extern "C" __attribute__((visibility("default"))) void mmla_debug_mfcc_tc_dump(float* dev_buffer, long long* dev_stamps) {
    g_tc_dump = dev_buffer;
    g_tc_prof = dev_stamps;
}

extern "C" __attribute__((visibility("default"))) int32_t mmla_psf_num_frames(int64_t n, const MmlaMfccParams* p) {
    if (!p || p->frame_step <= 0) return -1;
    if (n <= p->frame_len) return 1;
    return static_cast<int32_t>(1 + (n - p->frame_len + p->frame_step - 1) / p->frame_step);
}

extern "C" __attribute__((visibility("default"))) int mmla_psf_mfcc(const int16_t* pcm, int64_t pcm_total,
                             const int64_t* clip_off_host, const int32_t* clip_len_host,
                             int64_t n_clips, int32_t clip_len, int64_t clip_stride,
                             const MmlaMfccParams* pp, float* out, int64_t out_clip_stride,
                             void* stream) {
    return mmla_psf_mfcc_rows(pcm, pcm_total, clip_off_host, clip_len_host, n_clips, clip_len, clip_stride, pp, out,
                              out_clip_stride, 0, stream);
}

extern "C" __attribute__((visibility("default"))) int mmla_psf_mfcc_rows(const int16_t* pcm, int64_t pcm_total,
                             const int64_t* clip_off_host, const int32_t* clip_len_host,
                             int64_t n_clips, int32_t clip_len, int64_t clip_stride,
                             const MmlaMfccParams* pp, float* out, int64_t out_clip_stride,
                             int32_t out_row_stride, void* stream) {
    MMLA_REQUIRE(pp != nullptr && pcm != nullptr && out != nullptr, MMLA_EINVAL, "mfcc: null argument");
    const MmlaMfccParams& p = *pp;
    {
        const int dim = p.numcep * (p.with_deltas ? 3 : 1);
        if (out_row_stride == 0) out_row_stride = dim;
        MMLA_REQUIRE(out_row_stride >= dim, MMLA_EINVAL, "mfcc: out_row_stride %d < row width %d", out_row_stride, dim);
    }
    MMLA_REQUIRE(n_clips >= 0, MMLA_EINVAL, "mfcc: negative n_clips");
    if (n_clips == 0) return MMLA_OK;
    MMLA_REQUIRE(p.nfft == kNfft, MMLA_EUNSUP, "mfcc: nfft=%d unsupported (the warp FFT is 512-point)", p.nfft);
    MMLA_REQUIRE(p.frame_len >= 1 && p.frame_len <= kNfft, MMLA_EUNSUP, "mfcc: frame_len=%d must be in [1,512]", p.frame_len);
    MMLA_REQUIRE(p.frame_step >= 1 && (kBatch - 1) * p.frame_step + p.frame_len <= kMaxBatchSamples, MMLA_EUNSUP,
                 "mfcc: frame_step=%d too large for the %d-sample batch buffer", p.frame_step, kMaxBatchSamples);
    MMLA_REQUIRE(p.nfilt >= 1 && p.nfilt <= kMaxFilt, MMLA_EUNSUP, "mfcc: nfilt=%d must be in [1,%d]", p.nfilt, kMaxFilt);
    MMLA_REQUIRE(p.numcep >= 1 && p.numcep <= 14 && p.numcep <= p.nfilt, MMLA_EUNSUP,
                 "mfcc: numcep=%d must be in [1,14] and <= nfilt", p.numcep);
    MMLA_REQUIRE(p.window >= 0 && p.window <= 2, MMLA_EINVAL, "mfcc: bad window id %d", p.window);
    MMLA_REQUIRE((reinterpret_cast<uintptr_t>(pcm) & 15) == 0, MMLA_EINVAL, "mfcc: pcm must be 16-byte aligned");
    MMLA_REQUIRE((clip_off_host == nullptr) == (clip_len_host == nullptr), MMLA_EINVAL,
                 "mfcc: clip_off_host and clip_len_host must both be given or both be NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    if (clip_off_host == nullptr) {
        MMLA_REQUIRE(clip_len >= 0 && clip_stride >= 0, MMLA_EINVAL, "mfcc: bad uniform clip geometry");
        MMLA_REQUIRE((n_clips - 1) * clip_stride + clip_len <= pcm_total, MMLA_EINVAL, "mfcc: clips exceed pcm_total_samples");
    } else {
        for (int64_t c = 0; c < n_clips; ++c)
            MMLA_REQUIRE(clip_len_host[c] >= 0 && clip_off_host[c] >= 0 && clip_off_host[c] + clip_len_host[c] <= pcm_total,
                         MMLA_EINVAL, "mfcc: clip %lld exceeds pcm_total_samples", static_cast<long long>(c));
    }
    {
        int handled = 0;
        const int trc = mmla_mfcc_tc_try(pcm, pcm_total, clip_off_host, clip_len_host, n_clips, clip_len, clip_stride, p, out,
                                         out_clip_stride, out_row_stride, st, g_tc_dump, g_tc_prof, &handled);
        if (trc != MMLA_OK) return trc;
        if (handled) return MMLA_OK;
    }

    const MfccTables* tab = nullptr;
    int rc = get_tables(p, &tab);
    if (rc != MMLA_OK) return rc;

    MfccKernelParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.pcm = pcm;
    kp.tables = tab;
    kp.out = out;
    kp.clip_stride = clip_stride;
    kp.out_clip_stride = out_clip_stride;
    kp.out_row_stride = out_row_stride;
    kp.clip_len = clip_len;
    kp.frame_len = p.frame_len; kp.frame_step = p.frame_step; kp.nfilt = p.nfilt; kp.numcep = p.numcep;
    kp.append_energy = p.append_energy; kp.with_deltas = p.with_deltas; kp.pad_frames = p.pad_frames;
    kp.windowed = p.window != MMLA_WINDOW_RECT;
    kp.preemph = p.preemph;

    void* dev_tmp = nullptr;
    if (clip_off_host == nullptr) {
        MMLA_REQUIRE(clip_len >= 0 && clip_stride >= 0, MMLA_EINVAL, "mfcc: bad uniform clip geometry");   // stride < len = overlapping windows
        MMLA_REQUIRE((n_clips - 1) * clip_stride + clip_len <= pcm_total, MMLA_EINVAL, "mfcc: clips exceed pcm_total_samples");
        const int T = mmla_psf_num_frames(clip_len, &p);
        const int n_real = p.pad_frames > 0 ? (T < p.pad_frames ? T : p.pad_frames) : T;
        kp.tiles_per_clip = (n_real + kTileOut - 1) / kTileOut;
        if (kp.tiles_per_clip < 1) kp.tiles_per_clip = 1;
        kp.n_units = n_clips * kp.tiles_per_clip;
    } else {
        // ragged clips: explicit (clip, tile) unit table + per-clip offsets/lengths on the device
        std::vector<int2> units;
        for (int64_t c = 0; c < n_clips; ++c) {
            MMLA_REQUIRE(clip_len_host[c] >= 0 && clip_off_host[c] >= 0 &&
                             clip_off_host[c] + clip_len_host[c] <= pcm_total,
                         MMLA_EINVAL, "mfcc: clip %lld exceeds pcm_total_samples", static_cast<long long>(c));
            const int T = mmla_psf_num_frames(clip_len_host[c], &p);
            const int n_real = p.pad_frames > 0 ? (T < p.pad_frames ? T : p.pad_frames) : T;
            int tiles = (n_real + kTileOut - 1) / kTileOut;
            if (tiles < 1) tiles = 1;
            for (int t = 0; t < tiles; ++t) units.push_back(make_int2(static_cast<int>(c), t));
        }
        const size_t b_units = units.size() * sizeof(int2);
        const size_t b_off = static_cast<size_t>(n_clips) * sizeof(int64_t);
        const size_t b_len = static_cast<size_t>(n_clips) * sizeof(int32_t);
        const size_t o_off = (b_units + 15) & ~static_cast<size_t>(15);
        const size_t o_len = o_off + ((b_off + 15) & ~static_cast<size_t>(15));
        MMLA_CUDA_CHECK(cudaMallocAsync(&dev_tmp, o_len + b_len, st));
        char* base = static_cast<char*>(dev_tmp);
        // pageable host memory: cudaMemcpyAsync stages it before returning, so `units` may die
        MMLA_CUDA_CHECK(cudaMemcpyAsync(base, units.data(), b_units, cudaMemcpyHostToDevice, st));
        MMLA_CUDA_CHECK(cudaMemcpyAsync(base + o_off, clip_off_host, b_off, cudaMemcpyHostToDevice, st));
        MMLA_CUDA_CHECK(cudaMemcpyAsync(base + o_len, clip_len_host, b_len, cudaMemcpyHostToDevice, st));
        kp.units = reinterpret_cast<const int2*>(base);
        kp.clip_off = reinterpret_cast<const int64_t*>(base + o_off);
        kp.clip_len_arr = reinterpret_cast<const int32_t*>(base + o_len);
        kp.n_units = static_cast<long long>(units.size());
    }

    const int sms = mmla_num_sms();
    MMLA_REQUIRE(sms > 0, MMLA_ECUDA, "mfcc: no CUDA device");
    static MmlaPerDeviceOnce attr_once;                          // cudaFuncSetAttribute is per device
    const bool attr_set = !attr_once.first();
    if (!attr_set) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(mfcc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(Smem))));
    }
    long long grid = 2LL * sms;
    if (grid > kp.n_units) grid = kp.n_units;
    mfcc_fused_kernel<<<static_cast<unsigned>(grid), kThreads, sizeof(Smem), st>>>(kp);
    mmla_count_launch("mfcc_fused_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    if (dev_tmp) MMLA_CUDA_CHECK(cudaFreeAsync(dev_tmp, st));
    return MMLA_OK;
}
