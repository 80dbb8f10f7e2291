// Fused overlap-detection feature kernel for sm_100a:
//   int16 PCM -> /32768 -> pad/truncate 24000 -> reflect-centred STFT (n_fft 400, hop 160,
//   periodic Hann) -> |X|^2 -> slaney mel (sparse rows) -> power_to_db(ref=max, top_db 80)
//   -> min-max normalise -> zero-crossing rate -> RGB uint8 image, rows flipped.
// Replaces OverlapFeaturesGenerator.generate_mels / generate_zcr / normalize_matrix /
// generate_zcr_image (OverlapDetection/scripts/overlap_features_generator.py:65-151) and the
// PNG round trip plt.imsave(origin='lower') -> tf.image.decode_png(.,3)
// (OverlapDetection/scripts/record_on_pc.py:139,156-158).
//
// One persistent CTA per SM handles one clip at a time, entirely in shared memory:
//   PCM (TMA bulk copy, 48 KB)  ->  32-frame tiles:  windowed frames folded into even/odd
//   halves  ->  400-point real DFT as two [32x201]x[201x201] fp32 register-tiled contractions
//   against cos/sin tables streamed L2->smem with cp.async  ->  |X|^2 tile  ->  sparse mel
//   accumulation into the clip's [128x151] mel tile (kept in smem for the two-pass
//   max / min normalisation)  ->  dB, clip, normalise, image quantisation, coalesced stores.
// n_fft = 400 = 2^4*5^2 is not a power of two; the folded dense DFT costs 24 MFLOP/clip, ~1 %
// of the overlap classifier's 1.84 GFLOP/clip, so a mixed-radix FFT is not worth its
// complexity here.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kClip = 24000;          // hop * 150
constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kFrames = 151;
constexpr int kBins = 201;
constexpr int kBinsPad = 208;         // 52 threads x 4 bins
constexpr int kFold = 201;            // folded sample index n = 0..200
constexpr int kTileF = 32;            // frames per tile
constexpr int kMaxMels = 128;
constexpr int kMStride = 152;
constexpr int kChunk = 8;             // table rows per cp.async stage
constexpr int kMaxMelNnz = 1024;

struct OverlapTables {
    float cosT[kFold][kBinsPad];      // cos(2 pi k n / 400), zero padded columns
    float sinT[kFold][kBinsPad];      // sin(2 pi k n / 400)
    float window[kNfft];              // periodic Hann
    int mel_start[kMaxMels];
    int mel_len[kMaxMels];
    int mel_off[kMaxMels];
    float mel_w[kMaxMelNnz];
};

struct Smem {
    alignas(16) int16_t pcm[kClip + 16];
    alignas(16) float Ae[kFold][kTileF];          // even fold  (overlaid by the |X|^2 tile)
    alignas(16) float Ao[kFold][kTileF];          // odd fold
    alignas(16) float tabC[2][kChunk][kBinsPad];
    alignas(16) float tabS[2][kChunk][kBinsPad];
    float M[kMaxMels][kMStride];
    float zcr[kFrames + 1];
    uint32_t zq[kFrames + 1];                      // image channel 0 per frame: (uint8) trunc(float64(zcr) * 255), computed once
    float red[kThreads / 32];
    float bcast[4];
    int mel_start[kMaxMels], mel_len[kMaxMels], mel_off[kMaxMels];
    float mel_w[kMaxMelNnz];
    float window[kNfft];
    alignas(8) uint64_t bar;
};

struct Params {
    const int16_t* pcm;
    const long long* clip_off;
    const int* clip_len_arr;
    const OverlapTables* tab;
    const unsigned char* dft_b;      // tensor-core path: [13 K steps][cos hi | cos lo | sin hi | sin lo][6656 B], UMMA K-major
    long long n_clips, clip_stride;
    int clip_len, n_mels;
    float* s_db;
    float* s_db_norm;
    float* zcr;
    uint8_t* image;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NT, class SmemT>
__device__ __forceinline__ float block_reduce(SmemT& s, float v, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, t) : fminf(v, t);
    }
    __syncthreads();
    if (lane == 0) s.red[warp] = v;
    __syncthreads();
    float r = s.red[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) r = is_max ? fmaxf(r, s.red[w]) : fminf(r, s.red[w]);
    return r;
}

// power_to_db(ref=max, amin=1e-10, top_db=80) + normalize_matrix + the three outputs, from the clip's mel tile s.M
// (overlap_features_generator.py:82-83,103-117,142-151); every thread of the CTA takes part.
template <int NT, class SmemT>
__device__ __forceinline__ void finish_clip(SmemT& s, const Params& p, long long clip, int n_mels, int tid, unsigned char* scratch,
                                            int scratch_dirty_bytes) {
    const int total = n_mels * kFrames;
    float vmax = 0.f;
    for (int e = tid; e < total; e += NT) vmax = fmaxf(vmax, s.M[e / kFrames][e % kFrames]);
    vmax = block_reduce<NT>(s, vmax, true);
    // numpy-1.21 semantics: the reference term is evaluated in float64, then applied in float32
    const float ref_db = static_cast<float>(10.0 * log10(fmax(1e-10, static_cast<double>(vmax))));
    const float max_db = 10.0f * log10f(fmaxf(1e-10f, vmax)) - ref_db;
    const float floor_db = max_db - 80.0f;
    float vmin = max_db;
    for (int e = tid; e < total; e += NT) {
        float* q = &s.M[e / kFrames][e % kFrames];
        float db = 10.0f * log10f(fmaxf(1e-10f, *q)) - ref_db;
        db = fmaxf(db, floor_db);
        *q = db;
        vmin = fminf(vmin, db);
    }
    vmin = block_reduce<NT>(s, vmin, false);
    const float diff = max_db - vmin;
    __syncthreads();
    if (p.s_db) {
        float* dst = p.s_db + clip * static_cast<long long>(total);
        for (int e = tid; e < total; e += NT) dst[e] = s.M[e / kFrames][e % kFrames];
    }
    if (p.s_db_norm) {
        float* dst = p.s_db_norm + clip * static_cast<long long>(total);
        for (int e = tid; e < total; e += NT) dst[e] = (s.M[e / kFrames][e % kFrames] - vmin) / diff;
    }
    if (p.image) {
        // uint8 [n_mels][151][3], row r = mel (n_mels-1-r); value trunc(float64(v) * 255).  Channels 1 and 2 carry the same
        // value, channel 0 the frame's ZCR byte: one quantisation per PIXEL into a shared-memory copy of the image, then
        // 16-byte stores (packing per 32-bit word of the row-major byte stream cost ~200 instructions per word: two
        // divisions and a quantisation per byte — 45 % of the kernel's stall samples, profiles/r02/overlap_features_tc_v2_*).
        auto quant = [&](int r, int t) -> uint32_t {
            // trunc(float64(v) * 255) for the float32 v = 1 - nrm in [0, 1] in integer arithmetic: the 24-bit significand
            // times 255 fits 32 bits, so the result is exact (FP64 multiplies are slow on this part).  NaN (constant
            // clip), negatives and values below 2^-8 map to 0 exactly as the float64 path did.
            const float nrm = (s.M[n_mels - 1 - r][t] - vmin) / diff;
            const uint32_t bits = __float_as_uint(1.0f - nrm);
            const int ex = static_cast<int>((bits >> 23) & 0xFFu);
            uint32_t q = 0u;
            if (!(bits >> 31) && ex != 0xFF && ex >= 119) q = ex >= 127 ? 255u : ((((bits & 0x7FFFFFu) | 0x800000u) * 255u) >> (150 - ex));
            return q & 0xFFu;
        };
        uint8_t* gimg = p.image + clip * static_cast<long long>(total) * 3;
        const int bytes = total * 3;
        if ((bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(gimg) & 15) == 0) {
            for (int pix = tid; pix < total; pix += NT) {
                const int r = pix / kFrames, t = pix - r * kFrames;
                const uint32_t g = quant(r, t);
                scratch[3 * pix] = static_cast<uint8_t>(s.zq[t]);
                scratch[3 * pix + 1] = static_cast<uint8_t>(g);
                scratch[3 * pix + 2] = static_cast<uint8_t>(g);
            }
            __syncthreads();
            const uint4* src = reinterpret_cast<const uint4*>(scratch);
            uint4* dst = reinterpret_cast<uint4*>(gimg);
            for (int i = tid; i < bytes / 16; i += NT) dst[i] = src[i];
            __syncthreads();
            // the part of the scratch that doubles as a tensor-core operand ring must hold finite values again
            for (int i = tid; i < scratch_dirty_bytes / 16; i += NT) reinterpret_cast<uint4*>(scratch)[i] = make_uint4(0u, 0u, 0u, 0u);
        } else {
            uint32_t* dst = reinterpret_cast<uint32_t*>(gimg);   // odd sizes: per 32-bit word of the byte stream
            const int words = bytes / 4;
            for (int wd = tid; wd < words; wd += NT) {
                uint32_t packed = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int byte = 4 * wd + b;
                    const int pix = byte / 3, c = byte - 3 * pix;
                    const int r = pix / kFrames, t = pix - r * kFrames;
                    packed |= (c == 0 ? (s.zq[t] & 0xFFu) : quant(r, t)) << (8 * b);
                }
                dst[wd] = packed;
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1) overlap_features_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x;
    const OverlapTables& T = *p.tab;

    for (int i = tid; i < kMaxMels; i += kThreads) {
        s.mel_start[i] = T.mel_start[i];
        s.mel_len[i] = T.mel_len[i];
        s.mel_off[i] = T.mel_off[i];
    }
    for (int i = tid; i < kMaxMelNnz; i += kThreads) s.mel_w[i] = T.mel_w[i];
    for (int i = tid; i < kNfft; i += kThreads) s.window[i] = T.window[i];
    if (tid == 0) {
        mbar_init(&s.bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    const int n_mels = p.n_mels;

    for (long long clip = blockIdx.x; clip < p.n_clips; clip += gridDim.x) {
        const long long off = p.clip_off ? p.clip_off[clip] : clip * p.clip_stride;
        const int len = p.clip_len_arr ? p.clip_len_arr[clip] : p.clip_len;
        const int nvalid = min(len, kClip);
        const long long a0 = off & ~7LL;
        const int shift = static_cast<int>(off - a0);
        // ---- PCM -> smem (TMA bulk copy) ----------------------------------------------------------
        if (tid == 0) {
            uint32_t bytes = static_cast<uint32_t>(((off + nvalid + 7) & ~7LL) - a0) * 2u;
            if (bytes == 0) bytes = 16;
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(&s.bar, bytes);
            tma_bulk_g2s(&s.pcm[0], p.pcm + a0, bytes, &s.bar);
        }
        mbar_wait(&s.bar, phase);
        phase ^= 1u;
        const unsigned short* px = reinterpret_cast<const unsigned short*>(&s.pcm[0]) + shift;
        auto sample = [&](int i) -> float {          // y[i], i in [0, 24000): librosa.load scale
            return i < nvalid ? s16_bits_to_float(px[i]) * (1.0f / 32768.0f) : 0.0f;
        };

        // ---- zero-crossing rate (edge-padded frames of 400, hop 160) -----------------------------
        if (tid < kFrames) {
            const int base = tid * kHop - kNfft / 2;
            int prev_neg = sample(min(max(base, 0), kClip - 1)) < 0.0f;
            int cnt = 0;
            for (int j = 1; j < kNfft; ++j) {
                const int neg = sample(min(max(base + j, 0), kClip - 1)) < 0.0f;
                cnt += neg != prev_neg;
                prev_neg = neg;
            }
            const double z = static_cast<double>(cnt) / 400.0;      // np.mean over 400 booleans
            s.zcr[tid] = static_cast<float>(z);
            {
                const double v255 = z * 255.0;
                s.zq[tid] = (v255 >= 0.0) ? static_cast<uint32_t>(static_cast<int>(v255)) & 0xFFu : 0u;
            }
            if (p.zcr) p.zcr[clip * kFrames + tid] = static_cast<float>(z);
        }

        // ---- STFT power + mel, 32 frames at a time -----------------------------------------------
        for (int t0 = 0; t0 < kFrames; t0 += kTileF) {
            __syncthreads();                                         // previous tile's S reads done
            // folded, windowed frames: Ae[n][f] = w[n]y[n] + w[400-n]y[400-n], Ao = difference
            for (int e = tid; e < kFold * kTileF; e += kThreads) {
                const int n = e / kTileF, f = e % kTileF;
                const int t = t0 + f;
                float ve = 0.f, vo = 0.f;
                if (t < kFrames) {
                    const int base = t * kHop - kNfft / 2;           // reflect-padded origin
                    auto refl = [&](int i) { return i < 0 ? -i : (i >= kClip ? 2 * (kClip - 1) - i : i); };
                    const float a = s.window[n] * sample(refl(base + n));
                    if (n == 0 || n == kNfft / 2) {
                        ve = a;
                    } else {
                        const float b = s.window[kNfft - n] * sample(refl(base + kNfft - n));
                        ve = a + b;
                        vo = a - b;
                    }
                }
                s.Ae[n][f] = ve;
                s.Ao[n][f] = vo;
            }
            // contraction: thread = (kc: 4 bins, fg: 8 frames); 208 of 256 threads active
            const int kc = tid % 52, fg = tid / 52;
            const bool active = fg < 4;
            float re[8][4], im[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) re[i][j] = im[i][j] = 0.f;
            auto stage = [&](int chunk, int buf) {
                // kChunk rows x 208 floats of each table = 2 x 416 float4
                const int row0 = chunk * kChunk;
                for (int e = tid; e < 2 * kChunk * (kBinsPad / 4); e += kThreads) {
                    const int which = e / (kChunk * (kBinsPad / 4));
                    const int r = (e / (kBinsPad / 4)) % kChunk;
                    const int c4 = e % (kBinsPad / 4);
                    const int row = min(row0 + r, kFold - 1);
                    const float* src = (which ? &T.sinT[row][0] : &T.cosT[row][0]) + 4 * c4;
                    float* dst = (which ? &s.tabS[buf][r][0] : &s.tabC[buf][r][0]) + 4 * c4;
                    cp_async16(dst, src);
                }
                cp_async_commit();
            };
            constexpr int kNChunks = (kFold + kChunk - 1) / kChunk;  // 26
            stage(0, 0);
            for (int ch = 0; ch < kNChunks; ++ch) {
                const int buf = ch & 1;
                if (ch + 1 < kNChunks) {
                    stage(ch + 1, buf ^ 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();                                     // chunk landed; Ae/Ao ready
                if (active) {
                    const int rows = min(kChunk, kFold - ch * kChunk);
                    for (int r = 0; r < rows; ++r) {
                        const int n = ch * kChunk + r;
                        const float4 c = *reinterpret_cast<const float4*>(&s.tabC[buf][r][4 * kc]);
                        const float4 sn = *reinterpret_cast<const float4*>(&s.tabS[buf][r][4 * kc]);
                        const float4 e0 = *reinterpret_cast<const float4*>(&s.Ae[n][8 * fg]);
                        const float4 e1 = *reinterpret_cast<const float4*>(&s.Ae[n][8 * fg + 4]);
                        const float4 o0 = *reinterpret_cast<const float4*>(&s.Ao[n][8 * fg]);
                        const float4 o1 = *reinterpret_cast<const float4*>(&s.Ao[n][8 * fg + 4]);
                        const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                        const float ov[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
                        const float cv[4] = {c.x, c.y, c.z, c.w};
                        const float sv[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                re[i][j] = fmaf(ev[i], cv[j], re[i][j]);
                                im[i][j] = fmaf(ov[i], sv[j], im[i][j]);
                            }
                    }
                }
                __syncthreads();                                     // buffer may be restaged
            }
            // |X|^2 tile S[f][k], overlaid on Ae (all reads of Ae/Ao are behind the barrier above)
            float* S = &s.Ae[0][0];                                  // [kTileF][kBinsPad]: spans Ae and the head of Ao (both dead)
            if (active) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        S[(8 * fg + i) * kBinsPad + 4 * kc + j] = re[i][j] * re[i][j] + im[i][j] * im[i][j];
            }
            __syncthreads();
            // sparse mel rows: M[m][t] = sum_k w[m][k] S[t][k]
            for (int e = tid; e < n_mels * kTileF; e += kThreads) {
                const int m = e / kTileF, f = e % kTileF;
                const int t = t0 + f;
                if (t < kFrames) {
                    const float* w = &s.mel_w[s.mel_off[m]];
                    const float* x = &S[f * kBinsPad + s.mel_start[m]];
                    float acc = 0.f;
                    for (int i = 0; i < s.mel_len[m]; ++i) acc = fmaf(w[i], x[i], acc);
                    s.M[m][t] = acc;
                }
            }
        }
        __syncthreads();

        // scratch: Ae | Ao | tabC | tabS are contiguous (78 KB >= the 58 KB image) and rewritten by the next tile before use
        finish_clip<kThreads>(s, p, clip, n_mels, tid, reinterpret_cast<unsigned char*>(&s.Ae[0][0]), 0);
    }
}

// =================================================================================================
// Tensor-core variant: the folded 400-point DFT as a [frames x 208] x [208 x 208] product pair on tcgen05
// =================================================================================================
// X[k] = sum_{n<=200} e[n] cos(2 pi k n / 400) - i sum_{n<200} o[n] sin(2 pi k n / 400), e / o the even / odd folds of the
// Hann-windowed, reflect-padded frame.  Rows of the A operands are FRAMES (M = 128: frames 0..127 of the clip, then
// 128..150), K is the folded sample index in 13 steps of 16, N = 208 bins (201 used).  fp32-grade accuracy comes from the
// same fp16 hi + lo split as csrc/mfcc_tc.cu: D += A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32 accumulation in TMEM (D_cos in
// columns [0, 208), D_sin in [208, 416)).  Samples enter at int16 scale x 0.5, so e <= 32768 and the lo halves stay in the
// fp16 normal range; the power is scaled back by 2^-28.
//   warps 0-3   epilogue: zero-crossing counts while the first product runs, then tcgen05.ld of their 32 frames' spectrum
//               (thread = frame), |X|^2 into a per-thread scratch (local memory: lane-interleaved, so a warp's access is
//               one line), sparse slaney mel rows -> the clip's mel tile in shared memory
//   warps 4-11  A producers: per K step 128 frames x 16 folded samples; a lane owns one pair of samples of one frame
//               (eight lanes per frame read 32 contiguous bytes), window, fold, hi / lo split, 4-byte stores into the
//               UMMA K-major no-swizzle layout (K-group stride 2112 B: conflict-free)
//   warp 12     TMA: the step's four B tiles (cos hi | cos lo | sin hi | sin lo, host-arranged, 26 KB) by one bulk copy
//   warp 13     MMA issue: six N = 208 products per step (104 cycles each by the N/2 law: the tensor pipe is busy)
// after which all 448 threads run the clip-level dB / min-max / image pass (finish_clip).
constexpr int kTcThreads = 448;
constexpr int kKSteps = 13;                            // 208 folded samples / 16
constexpr int kALbo = 2112;                            // K-group stride of an A tile (2048 + 64: shifts banks by 16)
constexpr int kATile = 2 * kALbo;                      // one operand (e|o, hi|lo) of one K step: 128 rows x 16 halves
constexpr int kBLbo = 26 * 128;                        // K-group stride of a B tile: 26 row groups of 8 bins
constexpr int kBTile = 2 * kBLbo;                      // 6656 B: 208 bins x 16 halves
constexpr int kStageA = 4 * kATile;                    // e_hi, e_lo, o_hi, o_lo
constexpr int kStageB = 4 * kBTile;                    // cos_hi, cos_lo, sin_hi, sin_lo
constexpr int kStages = 2;
constexpr float kAScale = 0.5f;                        // int16 scale x 0.5 = librosa's float x 2^14
constexpr float kPowScale = 1.0f / (16384.0f * 16384.0f);

struct SmemTc {
    alignas(128) unsigned char a[kStages][kStageA];
    alignas(128) unsigned char b[kStages][kStageB];
    alignas(16) int16_t pcm[kClip + 16];
    float M[kMaxMels][kMStride];
    float zcr[kFrames + 1];
    uint32_t zq[kFrames + 1];                      // image channel 0 per frame: (uint8) trunc(float64(zcr) * 255), computed once
    float red[kTcThreads / 32];
    int mel_start[kMaxMels], mel_len[kMaxMels], mel_off[kMaxMels];
    float mel_w[kMaxMelNnz];
    float window[kNfft];
    alignas(8) uint64_t bar_pcm, full[kStages], empty[kStages], d_full;
    uint32_t tmem_base;
};
static_assert(sizeof(SmemTc) <= 227 * 1024, "SmemTc exceeds the 227 KB a CTA can own");
static_assert(offsetof(SmemTc, b) == kStages * kStageA && kStages * (kStageA + kStageB) >= kMaxMels * kFrames * 3,
              "the image scratch needs the A and B rings back to back");

__device__ __forceinline__ bool elect_one_tc() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t desc_k_major(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16_tc(uint32_t d_tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
                 "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_commit_tc(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tc(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: traps instead of hanging the GPU on a protocol error
// (try_wait with a suspend-time hint parks the warp in hardware: the plain polling loop was 18 % of the kernel's executed
//  instructions, taken from the issue slots of the producer and epilogue warps — profiles/r02/overlap_features_tc_v3_*)
__device__ __noinline__ void wait_tc(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
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
__device__ __forceinline__ void tmem_ld16_tc(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_tc() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void split2_tc(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(kTcThreads, 1) overlap_features_tc_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SmemTc& s = *reinterpret_cast<SmemTc*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const OverlapTables& T = *p.tab;
    for (int i = tid; i < kMaxMels; i += kTcThreads) {
        s.mel_start[i] = T.mel_start[i];
        s.mel_len[i] = T.mel_len[i];
        s.mel_off[i] = T.mel_off[i];
    }
    for (int i = tid; i < kMaxMelNnz; i += kTcThreads) s.mel_w[i] = T.mel_w[i];
    for (int i = tid; i < kNfft; i += kTcThreads) s.window[i] = T.window[i];
    {   // A tiles: finite everywhere
        uint4* z = reinterpret_cast<uint4*>(&s.a[0][0]);
        for (int i = tid; i < kStages * kStageA / 16; i += kTcThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        mbar_init(&s.bar_pcm, 1);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&s.full[i], 9);                   // 8 producer warps + the TMA warp's expect_tx arrive
            mbar_init(&s.empty[i], 1);
        }
        mbar_init(&s.d_full, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;
    const int n_mels = p.n_mels;
    uint32_t pcm_phase = 0;
    uint32_t step = 0;                                  // K steps so far (ring position); the same count in every role
    uint32_t tiles_done = 0;                            // M tiles so far (d_full phase)
    float* S = reinterpret_cast<float*>(&s.b[0][0]);    // [208 bins][64 frames] power tile, overlaid on the idle B ring
    static_assert(kStages * kStageB >= kBinsPad * 64 * 4, "power tile does not fit the B ring");

    for (long long clip = blockIdx.x; clip < p.n_clips; clip += gridDim.x) {
        const long long off = p.clip_off ? p.clip_off[clip] : clip * p.clip_stride;
        const int len = p.clip_len_arr ? p.clip_len_arr[clip] : p.clip_len;
        const int nvalid = min(len, kClip);
        const long long a0 = off & ~7LL;
        const int shift = static_cast<int>(off - a0);
        if (tid == 0) {
            uint32_t bytes = static_cast<uint32_t>(((off + nvalid + 7) & ~7LL) - a0) * 2u;
            if (bytes == 0) bytes = 16;
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(&s.bar_pcm, bytes);
            tma_bulk_g2s(&s.pcm[0], p.pcm + a0, bytes, &s.bar_pcm);
        }
        mbar_wait(&s.bar_pcm, pcm_phase);
        pcm_phase ^= 1u;
        unsigned short* px = reinterpret_cast<unsigned short*>(&s.pcm[0]) + shift;
        for (int i = nvalid + tid; i < kClip; i += kTcThreads) px[i] = 0;       // zero padding to 24000: no bounds tests later
        __syncthreads();

        // ---- zero-crossing rate: edge padding adds no crossings, so a frame's count is a difference of the running
        //      count C(i) = #{1 <= j <= i : sign(y[j]) != sign(y[j-1])} at its clamped ends (block scan over 54-sample runs)
        {
            constexpr int kSeg = 54;                                            // 448 x 54 >= 24000
            int* zc = reinterpret_cast<int*>(&s.a[0][0]);                       // scratch: [448] run totals, [302] end counts
            const int i0 = tid * kSeg, i1 = min(i0 + kSeg, kClip);
            int cnt = 0;
            if (i0 < kClip) {
                int prev = i0 > 0 ? px[i0 - 1] >> 15 : px[0] >> 15;
                for (int i = i0; i < i1; ++i) {
                    const int cur = px[i] >> 15;
                    cnt += cur != prev;
                    prev = cur;
                }
            }
            int incl = cnt;                                                     // inclusive scan: warp, then across warps
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int* wsum = zc + 800;
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            int base_w = 0;
            for (int w = 0; w < warp; ++w) base_w += wsum[w];
            zc[tid] = base_w + incl - cnt;                                      // exclusive prefix: crossings before run `tid`
            __syncthreads();
            if (tid < 2 * kFrames) {
                const int t = tid >> 1;
                const int b = t * kHop - kNfft / 2;
                const int i = (tid & 1) ? min(b + kNfft - 1, kClip - 1) : max(b, 0);
                const int r = i / kSeg;
                int c = zc[r];
                int prev = r * kSeg > 0 ? px[r * kSeg - 1] >> 15 : px[0] >> 15;
                for (int j = r * kSeg; j <= i; ++j) {
                    const int cur = px[j] >> 15;
                    c += cur != prev;
                    prev = cur;
                }
                zc[448 + tid] = c;                                              // C(i)
            }
            __syncthreads();
            if (tid < kFrames) {
                const double z = static_cast<double>(zc[448 + 2 * tid + 1] - zc[448 + 2 * tid]) / 400.0;
                s.zcr[tid] = static_cast<float>(z);
                {
                    const double v255 = z * 255.0;
                    s.zq[tid] = (v255 >= 0.0) ? static_cast<uint32_t>(static_cast<int>(v255)) & 0xFFu : 0u;
                }
                if (p.zcr) p.zcr[clip * kFrames + tid] = static_cast<float>(z);
            }
            __syncthreads();
            for (int i = tid; i < 1024; i += kTcThreads) zc[i] = 0;             // the A tile's rows must stay finite
            __syncthreads();
        }

        for (int mt = 0; mt < 2; ++mt) {
            if (warp == 12) {
                // ================= TMA: the four B tiles of every K step =================
                if (lane == 0) {
                    fence_proxy_async_smem();                                   // the ring held the power tile
                    for (int ks = 0; ks < kKSteps; ++ks) {
                        const uint32_t st = (step + ks) % kStages, use = (step + ks) / kStages;
                        if (use >= 1) wait_tc(&s.empty[st], (use - 1) & 1);
                        mbar_arrive_expect_tx(&s.full[st], kStageB);
                        tma_bulk_g2s(&s.b[st][0], p.dft_b + static_cast<size_t>(ks) * kStageB, kStageB, &s.full[st]);
                    }
                }
                __syncwarp();
            } else if (warp == 13) {
                // ================= MMA issue =================
                constexpr uint32_t kIdesc = (1u << 4) | (static_cast<uint32_t>(kBinsPad >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
                const uint64_t dA = desc_k_major(smem_u32(&s.a[0][0]), kALbo, 128);
                const uint64_t dB = desc_k_major(smem_u32(&s.b[0][0]), kBLbo, 128);
                for (int ks = 0; ks < kKSteps; ++ks) {
                    const uint32_t st = (step + ks) % kStages, use = (step + ks) / kStages;
                    wait_tc(&s.full[st], use & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one_tc()) {
                        const uint64_t a = dA + static_cast<uint64_t>((st * kStageA) >> 4);
                        const uint64_t b = dB + static_cast<uint64_t>((st * kStageB) >> 4);
                        const uint32_t acc = ks != 0 ? 1u : 0u;
                        // cos part: e_hi Bc_hi + e_lo Bc_hi + e_hi Bc_lo ; sin part: o_hi Bs_hi + o_lo Bs_hi + o_hi Bs_lo
                        umma_f16_tc(tmem, a, b, kIdesc, acc);
                        umma_f16_tc(tmem, a + (kATile >> 4), b, kIdesc, 1u);
                        umma_f16_tc(tmem, a, b + (kBTile >> 4), kIdesc, 1u);
                        umma_f16_tc(tmem + kBinsPad, a + ((2 * kATile) >> 4), b + ((2 * kBTile) >> 4), kIdesc, acc);
                        umma_f16_tc(tmem + kBinsPad, a + ((3 * kATile) >> 4), b + ((2 * kBTile) >> 4), kIdesc, 1u);
                        umma_f16_tc(tmem + kBinsPad, a + ((2 * kATile) >> 4), b + ((3 * kBTile) >> 4), kIdesc, 1u);
                        umma_commit_tc(&s.empty[st]);
                        if (ks == kKSteps - 1) umma_commit_tc(&s.d_full);
                    }
                    __syncwarp();
                }
            } else if (warp >= 4) {
                // ================= A producers =================
                const int pt = tid - 128;                                    // 0..255
                const int fs = pt >> 3, ps = pt & 7;                         // frame slot 0..31, sample pair 0..7
                const int rows = mt == 0 ? 128 : kFrames - 128;              // frames of this tile
                for (int ks = 0; ks < kKSteps; ++ks) {
                    const uint32_t st = (step + ks) % kStages, use = (step + ks) / kStages;
                    if (use >= 1) wait_tc(&s.empty[st], (use - 1) & 1);
                    unsigned char* at = &s.a[st][0];
                    const int n = 16 * ks + 2 * ps;                      // folded sample indices n, n + 1
                    const float w0 = n <= 200 ? s.window[n] * kAScale : 0.f;
                    const float w1 = n + 1 <= 200 ? s.window[n + 1] * kAScale : 0.f;
                    const bool m0 = n >= 1 && n <= 199, m1 = n + 1 <= 199;   // samples with a mirror partner
                    const uint32_t koff = (ps >> 2) * kALbo + (ps & 3) * 4;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int f = fs + 32 * q;                       // row of the tile
                        if (mt == 1 && f >= 32 && f >= rows) break;      // tile 1 has 23 frames: rows 32.. keep tile 0's (finite,
                                                                         // unused) values, rows are independent in the product
                        const int t = 128 * mt + f;                      // frame of the clip
                        float e0 = 0.f, e1 = 0.f, o0 = 0.f, o1 = 0.f;
                        if (t < kFrames) {
                            const int base = t * kHop - kNfft / 2;
                            float a_0, a_1, b_0, b_1;
                            if (t >= 2 && t <= 148) {                    // interior frame: no reflection
                                a_0 = s16_bits_to_float(px[base + n]);
                                a_1 = s16_bits_to_float(px[base + n + 1]);
                                b_0 = s16_bits_to_float(px[base + kNfft - n]);
                                b_1 = s16_bits_to_float(px[base + kNfft - n - 1]);
                            } else {
                                auto smp = [&](int i) -> float {         // librosa centre padding: reflect
                                    i = i < 0 ? -i : (i >= kClip ? 2 * (kClip - 1) - i : i);
                                    return s16_bits_to_float(px[i]);
                                };
                                a_0 = smp(base + n);
                                a_1 = smp(base + n + 1);
                                b_0 = smp(base + kNfft - n);
                                b_1 = smp(base + kNfft - n - 1);
                            }
                            if (!m0) b_0 = 0.f;
                            if (!m1) b_1 = 0.f;
                            e0 = w0 * (a_0 + b_0);
                            e1 = w1 * (a_1 + b_1);
                            o0 = m0 ? w0 * (a_0 - b_0) : 0.f;
                            o1 = m1 ? w1 * (a_1 - b_1) : 0.f;
                        }
                        uint32_t ehi, elo, ohi, olo;
                        split2_tc(e0, e1, ehi, elo);
                        split2_tc(o0, o1, ohi, olo);
                        const uint32_t roff = koff + (f >> 3) * 128 + (f & 7) * 16;
                        *reinterpret_cast<uint32_t*>(at + roff) = ehi;
                        *reinterpret_cast<uint32_t*>(at + kATile + roff) = elo;
                        *reinterpret_cast<uint32_t*>(at + 2 * kATile + roff) = ohi;
                        *reinterpret_cast<uint32_t*>(at + 3 * kATile + roff) = olo;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_tc(&s.full[st]);
                }
            }
            step += kKSteps;
            // ---- spectrum -> power -> mel, 64 frames at a time; the power tile S[bin][frame] overlays the B ring, which
            //      is idle once the tile's last product has completed (d_full) ----
            const int n_halves = mt == 0 ? 2 : 1;
            for (int h = 0; h < n_halves; ++h) {
                if (h == 0) {                                                   // every product of the tile has completed:
                    wait_tc(&s.d_full, tiles_done & 1);                         // accumulators readable, B ring idle
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                if (warp < 4 && (warp >> 1) == h) {
                    const uint32_t trow = tmem + (static_cast<uint32_t>(32 * warp) << 16);
                    const int fh = 32 * (warp & 1) + lane;
#pragma unroll 1
                    for (int c = 0; c < kBinsPad / 16; ++c) {
                        uint32_t re[16], im[16];
                        tmem_ld16_tc(trow + 16 * c, re);
                        tmem_ld16_tc(trow + kBinsPad + 16 * c, im);
                        tmem_wait_tc();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float a = __uint_as_float(re[j]), b = __uint_as_float(im[j]);
                            S[(16 * c + j) * 64 + fh] = fmaf(a, a, b * b) * kPowScale;
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                }
                __syncthreads();
                const int t0 = 128 * mt + 64 * h;
                for (int e = tid; e < n_mels * 64; e += kTcThreads) {
                    const int m = e >> 6, fh = e & 63;
                    const int t = t0 + fh;
                    if (t < kFrames) {
                        const float* w = &s.mel_w[s.mel_off[m]];
                        const float* x = &S[s.mel_start[m] * 64 + fh];
                        const int ln = s.mel_len[m];
                        float acc = 0.f;
                        // (1..10 taps: the compiler's unroll-by-8 with remainder chains cost ~60 instructions of control per
                        //  item; a plain loop is shorter)
#pragma unroll 1
                        for (int i = 0; i < ln; ++i) acc = fmaf(w[i], x[i * 64], acc);
                        s.M[m][t] = acc;
                    }
                }
                __syncthreads();
            }
            ++tiles_done;
        }
        // scratch: the A and B operand rings are contiguous (87 KB >= the 58 KB image) and idle here
        finish_clip<kTcThreads>(s, p, clip, n_mels, tid, &s.a[0][0], kStages * kStageA);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// -------------------------------------------------------------------------------------------------
// host: tables (librosa.filters.mel slaney, float32 storage semantics) and launch
// -------------------------------------------------------------------------------------------------
double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

int build_tables(int n_mels, OverlapTables& t) {
    memset(&t, 0, sizeof(t));
    const double PI = 3.14159265358979323846;
    for (int n = 0; n < kFold; ++n)
        for (int k = 0; k < kBins; ++k) {
            const int r = (n * k) % kNfft;                           // exact argument reduction
            t.cosT[n][k] = static_cast<float>(cos(2.0 * PI * r / kNfft));
            t.sinT[n][k] = static_cast<float>(sin(2.0 * PI * r / kNfft));
        }
    for (int n = 0; n < kNfft; ++n) t.window[n] = static_cast<float>(0.5 - 0.5 * cos(2.0 * PI * n / kNfft));
    // librosa.filters.mel(sr=16000, n_fft=400, n_mels, fmin=0, fmax=8000, htk=False, norm='slaney')
    const double sr = 16000.0, f_hi = sr / 2;
    std::vector<double> mel_f(n_mels + 2);
    const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(f_hi);
    for (int i = 0; i < n_mels + 2; ++i) {
        const double m = (i == n_mels + 1) ? m_hi : m_lo + i * ((m_hi - m_lo) / (n_mels + 1));
        mel_f[i] = mel_to_hz(m);
    }
    int nnz = 0;
    for (int i = 0; i < n_mels; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        int first = -1, last = -1;
        std::vector<float> row(kBins, 0.f);
        for (int k = 0; k < kBins; ++k) {
            const double fk = (sr / 2) * k / (kBins - 1);            // np.linspace(0, sr/2, 201)
            const double lower = -(mel_f[i] - fk) / fd0;
            const double upper = (mel_f[i + 2] - fk) / fd1;
            const double w = fmax(0.0, fmin(lower, upper));
            const float w32 = static_cast<float>(w);                 // stored into the float32 array
            const float wn = static_cast<float>(static_cast<double>(w32) * enorm);   // weights *= enorm
            row[k] = wn;
            if (wn != 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        }
        const int len = first < 0 ? 0 : last - first + 1;
        if (nnz + len > kMaxMelNnz) {
            mmla_set_error("overlap: mel basis has more than %d non-zeros", kMaxMelNnz);
            return MMLA_EUNSUP;
        }
        t.mel_start[i] = first < 0 ? 0 : first;
        t.mel_len[i] = len;
        t.mel_off[i] = nnz;
        for (int k = 0; k < len; ++k) t.mel_w[nnz + k] = row[first + k];
        nnz += len;
    }
    return MMLA_OK;
}

std::mutex g_mu;
std::map<std::pair<int, int>, OverlapTables*> g_cache;   // (device, n_mels)
std::map<int, unsigned char*> g_dft_b;                   // device -> tensor-core DFT operand table

// B operands of the tensor-core DFT: per K step the tiles cos hi | cos lo | sin hi | sin lo, each [208 bins x 16 folded
// samples] fp16 in the UMMA K-major no-swizzle layout (core matrix = 8 bins x 16 B; K-group stride 26 x 128 B).
int get_dft_b(const unsigned char** out) {
    int dev = 0;
    MMLA_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_dft_b.find(dev);
    if (it != g_dft_b.end()) {
        *out = it->second;
        return MMLA_OK;
    }
    std::vector<unsigned char> host(static_cast<size_t>(kKSteps) * kStageB, 0);
    const double PI = 3.14159265358979323846;
    auto put = [&](unsigned char* hi, unsigned char* lo, int col, int kk, double v) {
        const size_t off = static_cast<size_t>(kk / 8) * kBLbo + static_cast<size_t>(col / 8) * 128 + (col % 8) * 16 + (kk % 8) * 2;
        const __half h = __float2half_rn(static_cast<float>(v));
        const __half l = __float2half_rn(static_cast<float>(v - static_cast<double>(__half2float(h))));
        memcpy(hi + off, &h, 2);
        memcpy(lo + off, &l, 2);
    };
    for (int ks = 0; ks < kKSteps; ++ks) {
        unsigned char* base = host.data() + static_cast<size_t>(ks) * kStageB;
        for (int kk = 0; kk < 16; ++kk) {
            const int n = 16 * ks + kk;
            for (int k = 0; k < kBins; ++k) {
                if (n > 200) continue;
                const int r = (n * k) % kNfft;                               // exact argument reduction
                put(base, base + kBTile, k, kk, cos(2.0 * PI * r / kNfft));
                if (n >= 1 && n <= 199) put(base + 2 * kBTile, base + 3 * kBTile, k, kk, sin(2.0 * PI * r / kNfft));
            }
        }
    }
    unsigned char* d = nullptr;
    cudaError_t e = cudaMalloc(&d, host.size());
    if (e == cudaSuccess) e = cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        mmla_set_error("overlap: DFT operand upload failed: %s", cudaGetErrorString(e));
        return MMLA_ECUDA;
    }
    g_dft_b[dev] = d;
    *out = d;
    return MMLA_OK;
}

int get_tables(int n_mels, const OverlapTables** out) {
    int dev = 0;
    MMLA_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> g(g_mu);
    auto key = std::make_pair(dev, n_mels);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) {
        *out = it->second;
        return MMLA_OK;
    }
    OverlapTables* host = new OverlapTables;
    int rc = build_tables(n_mels, *host);
    if (rc != MMLA_OK) {
        delete host;
        return rc;
    }
    OverlapTables* devp = nullptr;
    cudaError_t e = cudaMalloc(&devp, sizeof(OverlapTables));
    if (e == cudaSuccess) e = cudaMemcpy(devp, host, sizeof(OverlapTables), cudaMemcpyHostToDevice);
    delete host;
    if (e != cudaSuccess) {
        mmla_set_error("overlap tables upload failed: %s", cudaGetErrorString(e));
        return MMLA_ECUDA;
    }
    g_cache[key] = devp;
    *out = devp;
    return MMLA_OK;
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int mmla_overlap_features(
    const int16_t* pcm, int64_t pcm_total, const int64_t* clip_off_host, const int32_t* clip_len_host,
    int64_t n_clips, int32_t clip_len, int64_t clip_stride, int32_t n_mels, float* s_db, float* s_db_norm,
    float* zcr, uint8_t* image, void* stream) {
    MMLA_REQUIRE(pcm != nullptr, MMLA_EINVAL, "overlap: null pcm");
    MMLA_REQUIRE(n_clips >= 0, MMLA_EINVAL, "overlap: negative n_clips");
    if (n_clips == 0) return MMLA_OK;
    MMLA_REQUIRE(n_mels >= 1 && n_mels <= kMaxMels, MMLA_EUNSUP, "overlap: n_mels=%d must be in [1,%d]", n_mels, kMaxMels);
    MMLA_REQUIRE((n_mels * kFrames * 3) % 4 == 0 || image == nullptr, MMLA_EUNSUP,
                 "overlap: image output needs n_mels*453 divisible by 4");
    MMLA_REQUIRE((reinterpret_cast<uintptr_t>(pcm) & 15) == 0, MMLA_EINVAL, "overlap: pcm must be 16-byte aligned");
    MMLA_REQUIRE((reinterpret_cast<uintptr_t>(image) & 3) == 0, MMLA_EINVAL, "overlap: image must be 4-byte aligned");
    MMLA_REQUIRE((clip_off_host == nullptr) == (clip_len_host == nullptr), MMLA_EINVAL,
                 "overlap: clip_off_host and clip_len_host must both be given or both be NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const OverlapTables* tab = nullptr;
    int rc = get_tables(n_mels, &tab);
    if (rc != MMLA_OK) return rc;

    Params kp;
    memset(&kp, 0, sizeof(kp));
    kp.pcm = pcm;
    kp.tab = tab;
    kp.n_clips = n_clips;
    kp.clip_stride = clip_stride;
    kp.clip_len = clip_len;
    kp.n_mels = n_mels;
    kp.s_db = s_db; kp.s_db_norm = s_db_norm; kp.zcr = zcr; kp.image = image;
    void* dev_tmp = nullptr;
    if (clip_off_host) {
        for (int64_t c = 0; c < n_clips; ++c)
            MMLA_REQUIRE(clip_len_host[c] >= 0 && clip_off_host[c] >= 0 && clip_off_host[c] + clip_len_host[c] <= pcm_total,
                         MMLA_EINVAL, "overlap: clip %lld exceeds pcm_total_samples", static_cast<long long>(c));
        const size_t b_off = static_cast<size_t>(n_clips) * sizeof(int64_t);
        const size_t b_len = static_cast<size_t>(n_clips) * sizeof(int32_t);
        MMLA_CUDA_CHECK(cudaMallocAsync(&dev_tmp, b_off + b_len, st));
        char* base = static_cast<char*>(dev_tmp);
        MMLA_CUDA_CHECK(cudaMemcpyAsync(base, clip_off_host, b_off, cudaMemcpyHostToDevice, st));
        MMLA_CUDA_CHECK(cudaMemcpyAsync(base + b_off, clip_len_host, b_len, cudaMemcpyHostToDevice, st));
        kp.clip_off = reinterpret_cast<const long long*>(base);
        kp.clip_len_arr = reinterpret_cast<const int*>(base + b_off);
    } else {
        MMLA_REQUIRE(clip_len >= 0 && clip_stride >= 0, MMLA_EINVAL, "overlap: bad uniform clip geometry");   // stride < len = overlapping windows
        MMLA_REQUIRE((n_clips - 1) * clip_stride + clip_len <= pcm_total, MMLA_EINVAL, "overlap: clips exceed pcm_total_samples");
    }
    const int sms = mmla_num_sms();
    MMLA_REQUIRE(sms > 0, MMLA_ECUDA, "overlap: no CUDA device");
    long long grid = sms;
    if (grid > n_clips) grid = n_clips;
    const char* force = getenv("MMLA_OVERLAP_KERNEL");
    if (force && strcmp(force, "fp32") == 0) {
        // the CUDA-core contraction (r01 kernel): kept as the cross-check of the tensor-core path
        static MmlaPerDeviceOnce attr_once;                      // cudaFuncSetAttribute is per device
        if (attr_once.first())
            MMLA_CUDA_CHECK(cudaFuncSetAttribute(overlap_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(sizeof(Smem))));
        overlap_features_kernel<<<static_cast<unsigned>(grid), kThreads, sizeof(Smem), st>>>(kp);
        mmla_count_launch("overlap_features_kernel", st);
    } else {
        rc = get_dft_b(&kp.dft_b);
        if (rc != MMLA_OK) return rc;
        static MmlaPerDeviceOnce attr_once_tc;
        if (attr_once_tc.first())
            MMLA_CUDA_CHECK(cudaFuncSetAttribute(overlap_features_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(sizeof(SmemTc))));
        overlap_features_tc_kernel<<<static_cast<unsigned>(grid), kTcThreads, sizeof(SmemTc), st>>>(kp);
        mmla_count_launch("overlap_features_tc_kernel", st);
    }
    MMLA_CUDA_CHECK(cudaGetLastError());
    if (dev_tmp) MMLA_CUDA_CHECK(cudaFreeAsync(dev_tmp, st));
    return MMLA_OK;
}
