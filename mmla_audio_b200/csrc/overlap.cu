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
#include <math.h>
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
    double zcr255[kFrames + 1];
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

__device__ __forceinline__ float block_reduce(Smem& s, float v, bool is_max) {
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
    for (int w = 1; w < kThreads / 32; ++w) r = is_max ? fmaxf(r, s.red[w]) : fminf(r, s.red[w]);
    return r;
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
            s.zcr255[tid] = z * 255.0;
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

        // ---- power_to_db(ref=max, amin=1e-10, top_db=80) + normalize_matrix ---------------------
        const int total = n_mels * kFrames;
        float vmax = 0.f;
        for (int e = tid; e < total; e += kThreads) vmax = fmaxf(vmax, s.M[e / kFrames][e % kFrames]);
        vmax = block_reduce(s, vmax, true);
        // numpy-1.21 semantics: the reference term is evaluated in float64, then applied in float32
        const float ref_db = static_cast<float>(10.0 * log10(fmax(1e-10, static_cast<double>(vmax))));
        const float max_db = 10.0f * log10f(fmaxf(1e-10f, vmax)) - ref_db;
        const float floor_db = max_db - 80.0f;
        float vmin = max_db;
        for (int e = tid; e < total; e += kThreads) {
            float* q = &s.M[e / kFrames][e % kFrames];
            float db = 10.0f * log10f(fmaxf(1e-10f, *q)) - ref_db;
            db = fmaxf(db, floor_db);
            *q = db;
            vmin = fminf(vmin, db);
        }
        vmin = block_reduce(s, vmin, false);
        const float diff = max_db - vmin;
        __syncthreads();
        if (p.s_db) {
            float* dst = p.s_db + clip * static_cast<long long>(total);
            for (int e = tid; e < total; e += kThreads) dst[e] = s.M[e / kFrames][e % kFrames];
        }
        if (p.s_db_norm) {
            float* dst = p.s_db_norm + clip * static_cast<long long>(total);
            for (int e = tid; e < total; e += kThreads) dst[e] = (s.M[e / kFrames][e % kFrames] - vmin) / diff;
        }
        if (p.image) {
            // uint8 [n_mels][151][3], row r = mel (n_mels-1-r); value trunc(float64(v) * 255)
            uint32_t* dst = reinterpret_cast<uint32_t*>(p.image + clip * static_cast<long long>(total) * 3);
            const int words = total * 3 / 4;
            for (int wd = tid; wd < words; wd += kThreads) {
                uint32_t packed = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int byte = 4 * wd + b;
                    const int pix = byte / 3, c = byte - 3 * pix;
                    const int r = pix / kFrames, t = pix - r * kFrames;
                    double v255;
                    if (c == 0) {
                        v255 = s.zcr255[t];
                    } else {
                        const float nrm = (s.M[n_mels - 1 - r][t] - vmin) / diff;
                        v255 = static_cast<double>(1.0f - nrm) * 255.0;
                    }
                    // (x*255).astype(uint8): truncation; NaN (constant clip) maps to 0 here
                    const uint32_t q = (v255 >= 0.0) ? static_cast<uint32_t>(static_cast<int>(v255)) & 0xFFu : 0u;
                    packed |= q << (8 * b);
                }
                dst[wd] = packed;
            }
        }
        __syncthreads();
    }
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
    static MmlaPerDeviceOnce attr_once;                          // cudaFuncSetAttribute is per device
    const bool attr_set = !attr_once.first();
    if (!attr_set) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(overlap_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(Smem))));
    }
    long long grid = sms;
    if (grid > n_clips) grid = n_clips;
    overlap_features_kernel<<<static_cast<unsigned>(grid), kThreads, sizeof(Smem), st>>>(kp);
    mmla_count_launch("overlap_features_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    if (dev_tmp) MMLA_CUDA_CHECK(cudaFreeAsync(dev_tmp, st));
    return MMLA_OK;
}
