// Stationary spectral-gating noise reduction on the device.
//
// Replaces, for batches of clips resident in HBM,
//   noise, sr = librosa.load(NOISE_PATH, sr=None); y, sr = librosa.load(filepath, sr=None)
//   noise_reduced_wav = nr.reduce_noise(y_noise=noise, y=y, sr=sr, stationary=True)
//   sf.write(filepath, noise_reduced_wav, 16000)
//     OverlapDetection/scripts/record_on_pc.py:208-212 (save_wave_file), :127-132;
//     overlap_detection_post_processing.py:128-133 (standardize_audio); the SpeakerIdentification copies.
// `noisereduce` (2.x, un-vendored) with every other argument at its default: n_fft = win_length = 1024, hop = 256,
// scipy.signal.stft / istft (periodic Hann, boundary='zeros', padded=False, 'spectrum' scaling), dB = 20 log10(|Z| + eps)
// floored at (row max - 80), threshold per bin = mean + 1.5 std of the noise dB over time, mask = dB > threshold
// (prop_decrease = 1), mask smoothed by the normalised outer product of two triangles (500 Hz -> 16 bins, 50 ms -> 3
// frames: 33 x 7 taps, fftconvolve 'same'), Z * mask, istft, the clip cut back out of its 30000-sample zero padding
// (chunk padding), written as PCM_16 (libsndfile: lrint(y * 32767)).
//
// Layout: the clip sits at sample 30000 of a zero chunk, so only the frames whose window touches the clip carry
// signal (frames kFirst .. kFirst + F - 1); every other frame of the chunk is all-zero, its dB sits on the floor, and
// its mask value is the same for every such frame ("edge mask" = floor > threshold).  Kernels:
//   ng_stft_kernel     one CTA per (frame, clip): window, 1024-point Stockham radix-4 FFT in shared memory, spectrum + dB
//   ng_profile_kernel  noise statistics per bin -> threshold (run once per noise recording)
//   ng_mask_kernel     per (clip, bin): row max -> floor -> 0/1 mask and the edge mask
//   ng_smooth_kernel   separable 33-tap (frequency) x 7-tap (time) triangle
//   ng_istft_kernel    Z * mask -> inverse FFT -> window
//   ng_ola_kernel      overlap-add / sum of squared windows -> int16
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kN = 1024, kHop = 256, kBins = 513;
constexpr int kPad = 30000;                              // noisereduce chunk padding
constexpr int kFirst = (kPad - kN / 2) / kHop + 1;       // first frame whose window reaches the clip: 116
constexpr float kEps = 2.220446049250313e-16f;
constexpr float kWinSum = 512.f;                         // sum of the periodic Hann window of length 1024
constexpr int kNf = 16, kNt = 3;                         // smoothing half-widths (bins, frames)

__host__ __device__ inline int frames_for_clip(int len) {
    if (len <= 0) return 0;
    const int last = (kPad + len - 1 + kN / 2) / kHop;   // last frame whose window reaches the clip
    return last - kFirst + 1;
}

// 1024-point complex FFT, Stockham autosort, radix 4, 256 threads, two shared buffers of 1024 float2.
// tw[k] = exp(-2 pi i k / 1024).  INV: conjugate transform (no 1/N scaling).  The result is in `a`.
template <bool INV>
__device__ __forceinline__ void fft1024(float2* a, float2* b, const float2* __restrict__ tw, int tid) {
    float2* src = a;
    float2* dst = b;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int ns = 1 << (2 * s);
        const int k = tid & (ns - 1);
        float2 v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            v[r] = src[tid + r * 256];
            if (r > 0 && s > 0) {
                float2 w = tw[(r * k * (256 >> (2 * s))) & 1023];                // exp(-2 pi i r k / (4 ns))
                if (INV) w.y = -w.y;
                v[r] = make_float2(v[r].x * w.x - v[r].y * w.y, v[r].x * w.y + v[r].y * w.x);
            }
        }
        const float2 a0 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y), a1 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
        const float2 a2 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
        const float2 d = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
        const float2 a3 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);   // (v1 - v3) * (+i | -i)
        const int j0 = ((tid - k) << 2) + k;
        dst[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
        dst[j0 + ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
        dst[j0 + 2 * ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
        dst[j0 + 3 * ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
        __syncthreads();
        float2* t = src;
        src = dst;
        dst = t;
    }
    // five passes: the result sits in `b`; copy back so callers always read `a`
    for (int i = tid; i < kN; i += 256) a[i] = b[i];
    __syncthreads();
}

struct NgTables {
    float2 tw[kN];
    float win[kN];
};

// frame m of a signal with scipy's boundary='zeros': samples [256 m - 512 - origin, ...) where `origin` is the offset of
// sample 0 of `x` inside the (virtual) zero-extended chunk.
__global__ void __launch_bounds__(256) ng_stft_kernel(const int16_t* __restrict__ pcm, long long clip_stride, int clip_len,
                                                      const int32_t* __restrict__ clip_len_dev, int origin, int first_frame,
                                                      int frames_alloc, const NgTables* __restrict__ tab,
                                                      float2* __restrict__ spec, float* __restrict__ db) {
    __shared__ float2 a[kN], b[kN];
    __shared__ float2 tw[kN];
    const int tid = threadIdx.x;
    const long long c = blockIdx.y;
    const int len = clip_len_dev ? clip_len_dev[c] : clip_len;
    const int m = blockIdx.x;
    const int16_t* x = pcm + c * clip_stride;
    const int start = kHop * (first_frame + m) - kN / 2 - origin;
    for (int i = tid; i < kN; i += 256) {
        tw[i] = tab->tw[i];
        const int n = start + i;
        const float v = (n >= 0 && n < len) ? static_cast<float>(x[n]) * (1.f / 32768.f) : 0.f;
        a[i] = make_float2(v * tab->win[i], 0.f);
    }
    __syncthreads();
    fft1024<false>(a, b, tw, tid);
    const long long base = (c * frames_alloc + m) * kBins;
    for (int k = tid; k < kBins; k += 256) {
        const float2 z = a[k];
        if (spec) spec[base + k] = z;
        const float mag = sqrtf(z.x * z.x + z.y * z.y) * (1.f / kWinSum);
        db[base + k] = 20.f * log10f(mag + kEps);
    }
}

// threshold[k] = mean_t + 1.5 std_t of the noise dB, after flooring every row at (row max - 80)
__global__ void __launch_bounds__(128) ng_profile_kernel(const float* __restrict__ db, int frames, float n_std, float* __restrict__ thresh) {
    __shared__ float red[128];
    const int k = blockIdx.x, tid = threadIdx.x;
    float mx = -INFINITY;
    for (int m = tid; m < frames; m += 128) mx = fmaxf(mx, db[static_cast<long long>(m) * kBins + k]);
    red[tid] = mx;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if (tid < s) red[tid] = fmaxf(red[tid], red[tid + s]);
        __syncthreads();
    }
    const float floor_db = red[0] - 80.f;
    __syncthreads();
    float sum = 0.f;
    for (int m = tid; m < frames; m += 128) sum += fmaxf(db[static_cast<long long>(m) * kBins + k], floor_db);
    red[tid] = sum;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    const float mean = red[0] / frames;
    __syncthreads();
    float ss = 0.f;
    for (int m = tid; m < frames; m += 128) {
        const float d = fmaxf(db[static_cast<long long>(m) * kBins + k], floor_db) - mean;
        ss = fmaf(d, d, ss);
    }
    red[tid] = ss;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        __syncthreads();
    }
    if (tid == 0) thresh[k] = mean + n_std * sqrtf(red[0] / frames);
}

// mask[c][m][k] = floor(db) > thresh, in place over db; edge[c][k] = the mask of an all-zero frame of this clip's chunk
__global__ void __launch_bounds__(256) ng_mask_kernel(float* __restrict__ db, int frames_alloc, int clip_len,
                                                      const int32_t* __restrict__ clip_len_dev, const float* __restrict__ thresh,
                                                      float* __restrict__ edge) {
    const long long c = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= kBins) return;
    const int F = frames_for_clip(clip_len_dev ? clip_len_dev[c] : clip_len);
    float* col = db + c * frames_alloc * kBins + k;
    const float zero_db = 20.f * log10f(kEps);
    float mx = zero_db;                                                      // the chunk always contains all-zero frames
    for (int m = 0; m < F; ++m) mx = fmaxf(mx, col[static_cast<long long>(m) * kBins]);
    const float floor_db = mx - 80.f, th = thresh[k];
    for (int m = 0; m < F; ++m) col[static_cast<long long>(m) * kBins] = fmaxf(col[static_cast<long long>(m) * kBins], floor_db) > th ? 1.f : 0.f;
    edge[c * kBins + k] = fmaxf(zero_db, floor_db) > th ? 1.f : 0.f;
}

// smoothed[c][m][k] = sum_{df, dt} tri16(df) tri3(dt) mask[m + dt][k + df] / 68   (zero outside the bin range; frames
// outside [0, F) are all-zero frames of the chunk and carry the edge mask)
__global__ void __launch_bounds__(256) ng_smooth_kernel(const float* __restrict__ mask, const float* __restrict__ edge,
                                                        int frames_alloc, int clip_len, const int32_t* __restrict__ clip_len_dev,
                                                        float* __restrict__ out) {
    const long long c = blockIdx.z;
    const int m = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int F = frames_for_clip(clip_len_dev ? clip_len_dev[c] : clip_len);
    if (k >= kBins || m >= F) return;
    const float* mk = mask + c * frames_alloc * kBins;
    const float* ed = edge + c * kBins;
    float acc = 0.f;
#pragma unroll
    for (int dt = -kNt; dt <= kNt; ++dt) {
        const int mm = m + dt;
        const float wt = static_cast<float>(kNt + 1 - (dt < 0 ? -dt : dt)) * (1.f / (kNt + 1));
        const float* row = (mm >= 0 && mm < F) ? mk + static_cast<long long>(mm) * kBins : ed;
        float rs = 0.f;
        for (int df = -kNf; df <= kNf; ++df) {
            const int kk = k + df;
            if (kk < 0 || kk >= kBins) continue;
            rs = fmaf(static_cast<float>(kNf + 1 - (df < 0 ? -df : df)) * (1.f / (kNf + 1)), row[kk], rs);
        }
        acc = fmaf(wt, rs, acc);
    }
    out[(c * frames_alloc + m) * kBins + k] = acc * (1.f / ((kNf + 1) * (kNt + 1)));
}

__global__ void __launch_bounds__(256) ng_istft_kernel(const float2* __restrict__ spec, const float* __restrict__ smask,
                                                       int frames_alloc, int clip_len, const int32_t* __restrict__ clip_len_dev,
                                                       const NgTables* __restrict__ tab, float* __restrict__ frames_out) {
    __shared__ float2 a[kN], b[kN];
    __shared__ float2 tw[kN];
    const int tid = threadIdx.x;
    const long long c = blockIdx.y;
    const int m = blockIdx.x;
    const int F = frames_for_clip(clip_len_dev ? clip_len_dev[c] : clip_len);
    if (m >= F) return;
    const long long base = (c * frames_alloc + m) * kBins;
    for (int i = tid; i < kN; i += 256) tw[i] = tab->tw[i];
    for (int k = tid; k < kBins; k += 256) {
        float2 z = spec[base + k];
        const float g = smask[base + k];
        z.x *= g;
        z.y = (k == 0 || k == kN / 2) ? 0.f : z.y * g;                       // irfft ignores the imaginary part of DC / Nyquist
        a[k] = z;
        if (k > 0 && k < kN / 2) a[kN - k] = make_float2(z.x, -z.y);
    }
    __syncthreads();
    fft1024<true>(a, b, tw, tid);
    float* o = frames_out + (c * frames_alloc + m) * kN;
    for (int i = tid; i < kN; i += 256) o[i] = a[i].x * (1.f / kN) * tab->win[i];
}

// y[n] = sum_m w * irfft(...) / sum_m w^2 over the (up to four) frames covering chunk sample 30000 + n, then PCM_16
__global__ void __launch_bounds__(256) ng_ola_kernel(const float* __restrict__ frames_out, int frames_alloc, int clip_len,
                                                     const int32_t* __restrict__ clip_len_dev, const NgTables* __restrict__ tab,
                                                     int16_t* __restrict__ out, long long out_stride) {
    const long long c = blockIdx.y;
    const int len = clip_len_dev ? clip_len_dev[c] : clip_len;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= len) return;
    const int F = frames_for_clip(len);
    const int p = kPad + n + kN / 2;                                        // coordinate in the boundary-padded chunk
    const int m_hi = p / kHop, m_lo = (p - (kN - 1) + kHop - 1) / kHop;
    const float* fo = frames_out + c * frames_alloc * kN;
    float acc = 0.f, norm = 0.f;
    for (int m = m_lo; m <= m_hi; ++m) {
        const int i = p - kHop * m;
        const float w = tab->win[i];
        norm = fmaf(w, w, norm);                                             // every chunk frame counts, zero or not
        const int mm = m - kFirst;
        if (mm >= 0 && mm < F) acc += fo[static_cast<long long>(mm) * kN + i];
    }
    const float y = acc / (norm > 1e-10f ? norm : 1.f);
    const float q = rintf(y * 32767.f);                                      // libsndfile float -> PCM_16
    out[c * out_stride + n] = static_cast<int16_t>(fminf(fmaxf(q, -32768.f), 32767.f));
}

NgTables* g_tables[64] = {};

int get_tables(const NgTables** out) {
    int dev = 0;
    MMLA_CUDA_CHECK(cudaGetDevice(&dev));
    MMLA_REQUIRE(dev >= 0 && dev < 64, MMLA_EUNSUP, "noise_gate: device ordinal out of range");
    if (!g_tables[dev]) {
        NgTables* h = new NgTables;
        const double PI = 3.14159265358979323846;
        for (int i = 0; i < kN; ++i) {
            h->tw[i] = make_float2(static_cast<float>(cos(2.0 * PI * i / kN)), static_cast<float>(-sin(2.0 * PI * i / kN)));
            h->win[i] = static_cast<float>(0.5 - 0.5 * cos(2.0 * PI * i / kN));     // periodic Hann (scipy get_window('hann', N))
        }
        NgTables* d = nullptr;
        cudaError_t e = cudaMalloc(&d, sizeof(NgTables));
        if (e == cudaSuccess) e = cudaMemcpy(d, h, sizeof(NgTables), cudaMemcpyHostToDevice);
        delete h;
        if (e != cudaSuccess) {
            mmla_set_error("noise_gate: table upload failed: %s", cudaGetErrorString(e));
            return MMLA_ECUDA;
        }
        g_tables[dev] = d;
    }
    *out = g_tables[dev];
    return MMLA_OK;
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int mmla_noise_profile(const int16_t* noise, int64_t n_samples, float n_std_thresh,
                                                                         float* thresh_out, void* stream) {
    MMLA_REQUIRE(noise && thresh_out, MMLA_EINVAL, "noise_profile: null argument");
    MMLA_REQUIRE(n_samples >= 1, MMLA_EINVAL, "noise_profile: empty noise recording");
    MMLA_REQUIRE(mmla_num_sms() > 0, MMLA_ECUDA, "noise_profile: no CUDA device");
    if (n_samples > 600000) n_samples = 600000;                             // clip_noise_stationary: y_noise[:chunk_size]
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const NgTables* tab = nullptr;
    int rc = get_tables(&tab);
    if (rc != MMLA_OK) return rc;
    const int frames = static_cast<int>((n_samples + kHop) / kHop);         // (n + 2*512 - 768) // 256
    float* db = nullptr;
    MMLA_CUDA_CHECK(cudaMallocAsync(&db, static_cast<size_t>(frames) * kBins * sizeof(float), st));
    ng_stft_kernel<<<dim3(frames, 1), 256, 0, st>>>(noise, 0, static_cast<int>(n_samples), nullptr, 0, 0, frames, tab, nullptr, db);
    mmla_count_launch("ng_stft_kernel", st);
    ng_profile_kernel<<<kBins, 128, 0, st>>>(db, frames, n_std_thresh, thresh_out);
    mmla_count_launch("ng_profile_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    MMLA_CUDA_CHECK(cudaFreeAsync(db, st));
    return MMLA_OK;
}

extern "C" __attribute__((visibility("default"))) int mmla_noise_gate(const int16_t* pcm, int64_t n_clips, int32_t clip_len,
                                                                      int64_t clip_stride, const int32_t* clip_len_dev,
                                                                      const float* thresh, int16_t* out, int64_t out_stride,
                                                                      void* stream) {
    MMLA_REQUIRE(pcm && thresh && out, MMLA_EINVAL, "noise_gate: null argument");
    MMLA_REQUIRE(n_clips >= 0 && clip_len >= 0 && clip_stride >= 0 && out_stride >= clip_len, MMLA_EINVAL, "noise_gate: bad geometry");
    MMLA_REQUIRE(clip_len <= 600000, MMLA_EUNSUP, "noise_gate: clips longer than one noisereduce chunk (600000 samples) are not supported");
    MMLA_REQUIRE(mmla_num_sms() > 0, MMLA_ECUDA, "noise_gate: no CUDA device");
    if (n_clips == 0 || clip_len == 0) return MMLA_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const NgTables* tab = nullptr;
    int rc = get_tables(&tab);
    if (rc != MMLA_OK) return rc;
    const int F = frames_for_clip(clip_len);
    // scratch per clip: spectrum F x 513 float2, mask + smoothed mask F x 513 float each, edge 513, frames F x 1024 float
    const size_t per_clip = static_cast<size_t>(F) * (kBins * (8 + 4 + 4) + kN * 4) + kBins * 4;
    long long micro = static_cast<long long>((3ull << 30) / per_clip);       // <= 3 GB of scratch at a time
    if (micro < 1) micro = 1;
    if (micro > n_clips) micro = n_clips;
    if (micro > 65535) micro = 65535;
    char* ws = nullptr;
    MMLA_CUDA_CHECK(cudaMallocAsync(&ws, per_clip * micro, st));
    float2* spec = reinterpret_cast<float2*>(ws);
    float* mask = reinterpret_cast<float*>(spec + micro * F * kBins);
    float* smask = mask + micro * F * kBins;
    float* fout = smask + micro * F * kBins;
    float* edge = fout + micro * F * kN;
    for (long long c0 = 0; c0 < n_clips; c0 += micro) {
        const int nb = static_cast<int>(std::min<long long>(micro, n_clips - c0));
        const int16_t* x = pcm + c0 * clip_stride;
        const int32_t* ld = clip_len_dev ? clip_len_dev + c0 : nullptr;
        ng_stft_kernel<<<dim3(F, nb), 256, 0, st>>>(x, clip_stride, clip_len, ld, kPad, kFirst, F, tab, spec, mask);
        mmla_count_launch("ng_stft_kernel", st);
        ng_mask_kernel<<<dim3((kBins + 255) / 256, nb), 256, 0, st>>>(mask, F, clip_len, ld, thresh, edge);
        mmla_count_launch("ng_mask_kernel", st);
        ng_smooth_kernel<<<dim3((kBins + 255) / 256, F, nb), 256, 0, st>>>(mask, edge, F, clip_len, ld, smask);
        mmla_count_launch("ng_smooth_kernel", st);
        ng_istft_kernel<<<dim3(F, nb), 256, 0, st>>>(spec, smask, F, clip_len, ld, tab, fout);
        mmla_count_launch("ng_istft_kernel", st);
        ng_ola_kernel<<<dim3((clip_len + 255) / 256, nb), 256, 0, st>>>(fout, F, clip_len, ld, tab, out + c0 * out_stride, out_stride);
        mmla_count_launch("ng_ola_kernel", st);
        MMLA_CUDA_CHECK(cudaGetLastError());
    }
    MMLA_CUDA_CHECK(cudaFreeAsync(ws, st));
    return MMLA_OK;
}
