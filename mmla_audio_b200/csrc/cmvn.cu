// Cepstral mean / variance normalisation over the frames of each clip, in place.
// BASELINE.json north_star (3) lists CMVN next to the mel / log / DCT pass; the reference itself applies none
// (speaker features are raw MFCC + delta + delta-delta, speaker_identification.py:386-395), so
// this is an option that is off by default.  It needs the mean over ALL frames of a clip, i.e. a reduction across the
// frame tiles the MFCC kernels work on, so it is a second, HBM-bound pass over the cepstra (13 floats per frame).
#include "common.cuh"

namespace {
constexpr int kMaxDim = 64;

__global__ void __launch_bounds__(256) cmvn_kernel(float* __restrict__ feat, long long clip_stride, int row_stride, int dim,
                                                   int n_rows_uniform, const int32_t* __restrict__ n_rows_dev, int variance) {
    __shared__ float part[8][kMaxDim];
    __shared__ float mean[kMaxDim], inv_std[kMaxDim];
    const long long clip = blockIdx.x;
    const int T = n_rows_dev ? n_rows_dev[clip] : n_rows_uniform;
    if (T <= 0) return;
    float* f = feat + clip * clip_stride;
    const int lane = threadIdx.x & 31, phase = threadIdx.x >> 5;
    // pass 1: column means
    float s0 = 0.f, s1 = 0.f;
    for (int t = phase; t < T; t += 8) {
        const float* row = f + static_cast<long long>(t) * row_stride;
        if (lane < dim) s0 += row[lane];
        if (lane + 32 < dim) s1 += row[lane + 32];
    }
    part[phase][lane] = s0;
    part[phase][lane + 32] = s1;
    __syncthreads();
    if (threadIdx.x < kMaxDim) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
        mean[threadIdx.x] = s / static_cast<float>(T);
    }
    __syncthreads();
    // pass 2: population variance of the centred values (two-pass form: no cancellation)
    if (variance) {
        const float m0 = mean[lane], m1 = mean[lane + 32];
        s0 = 0.f; s1 = 0.f;
        for (int t = phase; t < T; t += 8) {
            const float* row = f + static_cast<long long>(t) * row_stride;
            if (lane < dim) { const float d = row[lane] - m0; s0 = fmaf(d, d, s0); }
            if (lane + 32 < dim) { const float d = row[lane + 32] - m1; s1 = fmaf(d, d, s1); }
        }
        __syncthreads();
        part[phase][lane] = s0;
        part[phase][lane + 32] = s1;
        __syncthreads();
        if (threadIdx.x < kMaxDim) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
            const float var = s / static_cast<float>(T);
            inv_std[threadIdx.x] = var > 0.f ? rsqrtf(var) : 1.f;     // constant column: centred only
        }
    } else if (threadIdx.x < kMaxDim) {
        inv_std[threadIdx.x] = 1.f;
    }
    __syncthreads();
    const float m0 = mean[lane], m1 = mean[lane + 32], i0 = inv_std[lane], i1 = inv_std[lane + 32];
    for (int t = phase; t < T; t += 8) {
        float* row = f + static_cast<long long>(t) * row_stride;
        if (lane < dim) row[lane] = (row[lane] - m0) * i0;
        if (lane + 32 < dim) row[lane + 32] = (row[lane + 32] - m1) * i1;
    }
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int mmla_cmvn(float* feat, int64_t n_clips, int64_t clip_stride,
                                                                int32_t row_stride, int32_t dim, int32_t n_rows,
                                                                const int32_t* n_rows_dev, int32_t variance, void* stream) {
    MMLA_REQUIRE(feat, MMLA_EINVAL, "cmvn: null feature pointer");
    MMLA_REQUIRE(dim >= 1 && dim <= kMaxDim && row_stride >= dim && clip_stride >= 0, MMLA_EINVAL, "cmvn: bad geometry (dim %d)", dim);
    MMLA_REQUIRE(n_clips >= 0 && n_clips < (1LL << 31), MMLA_EINVAL, "cmvn: bad clip count");
    MMLA_REQUIRE(n_rows_dev || n_rows >= 0, MMLA_EINVAL, "cmvn: bad row count");
    MMLA_REQUIRE(mmla_num_sms() > 0, MMLA_ECUDA, "cmvn: no CUDA device");
    if (n_clips == 0) return MMLA_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cmvn_kernel<<<static_cast<unsigned>(n_clips), 256, 0, st>>>(feat, clip_stride, row_stride, dim, n_rows, n_rows_dev, variance);
    mmla_count_launch("cmvn_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}
