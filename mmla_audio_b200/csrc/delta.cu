// Standalone delta(feat, N): the reference's regression-delta helper
// (SpeakerIdentification/scripts/speaker_identification.py:141-151, duplicated at
// speaker_identification_post_processing.py:32-42) for callers that pass their own features.
// The fused MFCC kernel has its own in-smem copy of this for the hot path.
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) delta_kernel(const float* __restrict__ x, long long T, int D, int N,
                                                    float denom, float* __restrict__ out) {
    const long long total = T * D;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long t = e / D;
        const int c = static_cast<int>(e - t * D);
        float acc = 0.f;
        for (int k = -N; k <= N; ++k) {
            long long tt = t + k;
            tt = tt < 0 ? 0 : (tt > T - 1 ? T - 1 : tt);
            acc = fmaf(static_cast<float>(k), x[tt * D + c], acc);
        }
        out[e] = acc / denom;
    }
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int mmla_delta(const float* feat, int64_t n_frames, int32_t dim,
                                                                 int32_t N, float* out, void* stream) {
    MMLA_REQUIRE(feat && out, MMLA_EINVAL, "delta: null argument");
    MMLA_REQUIRE(n_frames >= 0 && dim >= 1 && N >= 1 && N <= 64, MMLA_EINVAL, "delta: bad shape/N");
    if (n_frames == 0) return MMLA_OK;
    const int sms = mmla_num_sms();
    MMLA_REQUIRE(sms > 0, MMLA_ECUDA, "delta: no CUDA device");
    float denom = 0.f;
    for (int i = 1; i <= N; ++i) denom += 2.f * i * i;
    long long grid = (n_frames * dim + 255) / 256;
    if (grid > 8LL * sms) grid = 8LL * sms;
    delta_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(feat, n_frames, dim, N, denom, out);
    mmla_count_launch("delta_kernel", static_cast<cudaStream_t>(stream));
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}
