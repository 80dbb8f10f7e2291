// Label histogram — the counting loops of the reference's visualization() functions
// (OverlapDetection/scripts/overlap_degree_distribution.py:49-61,
//  SpeakerIdentification/scripts/speaker_time_distribution.py:52-80) as one device pass.
// Per-CTA shared-memory bins, one global 64-bit atomic per (CTA, bin).
#include "common.cuh"

namespace {
constexpr int kMaxBins = 1024;

__global__ void __launch_bounds__(256) tally_kernel(const int32_t* __restrict__ labels, long long n,
                                                    int n_classes, unsigned long long* __restrict__ counts) {
    __shared__ unsigned int bins[kMaxBins + 1];
    const int nb = n_classes + 1;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) bins[i] = 0u;
    __syncthreads();
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int l = labels[i];
        atomicAdd(&bins[(l >= 0 && l < n_classes) ? l : n_classes], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
        if (bins[i]) atomicAdd(&counts[i], static_cast<unsigned long long>(bins[i]));
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int mmla_tally(const int32_t* labels, int64_t n, int32_t n_classes, int64_t* counts, void* stream) {
    MMLA_REQUIRE(counts != nullptr && (labels != nullptr || n == 0), MMLA_EINVAL, "tally: null argument");
    MMLA_REQUIRE(n >= 0 && n_classes >= 1 && n_classes <= kMaxBins, MMLA_EINVAL, "tally: n_classes must be in [1,%d]", kMaxBins);
    if (n == 0) return MMLA_OK;
    const int sms = mmla_num_sms();
    MMLA_REQUIRE(sms > 0, MMLA_ECUDA, "tally: no CUDA device");
    // each CTA's 32-bit shared bins see at most n/grid + 256 labels; keep that below 2^32
    long long grid = (n + 255) / 256;
    if (grid > 4LL * sms) grid = 4LL * sms;
    tally_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        labels, n, n_classes, reinterpret_cast<unsigned long long*>(counts));
    mmla_count_launch("tally_kernel", static_cast<cudaStream_t>(stream));
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}
