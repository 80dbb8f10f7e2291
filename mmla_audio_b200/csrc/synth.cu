// On-device synthetic PCM generator (stands in for the reference's PyAudio capture,
// OverlapDetection/scripts/record_on_pc.py:115-124).  Integer-only; bit-identical twin of
// oracle/synth.py so any clip can be regenerated on the CPU for parity checks.
#include "common.cuh"

namespace {

constexpr uint32_t kIncPerHz = 268435u;   // floor(2^32 / 16000)
constexpr int kHarm = 8;

__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x85EBCA6Bu;
    x ^= x >> 13;
    x *= 0xC2B2AE35u;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t param(uint32_t key, uint32_t slot) { return hash32(key + slot * 0x9E3779B9u); }

struct Speaker {
    uint32_t inc, am_inc, am_ph0, onset, offset;
    int amp;
};

__global__ void __launch_bounds__(256) synth_kernel(int16_t* __restrict__ pcm, long long first_clip,
                                                    long long n_clips, int clip_len, long long clip_stride,
                                                    uint32_t seed, const int16_t* __restrict__ sine_table) {
    __shared__ int16_t tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = sine_table[i];
    __syncthreads();
    const uint32_t half = static_cast<uint32_t>(max(clip_len / 2, 1));
    for (long long c = blockIdx.x; c < n_clips; c += gridDim.x) {
        const unsigned long long clip = static_cast<unsigned long long>(first_clip + c);
        const uint32_t lo = static_cast<uint32_t>(clip), hi = static_cast<uint32_t>(clip >> 32);
        uint32_t key = hash32(seed + lo * 0x9E3779B9u);
        key = hash32(key ^ (hi * 0x85EBCA6Bu));
        const int nspk = 1 + static_cast<int>(param(key, 0) % 3u);
        Speaker sp[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const uint32_t b = 1 + 6 * s;
            sp[s].inc = (85u + param(key, b + 0) % 171u) * kIncPerHz;
            sp[s].am_inc = (3u + param(key, b + 1) % 4u) * kIncPerHz;
            sp[s].am_ph0 = param(key, b + 2);
            sp[s].onset = param(key, b + 3) % half;
            sp[s].offset = half + param(key, b + 4) % half;
            if (s == 0) sp[s].onset /= 4u;
            sp[s].amp = static_cast<int>(4000u + param(key, b + 5) % 6000u);
        }
        const uint32_t noise_key = param(key, 31);
        int16_t* dst = pcm + c * clip_stride;
        for (int n = threadIdx.x; n < clip_len; n += blockDim.x) {
            const uint32_t un = static_cast<uint32_t>(n);
            int acc = 0;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                if (s < nspk && un >= sp[s].onset && un < sp[s].offset) {
                    const uint32_t ph = sp[s].inc * un;
                    int hsum = 0;
#pragma unroll
                    for (int h = 1; h <= kHarm; ++h)
                        hsum += (static_cast<int>(tab[(ph * static_cast<uint32_t>(h)) >> 22]) * (32768 / h)) >> 15;
                    const uint32_t am_ph = sp[s].am_ph0 + sp[s].am_inc * un;
                    const int env = (static_cast<int>(tab[am_ph >> 22]) + 32768) >> 1;
                    long long v = (static_cast<long long>(hsum) * env) >> 15;
                    v = (v * sp[s].amp) >> 17;
                    acc += static_cast<int>(v);
                }
            }
            acc += static_cast<int>(hash32(noise_key + un * 0x9E3779B9u) & 0x3FFu) - 512;
            acc = max(-32768, min(32767, acc));
            dst[n] = static_cast<int16_t>(acc);
        }
    }
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int mmla_synth_pcm(int16_t* pcm, int64_t first_clip, int64_t n_clips, int32_t clip_len,
                              int64_t clip_stride, uint32_t seed, const int16_t* sine_table, void* stream) {
    MMLA_REQUIRE(pcm && sine_table, MMLA_EINVAL, "synth: null argument");
    MMLA_REQUIRE(n_clips >= 0 && clip_len >= 0 && clip_stride >= clip_len, MMLA_EINVAL, "synth: bad geometry");
    if (n_clips == 0 || clip_len == 0) return MMLA_OK;
    const int sms = mmla_num_sms();
    MMLA_REQUIRE(sms > 0, MMLA_ECUDA, "synth: no CUDA device");
    long long grid = 8LL * sms;
    if (grid > n_clips) grid = n_clips;
    synth_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pcm, first_clip, n_clips, clip_len, clip_stride, seed, sine_table);
    mmla_count_launch("synth_kernel", static_cast<cudaStream_t>(stream));
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}
