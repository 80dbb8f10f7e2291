// Speaker-classifier stem on tcgen05 (TF32 mode): Conv1D(32, kernel 4, padding 'same') over the
// channel-padded feature tensor — the first layer of `res_model`
// (SpeakerIdentification/scripts/speaker_identification.py:196, `Conv1D(32, 4, padding='same')`).
//
//   x [B][256][40] fp32 (MFCC | delta | delta-delta | 0)  ->  y [B][256][32] fp32,
//   y[t] = b + sum_{j=0..3} W[j] x[t + j - 1]        (Keras 'same' for k = 4: one zero row before, two after)
//
// Same operand scheme as resunit_fused.cu: one CTA owns 128 consecutive time steps of one clip, x is read
// once (coalesced), rounded to TF32 and stored as the UMMA K-major no-swizzle "slab" operand
// [channel quad][row][16 B] with the 3 halo rows, so filter tap j is the SAME buffer read j rows further
// down — no im2col gather (the generic conv_tc kernel spent 0.18 ms per 4096 clips on this layer, 3x its
// HBM time).  The 20 KB of weights land by one TMA bulk copy; 20 tcgen05.mma (M=128, N=32, K=8) accumulate
// in 32 TMEM columns; the epilogue adds the bias and stores coalesced rows through a staging tile.
//
// FROM_CEP variant (the label pipeline, where the [256,39] feature tensor itself is not wanted): the input is the
// MFCC-13 rows straight from the feature kernel ([B][rows >= T][16] fp32) and the CTA builds its feature rows on the
// fly — delta and delta-delta exactly as the reference's `delta(feat, 2)` (speaker_identification.py:141-151, edge
// replicated inside the clip's T frames), zero rows from T to 256 (:391-395) — so the feature tensor never exists in
// HBM and the separate delta / padding pass disappears.
#include <string.h>

#include "conv_common.cuh"

namespace {

constexpr int kT = 256, kCin = 40, kCout = 32;
constexpr int kQuads = kCin / 4;                 // 10 channel quads
constexpr int kRows = 137;                       // 128 + 3 halo rows, padded to == 1 mod 8 (bank-friendly slab stride)
constexpr int kWBytes = 4 * kCin * kCout * 4;    // 20480: 40 K-slabs x [32][4] floats
constexpr int kThreadsStem = 256 + 32;           // warps 0..7 load / epilogue, warp 8 issues TMA + MMA
constexpr int kCepStride = 145;                  // 139 staged rows; 4*145 = 4 mod 32 spreads the transposing stores

struct StemSmem {
    alignas(128) unsigned char ab[kQuads * kRows * 16];     // operand slabs; later the output staging tile
    alignas(128) unsigned char w[kWBytes];
    alignas(16) float bias[kCout];
    // FROM_CEP, channel-major so that lanes = consecutive time rows read / write consecutive words:
    float cep[13][kCepStride];                                // cepstra of times t0-5 .. t0+133 (clamped into the clip)
    float dlt[13][kCepStride];                                // delta   of times t0-3 .. t0+131 (those inside the clip)
    alignas(8) uint64_t wfull, aready, done;
    uint32_t tmem_base;
};
static_assert(128 * (kCout + 4) * 4 <= kQuads * kRows * 16, "staging tile must fit in the operand buffer");

__device__ __forceinline__ uint32_t st_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void st_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    asm volatile("trap;");
}
__device__ __forceinline__ uint64_t st_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool st_elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

template <bool FROM_CEP>
__global__ void __launch_bounds__(kThreadsStem, 4) stem_fused_kernel(const float* __restrict__ x, const float* __restrict__ wg,
                                                                    const float* __restrict__ bias, float* __restrict__ y,
                                                                    int B, int n_frames, long long cep_clip_stride) {
    extern __shared__ unsigned char smem_dyn[];
    StemSmem& s = *reinterpret_cast<StemSmem*>(smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int clip = blockIdx.x >> 1, t0 = (blockIdx.x & 1) * 128;
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(kCout >> 3) << 17) |
                                (static_cast<uint32_t>(128 >> 4) << 24);   // f32 += tf32 x tf32, M=128, N=32

    // x rows t0-1 .. t0+129 (131 rows x 10 quads), issued before the setup barrier so HBM latency overlaps it
    constexpr int kLoads = (131 * kQuads + 255) / 256;   // 6 float4 per thread
    float4 v[kLoads];
    if (warp < 8) {
        if (!FROM_CEP) {
            const float* xc = x + static_cast<long long>(clip) * kT * kCin;
#pragma unroll
            for (int i = 0; i < kLoads; ++i) {
                const int idx = tid + i * 256;
                const int r = idx / kQuads, q = idx - r * kQuads;
                const int t = t0 - 1 + r;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < 131 && t >= 0 && t < kT) v[i] = *reinterpret_cast<const float4*>(xc + t * kCin + 4 * q);
            }
        } else {
            // cepstra rows of times t0-5 .. t0+133, clamped into the clip (the reference's edge replication); 16-float rows
            const float* cc = x + static_cast<long long>(clip) * cep_clip_stride;
            const int Tm1 = n_frames - 1;
#pragma unroll
            for (int i = 0; i < 3; ++i) {                 // 139 rows x 4 float4 = 556 loads
                const int idx = tid + i * 256;
                const int rr = idx >> 2, q4 = idx & 3;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rr < 139) {
                    int t = t0 - 5 + rr;
                    t = t < 0 ? 0 : (t > Tm1 ? Tm1 : t);
                    v[i] = *reinterpret_cast<const float4*>(cc + static_cast<long long>(t) * 16 + 4 * q4);
                }
            }
        }
        if (tid < kCout) s.bias[tid] = bias[tid];
    }
    if (tid == 0) {
        mbar_init(&s.wfull, 1);
        mbar_init(&s.aready, 1);
        mbar_init(&s.done, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;

    if (warp == 8) {
        // ================= weights by TMA, then the 20 MMAs (whole warp converged, one elected lane issues) =========
        if (lane == 0) {
            mbar_arrive_expect_tx(&s.wfull, kWBytes);
            tma_bulk_g2s(&s.w[0], wg, kWBytes, &s.wfull);
        }
        __syncwarp();
        st_wait(&s.wfull, 0u);
        st_wait(&s.aready, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t dA = st_desc(smem_u32(&s.ab[0]), kRows * 16, 128);
        const uint64_t dB = st_desc(smem_u32(&s.w[0]), kCout * 16, 128);
        if (st_elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 4; ++tap)
#pragma unroll
                for (int kq = 0; kq < kQuads / 2; ++kq) {
                    // K = 8 channels = slabs 2kq, 2kq+1 of the tap: A rows shifted down by `tap`, B slab tap*10 + 2kq
                    const uint64_t ad = dA + static_cast<uint64_t>(2 * kq * kRows + tap);
                    const uint64_t bd = dB + static_cast<uint64_t>((tap * kQuads + 2 * kq) * kCout);
                    const uint32_t acc = (tap | kq) != 0 ? 1u : 0u;
                    asm volatile(
                        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                        "l"(ad), "l"(bd), "r"(kIdesc), "r"(acc)
                        : "memory");
                }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s.done)) : "memory");
        }
        __syncwarp();
    } else {
        // ================= warps 0..7: operand slabs, epilogue =================
        if (!FROM_CEP) {
#pragma unroll
            for (int i = 0; i < kLoads; ++i) {
                const int idx = tid + i * 256;
                const int r = idx / kQuads, q = idx - r * kQuads;
                if (r < 131)
                    *reinterpret_cast<uint4*>(&s.ab[0] + (q * kRows + r) * 16) =
                        make_uint4(st_tf32(v[i].x), st_tf32(v[i].y), st_tf32(v[i].z), st_tf32(v[i].w));
            }
        } else {
            // (An earlier version looped over (row, channel) ELEMENTS with a division, five clamps and a three-way channel
            //  test each: 61 instructions per feature element, the kernel was issue-bound at 0.129 ms.  Here a thread owns
            //  a time row, clamps its neighbour indices once and walks the 13 channels with compile-time indices.)
            const int Tm1 = n_frames - 1;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int idx = tid + i * 256;
                const int rr = idx >> 2, q4 = idx & 3;
                if (rr < 139) {
                    const float vv[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (4 * q4 + u < 13) s.cep[4 * q4 + u][rr] = vv[u];
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // delta of times t0-3 .. t0+131 that lie inside the clip: row dd <-> time td = t0 - 3 + dd; the cep row of
            // time t is t - (t0 - 5).  Neighbour times are clamped into [0, T-1] (the reference's edge replication) and
            // |clamped - requested| <= 2, so they stay inside the staged window.  Work item = (row, channel half).
            for (int it = tid; it < 2 * 135; it += 256) {
                const int hi = it >= 135 ? 1 : 0;
                const int dd = it - 135 * hi;
                const int td = t0 - 3 + dd;
                if (td < 0 || td > Tm1) continue;         // never read: consumers index by times clamped into [0, T-1]
                const int base = 5 - t0;                  // cep row of time t = t + base
                const int m2 = max(td - 2, 0) + base, m1 = max(td - 1, 0) + base;
                const int p1 = min(td + 1, Tm1) + base, p2 = min(td + 2, Tm1) + base;
#pragma unroll
                for (int cc = 0; cc < 7; ++cc) {
                    const int c = cc + 7 * hi;            // channels [0,7) or [7,13)
                    if (c < 13) {
                        float acc = -2.f * s.cep[c][m2];
                        acc = fmaf(-1.f, s.cep[c][m1], acc);
                        acc = fmaf(1.f, s.cep[c][p1], acc);
                        acc = fmaf(2.f, s.cep[c][p2], acc);
                        s.dlt[c][dd] = acc * 0.1f;
                    }
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // feature rows of times t0-1 .. t0+129 -> TF32 slabs; rows outside [0, T) are the zero padding.
            // Work item = (row r, channel third): MFCC | delta | delta-delta, 13 channels each; a third's channels are
            // gathered into the 40-wide row with compile-time positions and stored as 16-byte slab entries.
            for (int it = tid; it < 3 * 131; it += 256) {
                const int third = it / 131;
                const int r = it - 131 * third;
                const int t = t0 - 1 + r;
                float f[13];
#pragma unroll
                for (int c = 0; c < 13; ++c) f[c] = 0.f;
                if (t >= 0 && t <= Tm1) {
                    if (third == 0) {
#pragma unroll
                        for (int c = 0; c < 13; ++c) f[c] = s.cep[c][t + 5 - t0];
                    } else if (third == 1) {
#pragma unroll
                        for (int c = 0; c < 13; ++c) f[c] = s.dlt[c][t + 3 - t0];
                    } else {
                        const int base = 3 - t0;              // dlt row of time t = t + base
                        const int m2 = max(t - 2, 0) + base, m1 = max(t - 1, 0) + base;
                        const int p1 = min(t + 1, Tm1) + base, p2 = min(t + 2, Tm1) + base;
#pragma unroll
                        for (int c = 0; c < 13; ++c) {
                            float acc = -2.f * s.dlt[c][m2];
                            acc = fmaf(-1.f, s.dlt[c][m1], acc);
                            acc = fmaf(1.f, s.dlt[c][p1], acc);
                            acc = fmaf(2.f, s.dlt[c][p2], acc);
                            f[c] = acc * 0.1f;
                        }
                    }
                }
                // channel ch = 13*third + c lives in slab quad ch/4, element ch%4: scalar 4-byte stores (lanes = rows,
                // 16 B apart: 4-way conflicts at worst, 39 stores per row instead of 10 gathers of mixed origin)
#pragma unroll
                for (int c = 0; c < 13; ++c) {
                    const int ch = 13 * third + c;        // third is warp-uniform except at the two seams
                    *reinterpret_cast<uint32_t*>(&s.ab[0] + ((ch >> 2) * kRows + r) * 16 + (ch & 3) * 4) = st_tf32(f[c]);
                }
                if (third == 2) *reinterpret_cast<uint32_t*>(&s.ab[0] + (9 * kRows + r) * 16 + 12) = 0u;   // channel 39
            }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s.aready)) : "memory");

        // epilogue: TMEM lane = output row; warps 0..3 take columns 0..15, warps 4..7 columns 16..31
        const int row = 32 * (warp & 3) + lane;
        const int col = (warp >> 2) * 16;
        constexpr int kStride = kCout + 4;
        float* stg = reinterpret_cast<float*>(&s.ab[0]);
        st_wait(&s.done, 0u);                                   // all MMAs retired: the operand buffer is dead
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + static_cast<uint32_t>(col)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(&s.bias[col + j]);
            *reinterpret_cast<float4*>(stg + row * kStride + col + j) =
                make_float4(__uint_as_float(r[j]) + bv.x, __uint_as_float(r[j + 1]) + bv.y, __uint_as_float(r[j + 2]) + bv.z,
                            __uint_as_float(r[j + 3]) + bv.w);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float* yc = y + (static_cast<long long>(clip) * kT + t0) * kCout;
#pragma unroll
        for (int i = 0; i < 128 * (kCout / 4) / 256; ++i) {      // 4 float4 per thread, coalesced rows
            const int idx = tid + i * 256;
            const int rr = idx >> 3, q = idx & 7;
            *reinterpret_cast<float4*>(yc + rr * kCout + 4 * q) = *reinterpret_cast<const float4*>(stg + rr * kStride + 4 * q);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
    (void)B;
}

}  // namespace

// x: [B][256][40] fp32 (channel 39 = 0), wg: the conv_tc-arranged weights of the padded stem (K = 160, N = 32),
// bias: [32], y: [B][256][32].
int mmla_launch_stem_fused(const float* x, const float* wg, const float* bias, float* y, long long B, cudaStream_t st) {
    MMLA_REQUIRE(B > 0 && B < (1LL << 22), MMLA_EINVAL, "stem_fused: bad batch");
    static MmlaPerDeviceOnce attr_once;                          // cudaFuncSetAttribute is per device
    const bool attr_set = !attr_once.first();
    const int smem = static_cast<int>(sizeof(StemSmem) + 128);
    if (!attr_set) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(stem_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    stem_fused_kernel<false><<<static_cast<unsigned>(2 * B), kThreadsStem, smem, st>>>(x, wg, bias, y, static_cast<int>(B), kT, 0);
    mmla_count_launch("stem_fused_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}

// cep: MFCC-13 rows [B][>= n_frames][16] fp32 (clip i at cep + i * cep_clip_stride), n_frames = psf frame count of every
// clip (1..256); the delta / delta-delta / zero-padding of the feature tensor happen inside the kernel.
int mmla_launch_stem_from_cepstra(const float* cep, long long cep_clip_stride, int n_frames, const float* wg, const float* bias,
                                  float* y, long long B, cudaStream_t st) {
    MMLA_REQUIRE(B > 0 && B < (1LL << 22), MMLA_EINVAL, "stem_fused: bad batch");
    MMLA_REQUIRE(n_frames >= 1 && n_frames <= kT, MMLA_EINVAL, "stem_fused: n_frames=%d must be in [1,256]", n_frames);
    MMLA_REQUIRE((cep_clip_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(cep) & 15) == 0, MMLA_EINVAL,
                 "stem_fused: cepstra rows must be 16-byte aligned");
    static MmlaPerDeviceOnce attr_once;                          // cudaFuncSetAttribute is per device
    const bool attr_set = !attr_once.first();
    const int smem = static_cast<int>(sizeof(StemSmem) + 128);
    if (!attr_set) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(stem_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    stem_fused_kernel<true><<<static_cast<unsigned>(2 * B), kThreadsStem, smem, st>>>(cep, wg, bias, y, static_cast<int>(B), n_frames,
                                                                                        cep_clip_stride);
    mmla_count_launch("stem_delta_fused_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}
