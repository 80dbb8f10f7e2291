// LSTM input projections on tcgen05 (TF32 mode): xp[d] = seq . W_in[d] + b[d] for both directions of
// `Bidirectional(LSTM(256))` (OverlapDetection/scripts/overlap_detector_temp.py:297,
// SpeakerIdentification/scripts/speaker_identification.py:213) — the [B*T, 128] x [128, 1024] products whose
// result `lstm_fused_kernel` adds to the recurrent pre-activations at every step.
//
// Output layout ("row-tiled", shared with lstm_fused.cu): xp[d][clip tile bt][t][column quad cq][row r][4], clip
// b = 128 bt + r, column = 4 cq + e in Keras order i|f|c|o.  Both kernels hold one clip per thread (TMEM lane = row),
// so with this layout a warp's 128-bit access covers 32 consecutive rows = 512 contiguous bytes, whereas in the
// natural [b][t][1024] layout the 32 lanes touch 32 different lines (the LSTM's gate epilogue spent 17k of its 60k
// cycles per step in exactly those loads).  The buffer is sized for whole 128-clip tiles.
//
// One CTA = (128 clips x one time step of seq, direction):
//   * the A tile [128 x 128] is read once (coalesced), rounded to TF32 and kept in shared memory for the whole
//     CTA as four 128 x 32 SWIZZLE_128B sub-tiles (the same operand form lstm_fused.cu uses for h);
//   * W_in streams from L2 through a 4-stage TMA ring of 16 KB chunks (host pre-arranged, TF32 pre-rounded):
//     four passes of 256 output columns, 16 tcgen05.mma (M=128, N=256, K=8) each, accumulating in TMEM; the two
//     256-column TMEM halves ping-pong so the MMAs of pass p+1 overlap the epilogue of pass p;
//   * the epilogue adds the bias and stores straight from registers (coalesced by construction, see above).
// The generic implicit-GEMM kernel (conv_tc.cu, built for gathered conv operands) needed 0.075 ms per direction
// for 4096 clips; this path is bound by the 134 MB it writes.
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kK = 128;                       // input features
constexpr int kN = 1024;                      // 4 gates x 256 units
constexpr int kSub = 128 * 128;               // bytes of one 128 x 32 TF32 sub-tile
constexpr int kChunkFloats = 4 * 256 * 4;     // K = 16 x 256 columns
constexpr int kChunkBytes = kChunkFloats * 4; // 16 KB
constexpr int kChunksPerPass = kK / 16;       // 8
constexpr int kPasses = kN / 256;             // 4
constexpr int kStages = 4;
constexpr int kEpi = 256;
constexpr int kProducers = 2;                 // TMA producer warps, alternate chunks (see lstm_fused.cu)
constexpr int kThreads = kEpi + 32 + 32 * kProducers;   // warp 8: MMA issuer, warps 9..: TMA producers

struct XpSmem {
    alignas(1024) unsigned char A[4][kSub];
    alignas(128) unsigned char ring[kStages][kChunkBytes];
    alignas(16) float bias[kN];
    alignas(8) uint64_t full[kStages], empty[kStages], tfull[2], tempty[2], aready;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t xp_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void xp_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    asm volatile("trap;");
}
__device__ __forceinline__ void xp_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool xp_elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t xp_desc_sw128(uint32_t addr) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (1ull << 16) | (static_cast<uint64_t>(1024u >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t xp_desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void xp_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct XpArgs {
    const float* seq;       // [B][T][128]
    const float* w[2];      // arranged chunk streams, kPasses * kChunksPerPass chunks each
    const float* b[2];      // [1024]
    float* xp[2];           // row-tiled, ceil(B/128) x T x 256 x 128 x 4 floats
    long long B;
    int T;
};

__global__ void __launch_bounds__(kThreads, 1) xproj_fused_kernel(const XpArgs a) {
    extern __shared__ unsigned char smem_dyn[];
    XpSmem& s = *reinterpret_cast<XpSmem*>(smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const long long bt = blockIdx.x / a.T;                        // clip tile
    const int t = static_cast<int>(blockIdx.x - bt * a.T);        // time step
    const long long b0 = bt * 128;
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(256 >> 3) << 17) |
                                (static_cast<uint32_t>(128 >> 4) << 24);   // f32 += tf32 x tf32, M=128, N=256

    // A tile loads first (HBM latency overlaps the setup barrier): 128 rows x 32 float4, 16 per thread
    float4 v[16];
    if (warp < 8) {
        const int q = tid & 7, rb = tid >> 3;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int kc = j >> 2, r = rb + 32 * (j & 3);
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b0 + r < a.B) v[j] = *reinterpret_cast<const float4*>(a.seq + ((b0 + r) * a.T + t) * kK + 32 * kc + 4 * q);
        }
        for (int i = tid; i < kN; i += kEpi) s.bias[i] = a.b[dir][i];
    }
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.tfull[i], 1);
            mbar_init(&s.tempty[i], 8);
        }
        mbar_init(&s.aready, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;
    constexpr int kTotalChunks = kPasses * kChunksPerPass;

    if (warp > 8) {
        // two producer warps, alternate chunks (a single thread gets one bulk copy through per ~420 cycles: 39 B/clk
        // at 16 KB, and the MMAs want 64 B/clk; see lstm_fused.cu / scripts/microbench/tma_stream.cu)
        if (lane == 0) {
            for (int g = warp - 9; g < kTotalChunks; g += kProducers) {
                const int stg = g % kStages, use = g / kStages;
                if (use > 0) xp_wait(&s.empty[stg], static_cast<uint32_t>((use - 1) & 1));
                mbar_arrive_expect_tx(&s.full[stg], kChunkBytes);
                tma_bulk_g2s(&s.ring[stg][0], a.w[dir] + static_cast<long long>(g) * kChunkFloats, kChunkBytes, &s.full[stg]);
            }
        }
        __syncwarp();
    } else if (warp == 8) {
        // whole warp converged, one elected lane issues (uniform-register descriptors)
        const uint64_t dB0 = xp_desc_noswz(smem_u32(&s.ring[0][0]), 256 * 16, 128);
        constexpr uint32_t kStageUnits = kChunkBytes / 16;
        xp_wait(&s.aready, 0u);
        // two chunks (4 MMAs) per trip, ring position as counters: see the MMA loop of lstm_fused.cu
        static_assert(kStages == 4, "chunk pairs use stages {0,1} and {2,3}");
        const uint64_t dA0 = xp_desc_sw128(smem_u32(&s.A[0][0]));
        int stg = 0;
        uint32_t par = 0;
        for (int pass = 0; pass < kPasses; ++pass) {
            const int buf = pass & 1;
            if (pass >= 2) xp_wait(&s.tempty[buf], static_cast<uint32_t>(((pass >> 1) - 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t dcol = tmem + static_cast<uint32_t>(buf * 256);
            uint64_t dA = dA0;
            for (int kc = 0; kc < 4; ++kc, dA += static_cast<uint64_t>(kSub / 16)) {
                const uint64_t bd0 = dB0 + static_cast<uint64_t>(stg * kStageUnits);
                xp_wait(&s.full[stg], par);
                xp_wait(&s.full[stg + 1], par);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (xp_elect_one()) {
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const uint64_t ad = dA + static_cast<uint64_t>(m * 2);                       // 32 B per K=8 step
                        const uint64_t bd = bd0 + static_cast<uint64_t>((m >> 1) * kStageUnits + (m & 1) * 2 * 256);
                        const uint32_t acc = (kc | m) != 0 ? 1u : 0u;
                        asm volatile(
                            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(dcol),
                            "l"(ad), "l"(bd), "r"(kIdesc), "r"(acc)
                            : "memory");
                    }
                    xp_commit(&s.empty[stg]);
                    xp_commit(&s.empty[stg + 1]);
                    if (kc == 3) xp_commit(&s.tfull[buf]);
                }
                stg ^= 2;
                if (stg == 0) par ^= 1u;
            }
        }
    } else {
        // ---- A operand: TF32, SWIZZLE_128B (16-byte chunk index XOR row % 8) ----
        {
            const int q = tid & 7, rb = tid >> 3;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int kc = j >> 2, r = rb + 32 * (j & 3);
                *reinterpret_cast<uint4*>(&s.A[kc][0] + r * 128 + ((q ^ (r & 7)) << 4)) =
                    make_uint4(xp_tf32(v[j].x), xp_tf32(v[j].y), xp_tf32(v[j].z), xp_tf32(v[j].w));
            }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) xp_arrive(&s.aready);

        const int quarter = warp & 3, chalf = warp >> 2;          // TMEM lanes 32*quarter..; 128-column half of a pass
        const int row = 32 * quarter + lane;
        const bool row_ok = b0 + row < a.B;
        float* out = a.xp[dir] + (bt * a.T + t) * (256LL * 512) + row * 4;
        for (int pass = 0; pass < kPasses; ++pass) {
            const int buf = pass & 1;
            xp_wait(&s.tfull[buf], static_cast<uint32_t>((pass >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int grp = 0; grp < 4; ++grp) {                   // 32 columns at a time: two loads in flight, one wait
                const int c0 = 128 * chalf + 32 * grp;            // column of the pass this thread reads
                uint32_t r[32];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                        : "=r"(r[16 * hh + 0]), "=r"(r[16 * hh + 1]), "=r"(r[16 * hh + 2]), "=r"(r[16 * hh + 3]), "=r"(r[16 * hh + 4]),
                          "=r"(r[16 * hh + 5]), "=r"(r[16 * hh + 6]), "=r"(r[16 * hh + 7]), "=r"(r[16 * hh + 8]), "=r"(r[16 * hh + 9]),
                          "=r"(r[16 * hh + 10]), "=r"(r[16 * hh + 11]), "=r"(r[16 * hh + 12]), "=r"(r[16 * hh + 13]),
                          "=r"(r[16 * hh + 14]), "=r"(r[16 * hh + 15])
                        : "r"(tmem + (static_cast<uint32_t>(32 * quarter) << 16) + static_cast<uint32_t>(buf * 256 + c0 + 16 * hh)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (grp == 3) {                                   // this warp has drained the pass: MMAs of pass+2 may start
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) xp_arrive(&s.tempty[buf]);
                }
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const int col = 256 * pass + c0 + j;
                        const float4 bv = *reinterpret_cast<const float4*>(&s.bias[col]);
                        *reinterpret_cast<float4*>(out + (col >> 2) * 512) =
                            make_float4(__uint_as_float(r[j]) + bv.x, __uint_as_float(r[j + 1]) + bv.y,
                                        __uint_as_float(r[j + 2]) + bv.z, __uint_as_float(r[j + 3]) + bv.w);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace

// Host: arrange W [128][1024] (Keras LSTM kernel, columns i|f|c|o) into the chunk stream: chunk (pass, kc, kh) ->
// [4 slabs][256 n][4] with k = 32 kc + 16 kh + 4 slab + e and output column 256 pass + n; TF32-rounded.
long long mmla_xproj_arranged_floats() { return static_cast<long long>(kPasses) * kChunksPerPass * kChunkFloats; }
void mmla_xproj_arrange_weights(const float* W, float* out) {
    for (int pass = 0; pass < kPasses; ++pass)
        for (int kc = 0; kc < 4; ++kc)
            for (int kh = 0; kh < 2; ++kh) {
                float* chunk = out + static_cast<long long>((pass * 4 + kc) * 2 + kh) * kChunkFloats;
                for (int slab = 0; slab < 4; ++slab)
                    for (int n = 0; n < 256; ++n)
                        for (int e = 0; e < 4; ++e) {
                            const int k = kc * 32 + kh * 16 + slab * 4 + e;
                            float v = W[static_cast<long long>(k) * kN + 256 * pass + n];
                            uint32_t u;
                            memcpy(&u, &v, 4);
                            if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & ~0x1FFFu;
                            memcpy(&v, &u, 4);
                            chunk[(slab * 256 + n) * 4 + e] = v;
                        }
            }
}

// seq: [B][T][128]; xp_f / xp_b: row-tiled outputs (see the top of this file), mmla_xproj_tiled_floats(B, T) floats each.
long long mmla_xproj_tiled_floats(long long B, int T) { return ((B + 127) / 128) * T * 256LL * 512; }
int mmla_launch_xproj_fused(const float* seq, const float* w_f, const float* w_b, const float* b_f, const float* b_b,
                            float* xp_f, float* xp_b, long long B, int T, cudaStream_t st) {
    MMLA_REQUIRE(B > 0 && T > 0 && B * T < (1LL << 30), MMLA_EINVAL, "xproj_fused: bad geometry");
    static MmlaPerDeviceOnce attr_once;                          // cudaFuncSetAttribute is per device
    const bool attr_set = !attr_once.first();
    const int smem = static_cast<int>(sizeof(XpSmem) + 1024);
    if (!attr_set) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(xproj_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    XpArgs a;
    a.seq = seq;
    a.w[0] = w_f; a.w[1] = w_b;
    a.b[0] = b_f; a.b[1] = b_b;
    a.xp[0] = xp_f; a.xp[1] = xp_b;
    a.B = B; a.T = T;
    const dim3 grid(static_cast<unsigned>(((B + 127) / 128) * T), 2);
    xproj_fused_kernel<<<grid, kThreads, smem, st>>>(a);
    mmla_count_launch("xproj_fused_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}
