// Fused 1-D residual unit for sm_100a (speaker classifier, TF32 tensor-core mode):
//
//   plain :  y = x + Conv1D_k3( ReLU(BN2( Conv1D_k3( ReLU(BN1(x)) ) )) )               (Cin == Cout)
//   pooled:  y = Conv1D_k1_s2(x) + Conv1D_k3( ReLU(BN2( Conv1D_k3( ReLU(BN1( MaxPool1D_2(x) )) ) )) )
//
// — `res_unit(x, filters, pool)` of SpeakerIdentification/scripts/speaker_identification.py:168-190.
// In the pooled unit the max-pool happens while x is loaded, and the stride-2 1x1 shortcut is
// simply extra K-chunks (A = raw even rows of x) accumulated into conv2's TMEM accumulator.
//
// The im2col expansion never exists.  One CTA owns a 128-row tile = G = 128/T whole clips with rows
// interleaved time-major (row = t*G + g), so that a filter tap is a UNIFORM shift of G rows:
//   1. x is read once (coalesced), BN1 + ReLU + TF32 rounding applied once per element, and stored
//      as the tcgen05 A operand in the UMMA K-major no-swizzle "slab" layout
//          [channel quad][row][16 B]      (rows 16 B apart, G zero halo rows at both ends);
//   2. conv1 = 3 taps x C/8 tcgen05.mma (M=128, N=C, K=8): the A descriptor of tap j simply starts
//      G*j rows further down the same buffer — Keras 'same' zero padding is the halo rows;
//   3. epilogue 1 (TMEM -> +bias -> BN2 -> ReLU -> TF32) writes the intermediate straight into a
//      second slab buffer; conv2 runs from it the same way;
//   4. epilogue 2 stages the accumulator through shared memory and writes y = acc + bias + x with
//      coalesced 128-bit accesses (x re-read from L2).
// Weights stream through a 4-stage TMA ring (dedicated producer warp), MMAs are issued by a
// dedicated warp, warps 0..7 load / transform / run the epilogues.  HBM traffic per unit: x in,
// y out — versus five activation tensors and two 3x im2col gathers for the unfused pair.
#include <string.h>

#include "conv_common.cuh"

namespace {

constexpr int kRtot = 137;                 // rows per slab: 128 + 2*G halo (G <= 4), == 1 mod 8 (bank-friendly)
constexpr int kStagesRU = 4;
constexpr int kEpi = 256;
constexpr int kThreadsRU = kEpi + 64;

struct ResUnitArgs {
    const float* x;
    float* y;
    const float* bn1_scale; const float* bn1_shift;
    const float* bn2_scale; const float* bn2_shift;
    const float* w1; const float* w2;      // conv_tc-arranged K-chunk streams ([K/32][8 slabs][C][4])
    const float* ws;                       // shortcut weights (pooled unit), same arrangement
    const float* b1; const float* b2; const float* bs;
    int B, T, G;                           // T = OUTPUT time steps (input has 2T when pooled)
};

template <int CIN, int C, bool POOL>
struct RuSmem {
    // ONE operand buffer, used three times: ReLU(BN1(x)) for conv1; then (conv1's MMAs have retired
    // before epilogue 1 runs) ReLU(BN2(conv1)) for conv2; then the output staging tile.
    alignas(128) unsigned char ab[(C / 4) * kRtot * 16];
    alignas(128) unsigned char a0[POOL ? (CIN / 4) * 128 * 16 : 16];   // raw x[2t] for the shortcut GEMM
    alignas(128) unsigned char ring[kStagesRU][(C == 128 ? 4 : 8) * C * 16];   // weight chunks: 16|32 K x C
    alignas(16) float prm[4][C];                               // b1, BN2 scale, BN2 shift, b2 (+ bs): fetched once at
                                                               // kernel start so the epilogues never wait on L2
    alignas(8) uint64_t full[kStagesRU];
    alignas(8) uint64_t empty[kStagesRU];
    alignas(8) uint64_t a_ready[2];                            // operand buffer written   (epilogue -> MMA)
    alignas(8) uint64_t tfull[2];                              // accumulator complete     (MMA -> epilogue)
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t ru_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void ru_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    asm volatile("trap;");
}
__device__ __forceinline__ void ru_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ru_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ uint64_t ru_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool ru_elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void ru_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int CIN, int C, bool POOL>
__global__ void __launch_bounds__(kThreadsRU, (C == 32 ? 4 : (C == 64 ? 3 : 2))) resunit_fused_kernel(const ResUnitArgs a) {
    static_assert(POOL || CIN == C, "plain units keep the channel count");
    constexpr int kQuads = C / 4;
    constexpr int kQuadsIn = CIN / 4;
    constexpr int kChunkK = C == 128 ? 16 : 32;                  // K per ring chunk (8 KB chunks keep 2 CTAs/SM at C=128)
    constexpr int kChunks1 = 3 * CIN / kChunkK;                  // conv1: K = 3*CIN
    constexpr int kChunksS = POOL ? CIN / kChunkK : 0;           // shortcut: K = CIN
    constexpr int kChunks2 = 3 * C / kChunkK;                    // conv2: K = 3*C
    constexpr uint32_t kChunkBytes = (kChunkK / 4) * C * 16;
    constexpr int kCols = 2 * C < 32 ? 32 : 2 * C;               // two accumulators side by side
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(C >> 3) << 17) |
                                (static_cast<uint32_t>(128 >> 4) << 24);
    extern __shared__ unsigned char smem_dyn[];
    // offset applied to the __shared__ array itself so accesses stay LDS/STS (an integer round-trip makes them generic)
    RuSmem<CIN, C, POOL>& s = *reinterpret_cast<RuSmem<CIN, C, POOL>*>(smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = a.G, T = a.T;
    const int clip0 = blockIdx.x * G;

    // ---- warps 0..7 issue their x loads BEFORE the setup barrier (mbarrier init, TMEM allocation), so the HBM latency
    //      overlaps it; lanes run over channel quads (coalesced), all loads are in flight before the first use ----
    constexpr int kRowsPerPass = kEpi / kQuadsIn;             // rows covered by 256 threads at once
    constexpr int kIters = 128 / kRowsPerPass;
    const int qd = tid % kQuadsIn, rsub = (tid % kEpi) / kQuadsIn;
    float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sh = sc;
    float4 v[kIters], v2[POOL ? kIters : 1];
    if (warp < 8) {
        sc = *reinterpret_cast<const float4*>(a.bn1_scale + 4 * qd);
        sh = *reinterpret_cast<const float4*>(a.bn1_shift + 4 * qd);
        const int Tin = POOL ? 2 * T : T;
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
            const int r = rsub + i * kRowsPerPass;
            const int t = r / G, g = r - t * G;                  // row = t*G + g
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (POOL) v2[i] = v[i];
            if (clip0 + g < a.B) {
                const float* src = a.x + (static_cast<long long>(clip0 + g) * Tin + (POOL ? 2 * t : t)) * CIN + 4 * qd;
                v[i] = *reinterpret_cast<const float4*>(src);
                if (POOL) v2[i] = *reinterpret_cast<const float4*>(src + CIN);
            }
        }
    }

    if (tid == 0) {
        for (int i = 0; i < kStagesRU; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.a_ready[i], 1);
            mbar_init(&s.tfull[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)),
                     "r"(static_cast<uint32_t>(kCols))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;

    if (warp == 8) {
        // ================= TMA producer: conv1 then conv2 weight chunks =================
        if (lane == 0) {
            for (int g = 0; g < kChunks1 + kChunksS + kChunks2; ++g) {
                const int stg = g % kStagesRU, use = g / kStagesRU;
                if (use > 0) ru_wait(&s.empty[stg], static_cast<uint32_t>((use - 1) & 1));
                // the conv_tc arrangement is a plain sequence of K-slabs ([K/4][C][4]), so any
                // multiple-of-4 K granularity is a contiguous slice of it.  Order: conv1, shortcut, conv2.
                const float* src = g < kChunks1              ? a.w1 + static_cast<long long>(g) * (kChunkK * C)
                                   : g < kChunks1 + kChunksS ? a.ws + static_cast<long long>(g - kChunks1) * (kChunkK * C)
                                                             : a.w2 + static_cast<long long>(g - kChunks1 - kChunksS) * (kChunkK * C);
                mbar_arrive_expect_tx(&s.full[stg], kChunkBytes);
                tma_bulk_g2s(&s.ring[stg][0], src, kChunkBytes, &s.full[stg]);
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ================= MMA issuer =================
        // The whole warp runs this loop with warp-uniform values and ONE elected lane issues: descriptors then live in
        // uniform registers and UTCHMMA issues back to back.  (Issued from an `if (lane == 0)` branch with descriptors
        // rebuilt per instruction, each MMA cost 80-200 cycles of R2UR waterfall instead of its ~45-64 cycle floor —
        // measured with scripts/microbench/umma_rate.cu.)
        {
            int g = 0;
            const uint64_t dA_main = ru_desc(smem_u32(&s.ab[0]), kRtot * 16, 128);
            const uint64_t dA_short = ru_desc(smem_u32(&s.a0[0]), 128 * 16, 128);
            const uint64_t dB0 = ru_desc(smem_u32(&s.ring[0][0]), C * 16, 128);
            constexpr uint32_t kStageUnits = sizeof(s.ring[0]) / 16;          // descriptor address units per ring stage
            // one K-chunk of MMAs: A slabs start at (slab0 + 2kk) in a buffer with `rows` rows per slab,
            // shifted down by `shift` rows; accumulates into TMEM column block `dcol`
            auto chunk_mma = [&](uint64_t dA, int rows, int slab0, int shift, uint32_t dcol, bool first) {
                const int stg = g % kStagesRU;
                ru_wait(&s.full[stg], static_cast<uint32_t>((g / kStagesRU) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bd0 = dB0 + static_cast<uint64_t>(stg * kStageUnits);
                const uint64_t ad0 = dA + static_cast<uint64_t>(slab0 * rows + shift);
                if (ru_elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < kChunkK / 8; ++kk) {
                        const uint64_t ad = ad0 + static_cast<uint64_t>(2 * kk * rows);
                        const uint64_t bd = bd0 + static_cast<uint64_t>(kk * 2 * C);
                        const uint32_t acc = (!first || kk != 0) ? 1u : 0u;
                        asm volatile(
                            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + dcol),
                            "l"(ad), "l"(bd), "r"(kIdesc), "r"(acc)
                            : "memory");
                    }
                    ru_commit(&s.empty[stg]);
                }
                __syncwarp();
                ++g;
            };
            // conv1: K index = tap*CIN + channel; tap j reads rows shifted by j*G (halo rows = zero padding)
            ru_wait(&s.a_ready[0], 0u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int ch = 0; ch < kChunks1; ++ch) {
                const int k0 = ch * kChunkK, tap = k0 / CIN, c0 = k0 - tap * CIN;
                chunk_mma(dA_main, kRtot, c0 >> 2, tap * G, 0u, ch == 0);
            }
            if (ru_elect_one()) ru_commit(&s.tfull[0]);
            __syncwarp();
            // shortcut (pooled unit): raw x[2t] x Ws starts conv2's accumulator while epilogue 1 runs
            for (int ch = 0; ch < kChunksS; ++ch) chunk_mma(dA_short, 128, (ch * kChunkK) >> 2, 0, C, ch == 0);
            // conv2
            ru_wait(&s.a_ready[1], 0u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int ch = 0; ch < kChunks2; ++ch) {
                const int k0 = ch * kChunkK, tap = k0 / C, c0 = k0 - tap * C;
                chunk_mma(dA_main, kRtot, c0 >> 2, tap * G, C, ch == 0 && !POOL);
            }
            if (ru_elect_one()) ru_commit(&s.tfull[1]);
        }
        __syncwarp();
    } else {
        // ================= warps 0..7: load/transform, epilogues =================
        for (int i = tid; i < C; i += kEpi) {
            s.prm[0][i] = a.b1[i];
            s.prm[1][i] = a.bn2_scale[i];
            s.prm[2][i] = a.bn2_shift[i];
            s.prm[3][i] = a.b2[i] + (POOL ? a.bs[i] : 0.f);
        }
        // zero the halo rows (rows [0,G) and [G+128, G+128+G)); epilogue 1 never touches them
        for (int i = tid; i < kQuads * 2 * G; i += kEpi) {
            const int qd = i / (2 * G), h = i % (2 * G);
            const int row = h < G ? h : 128 + h;
            *reinterpret_cast<uint4*>(&s.ab[0] + (qd * kRtot + row) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        // ---- [MaxPool2] -> ReLU(BN1(.)) -> TF32 -> operand buffer (x was loaded before the setup barrier) ----
        {
#pragma unroll
            for (int i = 0; i < kIters; ++i) {
                const int r = rsub + i * kRowsPerPass;
                const int g = r % G;
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);       // clips past the batch end stay zero
                if (clip0 + g < a.B) {
                    float4 p = v[i];
                    if (POOL) {
                        // shortcut operand: the raw even row; conv path: max over the pair
                        *reinterpret_cast<uint4*>(&s.a0[0] + (qd * 128 + r) * 16) =
                            make_uint4(ru_tf32(p.x), ru_tf32(p.y), ru_tf32(p.z), ru_tf32(p.w));
                        p.x = fmaxf(p.x, v2[i].x); p.y = fmaxf(p.y, v2[i].y);
                        p.z = fmaxf(p.z, v2[i].z); p.w = fmaxf(p.w, v2[i].w);
                    }
                    w.x = fmaxf(fmaf(p.x, sc.x, sh.x), 0.f);
                    w.y = fmaxf(fmaf(p.y, sc.y, sh.y), 0.f);
                    w.z = fmaxf(fmaf(p.z, sc.z, sh.z), 0.f);
                    w.w = fmaxf(fmaf(p.w, sc.w, sh.w), 0.f);
                } else if (POOL) {
                    *reinterpret_cast<uint4*>(&s.a0[0] + (qd * 128 + r) * 16) = make_uint4(0u, 0u, 0u, 0u);
                }
                *reinterpret_cast<uint4*>(&s.ab[0] + (qd * kRtot + G + r) * 16) =
                    make_uint4(ru_tf32(w.x), ru_tf32(w.y), ru_tf32(w.z), ru_tf32(w.w));
            }
        }
        fence_proxy_async_smem();
        ru_epi_sync();
        if (tid == 0) ru_arrive(&s.a_ready[0]);

        // TMEM lane = tile row; warps 0..3 take the lower half of the columns, warps 4..7 the upper
        const int row = 32 * (warp & 3) + lane;
        constexpr int kColsPerWarp = C / 2;
        const int cbase = (warp >> 2) * kColsPerWarp;
        auto ld16 = [&](uint32_t taddr, float (&z)[16]) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                "%15}, [%16];\n"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] = __uint_as_float(r[j]);
        };

        // ---- epilogue 1: conv1 accumulator -> +b1 -> BN2 -> ReLU -> TF32 -> operand buffer -------
        ru_wait(&s.tfull[0], 0u);                                 // conv1's MMAs (readers of the buffer) have retired
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < kColsPerWarp; c0 += 16) {
            const int col = cbase + c0;
            float z[16];
            ld16(tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + static_cast<uint32_t>(col), z);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(&s.prm[0][col + j]);
                const float4 sc = *reinterpret_cast<const float4*>(&s.prm[1][col + j]);
                const float4 sh = *reinterpret_cast<const float4*>(&s.prm[2][col + j]);
                const float v0 = fmaxf(fmaf(z[j + 0] + bv.x, sc.x, sh.x), 0.f);
                const float v1 = fmaxf(fmaf(z[j + 1] + bv.y, sc.y, sh.y), 0.f);
                const float v2 = fmaxf(fmaf(z[j + 2] + bv.z, sc.z, sh.z), 0.f);
                const float v3 = fmaxf(fmaf(z[j + 3] + bv.w, sc.w, sh.w), 0.f);
                *reinterpret_cast<uint4*>(&s.ab[0] + (((col + j) >> 2) * kRtot + G + row) * 16) =
                    make_uint4(ru_tf32(v0), ru_tf32(v1), ru_tf32(v2), ru_tf32(v3));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        fence_proxy_async_smem();
        ru_epi_sync();
        if (tid == 0) ru_arrive(&s.a_ready[1]);

        // ---- epilogue 2: conv2 accumulator + b2 -> staging (operand buffers are dead) -> y = . + x ----
        constexpr int kStride = C + 4;                            // floats per staged row (conflict-free STS.128)
        static_assert(kStride * 128 * 4 <= sizeof(s.ab) + sizeof(s.a0) + sizeof(s.ring),
                      "staging tile must fit in the (now idle) operand buffers + weight ring");
        float* stg = reinterpret_cast<float*>(&s.ab[0]);          // may run over into a0 / the ring head (all idle now)
        ru_wait(&s.tfull[1], 0u);                                 // every MMA has retired, every weight chunk consumed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < kColsPerWarp; c0 += 16) {
            const int col = cbase + c0;
            float z[16];
            ld16(tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + static_cast<uint32_t>(C + col), z);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(&s.prm[3][col + j]);
                *reinterpret_cast<float4*>(stg + row * kStride + col + j) =
                    make_float4(z[j] + bv.x, z[j + 1] + bv.y, z[j + 2] + bv.z, z[j + 3] + bv.w);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        ru_epi_sync();
        {
            constexpr int kIters = 128 * kQuads / kEpi;           // float4 per thread
            float4 xr[kIters];
#pragma unroll
            for (int i = 0; i < kIters; ++i) {                    // all residual loads first
                const int idx = tid + i * kEpi;
                const int r = idx / kQuads, qd = idx - r * kQuads;
                const int t = r / G, g = r - t * G;
                xr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!POOL && clip0 + g < a.B)                    // pooled units got their residual from the shortcut GEMM
                    xr[i] = *reinterpret_cast<const float4*>(a.x + (static_cast<long long>(clip0 + g) * T + t) * C + 4 * qd);
            }
#pragma unroll
            for (int i = 0; i < kIters; ++i) {
                const int idx = tid + i * kEpi;
                const int r = idx / kQuads, qd = idx - r * kQuads;
                const int t = r / G, g = r - t * G;
                if (clip0 + g < a.B) {
                    float4 v = *reinterpret_cast<const float4*>(stg + r * kStride + 4 * qd);
                    v.x += xr[i].x; v.y += xr[i].y; v.z += xr[i].z; v.w += xr[i].w;
                    *reinterpret_cast<float4*>(a.y + (static_cast<long long>(clip0 + g) * T + t) * C + 4 * qd) = v;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(static_cast<uint32_t>(kCols))
                     : "memory");
}

template <int CIN, int C, bool POOL>
int launch_ru(const ResUnitArgs& a, cudaStream_t st) {
    static MmlaPerDeviceOnce attr_once;                          // cudaFuncSetAttribute is per device
    const bool attr_set = !attr_once.first();
    const int smem = static_cast<int>(sizeof(RuSmem<CIN, C, POOL>) + 128);
    if (!attr_set) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(resunit_fused_kernel<CIN, C, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    const unsigned grid = static_cast<unsigned>((a.B + a.G - 1) / a.G);
    resunit_fused_kernel<CIN, C, POOL><<<grid, kThreadsRU, smem, st>>>(a);
    mmla_count_launch("resunit_fused_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}

}  // namespace

// x: [B][Tin][Cin], y: [B][T][C] fp32 NHWC (H = 1); T = output steps in {128, 64, 32}, Tin = T (plain,
// Cin == C) or 2T (pooled).  Weights are the conv_tc-arranged streams of the k=3 convolutions (and of
// the 1x1 stride-2 shortcut when pooled: ws/bs non-null).
int mmla_launch_resunit_fused(const float* x, float* y, long long B, int T, int Cin, int C, const float* bn1_scale,
                              const float* bn1_shift, const float* w1, const float* b1, const float* bn2_scale,
                              const float* bn2_shift, const float* w2, const float* b2, const float* ws, const float* bs,
                              cudaStream_t st) {
    MMLA_REQUIRE(T >= 32 && T <= 128 && 128 % T == 0, MMLA_EUNSUP, "resunit_fused: T=%d unsupported", T);
    MMLA_REQUIRE(B > 0 && B < (1LL << 24), MMLA_EINVAL, "resunit_fused: bad batch");
    ResUnitArgs a;
    a.x = x; a.y = y;
    a.bn1_scale = bn1_scale; a.bn1_shift = bn1_shift; a.bn2_scale = bn2_scale; a.bn2_shift = bn2_shift;
    a.w1 = w1; a.w2 = w2; a.ws = ws; a.b1 = b1; a.b2 = b2; a.bs = bs;
    a.B = static_cast<int>(B); a.T = T; a.G = 128 / T;
    const bool pool = ws != nullptr;
    if (!pool && Cin == C) {
        if (C == 32) return launch_ru<32, 32, false>(a, st);
        if (C == 64) return launch_ru<64, 64, false>(a, st);
        if (C == 128) return launch_ru<128, 128, false>(a, st);
    } else if (pool) {
        if (Cin == 32 && C == 32) return launch_ru<32, 32, true>(a, st);
        if (Cin == 32 && C == 64) return launch_ru<32, 64, true>(a, st);
        if (Cin == 64 && C == 128) return launch_ru<64, 128, true>(a, st);
    }
    mmla_set_error("resunit_fused: Cin=%d C=%d pool=%d unsupported", Cin, C, pool ? 1 : 0);
    return MMLA_EUNSUP;
}
