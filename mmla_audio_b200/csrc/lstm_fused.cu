// Persistent fused BiLSTM(256) recurrence for sm_100a (TF32 tensor-core mode).
//
// Keras `Bidirectional(LSTM(256))` as used by both reference classifiers
// (OverlapDetection/scripts/overlap_detector_temp.py:297,
//  SpeakerIdentification/scripts/speaker_identification.py:213): gates i,f,c,o, sigmoid
// recurrent activation, the backward layer walks t = T-1..0, only the last state is kept.
//
// One CTA owns 128 clips of one direction for ALL time steps:
//   * h_{t-1} [128 x 256] lives in shared memory as the tcgen05 A operand (eight 128x32 TF32
//     sub-tiles, UMMA K-major SWIZZLE_128B);
//   * the recurrent weights U [256 x 1024] are streamed every step from L2 by the TMA bulk-copy
//     engine through an 8-stage ring of 8 KB chunks, one per MMA (host pre-arranged: gate columns regrouped so a
//     512-column accumulator pass holds i|f|c~|o for the same 128 units, TF32 pre-rounded);
//   * z = h U accumulates in TMEM (512 fp32 columns = the whole tensor memory of the SM), two
//     passes of 128 units per step, tcgen05.mma M=128 N=256 K=8;
//   * the epilogue reads z from TMEM, adds the pre-computed input projection xp[:, t, :], applies
//     the gates, updates c (global scratch) and h (global, fp32) and the new h is re-staged into
//     the shared-memory operand for the next step.
// This replaces T x (GEMM launch + gate launch) per direction by ONE launch for both directions
// and removes the [B,1024] pre-activation round trip through HBM.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kU = 256;                     // LSTM units
constexpr int kRows = 128;                  // clips per CTA
constexpr int kSubTile = 128 * 128;         // bytes of one 128x32 TF32 sub-tile
constexpr int kBChunkFloats = 2 * 256 * 4;  // one weight chunk = one MMA: 2 K-slabs (K=8) x 256 columns x 4
constexpr int kBChunkBytes = kBChunkFloats * 4;   // 8 KB
constexpr int kChunksPerStep = 128;         // 2 halves x 2 N-tiles x 8 K-sub-tiles x 4 MMAs
constexpr int kStages = 8;                  // deep ring of small chunks: 7 TMA copies in flight hide L2 latency

struct LstmSmem {
    alignas(1024) unsigned char H[8][kSubTile];          // h_{t-1}, SWIZZLE_128B K-major
    alignas(128) unsigned char Bst[kStages][kBChunkBytes];   // weight ring (no-swizzle slab layout)
    alignas(8) uint64_t full[kStages];
    alignas(8) uint64_t empty[kStages];
    alignas(8) uint64_t accum;
    uint32_t tmem_base;
};

struct LstmArgs {
    const float* xp[2];     // [B][T][1024] input projections (+bias), Keras column order i|f|c|o
    const float* wr[2];     // arranged recurrent weights, kChunksPerStep x kBChunkFloats
    float* h[2];            // [B][256] running / final hidden state
    float* c[2];            // [B][256] cell state
    int B, T;
};

__device__ __forceinline__ uint32_t tf32_round(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 2.f * fast_sigmoid(2.f * x) - 1.f; }

__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    asm volatile("trap;");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (1ull << 16) | (static_cast<uint64_t>(1024u >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(256, 1) lstm_fused_kernel(const LstmArgs a) {
    extern __shared__ unsigned char smem_dyn[];
    LstmSmem& s = *reinterpret_cast<LstmSmem*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const int b0 = blockIdx.x * kRows;
    const float* xp = a.xp[dir];
    const float* wr = a.wr[dir];
    float* hg = a.h[dir];
    float* cg = a.c[dir];
    const int T = a.T;
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(256 >> 3) << 17) |
                                (static_cast<uint32_t>(128 >> 4) << 24);   // f32 += tf32 x tf32, M=128, N=256

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], 1);
        }
        mbar_init(&s.accum, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;

    // epilogue mapping: TMEM lane = row; warps 0..3 take units [0,64) of a half, warps 4..7 [64,128)
    const int row = 32 * (warp & 3) + lane;
    const int brow = b0 + row;
    const bool row_ok = brow < a.B;
    const int ubase = 64 * (warp >> 2);
    uint32_t accum_phase = 0;
    long long produced = 0, consumed = 0;            // weight chunks issued / used (thread 0 only)
    const long long total_chunks = static_cast<long long>(T > 1 ? T - 1 : 0) * kChunksPerStep;

    // One 16-unit slice of the cell update for this thread's row.  All global loads (four gate
    // slices of xp and the old cell state: 20 x 128-bit per thread) are issued before the first
    // TMEM read, so their latency overlaps; gates are then consumed one TMEM slice at a time.
    auto ld16 = [&](uint32_t taddr, float (&z)[16]) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
            "[%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = __uint_as_float(r[j]);
    };
    auto cell16 = [&](int t, int half, int uoff, bool first) {      // uoff: unit offset inside the half
        const int u = 128 * half + uoff;
        const long long rrow = row_ok ? brow : 0;                     // masked rows read row 0, never write
        const float* xr = xp + (rrow * T + t) * 1024 + u;
        float* cp = cg + rrow * kU + u;
        float* hp = hg + rrow * kU + u;
        float xi[16], xf[16], xc[16], xo[16], cv[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            *reinterpret_cast<float4*>(&xi[j]) = __ldcg(reinterpret_cast<const float4*>(xr + j));
            *reinterpret_cast<float4*>(&xc[j]) = __ldcg(reinterpret_cast<const float4*>(xr + 512 + j));
            *reinterpret_cast<float4*>(&xf[j]) = __ldcg(reinterpret_cast<const float4*>(xr + 256 + j));
            *reinterpret_cast<float4*>(&xo[j]) = __ldcg(reinterpret_cast<const float4*>(xr + 768 + j));
            *reinterpret_cast<float4*>(&cv[j]) =
                first ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcg(reinterpret_cast<const float4*>(cp + j));
        }
        const uint32_t tbase = tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + static_cast<uint32_t>(uoff);
        float z[16], pr[16];
        if (!first) ld16(tbase, z);                                   // i
#pragma unroll
        for (int j = 0; j < 16; ++j) pr[j] = fast_sigmoid((first ? 0.f : z[j]) + xi[j]);
        if (!first) ld16(tbase + 256, z);                             // c~
#pragma unroll
        for (int j = 0; j < 16; ++j) pr[j] *= fast_tanh((first ? 0.f : z[j]) + xc[j]);
        if (!first) ld16(tbase + 128, z);                             // f
#pragma unroll
        for (int j = 0; j < 16; ++j) cv[j] = fast_sigmoid((first ? 0.f : z[j]) + xf[j]) * cv[j] + pr[j];
        if (!first) ld16(tbase + 384, z);                             // o
#pragma unroll
        for (int j = 0; j < 16; ++j) pr[j] = fast_sigmoid((first ? 0.f : z[j]) + xo[j]) * fast_tanh(cv[j]);
        if (row_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                *reinterpret_cast<float4*>(cp + j) = *reinterpret_cast<float4*>(&cv[j]);
                *reinterpret_cast<float4*>(hp + j) = *reinterpret_cast<float4*>(&pr[j]);
            }
        }
    };

    // re-stage h (global, fp32) into the SWIZZLE_128B TF32 operand tiles
    auto restage_h = [&]() {
        const int q = tid & 7, rb = tid >> 3;
#pragma unroll 2
        for (int kc = 0; kc < 8; ++kc)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = rb + 32 * i;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (b0 + r < a.B) v = __ldcg(reinterpret_cast<const float4*>(hg + static_cast<long long>(b0 + r) * kU + 32 * kc + 4 * q));
                *reinterpret_cast<uint4*>(&s.H[kc][0] + r * 128 + ((q ^ (r & 7)) << 4)) =
                    make_uint4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
            }
        fence_proxy_async_smem();
    };

    auto issue_weight_chunk = [&]() {                 // thread 0: next chunk of the (periodic) weight stream
        const int stg = static_cast<int>(produced % kStages);
        const long long use = produced / kStages;
        if (use > 0) wait_or_trap(&s.empty[stg], static_cast<uint32_t>((use - 1) & 1));
        // (no proxy fence: the ring is written by TMA and read by the tensor core — async proxy only)
        mbar_arrive_expect_tx(&s.full[stg], kBChunkBytes);
        tma_bulk_g2s(&s.Bst[stg][0], wr + (produced % kChunksPerStep) * kBChunkFloats, kBChunkBytes, &s.full[stg]);
        ++produced;
    };

    if (tid == 0)                                                 // weights start flowing during step 0
        while (produced < total_chunks && produced < kStages - 2) issue_weight_chunk();
    if (warp == 0) __syncwarp();

    for (int step = 0; step < T; ++step) {
        const int t = dir == 0 ? step : T - 1 - step;
        if (step == 0) {
            // h0 = c0 = 0: pre-activations are the input projection alone
            for (int half = 0; half < 2; ++half)
                for (int uo = 0; uo < 64; uo += 16) cell16(t, half, ubase + uo, true);
        } else {
            for (int half = 0; half < 2; ++half) {
                // Warp 0 stays converged around its issuing lane: if lanes 1..31 ran ahead into the
                // blocking mbarrier wait below, the suspended warp would stall lane 0's issue loop.
                if (tid == 0) {
                    // 64 MMAs: N-tile j (0: i|f, 1: c~|o) x K sub-tile kc x 4 K-steps, one 8 KB chunk each
                    for (int j = 0; j < 2; ++j)
                        for (int kc = 0; kc < 8; ++kc)
                            for (int kk = 0; kk < 4; ++kk) {
                                // refill lags two slots behind the ring size so the slot being re-armed
                                // belongs to an MMA committed two iterations ago (never blocks on a fresh one)
                                while (produced < total_chunks && produced - consumed < kStages - 2) issue_weight_chunk();
                                const int stg = static_cast<int>(consumed % kStages);
                                wait_or_trap(&s.full[stg], static_cast<uint32_t>((consumed / kStages) & 1));
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                                const uint64_t ad = desc_sw128(smem_u32(&s.H[kc][0]) + kk * 32);
                                const uint64_t bd = desc_noswz(smem_u32(&s.Bst[stg][0]), 256 * 16, 128);
                                const uint32_t acc = (kc | kk) != 0 ? 1u : 0u;
                                asm volatile(
                                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + j * 256),
                                    "l"(ad), "l"(bd), "r"(kIdesc), "r"(acc)
                                    : "memory");
                                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                                 smem_u32(&s.empty[stg]))
                                             : "memory");
                                ++consumed;
                            }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_u32(&s.accum))
                                 : "memory");
                }
                if (warp == 0) __syncwarp();
                wait_or_trap(&s.accum, accum_phase);
                accum_phase ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // epilogue of this half: TMEM columns [0,128) i, [128,256) f, [256,384) c~, [384,512) o
                for (int uo = 0; uo < 64; uo += 16) cell16(t, half, ubase + uo, false);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();                       // TMEM drained before the next half's MMAs overwrite it
            }
        }
        if (step + 1 < T) {
            __syncthreads();                           // every thread's h writes are done (all MMAs reading H too)
            restage_h();
            __syncthreads();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace

// Host: arrange U [256][1024] (Keras recurrent kernel, columns i|f|c|o) into the chunk stream the
// kernel consumes: chunk (half, j, kc) -> [8 slabs][256 n][4], n<128: gate 2j, else gate 2j+1,
// unit = 128*half + (n&127); TF32-rounded.  out: 128 chunks x 2048 floats (one MMA each).
long long mmla_lstm_arranged_floats() { return static_cast<long long>(kChunksPerStep) * kBChunkFloats; }
void mmla_lstm_arrange_weights(const float* U, float* out) {
    for (int half = 0; half < 2; ++half)
        for (int j = 0; j < 2; ++j)
            for (int kc = 0; kc < 8; ++kc)
                for (int kk = 0; kk < 4; ++kk) {
                    float* chunk = out + (static_cast<long long>(((half * 2 + j) * 8 + kc) * 4 + kk)) * kBChunkFloats;
                    for (int slab = 0; slab < 2; ++slab)
                        for (int n = 0; n < 256; ++n)
                            for (int e = 0; e < 4; ++e) {
                                const int k = kc * 32 + kk * 8 + slab * 4 + e;
                                const int gate = 2 * j + (n >= 128 ? 1 : 0);
                                const int unit = 128 * half + (n & 127);
                                float v = U[static_cast<long long>(k) * 1024 + gate * 256 + unit];
                                uint32_t u;
                                memcpy(&u, &v, 4);
                                if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & ~0x1FFFu;
                                memcpy(&v, &u, 4);
                                chunk[(slab * 256 + n) * 4 + e] = v;
                            }
                }
}

int mmla_launch_lstm_fused(const float* xp_f, const float* xp_b, const float* wr_f, const float* wr_b, float* h_f,
                           float* h_b, float* c_f, float* c_b, long long B, int T, cudaStream_t st) {
    MMLA_REQUIRE(B > 0 && B < (1LL << 22) && T >= 1, MMLA_EINVAL, "lstm_fused: bad batch/T");
    static bool attr_set = false;
    const int smem = static_cast<int>(sizeof(LstmSmem) + 1024);
    if (!attr_set) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(lstm_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    LstmArgs a;
    a.xp[0] = xp_f; a.xp[1] = xp_b; a.wr[0] = wr_f; a.wr[1] = wr_b;
    a.h[0] = h_f; a.h[1] = h_b; a.c[0] = c_f; a.c[1] = c_b;
    a.B = static_cast<int>(B); a.T = T;
    const dim3 grid(static_cast<unsigned>((B + kRows - 1) / kRows), 2);
    lstm_fused_kernel<<<grid, 256, smem, st>>>(a);
    mmla_count_launch();
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}
