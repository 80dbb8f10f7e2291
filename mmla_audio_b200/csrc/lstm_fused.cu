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
//   * the recurrent weights U [256 x 1024] are streamed every step from L2 by a dedicated TMA
//     producer warp through a 4-stage ring of 16 KB chunks (host pre-arranged: gate columns
//     regrouped so a 256-column accumulator pass holds i|f|c~|o for the same 64 units, TF32
//     pre-rounded);
//   * z = h U accumulates in TMEM: four passes of 64 units per step, tcgen05.mma M=128 N=256 K=8
//     issued by a dedicated MMA warp; the two 256-column TMEM halves ping-pong so the MMAs of pass
//     p+1 overlap the gate epilogue (warps 0..7) of pass p;
//   * the epilogue reads z from TMEM, adds the pre-computed input projection xp[:, t, :], applies
//     the gates, updates c (global scratch) and h (global, fp32) and the new h is re-staged into
//     the shared-memory operand for the next step.
// This replaces T x (GEMM launch + gate launch) per direction by ONE launch for both directions
// and removes the [B,1024] pre-activation round trip through HBM.
//
// A 128-clip tile is shared by a CLUSTER OF TWO CTAs: CTA `half` computes the gate columns of units
// [128 half, 128 half + 128) (two of the four 64-unit passes) for all 128 clips, so 4096 clips give
// 128 CTAs instead of 64 (148 SMs) and each CTA streams half of U per step.  h(t) goes through global
// memory as before; the two halves hand it over with one remote mbarrier arrive per step (DSMEM).
//
// Global layouts.  The gate epilogue holds one clip per thread (TMEM lane = row), so everything it touches every step
// is "row-tiled": [clip tile][column quad][row 0..127][4 floats] — a warp's 128-bit access is then 512 contiguous
// bytes.  That is the layout of xp (written that way by xproj_fused.cu: [tile][t][256 quads][128][4]), of the cell
// state c ([tile][64 quads][128][4]) and of the h exchange buffer hx (same shape, two copies used alternately so a
// CTA never overwrites the h its peer may still be re-staging).  In the natural [b][..] layouts every lane touched
// its own 128-byte line: 17k of the 60k cycles of a step went into those loads (scripts/prof_lstm.py).  Only the
// final h is written in the caller's [B][256] layout.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kU = 256;                     // LSTM units
constexpr int kRows = 128;                  // clips per CTA
constexpr int kSubTile = 128 * 128;         // bytes of one 128x32 TF32 sub-tile
constexpr int kBChunkFloats = 4 * 256 * 4;  // one weight chunk = two MMAs: 4 K-slabs (K=16) x 256 columns x 4
constexpr int kBChunkBytes = kBChunkFloats * 4;   // 16 KB
constexpr int kChunksPerStep = 64;          // 4 passes x 8 K-sub-tiles x 2 halves (fp16 operands: 4 x 4 x 2 = 32 chunks of K = 32)
constexpr int kStages = 6;                   // 96 KB in flight: the stream into a CTA is ring bytes / ~2 000 cycles of L2 latency
constexpr int kEpiThreads = 256;            // warps 0..7: gate epilogue + h re-staging
constexpr int kProducers = 2;               // TMA producer warps (see the producer loop)
constexpr int kThreads = kEpiThreads + 32 + 32 * kProducers;  // warp 8: MMA issuer, warps 9..: TMA producers

struct LstmSmem {
    alignas(1024) unsigned char H[8][kSubTile];              // h_{t-1}, SWIZZLE_128B K-major
    alignas(128) unsigned char Bst[kStages][kBChunkBytes];   // weight ring (no-swizzle slab layout)
    alignas(8) uint64_t full[kStages];
    alignas(8) uint64_t empty[kStages];
    alignas(8) uint64_t tfull[2];                            // accumulator pass ready   (MMA -> epilogue)
    alignas(8) uint64_t tempty[2];                           // accumulator pass drained (epilogue -> MMA)
    alignas(8) uint64_t hready;                              // h re-staged for the next step
    alignas(8) uint64_t peer[2];                             // the other half's h(t) is in global memory (even / odd steps)
    uint32_t tmem_base;
};

struct LstmArgs {
    const float* xp[2];     // row-tiled input projections (+bias), columns in Keras order i|f|c|o
    const float* wr[2];     // arranged recurrent weights, kChunksPerStep x kBChunkFloats
    float* h[2];            // [B][256] final hidden state
    float* c[2];            // row-tiled cell state, ceil(B/128) x 64 x 128 x 4 floats
    float* hx[2];           // row-tiled h exchange, 2 x the size of c
    int B, T;
    int f16;                // fp16 operands (h in (-1, 1) and the recurrent weights as halves, fp32 accumulation): wr = fp16 chunk stream
    long long* stamps;      // diagnostics: clock64 timeline of CTA (stamp_cta, dir 0), 16 slots per step (null = off)
    int stamp_cta;
};

__device__ __forceinline__ uint32_t tf32_round(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
// Gate non-linearities on the MUFU pipe: tanh.approx.f32 (max relative error ~2^-11, the same
// order as the TF32 operand rounding of this mode) and sigmoid(x) = 0.5 + 0.5 tanh(x/2).
__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }

__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    asm volatile("trap;");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // epilogue warps only
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (1ull << 16) | (static_cast<uint64_t>(1024u >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool lstm_elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit_to(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-specialised: warp 8 streams the weights (TMA), warp 9 issues the MMAs, warps 0..7 run the
// gate epilogue.  A step is four accumulator passes of 64 units ([i|f|c~|o] x 64 = 256 TMEM
// columns); the two TMEM halves ping-pong so the MMAs of pass p+1 overlap the epilogue of pass p.
// F16: h_{t-1} and U as fp16 (`kind::f16`, K = 16 per MMA).  h = o * tanh(c) lies in (-1, 1), so fp16 keeps exactly the 11
// significant bits TF32 keeps and no range question arises; a 16 KB weight chunk then covers K = 32 instead of 16, i.e. the
// stream a step pulls through the ring — what paces the recurrence, see kStages — is half as long.  The operand tiles become
// four 128 x 64-half sub-tiles (the same 128-byte swizzled rows).
// NP: CTAs per 128-clip tile (the cluster size).  2: two 64-unit passes per CTA and step (4096 speaker clips = 128 CTAs); 4: one
// pass each — chosen for small batches (the overlap net's 512 clips: 32 CTAs instead of 16), where a step is paced by the gate
// epilogue's MUFU work of ONE SM and not by anything shared.
template <bool F16, int NP>
__global__ void __launch_bounds__(kThreads, 1) lstm_fused_kernel(const LstmArgs a) {
    extern __shared__ unsigned char smem_dyn[];
    // offset applied to the __shared__ array itself so accesses stay LDS/STS (an integer round-trip makes them generic)
    LstmSmem& s = *reinterpret_cast<LstmSmem*>(smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const int half = blockIdx.x % NP;                         // rank in the cluster = which 256 / NP units
    const int b0 = (blockIdx.x / NP) * kRows;
    constexpr int kPasses = 4 / NP;                           // 64-unit passes per CTA and step
    constexpr int kChunksCta = kChunksPerStep / (F16 ? 2 : 1) / NP;   // weight chunks per CTA and step
    constexpr int kSubTiles = F16 ? 4 : 8;                    // 128-byte-row K sub-tiles of the h operand
    const float* xp = a.xp[dir];
    const float* wr = a.wr[dir];
    float* hg = a.h[dir];
    const long long tile = blockIdx.x / NP;
    const long long tile_floats = 64LL * 512;                  // one row-tiled [64 quads][128 rows][4] block
    float* cg = a.c[dir] + tile * tile_floats;
    float* hx0 = a.hx[dir] + tile * tile_floats;                // buffer of even steps; odd steps: + hx_stride
    const long long hx_stride = ((a.B + kRows - 1) / kRows) * tile_floats;
    const int T = a.T;
    constexpr uint32_t kIdesc = (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | (static_cast<uint32_t>(256 >> 3) << 17) |
                                (static_cast<uint32_t>(128 >> 4) << 24);   // f32 += tf32 x tf32 (or f16 x f16), M=128, N=256

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.tfull[i], 1);
            mbar_init(&s.tempty[i], 8);           // one arrival per epilogue warp
        }
        mbar_init(&s.hready, 1);
        mbar_init(&s.peer[0], NP - 1);           // one arrival per peer CTA
        mbar_init(&s.peer[1], NP - 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // both CTAs' barriers must be initialised before either arrives remotely
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;
    const bool stamping = a.stamps != nullptr && static_cast<int>(blockIdx.x) == a.stamp_cta && dir == 0;
    auto stamp = [&](int step, int slot) {
        if (stamping) a.stamps[step * 16 + slot] = clock64();
    };
    const int n_rec = T > 1 ? T - 1 : 0;                          // recurrent steps
    const long long total_chunks = static_cast<long long>(n_rec) * kChunksCta;

    if (warp > 8) {
        // ================= TMA producers: the periodic weight stream =================
        // Two producer WARPS, alternate chunks: a thread gets one cp.async.bulk through every ~420 cycles whatever its
        // size (scripts/microbench/tma_stream.cu), so ONE producer moves 16 KB chunks at 39 B/clk while the MMAs of a
        // chunk (2 x N=256) need it in 256 cycles = 64 B/clk.  (Several lanes of one warp do not help: their spin
        // waits serialise.)
        if (lane == 0) {
            for (long long g = warp - 9; g < total_chunks; g += kProducers) {
                const int stg = static_cast<int>(g % kStages);
                const long long use = g / kStages;
                if (use > 0) wait_or_trap(&s.empty[stg], static_cast<uint32_t>((use - 1) & 1));
                mbar_arrive_expect_tx(&s.full[stg], kBChunkBytes);
                tma_bulk_g2s(&s.Bst[stg][0], wr + (half * kChunksCta + g % kChunksCta) * kBChunkFloats, kBChunkBytes,
                             &s.full[stg]);
            }
        }
        __syncwarp();
    } else if (warp == 8) {
        // ================= MMA issuer (warp-uniform loop, one elected lane issues; see resunit_fused.cu) =================
        {
            // Ring position as counters, descriptors advanced by adds, two chunks (4 MMAs) per trip: what sits between
            // the last MMA of a trip and the first of the next is a bubble once the tensor pipe's short queue (~4
            // MMAs) has drained — with one chunk (2 MMAs, 256 cycles) per trip a pass took 7.7k cycles instead of 4.1k.
            static_assert(kStages % 2 == 0, "chunk pairs use stages {0,1}, {2,3}, ...");
            int stg = 0;
            uint32_t par = 0;
            const uint64_t dB0 = desc_noswz(smem_u32(&s.Bst[0][0]), 256 * 16, 128);
            const uint64_t dA0 = desc_sw128(smem_u32(&s.H[0][0]));
            constexpr uint32_t kStageUnits = kBChunkBytes / 16;
            int P = 0;                                            // this CTA's pass counter
            for (int rs = 0; rs < n_rec; ++rs) {
                wait_or_trap(&s.hready, static_cast<uint32_t>(rs & 1));        // h_{t-1} staged in smem
                if (lane == 0) stamp(rs + 1, 8);
                for (int pass = 0; pass < kPasses; ++pass, ++P) {
                    const int buf = P & 1;
                    if (P >= 2) wait_or_trap(&s.tempty[buf], static_cast<uint32_t>(((P >> 1) - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t dcol = tmem + static_cast<uint32_t>(buf * 256);
                    uint64_t dA = dA0;
                    for (int kc = 0; kc < kSubTiles; ++kc, dA += static_cast<uint64_t>(kSubTile / 16)) {
                        // K-sub-tile kc: chunk 2kc covers its K columns 0..15, chunk 2kc+1 columns 16..31 (+64 B inside
                        // the swizzled row = 4 descriptor units); 32 B per K=8 MMA
                        const uint64_t bd0 = dB0 + static_cast<uint64_t>(stg * kStageUnits);
                        wait_or_trap(&s.full[stg], par);
                        wait_or_trap(&s.full[stg + 1], par);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (lstm_elect_one()) {
#pragma unroll
                            for (int m = 0; m < 4; ++m) {
                                const uint64_t ad = dA + static_cast<uint64_t>(m * 2);
                                const uint64_t bd = bd0 + static_cast<uint64_t>((m >> 1) * kStageUnits + (m & 1) * 2 * 256);
                                const uint32_t acc = (kc | m) != 0 ? 1u : 0u;
                                if constexpr (F16)
                                    asm volatile(
                                        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(dcol),
                                        "l"(ad), "l"(bd), "r"(kIdesc), "r"(acc)
                                        : "memory");
                                else
                                asm volatile(
                                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(dcol),
                                    "l"(ad), "l"(bd), "r"(kIdesc), "r"(acc)
                                    : "memory");
                            }
                            umma_commit_to(&s.empty[stg]);
                            umma_commit_to(&s.empty[stg + 1]);
                            if (kc == kSubTiles - 1) umma_commit_to(&s.tfull[buf]);
                        }
                        stg += 2;
                        if (stg == kStages) { stg = 0; par ^= 1u; }
                    }
                    if (lane == 0) stamp(rs + 1, 9 + pass);
                }
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue warps 0..7 =================
        // TMEM lane = row; warps 0..3 take units [0,32) of a 64-unit pass, warps 4..7 [32,64)
        const int row = 32 * (warp & 3) + lane;
        const int brow = b0 + row;
        const bool row_ok = brow < a.B;
        const int usub = 32 * (warp >> 2);

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
        // 16 units of the cell update for this thread's row.  All global loads (four gate slices of
        // xp + the old cell state, 20 x 128-bit) are issued before the first TMEM read.
        auto cell16 = [&](int step, int t, int u, uint32_t tcol, bool first) {  // u: absolute unit; tcol: TMEM column of gate i
            const float* xr = xp + ((tile * T + t) * 256 + (u >> 2)) * 512 + row * 4;   // gate g: + 64 g quads
            float* cp = cg + (u >> 2) * 512 + row * 4;
            float* hp = hx0 + (step & 1) * hx_stride + (u >> 2) * 512 + row * 4;
            float xi[16], xf[16], xc[16], xo[16], cv[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                *reinterpret_cast<float4*>(&xi[j]) = __ldcg(reinterpret_cast<const float4*>(xr + (j >> 2) * 512));
                *reinterpret_cast<float4*>(&xc[j]) = __ldcg(reinterpret_cast<const float4*>(xr + (128 + (j >> 2)) * 512));
                *reinterpret_cast<float4*>(&xf[j]) = __ldcg(reinterpret_cast<const float4*>(xr + (64 + (j >> 2)) * 512));
                *reinterpret_cast<float4*>(&xo[j]) = __ldcg(reinterpret_cast<const float4*>(xr + (192 + (j >> 2)) * 512));
                *reinterpret_cast<float4*>(&cv[j]) =
                    first ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcg(reinterpret_cast<const float4*>(cp + (j >> 2) * 512));
            }
            const uint32_t tbase = tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + tcol;
            float z[16], pr[16];
            if (!first) ld16(tbase, z);                                // i
#pragma unroll
            for (int j = 0; j < 16; ++j) pr[j] = fast_sigmoid((first ? 0.f : z[j]) + xi[j]);
            if (!first) ld16(tbase + 128, z);                          // c~
#pragma unroll
            for (int j = 0; j < 16; ++j) pr[j] *= fast_tanh((first ? 0.f : z[j]) + xc[j]);
            if (!first) ld16(tbase + 64, z);                           // f
#pragma unroll
            for (int j = 0; j < 16; ++j) cv[j] = fast_sigmoid((first ? 0.f : z[j]) + xf[j]) * cv[j] + pr[j];
            if (!first) ld16(tbase + 192, z);                          // o
#pragma unroll
            for (int j = 0; j < 16; ++j) pr[j] = fast_sigmoid((first ? 0.f : z[j]) + xo[j]) * fast_tanh(cv[j]);
            if (step + 1 < T) {                                        // padded rows live in the tile too: no masking
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    *reinterpret_cast<float4*>(cp + (j >> 2) * 512) = *reinterpret_cast<float4*>(&cv[j]);
                    *reinterpret_cast<float4*>(hp + (j >> 2) * 512) = *reinterpret_cast<float4*>(&pr[j]);
                }
            } else if (row_ok) {                                       // last step: the caller's [B][256] layout
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(hg + static_cast<long long>(brow) * kU + u + j) = *reinterpret_cast<float4*>(&pr[j]);
            }
        };
        // re-stage h (global, fp32) into the SWIZZLE_128B TF32 operand tiles; loads are issued
        // 16 at a time before their first use
        auto restage_h = [&](int step) {
            // lanes run over rows (coalesced 512-byte reads of the row-tiled exchange buffer); quad uq of the 64 is
            // K-sub-tile uq/8, 16-byte column uq%8 of the swizzled operand row
            const float* src = hx0 + (step & 1) * hx_stride;
            const int r = tid & 127, uq0 = tid >> 7;
            const bool ok = b0 + r < a.B;
#pragma unroll
            for (int i0 = 0; i0 < 32; i0 += 16) {
                float4 v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int uq = uq0 + 2 * (i0 + j);
                    v[j] = __ldcg(reinterpret_cast<const float4*>(src + uq * 512 + r * 4));
                    if (!ok) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int uq = uq0 + 2 * (i0 + j);
                    if constexpr (F16) {
                        // four units = 8 bytes: half of the 16-byte column (uq >> 1) & 7 of K sub-tile uq >> 4 (64 halves per row)
                        const __half2 lo = __floats2half2_rn(v[j].x, v[j].y), hi = __floats2half2_rn(v[j].z, v[j].w);
                        *reinterpret_cast<uint2*>(&s.H[uq >> 4][0] + r * 128 + ((((uq >> 1) & 7) ^ (r & 7)) << 4) + (uq & 1) * 8) =
                            make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
                        continue;
                    }
                    const int kc = uq >> 3, q = uq & 7;
                    *reinterpret_cast<uint4*>(&s.H[kc][0] + r * 128 + ((q ^ (r & 7)) << 4)) =
                        make_uint4(tf32_round(v[j].x), tf32_round(v[j].y), tf32_round(v[j].z), tf32_round(v[j].w));
                }
            }
            fence_proxy_async_smem();
        };
        // pull the xp rows of time step t into L2 one step ahead (they stream from HBM otherwise):
        // 128 rows x 4 KB = 4096 lines, 16 per thread
        auto prefetch_xp = [&](int t) {
            // this CTA's 256 / NP units of every gate: 4 runs of 64 / NP quads x 2 KB (16 lines per quad), 16 / NP lines per thread
            const float* base = xp + (tile * T + t) * (256LL * 512);
            constexpr int kQuads = 64 / NP, kLinesGate = kQuads * 16;
#pragma unroll
            for (int i = 0; i < 16 / NP; ++i) {
                const int line = tid + i * kEpiThreads;               // 0 .. 4 kLinesGate - 1
                const int g = line / kLinesGate, l = line % kLinesGate;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (64 * g + kQuads * half) * 512 + l * 32));
            }
        };

        for (int step = 0; step < T; ++step) {
            const int t = dir == 0 ? step : T - 1 - step;
            if (tid == 0) stamp(step, 0);
            if (step + 1 < T) prefetch_xp(dir == 0 ? t + 1 : t - 1);
            if (tid == 0) stamp(step, 1);
            if (step == 0) {
                for (int pass = 0; pass < kPasses; ++pass)             // h0 = c0 = 0: z is the input projection alone
                    for (int uo = 0; uo < 32; uo += 16) cell16(step, t, 64 * (kPasses * half + pass) + usub + uo, 0u, true);
            } else {
                for (int pass = 0; pass < kPasses; ++pass) {
                    const long long P = static_cast<long long>(step - 1) * kPasses + pass;
                    const int buf = static_cast<int>(P & 1);
                    wait_or_trap(&s.tfull[buf], static_cast<uint32_t>((P >> 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (tid == 0) stamp(step, 2 + 2 * pass);
                    // pass columns: [0,64) i, [64,128) f, [128,192) c~, [192,256) o
                    for (int uo = 0; uo < 32; uo += 16)
                        cell16(step, t, 64 * (kPasses * half + pass) + usub + uo, static_cast<uint32_t>(buf * 256 + usub + uo), false);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s.tempty[buf]);        // this warp has drained the pass
                    if (tid == 0) stamp(step, 3 + 2 * pass);
                }
            }
            if (step + 1 < T) {
                // hand this half of h(t) to the peer CTA and wait for the other half: global stores -> gpu-scope
                // fence -> one remote mbarrier arrive (release.cluster) -> local wait (acquire.cluster)
                __threadfence();
                epi_bar_sync();            // all of this CTA's h(t) written; every MMA of this step has completed
                if (tid == 0) {
                    uint32_t remote;
                    // two barriers used alternately: a barrier can then never run two phases ahead of its waiter
                    // (the peer cannot pass step s+1's hand-over without this CTA's arrive for s+1)
                    uint64_t* pb = &s.peer[step & 1];
#pragma unroll
                    for (int pr = 1; pr < NP; ++pr) {                  // every peer's barrier expects NP - 1 arrivals
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(pb)), "r"((half + pr) % NP));
                        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
                    }
                    const uint32_t parity = static_cast<uint32_t>((step >> 1) & 1);
                    uint32_t ok = 0;
                    for (uint32_t i = 0; i < (1u << 24) && !ok; ++i)
                        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
                                     "selp.u32 %0, 1, 0, p;\n}\n"
                                     : "=r"(ok)
                                     : "r"(smem_u32(pb)), "r"(parity)
                                     : "memory");
                    if (!ok) asm volatile("trap;");
                }
                epi_bar_sync();
                if (tid == 0) stamp(step, 6);
                restage_h(step);
                epi_bar_sync();
                if (tid == 0) mbar_arrive(&s.hready);
                if (tid == 0) stamp(step, 7);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    // neither CTA leaves while the peer could still arrive on its barrier
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace

// Host: arrange U [256][1024] (Keras recurrent kernel, columns i|f|c|o) into the chunk stream the
// kernel consumes: chunk (pass, kc, kh) -> [4 slabs][256 n][4] with k = 32 kc + 16 kh + 4 slab + e,
// column n: gate = n/64, unit = 64*pass + n%64; TF32-rounded.  out: 64 chunks x 4096 floats.
long long mmla_lstm_arranged_floats() { return static_cast<long long>(kChunksPerStep) * kBChunkFloats; }
void mmla_lstm_arrange_weights(const float* U, float* out) {
    for (int pass = 0; pass < 4; ++pass)
        for (int kc = 0; kc < 8; ++kc)
            for (int kh = 0; kh < 2; ++kh) {
                float* chunk = out + (static_cast<long long>((pass * 8 + kc) * 2 + kh)) * kBChunkFloats;
                for (int slab = 0; slab < 4; ++slab)
                    for (int n = 0; n < 256; ++n)
                        for (int e = 0; e < 4; ++e) {
                            const int k = kc * 32 + kh * 16 + slab * 4 + e;
                            const int gate = n / 64;
                            const int unit = 64 * pass + (n % 64);
                            float v = U[static_cast<long long>(k) * 1024 + gate * 256 + unit];
                            uint32_t u;
                            memcpy(&u, &v, 4);
                            if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & ~0x1FFFu;
                            memcpy(&v, &u, 4);
                            chunk[(slab * 256 + n) * 4 + e] = v;
                        }
            }
}

// fp16 form: chunk (pass, kc, kh) -> [4 octets][256 n][8 halves] with k = 64 kc + 32 kh + 8 octet + e (K = 32 per 16 KB chunk),
// columns as above; round-to-nearest-even halves.  out: 32 chunks x 8192 halves (half the bytes of the TF32 stream).
long long mmla_lstm_arranged_halves() { return 32LL * 8192; }
void mmla_lstm_arrange_weights_f16(const float* U, uint16_t* out) {
    for (int pass = 0; pass < 4; ++pass)
        for (int kc = 0; kc < 4; ++kc)
            for (int kh = 0; kh < 2; ++kh) {
                uint16_t* chunk = out + (static_cast<long long>((pass * 4 + kc) * 2 + kh)) * 8192;
                for (int oct = 0; oct < 4; ++oct)
                    for (int n = 0; n < 256; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const int k = kc * 64 + kh * 32 + oct * 8 + e;
                            const int gate = n / 64;
                            const int unit = 64 * pass + (n % 64);
                            float v = U[static_cast<long long>(k) * 1024 + gate * 256 + unit];
                            v = v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v);
                            const __half h = __float2half_rn(v);
                            memcpy(&chunk[(oct * 256 + n) * 8 + e], &h, 2);
                        }
            }
}

static long long* g_lstm_stamps = nullptr;
static int g_lstm_stamp_cta = 0;
extern "C" __attribute__((visibility("default"))) void mmla_debug_lstm_stamps(long long* dev_stamps, int32_t cta) {
    g_lstm_stamps = dev_stamps;
    g_lstm_stamp_cta = cta;
}

// Scratch floats per direction for the row-tiled cell state (1x) and the h exchange (2x): 3 x this.
long long mmla_lstm_tile_floats(long long B) { return ((B + kRows - 1) / kRows) * 64LL * 512; }

// xp_*: row-tiled input projections (xproj_fused.cu); h_*: [B][256] final hidden state (output);
// scratch_*: 3 * mmla_lstm_tile_floats(B) floats each.
// f16 != 0: wr_f / wr_b point at the fp16 chunk streams (mmla_lstm_arrange_weights_f16)
int mmla_launch_lstm_fused(const float* xp_f, const float* xp_b, const float* wr_f, const float* wr_b, float* h_f,
                           float* h_b, float* scratch_f, float* scratch_b, long long B, int T, cudaStream_t st, int f16) {
    MMLA_REQUIRE(B > 0 && B < (1LL << 22) && T >= 1, MMLA_EINVAL, "lstm_fused: bad batch/T");
    static MmlaPerDeviceOnce attr_once;                          // cudaFuncSetAttribute is per device
    const bool attr_set = !attr_once.first();
    const int smem = static_cast<int>(sizeof(LstmSmem) + 1024);
    if (!attr_set) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(lstm_fused_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(lstm_fused_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(lstm_fused_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(lstm_fused_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    // four CTAs per tile while they all fit on the chip at once (a second wave would double the recurrence's latency)
    const long long tiles = (B + kRows - 1) / kRows;
    int np = (2 * tiles * 4 <= mmla_num_sms()) ? 4 : 2;
    if (const char* e = getenv("MMLA_LSTM_CLUSTER")) {
        const int v = atoi(e);
        if (v == 2 || v == 4) np = v;
    }
    LstmArgs a;
    a.xp[0] = xp_f; a.xp[1] = xp_b; a.wr[0] = wr_f; a.wr[1] = wr_b;
    a.h[0] = h_f; a.h[1] = h_b;
    a.c[0] = scratch_f; a.c[1] = scratch_b;
    a.hx[0] = scratch_f + mmla_lstm_tile_floats(B); a.hx[1] = scratch_b + mmla_lstm_tile_floats(B);
    a.B = static_cast<int>(B); a.T = T;
    a.f16 = f16;
    a.stamps = g_lstm_stamps; a.stamp_cta = g_lstm_stamp_cta;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(np) * static_cast<unsigned>(tiles), 2, 1);   // np CTAs (unit ranges) per 128-clip tile
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = static_cast<size_t>(smem);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(np);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (f16 && np == 4) MMLA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, lstm_fused_kernel<true, 4>, a));
    else if (f16) MMLA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, lstm_fused_kernel<true, 2>, a));
    else if (np == 4) MMLA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, lstm_fused_kernel<false, 4>, a));
    else MMLA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, lstm_fused_kernel<false, 2>, a));
    mmla_count_launch(f16 ? "lstm_fused_f16_kernel" : "lstm_fused_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}
