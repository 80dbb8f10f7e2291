// One whole ResNet STAGE of the speaker classifier in a single launch (sm_100a, TF32 tensor-core mode):
//
//   y0 = Conv1D_k1_s2(x) + Conv_k3( ReLU(BN2( Conv_k3( ReLU(BN1( MaxPool1D_2(x) )) ) )) )        pooled unit
//   y1 = y0 + Conv_k3( ReLU(BN2'( Conv_k3( ReLU(BN1'(y0)) ) )) )                                 plain unit
//   y2 = y1 + Conv_k3( ReLU(BN2''( Conv_k3( ReLU(BN1''(y1)) ) )) )                               plain unit
//
// — three consecutive `res_unit(x, filters, pool)` calls of SpeakerIdentification/scripts/
// speaker_identification.py:168-190, as the model builder stacks them (:205-216).
//
// Why a stage and not a unit (resunit_fused.cu): measured on B200, the per-unit kernel is bound by streaming
// its conv weights from L2 — every 128-row tile pulls the whole K x C weight stream through a small TMA ring,
// so each weight byte feeds only 128 rows of MMA (C = 128: 393 KB of weights per 64 KB tile; 1024 tiles x 3
// units = 1.3 GB of L2 -> SM traffic per stage, ~110 us at the chip's L2 throughput before any math).  This
// kernel changes the two ratios that matter:
//   * a CTA owns TILES (2 or 4) 128-row tiles and every weight chunk that comes through the ring feeds the
//     MMAs of all of them before the slot is released: L2 -> SM weight traffic / TILES;
//   * a tile is G = 128/T WHOLE clips (rows interleaved time-major, row = t*G + g, so a filter tap is a uniform
//     shift of G rows and 'same' padding is G zero halo rows), hence nothing a later unit needs lives in another
//     tile: the three units run back to back on the tile and the activations between them never touch HBM —
//     x in once, y out once per stage instead of per unit.
//
// The residual needs no buffer: conv2's TMEM accumulator is simply never cleared.  It starts as the shortcut
// GEMM (raw x[2t] x Ws, pooled unit), every conv2 accumulates on top, and the biases are added on read:
//      y_u = acc2 + (bs + b2_0 + .. + b2_u).
// Epilogue 2 of unit u therefore only has to produce the next unit's operand ReLU(BN1'(y_u)) -> TF32, exactly
// like epilogue 1 does for conv2; the last one stages y through shared memory for coalesced 128-bit stores.
//
// Operand scheme as in resunit_fused.cu: UMMA K-major no-swizzle slabs [channel quad][row][16 B], one buffer
// per tile, reused for every operand of the stage (conv MMAs have retired before an epilogue rewrites it).
// Warps 0..7 load / transform / run the epilogues, warp 8 streams weights (TMA bulk copies into an mbarrier
// ring), warp 9 issues tcgen05.mma (warp-uniform loop, one elected lane, descriptors in uniform registers).
#include <string.h>

#include "conv_common.cuh"

namespace {

constexpr int kRtotS = 137;                // rows per slab: 128 + 2*G halo (G <= 4), == 1 mod 8 (bank-friendly)
constexpr int kEpiS = 256;
constexpr int kThreadsS = kEpiS + 64;
constexpr int kUnitsS = 3;

struct StageUnit {
    const float* bn1_scale; const float* bn1_shift;
    const float* bn2_scale; const float* bn2_shift;
    const float* w1; const float* w2;      // conv_tc-arranged K-slab streams ([K/4][C][4])
    const float* b1; const float* b2;
};

struct StageArgs {
    const float* x;                        // [B][2T][CIN]
    float* y;                              // [B][T][C]
    StageUnit u[kUnitsS];
    const float* ws; const float* bs;      // stride-2 1x1 shortcut of the pooled unit
    // optional tail (last stage of the net): instead of y, write AveragePooling1D(4)(ReLU(BN(y))) -> pooled [B][T/4][C]
    const float* fin_scale; const float* fin_shift;
    float* pooled;
    int B, T, G;                           // T = OUTPUT time steps
    // tile -> CTA map: CTAs [0, n_full) own kTiles tiles each (whole waves); the rest own tail_tiles (<= kTiles) each,
    // so the last wave is not a wave of full-size CTAs on a fraction of the SMs
    int n_full, tail_tiles, n_tiles;
    long long* stamps;                     // diagnostics: clock64 timeline of CTA `stamp_cta` (null = off)
    int stamp_cta;
    // STEM variant (first stage, label pipeline): x is not read; the stem Conv1D(32, 4) runs in here from the MFCC-13 rows
    const float* cep;                      // [B][>= n_frames rows][16] fp32
    long long cep_clip_stride;             // floats
    int n_frames;                          // T of the clip's features (<= 256); rows beyond are the zero padding
    const float* stem_w;                   // conv_tc-arranged [4 taps x 10 quads][32][4] TF32
    const float* stem_b;                   // [32]
};

// Stem region of the STEM variant, laid over [ab | ring | stem_extra] (all dead once the stem MMAs have retired):
constexpr int kStemFRows = 137;                                // feature slab rows: pooled time t = -1 .. 128 (+ bank padding)
constexpr int kStemFBytes = (10 * kStemFRows * 16 + 127) / 128 * 128;   // one parity's slab [10 quads][rows][16 B]
constexpr int kStemWBytes = 4 * 40 * 32 * 4;                   // 20480
constexpr int kStemCepStride = 264;                            // >= 256 frames
constexpr int kStemCepBytes = (13 * kStemCepStride * 4 + 127) / 128 * 128;
constexpr int kStemOffFe = 0, kStemOffFo = kStemFBytes, kStemOffW = 2 * kStemFBytes, kStemOffCep = kStemOffW + kStemWBytes,
              kStemOffDlt = kStemOffCep + kStemCepBytes, kStemBytes = kStemOffDlt + kStemCepBytes;

template <int CIN, int C, bool STEM = false>
struct StageCfg {
    static_assert(!STEM || (CIN == 32 && C == 32), "the stem feeds the first stage");
    static constexpr int kTiles = 2;                           // 128-row tiles per CTA
    static constexpr int kChunkK = 32;                         // K per ring chunk
    static constexpr int kStages = C == 32 ? 4 : (C == 64 ? 2 : 3);
    static constexpr int kMinCtas = STEM ? 2 : (C == 32 ? 3 : (C == 64 ? 2 : 1));
    static constexpr int kChunkBytes = kChunkK * C * 4;
    // raw x[2t] of ONE tile (shortcut GEMM operand): slabs of 129 rows — with 128, the loader's lanes (which run over
    // channel quads for coalesced global loads) would all hit the same banks (16-way conflict, measured 4k cycles/tile)
    static constexpr int kA0Rows = 129;
    static constexpr int kA0Bytes = ((CIN / 4) * kA0Rows * 16 + 127) / 128 * 128;
    static constexpr int kExtra = kA0Bytes / kChunkBytes;      // ring stages that open up once the shortcut is done
    static constexpr int kStagesTot = kStages + kExtra;
    static constexpr int kSlabBytes = (C / 4) * kRtotS * 16;
    static constexpr int kStageTileBytes = 128 * (C + 4) * 4;  // output staging, rows C + 4 floats apart
    static constexpr int kTileBytes = ((kSlabBytes > kStageTileBytes ? kSlabBytes : kStageTileBytes) + 127) / 128 * 128;
    // With one CTA per SM the raw x tile ([G clips][2T rows][CIN] fp32, contiguous in HBM) is fetched by ONE bulk copy
    // per tile into that tile's (still empty) operand buffer, all tiles in flight from the first cycle of the CTA.
    static constexpr bool kTmaX = C >= 64;
    static constexpr int kRawTileBytes = 256 * CIN * 4;
    static_assert(!kTmaX || kRawTileBytes <= kTileBytes, "raw x tile is staged in the operand buffer");
};

template <int CIN, int C, bool STEM = false>
struct StageSmem {
    using Cfg = StageCfg<CIN, C, STEM>;
    alignas(128) unsigned char ab[Cfg::kTiles][Cfg::kTileBytes];   // per tile: the current MMA A operand / output staging
    // weight ring; its last kExtra stages double as `a0`, the shortcut GEMM's operand (raw x[2t] of one tile), and
    // join the ring when the last shortcut MMA has retired (short_done)
    alignas(128) unsigned char ring[Cfg::kStages * Cfg::kChunkBytes + Cfg::kA0Bytes];
    static constexpr int kMainBytes = Cfg::kTiles * Cfg::kTileBytes + Cfg::kStages * Cfg::kChunkBytes + Cfg::kA0Bytes;
    static_assert(kMainBytes % 128 == 0, "ab | ring | stem_extra are contiguous");
    alignas(128) unsigned char stem_extra[STEM ? (kStemBytes > kMainBytes ? kStemBytes - kMainBytes : 128) : 128];
    alignas(16) float prm0[3][C];          // STEM: stem bias, BN1 scale, BN1 shift of unit 0
    // per unit: [0] BN2 scale, [1] BN2 shift + b1*scale  (epilogue 1: ReLU(BN2(acc1 + b1)) = ReLU(acc1*[0] + [1]));
    //           [2] next BN1 scale, [3] next BN1 shift + run_u*scale, run_u = bs + b2_0 + .. + b2_u  (epilogue 2);
    //           [4] run_u  (y_u = acc2 + run_u, used by the final store)
    alignas(16) float prm[kUnitsS][5][C];
    alignas(8) uint64_t full[Cfg::kStagesTot];
    alignas(8) uint64_t empty[Cfg::kStagesTot];
    alignas(8) uint64_t a0_ready, a0_free; // shortcut operand of tile j written / its MMAs retired
    alignas(8) uint64_t short_done;        // every shortcut MMA retired: a0's bytes become ring stages
    alignas(8) uint64_t x_full[Cfg::kTiles];   // raw x tile landed (kTmaX)
    alignas(8) uint64_t a_ready[2];        // operand buffers written   (epilogue -> MMA): conv1, conv2
    alignas(8) uint64_t tfull[2];          // accumulators complete     (MMA -> epilogue): conv1, conv2
    alignas(8) uint64_t stem_wfull, f_ready, f_free, stem_done;   // STEM: weights landed / feature slabs built / consumed / all done
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t rs_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
// ReLU + round-to-nearest TF32 in one integer add/max (VIADDMNMX): a negative float is a negative int, and for x >= 0
// adding half a TF32 ulp to the bit pattern rounds the 10-bit mantissa (carry into the exponent included).  The low 13
// bits are left as they fall: kind::tf32 reads only the upper 19 bits of each operand word.
__device__ __forceinline__ uint32_t rs_relu_tf32(float x) {
    return static_cast<uint32_t>(max(static_cast<int>(__float_as_uint(x)) + 0x1000, 0));
}
__device__ __forceinline__ void rs_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    asm volatile("trap;");
}
__device__ __forceinline__ void rs_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rs_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ uint64_t rs_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool rs_elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void rs_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int CIN, int C, bool STEM>
__global__ void __launch_bounds__(kThreadsS, StageCfg<CIN, C, STEM>::kMinCtas) resstage_fused_kernel(const __grid_constant__ StageArgs a) {
    using Cfg = StageCfg<CIN, C, STEM>;
    using Smem = StageSmem<CIN, C, STEM>;
    constexpr int kTiles = Cfg::kTiles;
    constexpr int kQuads = C / 4;
    constexpr int kQuadsIn = CIN / 4;
    constexpr int kChunkK = Cfg::kChunkK;
    constexpr int kStages = Cfg::kStages;          // ring stages available from the start
    constexpr int kStagesTot = Cfg::kStagesTot;    // ... once the shortcut operand buffer has been released
    constexpr int kChunksS = CIN / kChunkK;                      // shortcut: K = CIN
    constexpr int kChunks1 = 3 * CIN / kChunkK;                  // conv1 of the pooled unit: K = 3*CIN
    constexpr int kChunks2 = 3 * C / kChunkK;                    // every other conv: K = 3*C
    static_assert(kChunksS >= 1 && kChunksS < kStages, "the shortcut weights stay in the ring while all tiles use them");
    constexpr uint32_t kChunkBytes = Cfg::kChunkBytes;
    constexpr int kCols = kTiles * 2 * C;                        // per tile: conv1 | running-y accumulator
    static_assert(kCols <= 512 && (kCols & (kCols - 1)) == 0 && kCols >= 32, "TMEM allocation");
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(C >> 3) << 17) |
                                (static_cast<uint32_t>(128 >> 4) << 24);
    extern __shared__ unsigned char smem_dyn[];
    // offset applied to the __shared__ array itself so accesses stay LDS/STS (an integer round-trip makes them generic)
    Smem& s = *reinterpret_cast<Smem*>(smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = a.G, T = a.T;
    const int bx = static_cast<int>(blockIdx.x);
    const int tile0 = bx < a.n_full ? bx * kTiles : a.n_full * kTiles + (bx - a.n_full) * a.tail_tiles;
    const int nt = bx < a.n_full ? kTiles : min(a.tail_tiles, a.n_tiles - tile0);   // this CTA's tiles (warp-uniform)
    const int clip_base = tile0 * G;                             // tile j holds clips [clip_base + j*G, +G)

    if (tid == 0) {
        for (int i = 0; i < kStagesTot; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], 1);
        }
        mbar_init(&s.short_done, 1);
        for (int i = 0; i < kTiles; ++i) mbar_init(&s.x_full[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.a_ready[i], 1);
            mbar_init(&s.tfull[i], 1);
        }
        mbar_init(&s.a0_ready, 1);
        mbar_init(&s.a0_free, 1);
        mbar_init(&s.stem_wfull, 1);
        mbar_init(&s.f_ready, 1);
        mbar_init(&s.f_free, 1);
        mbar_init(&s.stem_done, 1);
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
    const bool stamping = a.stamps != nullptr && static_cast<int>(blockIdx.x) == a.stamp_cta;
    auto stamp = [&](int slot) {
        if (stamping) a.stamps[slot] = clock64();
    };

    if (warp == 8) {
        // ================= TMA producer: shortcut, then per unit conv1 and conv2 weight chunks =================
        if (lane == 0) {
            if (STEM) {
                // stem weights first; the ring shares their bytes and starts when the last stem MMA has retired
                mbar_arrive_expect_tx(&s.stem_wfull, kStemWBytes);
                tma_bulk_g2s(&s.ab[0][0] + kStemOffW, a.stem_w, kStemWBytes, &s.stem_wfull);
                rs_wait(&s.stem_done, 0u);
            }
            if (Cfg::kTmaX) {
                const int clip_floats = 2 * T * CIN;
                for (int j = 0; j < nt; ++j) {
                    const int clip0 = clip_base + j * G;
                    const int valid = a.B - clip0 < G ? (a.B - clip0 < 0 ? 0 : a.B - clip0) : G;
                    if (valid > 0) {
                        const uint32_t bytes = static_cast<uint32_t>(valid * clip_floats * 4);
                        mbar_arrive_expect_tx(&s.x_full[j], bytes);
                        tma_bulk_g2s(&s.ab[j][0], a.x + static_cast<long long>(clip0) * clip_floats, bytes, &s.x_full[j]);
                    } else {
                        rs_arrive(&s.x_full[j]);
                    }
                }
            }
            int g = 0;
            // the conv_tc arrangement is a plain sequence of K-slabs ([K/4][C][4]): any multiple-of-4 K
            // granularity is a contiguous slice of it
            auto stream = [&](const float* w, int chunks) {
                for (int ch = 0; ch < chunks; ++ch, ++g) {
                    const int stg = g % kStagesTot, use = g / kStagesTot;
                    if (use > 0) rs_wait(&s.empty[stg], static_cast<uint32_t>((use - 1) & 1));
                    else if (stg >= kStages) rs_wait(&s.short_done, 0u);
                    mbar_arrive_expect_tx(&s.full[stg], kChunkBytes);
                    tma_bulk_g2s(&s.ring[stg * kChunkBytes], w + static_cast<long long>(ch) * (kChunkK * C), kChunkBytes, &s.full[stg]);
                }
            };
            stream(a.ws, kChunksS);
#pragma unroll 1
            for (int u = 0; u < kUnitsS; ++u) {
                stream(a.u[u].w1, u == 0 ? kChunks1 : kChunks2);
                stream(a.u[u].w2, kChunks2);
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ================= MMA issuer =================
        if (lane == 0) stamp(0);
        const uint64_t dA_main = rs_desc(smem_u32(&s.ab[0][0]), kRtotS * 16, 128);
        const uint64_t dA_short = rs_desc(smem_u32(&s.ring[kStages * kChunkBytes]), Cfg::kA0Rows * 16, 128);
        const uint64_t dB0 = rs_desc(smem_u32(&s.ring[0]), C * 16, 128);
        constexpr uint32_t kStageUnits = Cfg::kChunkBytes / 16;      // descriptor address units per ring stage
        constexpr uint32_t kTileUnits = Cfg::kTileBytes / 16;        // ... per tile operand buffer
        if (STEM) {
            // ---- stem: Conv1D(32, 4, 'same') of the clip's [256, 40] feature rows, split by output-row parity so that
            //      both x[2t] (-> TMEM columns of acc1) and x[2t+1] (-> acc2) come out with TMEM lane = pooled time t:
            //        x[2t]   = W0 f[2t-1] + W1 f[2t]   + W2 f[2t+1] + W3 f[2t+2]
            //        x[2t+1] = W0 f[2t]   + W1 f[2t+1] + W2 f[2t+2] + W3 f[2t+3]
            //      with the features de-interleaved into Fe[t] = f[2t], Fo[t] = f[2t+1] (slab row = t + 1) every tap is a
            //      whole-slab row shift again.  The max-pool and the shortcut operand then need no cross-lane traffic.
            const uint64_t dFe = rs_desc(smem_u32(&s.ab[0][0] + kStemOffFe), kStemFRows * 16, 128);
            const uint64_t dFo = rs_desc(smem_u32(&s.ab[0][0] + kStemOffFo), kStemFRows * 16, 128);
            const uint64_t dW = rs_desc(smem_u32(&s.ab[0][0] + kStemOffW), C * 16, 128);
            rs_wait(&s.stem_wfull, 0u);
#pragma unroll 1
            for (int j = 0; j < nt; ++j) {
                rs_wait(&s.f_ready, static_cast<uint32_t>(j & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (rs_elect_one()) {
#pragma unroll
                    for (int par = 0; par < 2; ++par) {            // output row parity: 0 -> acc1 columns, 1 -> acc2 columns
                        const uint32_t dcol = tmem + static_cast<uint32_t>(j * 2 * C + par * C);
#pragma unroll
                        for (int tap = 0; tap < 4; ++tap) {
                            // feature index 2t + par + tap - 1 = 2 (t + sh) + odd  ->  slab Fo/Fe, start row t + sh + 1
                            const int e = par + tap - 1;           // -1 .. 3
                            const int odd = e & 1;
                            const int sh = (e - odd) / 2;          // -1, 0, 1
                            const uint64_t dF = odd ? dFo : dFe;
#pragma unroll
                            for (int kq = 0; kq < 5; ++kq) {
                                const uint64_t ad = dF + static_cast<uint64_t>(2 * kq * kStemFRows + sh + 1);
                                const uint64_t bd = dW + static_cast<uint64_t>((tap * 10 + 2 * kq) * C);
                                const uint32_t acc = (tap | kq) != 0 ? 1u : 0u;
                                asm volatile(
                                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(dcol),
                                    "l"(ad), "l"(bd), "r"(kIdesc), "r"(acc)
                                    : "memory");
                            }
                        }
                    }
                    rs_commit(&s.f_free);
                    if (j == nt - 1) rs_commit(&s.stem_done);
                }
                __syncwarp();
            }
        }
        // ---- shortcut: raw x[2t] x Ws initialises the running-y accumulator of each tile as its operand arrives;
        //      the Ws chunks (ring slots 0..kChunksS-1) are released after the last tile has used them ----
#pragma unroll 1
        for (int j = 0; j < nt; ++j) {
            rs_wait(&s.a0_ready, static_cast<uint32_t>(j & 1));
            if (j == 0)
                for (int c = 0; c < kChunksS; ++c) rs_wait(&s.full[c], 0u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (rs_elect_one()) {
                const uint32_t dcol = tmem + static_cast<uint32_t>(j * 2 * C + C);
#pragma unroll
                for (int c = 0; c < kChunksS; ++c) {
#pragma unroll
                    for (int kk = 0; kk < kChunkK / 8; ++kk) {
                        const uint64_t ad = dA_short + static_cast<uint64_t>((c * (kChunkK / 4) + 2 * kk) * Cfg::kA0Rows);
                        const uint64_t bd = dB0 + static_cast<uint64_t>(c * kStageUnits + kk * 2 * C);
                        const uint32_t acc = (c != 0 || kk != 0) ? 1u : 0u;
                        asm volatile(
                            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(dcol),
                            "l"(ad), "l"(bd), "r"(kIdesc), "r"(acc)
                            : "memory");
                    }
                }
                rs_commit(&s.a0_free);
                if (j == nt - 1) {
                    for (int c = 0; c < kChunksS; ++c) rs_commit(&s.empty[c]);
                    rs_commit(&s.short_done);
                }
            }
            __syncwarp();
        }
        if (lane == 0) stamp(1);
        // Ring position kept as counters and descriptors advanced by adds: everything between the last MMA of a chunk
        // and the first MMA of the next one is a bubble the tensor pipe sees once its short queue (~4 MMAs) has drained.
        int stg = kChunksS;                                          // kChunksS < kStagesTot: no wrap yet
        uint32_t par = 0;
        uint64_t bd0 = dB0 + static_cast<uint64_t>(kChunksS * kStageUnits);
        // one K-chunk of MMAs for EVERY tile: A slabs start at descriptor `ad0` (slab + tap shift already applied);
        // accumulates into TMEM column block `dcol` of each tile
        auto chunk_mma = [&](uint64_t ad0, uint32_t dcol, bool first) {
            rs_wait(&s.full[stg], par);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (rs_elect_one()) {
#pragma unroll
                for (int j = 0; j < kTiles; ++j) {
                    if (j >= nt) break;
#pragma unroll
                    for (int kk = 0; kk < kChunkK / 8; ++kk) {
                        const uint64_t ad = ad0 + static_cast<uint64_t>(j * kTileUnits + 2 * kk * kRtotS);
                        const uint64_t bd = bd0 + static_cast<uint64_t>(kk * 2 * C);
                        const uint32_t acc = (!first || kk != 0) ? 1u : 0u;
                        asm volatile(
                            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + dcol + j * 2 * C),
                            "l"(ad), "l"(bd), "r"(kIdesc), "r"(acc)
                            : "memory");
                    }
                }
                rs_commit(&s.empty[stg]);
            }
            ++stg;
            bd0 += kStageUnits;
            if (stg == kStagesTot) {
                stg = 0;
                par ^= 1u;
                bd0 = dB0;
            }
        };
        // a k=3 convolution over `cin` input channels: K index = tap*cin + channel; tap j reads rows shifted by j*G
        // (halo rows = zero padding)
        auto conv = [&](int cin, uint32_t dcol, bool clear) {
            uint64_t tap0 = dA_main;
            for (int tap = 0; tap < 3; ++tap, tap0 += static_cast<uint64_t>(G)) {
                uint64_t ad = tap0;
                for (int c0 = 0; c0 < cin; c0 += kChunkK, ad += static_cast<uint64_t>((kChunkK / 4) * kRtotS))
                    chunk_mma(ad, dcol, clear && tap == 0 && c0 == 0);
            }
        };
#pragma unroll 1
        for (int u = 0; u < kUnitsS; ++u) {
            const uint32_t upar = static_cast<uint32_t>(u & 1);
            rs_wait(&s.a_ready[0], upar);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) stamp(2 + 4 * u);
            conv(u == 0 ? CIN : C, 0u, true);
            if (rs_elect_one()) rs_commit(&s.tfull[0]);
            __syncwarp();
            if (lane == 0) stamp(3 + 4 * u);
            // conv2 accumulates on top of the shortcut (u = 0) / of y_{u-1}: that IS the residual add
            rs_wait(&s.a_ready[1], upar);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) stamp(4 + 4 * u);
            conv(C, static_cast<uint32_t>(C), false);
            if (rs_elect_one()) rs_commit(&s.tfull[1]);
            __syncwarp();
            if (lane == 0) stamp(5 + 4 * u);
        }
    } else {
        // ================= warps 0..7: load/transform, epilogues =================
        constexpr int kRowsPerPass = kEpiS / kQuadsIn;            // rows covered by 256 threads at once
        constexpr int kIters = 128 / kRowsPerPass;
        const int qd = tid % kQuadsIn, rsub = tid / kQuadsIn;
        // x loads of the C = 32 stage (3 CTAs/SM hide the latency; its raw tile does not fit the operand buffer): lanes
        // run over channel quads (coalesced), tile j+1 is fetched while tile j's shortcut MMAs run.  The other stages
        // read the tile the producer warp's bulk copy staged in shared memory (RS_READ_TILE).
        constexpr int kPre = 1;
        float4 v[kPre][kIters], v2[kPre][kIters];
#define RS_LOAD_TILE(J)                                                                                                  \
    {                                                                                                                    \
        const int clip0_ = clip_base + (J) * G;                                                                          \
        _Pragma("unroll") for (int i = 0; i < kIters; ++i) {                                                             \
            const int r = rsub + i * kRowsPerPass;                                                                       \
            const int t = r / G, gg = r - t * G; /* row = t*G + g */                                                     \
            v[(J) % kPre][i] = make_float4(0.f, 0.f, 0.f, 0.f);                                                          \
            v2[(J) % kPre][i] = v[(J) % kPre][i];                                                                        \
            if (clip0_ + gg < a.B) {                                                                                     \
                const float* src = a.x + (static_cast<long long>(clip0_ + gg) * (2 * T) + 2 * t) * CIN + 4 * qd;         \
                v[(J) % kPre][i] = *reinterpret_cast<const float4*>(src);                                                \
                v2[(J) % kPre][i] = *reinterpret_cast<const float4*>(src + CIN);                                         \
            }                                                                                                            \
        }                                                                                                                \
    }
#define RS_READ_TILE(J)                                                                                                  \
    {                                                                                                                    \
        const int clip0_ = clip_base + (J) * G;                                                                          \
        const float* raw = reinterpret_cast<const float*>(&s.ab[J][0]);                                                  \
        _Pragma("unroll") for (int i = 0; i < kIters; ++i) {                                                             \
            const int r = rsub + i * kRowsPerPass;                                                                       \
            const int t = r / G, gg = r - t * G;                                                                         \
            v[0][i] = make_float4(0.f, 0.f, 0.f, 0.f);                                                                   \
            v2[0][i] = v[0][i];                                                                                          \
            if (clip0_ + gg < a.B) {                                                                                     \
                const float* src = raw + (gg * (2 * T) + 2 * t) * CIN + 4 * qd;                                          \
                v[0][i] = *reinterpret_cast<const float4*>(src);                                                         \
                v2[0][i] = *reinterpret_cast<const float4*>(src + CIN);                                                  \
            }                                                                                                            \
        }                                                                                                                \
    }
        if (tid == 0) stamp(16);
        if (!STEM && !Cfg::kTmaX) RS_LOAD_TILE(0)
        const float4 sc1 = *reinterpret_cast<const float4*>(a.u[0].bn1_scale + 4 * qd);
        const float4 sh1 = *reinterpret_cast<const float4*>(a.u[0].bn1_shift + 4 * qd);
        // Parameters: all global loads first, then the shared stores.  (Interleaved, every load has to wait for the
        // store before it — the pointers are generic, so the compiler must assume they may alias shared memory —
        // and ~17 serialised L2 round trips cost 10k cycles at the head of every CTA.)
        for (int i = tid; i < C; i += kEpiS) {
            float pb1[kUnitsS], pb2[kUnitsS], ps1[kUnitsS], ph1[kUnitsS], ps2[kUnitsS], ph2[kUnitsS];
            const float pbs = __ldg(a.bs + i);
#pragma unroll
            for (int u = 0; u < kUnitsS; ++u) {
                pb1[u] = __ldg(a.u[u].b1 + i); pb2[u] = __ldg(a.u[u].b2 + i);
                ps1[u] = __ldg(a.u[u].bn1_scale + i); ph1[u] = __ldg(a.u[u].bn1_shift + i);
                ps2[u] = __ldg(a.u[u].bn2_scale + i); ph2[u] = __ldg(a.u[u].bn2_shift + i);
            }
            if (STEM) {
                s.prm0[0][i] = __ldg(a.stem_b + i);
                s.prm0[1][i] = ps1[0];
                s.prm0[2][i] = ph1[0];
            }
            float run = pbs;
#pragma unroll
            for (int u = 0; u < kUnitsS; ++u) {
                run += pb2[u];
                const float scn = u + 1 < kUnitsS ? ps1[u + 1 < kUnitsS ? u + 1 : u] : 0.f;
                const float shn = u + 1 < kUnitsS ? ph1[u + 1 < kUnitsS ? u + 1 : u] : 0.f;
                s.prm[u][0][i] = ps2[u];
                s.prm[u][1][i] = fmaf(pb1[u], ps2[u], ph2[u]);
                s.prm[u][2][i] = scn;
                s.prm[u][3][i] = fmaf(run, scn, shn);
                s.prm[u][4][i] = run;                             // y_u = acc2 + bs + b2_0 + .. + b2_u
            }
        }
        // zero the halo rows (rows [0,G) and [G+128, G+128+G)) of a tile; the epilogues never touch them
        auto zero_halo = [&](int j) {
            for (int i = tid; i < kQuads * 2 * G; i += kEpiS) {
                const int q = i / (2 * G), h = i - q * (2 * G);
                const int row = h < G ? h : 128 + h;
                *reinterpret_cast<uint4*>(&s.ab[j][0] + (q * kRtotS + row) * 16) = make_uint4(0u, 0u, 0u, 0u);
            }
        };
        if (!STEM && !Cfg::kTmaX)
            for (int j = 0; j < nt; ++j) zero_halo(j);
        unsigned char* a0 = &s.ring[kStages * kChunkBytes];
        if (STEM) {
            // ================= stem, phase A: per clip, MFCC-13 rows -> delta, delta-delta -> parity-split feature slabs ======
            // (G = 1: a tile is one clip.)  Same arithmetic as stem_fused.cu's FROM_CEP path — `delta(feat, 2)` of
            // speaker_identification.py:141-151 with edge replication inside the clip's T frames, zero rows from T to 256.
            unsigned char* R = &s.ab[0][0];
            float (*cep)[kStemCepStride] = reinterpret_cast<float (*)[kStemCepStride]>(R + kStemOffCep);
            float (*dlt)[kStemCepStride] = reinterpret_cast<float (*)[kStemCepStride]>(R + kStemOffDlt);
            const int Tn = a.n_frames, Tm1 = Tn - 1;
            // rows 0 (t = -1) and 129 (t = 128) of both slabs are the conv's zero padding; nothing below rewrites them
            for (int i = tid; i < 2 * 2 * 10; i += kEpiS) {
                const int par = i / 20, rem = i - 20 * par;
                const int q = rem >> 1, row = (rem & 1) ? 129 : 0;
                *reinterpret_cast<uint4*>(R + (par ? kStemOffFo : kStemOffFe) + (q * kStemFRows + row) * 16) = make_uint4(0u, 0u, 0u, 0u);
            }
            // cepstra rows of BOTH clips are fetched up front (<= 256 rows x 4 float4 = 4 per thread and clip): one exposed
            // HBM latency per CTA
            float4 cv[kTiles][4];
#pragma unroll
            for (int j = 0; j < kTiles; ++j) {
                const int clip = clip_base + j;
                const float* cc = a.cep + static_cast<long long>(j < nt && clip < a.B ? clip : 0) * a.cep_clip_stride;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int idx = tid + i * kEpiS;
                    cv[j][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < nt && idx < Tn * 4) cv[j][i] = *reinterpret_cast<const float4*>(cc + static_cast<long long>(idx >> 2) * 16 + 4 * (idx & 3));
                }
            }
#pragma unroll
            for (int j = 0; j < kTiles; ++j) {
                if (j >= nt) break;
                const int clip = clip_base + j;
                const bool live = clip < a.B;
                // cepstra rows 0 .. T-1 (16-float rows, 13 used), channel-major so that lanes = consecutive frames
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int idx = tid + i * kEpiS;
                    const int rr = idx >> 2, q4 = idx & 3;
                    if (idx < Tn * 4) {
                        const float e[4] = {cv[j][i].x, cv[j][i].y, cv[j][i].z, cv[j][i].w};
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (4 * q4 + u < 13) cep[4 * q4 + u][rr] = e[u];
                    }
                }
                rs_epi_sync();
                {   // delta: one thread per frame
                    const int td = tid;
                    if (td < Tn) {
                        const int m2 = max(td - 2, 0), m1 = max(td - 1, 0), p1 = min(td + 1, Tm1), p2 = min(td + 2, Tm1);
#pragma unroll
                        for (int c = 0; c < 13; ++c) {
                            float acc = -2.f * cep[c][m2];
                            acc = fmaf(-1.f, cep[c][m1], acc);
                            acc = fmaf(1.f, cep[c][p1], acc);
                            acc = fmaf(2.f, cep[c][p2], acc);
                            dlt[c][td] = acc * 0.1f;
                        }
                    }
                }
                rs_epi_sync();
                if (j > 0) rs_wait(&s.f_free, static_cast<uint32_t>((j - 1) & 1));   // stem MMAs of clip j-1 retired
                // feature rows 0 .. T-1 of the clip's [256, 40] tensor -> slab of the row's parity, slab row sI/2 + 1.  Work
                // item = (row, third of the channels: MFCC | delta | delta-delta): 3T items over 256 threads, so the longest
                // thread does 2 items instead of 39 channels.  Rows T .. 255 (zero padding) are written once: the slabs
                // are reused by the CTA's next clip and n_frames is the same for all clips.
                if (j == 0) {
                    for (int i = tid; i < (256 - Tn) * 10; i += kEpiS) {
                        const int q = i % 10, sI = Tn + i / 10;
                        *reinterpret_cast<uint4*>(R + ((sI & 1) ? kStemOffFo : kStemOffFe) + (q * kStemFRows + (sI >> 1) + 1) * 16) =
                            make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                for (int it = tid; it < 3 * Tn; it += kEpiS) {
                    const int third = (it >= Tn ? 1 : 0) + (it >= 2 * Tn ? 1 : 0);
                    const int sI = it - third * Tn;
                    unsigned char* F = R + ((sI & 1) ? kStemOffFo : kStemOffFe) + ((sI >> 1) + 1) * 16;
                    float f[13];
                    if (!live) {
#pragma unroll
                        for (int c = 0; c < 13; ++c) f[c] = 0.f;
                    } else if (third == 0) {
#pragma unroll
                        for (int c = 0; c < 13; ++c) f[c] = cep[c][sI];
                    } else if (third == 1) {
#pragma unroll
                        for (int c = 0; c < 13; ++c) f[c] = dlt[c][sI];
                    } else {
                        const int m2 = max(sI - 2, 0), m1 = max(sI - 1, 0), p1 = min(sI + 1, Tm1), p2 = min(sI + 2, Tm1);
#pragma unroll
                        for (int c = 0; c < 13; ++c) {
                            float acc = -2.f * dlt[c][m2];
                            acc = fmaf(-1.f, dlt[c][m1], acc);
                            acc = fmaf(1.f, dlt[c][p1], acc);
                            acc = fmaf(2.f, dlt[c][p2], acc);
                            f[c] = acc * 0.1f;
                        }
                    }
                    // channel ch = 13 third + c -> quad ch/4, element ch%4 (compile-time per branch)
                    if (third == 0) {
#pragma unroll
                        for (int c = 0; c < 13; ++c)
                            *reinterpret_cast<uint32_t*>(F + (c >> 2) * (kStemFRows * 16) + (c & 3) * 4) = rs_tf32(f[c]);
                    } else if (third == 1) {
#pragma unroll
                        for (int c = 0; c < 13; ++c)
                            *reinterpret_cast<uint32_t*>(F + ((13 + c) >> 2) * (kStemFRows * 16) + ((13 + c) & 3) * 4) = rs_tf32(f[c]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 13; ++c)
                            *reinterpret_cast<uint32_t*>(F + ((26 + c) >> 2) * (kStemFRows * 16) + ((26 + c) & 3) * 4) = rs_tf32(f[c]);
                        *reinterpret_cast<uint32_t*>(F + 9 * (kStemFRows * 16) + 12) = 0u;     // channel 39
                    }
                }
                fence_proxy_async_smem();
                rs_epi_sync();
                if (tid == 0) rs_arrive(&s.f_ready);
                if (tid == 0) stamp(40 + j);
            }
            // ================= stem, phase B: x[2t] | x[2t+1] (TMEM, lane = t) -> shortcut operand, conv1 operand ==========
            rs_wait(&s.stem_done, 0u);                            // every stem MMA retired: the stem region is dead
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 0) stamp(44);
            const int brow = 32 * (warp & 3) + lane;              // TMEM lane = pooled time t
            const int bcol = 16 * (warp >> 2);                    // this warp's 16 of the 32 channels
#pragma unroll 1
            for (int j = 0; j < nt; ++j) {
                zero_halo(j);
                uint32_t xe[16], xo[16];
                const uint32_t taddr = tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + static_cast<uint32_t>(j * 2 * C + bcol);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                    "%15}, [%16];\n"
                    : "=r"(xe[0]), "=r"(xe[1]), "=r"(xe[2]), "=r"(xe[3]), "=r"(xe[4]), "=r"(xe[5]), "=r"(xe[6]), "=r"(xe[7]),
                      "=r"(xe[8]), "=r"(xe[9]), "=r"(xe[10]), "=r"(xe[11]), "=r"(xe[12]), "=r"(xe[13]), "=r"(xe[14]), "=r"(xe[15])
                    : "r"(taddr));
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                    "%15}, [%16];\n"
                    : "=r"(xo[0]), "=r"(xo[1]), "=r"(xo[2]), "=r"(xo[3]), "=r"(xo[4]), "=r"(xo[5]), "=r"(xo[6]), "=r"(xo[7]),
                      "=r"(xo[8]), "=r"(xo[9]), "=r"(xo[10]), "=r"(xo[11]), "=r"(xo[12]), "=r"(xo[13]), "=r"(xo[14]), "=r"(xo[15])
                    : "r"(taddr + static_cast<uint32_t>(C)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (j > 0) rs_wait(&s.a0_free, static_cast<uint32_t>((j - 1) & 1));   // shortcut MMAs of tile j-1 retired
#pragma unroll
                for (int k = 0; k < 16; k += 4) {
                    const int col = bcol + k;
                    const float4 bv = *reinterpret_cast<const float4*>(&s.prm0[0][col]);
                    const float4 sc = *reinterpret_cast<const float4*>(&s.prm0[1][col]);
                    const float4 sh = *reinterpret_cast<const float4*>(&s.prm0[2][col]);
                    const float e0 = __uint_as_float(xe[k]) + bv.x, e1 = __uint_as_float(xe[k + 1]) + bv.y;
                    const float e2 = __uint_as_float(xe[k + 2]) + bv.z, e3 = __uint_as_float(xe[k + 3]) + bv.w;
                    const float o0 = __uint_as_float(xo[k]) + bv.x, o1 = __uint_as_float(xo[k + 1]) + bv.y;
                    const float o2 = __uint_as_float(xo[k + 2]) + bv.z, o3 = __uint_as_float(xo[k + 3]) + bv.w;
                    *reinterpret_cast<uint4*>(a0 + ((col >> 2) * Cfg::kA0Rows + brow) * 16) =
                        make_uint4(rs_tf32(e0), rs_tf32(e1), rs_tf32(e2), rs_tf32(e3));
                    *reinterpret_cast<uint4*>(&s.ab[j][0] + ((col >> 2) * kRtotS + G + brow) * 16) =
                        make_uint4(rs_relu_tf32(fmaf(fmaxf(e0, o0), sc.x, sh.x)), rs_relu_tf32(fmaf(fmaxf(e1, o1), sc.y, sh.y)),
                                   rs_relu_tf32(fmaf(fmaxf(e2, o2), sc.z, sh.z)), rs_relu_tf32(fmaf(fmaxf(e3, o3), sc.w, sh.w)));
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                fence_proxy_async_smem();
                rs_epi_sync();
                if (tid == 0) rs_arrive(&s.a0_ready);
                if (tid == 0) stamp(32 + j);
            }
        } else {
        // ---- per tile: raw x[2t] -> shortcut operand; ReLU(BN1(max(x[2t], x[2t+1]))) -> conv1 operand ----
#pragma unroll
        for (int j = 0; j < kTiles; ++j) {
            if (j >= nt) break;
            if (Cfg::kTmaX) {
                rs_wait(&s.x_full[j], 0u);
                RS_READ_TILE(j)
                rs_epi_sync();                                    // everyone holds its part: the buffer may be rewritten
                zero_halo(j);
            }
            if (j > 0) rs_wait(&s.a0_free, static_cast<uint32_t>((j - 1) & 1));   // shortcut MMAs of tile j-1 retired
#pragma unroll
            for (int i = 0; i < kIters; ++i) {
                const int r = rsub + i * kRowsPerPass;
                float4 p = v[j % kPre][i];
                const float4 p2 = v2[j % kPre][i];
                *reinterpret_cast<uint4*>(a0 + (qd * Cfg::kA0Rows + r) * 16) =
                    make_uint4(rs_tf32(p.x), rs_tf32(p.y), rs_tf32(p.z), rs_tf32(p.w));
                p.x = fmaxf(p.x, p2.x); p.y = fmaxf(p.y, p2.y);
                p.z = fmaxf(p.z, p2.z); p.w = fmaxf(p.w, p2.w);
                *reinterpret_cast<uint4*>(&s.ab[j][0] + (qd * kRtotS + G + r) * 16) =
                    make_uint4(rs_tf32(fmaxf(fmaf(p.x, sc1.x, sh1.x), 0.f)), rs_tf32(fmaxf(fmaf(p.y, sc1.y, sh1.y), 0.f)),
                               rs_tf32(fmaxf(fmaf(p.z, sc1.z, sh1.z), 0.f)), rs_tf32(fmaxf(fmaf(p.w, sc1.w, sh1.w), 0.f)));
            }
            if (!Cfg::kTmaX && j + 1 < nt) RS_LOAD_TILE(j + 1)  // latency overlaps the barrier round trip below
            fence_proxy_async_smem();
            rs_epi_sync();
            if (tid == 0) rs_arrive(&s.a0_ready);
            if (tid == 0) stamp(32 + j);
        }
        }
        if (tid == 0) rs_arrive(&s.a_ready[0]);

        // TMEM lane = tile row; warps 0..3 take the lower half of the columns, warps 4..7 the upper
        const int row = 32 * (warp & 3) + lane;
        constexpr int kColsPerWarp = C / 2;
        const int cbase = (warp >> 2) * kColsPerWarp;
        const uint32_t trow = tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16);
        // this warp's kColsPerWarp accumulator columns of one tile: every tcgen05.ld in flight before the single wait
        auto ld_cols = [&](uint32_t taddr, uint32_t (&z)[kColsPerWarp]) {
#pragma unroll
            for (int c0 = 0; c0 < kColsPerWarp; c0 += 16)
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                    "%15}, [%16];\n"
                    : "=r"(z[c0 + 0]), "=r"(z[c0 + 1]), "=r"(z[c0 + 2]), "=r"(z[c0 + 3]), "=r"(z[c0 + 4]), "=r"(z[c0 + 5]),
                      "=r"(z[c0 + 6]), "=r"(z[c0 + 7]), "=r"(z[c0 + 8]), "=r"(z[c0 + 9]), "=r"(z[c0 + 10]), "=r"(z[c0 + 11]),
                      "=r"(z[c0 + 12]), "=r"(z[c0 + 13]), "=r"(z[c0 + 14]), "=r"(z[c0 + 15])
                    : "r"(taddr + static_cast<uint32_t>(c0)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        };
        // accumulator block `dcol` of every tile -> * scale + shift -> ReLU -> TF32 -> that tile's operand buffer
        auto to_operand = [&](uint32_t dcol, const float* scale, const float* shift) {
#pragma unroll 1
            for (int j = 0; j < nt; ++j) {
                uint32_t z[kColsPerWarp];
                ld_cols(trow + dcol + static_cast<uint32_t>(j * 2 * C + cbase), z);
#pragma unroll
                for (int k = 0; k < kColsPerWarp; k += 4) {
                    const int col = cbase + k;
                    const float4 sc = *reinterpret_cast<const float4*>(scale + col);
                    const float4 sh = *reinterpret_cast<const float4*>(shift + col);
                    *reinterpret_cast<uint4*>(&s.ab[j][0] + ((col >> 2) * kRtotS + G + row) * 16) =
                        make_uint4(rs_relu_tf32(fmaf(__uint_as_float(z[k + 0]), sc.x, sh.x)),
                                   rs_relu_tf32(fmaf(__uint_as_float(z[k + 1]), sc.y, sh.y)),
                                   rs_relu_tf32(fmaf(__uint_as_float(z[k + 2]), sc.z, sh.z)),
                                   rs_relu_tf32(fmaf(__uint_as_float(z[k + 3]), sc.w, sh.w)));
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            fence_proxy_async_smem();
            rs_epi_sync();
        };

#pragma unroll 1
        for (int u = 0; u < kUnitsS; ++u) {
            const uint32_t par = static_cast<uint32_t>(u & 1);
            // ---- epilogue 1: conv1 accumulator -> +b1 -> BN2 -> ReLU -> TF32 -> operand buffer ----
            rs_wait(&s.tfull[0], par);                            // conv1's MMAs (readers of the buffers) have retired
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 0) stamp(18 + 4 * u);
            to_operand(0u, &s.prm[u][0][0], &s.prm[u][1][0]);
            if (tid == 0) rs_arrive(&s.a_ready[1]);
            if (tid == 0) stamp(19 + 4 * u);
            // ---- epilogue 2: y_u = running accumulator + running bias; next unit's operand = ReLU(BN1'(y_u)) ----
            rs_wait(&s.tfull[1], par);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 0) stamp(20 + 4 * u);
            if (u + 1 < kUnitsS) {
                to_operand(static_cast<uint32_t>(C), &s.prm[u][2][0], &s.prm[u][3][0]);
                if (tid == 0) rs_arrive(&s.a_ready[0]);
                if (tid == 0) stamp(21 + 4 * u);
            }
        }

        // ---- last epilogue: y_2 -> staging (operand buffers are dead) -> coalesced 128-bit stores ----
        constexpr int kStride = C + 4;                            // floats per staged row (conflict-free STS/LDS.128)
#pragma unroll 1
        for (int j = 0; j < nt; ++j) {
            float* stg = reinterpret_cast<float*>(&s.ab[j][0]);
            uint32_t z[kColsPerWarp];
            ld_cols(trow + static_cast<uint32_t>(j * 2 * C + C + cbase), z);
#pragma unroll
            for (int k = 0; k < kColsPerWarp; k += 4) {
                const int col = cbase + k;
                const float4 bv = *reinterpret_cast<const float4*>(&s.prm[kUnitsS - 1][4][col]);
                *reinterpret_cast<float4*>(stg + row * kStride + col) =
                    make_float4(__uint_as_float(z[k]) + bv.x, __uint_as_float(z[k + 1]) + bv.y, __uint_as_float(z[k + 2]) + bv.z,
                                __uint_as_float(z[k + 3]) + bv.w);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        fence_proxy_async_smem();                                 // staged rows are read by the bulk-copy engine
        rs_epi_sync();
        if (tid == 0) stamp(30);
        if (a.pooled != nullptr) {
            // tail of the net folded in: BN -> ReLU -> mean over 4 consecutive time steps, from the staged rows
            // (row = t*G + g), summed in the order of the stand-alone bn_relu_avgpool4_kernel; y itself is never written
            const int q = tid % kQuads;                           // constant per thread (kEpiS % kQuads == 0)
            const float4 fsc = *reinterpret_cast<const float4*>(a.fin_scale + 4 * q);
            const float4 fsh = *reinterpret_cast<const float4*>(a.fin_shift + 4 * q);
            const int To = T / 4;
#pragma unroll 1
            for (int idx = tid; idx < nt * 32 * kQuads; idx += kEpiS) {
                const int j = idx / (32 * kQuads), pr = (idx / kQuads) % 32;
                const int to = pr / G, gg = pr - to * G;
                const int clip = clip_base + j * G + gg;
                if (clip >= a.B) continue;
                const float* stg = reinterpret_cast<const float*>(&s.ab[j][0]);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 v = *reinterpret_cast<const float4*>(stg + ((4 * to + k) * G + gg) * kStride + 4 * q);
                    acc.x += fmaxf(fmaf(v.x, fsc.x, fsh.x), 0.f);
                    acc.y += fmaxf(fmaf(v.y, fsc.y, fsh.y), 0.f);
                    acc.z += fmaxf(fmaf(v.z, fsc.z, fsh.z), 0.f);
                    acc.w += fmaxf(fmaf(v.w, fsc.w, fsh.w), 0.f);
                }
                *reinterpret_cast<float4*>(a.pooled + (static_cast<long long>(clip) * To + to) * C + 4 * q) =
                    make_float4(acc.x * 0.25f, acc.y * 0.25f, acc.z * 0.25f, acc.w * 0.25f);
            }
        } else {
        // one bulk copy per output row (C floats, contiguous in y): asynchronous, so the warps are done once the
        // copies are issued; they only wait for the shared-memory READS before the CTA (and its smem) goes away
#pragma unroll
        for (int i = 0; i < kTiles * 128 / kEpiS; ++i) {
            const int idx = tid + i * kEpiS;
            const int j = idx >> 7, r = idx & 127;
            const int t = r / G, gg = r - t * G;
            const int clip = clip_base + j * G + gg;
            if (j < nt && clip < a.B) {
                float* dst = a.y + (static_cast<long long>(clip) * T + t) * C;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                             "r"(smem_u32(&s.ab[j][0] + r * kStride * 4)), "r"(static_cast<uint32_t>(C * 4))
                             : "memory");
            }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        if (tid == 0) stamp(31);
#undef RS_LOAD_TILE
#undef RS_READ_TILE
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(static_cast<uint32_t>(kCols))
                     : "memory");
}

template <int CIN, int C, bool STEM = false>
int launch_stage(const StageArgs& a, cudaStream_t st) {
    static MmlaPerDeviceOnce attr_once;                          // cudaFuncSetAttribute is per device
    const bool attr_set = !attr_once.first();
    const int smem = static_cast<int>(sizeof(StageSmem<CIN, C, STEM>) + 128);
    static_assert(sizeof(StageSmem<CIN, C, STEM>) + 128 <= 227 * 1024, "stage kernel shared memory");
    static_assert(StageCfg<CIN, C, STEM>::kExtra >= 1, "the shortcut operand buffer becomes at least one ring stage");
    if (!attr_set) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(resstage_fused_kernel<CIN, C, STEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    constexpr int kTilesCta = StageCfg<CIN, C, STEM>::kTiles;
    StageArgs b = a;
    const int tiles = (a.B + a.G - 1) / a.G;
    const int slots = mmla_num_sms() * StageCfg<CIN, C, STEM>::kMinCtas;       // CTAs resident at once
    b.n_tiles = tiles;
    b.n_full = (tiles / (slots * kTilesCta)) * slots;                          // whole waves of full-size CTAs
    const int rest = tiles - b.n_full * kTilesCta;
    b.tail_tiles = rest > 0 ? (rest + slots - 1) / slots : 1;                   // <= kTilesCta
    const unsigned grid = static_cast<unsigned>(b.n_full + (rest + b.tail_tiles - 1) / b.tail_tiles);
    resstage_fused_kernel<CIN, C, STEM><<<grid, kThreadsS, smem, st>>>(b);
    mmla_count_launch(STEM ? "stem_resstage_fused_kernel" : "resstage_fused_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}

}  // namespace

static long long* g_stage_stamps = nullptr;
static int g_stage_stamp_cta = 0;
extern "C" __attribute__((visibility("default"))) void mmla_debug_resstage_stamps(long long* dev_stamps, int32_t cta) {
    g_stage_stamps = dev_stamps;
    g_stage_stamp_cta = cta;
}

// x: [B][2T][Cin], y: [B][T][C] fp32 NHWC (H = 1); T = output steps in {128, 64, 32}.
// p[u] = {bn1_scale, bn1_shift, w1, b1, bn2_scale, bn2_shift, w2, b2} of unit u (u = 0: the pooled unit);
// w* are the conv_tc-arranged streams of the k=3 convolutions, ws / bs the stride-2 1x1 shortcut of unit 0.
// fin_scale / fin_shift / pooled (all or none): the net's tail BN -> ReLU -> AveragePooling1D(4) applied to the stage
// output, written to pooled [B][T/4][C]; y is then not written (and may be null).
int mmla_launch_resstage_fused(const float* x, float* y, long long B, int T, int Cin, int C, const float* const (*p)[8],
                               const float* ws, const float* bs, const float* fin_scale, const float* fin_shift,
                               float* pooled, cudaStream_t st) {
    MMLA_REQUIRE(T >= 32 && T <= 128 && 128 % T == 0, MMLA_EUNSUP, "resstage_fused: T=%d unsupported", T);
    MMLA_REQUIRE(B > 0 && B < (1LL << 24), MMLA_EINVAL, "resstage_fused: bad batch");
    MMLA_REQUIRE(ws && bs, MMLA_EINVAL, "resstage_fused: the first unit of a stage is the pooled one");
    StageArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.y = y;
    for (int u = 0; u < kUnitsS; ++u) {
        a.u[u].bn1_scale = p[u][0]; a.u[u].bn1_shift = p[u][1]; a.u[u].w1 = p[u][2]; a.u[u].b1 = p[u][3];
        a.u[u].bn2_scale = p[u][4]; a.u[u].bn2_shift = p[u][5]; a.u[u].w2 = p[u][6]; a.u[u].b2 = p[u][7];
    }
    a.ws = ws; a.bs = bs;
    MMLA_REQUIRE(pooled == nullptr || (fin_scale && fin_shift && T % 4 == 0), MMLA_EINVAL, "resstage_fused: bad pooled tail");
    MMLA_REQUIRE(pooled != nullptr || y != nullptr, MMLA_EINVAL, "resstage_fused: no output");
    a.fin_scale = fin_scale; a.fin_shift = fin_shift; a.pooled = pooled;
    a.B = static_cast<int>(B); a.T = T; a.G = 128 / T;
    if (g_stage_stamps) {                  // one 64-slot row per stage, keyed by C: 32 -> row 0, 64 -> 1, 128 -> 2
        a.stamps = g_stage_stamps + 64 * (C == 32 ? 0 : (C == 64 ? 1 : 2));
        a.stamp_cta = g_stage_stamp_cta;
    }
    if (Cin == 32 && C == 32) return launch_stage<32, 32>(a, st);
    if (Cin == 32 && C == 64) return launch_stage<32, 64>(a, st);
    if (Cin == 64 && C == 128) return launch_stage<64, 128>(a, st);
    mmla_set_error("resstage_fused: Cin=%d C=%d unsupported", Cin, C);
    return MMLA_EUNSUP;
}

// First stage with the stem folded in (label pipeline): cepstra [B][rows >= n_frames][16] fp32 (cep_clip_stride floats per
// clip) -> Conv1D(32, 4, 'same') of the [256, 39] feature rows built on the fly (delta, delta-delta, zero rows from
// n_frames to 256) -> the stage's three residual units -> y [B][128][32].  stem_w: the stem's conv_tc-arranged weights
// with 40 input channels (channel 39 = 0), stem_b: [32].  Same results as stem_fused.cu (FROM_CEP) + the plain stage.
int mmla_launch_resstage_stem_fused(const float* cepstra, long long cep_clip_stride, int n_frames, const float* stem_w,
                                    const float* stem_b, float* y, long long B, const float* const (*p)[8], const float* ws,
                                    const float* bs, cudaStream_t st) {
    MMLA_REQUIRE(B > 0 && B < (1LL << 24), MMLA_EINVAL, "resstage_stem_fused: bad batch");
    MMLA_REQUIRE(cepstra && stem_w && stem_b && y && ws && bs, MMLA_EINVAL, "resstage_stem_fused: null argument");
    MMLA_REQUIRE(n_frames >= 1 && n_frames <= 256 && cep_clip_stride >= static_cast<long long>(n_frames) * 16, MMLA_EINVAL,
                 "resstage_stem_fused: bad cepstra geometry");
    StageArgs a;
    memset(&a, 0, sizeof(a));
    a.y = y;
    for (int u = 0; u < kUnitsS; ++u) {
        a.u[u].bn1_scale = p[u][0]; a.u[u].bn1_shift = p[u][1]; a.u[u].w1 = p[u][2]; a.u[u].b1 = p[u][3];
        a.u[u].bn2_scale = p[u][4]; a.u[u].bn2_shift = p[u][5]; a.u[u].w2 = p[u][6]; a.u[u].b2 = p[u][7];
    }
    a.ws = ws; a.bs = bs;
    a.B = static_cast<int>(B); a.T = 128; a.G = 1;
    a.cep = cepstra; a.cep_clip_stride = cep_clip_stride; a.n_frames = n_frames; a.stem_w = stem_w; a.stem_b = stem_b;
    if (g_stage_stamps) {
        a.stamps = g_stage_stamps;
        a.stamp_cta = g_stage_stamp_cta;
    }
    return launch_stage<32, 32, true>(a, st);
}
