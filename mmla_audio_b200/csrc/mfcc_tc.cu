// Tensor-core MFCC for sm_100a: python_speech_features-style MFCC where the 512-point DFT of every
// frame runs on the tcgen05 tensor cores as a two-stage Cooley-Tukey transform (512 = 32 x 16)
// with fp16 hi+lo operand splitting (three MMA passes per stage, fp32 accumulation in TMEM),
// i.e. fp32-grade accuracy at tensor-core rate.  Everything else (pre-emphasis, twiddles, power
// spectrum, mel filterbank, log, DCT-II, lifter) is fused around it; frames, spectra and log-mels
// never touch HBM.
//
// Replaces (same outputs as csrc/mfcc.cu, which stays as the general-parameter kernel):
//   mfcc(sig, rate, winlen=0.025, winstep=0.01, nfft=512)
//     SpeakerIdentification/scripts/speaker_identification.py:89,285,341,386
//     SpeakerIdentification/scripts/speaker_identification_post_processing.py:256
//   delta(feat, 2) twice + concatenate + zero-pad rows (mfcc_finish_kernel)
//     SpeakerIdentification/scripts/speaker_identification.py:141-151,387-395
//
// Index maps:  n = 16 n1 + n2  (n1 < 32, n2 < 16),  k = k1 + 32 k2  (k1 < 32, k2 < 16)
//   stage 1   S[n2][k1] = sum_n1 y[16 n1 + n2] W32^(n1 k1)            k1 = 0..16 (real input)
//   twiddle   T[n2][k1] = S[n2][k1] W512^(n2 k1)
//   stage 2   X[k1 + 32 k2] = sum_n2 T[n2][k1] W16^(n2 k2)
// Bins 0..256 are the outputs with k1 <= 16 plus their conjugates (|X[512-k]| = |X[k]|).
//
// Work unit: a "group" = 16 consecutive frames of one clip; a tile = 4 groups = 64 frames.
//   * TMA warp: bulk-copies the group's int16 PCM (2920 samples) into a 2-slot raw ring.
//   * signal warps (4): int16 -> float, pre-emphasis, x0.5, split into fp16 hi / lo, stored ONCE
//     per sample as 8-sample chunks de-interleaved by chunk parity h.  In that layout frame f,
//     sample 16 n1 + 8 h + r sits at plane_h + 160 f + 16 n1 (bytes) + 2 r, which is exactly the
//     UMMA MN-major no-swizzle canonical layout with SBO = 160 B (next frame) and LBO = 128 B
//     (next 8 n1): the 2.5x overlapping frames are never materialised, the tensor core reads
//     them through the descriptor stride.  The frame's zero padding (n1 >= 25) is zero rows of B1.
//   * MMA warp: stage 1 = per (group, h) a 128x32x32 product, rows (frame, r), K = n1, columns
//     (k1, re/im), accumulators D1 in TMEM columns [0,256); stage 2 = per pair of k1 a 128x32x32
//     product, rows (k1 parity, frame), K = (n2, re/im), columns (k2 pair: re re im im), D2 in [256,512).
//     Each product is hi*hi + lo*hi + hi*lo (three kind::f16 passes).
//   * convert warps (4): read D1 (tcgen05.ld), apply the twiddle (fp32), scale by 2^-5, split to
//     fp16 hi / lo and transpose into the stage-2 A operand (K-major, no swizzle) in smem.
//   * epilogue warps (4): read D2, |X|^2, triangular mel filters with COMPILE-TIME bin->filter
//     maps and immediate weights (mfcc_tc_tables.inc), the k1 = 16 column by a small fp32 DFT,
//     exchange the two k1-parity partial sums through smem, log, DCT x lifter, store.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"
#include "mfcc_tc_tables.inc"

namespace {

constexpr int kFrameLen = 400, kStep = 160;
constexpr int kGroupFrames = 16, kTileGroups = 4, kTileFrames = 64;
constexpr int kGroupSamples = 2912;                 // 15*160 + 32*16: samples one group's MMAs touch
constexpr int kChunks = kGroupSamples / 8;          // 364 8-sample chunks
constexpr int kPlaneBytes = 2944;                   // 184 chunks x 16 B
constexpr int kPlaneH = kPlaneBytes + 64;           // h = 1 plane: = 64 mod 128 past h = 0, i.e. 16 banks apart
constexpr int kPlaneLo = kPlaneH + kPlaneBytes;     // the fp16 "lo" pair of planes follows the "hi" pair
constexpr int kSlotBytes = 2 * kPlaneLo;            // [hi|lo][h] = 11904 bytes
constexpr int kRawBytes = 5888;                     // 8 + 2912 samples, padded
constexpr int kA2Lbo = 2112;                        // K-group stride of the stage-2 A operand
constexpr int kA2Bytes = 4 * kA2Lbo;                // one (k1 pair, hi|lo) operand: 128 rows x 32 halves
constexpr float kS1 = 0.5f;                         // stage-1 operand scale (|y| <= 64553 -> fp16 range)
constexpr float kS2 = 0.03125f;                     // stage-2 operand scale (|S| <= 8.1e5 -> fp16 range)
constexpr float kEps = 2.220446049250313e-16f;
constexpr int kThreads = 448;                       // warps 0-3 epilogue, 4-7 convert, 8-11 signal, 12 TMA, 13 MMA
constexpr int kXchStride = 44;                     // >= nfilt + 3; 176-byte rows: 128-bit accesses conflict-free

struct TcSmem {
    alignas(128) unsigned char a2[8][2][kA2Bytes];
    alignas(128) unsigned char planes[kTileGroups][kSlotBytes];
    alignas(128) unsigned char raw[2][kRawBytes];
    alignas(128) unsigned char b1[2][2][2048];       // [h][hi|lo]: B1[col][n1], K-major core matrices; the h = 1 copy
                                                     // carries the W64^k1 half of the twiddle (angle (2 n1 + 1) k1 / 64)
    alignas(128) unsigned char b2[2][2048];          // hi, lo: B2[(k2,c')][(n2,c)]
    float2 k16[8][16];                               // [k2][n2] = s2 * W512^(n2 (16 + 32 k2))
    float dct[40][16];                               // [m][c] DCT-II(ortho) x lifter; c 0..6 at 0..6, 7..12 at 8..13
    float s16[2][kTileFrames][20];                   // S[n2][16] of the tile (k1 = 16 column), double buffered
    alignas(16) float xch[kTileFrames][kXchStride];  // mel partial sums / log-mels between k1 parities
    alignas(8) uint64_t raw_full[2], raw_empty[2];
    alignas(8) uint64_t plane_full[kTileGroups], plane_empty[kTileGroups];
    alignas(8) uint64_t d1_full[kTileGroups], d1_empty[kTileGroups];
    alignas(8) uint64_t s2_done, d2_empty;
    uint32_t tmem_base;
};
static_assert(sizeof(TcSmem) <= 227 * 1024, "TcSmem exceeds the 227 KB a CTA can own");
constexpr int kConstSmemBytes = 6 * 2048 + 1024 + 2560;   // b1 [h][hi|lo], b2 hi|lo, k16, dct -> shared memory
constexpr int kConstBytes = kConstSmemBytes + 1024;       // + twr [k1][r] (read once into registers)

struct TcParams {
    const int16_t* pcm;
    const int64_t* clip_off;       // device, or null (uniform)
    const int32_t* clip_len_arr;   // device, or null (uniform)
    const int2* groups;            // device (clip, first frame) per group, or null (uniform arithmetic)
    const unsigned char* consts;
    float* out;
    float* dbg;
    long long* prof;               // DBG builds: clock64 stamps of CTA 0, 32 slots per tile (after the dump area)
    long long n_groups;
    long long clip_stride;
    long long out_clip_stride;
    int groups_per_clip;
    int clip_len;
    int row_stride;                // floats between output rows (>= 13 or 39)
    int zero_tail;                 // columns [13, 13 + zero_tail) of every row are zeroed by the epilogue (no finish pass)
    int pad_frames;
    int append_energy;
    float preemph;
};

struct Group {
    long long clip, clip_off;
    int len, f0, n_real;
    bool active;
};

__device__ __forceinline__ Group decode_group(const TcParams& p, int G) {
    Group g;
    g.active = G < static_cast<int>(p.n_groups);
    g.clip = 0; g.clip_off = 0; g.len = 0; g.f0 = 0; g.n_real = 0;
    if (!g.active) return g;
    if (p.groups) {
        const int2 e = p.groups[G];
        g.clip = e.x;
        g.f0 = e.y;
    } else {
        const int c = static_cast<int>(static_cast<unsigned>(G) / static_cast<unsigned>(p.groups_per_clip));
        g.clip = c;
        g.f0 = (G - c * p.groups_per_clip) * kGroupFrames;
    }
    g.clip_off = p.clip_off ? p.clip_off[g.clip] : g.clip * p.clip_stride;
    g.len = p.clip_len_arr ? p.clip_len_arr[g.clip] : p.clip_len;
    const int T = g.len <= kFrameLen ? 1 : 1 + (g.len - kFrameLen + kStep - 1) / kStep;
    g.n_real = p.pad_frames > 0 ? min(T, p.pad_frames) : T;
    return g;
}

// One out-of-line copy of the group arithmetic for the roles of mfcc_tc2_kernel (code size: its roles share a small
// instruction cache).  mfcc_tc_kernel keeps the inlined form: the out-of-line call, a compact signal loop and a rolled TMA
// loop were measured on it and ran 3-6 % slower (profiles/r02/experiment_notes.txt).
__device__ __noinline__ Group decode_group2(const TcParams& p, int G) { return decode_group(p, G); }

// mbarrier wait that parks the warp in hardware (suspend-time hint) instead of polling, so waiting
// roles do not steal issue slots from working ones; traps instead of hanging the GPU on a logic error.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
__device__ __noinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t i = 0; i < (1u << 22); ++i)
        if (mbar_try_wait_hint(bar, parity, 100000u)) return;
    asm volatile("trap;");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// One lane of the (converged) warp; the surrounding control flow stays warp-uniform so descriptors live in
// uniform registers and UTCHMMA issues without a per-instruction R2UR waterfall.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit_to(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
        : "memory");
}
// tcgen05.ld of 32 consecutive columns of this thread's TMEM lane; asynchronous until tmem_wait32.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// Waits for the outstanding tcgen05.ld; the registers are in/out operands so no use can be hoisted above it.
__device__ __forceinline__ void tmem_wait32(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // epilogue warps only

// fp16 hi + lo split of two floats: hi = rn(v), lo = rn(v - hi); low half of each word = first value.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---------------------------------------------------------------------------------------------
// Epilogue arithmetic with compile-time bin -> filter maps
// ---------------------------------------------------------------------------------------------
template <int NF, int B>
__device__ __forceinline__ void mel_add(float q, float (&mel)[NF], float& edge) {
    using Tab = MfccTcTab<NF>;
    constexpr int fa = Tab::fa[B];
    constexpr int fb = Tab::fb[B];
    constexpr float we = Tab::we[B];
    if constexpr (fa >= 0) {
        constexpr float w = Tab::wa[B];
        mel[fa] = fmaf(w, q, mel[fa]);
    }
    if constexpr (fb >= 0) {
        constexpr float w = Tab::wb[B];
        mel[fb] = fmaf(w, q, mel[fb]);
    }
    if constexpr (we != 0.f) edge = fmaf(we, q, edge);     // first / last segment, bin 256: the energy remainder
}

// |X|^2 of two bins at once with the packed fp32 pipe: q = re*re + im*im on (bin 2i, bin 2i+1).
__device__ __forceinline__ void sq2(uint32_t re0, uint32_t re1, uint32_t im0, uint32_t im1, float& q0, float& q1) {
    asm("{\n.reg .b64 a, b, c;\n"
        "mov.b64 a, {%2, %3};\n"
        "mov.b64 b, {%4, %5};\n"
        "mul.rn.f32x2 c, b, b;\n"
        "fma.rn.f32x2 c, a, a, c;\n"
        "mov.b64 {%0, %1}, c;\n}\n"
        : "=f"(q0), "=f"(q1)
        : "r"(re0), "r"(re1), "r"(im0), "r"(im1));
}
template <int NF, int P, int J, int K2>
__device__ __forceinline__ void bin_add(float q, float (&mel)[NF], float& edge) {
    constexpr int k1 = 2 * J + P;
    constexpr bool valid = (k1 != 0) || (K2 <= 8);         // k1 = 0: k2 = 9..15 duplicate k2 = 7..1
    if constexpr (valid) {
        constexpr int b = (k1 == 0) ? 32 * K2 : (K2 < 8 ? k1 + 32 * K2 : 512 - k1 - 32 * K2);
        mel_add<NF, b>(q, mel, edge);
    }
}
// D2 block J of this row: columns are grouped (re 2i, re 2i+1, im 2i, im 2i+1) so that both squares
// of two neighbouring k2 are one packed multiply + one packed fma.
template <int NF, int P, int J, int I>
__device__ __forceinline__ void pair_acc(const uint32_t (&v)[32], float (&mel)[NF], float& edge) {
    constexpr int k1 = 2 * J + P;
    if constexpr (k1 != 0 || 2 * I <= 8) {
        float q0, q1;
        sq2(v[4 * I], v[4 * I + 1], v[4 * I + 2], v[4 * I + 3], q0, q1);
        bin_add<NF, P, J, 2 * I>(q0, mel, edge);
        bin_add<NF, P, J, 2 * I + 1>(q1, mel, edge);
    }
}
template <int NF, int P, int J, int... I>
__device__ __forceinline__ void block_acc(const uint32_t (&v)[32], float (&mel)[NF], float& edge,
                                          std::integer_sequence<int, I...>) {
    (pair_acc<NF, P, J, I>(v, mel, edge), ...);
}
// Two D2 blocks per step through two register buffers: the tcgen05.ld of the next block is in
// flight while the current one is accumulated.
template <int NF, int P, int J>
__device__ __forceinline__ void d2_pair(uint32_t tbase, uint32_t (&va)[32], uint32_t (&vb)[32], float (&mel)[NF],
                                        float& edge) {
    using Seq = std::make_integer_sequence<int, 8>;
    tmem_wait32(va);                                       // block J landed
    tmem_ld32_issue(tbase + 32 * (J + 1), vb);
    block_acc<NF, P, J>(va, mel, edge, Seq{});
    tmem_wait32(vb);                                       // block J+1 landed
    if constexpr (J + 2 < 8) tmem_ld32_issue(tbase + 32 * (J + 2), va);
    block_acc<NF, P, J + 1>(vb, mel, edge, Seq{});
}

template <int NF, int P>
__device__ __forceinline__ void epilogue_tile(const TcParams& p, TcSmem& s, uint32_t tmem, int q, int lane, int it,
                                              long long tile) {
    const int f = (32 * q + lane) & 63;                     // frame of the tile this lane owns
    float mel[NF];
#pragma unroll
    for (int m = 0; m < NF; ++m) mel[m] = 0.f;
    float edge = 0.f;
    {
        const uint32_t tbase = tmem + (static_cast<uint32_t>(32 * q) << 16) + 256u;
        uint32_t va[32], vb[32];
        tmem_ld32_issue(tbase, va);
        d2_pair<NF, P, 0>(tbase, va, vb, mel, edge);
        d2_pair<NF, P, 2>(tbase, va, vb, mel, edge);
        d2_pair<NF, P, 4>(tbase, va, vb, mel, edge);
        d2_pair<NF, P, 6>(tbase, va, vb, mel, edge);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.d2_empty);                // D2 drained: stage 2 of the next tile may start
    // ---- k1 = 16 column: X[16 + 32 k2] = sum_n2 S16[n2] W512^(n2 (16 + 32 k2)), k2 = 4P .. 4P+3 ----
    {
        const float4* sr = reinterpret_cast<const float4*>(&s.s16[it & 1][f][0]);
        float q16[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4* cf = reinterpret_cast<const float4*>(&s.k16[4 * P + i][0]);
            float re = 0.f, im = 0.f;
#pragma unroll
            for (int n4 = 0; n4 < 4; ++n4) {
                const float4 sv = sr[n4];
                const float4 c0 = cf[2 * n4], c1 = cf[2 * n4 + 1];
                re = fmaf(sv.x, c0.x, re); im = fmaf(sv.x, c0.y, im);
                re = fmaf(sv.y, c0.z, re); im = fmaf(sv.y, c0.w, im);
                re = fmaf(sv.z, c1.x, re); im = fmaf(sv.z, c1.y, im);
                re = fmaf(sv.w, c1.z, re); im = fmaf(sv.w, c1.w, im);
            }
            q16[i] = fmaf(re, re, im * im);
        }
        mel_add<NF, 16 + 32 * (4 * P + 0)>(q16[0], mel, edge);
        mel_add<NF, 16 + 32 * (4 * P + 1)>(q16[1], mel, edge);
        mel_add<NF, 16 + 32 * (4 * P + 2)>(q16[2], mel, edge);
        mel_add<NF, 16 + 32 * (4 * P + 3)>(q16[3], mel, edge);
    }
    // ---- combine the two k1 parities in smem, then log + DCT as compact loops -------------------
    // Row layout (176-byte rows, every access 128-bit => conflict-free): [0, NFP) mel sums (zero padded
    // past NF), [NFP] energy remainder, [NFP + 1 + parity] energy share of that parity's filter half.
    constexpr int NFP = (NF + 3) & ~3, NG = NFP / 4;
    static_assert(NFP + 3 <= kXchStride, "xch row too short");
    float4* x4 = reinterpret_cast<float4*>(&s.xch[f][0]);
    {
        float part[kXchStride];
#pragma unroll
        for (int m = 0; m < kXchStride; ++m) part[m] = 0.f;
#pragma unroll
        for (int m = 0; m < NF; ++m) part[m] = mel[m];
        part[NFP] = edge;
        epi_bar();                                           // previous tile's readers are done
        if (P == 1) {
#pragma unroll
            for (int i = 0; i < kXchStride / 4; ++i) x4[i] = make_float4(part[4 * i], part[4 * i + 1], part[4 * i + 2], part[4 * i + 3]);
        }
        epi_bar();
        if (P == 0) {
#pragma unroll
            for (int i = 0; i <= NG; ++i) {
                float4 v = x4[i];
                v.x += part[4 * i]; v.y += part[4 * i + 1]; v.z += part[4 * i + 2]; v.w += part[4 * i + 3];
                x4[i] = v;
            }
        }
        epi_bar();
    }
    float e_tail;
    {   // each parity takes half of the filter quads: energy share, zero guard, log
        constexpr int g0 = P == 0 ? 0 : (NG + 1) / 2, g1 = P == 0 ? (NG + 1) / 2 : NG;
        float esum = 0.f;
        float4 v[g1 - g0];
#pragma unroll
        for (int i = g0; i < g1; ++i) v[i - g0] = x4[i];
#pragma unroll
        for (int i = g0; i < g1; ++i) {
            float4& q = v[i - g0];
            esum += (q.x + q.y) + (q.z + q.w);                // padding entries are exact zeros
            q.x = __logf(q.x == 0.f ? kEps : q.x); q.y = __logf(q.y == 0.f ? kEps : q.y);
            q.z = __logf(q.z == 0.f ? kEps : q.z); q.w = __logf(q.w == 0.f ? kEps : q.w);
            x4[i] = q;
        }
        e_tail = esum;
    }
    // the two energy shares meet through one word each (different banks per lane: stride 44 words, +1)
    s.xch[f][NFP + 1 + P] = e_tail;
    epi_bar();
    const int g = f >> 4;
    const Group gr = decode_group(p, static_cast<int>(tile) * kTileGroups + g);
    const int t = gr.f0 + (f & 15);
    constexpr int nc = P == 0 ? 7 : 6;
    float c[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) c[i] = 0.f;
#pragma unroll 2
    for (int i = 0; i < NG; ++i) {
        const float4 lm4 = x4[i];
        const float lm[4] = {lm4.x, lm4.y, lm4.z, lm4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4* w4 = reinterpret_cast<const float4*>(&s.dct[4 * i + u][8 * P]);   // rows >= NF are zero
            const float4 wa = w4[0], wb = w4[1];
            c[0] = fmaf(wa.x, lm[u], c[0]); c[1] = fmaf(wa.y, lm[u], c[1]); c[2] = fmaf(wa.z, lm[u], c[2]);
            c[3] = fmaf(wa.w, lm[u], c[3]); c[4] = fmaf(wb.x, lm[u], c[4]); c[5] = fmaf(wb.y, lm[u], c[5]);
            c[6] = fmaf(wb.z, lm[u], c[6]);
        }
    }
    if (P == 0 && p.append_energy) {
        const float4 tail = x4[NG];                          // [edge, share 0, share 1, -]
        const float e = tail.x + tail.y + tail.z;            // = 8 * sum_b q_b = the psf frame energy
        c[0] = __logf(e == 0.f ? kEps : e);
    }
    if (gr.active && t < gr.n_real) {
        float* o = p.out + gr.clip * p.out_clip_stride + static_cast<long long>(t) * p.row_stride + 7 * P;
#pragma unroll
        for (int i = 0; i < nc; ++i) o[i] = c[i];
        if (P == 1)
            for (int i = 0; i < p.zero_tail; ++i) o[6 + i] = 0.f;
    }
}

#define TC_STAMP(slot)                                                                         \
    do {                                                                                       \
        if constexpr (DBG) {                                                                   \
            if (p.prof && blockIdx.x == 0 && lane == 0 && it < 64) p.prof[it * 32 + (slot)] = clock64(); \
        }                                                                                      \
    } while (0)

template <int NF, bool DBG>
__global__ void __launch_bounds__(kThreads, 1) mfcc_tc_kernel(const __grid_constant__ TcParams p) {
    // no static __shared__ in this kernel, so the dynamic window starts at the CTA's shared base (1 KB aligned);
    // using the array directly keeps every access in the shared address space (LDS/STS, not generic LD/ST)
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    TcSmem& s = *reinterpret_cast<TcSmem*>(smem_dyn);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n_tiles = (p.n_groups + kTileGroups - 1) / kTileGroups;
    const int my_tiles = blockIdx.x < n_tiles ? static_cast<int>((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    // ---- one-time setup -------------------------------------------------------------------------
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.consts);
        uint4* dst = reinterpret_cast<uint4*>(&s.b1[0][0][0]);       // b1, b2, k16 are contiguous
        for (int i = tid; i < kConstSmemBytes / 16; i += kThreads) dst[i] = src[i];
        uint4* z = reinterpret_cast<uint4*>(&s.a2[0][0][0]);          // operand buffers: finite everywhere
        for (int i = tid; i < static_cast<int>(sizeof(s.a2) + sizeof(s.planes)) / 16; i += kThreads)
            z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.raw_full[i], 1);
            mbar_init(&s.raw_empty[i], 4);
        }
        for (int i = 0; i < kTileGroups; ++i) {
            mbar_init(&s.plane_full[i], 4);
            mbar_init(&s.plane_empty[i], 1);
            mbar_init(&s.d1_full[i], 1);
            mbar_init(&s.d1_empty[i], 4);
        }
        mbar_init(&s.s2_done, 1);
        mbar_init(&s.d2_empty, 4);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();                                // zero-filled operands + constants -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s.tmem_base;

    if (warp == 12) {
        // ================= TMA producer: raw PCM of each group =================
        if (lane == 0) {
            for (int it = 0; it < my_tiles; ++it) {
                const long long tile = blockIdx.x + static_cast<long long>(it) * gridDim.x;
                for (int g = 0; g < kTileGroups; ++g) {
                    const int gi = it * kTileGroups + g, rs = gi & 1;
                    if (gi >= 2) wait_or_trap(&s.raw_empty[rs], static_cast<uint32_t>(((gi >> 1) - 1) & 1));
                    if (g == 0) TC_STAMP(0);
                    const Group gr = decode_group(p, static_cast<int>(tile) * kTileGroups + g);
                    if (gr.active) {
                        const long long gs0 = gr.clip_off + static_cast<long long>(gr.f0) * kStep;
                        const long long gA = gs0 >= 8 ? gs0 - 8 : 0;
                        long long gB = gs0 + kGroupSamples;
                        const long long clip_end = gr.clip_off + ((gr.len + 7) & ~7);
                        if (gB > clip_end) gB = clip_end;
                        const uint32_t bytes = static_cast<uint32_t>((gB - gA) * 2);
                        mbar_arrive_expect_tx(&s.raw_full[rs], bytes);
                        tma_bulk_g2s(&s.raw[rs][0], p.pcm + gA, bytes, &s.raw_full[rs]);
                    } else {
                        mbar_arrive(&s.raw_full[rs]);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 13) {
        // ================= MMA issuer (whole warp runs the loop, one elected lane issues) =================
        {
            constexpr uint32_t kIdesc2 = (1u << 4) | (static_cast<uint32_t>(32 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
            constexpr uint32_t kIdesc1 = kIdesc2 | (1u << 15);     // A is MN-major in stage 1
            const uint32_t b1a = smem_u32(&s.b1[0][0][0]), b2a = smem_u32(&s.b2[0][0]);
            // Base descriptors once; every MMA then adds a compile-time offset (>> 4) to the 14-bit address field
            // (all operands live below 256 KB, so the field cannot carry).  Measured (scripts/microbench/umma_rate.cu):
            // an M=128 K=16 SS-mode MMA costs max(N/2, (A+B bytes)/128) ~ 45 cycles at N=32 when issued like this,
            // 80-200 cycles when the descriptors are rebuilt per instruction.
            const uint64_t dA1 = desc_noswz(smem_u32(&s.planes[0][0]), 128, kStep);
            const uint64_t dB1 = desc_noswz(b1a, 128, 512);
            const uint64_t dA2 = desc_noswz(smem_u32(&s.a2[0][0][0]), kA2Lbo, 128);
            const uint64_t dB2 = desc_noswz(b2a, 128, 512);
            auto stage1 = [&](int it) {
#pragma unroll 1
                for (int g = 0; g < kTileGroups; ++g) {
                    wait_or_trap(&s.plane_full[g], static_cast<uint32_t>(it & 1));
                    if (it >= 1) wait_or_trap(&s.d1_empty[g], static_cast<uint32_t>((it - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (g == 0) TC_STAMP(1);
                    if (g == 3) TC_STAMP(2);
                    const uint64_t dAg = dA1 + static_cast<uint64_t>((g * kSlotBytes) >> 4);
                    const uint32_t dcol = tmem + g * 64;
                    if (elect_one()) {
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                                for (int ks = 0; ks < 2; ++ks)
                                    umma_f16(dcol + h * 32,
                                             dAg + static_cast<uint64_t>(((pass == 1 ? kPlaneLo : 0) + h * kPlaneH + ks * 256) >> 4),
                                             dB1 + static_cast<uint64_t>((h * 4096 + (pass == 2 ? 2048 : 0) + ks * 256) >> 4), kIdesc1,
                                             (pass | ks) != 0 ? 1u : 0u);
                        umma_commit_to(&s.plane_empty[g]);
                        umma_commit_to(&s.d1_full[g]);
                    }
                    __syncwarp();
                }
            };
            auto stage2 = [&](int it) {
                wait_or_trap(&s.d1_empty[kTileGroups - 1], static_cast<uint32_t>(it & 1));   // conversion complete
                if (it >= 1) wait_or_trap(&s.d2_empty, static_cast<uint32_t>((it - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                TC_STAMP(3);
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    const uint64_t dAj = dA2 + static_cast<uint64_t>((j * 2 * kA2Bytes) >> 4);
                    const uint32_t dcol = tmem + 256 + j * 32;
                    if (elect_one()) {
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks)
                                umma_f16(dcol, dAj + static_cast<uint64_t>(((pass == 1 ? kA2Bytes : 0) + ks * 2 * kA2Lbo) >> 4),
                                         dB2 + static_cast<uint64_t>(((pass == 2 ? 2048 : 0) + ks * 256) >> 4), kIdesc2,
                                         (pass | ks) != 0 ? 1u : 0u);
                        if (j == 7) umma_commit_to(&s.s2_done);
                    }
                    __syncwarp();
                }
                TC_STAMP(4);
            };
            if (my_tiles > 0) stage1(0);
            for (int it = 0; it < my_tiles; ++it) {
                if (it + 1 < my_tiles) stage1(it + 1);
                stage2(it);
            }
        }
        __syncwarp();
    } else if (warp >= 8) {
        // ================= signal warps: PCM -> pre-emphasised fp16 hi/lo planes =================
        const int ts = tid - 256;
        const float pre = p.preemph;
        for (int it = 0; it < my_tiles; ++it) {
            const long long tile = blockIdx.x + static_cast<long long>(it) * gridDim.x;
            for (int g = 0; g < kTileGroups; ++g) {
                const int gi = it * kTileGroups + g, rs = gi & 1;
                const Group gr = decode_group(p, static_cast<int>(tile) * kTileGroups + g);
                wait_or_trap(&s.raw_full[rs], static_cast<uint32_t>((gi >> 1) & 1));
                if (it >= 1) wait_or_trap(&s.plane_empty[g], static_cast<uint32_t>((it - 1) & 1));
                if (warp == 8) TC_STAMP(8 + 2 * g);
                const long long gs0 = gr.clip_off + static_cast<long long>(gr.f0) * kStep;
                const int delta = gs0 >= 8 ? 8 : 0;                      // raw index of the group's first sample
                const int n_base = gr.f0 * kStep;
                const int len = gr.active ? gr.len : 0;
                const uint4* raw4 = reinterpret_cast<const uint4*>(&s.raw[rs][0]);
                const unsigned short* raw16 = reinterpret_cast<const unsigned short*>(&s.raw[rs][0]);
                unsigned char* slot = &s.planes[g][0];
                // chunks ts, ts+128, ts+256 of the group: all loads first, then three independent convert chains
                uint4 w[3];
                uint32_t prev16[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int c8 = ts + 128 * u;
                    const int r0 = delta + 8 * c8;
                    w[u] = make_uint4(0u, 0u, 0u, 0u);
                    prev16[u] = 0u;
                    if (c8 < kChunks && n_base + 8 * c8 < len) w[u] = raw4[r0 >> 3];
                    // the sample before this chunk is the last one of the neighbouring lane's chunk; only lane 0
                    // reads it from shared memory (a 2-byte load per lane would be a 4-way bank conflict)
                    const uint32_t up = __shfl_up_sync(0xffffffffu, w[u].w, 1) >> 16;
                    prev16[u] = lane > 0 ? up : ((c8 < kChunks && r0 > 0) ? raw16[r0 - 1] : 0u);
                }
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int c8 = ts + 128 * u;
                    if (c8 >= kChunks) break;
                    const int n0 = n_base + 8 * c8;
                    uint4 hi4 = make_uint4(0u, 0u, 0u, 0u), lo4 = hi4;
                    if (n0 < len) {
                        float x[9];
                        x[0] = n0 > 0 ? s16_bits_to_float(prev16[u]) : 0.f;          // y[0] = x[0]
                        x[1] = s16_bits_to_float(w[u].x & 0xffffu); x[2] = s16_bits_to_float(w[u].x >> 16);
                        x[3] = s16_bits_to_float(w[u].y & 0xffffu); x[4] = s16_bits_to_float(w[u].y >> 16);
                        x[5] = s16_bits_to_float(w[u].z & 0xffffu); x[6] = s16_bits_to_float(w[u].z >> 16);
                        x[7] = s16_bits_to_float(w[u].w & 0xffffu); x[8] = s16_bits_to_float(w[u].w >> 16);
                        float y[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float v = kS1 * fmaf(-pre, x[j], x[j + 1]);
                            y[j] = (n0 + j < len) ? v : 0.f;                        // zero padded tail
                        }
                        split2(y[0], y[1], hi4.x, lo4.x);
                        split2(y[2], y[3], hi4.y, lo4.y);
                        split2(y[4], y[5], hi4.z, lo4.z);
                        split2(y[6], y[7], hi4.w, lo4.w);
                    }
                    const int h = c8 & 1, j16 = (c8 >> 1) * 16;
                    *reinterpret_cast<uint4*>(slot + h * kPlaneH + j16) = hi4;
                    *reinterpret_cast<uint4*>(slot + kPlaneLo + h * kPlaneH + j16) = lo4;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (warp == 8) TC_STAMP(9 + 2 * g);
                if (lane == 0) {
                    mbar_arrive(&s.plane_full[g]);
                    mbar_arrive(&s.raw_empty[rs]);
                }
            }
        }
    } else if (warp >= 4) {
        // ================= convert warps: D1 -> twiddle -> fp16 hi/lo stage-2 operand =================
        const int q = warp - 4;
        const int fl = 4 * q + (lane >> 3), r = lane & 7;
        // s2 * W512^(r k1): the part of the twiddle W512^((8 h + r) k1) that depends on the lane; the W64^(h k1)
        // part is folded into the h = 1 copy of B1, so the 15 factors stay in registers for the whole kernel
        float2 tw[16];
        {
            const float2* twr = reinterpret_cast<const float2*>(p.consts + kConstSmemBytes);
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) tw[k1] = __ldg(&twr[k1 * 8 + r]);
        }
        for (int it = 0; it < my_tiles; ++it) {
            const long long tile = blockIdx.x + static_cast<long long>(it) * gridDim.x;
            if (it >= 1) wait_or_trap(&s.s2_done, static_cast<uint32_t>((it - 1) & 1));    // A2 free again
            // blocks b = 2 g + h; the tcgen05.ld of block b+1 is in flight while block b is converted
            uint32_t va[32], vb[32];
            const uint32_t tlane = tmem + (static_cast<uint32_t>(32 * q) << 16);
            auto convert_block = [&](const uint32_t (&vr)[32], int g, int h) {
                const int frow = 16 * g + fl;                                 // frame of the tile
                float v[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(vr[c]);
                if constexpr (DBG) {
                    if (p.dbg) {
                        float* d = p.dbg + (tile * 2) * 128 * 256 + (32 * q + lane) * 256 + (2 * g + h) * 32;
#pragma unroll
                        for (int c = 0; c < 32; ++c) d[c] = v[c];
                    }
                }
                const int n2 = 8 * h + r;
                s.s16[it & 1][frow][n2] = v[1];
                const uint32_t koff = (n2 >> 2) * kA2Lbo + (n2 & 3) * 4 + (frow >> 3) * 128 + (frow & 7) * 16;
#pragma unroll
                for (int k1 = 0; k1 < 16; ++k1) {
                    float tr, ti;
                    if (k1 == 0) {
                        tr = v[0] * kS2;
                        ti = 0.f;
                    } else {
                        const float2 w = tw[k1];
                        const float a = v[2 * k1], b = v[2 * k1 + 1];
                        tr = fmaf(a, w.x, -b * w.y);
                        ti = fmaf(a, w.y, b * w.x);
                    }
                    uint32_t hi, lo;
                    split2(tr, ti, hi, lo);
                    const uint32_t off = koff + (k1 & 1) * 1024;               // rows 64.. of the pair: +8 row groups
                    *reinterpret_cast<uint32_t*>(&s.a2[k1 >> 1][0][off]) = hi;
                    *reinterpret_cast<uint32_t*>(&s.a2[k1 >> 1][1][off]) = lo;
                }
            };
            auto block_done = [&](int g) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                fence_proxy_async_smem();
                __syncwarp();
                if (warp == 4) TC_STAMP(17 + 2 * g);
                if (lane == 0) mbar_arrive(&s.d1_empty[g]);
            };
            wait_or_trap(&s.d1_full[0], static_cast<uint32_t>(it & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp == 4) TC_STAMP(16);
            tmem_ld32_issue(tlane, va);
#pragma unroll 1
            for (int g = 0; g < kTileGroups; ++g) {
                tmem_wait32(va);                                              // block (g, 0)
                tmem_ld32_issue(tlane + (2 * g + 1) * 32, vb);
                convert_block(va, g, 0);
                tmem_wait32(vb);                                              // block (g, 1)
                if (g + 1 < kTileGroups) {
                    wait_or_trap(&s.d1_full[g + 1], static_cast<uint32_t>(it & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (warp == 4) TC_STAMP(16 + 2 * (g + 1));
                    tmem_ld32_issue(tlane + (2 * g + 2) * 32, va);
                }
                convert_block(vb, g, 1);
                block_done(g);
            }
        }
    } else {
        // ================= epilogue warps =================
        const int q = warp;
        for (int it = 0; it < my_tiles; ++it) {
            const long long tile = blockIdx.x + static_cast<long long>(it) * gridDim.x;
            // s2_done(it) also covers the convert warps' s16 stores: stage 2 was issued after d1_empty[3](it)
            wait_or_trap(&s.s2_done, static_cast<uint32_t>(it & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp == 0) TC_STAMP(24);
            if constexpr (DBG) {
                if (p.dbg) {
                    float* d = p.dbg + (tile * 2 + 1) * 128 * 256 + (32 * q + lane) * 256;
                    for (int j = 0; j < 8; ++j) {
                        uint32_t vr[32];
                        tmem_ld32_issue(tmem + (static_cast<uint32_t>(32 * q) << 16) + 256u + 32 * j, vr);
                        tmem_wait32(vr);
#pragma unroll
                        for (int c = 0; c < 32; ++c) d[32 * j + c] = __uint_as_float(vr[c]);
                    }
                }
            }
            if (q < 2) epilogue_tile<NF, 0>(p, s, tmem, q, lane, it, tile);
            else epilogue_tile<NF, 1>(p, s, tmem, q, lane, it, tile);
            if (warp == 0) TC_STAMP(26);
            if (warp == 3) TC_STAMP(27);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

#include "mfcc_tc2.inc"

// ---------------------------------------------------------------------------------------------
// Finish pass: delta, delta-delta (reference `delta(feat, 2)` applied twice, edge replicated) and
// zero rows up to pad_frames.  One CTA handles 128 rows of one clip.
// ---------------------------------------------------------------------------------------------
constexpr int kFinRows = 128;
struct FinParams {
    float* out;
    const int32_t* clip_len_arr;
    long long out_clip_stride;
    int clip_len, row_stride, pad_frames, with_deltas, tiles_per_clip;
};
__global__ void __launch_bounds__(256) mfcc_finish_kernel(const FinParams p) {
    __shared__ float c[kFinRows + 8][13];
    __shared__ float d[kFinRows + 4][13];
    const long long clip = blockIdx.x / p.tiles_per_clip;
    const int t0 = static_cast<int>(blockIdx.x - clip * p.tiles_per_clip) * kFinRows;
    const int len = p.clip_len_arr ? p.clip_len_arr[clip] : p.clip_len;
    const int T = len <= kFrameLen ? 1 : 1 + (len - kFrameLen + kStep - 1) / kStep;
    const int n_real = p.pad_frames > 0 ? min(T, p.pad_frames) : T;
    const int rows_total = p.pad_frames > 0 ? p.pad_frames : T;
    float* o = p.out + clip * p.out_clip_stride;
    const int tid = threadIdx.x;
    if (p.with_deltas && t0 < n_real) {
        for (int i = tid; i < (kFinRows + 8) * 13; i += 256) {
            const int rr = i / 13, cc = i - rr * 13;
            int t = t0 - 4 + rr;
            t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
            c[rr][cc] = o[static_cast<long long>(t) * p.row_stride + cc];
        }
        __syncthreads();
        for (int i = tid; i < (kFinRows + 4) * 13; i += 256) {
            const int rr = i / 13, cc = i - rr * 13;
            int sidx = t0 - 2 + rr;                                   // delta row, clamped like the edge padding
            sidx = sidx < 0 ? 0 : (sidx > T - 1 ? T - 1 : sidx);
            float acc = 0.f;
#pragma unroll
            for (int k = -2; k <= 2; ++k) {
                int tt = sidx + k;
                tt = tt < 0 ? 0 : (tt > T - 1 ? T - 1 : tt);
                acc = fmaf(static_cast<float>(k), c[tt - (t0 - 4)][cc], acc);
            }
            d[rr][cc] = acc * 0.1f;
        }
        __syncthreads();
        for (int i = tid; i < kFinRows * 26; i += 256) {
            const int rr = i / 26, cc = i - rr * 26;
            const int t = t0 + rr;
            if (t >= n_real) continue;
            float v;
            if (cc < 13) {
                v = d[rr + 2][cc];
            } else {
                float acc = 0.f;
#pragma unroll
                for (int k = -2; k <= 2; ++k) acc = fmaf(static_cast<float>(k), d[rr + 2 + k][cc - 13], acc);
                v = acc * 0.1f;
            }
            o[static_cast<long long>(t) * p.row_stride + 13 + cc] = v;
        }
    }
    // zero the extra columns [dim, row_stride) of this tile's real rows
    const int dim = p.with_deltas ? 39 : 13;
    if (p.row_stride > dim) {
        const int extra = p.row_stride - dim;
        const int r1 = min(t0 + kFinRows, n_real);
        for (int i = tid; i < (r1 - t0) * extra; i += 256) {
            const int rr = i / extra, cc = i - rr * extra;
            o[static_cast<long long>(t0 + rr) * p.row_stride + dim + cc] = 0.f;
        }
    }
    // zero rows [n_real, rows_total) of this tile
    const int z0 = max(t0, n_real), z1 = min(t0 + kFinRows, rows_total);
    if (z1 > z0) {
        float* zp = o + static_cast<long long>(z0) * p.row_stride;
        const int n = (z1 - z0) * p.row_stride;
        for (int i = tid; i < n; i += 256) zp[i] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
std::mutex g_mu;
std::map<std::pair<int, int>, unsigned char*> g_consts;      // per (device, nfilt)

void put_split(unsigned char* hi, unsigned char* lo, int n, int k, double v) {
    // B[n][k] in the UMMA K-major no-swizzle layout: core matrix (n/8, k/8) = 8 rows x 16 B
    const size_t off = static_cast<size_t>(n / 8) * 512 + static_cast<size_t>(k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
    const __half h = __float2half_rn(static_cast<float>(v));
    const __half l = __float2half_rn(static_cast<float>(v - static_cast<double>(__half2float(h))));
    memcpy(hi + off, &h, 2);
    memcpy(lo + off, &l, 2);
}

int get_consts(int nfilt, const unsigned char** out) {
    int dev = 0;
    MMLA_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_consts.find(std::make_pair(dev, nfilt));
    if (it != g_consts.end()) {
        *out = it->second;
        return MMLA_OK;
    }
    std::vector<unsigned char> host(kConstBytes, 0);
    unsigned char* b1 = host.data();                 // [h][hi|lo][2048]
    unsigned char* b2h = b1 + 4 * 2048;
    unsigned char* b2l = b2h + 2048;
    float2* k16 = reinterpret_cast<float2*>(b2l + 2048);
    float* dct = reinterpret_cast<float*>(k16 + 8 * 16);
    float2* twr = reinterpret_cast<float2*>(dct + 40 * 16);
    const double PI = 3.14159265358979323846;
    // stage 1: column 0 = Re S[.][0], column 1 = Re S[.][16], columns 2j, 2j+1 = Re, Im of S[.][j] W64^(h j); rows
    // n1 >= 25 are the frame's zero padding
    for (int h = 0; h < 2; ++h) {
        unsigned char* bh = b1 + h * 4096;
        unsigned char* bl = bh + 2048;
        for (int n1 = 0; n1 < 32; ++n1) {
            const double live = n1 < 25 ? 1.0 : 0.0;
            put_split(bh, bl, 0, n1, live);
            put_split(bh, bl, 1, n1, live * ((n1 & 1) ? -1.0 : 1.0));
            for (int j = 1; j < 16; ++j) {
                const double th = 2.0 * PI * (((2 * n1 + h) * j) % 64) / 64.0;
                put_split(bh, bl, 2 * j, n1, live * cos(th));
                put_split(bh, bl, 2 * j + 1, n1, -live * sin(th));
            }
        }
    }
    // stage 2: complex DFT-16 as a real 32 x 32 product, K = (n2, re/im), N = (k2, re/im)
    for (int n2 = 0; n2 < 16; ++n2)
        for (int k2 = 0; k2 < 16; ++k2) {
            const double th = 2.0 * PI * ((n2 * k2) % 16) / 16.0;
            const int cre = 4 * (k2 >> 1) + (k2 & 1), cim = cre + 2;     // columns: re 2i, re 2i+1, im 2i, im 2i+1
            put_split(b2h, b2l, cre, 2 * n2, cos(th));
            put_split(b2h, b2l, cre, 2 * n2 + 1, sin(th));
            put_split(b2h, b2l, cim, 2 * n2, -sin(th));
            put_split(b2h, b2l, cim, 2 * n2 + 1, cos(th));
        }
    for (int k1 = 0; k1 < 16; ++k1)
        for (int r = 0; r < 8; ++r) {
            const double th = 2.0 * PI * ((r * k1) % 512) / 512.0;
            twr[k1 * 8 + r] = make_float2(static_cast<float>(kS2 * cos(th)), static_cast<float>(-kS2 * sin(th)));
        }
    for (int k2 = 0; k2 < 8; ++k2)
        for (int n2 = 0; n2 < 16; ++n2) {
            const double th = 2.0 * PI * ((n2 * (16 + 32 * k2)) % 512) / 512.0;
            k16[k2 * 16 + n2] = make_float2(static_cast<float>(kS2 * cos(th)), static_cast<float>(-kS2 * sin(th)));
        }
    // DCT-II ortho (scipy.fftpack.dct norm='ortho') with the psf lifter (L = 22) folded in; rows >= nfilt stay zero
    for (int c = 0; c < 13; ++c) {
        const double scale = c == 0 ? sqrt(1.0 / nfilt) : sqrt(2.0 / nfilt);
        const double lift = 1.0 + (22 / 2.0) * sin(PI * c / 22);
        for (int m = 0; m < nfilt; ++m)
            dct[m * 16 + (c < 7 ? c : c + 1)] = static_cast<float>(lift * scale * cos(PI * c * (2 * m + 1) / (2.0 * nfilt)));
    }
    unsigned char* devp = nullptr;
    cudaError_t e = cudaMalloc(&devp, kConstBytes);
    if (e == cudaSuccess) e = cudaMemcpy(devp, host.data(), kConstBytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        mmla_set_error("mfcc_tc constants upload failed: %s", cudaGetErrorString(e));
        return MMLA_ECUDA;
    }
    g_consts[std::make_pair(dev, nfilt)] = devp;
    *out = devp;
    return MMLA_OK;
}

template <int NF, bool DBG>
int launch_tc(const TcParams& kp, long long grid, cudaStream_t st) {
    static MmlaPerDeviceOnce attr_once;                          // cudaFuncSetAttribute is per device
    const bool attr_set = !attr_once.first();
    const int smem = static_cast<int>(sizeof(TcSmem));
    if (!attr_set) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(mfcc_tc_kernel<NF, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    mfcc_tc_kernel<NF, DBG><<<static_cast<unsigned>(grid), kThreads, smem, st>>>(kp);
    mmla_count_launch("mfcc_tc_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}

std::map<std::pair<int, int>, unsigned char*> g_consts2;     // per (device, nfilt), mfcc_tc2_kernel

void put_split2(unsigned char* hi, unsigned char* lo, int n, int k, double v) {
    // B[n][k], 64 x 64, UMMA K-major no-swizzle: core matrix (n/8, k/8) = 8 rows x 16 B, N-group stride 1024 B
    const size_t off = static_cast<size_t>(n / 8) * 1024 + static_cast<size_t>(k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
    const __half h = __float2half_rn(static_cast<float>(v));
    const __half l = __float2half_rn(static_cast<float>(v - static_cast<double>(__half2float(h))));
    memcpy(hi + off, &h, 2);
    memcpy(lo + off, &l, 2);
}

int get_consts2(int nfilt, const unsigned char** out) {
    int dev = 0;
    MMLA_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_consts2.find(std::make_pair(dev, nfilt));
    if (it != g_consts2.end()) {
        *out = it->second;
        return MMLA_OK;
    }
    std::vector<unsigned char> host(k2ConstBytes, 0);
    unsigned char* bh = host.data();
    unsigned char* bl = bh + k2B1Bytes;
    float2* tw = reinterpret_cast<float2*>(bl + k2B1Bytes);
    float* dct = reinterpret_cast<float*>(tw + 8 * k2TwStride);
    const double PI = 3.14159265358979323846;
    // stage 1: column 0 = Re S[.][0], column 1 = Re S[.][32], columns 2j, 2j+1 = Re, Im of S[.][j]; rows n1 >= 50 are the
    // frame's zero padding (8 * 50 = 400 samples)
    for (int n1 = 0; n1 < 64; ++n1) {
        const double live = n1 < 50 ? 1.0 : 0.0;
        put_split2(bh, bl, 0, n1, live);
        put_split2(bh, bl, 1, n1, live * ((n1 & 1) ? -1.0 : 1.0));
        for (int j = 1; j < 32; ++j) {
            const double th = 2.0 * PI * ((n1 * j) % 64) / 64.0;
            put_split2(bh, bl, 2 * j, n1, live * cos(th));
            put_split2(bh, bl, 2 * j + 1, n1, -live * sin(th));
        }
    }
    for (int r = 0; r < 8; ++r)
        for (int k1 = 0; k1 <= 32; ++k1) {
            const double th = 2.0 * PI * ((r * k1) % 512) / 512.0;
            tw[r * k2TwStride + k1] = make_float2(static_cast<float>(kS2 * cos(th)), static_cast<float>(-kS2 * sin(th)));
        }
    for (int c = 0; c < 13; ++c) {
        const double scale = c == 0 ? sqrt(1.0 / nfilt) : sqrt(2.0 / nfilt);
        const double lift = 1.0 + (22 / 2.0) * sin(PI * c / 22);
        for (int m = 0; m < nfilt; ++m)
            dct[m * 16 + c] = static_cast<float>(lift * scale * cos(PI * c * (2 * m + 1) / (2.0 * nfilt)));
    }
    unsigned char* devp = nullptr;
    cudaError_t e = cudaMalloc(&devp, k2ConstBytes);
    if (e == cudaSuccess) e = cudaMemcpy(devp, host.data(), k2ConstBytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        mmla_set_error("mfcc_tc2 constants upload failed: %s", cudaGetErrorString(e));
        return MMLA_ECUDA;
    }
    g_consts2[std::make_pair(dev, nfilt)] = devp;
    *out = devp;
    return MMLA_OK;
}

template <int NF>
int launch_tc2(const TcParams& kp, long long grid, cudaStream_t st) {
    static MmlaPerDeviceOnce attr_once;
    const int smem = static_cast<int>(sizeof(Tc2Smem));
    if (attr_once.first()) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(mfcc_tc2_kernel<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    mfcc_tc2_kernel<NF><<<static_cast<unsigned>(grid), k2Threads, smem, st>>>(kp);
    mmla_count_launch("mfcc_tc2_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}

}  // namespace

// Returns MMLA_OK and *handled = 1 when the tensor-core path ran; *handled = 0 when the
// parameters are outside what it is specialised for (the caller then uses the general kernel).
// `dbg` (device, or null): per tile [D1 | D2] raw accumulators, 2 x 128 x 256 floats.
// `prof` (device, or null): clock64 stamps of CTA 0, 64 tiles x 32 slots.
int mmla_mfcc_tc_try(const int16_t* pcm, int64_t pcm_total, const int64_t* clip_off_host, const int32_t* clip_len_host,
                     int64_t n_clips, int32_t clip_len, int64_t clip_stride, const MmlaMfccParams& p, float* out,
                     int64_t out_clip_stride, int32_t out_row_stride, cudaStream_t st, float* dbg, long long* prof,
                     int* handled) {
    *handled = 0;
    const char* force = getenv("MMLA_MFCC_KERNEL");
    if (force && strcmp(force, "fft") == 0) return MMLA_OK;
    if (p.samplerate != 16000 || p.nfft != 512 || p.frame_len != kFrameLen || p.frame_step != kStep) return MMLA_OK;
    if (p.window != MMLA_WINDOW_RECT || p.numcep != 13 || p.ceplifter != 22) return MMLA_OK;
    if (p.nfilt != 26 && p.nfilt != 40) return MMLA_OK;
    if (p.lowfreq != 0.f || (p.highfreq != 8000.f && p.highfreq != 0.f)) return MMLA_OK;
    auto frames_of = [](long long len) { return len <= kFrameLen ? 1LL : 1 + (len - kFrameLen + kStep - 1) / kStep; };
    std::vector<int2> groups;
    long long n_groups = 0;
    int groups_per_clip = 0;
    long long fin_tiles_per_clip = 1;
    if (clip_off_host == nullptr) {
        if ((clip_stride & 7) != 0 && n_clips > 1) return MMLA_OK;            // clip starts must be 16-byte aligned
        if (clip_len < 1) return MMLA_OK;
        const long long T = frames_of(clip_len);
        if (p.with_deltas && p.pad_frames > 0 && T > p.pad_frames) return MMLA_OK;   // truncated context: general kernel
        const long long n_real = p.pad_frames > 0 ? std::min<long long>(T, p.pad_frames) : T;
        groups_per_clip = static_cast<int>((n_real + kGroupFrames - 1) / kGroupFrames);
        n_groups = n_clips * groups_per_clip;
        const long long rows_total = p.pad_frames > 0 ? p.pad_frames : T;
        fin_tiles_per_clip = (rows_total + kFinRows - 1) / kFinRows;
    } else {
        long long max_rows = 1;
        for (int64_t c = 0; c < n_clips; ++c) {
            if ((clip_off_host[c] & 7) != 0 || clip_len_host[c] < 1) return MMLA_OK;
            const long long T = frames_of(clip_len_host[c]);
            if (p.with_deltas && p.pad_frames > 0 && T > p.pad_frames) return MMLA_OK;
            const long long n_real = p.pad_frames > 0 ? std::min<long long>(T, p.pad_frames) : T;
            for (long long f0 = 0; f0 < n_real; f0 += kGroupFrames) groups.push_back(make_int2(static_cast<int>(c), static_cast<int>(f0)));
            max_rows = std::max(max_rows, p.pad_frames > 0 ? static_cast<long long>(p.pad_frames) : T);
        }
        n_groups = static_cast<long long>(groups.size());
        fin_tiles_per_clip = (max_rows + kFinRows - 1) / kFinRows;
    }
    if (n_groups == 0) return MMLA_OK;

    // MMLA_MFCC_TC=2 selects the 64 x 8 formulation (mfcc_tc2_kernel); the debug dumps / clock stamps exist only in the
    // 32 x 16 kernel
    const char* form = getenv("MMLA_MFCC_TC");
    const bool use_tc2 = form && strcmp(form, "2") == 0 && !dbg;
    const unsigned char* consts = nullptr;
    int rc = use_tc2 ? get_consts2(p.nfilt, &consts) : get_consts(p.nfilt, &consts);
    if (rc != MMLA_OK) return rc;

    TcParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.pcm = pcm;
    kp.consts = consts;
    kp.out = out;
    kp.dbg = dbg;
    kp.prof = prof;
    kp.n_groups = n_groups;
    kp.clip_stride = clip_stride;
    kp.out_clip_stride = out_clip_stride;
    kp.groups_per_clip = groups_per_clip;
    kp.clip_len = clip_len;
    kp.row_stride = out_row_stride;
    kp.zero_tail = (!p.with_deltas && p.pad_frames == 0 && out_row_stride > 13 && out_row_stride - 13 <= 3) ? out_row_stride - 13 : 0;
    kp.pad_frames = p.pad_frames;
    kp.append_energy = p.append_energy;
    kp.preemph = p.preemph;

    void* dev_tmp = nullptr;
    const int64_t* d_off = nullptr;
    const int32_t* d_len = nullptr;
    if (clip_off_host != nullptr) {
        const size_t b_g = groups.size() * sizeof(int2);
        const size_t b_off = static_cast<size_t>(n_clips) * sizeof(int64_t);
        const size_t b_len = static_cast<size_t>(n_clips) * sizeof(int32_t);
        const size_t o_off = (b_g + 15) & ~static_cast<size_t>(15);
        const size_t o_len = o_off + ((b_off + 15) & ~static_cast<size_t>(15));
        MMLA_CUDA_CHECK(cudaMallocAsync(&dev_tmp, o_len + b_len, st));
        char* base = static_cast<char*>(dev_tmp);
        MMLA_CUDA_CHECK(cudaMemcpyAsync(base, groups.data(), b_g, cudaMemcpyHostToDevice, st));
        MMLA_CUDA_CHECK(cudaMemcpyAsync(base + o_off, clip_off_host, b_off, cudaMemcpyHostToDevice, st));
        MMLA_CUDA_CHECK(cudaMemcpyAsync(base + o_len, clip_len_host, b_len, cudaMemcpyHostToDevice, st));
        kp.groups = reinterpret_cast<const int2*>(base);
        kp.clip_off = d_off = reinterpret_cast<const int64_t*>(base + o_off);
        kp.clip_len_arr = d_len = reinterpret_cast<const int32_t*>(base + o_len);
    }
    (void)d_off;
    (void)pcm_total;

    const int sms = mmla_num_sms();
    MMLA_REQUIRE(sms > 0, MMLA_ECUDA, "mfcc_tc: no CUDA device");
    const long long n_tiles = (n_groups + kTileGroups - 1) / kTileGroups;
    const long long grid = n_tiles < sms ? n_tiles : sms;
    if (use_tc2) rc = p.nfilt == 26 ? launch_tc2<26>(kp, grid, st) : launch_tc2<40>(kp, grid, st);
    else if (dbg || prof) rc = p.nfilt == 26 ? launch_tc<26, true>(kp, grid, st) : launch_tc<40, true>(kp, grid, st);
    else rc = p.nfilt == 26 ? launch_tc<26, false>(kp, grid, st) : launch_tc<40, false>(kp, grid, st);
    if (rc != MMLA_OK) return rc;

    const int dim = p.with_deltas ? 39 : 13;
    // plain cepstra with a few spare columns per row (e.g. the 16-float rows stem_fused.cu consumes): the epilogue zeroes
    // them itself and no second pass is needed
    const bool tail_in_epilogue = !p.with_deltas && p.pad_frames == 0 && out_row_stride > dim && out_row_stride - dim <= 3;
    const bool need_finish = p.with_deltas || p.pad_frames > 0 || (out_row_stride > dim && !tail_in_epilogue);
    if (need_finish) {
        FinParams fp;
        fp.out = out;
        fp.clip_len_arr = d_len;
        fp.out_clip_stride = out_clip_stride;
        fp.clip_len = clip_len;
        fp.row_stride = kp.row_stride;
        fp.pad_frames = p.pad_frames;
        fp.with_deltas = p.with_deltas;
        fp.tiles_per_clip = static_cast<int>(fin_tiles_per_clip);
        const long long fgrid = n_clips * fin_tiles_per_clip;
        MMLA_REQUIRE(fgrid < (1LL << 31), MMLA_EUNSUP, "mfcc_tc: too many finish tiles");
        mfcc_finish_kernel<<<static_cast<unsigned>(fgrid), 256, 0, st>>>(fp);
        mmla_count_launch("mfcc_finish_kernel", st);
        MMLA_CUDA_CHECK(cudaGetLastError());
    }
    if (dev_tmp) MMLA_CUDA_CHECK(cudaFreeAsync(dev_tmp, st));
    *handled = 1;
    return MMLA_OK;
}
