// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a, TF32 operands, FP32
// accumulation in tensor memory (TMEM).
//
//   Y[M, N] = gather(X)[M, K] * W[K, N] + bias (+ residual)          M = B*Ho*Wo, K = kh*kw*Cin
//
// One CTA computes a 128 x NT output tile:
//   * A operand (activations): all 256 threads gather the im2col tile from NHWC global memory,
//     apply the preceding BatchNorm + ELU/ReLU on the fly, round to TF32 and store it to shared
//     memory in the UMMA canonical K-major SWIZZLE_128B layout (128-byte rows, 16-byte chunks
//     XORed with row%8).  A warp load covers 4 rows x 128 contiguous bytes (coalesced) and each
//     quarter-warp STS.128 covers the eight swizzled chunks of one row (conflict free).
//   * B operand (weights): pre-arranged on the host in the same slab layout and pre-rounded to
//     TF32, so one TMA bulk copy (cp.async.bulk + mbarrier complete_tx) lands a whole
//     [32 x NT] K-chunk, ready for the tensor core.
//   * One elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=NT, K=8) four
//     times per 32-wide K chunk; tcgen05.commit releases the smem stage (2-stage ring) and, after
//     the last chunk, signals the epilogue.
//   * Epilogue: tcgen05.ld TMEM -> registers, + bias, staged through the (now dead) operand
//     buffers so the final stores (and residual loads) are fully coalesced 128-bit accesses.
// Used for every conv / LSTM projection of both classifiers when the net runs in TF32 mode
// (mmla_net_set_precision); the fp32 CUDA-core kernel in nets.cu is the bit-faithful path.
#include <math.h>
#include <string.h>

#include "conv_common.cuh"

namespace {

constexpr int kTcBK = 32;                  // K elements per pipeline stage (4 MMAs of K=8)
constexpr int kTcStages = 2;
constexpr int kAStageBytes = 8 * 128 * 16;  // 8 slabs x 128 rows x 16 B

// Round-to-nearest (ties away) to TF32's 10-bit mantissa with two integer ops — what
// cvt.rna.tf32.f32 does for finite inputs, without its multi-instruction special-case handling
// (activations here are finite; an Inf/NaN would propagate as a huge/NaN value either way).
__device__ __forceinline__ uint32_t f32_to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
// mbarrier wait that traps instead of hanging the GPU if a barrier is never satisfied (a wrong
// descriptor / byte count would otherwise spin forever); ~seconds of polling before giving up.
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    asm volatile("trap;");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE, Blackwell version bit set.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;                       // descriptor version (sm_100)
    return d;                              // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

template <int NT>
struct TcSmem {
    // A stages: 128 rows x 128 B in the UMMA K-major SWIZZLE_128B layout (8-row x 128 B atoms, the
    // 16-byte chunk index XORed with row%8).  B stages: no-swizzle slab layout as arranged on the
    // host.  After the last MMA the A|B region is reused as the epilogue staging tile.
    alignas(1024) unsigned char A[kTcStages][kAStageBytes];
    alignas(128) unsigned char B[kTcStages][8 * NT * 16];
    alignas(8) uint64_t full_b[kTcStages];
    alignas(8) uint64_t empty[kTcStages];
    alignas(8) uint64_t accum;
    uint32_t tmem_base;
};

// UMMA descriptor for the SWIZZLE_128B K-major A tile (LBO unused = 1, SBO = 1024 B between 8-row atoms).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>(1u) << 16;
    d |= static_cast<uint64_t>((1024u >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;                       // descriptor version (sm_100)
    d |= 2ull << 61;                       // layout type: SWIZZLE_128B
    return d;
}

template <int NT>
__global__ void __launch_bounds__(256) conv_tc_kernel(const ConvArgs a, const float* __restrict__ wg) {
    constexpr int kCols = NT < 32 ? 32 : NT;                       // TMEM columns (power of two >= 32)
    constexpr uint32_t kBStage = 8 * NT * 16;
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(NT >> 3) << 17) |
                                (static_cast<uint32_t>(128 >> 4) << 24);   // D=f32, A=B=tf32, K-major, N, M=128
    extern __shared__ unsigned char smem_dyn[];
    // SWIZZLE_128B atoms need a 1024-byte aligned base; the launch reserves 1 KB of slack for this
    // (offset added to the __shared__ array itself, not to an integer copy of its address, so the compiler keeps
    // every access in the shared address space: LDS/STS instead of generic LD/ST)
    TcSmem<NT>& s = *reinterpret_cast<TcSmem<NT>*>(smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m0 = static_cast<long long>(blockIdx.x) * 128;
    const int ntile = blockIdx.y;
    const int n0 = ntile * NT;
    const int nk = (a.K + kTcBK - 1) / kTcBK;

    if (tid == 0) {
        for (int i = 0; i < kTcStages; ++i) {
            mbar_init(&s.full_b[i], 1);
            mbar_init(&s.empty[i], 1);
        }
        mbar_init(&s.accum, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)),
                     "r"(static_cast<uint32_t>(kCols))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;

    // ---- A-gather bookkeeping -------------------------------------------------------------------
    // Thread t owns the 16-byte chunk q = t&7 (4 consecutive K elements) of rows rb+32i, i = 0..3,
    // so one warp load instruction covers 4 rows x 128 contiguous bytes (coalesced) and one warp
    // STS.128 covers, per quarter-warp, the eight swizzled chunks of one row (conflict free).
    const int q = tid & 7, rb = tid >> 3;
    // (the launcher guarantees M and the input element count fit in 31 bits: 32-bit index math)
    int hi0[4], wi0[4], xb[4], klo[4], khi[4];
    bool rv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int am = static_cast<int>(m0) + rb + 32 * i;
        rv[i] = am < static_cast<int>(a.M);
        hi0[i] = wi0[i] = xb[i] = klo[i] = khi[i] = 0;
        if (rv[i]) {
            const int hw = a.Ho * a.Wo;
            const int b = am / hw;
            const int r = am - b * hw;
            const int ho = r / a.Wo, wo = r - ho * a.Wo;
            hi0[i] = ho * a.stride - a.pad_t;
            wi0[i] = wo * a.stride - a.pad_l;
            xb[i] = b * a.H * a.W * a.Cin;
            // kh == 1, stride 1 convs: the im2col row is the contiguous input run
            // [xb + (hi0*W + wi0)*Cin + k], valid for k in [klo, khi) (columns inside the image)
            klo[i] = max(0, -wi0[i]) * a.Cin;
            khi[i] = min(min(a.kw, a.W - wi0[i]) * a.Cin, a.K);
        }
    }
    const bool fast = (a.Cin % 4 == 0) && !a.x_is_u8;             // a K-quad never straddles a filter tap
    const bool rowrun = !fast && a.kh == 1 && a.stride == 1 && !a.pre_scale;   // stems: contiguous im2col rows
    const float* xf = static_cast<const float*>(a.x);
    const unsigned char* xu = static_cast<const unsigned char*>(a.x);

    // load_raw() only issues the global loads of a chunk; finish() (BN + activation + TF32
    // rounding) runs one iteration later, so the loads fly across the barrier + MMA issue.
    auto load_raw = [&](int kc, float4 (&raw)[4], int& chan) {
        const int k = kc * kTcBK + 4 * q;
        chan = -1;
#pragma unroll
        for (int i = 0; i < 4; ++i) raw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fast) {
            if (k < a.K) {
                const int tap = k / a.Cin, c = k - tap * a.Cin;
                const int ki = tap / a.kw, kj = tap - ki * a.kw;
                chan = c;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int hi = hi0[i] + ki, wi = wi0[i] + kj;
                    if (rv[i] && hi >= 0 && hi < a.H && wi >= 0 && wi < a.W)
                        raw[i] = *reinterpret_cast<const float4*>(xf + (xb[i] + (hi * a.W + wi) * a.Cin + c));
                    else
                        raw[i].x = __int_as_float(0x7fc00000);   // NaN marks "padding": stays zero after BN
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float e[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int kk = k + j;
                    float v = 0.f;
                    if (rowrun) {
                        if (kk >= klo[i] && kk < khi[i]) {
                            const int idx = xb[i] + (hi0[i] * a.W + wi0[i]) * a.Cin + kk;
                            v = a.x_is_u8 ? static_cast<float>(xu[idx]) : xf[idx];
                        }
                    } else if (rv[i] && kk < a.K) {
                        {
                            const int tap = kk / a.Cin, c = kk - tap * a.Cin;
                            const int ki = tap / a.kw, kj = tap - ki * a.kw;
                            const int hi = hi0[i] + ki, wi = wi0[i] + kj;
                            if (hi >= 0 && hi < a.H && wi >= 0 && wi < a.W) {
                                const int idx = xb[i] + (hi * a.W + wi) * a.Cin + c;
                                v = a.x_is_u8 ? static_cast<float>(xu[idx]) : xf[idx];
                                if (a.pre_scale) v = apply_act_tc(fmaf(v, a.pre_scale[c], a.pre_shift[c]), a.pre_act);
                            }
                        }
                    }
                    e[j] = v;
                }
                raw[i] = make_float4(e[0], e[1], e[2], e[3]);
            }
        }
    };
    auto finish_store = [&](const float4 (&raw)[4], int chan, unsigned char* abase) {
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool bn = chan >= 0 && a.pre_scale;
        if (bn) {
            sc = *reinterpret_cast<const float4*>(a.pre_scale + chan);
            sh = *reinterpret_cast<const float4*>(a.pre_shift + chan);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 v = raw[i];
            if (chan >= 0) {
                if (__float_as_uint(v.x) == 0x7fc00000u) {
                    v = make_float4(0.f, 0.f, 0.f, 0.f);                     // padding pixel
                } else if (bn) {
                    v.x = apply_act_tc(fmaf(v.x, sc.x, sh.x), a.pre_act);
                    v.y = apply_act_tc(fmaf(v.y, sc.y, sh.y), a.pre_act);
                    v.z = apply_act_tc(fmaf(v.z, sc.z, sh.z), a.pre_act);
                    v.w = apply_act_tc(fmaf(v.w, sc.w, sh.w), a.pre_act);
                }
            }
            const int r = rb + 32 * i;
            *reinterpret_cast<uint4*>(abase + r * 128 + ((q ^ (r & 7)) << 4)) =
                make_uint4(f32_to_tf32(v.x), f32_to_tf32(v.y), f32_to_tf32(v.z), f32_to_tf32(v.w));
        }
    };

    float4 raw[4];
    int chan;
    load_raw(0, raw, chan);
    for (int kc = 0; kc < nk; ++kc) {
        const int st = kc & 1;
        const int use = kc >> 1;
        if (kc >= kTcStages) mbar_wait_or_trap(&s.empty[st], static_cast<uint32_t>((use - 1) & 1));   // MMAs of kc-2 done
        if (tid == 0) {   // (no proxy fence: B stages are only ever touched by the async proxy)
            mbar_arrive_expect_tx(&s.full_b[st], kBStage);
            tma_bulk_g2s(&s.B[st][0], wg + (static_cast<long long>(ntile) * nk + kc) * (NT * kTcBK), kBStage,
                         &s.full_b[st]);
        }
        if (warp == 0) __syncwarp();
        finish_store(raw, chan, &s.A[st][0]);
        fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the tensor core
        if (kc + 1 < nk) load_raw(kc + 1, raw, chan);   // next chunk's loads fly across the barrier + MMAs
        __syncthreads();
        if (warp == 0) {
            // the whole warp stays converged and ONE elected lane issues: warp-uniform descriptors live in uniform
            // registers and the four UTCHMMA go out back to back (from a `tid == 0` branch each cost an R2UR
            // waterfall of 80-200 cycles; scripts/microbench/umma_rate.cu)
            mbar_wait_or_trap(&s.full_b[st], static_cast<uint32_t>(use & 1));
            tc_fence_after();
            const uint64_t ad0 = umma_desc_sw128(smem_u32(&s.A[st][0]));
            const uint64_t bd0 = umma_desc(smem_u32(&s.B[st][0]), NT * 16, 128);
            if (tc_elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)                // K advances 32 B inside the atom row / two slabs of B
                    umma_tf32(tmem, ad0 + static_cast<uint64_t>(kk * 2), bd0 + static_cast<uint64_t>(kk * 2 * NT), kIdesc,
                              (kc | kk) != 0 ? 1u : 0u);
                umma_commit(&s.empty[st]);                   // frees this smem stage when the MMAs retire
                if (kc == nk - 1) umma_commit(&s.accum);     // accumulator complete
            }
            __syncwarp();
        }
    }

    // ---- epilogue: TMEM -> registers (+bias) -> smem staging -> coalesced (+residual) stores ----
    // The operand stages are dead once the accumulator barrier fires, so they become a
    // [128][HC+4] float staging tile (HC = columns per pass; +4 floats of padding keep the
    // per-row STS.128 conflict free).  The write-out then moves whole 16-byte words with
    // consecutive lanes on consecutive addresses; the residual is read the same way.
    {
        constexpr int HC = NT > 64 ? 64 : NT;                     // columns staged per pass
        constexpr int kPasses = NT / HC;
        constexpr int kStride = HC + 4;                           // floats
        static_assert(kStride * 128 * 4 <= kTcStages * (kAStageBytes + 8 * NT * 16), "staging tile must fit");
        float* stg = reinterpret_cast<float*>(&s.A[0][0]);
        const int quarter = warp & 3;
        const int chalf = warp >> 2;                              // warps 4..7 take the upper half of a pass
        const int row = quarter * 32 + lane;
        mbar_wait_or_trap(&s.accum, 0u);
        tc_fence_after();
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) {
            constexpr int kColsPerWarp = HC >= 32 ? HC / 2 : HC;  // HC=16: only warps 0..3 read TMEM
            const bool active = HC >= 32 || chalf == 0;
            if (active) {
#pragma unroll
                for (int c0 = 0; c0 < kColsPerWarp; c0 += 16) {
                    const int col = chalf * kColsPerWarp + c0;    // column within the pass
                    uint32_t r[16];
                    const uint32_t taddr = tmem + (static_cast<uint32_t>(quarter * 32) << 16) +
                                           static_cast<uint32_t>(pass * HC + col);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
                        "%13, %14, %15}, [%16];\n"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
                          "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const float* bias = a.bias + n0 + pass * HC + col;
                    float* dst = stg + row * kStride + col;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 bv = *reinterpret_cast<const float4*>(bias + j);
                        *reinterpret_cast<float4*>(dst + j) =
                            make_float4(__uint_as_float(r[j]) + bv.x, __uint_as_float(r[j + 1]) + bv.y,
                                        __uint_as_float(r[j + 2]) + bv.z, __uint_as_float(r[j + 3]) + bv.w);
                    }
                }
            }
            __syncthreads();
            constexpr int kQuadsPerRow = HC / 4;
            for (int idx = tid; idx < 128 * kQuadsPerRow; idx += 256) {
                const int r = idx / kQuadsPerRow, c4 = idx - r * kQuadsPerRow;
                const long long m = m0 + r;
                if (m < a.M) {   // (64-bit only for the final byte offsets)
                    float4 v = *reinterpret_cast<const float4*>(stg + r * kStride + 4 * c4);
                    const int n = n0 + pass * HC + 4 * c4;
                    if (a.res) {
                        const float4 rr = *reinterpret_cast<const float4*>(a.res + m * a.res_row_stride + n);
                        v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
                    }
                    *reinterpret_cast<float4*>(a.y + m * a.N + n) = v;
                }
            }
            if (pass + 1 < kPasses) __syncthreads();              // staging tile is rewritten by the next pass
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(static_cast<uint32_t>(kCols))
                     : "memory");
    }
}

template <int NT>
int launch_nt(const ConvArgs& a, const float* wg, cudaStream_t st) {
    static MmlaPerDeviceOnce attr_once;                          // cudaFuncSetAttribute is per device
    const bool attr_set = !attr_once.first();
    if (!attr_set) {
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(TcSmem<NT>) + 1024)));
    }
    const dim3 grid(static_cast<unsigned>((a.M + 127) / 128), static_cast<unsigned>(a.N / NT));
    conv_tc_kernel<NT><<<grid, 256, sizeof(TcSmem<NT>) + 1024, st>>>(a, wg);
    mmla_count_launch("conv_tc_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}

}  // namespace

// N tile used for an output width N (0: this layer is not eligible for the tensor-core path).
int mmla_tc_ntile(int n) {
    if (n == 16 || n == 32 || n == 64 || n == 128) return n;
    if (n > 128 && n % 128 == 0) return 128;
    return 0;
}

// Host: arrange W[K][N] (row-major, TF layout) for the TMA-fed B operand:
//   out[ntile][kchunk][slab 0..7][n 0..NT-1][4]  = W[kchunk*32 + slab*4 + j][ntile*NT + n], zero padded in K,
// rounded to TF32 (round-to-nearest, ties away — what cvt.rna.tf32.f32 does on the A side).
long long mmla_tc_arranged_floats(int K, int N) {
    const int nk = (K + kTcBK - 1) / kTcBK;
    return static_cast<long long>(nk) * kTcBK * N;
}
void mmla_tc_arrange_weights(const float* w, int K, int N, float* out) {
    const int NT = mmla_tc_ntile(N);
    const int nk = (K + kTcBK - 1) / kTcBK;
    const int ntiles = N / NT;
    for (int t = 0; t < ntiles; ++t)
        for (int kc = 0; kc < nk; ++kc)
            for (int slab = 0; slab < 8; ++slab)
                for (int n = 0; n < NT; ++n)
                    for (int j = 0; j < 4; ++j) {
                        const int k = kc * kTcBK + slab * 4 + j;
                        float v = k < K ? w[static_cast<long long>(k) * N + t * NT + n] : 0.f;
                        uint32_t u;
                        memcpy(&u, &v, 4);
                        if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & ~0x1FFFu;
                        memcpy(&v, &u, 4);
                        out[((((static_cast<long long>(t) * nk + kc) * 8 + slab) * NT + n) * 4) + j] = v;
                    }
}

// conv_slab.cu: stride-1 k > 1 convs without the im2col gather
bool mmla_conv_slab_eligible(const ConvArgs& a);
int mmla_launch_conv_slab(const ConvArgs& a, const float* wg, cudaStream_t st);

int mmla_launch_conv_tc(const ConvArgs& a, const float* wg, cudaStream_t st) {
    if (mmla_conv_slab_eligible(a)) return mmla_launch_conv_slab(a, wg, st);
    MMLA_REQUIRE(a.M < (1LL << 31) - 256 && (a.M / (a.Ho * a.Wo) + 1) * a.H * a.W * a.Cin < (1LL << 31), MMLA_EUNSUP,
                 "conv_tc: tensor too large for 32-bit indexing (reduce the micro-batch)");
    switch (mmla_tc_ntile(a.N)) {
        case 16: return launch_nt<16>(a, wg, st);
        case 32: return launch_nt<32>(a, wg, st);
        case 64: return launch_nt<64>(a, wg, st);
        case 128: return launch_nt<128>(a, wg, st);
        default:
            mmla_set_error("conv_tc: N=%d is not eligible for the tensor-core path", a.N);
            return MMLA_EUNSUP;
    }
}
