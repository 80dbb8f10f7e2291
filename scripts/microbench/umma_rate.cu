// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS mode) for several N and
// operand layouts, one CTA per SM, one issuing thread.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate scripts/microbench/umma_rate.cu && /tmp/umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (static_cast<uint64_t>(layout) << 61);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

struct Cfg { int N; int a_mode; int b_mode; int same_d; };   // a_mode: 0 K-major noswz, 1 K-major SW128, 2 MN-major noswz (stride 160)

__global__ void __launch_bounds__(128, 1) bench(const Cfg* cfgs, int ncfg, int iters, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x;
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    uint32_t phase = 0;
    if (tid == 0) {
        for (int c = 0; c < ncfg; ++c) {
            const Cfg cf = cfgs[c];
            uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(cf.N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
            if (cf.a_mode == 2) idesc |= (1u << 15);
            const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
            uint64_t ad, bd;
            if (cf.a_mode == 0) ad = desc(a_base, 128, 256, 0);             // [rowgroup][2 kgroups][8 x 16 B]
            else if (cf.a_mode == 1) ad = desc(a_base, 16, 1024, 2);         // SW128: 8 rows x 128 B atoms
            else if (cf.a_mode == 2) ad = desc(a_base, 128, 160, 0);         // MN-major, overlapping groups
            // "slab" layouts of the conv kernels: [k group][row][16 B], rows 16 B apart (SBO 128), slabs LBO apart;
            // a_mode 3..8: slab of 137 rows (start +0 / +16 / +64 B), slab of 136 rows (start +0 / +16 / +64 B)
            else if (cf.a_mode <= 5) ad = desc(a_base + (cf.a_mode == 3 ? 0 : (cf.a_mode == 4 ? 16 : 64)), 137 * 16, 128, 0);
            else ad = desc(a_base + (cf.a_mode == 6 ? 0 : (cf.a_mode == 7 ? 16 : 64)), 136 * 16, 128, 0);
            if (cf.b_mode == 0) bd = desc(b_base, 128, 256, 0);
            else if (cf.b_mode == 1) bd = desc(b_base, 16, 1024, 2);
            else bd = desc(b_base, cf.N * 16, 128, 0);                       // weight slabs [k group][N][16 B]
            const uint64_t astep = cf.a_mode >= 3 ? 0 : 2;   // slab modes keep their start alignment
            for (int rep = 0; rep < 2; ++rep) {
                const long long t0 = clock64();
                if (cf.same_d == 2) {
                    // fully unrolled groups of 16 with compile-time offsets: the pure issue rate
                    for (int i = 0; i < iters; i += 16) {
#pragma unroll
                        for (int u = 0; u < 16; ++u) umma(tmem + (u & 7) * 32, ad + (u & 3) * astep, bd, idesc, 1u);
                    }
                } else {
#pragma unroll 1
                    for (int i = 0; i < iters; ++i) {
                        const uint32_t dcol = cf.same_d ? 0u : static_cast<uint32_t>(i & 7) * 32u;
                        umma(tmem + dcol, ad + ((i & 3) * 2), bd, idesc, 1u);
                    }
                }
                const long long t1 = clock64();
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                while (!try_wait(&bar, phase)) {}
                phase ^= 1;
                const long long t2 = clock64();
                if (blockIdx.x == 0 && rep == 1) {
                    out[c * 2] = t1 - t0;
                    out[c * 2 + 1] = t2 - t0;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// Is the ~46-cycle issue floor per issuing thread or global?  `nwarps` warps (lane 0 of each) issue concurrently
// into their own TMEM columns; reports cycles per MMA as seen by each issuer (same value as 1 warp => per thread).
__global__ void __launch_bounds__(128, 1) bench_multi(int N, int nwarps, int iters, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x;
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const int w = tid >> 5;
    if ((tid & 31) == 0 && w < nwarps) {
        const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
        const uint64_t ad = desc(smem_u32(smem), 128, 256, 0), bd = desc(smem_u32(smem + 96 * 1024), 128, 256, 0);
        const uint32_t d0 = tmem + w * 128;
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 16) {
#pragma unroll
            for (int u = 0; u < 16; ++u) umma(d0, ad + (u & 3) * 2, bd, idesc, 1u);
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[w])) : "memory");
        while (!try_wait(&bar[w], 0)) {}
        const long long t2 = clock64();
        if (blockIdx.x == 0) {
            out[w * 2] = t1 - t0;
            out[w * 2 + 1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    Cfg h[] = {{32, 0, 0, 0}, {32, 1, 0, 0}, {32, 2, 0, 0}, {32, 0, 0, 1}, {32, 0, 0, 2}, {32, 1, 0, 2}, {32, 2, 0, 2}, {16, 0, 0, 2},
               {64, 0, 0, 2}, {64, 1, 0, 2}, {128, 0, 0, 2}, {128, 1, 0, 2}, {256, 0, 0, 2}, {256, 1, 0, 2}, {256, 1, 1, 2},
               {128, 0, 2, 2}, {128, 3, 2, 2}, {128, 4, 2, 2}, {128, 5, 2, 2}, {128, 6, 2, 2}, {128, 7, 2, 2}, {128, 8, 2, 2},
               {64, 0, 2, 2},  {64, 3, 2, 2},  {64, 4, 2, 2},  {64, 5, 2, 2},  {64, 6, 2, 2},  {64, 7, 2, 2},  {64, 8, 2, 2},
               {32, 3, 2, 2},  {32, 4, 2, 2},  {32, 6, 2, 2},  {32, 7, 2, 2}};
    const int n = sizeof(h) / sizeof(h[0]);
    Cfg* d;
    long long* o;
    cudaMalloc(&d, sizeof(h));
    cudaMalloc(&o, n * 16);
    static_assert(sizeof(h) / sizeof(h[0]) <= 64, "");
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 512;
    for (int grid : {1, 148}) {
        bench<<<grid, 128, 200 * 1024>>>(d, n, iters, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long r[128];
        cudaMemcpy(r, o, n * 16, cudaMemcpyDeviceToHost);
        printf("grid %d\n", grid);
        for (int c = 0; c < n; ++c)
            printf("  N=%3d a_mode=%d b_mode=%d same_d=%d : issue %.1f cyc/mma, complete %.1f cyc/mma\n", h[c].N, h[c].a_mode, h[c].b_mode,
                   h[c].same_d, double(r[2 * c]) / iters, double(r[2 * c + 1]) / iters);
    }
    cudaFuncSetAttribute(bench_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int N : {32, 128})
        for (int nw : {1, 2, 4}) {
            bench_multi<<<1, 128, 200 * 1024>>>(N, nw, iters, o);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long r[8];
            cudaMemcpy(r, o, 64, cudaMemcpyDeviceToHost);
            printf("multi-issuer N=%d warps=%d:", N, nw);
            for (int w = 0; w < nw; ++w) printf("  w%d issue %.1f complete %.1f", w, double(r[2 * w]) / iters, double(r[2 * w + 1]) / iters);
            printf("  (cycles per MMA of that issuer)\n");
        }
    return 0;
}
