// Microbenchmark: how fast can ONE SM pull an L2-resident stream into shared memory with cp.async.bulk
// (the weight rings of the conv / LSTM kernels)?  One producer thread per CTA keeps `depth` chunks of
// `chunk` bytes in flight through an mbarrier ring; reports bytes per SM clock, for 1 CTA and for one CTA
// on every SM, all reading the SAME 512 KB (as the kernels do) or per-CTA private regions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream.bin tma_stream.cu && ./tma_stream.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// nprod producers (lane 0 of warps 0..nprod-1, or lanes 0..nprod-1 of warp 0 when same_warp), each with its own ring
__global__ void __launch_bounds__(128, 1) stream(const float* src, long long region_floats, int private_regions, int chunk, int depth,
                                                  int total_chunks, long long* out, int nprod, int same_warp) {
    extern __shared__ __align__(1024) unsigned char smem_all[];
    __shared__ uint64_t full_all[4][16];
    const int pid = same_warp ? threadIdx.x : (threadIdx.x >> 5);
    const bool is_prod = same_warp ? (threadIdx.x < nprod) : ((threadIdx.x & 31) == 0 && pid < nprod);
    unsigned char* smem = smem_all + static_cast<size_t>(pid % 4) * chunk * depth;
    uint64_t* full = full_all[pid % 4];
    const float* base = src + (private_regions ? blockIdx.x * region_floats : 0);
    if (is_prod) {
        for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int per_region = static_cast<int>(region_floats * 4 / chunk);
        auto issue = [&](int g) {
            const int stg = g % depth;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[stg])), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(smem + static_cast<size_t>(stg) * chunk)),
                         "l"(base + static_cast<long long>(g % per_region) * (chunk / 4)), "r"(chunk), "r"(smem_u32(&full[stg]))
                         : "memory");
        };
        const long long t0 = clock64();
        for (int g = 0; g < depth; ++g) issue(g);
        for (int g = 0; g < total_chunks; ++g) {
            const int stg = g % depth;
            const uint32_t parity = (g / depth) & 1;
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok)
                             : "r"(smem_u32(&full[stg])), "r"(parity)
                             : "memory");
            if (g + depth < total_chunks) issue(g + depth);
        }
        const long long t1 = clock64();
        if (pid == 0) out[blockIdx.x] = t1 - t0;
    }
}

int main() {
    const long long region_floats = 512 * 1024 / 4;
    float* src;
    long long* out;
    cudaMalloc(&src, 148 * region_floats * 4);
    cudaMemset(src, 0, 148 * region_floats * 4);
    cudaMalloc(&out, 148 * 8);
    cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mode : {0, 1, 2, 3, 4})
        for (int grid : {1, 148})
            for (int chunk : {2048, 8192, 16384, 32768})
                for (int depth : {2, 4}) {
                    const int priv = 0;
                    const int nprod = mode == 0 ? 1 : (mode == 1 || mode == 3 ? 2 : 4), same_warp = mode >= 3;
                    if (static_cast<long long>(chunk) * depth * nprod > 192 * 1024) continue;
                    const int total = 64 * 1024 * 1024 / chunk / 16;   // 4 MB per producer
                    for (int rep = 0; rep < 2; ++rep)
                        stream<<<grid, 128, 200 * 1024>>>(src, region_floats, priv, chunk, depth, total, out, nprod, same_warp);
                    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
                    long long h[148];
                    cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
                    long long mx = 0;
                    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                    printf("%d producer(s) %s grid %3d chunk %5d depth %d : %.1f B/clk per SM, %.0f clk per copy per producer\n", nprod,
                           same_warp ? "(lanes of one warp)" : "(one per warp)     ", grid, chunk, depth,
                           double(total) * chunk * nprod / double(mx), double(mx) / total);
                }
    return 0;
}
