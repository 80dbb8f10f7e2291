// Stand-alone host->device bandwidth probe: one process per GPU, pinned host buffer, one cudaMemcpyAsync per chunk.
// Answers VERDICT r01 item 8: is the e2e floor of bench.py (23 GB/s per GPU with eight concurrent uploads) the box's
// PCIe / host-memory limit or the pipeline's staging?  Launch N copies at once (scripts/microbench/run_h2d.sh) and
// compare the per-GPU figure with bench.py's `e2e.h2d_only_ms`.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o h2d_bw.bin h2d_bw.cu
//   ./h2d_bw.bin <device> <MiB per copy> <copies per repetition> <repetitions> <start-after-epoch-seconds>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int dev = argc > 1 ? atoi(argv[1]) : 0;
    const size_t mib = argc > 2 ? atol(argv[2]) : 96;
    const int copies = argc > 3 ? atoi(argv[3]) : 2;
    const int reps = argc > 4 ? atoi(argv[4]) : 20;
    const double start_at = argc > 5 ? atof(argv[5]) : 0.0;
    CK(cudaSetDevice(dev));
    const size_t bytes = mib << 20;
    void *h = nullptr, *d = nullptr;
    CK(cudaMallocHost(&h, bytes * copies));
    CK(cudaMalloc(&d, bytes * copies));
    memset(h, 1, bytes * copies);
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int c = 0; c < copies; ++c) CK(cudaMemcpyAsync((char*)d + c * bytes, (char*)h + c * bytes, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    if (start_at > 0) {                                    // crude rendezvous: all processes start their timed loop together
        struct timespec ts;
        do { clock_gettime(CLOCK_REALTIME, &ts); } while (ts.tv_sec + ts.tv_nsec * 1e-9 < start_at);
    }
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < reps; ++r)
        for (int c = 0; c < copies; ++c)
            CK(cudaMemcpyAsync((char*)d + c * bytes, (char*)h + c * bytes, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"device\": %d, \"mib_per_copy\": %zu, \"copies\": %d, \"reps\": %d, \"ms\": %.3f, \"gb_per_s\": %.2f}\n", dev, mib, copies,
           reps, ms, (double)bytes * copies * reps / (ms * 1e-3) / 1e9);
    return 0;
}
