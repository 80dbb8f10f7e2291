// C-ABI plumbing shared by every entry point of libmmla_b200.so: error channel, device query,
// host-side CRC-32C (TF tensor-bundle checksums).
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

// Launch trace: one CUDA event after every launch on the traced stream.  Kernels of one stream run in order, so
// the time between consecutive events is the device time of the later kernel (plus any launch gap).
struct TraceRec {
    const char* name;
    cudaEvent_t ev;
};
std::mutex g_trace_mu;
std::atomic<bool> g_trace_on{false};
cudaStream_t g_trace_stream = nullptr;
cudaEvent_t g_trace_start = nullptr;
std::vector<TraceRec> g_trace;
}  // namespace

void mmla_count_launch(const char* kernel_name, cudaStream_t st) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_trace_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_trace_mu);
    if (!g_trace_on.load() || st != g_trace_stream) return;
    TraceRec r;
    r.name = kernel_name ? kernel_name : "?";
    if (cudaEventCreate(&r.ev) != cudaSuccess) return;
    cudaEventRecord(r.ev, st);
    g_trace.push_back(r);
}

extern "C" __attribute__((visibility("default"))) int mmla_trace_begin(void* stream) {
    std::lock_guard<std::mutex> lk(g_trace_mu);
    for (auto& r : g_trace) cudaEventDestroy(r.ev);
    g_trace.clear();
    if (g_trace_start) cudaEventDestroy(g_trace_start);
    g_trace_start = nullptr;
    g_trace_stream = static_cast<cudaStream_t>(stream);
    MMLA_CUDA_CHECK(cudaEventCreate(&g_trace_start));
    MMLA_CUDA_CHECK(cudaEventRecord(g_trace_start, g_trace_stream));
    g_trace_on.store(true);
    return MMLA_OK;
}

extern "C" __attribute__((visibility("default"))) int mmla_trace_end(char* names_host, int64_t names_bytes, float* ms_host,
                                                                    int32_t max_records) {
    std::lock_guard<std::mutex> lk(g_trace_mu);
    g_trace_on.store(false);
    if (!g_trace_start) {
        mmla_set_error("mmla_trace_end without mmla_trace_begin");
        return MMLA_EINVAL;
    }
    int n = 0;
    int64_t pos = 0;
    cudaEvent_t prev = g_trace_start;
    for (auto& r : g_trace) {
        if (n < max_records) {
            MMLA_CUDA_CHECK(cudaEventSynchronize(r.ev));
            float ms = 0.f;
            MMLA_CUDA_CHECK(cudaEventElapsedTime(&ms, prev, r.ev));
            if (ms_host) ms_host[n] = ms;
            const int64_t len = static_cast<int64_t>(strlen(r.name));
            if (names_host && pos + len + 1 < names_bytes) {
                memcpy(names_host + pos, r.name, static_cast<size_t>(len));
                names_host[pos + len] = '\n';
                pos += len + 1;
            }
            ++n;
        }
        prev = r.ev;
    }
    if (names_host && names_bytes > 0) names_host[pos < names_bytes ? pos : names_bytes - 1] = 0;
    for (auto& r : g_trace) cudaEventDestroy(r.ev);
    g_trace.clear();
    cudaEventDestroy(g_trace_start);
    g_trace_start = nullptr;
    return n;
}

void mmla_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int mmla_num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        mmla_set_error("cudaGetDevice failed (no CUDA device?)");
        return -1;
    }
    if (dev < 0 || dev >= 64) return -1;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            mmla_set_error("cudaDeviceGetAttribute(MultiProcessorCount) failed");
            return -1;
        }
        cached[dev] = n;
    }
    return cached[dev];
}

extern "C" __attribute__((visibility("default"))) const char* mmla_last_error(void) { return g_err; }

extern "C" __attribute__((visibility("default"))) int mmla_abi_version(void) { return 1; }

extern "C" __attribute__((visibility("default"))) uint32_t mmla_crc32c_host(const void* data, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    const unsigned char* p = static_cast<const unsigned char*>(data);
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

extern "C" __attribute__((visibility("default"))) int64_t mmla_launch_count(void) {
    return g_launches.load(std::memory_order_relaxed);
}
