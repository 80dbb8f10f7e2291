// C-ABI plumbing shared by every entry point of libmmla_b200.so: error channel, device query,
// host-side CRC-32C (TF tensor-bundle checksums).
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}

void mmla_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

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
