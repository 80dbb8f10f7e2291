// TEMPORARY: placeholders until overlap.cu / nets.cu land (removed in the same round).
#include "common.cuh"
#define EXPORT extern "C" __attribute__((visibility("default")))
EXPORT int mmla_overlap_features(const int16_t*, int64_t, const int64_t*, const int32_t*, int64_t, int32_t, int64_t,
                                 int32_t, float*, float*, float*, uint8_t*, void*) {
    mmla_set_error("overlap features: not built yet");
    return MMLA_EUNSUP;
}
EXPORT int mmla_net_create(int32_t, int32_t, int32_t, const float*, int64_t, MmlaNet**) {
    mmla_set_error("nets: not built yet");
    return MMLA_EUNSUP;
}
EXPORT void mmla_net_destroy(MmlaNet*) {}
EXPORT int64_t mmla_net_workspace_bytes(const MmlaNet*, int64_t) { return -1; }
EXPORT int mmla_net_forward(MmlaNet*, const void*, int32_t, int64_t, void*, int64_t, float*, int32_t*, void*) {
    mmla_set_error("nets: not built yet");
    return MMLA_EUNSUP;
}
