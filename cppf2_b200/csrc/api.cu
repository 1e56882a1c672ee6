// Library-level entry points: error strings, version, cached device properties.
#include "common.cuh"

#include <mutex>

namespace cppf {

const DeviceInfo &device_info() {
    static DeviceInfo info{148, 126ll << 20, 10, 0, 227 * 1024};
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return;
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return;
        info.sm_count = p.multiProcessorCount;
        info.l2_bytes = p.l2CacheSize;
        info.cc_major = p.major;
        info.cc_minor = p.minor;
        info.max_smem_optin = static_cast<int>(p.sharedMemPerBlockOptin);
    });
    return info;
}

}  // namespace cppf

CPPF_API const char *cppf_error_string(int code) {
    switch (code) {
        case CPPF_OK: return "ok";
        case CPPF_ERR_INVALID_ARGUMENT: return "invalid argument";
        case CPPF_ERR_CUDA: return "CUDA runtime or launch failure (see stderr)";
        case CPPF_ERR_WORKSPACE: return "workspace too small";
        case CPPF_ERR_UNSUPPORTED: return "unsupported (no caller in the reference for this entry point or size)";
        case CPPF_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown error";
    }
}

CPPF_API int cppf_version(void) { return 100; }

CPPF_API int cppf_device_info(int *sm_count, int64_t *l2_bytes, int *cc_major, int *cc_minor) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return CPPF_ERR_NO_DEVICE;
    }
    const cppf::DeviceInfo &d = cppf::device_info();
    if (sm_count) *sm_count = d.sm_count;
    if (l2_bytes) *l2_bytes = d.l2_bytes;
    if (cc_major) *cc_major = d.cc_major;
    if (cc_minor) *cc_minor = d.cc_minor;
    return CPPF_OK;
}
