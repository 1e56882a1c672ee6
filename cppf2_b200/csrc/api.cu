// Library-level entry points: error strings, version, cached device properties.
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace cppf {

// Properties of the CURRENT device, cached per device ordinal (one process may drive several GPUs, and several host
// threads may get here at once).
const DeviceInfo &device_info() {
    constexpr int kMaxDevices = 64;
    static DeviceInfo info[kMaxDevices];
    static std::once_flag once[kMaxDevices];
    static const DeviceInfo fallback{148, 126ll << 20, 10, 0, 227 * 1024};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return fallback;
    std::call_once(once[dev], [dev] {
        info[dev] = fallback;
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return;
        info[dev].sm_count = p.multiProcessorCount;
        info[dev].l2_bytes = p.l2CacheSize;
        info[dev].cc_major = p.major;
        info[dev].cc_minor = p.minor;
        info[dev].max_smem_optin = static_cast<int>(p.sharedMemPerBlockOptin);
    });
    return info[dev];
}

bool frame_pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("CPPF_FRAME_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

// True exactly once per (call site tag, device): the caller then performs its per-device one-time setup.
bool first_use_on_device(std::atomic<uint64_t> *seen) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    const uint64_t bit = 1ull << dev;
    return (seen->fetch_or(bit) & bit) == 0;
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
