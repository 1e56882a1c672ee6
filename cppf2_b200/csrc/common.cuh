// Shared device/host helpers of libcppf_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cppf_b200.h"

#define CPPF_API extern "C" __attribute__((visibility("default")))

#define CPPF_CUDA_TRY(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            fprintf(stderr, "[cppf_b200] %s failed at %s:%d: %s\n", #expr, __FILE__, __LINE__,           \
                    cudaGetErrorString(_e));                                                             \
            return CPPF_ERR_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

#define CPPF_LAUNCH_CHECK() CPPF_CUDA_TRY(cudaGetLastError())

namespace cppf {

// Programmatic dependent launch along the frame's kernel sequence (cppf_frame_pose): every frame kernel begins with
// pdl_wait() -- the previous kernel has completed and its writes are visible, i.e. plain stream order -- and then lets the
// NEXT kernel's CTAs be scheduled (pdl_trigger) as soon as all of this kernel's CTAs have started: launch latency, CTA
// scheduling and the next kernel's prologue overlap this kernel's tail instead of following it.  Without the launch
// attribute both instructions are no-ops, so the same kernels serve the single-job entry points.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
    pdl_wait();
    pdl_trigger();
}

bool frame_pdl_enabled();      // CPPF_FRAME_PDL=0 switches the attribute off (api.cu)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_frame_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = frame_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct DeviceInfo {
    int sm_count;
    int64_t l2_bytes;
    int cc_major, cc_minor;
    int max_smem_optin;
};

// Properties of the current device, cached per device ordinal; B200: 148 SMs, ~126 MB L2, 227 KB opt-in shared memory.
const DeviceInfo &device_info();

// Per-device one-time setup (cudaFuncSetAttribute opt-ins): thread-safe, and repeated for every device a process drives.
// A failed attempt is not remembered as done.
bool first_use_on_device(std::atomic<uint64_t> *seen);
#define CPPF_TRY_ONCE_PER_DEVICE(expr)                         \
    do {                                                       \
        static std::atomic<uint64_t> _seen{0};                 \
        if (cppf::first_use_on_device(&_seen)) {               \
            cudaError_t _e1 = (expr);                          \
            if (_e1 != cudaSuccess) {                          \
                _seen.store(0);                                \
                CPPF_CUDA_TRY(_e1);                            \
            }                                                  \
        }                                                      \
    } while (0)

// Row of the tuple-index matrix, int64 (reference dtype) or int32, arbitrary row stride.
struct IdxView {
    const void *ptr;
    int64_t stride;
    int is_i64;
    __device__ __forceinline__ int64_t at(int64_t row, int col) const {
        return is_i64 ? static_cast<const int64_t *>(ptr)[row * stride + col]
                      : static_cast<int64_t>(static_cast<const int32_t *>(ptr)[row * stride + col]);
    }
};

// ---- float32 recipes that must match torch-CPU bit for bit ----------------
// Every operation is an explicit round-to-nearest intrinsic so that neither -fmad nor the optimiser
// can contract or reassociate them.

__device__ __forceinline__ float norm3_torch(float x, float y, float z) {
    // torch.norm(dim=-1) over 3 contiguous floats on CPU: sqrt(fma(z,z,fma(y,y,x*x)))
    return __fsqrt_rn(__fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
}

__device__ __forceinline__ float norm3_numpy(float x, float y, float z) {
    // np.linalg.norm(axis=-1) on float32: sqrt((x*x + y*y) + z*z)
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// torch.cross(x, ab) on CPU: fma(x1, ab2, -(x2*ab1)) and cyclic
__device__ __forceinline__ void cross_torch(const float x[3], const float ab[3], float y[3]) {
    y[0] = __fmaf_rn(x[1], ab[2], -__fmul_rn(x[2], ab[1]));
    y[1] = __fmaf_rn(x[2], ab[0], -__fmul_rn(x[0], ab[2]));
    y[2] = __fmaf_rn(x[0], ab[1], -__fmul_rn(x[1], ab[0]));
}

// Unit pair direction `ab` and the in-plane unit vector `x` (train_dino.py:178-191 / :219-229).
// Returns false when |a-b| <= 1e-7 (masked pair).
__device__ __forceinline__ bool pair_frame(const float a[3], const float b[3], bool clamp_co, float ab[3], float x[3]) {
    ab[0] = __fsub_rn(a[0], b[0]);
    ab[1] = __fsub_rn(a[1], b[1]);
    ab[2] = __fsub_rn(a[2], b[2]);
    float n = norm3_torch(ab[0], ab[1], ab[2]);
    if (!(n > 1e-7f)) return false;
    float d = n < 1e-7f ? 1e-7f : n;
    ab[0] = __fdiv_rn(ab[0], d);
    ab[1] = __fdiv_rn(ab[1], d);
    ab[2] = __fdiv_rn(ab[2], d);
    float co[3] = {0.0f, -ab[2], ab[1]};
    float cn = norm3_torch(co[0], co[1], co[2]);
    if (cn < 1e-7f) {  // ab parallel to the x axis
        co[0] = -ab[1];
        co[1] = ab[0];
        co[2] = 0.0f;
        cn = norm3_torch(co[0], co[1], co[2]);
    }
    if (clamp_co && cn < 1e-7f) cn = 1e-7f;
    x[0] = __fdiv_rn(co[0], cn);
    x[1] = __fdiv_rn(co[1], cn);
    x[2] = __fdiv_rn(co[2], cn);
    return true;
}

// a / b with the correctly rounded quotient of IEEE division (what torch computes, train_dino.py:199) for a divisor that
// serves several quotients (the voxel size of a launch, the norm of a candidate direction): q = a * RN(1/b) followed by two residual corrections in FMA arithmetic -- the sequence the
// hardware division's fast path runs after its reciprocal refinement, minus the per-call reciprocal and range check (5
// instructions instead of ~11; three divisions are a quarter of a vote's instructions and the shared-memory mode is issue-bound).
// The operands here are finite and far from the denormal range (grid coordinates in metres over a voxel size); a NaN stays a
// NaN.  Bit-exactness of the grids against the torch-CPU golden vectors and the oracle is what the tests check.
__device__ __forceinline__ float div_by(float a, float b, float inv_b) {
    float q = __fmul_rn(a, inv_b);
    float r = __fmaf_rn(-b, q, a);
    q = __fmaf_rn(r, inv_b, q);
    r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, inv_b, q);
}

// Order-preserving float32 <-> uint32 key (for radix selection and atomic min/max on floats).
__device__ __host__ __forceinline__ uint32_t float_to_key(float f) {
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float key_to_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Counter-based uniform in [0,1) for the multinomial draw of row `ctr` (splitmix64 finaliser, 24 mantissa bits): the
// stand-alone sampler (targets.cu) and the sampling epilogue of the heads (heads_tc.cu) draw the same numbers.
__device__ __forceinline__ float uniform_from_counter(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (ctr + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
}

// Histogram add of a whole (converged) warp: lanes with `on` add 1 to h[digit].  Keys of one selection pass usually share
// their leading digits, so the common case is "every lane hits the same bin": one leader adds the group's size (a plain
// per-lane atomicAdd would serialise 32-deep on one shared-memory address).  Mixed digits fall back to per-lane adds, which
// then spread over the bins.  (__match_any_sync would aggregate every group, but costs ~100 cycles per call on this part:
// it made a 50 000-key selection by one CTA take 75 us.)
__device__ __forceinline__ void warp_hist_add(uint32_t *h, uint32_t digit, bool on) {
    const uint32_t active = __ballot_sync(0xffffffffu, on);
    if (active == 0u) return;
    const int leader = __ffs(active) - 1;
    const uint32_t d0 = __shfl_sync(0xffffffffu, digit, leader);
    const uint32_t same = __ballot_sync(0xffffffffu, on && digit == d0);
    if (same == active) {
        if ((threadIdx.x & 31) == leader) atomicAdd(&h[d0], static_cast<uint32_t>(__popc(active)));
    } else if (on) {
        atomicAdd(&h[digit], 1u);
    }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// corners / grid_res of a cloud (train_dino.py:172-173; guard eval.py:200): ONE CTA of 1024 threads, coalesced sweep,
// shared-memory tree reduction.  Shared by cppf_cloud_bounds, the SHOT search grid and the batched frame path.
__device__ __forceinline__ void cloud_bounds_shared(const float *__restrict__ pc, int64_t n, float res,
                                                  cppf_grid_geom *__restrict__ geom) {
    __shared__ float s_lo[3][32], s_hi[3][32];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    // flat float index i = 3*p + k: consecutive threads read consecutive floats; k = i % 3
    for (int64_t i = threadIdx.x; i < 3 * n; i += 3 * 1024) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            int64_t j = i + static_cast<int64_t>(u) * 1024;
            if (j < 3 * n) {
                float v = pc[j];
                int k = static_cast<int>(j % 3);
                // k is thread-dependent; keep the three accumulators in registers with selects
                lo[0] = (k == 0 && v < lo[0]) ? v : lo[0];
                lo[1] = (k == 1 && v < lo[1]) ? v : lo[1];
                lo[2] = (k == 2 && v < lo[2]) ? v : lo[2];
                hi[0] = (k == 0 && v > hi[0]) ? v : hi[0];
                hi[1] = (k == 1 && v > hi[1]) ? v : hi[1];
                hi[2] = (k == 2 && v > hi[2]) ? v : hi[2];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if (lane_id() == 0) {
            s_lo[k][threadIdx.x >> 5] = lo[k];
            s_hi[k][threadIdx.x >> 5] = hi[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float l = s_lo[k][threadIdx.x], h = s_hi[k][threadIdx.x];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
                h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
            }
            lo[k] = l;
            hi[k] = h;
        }
        if (threadIdx.x == 0) {
            uint32_t flags = (n <= 0) ? CPPF_STATUS_EMPTY : 0u;
            float ext_max = 0.0f;
            int64_t cells = 1;
            for (int k = 0; k < 3; ++k) {
                geom->lo[k] = lo[k];
                geom->hi[k] = hi[k];
                float ext = __fsub_rn(hi[k], lo[k]);
                ext_max = fmaxf(ext_max, ext);
                int64_t g = (n > 0) ? static_cast<int64_t>(__fdiv_rn(ext, res)) + 1 : 1;  // trunc toward zero
                geom->grid_res[k] = g;
                cells *= g;
            }
            if (__fdiv_rn(ext_max, res) > 1000.0f) flags |= CPPF_STATUS_GRID_GUARD;  // eval.py:200
            geom->res = res;
            geom->cells = cells;
            geom->flags = flags;
        }
    }
}

inline int div_up(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// Grid size for a grid-stride kernel: whole multiples of the SM count, capped by the work available.
inline int grid_for(int64_t work_items, int threads_per_block, int blocks_per_sm) {
    const DeviceInfo &d = device_info();
    int64_t need = (work_items + threads_per_block - 1) / threads_per_block;
    int64_t cap = static_cast<int64_t>(d.sm_count) * blocks_per_sm;
    if (need < 1) need = 1;
    return static_cast<int>(need < cap ? need : cap);
}

}  // namespace cppf
