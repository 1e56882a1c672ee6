// Instance cloud preparation on B200 -- the step in front of SHOT:
//   back-projection of the masked depth pixels     utils/util.py:2586-2607 + the callers' un-flip, eval.py:185-189
//   voxel down-sampling, one random member / voxel utils/util.py:39-46 (Open3D voxel_down_sample_and_trace + np.random.choice)
//   the 50 000-point cap                           eval.py:195-198 (np.random.randint with replacement)
// Order-preserving stream compaction (flags -> ascending index list) is the shared primitive: per-block counts, one
// block scanning the counts, per-block scatter.  All arithmetic that decides a float32 output or a voxel is float64 with
// explicit round-to-nearest intrinsics, like numpy / Open3D (which promote the float32 cloud to double).
#include "common.cuh"

namespace cppf {

constexpr int kCompactThreads = 256;
constexpr int kCompactItems = 8;                                  // per thread
constexpr int kCompactTile = kCompactThreads * kCompactItems;     // 2048 flags per block

// flags u8 [n] -> block_counts[b]
__global__ void __launch_bounds__(kCompactThreads) compact_count_kernel(const uint8_t *__restrict__ flags, int64_t n,
                                                                        int32_t *__restrict__ block_counts) {
    __shared__ int s_warp[kCompactThreads / 32];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kCompactTile + static_cast<int64_t>(threadIdx.x) * kCompactItems;
    int c = 0;
#pragma unroll
    for (int k = 0; k < kCompactItems; ++k) c += (base + k < n && flags[base + k]) ? 1 : 0;
    c = warp_sum(c);
    if (lane_id() == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kCompactThreads / 32; ++w) t += s_warp[w];
        block_counts[blockIdx.x] = t;
    }
}

// exclusive scan of block_counts[0..n_blocks) in place (single block), total -> *count_out
__global__ void __launch_bounds__(1024) compact_scan_kernel(int32_t *__restrict__ block_counts, int n_blocks,
                                                            int64_t *__restrict__ count_out) {
    __shared__ int s_part[1024];
    const int per = (n_blocks + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(lo + per, n_blocks);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += block_counts[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {     // Hillis-Steele inclusive scan of the per-thread sums
        const int v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = s_part[threadIdx.x] - sum;
    for (int i = lo; i < hi; ++i) {
        const int c = block_counts[i];
        block_counts[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) *count_out = s_part[1023];
}

// out_idx[block offset + local rank] = i for every set flag, ascending
__global__ void __launch_bounds__(kCompactThreads) compact_scatter_kernel(const uint8_t *__restrict__ flags, int64_t n,
                                                                          const int32_t *__restrict__ block_offsets,
                                                                          int32_t *__restrict__ out_idx) {
    __shared__ int s_warp[kCompactThreads / 32];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kCompactTile + static_cast<int64_t>(threadIdx.x) * kCompactItems;
    bool f[kCompactItems];
    int c = 0;
#pragma unroll
    for (int k = 0; k < kCompactItems; ++k) {
        f[k] = base + k < n && flags[base + k];
        c += f[k] ? 1 : 0;
    }
    int incl = c;                               // inclusive scan of the per-thread counts inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane_id() >= o) incl += v;
    }
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    int warp_base = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) warp_base += s_warp[w];
    int pos = block_offsets[blockIdx.x] + warp_base + incl - c;
#pragma unroll
    for (int k = 0; k < kCompactItems; ++k)
        if (f[k]) out_idx[pos++] = static_cast<int32_t>(base + k);
}

static int compact(const uint8_t *flags, int64_t n, int32_t *block_counts, int32_t *out_idx, int64_t *count_out, cudaStream_t s) {
    const int n_blocks = static_cast<int>((n + kCompactTile - 1) / kCompactTile);
    compact_count_kernel<<<n_blocks, kCompactThreads, 0, s>>>(flags, n, block_counts);
    compact_scan_kernel<<<1, 1024, 0, s>>>(block_counts, n_blocks, count_out);
    compact_scatter_kernel<<<n_blocks, kCompactThreads, 0, s>>>(flags, n, block_counts, out_idx);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

// ---- back-projection --------------------------------------------------------------------------------------------------
struct Kinv {
    double m[9];
};

__global__ void __launch_bounds__(256) backproject_flags_kernel(const void *__restrict__ depth, int depth_is_u16,
                                                                const uint8_t *__restrict__ mask, int64_t pixels,
                                                                uint8_t *__restrict__ flags) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < pixels; i += stride) {
        const bool pos = depth_is_u16 ? static_cast<const uint16_t *>(depth)[i] > 0 : static_cast<const float *>(depth)[i] > 0.0f;
        flags[i] = (mask[i] != 0 && pos) ? 1 : 0;      // np.logical_and(instance_mask, depth > 0)
    }
}

// pts = (Kinv @ [u, v, 1]) * z / xyz_z in float64, then float32 (the two sign flips of x and y cancel exactly)
__global__ void __launch_bounds__(256) backproject_points_kernel(const void *__restrict__ depth, int depth_is_u16, double depth_div,
                                                                 int W, Kinv kinv, const int32_t *__restrict__ pix,
                                                                 const int64_t *__restrict__ count, float *__restrict__ pc) {
    const int64_t n = *count;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int p = pix[i];
        const double u = static_cast<double>(p % W), v = static_cast<double>(p / W);
        const double raw = depth_is_u16 ? static_cast<double>(static_cast<const uint16_t *>(depth)[p])
                                        : static_cast<double>(static_cast<const float *>(depth)[p]);
        const double z = depth_div != 1.0 ? __ddiv_rn(raw, depth_div) : raw;     // depth / 1000.
        double xyz[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            xyz[k] = __dadd_rn(__dadd_rn(__dmul_rn(kinv.m[3 * k], u), __dmul_rn(kinv.m[3 * k + 1], v)), kinv.m[3 * k + 2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) pc[3 * i + k] = static_cast<float>(__ddiv_rn(__dmul_rn(xyz[k], z), xyz[2]));
    }
}

// ---- voxel down-sampling ----------------------------------------------------------------------------------------------
constexpr unsigned long long kEmptyKey = 0xffffffffffffffffull;

__global__ void __launch_bounds__(256) voxel_min_kernel(const float *__restrict__ pc, int64_t n, uint32_t *__restrict__ min_key) {
    float m[3] = {INFINITY, INFINITY, INFINITY};
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
#pragma unroll
        for (int k = 0; k < 3; ++k) m[k] = fminf(m[k], pc[3 * i + k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m[k] = fminf(m[k], __shfl_xor_sync(0xffffffffu, m[k], o));
        if (lane_id() == 0) atomicMin(&min_key[k], float_to_key(m[k]));
    }
}

__global__ void __launch_bounds__(256) voxel_table_init_kernel(unsigned long long *__restrict__ keys, unsigned long long *__restrict__ vals,
                                                               int64_t cap, uint32_t *__restrict__ min_key) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < cap; i += stride) {
        keys[i] = kEmptyKey;
        vals[i] = kEmptyKey;
    }
    if (blockIdx.x == 0 && threadIdx.x < 3) min_key[threadIdx.x] = 0xffffffffu;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Open3D: voxel_min_bound = min_bound - voxel_size/2; voxel = floor((p - voxel_min_bound) / voxel_size) in double.
// Per voxel the member with the smallest (priority, index) wins: priority = injected u01 (float bits) or a counter draw.
__global__ void __launch_bounds__(256) voxel_insert_kernel(const float *__restrict__ pc, int64_t n, double res,
                                                           const uint32_t *__restrict__ min_key, const float *__restrict__ prio,
                                                           uint64_t seed, unsigned long long *__restrict__ keys,
                                                           unsigned long long *__restrict__ vals, int64_t cap_mask,
                                                           int32_t *__restrict__ slot_of) {
    const double half = __dmul_rn(res, 0.5);
    double lo[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) lo[k] = __dsub_rn(static_cast<double>(key_to_float(min_key[k])), half);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned long long key = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double q = floor(__ddiv_rn(__dsub_rn(static_cast<double>(pc[3 * i + k]), lo[k]), res));
            const long long c = static_cast<long long>(q);
            key |= (static_cast<unsigned long long>(c) & 0x1fffffull) << (21 * k);
        }
        const uint32_t pr = prio ? __float_as_uint(prio[i]) : static_cast<uint32_t>(mix64(seed + 0x9E3779B97F4A7C15ull * (i + 1)) >> 32);
        const unsigned long long val = (static_cast<unsigned long long>(pr) << 32) | static_cast<unsigned long long>(i);
        int64_t slot = static_cast<int64_t>(mix64(key) & static_cast<uint64_t>(cap_mask));
        while (true) {
            const unsigned long long prev = atomicCAS(&keys[slot], kEmptyKey, key);
            if (prev == kEmptyKey || prev == key) break;
            slot = (slot + 1) & cap_mask;
        }
        atomicMin(&vals[slot], val);
        slot_of[i] = static_cast<int32_t>(slot);
    }
}

__global__ void __launch_bounds__(256) voxel_winner_kernel(const unsigned long long *__restrict__ vals, const int32_t *__restrict__ slot_of,
                                                           int64_t n, uint8_t *__restrict__ flags) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        flags[i] = (vals[slot_of[i]] & 0xffffffffull) == static_cast<unsigned long long>(i) ? 1 : 0;
}

// out[j] = pc[idx[j]] for j < *count (or count_host when count == nullptr); optional second gather of an int32 side array
__global__ void __launch_bounds__(256) gather_points_kernel(const float *__restrict__ pc, const int32_t *__restrict__ idx,
                                                            const int64_t *__restrict__ count, int64_t count_host,
                                                            float *__restrict__ out, const int32_t *__restrict__ side_in,
                                                            int32_t *__restrict__ side_out) {
    const int64_t n = count ? *count : count_host;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int64_t i = idx[j];
        out[3 * j] = pc[3 * i];
        out[3 * j + 1] = pc[3 * i + 1];
        out[3 * j + 2] = pc[3 * i + 2];
        if (side_in) side_out[j] = side_in[i];
    }
}

static size_t al256(size_t x) { return (x + 255) / 256 * 256; }
static int64_t table_cap(int64_t n) {
    int64_t c = 1024;
    while (c < 2 * n) c <<= 1;
    return c;
}

}  // namespace cppf

using namespace cppf;

CPPF_API int64_t cppf_cloud_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    const int64_t blocks = (n + kCompactTile - 1) / kCompactTile + 1;
    return static_cast<int64_t>(al256(static_cast<size_t>(n)) /* flags */ + al256(sizeof(int32_t) * blocks) + al256(sizeof(int32_t) * n) /* slots */ +
                                2 * al256(sizeof(unsigned long long) * table_cap(n)) + 256 /* min keys */);
}

CPPF_API int cppf_backproject(const void *depth, int depth_is_u16, double depth_div, const uint8_t *mask, int H, int W,
                              const double *kinv_host, float *pc, int32_t *pix, int64_t *count, void *ws, int64_t ws_bytes,
                              void *stream) {
    if (!depth || !mask || !kinv_host || !pc || !pix || !count || !ws || H <= 0 || W <= 0 || depth_div == 0.0) return CPPF_ERR_INVALID_ARGUMENT;
    const int64_t pixels = static_cast<int64_t>(H) * W;
    if (pixels > 0x7fffffff) return CPPF_ERR_UNSUPPORTED;
    if (ws_bytes < cppf_cloud_workspace_bytes(pixels)) return CPPF_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint8_t *flags = static_cast<uint8_t *>(ws);
    int32_t *block_counts = reinterpret_cast<int32_t *>(flags + al256(static_cast<size_t>(pixels)));
    backproject_flags_kernel<<<grid_for(pixels, 256, 8), 256, 0, s>>>(depth, depth_is_u16, mask, pixels, flags);
    const int rc = compact(flags, pixels, block_counts, pix, count, s);
    if (rc != CPPF_OK) return rc;
    Kinv k;
    for (int i = 0; i < 9; ++i) k.m[i] = kinv_host[i];
    backproject_points_kernel<<<grid_for(pixels / 4 + 1, 256, 8), 256, 0, s>>>(depth, depth_is_u16, depth_div, W, k, pix, count, pc);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_voxel_downsample(const float *pc, int64_t n, double res, const float *prio, uint64_t seed, float *pc_out,
                                   int32_t *kept_idx, int64_t *count, const int32_t *side_in, int32_t *side_out, void *ws,
                                   int64_t ws_bytes, void *stream) {
    if (!pc || !pc_out || !kept_idx || !count || !ws || n < 0 || !(res > 0.0)) return CPPF_ERR_INVALID_ARGUMENT;
    if (n > 0x7fffffff || (side_in && !side_out)) return CPPF_ERR_INVALID_ARGUMENT;
    if (ws_bytes < cppf_cloud_workspace_bytes(n)) return CPPF_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        CPPF_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int64_t), s));
        return CPPF_OK;
    }
    const int64_t blocks = (n + kCompactTile - 1) / kCompactTile + 1;
    const int64_t cap = table_cap(n);
    unsigned char *p = static_cast<unsigned char *>(ws);
    uint8_t *flags = p;
    p += al256(static_cast<size_t>(n));
    int32_t *block_counts = reinterpret_cast<int32_t *>(p);
    p += al256(sizeof(int32_t) * blocks);
    int32_t *slot_of = reinterpret_cast<int32_t *>(p);
    p += al256(sizeof(int32_t) * n);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(p);
    p += al256(sizeof(unsigned long long) * cap);
    unsigned long long *vals = reinterpret_cast<unsigned long long *>(p);
    p += al256(sizeof(unsigned long long) * cap);
    uint32_t *min_key = reinterpret_cast<uint32_t *>(p);
    voxel_table_init_kernel<<<grid_for(cap, 256, 8), 256, 0, s>>>(keys, vals, cap, min_key);
    voxel_min_kernel<<<grid_for(n, 256, 4), 256, 0, s>>>(pc, n, min_key);
    voxel_insert_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(pc, n, res, min_key, prio, seed, keys, vals, cap - 1, slot_of);
    voxel_winner_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(vals, slot_of, n, flags);
    const int rc = compact(flags, n, block_counts, kept_idx, count, s);
    if (rc != CPPF_OK) return rc;
    gather_points_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(pc, kept_idx, count, 0, pc_out, side_in, side_out);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_gather_points(const float *pc, const int32_t *idx, int64_t m, float *pc_out, const int32_t *side_in,
                                int32_t *side_out, void *stream) {
    if (!pc || !idx || !pc_out || m < 0 || (side_in && !side_out)) return CPPF_ERR_INVALID_ARGUMENT;
    if (m == 0) return CPPF_OK;
    gather_points_kernel<<<grid_for(m, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(pc, idx, nullptr, m, pc_out, side_in, side_out);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}
