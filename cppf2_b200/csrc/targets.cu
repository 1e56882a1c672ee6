// Decode + vote-target generation on B200 -- replaces eval.py:225-235 (softmax, multinomial draw,
// per-pair metric scale) and generate_target_pairs (dataset.py:118-135).
//
// dtype flow follows the reference when it is fed float32 pairs: the pair difference and its unit
// vector are float32 (numpy norm order, +1e-7 as a weak scalar), everything that meets the float64
// centre or the integer axes is float64, results are rounded to float32 at the end.  Every operation is
// an explicit _rn intrinsic, so targets_tr is bit-exact against numpy; targets_rot differs by at most
// 1 ulp of float32 (device acos vs libm acos, both < 1 ulp in float64).
#include "common.cuh"
#include "frame.cuh"
#include "targets.cuh"

namespace cppf {

__global__ void __launch_bounds__(256) generate_targets_kernel(const float *__restrict__ pairs, int64_t T, Axes ax,
                                                               const double *__restrict__ center,
                                                               float *__restrict__ targets_tr,
                                                               float *__restrict__ targets_rot) {
    const double ctr[3] = {center ? center[0] : 0.0, center ? center[1] : 0.0, center ? center[2] : 0.0};
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < T; t += stride) {
        const float2 *p = reinterpret_cast<const float2 *>(pairs + 6 * t);  // 24 B per tuple, 8 B aligned
        const float2 p0 = p[0], p1 = p[1], p2 = p[2];
        const float a[3] = {p0.x, p0.y, p1.x}, b[3] = {p1.y, p2.x, p2.y};
        float tr[2], rot[3];
        targets_of_pair(a, b, ctr, ax, tr, rot, targets_rot != nullptr);
        if (targets_tr) reinterpret_cast<float2 *>(targets_tr)[t] = make_float2(tr[0], tr[1]);
        if (targets_rot) {
            targets_rot[3 * t] = rot[0];
            targets_rot[3 * t + 1] = rot[1];
            targets_rot[3 * t + 2] = rot[2];
        }
    }
}

__global__ void __launch_bounds__(256) decode_targets_kernel(const float *__restrict__ pc, IdxView idx,
                                                             const uint8_t *__restrict__ bins, int64_t T, int num_bins,
                                                             Axes ax, float *__restrict__ targets_tr,
                                                             float *__restrict__ targets_rot,
                                                             float *__restrict__ pair_scale,
                                                             float *__restrict__ scaled_out) {
    decode_targets_body(pc, idx, bins, T, num_bins, ax, targets_tr, targets_rot, pair_scale, scaled_out, blockIdx.x, gridDim.x);
}

// ---------------------------------------------------------------------------------------------------
// softmax + one multinomial draw per row: one warp per row of `num_bins` (<= 32) logits, lane = bin,
// coalesced 128 B row reads, softmax and inclusive CDF by shuffles, draw = #(cdf <= u*total).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_bins_kernel(const float *__restrict__ logits, int64_t rows, int num_bins,
                                                          const float *__restrict__ u01, uint64_t seed,
                                                          uint8_t *__restrict__ bins) {
    const int lane = lane_id();
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = warp; row < rows; row += n_warps) {
        const float v = lane < num_bins ? logits[row * num_bins + lane] : -INFINITY;
        float m = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float e = lane < num_bins ? expf(v - m) : 0.0f;
        float cdf = e;  // inclusive scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float up = __shfl_up_sync(0xffffffffu, cdf, o);
            if (lane >= o) cdf += up;
        }
        const float total = __shfl_sync(0xffffffffu, cdf, 31);
        const float u = u01 ? u01[row] : uniform_from_counter(seed, static_cast<uint64_t>(row));
        const uint32_t below = __ballot_sync(0xffffffffu, lane < num_bins && cdf <= u * total);
        if (lane == 0) {
            int b = __popc(below);
            bins[row] = static_cast<uint8_t>(b < num_bins ? b : num_bins - 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Tuple sampling (eval.py:207: np.random.randint(0, N, (T, K)), with replacement): one counter-based
// 64-bit draw per index, mapped to [0, n) by multiply-shift.  Host-sampled indices remain injectable
// everywhere; this is the default when the caller passes none, and saves the 4*T*K-byte upload.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_tuples_kernel(int64_t n, int64_t count, uint64_t seed, int32_t *__restrict__ idx) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (static_cast<uint64_t>(i) + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        idx[i] = static_cast<int32_t>(__umul64hi(z, static_cast<uint64_t>(n)));
    }
}

// Batched frame path (frame.cuh): the tuple indices of every instance that brought none (eval.py:207), one launch.  Same
// generator and counters as sample_tuples_kernel with the instance's seed, so cppf_sample_tuples reproduces them.
__global__ void __launch_bounds__(256) frame_sample_tuples_kernel(const FrameTable *__restrict__ t) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.y) >= t->n_inst) return;
    const FrameInst &in = t->inst[blockIdx.y];
    if (in.idx_draw == nullptr) return;
    const int64_t count = static_cast<int64_t>(in.T) * 5;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
        uint64_t z = in.seed_idx + 0x9E3779B97F4A7C15ull * (static_cast<uint64_t>(i) + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        in.idx_draw[i] = static_cast<int32_t>(__umul64hi(z, static_cast<uint64_t>(in.n)));
    }
}

int frame_launch_sample_tuples(const FrameTable *t, int ni, int64_t T_cap, cudaStream_t s) {
    if (ni <= 0 || T_cap <= 0) return CPPF_OK;
    const int per_inst = std::max(1, std::min<int>(div_up(T_cap * 5, 256), (device_info().sm_count * 8 + ni - 1) / ni));
    CPPF_CUDA_TRY(launch_frame_kernel(frame_sample_tuples_kernel, dim3(per_inst, ni), dim3(256), 0, s, t));
    return CPPF_OK;
}

}  // namespace cppf

using namespace cppf;

CPPF_API int cppf_sample_tuples(int64_t n, int64_t T, int arity, uint64_t seed, int32_t *idx, void *stream) {
    if (!idx || n < 1 || n > 0x7fffffff || T < 0 || arity < 1) return CPPF_ERR_INVALID_ARGUMENT;
    if (T == 0) return CPPF_OK;
    const int64_t count = T * arity;
    sample_tuples_kernel<<<grid_for(count, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, count, seed, idx);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

static Axes axes_from_host(const double *axes_host) {
    Axes ax;
    for (int i = 0; i < 9; ++i) ax.v[i] = axes_host[i];
    return ax;
}

CPPF_API int cppf_sample_bins(const float *logits, int64_t T, int num_bins, const float *u01, uint64_t seed,
                              uint8_t *bins, void *stream) {
    if (!logits || !bins || T < 0 || num_bins < 2 || num_bins > 32) return CPPF_ERR_INVALID_ARGUMENT;
    if (T == 0) return CPPF_OK;
    const int64_t rows = T * 6;
    int blocks = grid_for(rows * 32, 256, 8);
    sample_bins_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, rows, num_bins, u01, seed, bins);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_decode_targets(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride,
                                 const uint8_t *bins, int64_t T, int num_bins, const double *axes_host,
                                 float *targets_tr, float *targets_rot, float *pair_scale, float *pred_pairs_scaled,
                                 void *stream) {
    if (!pc || !idx || !bins || !axes_host || T < 0 || num_bins < 2 || idx_stride < 2) return CPPF_ERR_INVALID_ARGUMENT;
    if (T == 0) return CPPF_OK;
    IdxView iv{idx, idx_stride, idx_is_i64};
    int blocks = grid_for(T, 256, 8);
    decode_targets_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pc, iv, bins, T, num_bins, axes_from_host(axes_host), targets_tr, targets_rot, pair_scale, pred_pairs_scaled);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_generate_targets(const float *pairs, int64_t T, const double *axes_host, const double *center,
                                   float *targets_tr, float *targets_rot, void *stream) {
    if (!pairs || !axes_host || T < 0) return CPPF_ERR_INVALID_ARGUMENT;
    if (T == 0) return CPPF_OK;
    int blocks = grid_for(T, 256, 8);
    generate_targets_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(pairs, T, axes_from_host(axes_host),
                                                                                  center, targets_tr, targets_rot);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}
