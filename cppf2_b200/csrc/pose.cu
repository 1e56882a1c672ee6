// Pose assembly on B200 -- replaces eval.py:284-313 (top-1 directions, Gram-Schmidt, third axis by
// cross product, lower-median scale of the kept tuples) and eval.py:358-363 (branch selection loss).
// Three tiny stream-ordered kernels; nothing returns to the host until the caller reads cppf_pose.
#include "common.cuh"

namespace cppf {

struct PoseScratch {  // head of the workspace, zeroed by the host wrapper
    double loss_sum;
    unsigned long long loss_cnt;
    unsigned int ticket;
    unsigned int pad;
};

// ---- directions + rotation matrix -------------------------------------------------------------------
__global__ void __launch_bounds__(256) pose_directions_kernel(const double *__restrict__ counts,
                                                              const float *__restrict__ sphere, int S,
                                                              const cppf_center *__restrict__ center,
                                                              const cppf_backvote_summary *__restrict__ summary,
                                                              int up_loc, int right_loc, cppf_pose *__restrict__ pose) {
    // first arg-max of float32(counts) per angle column: the reference accumulates in float32 (eval.py:39)
    __shared__ float s_val[2][8];
    __shared__ int s_idx[2][8];
    __shared__ int s_best[2];
    __shared__ float s_bestv[2];
    for (int c = 0; c < 2; ++c) {
        float bv = -1.0f;
        int bi = 0x7fffffff;
        for (int i = threadIdx.x; i < S; i += blockDim.x) {
            const float v = static_cast<float>(counts[c * S + i]);
            if (v > bv || (v == bv && i < bi)) {
                bv = v;
                bi = i;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) {
                bv = ov;
                bi = oi;
            }
        }
        if (lane_id() == 0) {
            s_val[c][threadIdx.x >> 5] = bv;
            s_idx[c][threadIdx.x >> 5] = bi;
        }
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int c = threadIdx.x;
        float bv = s_val[c][0];
        int bi = s_idx[c][0];
        for (int w = 1; w < 8; ++w)
            if (s_val[c][w] > bv || (s_val[c][w] == bv && s_idx[c][w] < bi)) {
                bv = s_val[c][w];
                bi = s_idx[c][w];
            }
        s_best[c] = bi;
        s_bestv[c] = bv;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int bu = s_best[0], br = s_best[1];
    float up[3] = {sphere[3 * bu], sphere[3 * bu + 1], sphere[3 * bu + 2]};
    float rt[3] = {sphere[3 * br], sphere[3 * br + 1], sphere[3 * br + 2]};
    // preds_right -= np.dot(up, right) * up ; preds_right /= (norm + 1e-9)   (float32 numpy, eval.py:295-296)
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(up[0], rt[0]), __fmul_rn(up[1], rt[1])), __fmul_rn(up[2], rt[2]));
    for (int k = 0; k < 3; ++k) rt[k] = __fsub_rn(rt[k], __fmul_rn(d, up[k]));
    const float nr = __fadd_rn(norm3_numpy(rt[0], rt[1], rt[2]), 1e-9f);
    for (int k = 0; k < 3; ++k) rt[k] = __fdiv_rn(rt[k], nr);
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 3; ++k) {
        R[3 * k + up_loc] = static_cast<double>(up[k]);
        R[3 * k + right_loc] = static_cast<double>(rt[k]);
    }
    const int other = 3 - up_loc - right_loc;
    const int c1 = (other + 1) % 3, c2 = (other + 2) % 3;  // R[:,other] = cross(R[:,other+1], R[:,other+2])
    const double a0 = R[c1], a1 = R[3 + c1], a2 = R[6 + c1], b0 = R[c2], b1 = R[3 + c2], b2 = R[6 + c2];
    R[other] = __dsub_rn(__dmul_rn(a1, b2), __dmul_rn(a2, b1));
    R[3 + other] = __dsub_rn(__dmul_rn(a2, b0), __dmul_rn(a0, b2));
    R[6 + other] = __dsub_rn(__dmul_rn(a0, b1), __dmul_rn(a1, b0));
    for (int i = 0; i < 9; ++i) pose->R[i] = R[i];
    for (int k = 0; k < 3; ++k) pose->t[k] = center->world[k];
    pose->bin_up = bu;
    pose->bin_right = br;
    pose->count_up = s_bestv[0];
    pose->count_right = s_bestv[1];
    pose->kept = summary ? summary->kept : 0;
    pose->status = (summary && summary->kept == 0) ? CPPF_STATUS_EMPTY : 0u;
}

// ---- lower median of the kept scale predictions, one CTA per axis ------------------------------------
__global__ void __launch_bounds__(1024) scale_median_kernel(const float *__restrict__ pred_scales,
                                                            const int32_t *__restrict__ kept_list,
                                                            const cppf_backvote_summary *__restrict__ summary,
                                                            const float *__restrict__ scale_override,
                                                            cppf_pose *__restrict__ pose) {
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_prefix;
    __shared__ unsigned long long s_k;
    const int axis = blockIdx.x;
    const int64_t M = summary->kept;
    if (scale_override || M <= 0) {
        if (threadIdx.x == 0) pose->scale[axis] = scale_override ? scale_override[axis] : 0.0f;
    } else {
        if (threadIdx.x == 0) {
            s_prefix = 0u;
            s_k = static_cast<unsigned long long>((M - 1) / 2);  // torch.median: lower middle element
        }
        for (int pass = 0; pass < 4; ++pass) {
            if (threadIdx.x < 256) s_hist[threadIdx.x] = 0u;
            __syncthreads();
            const int shift = 24 - 8 * pass;
            const uint32_t decided = pass == 0 ? 0u : ~((1u << (shift + 8)) - 1u);
            const uint32_t prefix = s_prefix;
            for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
                const uint32_t key = float_to_key(pred_scales[3 * static_cast<int64_t>(kept_list[i]) + axis]);
                if ((key & decided) == (prefix & decided)) atomicAdd(&s_hist[(key >> shift) & 0xffu], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long k = s_k, cum = 0;
                int d = 0;
                for (; d < 256; ++d) {
                    if (cum + s_hist[d] > k) break;
                    cum += s_hist[d];
                }
                if (d == 256) d = 255;
                s_prefix = prefix | (static_cast<uint32_t>(d) << shift);
                s_k = k - cum;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) pose->scale[axis] = key_to_float(s_prefix);
    }
}

// ---- branch loss: mean clip(|canon(pc[pair]) - pred_pairs|, 0, 0.1) over kept pairs (eval.py:358-363) -
__global__ void __launch_bounds__(256) pose_loss_kernel(const float *__restrict__ pc, IdxView idx,
                                                        const uint8_t *__restrict__ bins, int num_bins,
                                                        const int32_t *__restrict__ kept_list,
                                                        const cppf_backvote_summary *__restrict__ summary,
                                                        int loss_y_only, cppf_pose *__restrict__ pose,
                                                        PoseScratch *__restrict__ scratch) {
    __shared__ double s_sum[8];
    __shared__ bool s_last;
    const int64_t M = summary->kept;
    // scale norm: np.linalg.norm(pred_scale) on float32 -> float32 (eval.py:310)
    const float sx = pose->scale[0], sy = pose->scale[1], sz = pose->scale[2];
    const float sn = norm3_numpy(sx, sy, sz);
    const double snd = static_cast<double>(sn);
    double R[9], t[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = pose->R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = pose->t[i];
    const float denom = static_cast<float>(num_bins - 1);
    double acc = 0.0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < 2 * M; i += stride) {
        const int64_t m = kept_list[i >> 1];
        const int end = static_cast<int>(i & 1);
        const int64_t ip = idx.at(m, end);
        const double d0 = static_cast<double>(pc[3 * ip]) - t[0], d1 = static_cast<double>(pc[3 * ip + 1]) - t[1],
                     d2 = static_cast<double>(pc[3 * ip + 2]) - t[2];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (loss_y_only && k != 1) continue;
            const double canon = (d0 * R[k] + d1 * R[3 + k] + d2 * R[6 + k]) / snd;
            const float pred = __fsub_rn(__fdiv_rn(static_cast<float>(bins[6 * m + 3 * end + k]), denom), 0.5f);
            double l = fabs(canon - static_cast<double>(pred));
            l = l > 0.1 ? 0.1 : l;
            acc += l;
        }
    }
    acc = warp_sum(acc);
    if (lane_id() == 0) s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) acc += s_sum[w];
        atomicAdd(&scratch->loss_sum, acc);
        __threadfence();
        s_last = (atomicAdd(&scratch->ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const double total = *reinterpret_cast<volatile double *>(&scratch->loss_sum);
        const double cnt = static_cast<double>(2 * M) * (loss_y_only ? 1.0 : 3.0);
        pose->loss = M > 0 ? total / cnt : INFINITY;
        pose->scale_norm = sn;
    }
}

}  // namespace cppf

using namespace cppf;

CPPF_API int64_t cppf_pose_workspace_bytes(int64_t T) {
    (void)T;
    return 256;
}

CPPF_API int cppf_pose_finalize(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride,
                                const uint8_t *bins, int num_bins, const float *pred_scales, const int32_t *kept_list,
                                const cppf_backvote_summary *summary, const double *counts, const float *sphere, int S,
                                const cppf_center *center, int up_loc, int right_loc, int loss_y_only,
                                const float *scale_override, cppf_pose *pose, void *ws, int64_t ws_bytes, void *stream) {
    if (!pc || !idx || !bins || !kept_list || !summary || !counts || !sphere || !center || !pose || !ws)
        return CPPF_ERR_INVALID_ARGUMENT;
    if (!pred_scales && !scale_override) return CPPF_ERR_INVALID_ARGUMENT;
    if (up_loc < 0 || up_loc > 2 || right_loc < 0 || right_loc > 2 || up_loc == right_loc || S < 1 || idx_stride < 2)
        return CPPF_ERR_INVALID_ARGUMENT;
    if (ws_bytes < cppf_pose_workspace_bytes(0)) return CPPF_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PoseScratch *scratch = static_cast<PoseScratch *>(ws);
    CPPF_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(PoseScratch), s));
    pose_directions_kernel<<<1, 256, 0, s>>>(counts, sphere, S, center, summary, up_loc, right_loc, pose);
    CPPF_LAUNCH_CHECK();
    scale_median_kernel<<<3, 1024, 0, s>>>(pred_scales, kept_list, summary, scale_override, pose);
    CPPF_LAUNCH_CHECK();
    IdxView iv{idx, idx_stride, idx_is_i64};
    pose_loss_kernel<<<device_info().sm_count, 256, 0, s>>>(pc, iv, bins, num_bins, kept_list, summary, loss_y_only, pose,
                                                           scratch);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}
