// Rotation voting on B200 -- replaces vote_rotation (train_dino.py:218-239), get_topk_dir
// (eval.py:37-51) and the fibonacci-sphere binning they feed (utils/util.py:191-207).
//
// The reference materialises [M,R,3] candidate directions, multiplies them with all S=720 sphere
// points (a K=3 GEMM per 100k rows), thresholds at cos(2 deg) and column-sums hits / weight.  Here the
// candidates never leave registers and each one is tested only against the latitude band of the
// Fibonacci lattice it can possibly hit: sphere point i sits at y_i = 1 - 2i/(S-1), and
// dot(p, s) > cos_thr implies |p.y - s.y| <= |p - s| <= sqrt(2 - 2 cos_thr), i.e. at most
// `band` = ceil(chord*(S-1)/2)+2 indices either side (29 points instead of 720 at the default
// tolerance).  Inside the band the test is the reference's: float32 dot, strict '>', hit adds
// 1/weight in float64.  Bins are float64 in shared memory (720 x 8 B per angle column), flushed with
// one global atomicAdd(double) per non-zero bin per CTA.
#include "common.cuh"

namespace cppf {

struct RotFrame {
    float ab[3], x[3], y[3], tn, sg;
};

// train_dino.py:219-235; tan evaluated in double and rounded (see oracle/cppf_oracle.c).
__device__ __forceinline__ bool rotation_frame(const float a[3], const float b[3], float theta, RotFrame &f) {
    if (!pair_frame(a, b, true, f.ab, f.x)) return false;
    cross_torch(f.x, f.ab, f.y);
    f.tn = static_cast<float>(tan(static_cast<double>(theta)));
    f.sg = f.tn > 0.0f ? 1.0f : -1.0f;  // torch.where(tan > 0, 1., -1.)
    return true;
}

__device__ __forceinline__ void rotation_candidate(const RotFrame &f, float cr, float sr, float up[3]) {
    float v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float off = __fadd_rn(__fmul_rn(cr, f.x[k]), __fmul_rn(sr, f.y[k]));
        v[k] = __fadd_rn(__fmul_rn(f.tn, off), __fmul_rn(f.sg, f.ab[k]));
    }
    float nv = norm3_torch(v[0], v[1], v[2]);
    nv = nv < 1e-7f ? 1e-7f : nv;
    up[0] = __fdiv_rn(v[0], nv);
    up[1] = __fdiv_rn(v[1], nv);
    up[2] = __fdiv_rn(v[2], nv);
}

// Tests direction p against the lattice band and adds w to every bin it hits.
__device__ __forceinline__ void band_vote(const float p[3], double w, const float *__restrict__ s_sphere, int S,
                                          float cos_thr, int band, float half_sm1, double *__restrict__ bins) {
    int lo = 0, hi = S - 1;
    if (band < S) {
        const int ic = __float2int_rn((1.0f - p[1]) * half_sm1);
        lo = max(ic - band, 0);
        hi = min(ic + band, S - 1);
    }
    for (int i = lo; i <= hi; ++i) {
        const float d = __fmaf_rn(p[2], s_sphere[3 * i + 2], __fmaf_rn(p[1], s_sphere[3 * i + 1], __fmul_rn(p[0], s_sphere[3 * i])));
        if (d > cos_thr) atomicAdd(&bins[i], w);
    }
}

// ---- materialising vote_rotation (drop-in shim) ----------------------------------------------------
__global__ void __launch_bounds__(256) vote_rotation_kernel(const float *__restrict__ pc, IdxView idx,
                                                            const float *__restrict__ preds_rot, int64_t M,
                                                            const float *__restrict__ cos_tab,
                                                            const float *__restrict__ sin_tab, int R,
                                                            float *__restrict__ up, uint8_t *__restrict__ mask) {
    const int lane = lane_id();
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t m = warp; m < M; m += n_warps) {
        const int64_t ia = idx.at(m, 0), ib = idx.at(m, 1);
        const float a[3] = {pc[3 * ia], pc[3 * ia + 1], pc[3 * ia + 2]};
        const float b[3] = {pc[3 * ib], pc[3 * ib + 1], pc[3 * ib + 2]};
        RotFrame f;
        const bool ok = rotation_frame(a, b, preds_rot[m], f);
        if (lane == 0) mask[m] = ok ? 1 : 0;
        float *dst = up + m * R * 3;
        for (int r = lane; r < R; r += 32) {
            float v[3] = {0.f, 0.f, 0.f};
            if (ok) rotation_candidate(f, cos_tab[r], sin_tab[r], v);
            dst[3 * r] = v[0];
            dst[3 * r + 1] = v[1];
            dst[3 * r + 2] = v[2];
        }
    }
}

// ---- get_topk_dir histogram over explicit rows -----------------------------------------------------
__global__ void __launch_bounds__(256) sphere_hist_kernel(const float *__restrict__ pred, int64_t rows,
                                                          const double *__restrict__ wt,
                                                          const float *__restrict__ sphere, int S, float cos_thr,
                                                          int band, double *__restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_bins = reinterpret_cast<double *>(smem_raw);
    float *s_sphere = reinterpret_cast<float *>(s_bins + S);
    for (int i = threadIdx.x; i < S; i += blockDim.x) s_bins[i] = 0.0;
    for (int i = threadIdx.x; i < 3 * S; i += blockDim.x) s_sphere[i] = sphere[i];
    __syncthreads();
    const float half_sm1 = 0.5f * static_cast<float>(S - 1);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows; i += stride) {
        const float p[3] = {pred[3 * i], pred[3 * i + 1], pred[3 * i + 2]};
        const double w = wt ? __drcp_rn(wt[i]) : 1.0;
        band_vote(p, w, s_sphere, S, cos_thr, band, half_sm1, s_bins);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += blockDim.x)
        if (s_bins[i] != 0.0) atomicAdd(&counts[i], s_bins[i]);
}

// ---- fused vote_rotation + get_topk_dir over kept tuples --------------------------------------------
constexpr int kMaxTheta = 3;

struct ThetaCols {
    int col[kMaxTheta];
    int n;
};

__global__ void __launch_bounds__(256) rotation_hist_kernel(
    const float *__restrict__ pc, IdxView idx, const float *__restrict__ theta, int64_t theta_stride, ThetaCols cols,
    const int32_t *__restrict__ kept_list, const int64_t *__restrict__ kept_count, int64_t M,
    const int32_t *__restrict__ imp, const cppf_backvote_summary *__restrict__ summary, double margin,
    const float *__restrict__ cos_tab, const float *__restrict__ sin_tab, int R, const float *__restrict__ sphere, int S,
    float cos_thr, int band, double *__restrict__ counts, int part, int n_parts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_bins = reinterpret_cast<double *>(smem_raw);                   // [n_theta][S]
    float *s_sphere = reinterpret_cast<float *>(s_bins + cols.n * S);        // [S][3]
    float *s_cos = s_sphere + 3 * S;                                         // [R]
    float *s_sin = s_cos + R;
    for (int i = threadIdx.x; i < cols.n * S; i += blockDim.x) s_bins[i] = 0.0;
    for (int i = threadIdx.x; i < 3 * S; i += blockDim.x) s_sphere[i] = sphere[i];
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
        s_cos[i] = cos_tab[i];
        s_sin[i] = sin_tab[i];
    }
    __syncthreads();
    const int64_t n_items = kept_list ? (kept_count ? *kept_count : M) : M;
    const double imp_max = (imp && summary) ? static_cast<double>(summary->imp_max) : 1.0;
    const float half_sm1 = 0.5f * static_cast<float>(S - 1);
    const int lane = lane_id();
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    // tuple-sharded runs give every rank the kept items congruent to `part` modulo `n_parts`
    for (int64_t it = warp * n_parts + part; it < n_items; it += n_warps * n_parts) {
        const int64_t m = kept_list ? static_cast<int64_t>(kept_list[it]) : it;
        const int64_t ia = idx.at(m, 0), ib = idx.at(m, 1);
        const float a[3] = {pc[3 * ia], pc[3 * ia + 1], pc[3 * ia + 2]};
        const float b[3] = {pc[3 * ib], pc[3 * ib + 1], pc[3 * ib + 2]};
        double w = 1.0;
        if (imp) {  // imp_wt = imp / imp.max(); pair weight = imp_wt[i] + imp_wt[j] + margin (eval.py:274-275)
            const double wi = __ddiv_rn(static_cast<double>(imp[ia]), imp_max);
            const double wj = __ddiv_rn(static_cast<double>(imp[ib]), imp_max);
            w = __drcp_rn(__dadd_rn(__dadd_rn(wi, wj), margin));
        }
        for (int c = 0; c < cols.n; ++c) {
            RotFrame f;
            if (!rotation_frame(a, b, theta[m * theta_stride + cols.col[c]], f)) break;  // warp-uniform
            double *bins = s_bins + c * S;
            for (int r = lane; r < R; r += 32) {
                float p[3];
                rotation_candidate(f, s_cos[r], s_sin[r], p);
                band_vote(p, w, s_sphere, S, cos_thr, band, half_sm1, bins);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cols.n * S; i += blockDim.x)
        if (s_bins[i] != 0.0) atomicAdd(&counts[i], s_bins[i]);
}

}  // namespace cppf

using namespace cppf;

CPPF_API int cppf_sphere_band(int S, float cos_thr) {
    if (S < 2) return S;
    double chord = sqrt(fmax(0.0, 2.0 - 2.0 * static_cast<double>(cos_thr)));
    int band = static_cast<int>(ceil(chord * (S - 1) * 0.5)) + 2;
    return band;
}

CPPF_API int cppf_vote_rotation(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride,
                                const float *preds_rot, int64_t M, const float *cos_tab, const float *sin_tab, int R,
                                float *up, uint8_t *mask, void *stream) {
    if (!pc || !idx || !preds_rot || !cos_tab || !sin_tab || !up || !mask || M < 0 || R <= 0 || idx_stride < 2)
        return CPPF_ERR_INVALID_ARGUMENT;
    if (M == 0) return CPPF_OK;
    IdxView iv{idx, idx_stride, idx_is_i64};
    vote_rotation_kernel<<<grid_for(M * 32, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pc, iv, preds_rot, M, cos_tab, sin_tab, R, up, mask);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_sphere_hist(const float *pred, int64_t rows, const double *wt, const float *sphere, int S,
                              float cos_thr, int band, double *counts, void *stream) {
    if (!pred || !sphere || !counts || rows < 0 || S < 1) return CPPF_ERR_INVALID_ARGUMENT;
    if (rows == 0) return CPPF_OK;
    const size_t smem = static_cast<size_t>(S) * (sizeof(double) + 3 * sizeof(float));
    if (smem > static_cast<size_t>(device_info().max_smem_optin)) return CPPF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        CPPF_CUDA_TRY(cudaFuncSetAttribute(sphere_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    sphere_hist_kernel<<<grid_for(rows, 256, 4), 256, smem, static_cast<cudaStream_t>(stream)>>>(pred, rows, wt, sphere, S,
                                                                                               cos_thr, band, counts);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_rotation_hist_part(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride, const float *theta,
                                     int64_t theta_stride, const int *theta_cols_host, int n_theta, const int32_t *kept_list,
                                     const int64_t *kept_count, int64_t M, const int32_t *imp,
                                     const cppf_backvote_summary *summary, double margin, const float *cos_tab,
                                     const float *sin_tab, int R, const float *sphere, int S, float cos_thr, int band,
                                     double *counts, int part, int n_parts, void *stream) {
    if (!pc || !idx || !theta || !theta_cols_host || !cos_tab || !sin_tab || !sphere || !counts)
        return CPPF_ERR_INVALID_ARGUMENT;
    if (n_parts < 1 || part < 0 || part >= n_parts) return CPPF_ERR_INVALID_ARGUMENT;
    if (n_theta < 1 || n_theta > kMaxTheta || M < 0 || R <= 0 || S < 1 || idx_stride < 2) return CPPF_ERR_INVALID_ARGUMENT;
    if (imp && !summary) return CPPF_ERR_INVALID_ARGUMENT;
    if (M == 0) return CPPF_OK;
    ThetaCols cols;
    cols.n = n_theta;
    for (int i = 0; i < kMaxTheta; ++i) cols.col[i] = i < n_theta ? theta_cols_host[i] : 0;
    const size_t smem = static_cast<size_t>(n_theta) * S * sizeof(double) + 3 * static_cast<size_t>(S) * sizeof(float) +
                        2 * static_cast<size_t>(R) * sizeof(float);
    if (smem > static_cast<size_t>(device_info().max_smem_optin)) return CPPF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        CPPF_CUDA_TRY(cudaFuncSetAttribute(rotation_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    IdxView iv{idx, idx_stride, idx_is_i64};
    // one warp per kept tuple; the kept count is only known on the device, so size for M/8 (ratio 0.1) at least
    const int64_t guess = (kept_list ? (M / 8 + 1) : M) / n_parts + 1;
    rotation_hist_kernel<<<grid_for(guess * 32, 256, 4), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        pc, iv, theta, theta_stride, cols, kept_list, kept_count, M, imp, summary, margin, cos_tab, sin_tab, R, sphere, S,
        cos_thr, band, counts, part, n_parts);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_rotation_hist(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride, const float *theta,
                                int64_t theta_stride, const int *theta_cols_host, int n_theta, const int32_t *kept_list,
                                const int64_t *kept_count, int64_t M, const int32_t *imp,
                                const cppf_backvote_summary *summary, double margin, const float *cos_tab,
                                const float *sin_tab, int R, const float *sphere, int S, float cos_thr, int band,
                                double *counts, void *stream) {
    return cppf_rotation_hist_part(pc, idx, idx_is_i64, idx_stride, theta, theta_stride, theta_cols_host, n_theta, kept_list,
                                   kept_count, M, imp, summary, margin, cos_tab, sin_tab, R, sphere, S, cos_thr, band, counts,
                                   0, 1, stream);
}
