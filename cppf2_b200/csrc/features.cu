// Key-point descriptor sampling on B200 -- replaces interpolate_features (dataset.py:40-59): bilinear
// F.grid_sample(align_corners=False, zero padding) of a [C,h,w] patch-token map at N key-points, then L2
// normalisation over the channels (F.normalize, eps 1e-12).  One warp per key-point; the token map is read through
// arbitrary element strides, so the reference's permuted view of the ViT output ([h*w, C] tokens seen as [C,h,w],
// channel stride 1) is read coalesced without a transposing copy.  HBM-bound: 4 neighbour rows in, one row out.
#include "common.cuh"

namespace cppf {

__global__ void __launch_bounds__(256) interpolate_features_kernel(const float *__restrict__ desc, int C, int h, int w, int64_t sc,
                                                                   int64_t sh, int64_t sw, const float *__restrict__ pts, int64_t n,
                                                                   float stride, int normalize, float *__restrict__ out) {
    const int lane = lane_id();
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += n_warps) {
        // dataset.py:46-47: pixel centre -> normalised [-1,1]; grid_sample(align_corners=False): ((g + 1) * size - 1) / 2
        const float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fdiv_rn(__fadd_rn(pts[2 * i], 0.5f), static_cast<float>(w)), stride), 2.0f), 1.0f);
        const float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fdiv_rn(__fadd_rn(pts[2 * i + 1], 0.5f), static_cast<float>(h)), stride), 2.0f), 1.0f);
        const float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), static_cast<float>(w)), 1.0f), 2.0f);
        const float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), static_cast<float>(h)), 1.0f), 2.0f);
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy), x1 = x0 + 1, y1 = y0 + 1;
        // torch's weights (GridSampler.h): nw = (x1 - ix)(y1 - iy), ne = (ix - x0)(y1 - iy), sw = (x1 - ix)(iy - y0), se = (ix - x0)(iy - y0)
        const float ex = __fsub_rn(__fadd_rn(fx, 1.0f), ix), ey = __fsub_rn(__fadd_rn(fy, 1.0f), iy), tx = __fsub_rn(ix, fx), ty = __fsub_rn(iy, fy);
        const float wnw = __fmul_rn(ex, ey), wne = __fmul_rn(tx, ey), wsw = __fmul_rn(ex, ty), wse = __fmul_rn(tx, ty);
        const bool vx0 = x0 >= 0 && x0 < w, vx1 = x1 >= 0 && x1 < w, vy0 = y0 >= 0 && y0 < h, vy1 = y1 >= 0 && y1 < h;
        const float *pnw = desc + y0 * sh + x0 * sw, *pne = desc + y0 * sh + x1 * sw, *psw = desc + y1 * sh + x0 * sw, *pse = desc + y1 * sh + x1 * sw;
        float ss = 0.0f;
        float *o = out + i * C;
        for (int c = lane; c < C; c += 32) {
            float v = 0.0f;
            if (vy0 && vx0) v += pnw[c * sc] * wnw;
            if (vy0 && vx1) v += pne[c * sc] * wne;
            if (vy1 && vx0) v += psw[c * sc] * wsw;
            if (vy1 && vx1) v += pse[c * sc] * wse;
            o[c] = v;
            ss += v * v;
        }
        if (normalize) {
            ss = warp_sum(ss);
            const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
            __syncwarp();
            for (int c = lane; c < C; c += 32) o[c] *= inv;
        }
    }
}

}  // namespace cppf

using namespace cppf;

CPPF_API int cppf_interpolate_features(const float *desc, int C, int h, int w, int64_t stride_c, int64_t stride_h, int64_t stride_w,
                                       const float *pts, int64_t n, float stride, int normalize, float *out, void *stream) {
    if (!desc || !pts || !out || C <= 0 || h <= 0 || w <= 0 || n < 0 || !(stride > 0.0f)) return CPPF_ERR_INVALID_ARGUMENT;
    if (n == 0) return CPPF_OK;
    interpolate_features_kernel<<<grid_for(n * 32, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        desc, C, h, w, stride_c, stride_h, stride_w, pts, n, stride, normalize, out);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}
