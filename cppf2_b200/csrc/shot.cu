// SHOT-352 descriptors and normals on B200 -- replaces shot.compute / shot.estimate_normal
// (reference src_shot/shot.cpp:12-42, :45-100: PCL NormalEstimation + SHOTEstimation<SHOT352> with the
// default SHOT local reference frame, radius search, viewpoint at the origin).
//
// Search structure: instead of three kd-trees (normals, LRF, SHOT) one uniform grid with cell edge >=
// radius is built on the device: bounds -> cell histogram -> exclusive scan -> scatter into a cell-sorted
// float4 copy (x, y, z, original index).  A radius query then reads the 27 surrounding cells, which are 9
// contiguous runs of the sorted array (the 3 z-neighbours of a cell are adjacent), with coalesced 16-byte
// loads; queries are processed in sorted order, one warp per key-point, so consecutive warps hit the same
// runs in L1.  Membership uses FLANN's float arithmetic exactly (((dx*dx)+dy*dy)+dz*dz < float(r*r)),
// so the neighbour SETS are identical to the CPU restatement; only accumulation order differs.
//
// Per key-point:  normals (one warp): 9 float sums + count, warp-shuffle reduction, pcl::eigen33 closed
// form in float, flip towards the origin.  LRF in three stages: weighted scatter matrix in double over the
// compacted neighbour list (one warp), cyclic Jacobi (one THREAD: the solve is scalar work), sign votes (one
// warp, in the descriptor kernel).  Histogram: every neighbour's quadrilinear contributions go to a per-warp
// 352-bin histogram in shared memory (32-bit fixed point, native integer atomics: exact sums), then
// L2-normalised and written as one coalesced 1408-byte row.
#include "common.cuh"
#include "frame.cuh"

#include <math_constants.h>

namespace cppf {

constexpr int kShotWarps = 8;                    // key-points in flight per CTA
constexpr int kShotMaxCells = 1 << 21;           // cell table capacity (coarsened beyond that)
constexpr int kShotListCap = 768;                // cached neighbours per key-point (more: the cells are swept again)

__device__ __forceinline__ int shot_coord(float v, float lo, float inv, int dim) {
    int c = static_cast<int>((v - lo) * inv);
    return min(max(c, 0), dim - 1);
}

// one thread
__device__ __forceinline__ void shot_grid_setup(const cppf_grid_geom *__restrict__ bounds, float radius, ShotGrid *__restrict__ g) {
    float inv = 1.0f / radius;
    int dim[3];
    long long cells;
    for (;;) {
        cells = 1;
        for (int k = 0; k < 3; ++k) {
            float ext = bounds->hi[k] - bounds->lo[k];
            dim[k] = max(1, static_cast<int>(ext * inv) + 1);
            cells *= dim[k];
        }
        if (cells <= kShotMaxCells) break;
        inv *= 0.5f;  // coarser cells keep the 27-cell stencil valid (edge stays >= radius)
    }
    for (int k = 0; k < 3; ++k) {
        g->lo[k] = bounds->lo[k];
        g->dim[k] = dim[k];
    }
    g->inv = inv;
    g->cells = static_cast<int>(cells);
}

__global__ void shot_grid_setup_kernel(const cppf_grid_geom *__restrict__ bounds, float radius, ShotGrid *__restrict__ g) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    shot_grid_setup(bounds, radius, g);
}

__device__ __forceinline__ void shot_cell_count_one(const float *__restrict__ pc, int i, const ShotGrid *__restrict__ g,
                                                    int *__restrict__ cell_of, int *__restrict__ cell_count) {
    const float x = pc[3 * i], y = pc[3 * i + 1], z = pc[3 * i + 2];
    int cell = -1;
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
        cell = (shot_coord(x, g->lo[0], g->inv, g->dim[0]) * g->dim[1] + shot_coord(y, g->lo[1], g->inv, g->dim[1])) * g->dim[2] +
               shot_coord(z, g->lo[2], g->inv, g->dim[2]);
        atomicAdd(&cell_count[cell], 1);
    }
    cell_of[i] = cell;
}

__global__ void __launch_bounds__(256) shot_cell_count_kernel(const float *__restrict__ pc, int n,
                                                              const ShotGrid *__restrict__ g, int *__restrict__ cell_of,
                                                              int *__restrict__ cell_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    shot_cell_count_one(pc, i, g, cell_of, cell_count);
}

// exclusive scan of cell_count[0..cells) into cell_start[0..cells], single CTA
__device__ __forceinline__ void shot_cell_scan_body(const ShotGrid *__restrict__ g, const int *__restrict__ cell_count,
                                                    int *__restrict__ cell_start, int *__restrict__ cell_fill) {
    __shared__ int s_part[1024];
    const int cells = g->cells;
    const int per = (cells + 1023) / 1024;
    const int b = threadIdx.x * per, e = min(b + per, cells);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += cell_count[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan of the partials
        int v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = threadIdx.x ? s_part[threadIdx.x - 1] : 0;
    for (int i = b; i < e; ++i) {
        cell_start[i] = run;
        cell_fill[i] = run;
        run += cell_count[i];
    }
    if (threadIdx.x == 1023) cell_start[cells] = s_part[1023];
}

__global__ void __launch_bounds__(1024) shot_cell_scan_kernel(const ShotGrid *__restrict__ g, const int *__restrict__ cell_count,
                                                              int *__restrict__ cell_start, int *__restrict__ cell_fill) {
    shot_cell_scan_body(g, cell_count, cell_start, cell_fill);
}

__device__ __forceinline__ void shot_scatter_one(const float *__restrict__ pc, int i, const int *__restrict__ cell_of,
                                                 int *__restrict__ cell_fill, float4 *__restrict__ sorted) {
    const int cell = cell_of[i];
    if (cell < 0) return;
    const int pos = atomicAdd(&cell_fill[cell], 1);
    sorted[pos] = make_float4(pc[3 * i], pc[3 * i + 1], pc[3 * i + 2], __int_as_float(i));
}

__global__ void __launch_bounds__(256) shot_scatter_kernel(const float *__restrict__ pc, int n, const int *__restrict__ cell_of,
                                                           int *__restrict__ cell_fill, float4 *__restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    shot_scatter_one(pc, i, cell_of, cell_fill, sorted);
}

// The atomic scatter leaves each cell's points in arrival order, which changes from run to run.  Rank every
// point of a cell by its original index (one warp per cell, counting sort by comparison: segments hold
// ~100 points when the cloud is voxel-sampled at radius/10 as the reference does) so that the sorted copy,
// and with it every floating-point accumulation order downstream, is deterministic.
constexpr int kOrderStage = 256;                 // keys of one cell staged in shared memory per warp (more: read from L2)

__device__ __forceinline__ void shot_cell_order_body(const ShotGrid *__restrict__ gp, const int *__restrict__ cell_start,
                                                     const float4 *__restrict__ scattered, float4 *__restrict__ sorted,
                                                     int bid, int nblk) {
    // the original indices of the cell's points, staged once: the rank loop then reads shared-memory broadcasts instead of
    // issuing a dependent 16-byte L2 load per comparison
    __shared__ int s_key[8][kOrderStage];
    int *keys = s_key[(threadIdx.x >> 5) & 7];
    const int cells = gp->cells;
    const int lane = lane_id();
    const int warp = (bid * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (nblk * blockDim.x) >> 5;
    for (int c = warp; c < cells; c += n_warps) {
        const int b = cell_start[c], e = cell_start[c + 1];
        const int k = e - b;
        if (k <= 0) continue;
        const bool staged = k <= kOrderStage;
        if (staged)
            for (int u = lane; u < k; u += 32) keys[u] = __float_as_int(scattered[b + u].w);
        __syncwarp();
        for (int t = b + lane; t < e; t += 32) {
            const float4 mine = scattered[t];
            const int key = __float_as_int(mine.w);
            int rank = 0;
            if (staged) {
#pragma unroll 4
                for (int u = 0; u < k; ++u) rank += (keys[u] < key) ? 1 : 0;
            } else {
                for (int u = b; u < e; ++u) rank += (__float_as_int(scattered[u].w) < key) ? 1 : 0;
            }
            sorted[b + rank] = mine;
        }
        __syncwarp();                                 // the staging buffer is rewritten for the warp's next cell
    }
}

__global__ void __launch_bounds__(256) shot_cell_order_kernel(const ShotGrid *__restrict__ gp, const int *__restrict__ cell_start,
                                                              const float4 *__restrict__ scattered,
                                                              float4 *__restrict__ sorted) {
    shot_cell_order_body(gp, cell_start, scattered, sorted, blockIdx.x, gridDim.x);
}

// Points with non-finite coordinates never enter the grid: their outputs are NaN (PCL: isFinite checks).
__global__ void __launch_bounds__(256) shot_nan_fill_kernel(const int *__restrict__ cell_of, int n, float *__restrict__ normals,
                                                            float *__restrict__ desc) {
    const int i = blockIdx.x;
    if (i >= n || cell_of[i] >= 0) return;
    const float nanv = CUDART_NAN_F;
    if (normals && threadIdx.x < 3) normals[3 * i + threadIdx.x] = nanv;
    if (desc)
        for (int j = threadIdx.x; j < CPPF_SHOT_DIM; j += blockDim.x) desc[static_cast<size_t>(i) * CPPF_SHOT_DIM + j] = nanv;
}

// Visits the candidates of the 27-cell stencil around p, lane-strided; f(j, q) is called for every
// candidate slot j of the sorted array with its float4 (the caller applies the radius test).
template <typename F>
__device__ __forceinline__ void for_each_candidate(const ShotGrid &g, const int *__restrict__ cell_start,
                                                   const float4 *__restrict__ sorted, const float p[3], int lane, F &&f) {
    const int cx = shot_coord(p[0], g.lo[0], g.inv, g.dim[0]);
    const int cy = shot_coord(p[1], g.lo[1], g.inv, g.dim[1]);
    const int cz = shot_coord(p[2], g.lo[2], g.inv, g.dim[2]);
    const int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.dim[2] - 1);
    for (int x = max(cx - 1, 0); x <= min(cx + 1, g.dim[0] - 1); ++x)
        for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y) {
            const int row = (x * g.dim[1] + y) * g.dim[2];
            const int b = cell_start[row + z0], e = cell_start[row + z1 + 1];  // z-neighbours are contiguous
#pragma unroll 2
            for (int j = b + lane; j < e; j += 32) f(j, sorted[j]);
        }
}

// FLANN L2_Simple in float: ((dx*dx) + dy*dy) + dz*dz
__device__ __forceinline__ float flann_dist2(const float p[3], const float4 &q) {
    const float dx = __fsub_rn(p[0], q.x), dy = __fsub_rn(p[1], q.y), dz = __fsub_rn(p[2], q.z);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- pcl::computeRoots2 / computeRoots / eigen33 (float; common/impl/eigen.hpp) ------------------------
__device__ __forceinline__ void compute_roots2(float b, float c, float roots[3]) {
    roots[0] = 0.0f;
    float d = static_cast<float>(static_cast<double>(b * b) - 4.0 * static_cast<double>(c));
    if (d < 0.0f) d = 0.0f;
    const float sd = sqrtf(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}

__device__ void compute_roots(const float m[9], float roots[3]) {
    const float c0 = m[0] * m[4] * m[8] + 2.0f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
    const float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
    const float c2 = m[0] + m[4] + m[8];
    if (fabsf(c0) < 1.1920929e-07f) {
        compute_roots2(c2, c1, roots);
        return;
    }
    const float s_inv3 = static_cast<float>(1.0 / 3.0);
    const float s_sqrt3 = sqrtf(3.0f);
    const float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    const float rho = sqrtf(-a_over_3);
    const float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
    float sin_theta, cos_theta;
    sincosf(theta, &sin_theta, &cos_theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    float t;
    if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    if (roots[1] >= roots[2]) {
        t = roots[1]; roots[1] = roots[2]; roots[2] = t;
        if (roots[0] >= roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
}

__device__ __forceinline__ void cross3f(const float a[3], const float b[3], float out[3]) {
    out[0] = a[1] * b[2] - a[2] * b[1];
    out[1] = a[2] * b[0] - a[0] * b[2];
    out[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ void eigen33_smallest(const float mat[9], float evec[3]) {
    float scale = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(mat[i]));
    if (scale <= 1.17549435e-38f) scale = 1.0f;
    float s[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) s[i] = mat[i] / scale;
    float roots[3];
    compute_roots(s, roots);
    s[0] -= roots[0];
    s[4] -= roots[0];
    s[8] -= roots[0];
    float v1[3], v2[3], v3[3];
    cross3f(s + 0, s + 3, v1);
    cross3f(s + 0, s + 6, v2);
    cross3f(s + 3, s + 6, v3);
    const float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
    const float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
    const float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
    const float *v;
    float l;
    if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
    else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
    else { v = v3; l = l3; }
    const float len = sqrtf(l);
    evec[0] = v[0] / len;
    evec[1] = v[1] / len;
    evec[2] = v[2] / len;
}

// ---- normals: NormalEstimation::computePointNormal + flipNormalTowardsViewpoint(origin) ----------------
// normal of the point at sorted position s (one warp): NormalEstimation::computePointNormal + flipNormalTowardsViewpoint
__device__ __forceinline__ void shot_normal_point(const ShotGrid &g, const int *__restrict__ cell_start,
                                                  const float4 *__restrict__ sorted, float radius_sq,
                                                  float *__restrict__ normals, float4 *__restrict__ normals_sorted, int s, int lane) {
    const float4 pq = sorted[s];
    const float p[3] = {pq.x, pq.y, pq.z};
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for_each_candidate(g, cell_start, sorted, p, lane, [&](int, const float4 &q) {
        if (flann_dist2(p, q) < radius_sq) {
            acc[0] += q.x * q.x;
            acc[1] += q.x * q.y;
            acc[2] += q.x * q.z;
            acc[3] += q.y * q.y;
            acc[4] += q.y * q.z;
            acc[5] += q.z * q.z;
            acc[6] += q.x;
            acc[7] += q.y;
            acc[8] += q.z;
            ++cnt;
        }
    });
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = warp_sum(acc[i]);
    cnt = warp_sum(cnt);
    float nrm[3] = {CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F};
    if (cnt >= 3) {
        const float c = static_cast<float>(cnt);
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] /= c;
        float cov[9];
        cov[0] = acc[0] - acc[6] * acc[6];
        cov[1] = acc[1] - acc[6] * acc[7];
        cov[2] = acc[2] - acc[6] * acc[8];
        cov[4] = acc[3] - acc[7] * acc[7];
        cov[5] = acc[4] - acc[7] * acc[8];
        cov[8] = acc[5] - acc[8] * acc[8];
        cov[3] = cov[1];
        cov[6] = cov[2];
        cov[7] = cov[5];
        eigen33_smallest(cov, nrm);
        const float cos_theta = (-p[0]) * nrm[0] + (-p[1]) * nrm[1] + (-p[2]) * nrm[2];
        if (cos_theta < 0.0f) {
            nrm[0] = -nrm[0];
            nrm[1] = -nrm[1];
            nrm[2] = -nrm[2];
        }
    }
    if (lane == 0) {
        const int orig = __float_as_int(pq.w);
        normals[3 * orig] = nrm[0];
        normals[3 * orig + 1] = nrm[1];
        normals[3 * orig + 2] = nrm[2];
        if (normals_sorted) normals_sorted[s] = make_float4(nrm[0], nrm[1], nrm[2], 0.0f);
    }
}

__device__ __forceinline__ void shot_normals_body(const ShotGrid *__restrict__ gp, const int *__restrict__ cell_start,
                                                  const float4 *__restrict__ sorted, float radius_sq,
                                                  float *__restrict__ normals, float4 *__restrict__ normals_sorted, int bid,
                                                  int nblk) {
    const ShotGrid g = *gp;
    const int n_sorted = cell_start[g.cells];
    const int lane = lane_id();
    const int warp = (bid * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (nblk * blockDim.x) >> 5;
    for (int s = warp; s < n_sorted; s += n_warps) shot_normal_point(g, cell_start, sorted, radius_sq, normals, normals_sorted, s, lane);
}

__global__ void __launch_bounds__(kShotWarps * 32) shot_normals_kernel(const ShotGrid *__restrict__ gp,
                                                                      const int *__restrict__ cell_start,
                                                                      const float4 *__restrict__ sorted, int n_sorted_max,
                                                                      float radius_sq, float *__restrict__ normals,
                                                                      float4 *__restrict__ normals_sorted) {
    (void)n_sorted_max;
    shot_normals_body(gp, cell_start, sorted, radius_sq, normals, normals_sorted, blockIdx.x, gridDim.x);
}

// ---- cyclic Jacobi, symmetric 3x3, double (stands in for Eigen::SelfAdjointEigenSolver<Matrix3d>) ------
__device__ void jacobi_eigen3(const double Ain[6] /* xx xy xz yy yz zz */, double w[3], double V[9]) {
    double A[9] = {Ain[0], Ain[1], Ain[2], Ain[1], Ain[3], Ain[4], Ain[2], Ain[4], Ain[5]};
#pragma unroll
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
        const double diag = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
        if (off <= 1e-34 * diag || off == 0.0) break;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = p + 1; q < 3; ++q) {
                const double apq = A[3 * p + q];
                if (apq == 0.0) continue;
                const double theta = (A[3 * q + q] - A[3 * p + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[3 * k + p], akq = A[3 * k + q];
                    A[3 * k + p] = c * akp - s * akq;
                    A[3 * k + q] = s * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[3 * p + k], aqk = A[3 * q + k];
                    A[3 * p + k] = c * apk - s * aqk;
                    A[3 * q + k] = s * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[3 * k + p], vkq = V[3 * k + q];
                    V[3 * k + p] = c * vkp - s * vkq;
                    V[3 * k + q] = s * vkp + c * vkq;
                }
            }
    }
    // ascending eigenvalues (3-element sorting network on columns)
    double d[3] = {A[0], A[4], A[8]};
    int o0 = 0, o1 = 1, o2 = 2, t;
    if (d[o0] > d[o1]) { t = o0; o0 = o1; o1 = t; }
    if (d[o1] > d[o2]) { t = o1; o1 = o2; o2 = t; }
    if (d[o0] > d[o1]) { t = o0; o0 = o1; o1 = t; }
    double Vs[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        Vs[3 * k + 0] = V[3 * k + o0];
        Vs[3 * k + 1] = V[3 * k + o1];
        Vs[3 * k + 2] = V[3 * k + o2];
    }
    w[0] = d[o0];
    w[1] = d[o1];
    w[2] = d[o2];
#pragma unroll
    for (int i = 0; i < 9; ++i) V[i] = Vs[i];
}

// ---- LRF (shot_lrf.hpp::getLocalRF) in three stages -------------------------------------------------------------
// The eigen-solve is scalar work: inside the one-warp-per-key-point kernel all 32 lanes repeated it (~3 000 of the ~8 100
// warp instructions a key-point cost, ncu + SASS count of the sweep body).  It now runs one THREAD per key-point between
// two warp-per-key-point kernels; the scatter matrix and the axes cross in `lrf`, 8 doubles per sorted position:
//   after shot_lrf_cov_point:   [0..5] weighted scatter matrix (xx xy xz yy yz zz, not yet divided), [6] weight sum,
//                               [7] bits: valid | total << 32 (neighbours that are not the key-point itself | all in radius)
//   after shot_lrf_solve_one:   [0..2] eigenvector of the largest eigenvalue (x axis before sign disambiguation),
//                               [3..5] of the smallest (z axis), [6] 1.0 if the frame is usable else 0.0, [7] unchanged
// Same operations in the same order as the fused form, so descriptors are unchanged bit for bit.
constexpr int kLrfStride = 8;

// The in-radius neighbours of p in sweep order (positions in the sorted array) -> list[0 .. min(count, kShotListCap)); returns
// the count (warp-uniform).  Whole warp converged; same visiting order as for_each_candidate.
__device__ __forceinline__ int shot_neighbour_list(const ShotGrid &g, const int *__restrict__ cell_start,
                                                   const float4 *__restrict__ sorted, const float p[3], float radius_sq, int *list,
                                                   int lane) {
    int n_list = 0;
    const int cx = shot_coord(p[0], g.lo[0], g.inv, g.dim[0]);
    const int cy = shot_coord(p[1], g.lo[1], g.inv, g.dim[1]);
    const int cz = shot_coord(p[2], g.lo[2], g.inv, g.dim[2]);
    const int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.dim[2] - 1);
    for (int x = max(cx - 1, 0); x <= min(cx + 1, g.dim[0] - 1); ++x)
        for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y) {
            const int crow = (x * g.dim[1] + y) * g.dim[2];
            const int b = cell_start[crow + z0], e = cell_start[crow + z1 + 1];  // z-neighbours are contiguous
            for (int j0 = b; j0 < e; j0 += 64) {         // two candidates per lane and pass: both loads in flight together
                const int ja = j0 + lane, jb = ja + 32;
                float4 qa = make_float4(0.f, 0.f, 0.f, 0.f), qb = qa;
                if (ja < e) qa = sorted[ja];
                if (jb < e) qb = sorted[jb];
                const bool hit_a = ja < e && flann_dist2(p, qa) < radius_sq;
                const bool hit_b = jb < e && flann_dist2(p, qb) < radius_sq;
                const unsigned ma = __ballot_sync(0xffffffffu, hit_a), mb = __ballot_sync(0xffffffffu, hit_b);
                const unsigned below = (1u << lane) - 1u;
                if (ma | mb) {                           // appended in sweep order: the first 32 candidates, then the next 32
                    const int pos_a = n_list + __popc(ma & below);
                    if (hit_a && pos_a < kShotListCap) list[pos_a] = ja;
                    n_list += __popc(ma);
                    const int pos_b = n_list + __popc(mb & below);
                    if (hit_b && pos_b < kShotListCap) list[pos_b] = jb;
                    n_list += __popc(mb);
                }
            }
        }
    __syncwarp();
    return n_list;
}

// stage 1 (one warp): weighted scatter matrix in double.  The ~270 neighbours are first compacted out of the ~900 candidates
// (list: this warp's shared-memory scratch), so the double-precision part runs on full warps (9 passes instead of 28 sparse ones).
__device__ __forceinline__ void shot_lrf_cov_point(const ShotGrid &g, const int *__restrict__ cell_start,
                                                   const float4 *__restrict__ sorted, double radius, double *__restrict__ lrf,
                                                   int *list, int s, int lane) {
    const float radius_sq = static_cast<float>(radius * radius);
    const float4 pq = sorted[s];
    const float p[3] = {pq.x, pq.y, pq.z};
    double cov[6] = {0, 0, 0, 0, 0, 0}, wsum = 0.0;
    int valid = 0, total = 0;
    auto accumulate = [&](int, const float4 &q) {
        const float d2 = flann_dist2(p, q);
        if (d2 < radius_sq) {
            ++total;
            if (!(q.x == p[0] && q.y == p[1] && q.z == p[2])) {
                const double vx = static_cast<double>(__fsub_rn(q.x, p[0])), vy = static_cast<double>(__fsub_rn(q.y, p[1])),
                             vz = static_cast<double>(__fsub_rn(q.z, p[2]));
                const double w = radius - sqrt(static_cast<double>(d2));
                cov[0] += w * (vx * vx);
                cov[1] += w * (vx * vy);
                cov[2] += w * (vx * vz);
                cov[3] += w * (vy * vy);
                cov[4] += w * (vy * vz);
                cov[5] += w * (vz * vz);
                wsum += w;
                ++valid;
            }
        }
    };
    const int n_list = shot_neighbour_list(g, cell_start, sorted, p, radius_sq, list, lane);
    if (n_list <= kShotListCap) {
        for (int k = lane; k < n_list; k += 32) {
            const int j = list[k];
            accumulate(j, sorted[j]);
        }
    } else {
        for_each_candidate(g, cell_start, sorted, p, lane, accumulate);
    }
    __syncwarp();                                    // the list is rewritten for the warp's next key-point
#pragma unroll
    for (int i = 0; i < 6; ++i) cov[i] = warp_sum(cov[i]);
    wsum = warp_sum(wsum);
    valid = warp_sum(valid);
    total = warp_sum(total);
    // lanes 0..7 write one double each (one 64-byte row)
    const long long counts = static_cast<long long>(static_cast<unsigned int>(valid)) | (static_cast<long long>(total) << 32);
    double v = __longlong_as_double(counts);
    v = lane == 6 ? wsum : v;
#pragma unroll
    for (int i = 0; i < 6; ++i) v = lane == i ? cov[i] : v;
    if (lane < kLrfStride) lrf[static_cast<size_t>(s) * kLrfStride + lane] = v;
}

// stage 2 (one thread): normalise, eigen-solve, keep the two axes
__device__ __forceinline__ void shot_lrf_solve_one(double *__restrict__ row) {
    const long long counts = __double_as_longlong(row[7]);
    const int valid = static_cast<int>(counts & 0xffffffffll), total = static_cast<int>(counts >> 32);
    bool ok = valid >= 5 && total >= 5;
    if (ok) {
        const double wsum = row[6];
        double cov[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) cov[i] = row[i] / wsum;
        double w[3], V[9];
        jacobi_eigen3(cov, w, V);
        ok = isfinite(w[0]) && isfinite(w[1]) && isfinite(w[2]);
        row[0] = V[2];  // largest eigenvalue  -> x
        row[1] = V[5];
        row[2] = V[8];
        row[3] = V[0];  // smallest eigenvalue -> z
        row[4] = V[3];
        row[5] = V[6];
    }
    row[6] = ok ? 1.0 : 0.0;
}

// ---- stage 3: sign disambiguation + SHOT352 histogram --------------------------------------------------------------
// REAL = double reproduces PCL's double interpolation weights; REAL = float evaluates acos/atan2 and the
// weights in float (SHOT's quadrilinear interpolation is continuous across every bin boundary, so the
// descriptor moves by ~1e-7).
// SHOT-352 row of the point at sorted position s (one warp): LRF signs (shot_lrf.hpp::getLocalRF) + histogram (shot.hpp).
// hist [352] and list [kShotListCap] are this warp's shared-memory scratch.
template <typename REAL>
__device__ __forceinline__ void shot_descriptor_point(const ShotGrid &g, const int *__restrict__ cell_start,
                                                      const float4 *__restrict__ sorted, const float4 *__restrict__ normals_sorted,
                                                      const double *__restrict__ lrf, double radius, float *__restrict__ desc,
                                                      float *__restrict__ rf_out, unsigned int *hist, int *list, int s, int lane) {
    const float radius_sq = static_cast<float>(radius * radius);
    float hist_scale = 1.0f;
    auto hist_add = [&](int bin_index, float v) { atomicAdd(&hist[bin_index], __float2uint_rn(v * hist_scale)); };
    // REAL = float: quotients by the constants radius/2, 90 and 45 degrees become products with their reciprocals (<= 1 ulp of
    // the weight, the documented ~1e-7 of this mode), and the double square root of a float is the float square root bit for bit
    constexpr bool kFast = sizeof(REAL) == 4;
    const REAL r12 = static_cast<REAL>(radius / 2), r14 = static_cast<REAL>(radius / 4), r34 = static_cast<REAL>((radius * 3) / 4);
    const REAL inv_r12 = REAL(1) / r12;
    auto over = [&](REAL x, REAL c, REAL inv_c) { return kFast ? x * inv_c : x / c; };
    const REAL RAD_45 = static_cast<REAL>(0.78539816339744830961566084581988);
    const REAL RAD_90 = static_cast<REAL>(1.5707963267948966192313216916398);
    const REAL RAD_135 = static_cast<REAL>(2.3561944901923449288469825374596);
    const REAL RAD_7_8 = static_cast<REAL>(2.7488935718910690836548129603691);
    const REAL INV_RAD_90 = static_cast<REAL>(1.0 / 1.5707963267948966192313216916398);
    const REAL INV_RAD_45 = static_cast<REAL>(1.0 / 0.78539816339744830961566084581988);

    const float4 pq = sorted[s];
    const float p[3] = {pq.x, pq.y, pq.z};
    const int orig = __float_as_int(pq.w);
    float *out = desc + static_cast<size_t>(orig) * CPPF_SHOT_DIM;

    const double *row = lrf + static_cast<size_t>(s) * kLrfStride;   // warp-uniform loads
    const bool ok = row[6] != 0.0;
    const long long counts = __double_as_longlong(row[7]);
    const int valid = static_cast<int>(counts & 0xffffffffll), total = static_cast<int>(counts >> 32);
    if (!ok) {  // invalid LRF or fewer than 5 neighbours: NaN row (shot.hpp::computeFeature / computePointSHOT)
        if (rf_out && lane < 9) rf_out[static_cast<size_t>(orig) * 9 + lane] = CUDART_NAN_F;
        for (int j = lane; j < CPPF_SHOT_DIM; j += 32) out[j] = CUDART_NAN_F;
        return;
    }

    // the in-radius neighbours, in sweep order (positions in the sorted array)
    const int n_list = shot_neighbour_list(g, cell_start, sorted, p, radius_sq, list, lane);     // warp-uniform
    const bool listed = n_list <= kShotListCap;
    // passes B and C: the cached neighbours when they fit, else the cells again
    auto for_each_neighbour = [&](auto &&f) {
        if (listed) {
            for (int k = lane; k < n_list; k += 32) {
                const int j = list[k];
                f(j, sorted[j]);
            }
        } else {
            for_each_candidate(g, cell_start, sorted, p, lane, f);
        }
    };

    float fx[3], fy[3], fz[3];
    {
        double v1[3] = {row[0], row[1], row[2]};
        double v3[3] = {row[3], row[4], row[5]};
        // pass B: sign disambiguation votes
        int plus_t = 0, plus_n = 0;
        for_each_neighbour([&](int, const float4 &q) {
            if (flann_dist2(p, q) < radius_sq && !(q.x == p[0] && q.y == p[1] && q.z == p[2])) {
                const double vx = static_cast<double>(__fsub_rn(q.x, p[0])), vy = static_cast<double>(__fsub_rn(q.y, p[1])),
                             vz = static_cast<double>(__fsub_rn(q.z, p[2]));
                if (vx * v1[0] + vy * v1[1] + vz * v1[2] >= 0.0) ++plus_t;
                if (vx * v3[0] + vy * v3[1] + vz * v3[2] >= 0.0) ++plus_n;
            }
        });
        plus_t = 2 * warp_sum(plus_t) - valid;
        plus_n = 2 * warp_sum(plus_n) - valid;
        // exact ties fall back to PCL's search-order dependent rule (5 neighbours around the median
        // of the kd-tree order); without that order the axis is left as the solver produced it.
        if (plus_t < 0) { v1[0] = -v1[0]; v1[1] = -v1[1]; v1[2] = -v1[2]; }
        if (plus_n < 0) { v3[0] = -v3[0]; v3[1] = -v3[1]; v3[2] = -v3[2]; }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            fx[k] = static_cast<float>(v1[k]);
            fz[k] = static_cast<float>(v3[k]);
        }
        cross3f(fz, fx, fy);
    }
    if (rf_out && lane < 9) {
        const float v = lane < 3 ? fx[lane] : (lane < 6 ? fy[lane - 3] : fz[lane - 6]);
        rf_out[static_cast<size_t>(orig) * 9 + lane] = v;
    }

    // pass C: histogram (shot.hpp::createBinDistanceShape + interpolateSingleChannel)
    for (int j = lane; j < CPPF_SHOT_DIM; j += 32) hist[j] = 0u;
    {
        const unsigned m = 5u * static_cast<unsigned>(total) + 1u;          // total < 2^29 points
        const int k = min(24, __clz(m));                                    // 32 - bits(m)
        hist_scale = static_cast<float>(1u << k);
    }
    __syncwarp();
    for_each_neighbour([&](int j, const float4 &q) {
        const float d2 = flann_dist2(p, q);
        if (!(d2 < radius_sq)) return;
        const float4 nq = normals_sorted[j];
        if (!(isfinite(nq.x) && isfinite(nq.y) && isfinite(nq.z))) return;
        REAL cosine = static_cast<REAL>(nq.x * fz[0] + nq.y * fz[1] + nq.z * fz[2]);
        cosine = cosine > REAL(1) ? REAL(1) : (cosine < REAL(-1) ? REAL(-1) : cosine);
        REAL bin = ((REAL(1) + cosine) * REAL(10)) / REAL(2);
        const float dl[3] = {__fsub_rn(q.x, p[0]), __fsub_rn(q.y, p[1]), __fsub_rn(q.z, p[2])};
        const REAL distance = kFast ? static_cast<REAL>(__fsqrt_rn(d2)) : static_cast<REAL>(sqrt(static_cast<double>(d2)));
        if (fabs(static_cast<double>(distance)) < 1e-15) return;
        REAL xr = static_cast<REAL>(dl[0] * fx[0] + dl[1] * fx[1] + dl[2] * fx[2]);
        REAL yr = static_cast<REAL>(dl[0] * fy[0] + dl[1] * fy[1] + dl[2] * fy[2]);
        REAL zr = static_cast<REAL>(dl[0] * fz[0] + dl[1] * fz[1] + dl[2] * fz[2]);
        if (fabs(static_cast<double>(yr)) < 1e-30) yr = 0;
        if (fabs(static_cast<double>(xr)) < 1e-30) xr = 0;
        if (fabs(static_cast<double>(zr)) < 1e-30) zr = 0;
        const int bit4 = ((yr > 0) || ((yr == 0) && (xr < 0))) ? 1 : 0;
        const int bit3 = ((xr > 0) || ((xr == 0) && (yr > 0))) ? (1 - bit4) : bit4;
        int di = ((bit4 << 3) + (bit3 << 2)) << 1;
        const REAL ax = xr < 0 ? -xr : xr, ay = yr < 0 ? -yr : yr;
        if ((xr * yr > 0) || (xr == 0))
            di += (ax >= ay) ? 0 : 4;
        else
            di += (ax > ay) ? 4 : 0;
        di += zr > 0 ? 1 : 0;
        di += (distance > r12) ? 2 : 0;
        const int step = static_cast<int>(floor(static_cast<double>(bin) + 0.5));
        const int vol = di * 11;
        bin -= static_cast<REAL>(step);
        REAL wgt = REAL(1) - (bin < 0 ? -bin : bin);
        if (bin > 0)
            hist_add(vol + ((step + 1) % 10), static_cast<float>(bin));
        else
            hist_add(vol + ((step - 1 + 10) % 10), -static_cast<float>(bin));
        if (distance > r12) {
            const REAL rd = over(distance - r34, r12, inv_r12);
            if (distance > r34)
                wgt += REAL(1) - rd;
            else {
                wgt += REAL(1) + rd;
                hist_add((di - 2) * 11 + step, -static_cast<float>(rd));
            }
        } else {
            const REAL rd = over(distance - r14, r12, inv_r12);
            if (distance < r14)
                wgt += REAL(1) + rd;
            else {
                wgt += REAL(1) - rd;
                hist_add((di + 2) * 11 + step, static_cast<float>(rd));
            }
        }
        REAL ic = zr / distance;
        ic = ic < REAL(-1) ? REAL(-1) : (ic > REAL(1) ? REAL(1) : ic);
        const REAL incl = acos(ic);
        if (incl > RAD_90 || (fabs(static_cast<double>(incl - RAD_90)) < 1e-30 && zr <= 0)) {
            const REAL id = over(incl - RAD_135, RAD_90, INV_RAD_90);
            if (incl > RAD_135)
                wgt += REAL(1) - id;
            else {
                wgt += REAL(1) + id;
                hist_add((di + 1) * 11 + step, -static_cast<float>(id));
            }
        } else {
            const REAL id = over(incl - RAD_45, RAD_90, INV_RAD_90);
            if (incl < RAD_45)
                wgt += REAL(1) + id;
            else {
                wgt += REAL(1) - id;
                hist_add((di - 1) * 11 + step, static_cast<float>(id));
            }
        }
        if (yr != 0 || xr != 0) {
            const REAL az = atan2(yr, xr);
            const int sel = di >> 2;
            REAL ad = over(az - (-RAD_7_8 + RAD_45 * static_cast<REAL>(sel)), RAD_45, INV_RAD_45);
            ad = ad < REAL(-0.5) ? REAL(-0.5) : (ad > REAL(0.5) ? REAL(0.5) : ad);
            if (ad > 0) {
                wgt += REAL(1) - ad;
                hist_add(((di + 4) % 32) * 11 + step, static_cast<float>(ad));
            } else {
                wgt += REAL(1) + ad;
                hist_add(((di - 4 + 32) % 32) * 11 + step, -static_cast<float>(ad));
            }
        }
        hist_add(vol + step, static_cast<float>(wgt));
    });
    __syncwarp();
    // normalizeHistogram: float squares accumulated in double, divide by float(norm)
    double acc = 0.0;
    float hv[CPPF_SHOT_DIM / 32];
    const double inv_scale = 1.0 / static_cast<double>(hist_scale);      // a power of two: the product below is the exact quotient
#pragma unroll
    for (int u = 0; u < CPPF_SHOT_DIM / 32; ++u) {
        hv[u] = static_cast<float>(static_cast<double>(hist[lane + 32 * u]) * inv_scale);
        acc += static_cast<double>(hv[u] * hv[u]);
    }
    acc = warp_sum(acc);
    const float nrm = static_cast<float>(sqrt(acc));
#pragma unroll
    for (int u = 0; u < CPPF_SHOT_DIM / 32; ++u) out[lane + 32 * u] = hv[u] / nrm;
    __syncwarp();
}

__global__ void __launch_bounds__(kShotWarps * 32, 4) shot_lrf_cov_kernel(const ShotGrid *__restrict__ gp, const int *__restrict__ cell_start,
                                                                      const float4 *__restrict__ sorted, double radius,
                                                                      double *__restrict__ lrf) {
    __shared__ int s_list[kShotWarps][kShotListCap];
    const ShotGrid g = *gp;
    const int n_sorted = cell_start[g.cells];
    const int lane = lane_id();
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_sorted; s += n_warps)
        shot_lrf_cov_point(g, cell_start, sorted, radius, lrf, s_list[threadIdx.x >> 5], s, lane);
}

__global__ void __launch_bounds__(64) shot_lrf_solve_kernel(const ShotGrid *__restrict__ gp, const int *__restrict__ cell_start,
                                                            double *__restrict__ lrf) {
    const int n_sorted = cell_start[gp->cells];
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_sorted) shot_lrf_solve_one(lrf + static_cast<size_t>(s) * kLrfStride);
}

template <typename REAL>
__global__ void __launch_bounds__(kShotWarps * 32) shot_descriptor_kernel(
    const ShotGrid *__restrict__ gp, const int *__restrict__ cell_start, const float4 *__restrict__ sorted,
    const float4 *__restrict__ normals_sorted, const double *__restrict__ lrf, double radius, float *__restrict__ desc,
    float *__restrict__ rf_out) {
    // The histogram accumulates in 32-bit fixed point: shared-memory float atomicAdd is a compare-and-swap loop (60 % of this
    // kernel's stall samples, ncu; so is the 64-bit integer add), the 32-bit integer add is one native ATOMS.ADD.  A neighbour
    // adds at most 5 over all bins, so with `total` neighbours a scale of 2^k, k = 32 - bits(5 total + 1) (<= 24), cannot
    // overflow: 2^21 for the ~270 neighbours of a cloud sampled at radius/10.  Each contribution is rounded to 2^-k once;
    // the sum itself is exact, hence independent of the order the lanes arrive in and identical from run to run (PCL adds in
    // float in kd-tree order, which nothing pins either).
    __shared__ unsigned int s_hist[kShotWarps][CPPF_SHOT_DIM];
    // in-radius neighbours (positions in the sorted array, sweep order): the sign votes and the histogram walk this list
    // instead of sweeping the 27 cells again -- a surface sampled at radius/10 has ~270 neighbours among ~900 candidates
    __shared__ int s_list[kShotWarps][kShotListCap];
    const ShotGrid g = *gp;
    const int n_sorted = cell_start[g.cells];
    const int lane = lane_id();
    const int wib = threadIdx.x >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_sorted; s += n_warps)
        shot_descriptor_point<REAL>(g, cell_start, sorted, normals_sorted, lrf, radius, desc, rf_out, s_hist[wib], s_list[wib], s, lane);
}

// normals_sorted[s] = normals_in[orig(s)]: used when the caller supplies the normals
__global__ void __launch_bounds__(256) shot_gather_normals_kernel(const ShotGrid *__restrict__ gp, const int *__restrict__ cell_start,
                                                                  const float4 *__restrict__ sorted,
                                                                  const float *__restrict__ normals_in,
                                                                  float4 *__restrict__ normals_sorted) {
    const int n_sorted = cell_start[gp->cells];
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sorted) return;
    const int orig = __float_as_int(sorted[s].w);
    normals_sorted[s] = make_float4(normals_in[3 * orig], normals_in[3 * orig + 1], normals_in[3 * orig + 2], 0.0f);
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

size_t shot_carve(void *ws, int64_t n, ShotWorkspace *out) {
    unsigned char *base = static_cast<unsigned char *>(ws);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *p = base ? base + off : nullptr;
        off += align256(bytes);
        return p;
    };
    ShotWorkspace w;
    w.bounds = static_cast<cppf_grid_geom *>(take(sizeof(cppf_grid_geom)));
    w.grid = static_cast<ShotGrid *>(take(sizeof(ShotGrid)));
    w.cell_count = static_cast<int *>(take(sizeof(int) * (kShotMaxCells + 1)));
    w.cell_start = static_cast<int *>(take(sizeof(int) * (kShotMaxCells + 1)));
    w.cell_fill = static_cast<int *>(take(sizeof(int) * (kShotMaxCells + 1)));
    w.cell_of = static_cast<int *>(take(sizeof(int) * static_cast<size_t>(n)));
    w.sorted = static_cast<float4 *>(take(sizeof(float4) * static_cast<size_t>(n)));
    w.normals_sorted = static_cast<float4 *>(take(sizeof(float4) * static_cast<size_t>(n)));
    w.lrf = static_cast<double *>(take(sizeof(double) * kLrfStride * static_cast<size_t>(n)));
    if (out) *out = w;
    return off;
}

static int shot_build_grid(const float *pc, int64_t n, float radius, const ShotWorkspace &w, cudaStream_t s) {
    int rc = cppf_cloud_bounds(pc, n, radius, w.bounds, s);
    if (rc) return rc;
    shot_grid_setup_kernel<<<1, 32, 0, s>>>(w.bounds, radius, w.grid);
    CPPF_LAUNCH_CHECK();
    CPPF_CUDA_TRY(cudaMemsetAsync(w.cell_count, 0, sizeof(int) * (kShotMaxCells + 1), s));
    const int nb = div_up(n, 256);
    shot_cell_count_kernel<<<nb, 256, 0, s>>>(pc, static_cast<int>(n), w.grid, w.cell_of, w.cell_count);
    CPPF_LAUNCH_CHECK();
    shot_cell_scan_kernel<<<1, 1024, 0, s>>>(w.grid, w.cell_count, w.cell_start, w.cell_fill);
    CPPF_LAUNCH_CHECK();
    // scatter into the (not yet used) normals buffer, then order every cell by original index
    shot_scatter_kernel<<<nb, 256, 0, s>>>(pc, static_cast<int>(n), w.cell_of, w.cell_fill, w.normals_sorted);
    CPPF_LAUNCH_CHECK();
    shot_cell_order_kernel<<<grid_for(n * 8, 256, 8), 256, 0, s>>>(w.grid, w.cell_start, w.normals_sorted, w.sorted);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

// =====================================================================================================================
// Batched frame path (frame.cuh): SHOT-352 + normals of every instance of a frame in nine launches; blockIdx.y = instance.
// Same bodies as the single-cloud kernels, so descriptors are identical to cppf_shot_compute_ex(fast_math = 1).
// =====================================================================================================================
__device__ __forceinline__ bool frame_shot_instance(const FrameTable *t, const FrameInst *&in) {
    if (static_cast<int>(blockIdx.y) >= t->n_inst) return false;
    in = &t->inst[blockIdx.y];
    return in->shot_desc != nullptr && in->n > 0;
}

// bounds -> search grid header -> its cell counters zeroed (one CTA per instance)
__global__ void __launch_bounds__(1024) frame_shot_setup_kernel(const FrameTable *__restrict__ t) {
    const FrameInst *in;
    if (!frame_shot_instance(t, in)) return;
    const float cell_r = fmaxf(in->normal_r, in->shot_r);      // one grid serves both searches
    cloud_bounds_shared(in->pc, in->n, cell_r, in->sw.bounds);
    __syncthreads();
    if (threadIdx.x == 0) shot_grid_setup(in->sw.bounds, cell_r, in->sw.grid);
    __syncthreads();
    const int cells = in->sw.grid->cells;
    for (int i = threadIdx.x; i <= cells; i += blockDim.x) in->sw.cell_count[i] = 0;
}

__global__ void __launch_bounds__(256) frame_shot_count_kernel(const FrameTable *__restrict__ t) {
    const FrameInst *in;
    if (!frame_shot_instance(t, in)) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < in->n; i += gridDim.x * blockDim.x)
        shot_cell_count_one(in->pc, i, in->sw.grid, in->sw.cell_of, in->sw.cell_count);
}

__global__ void __launch_bounds__(1024) frame_shot_scan_kernel(const FrameTable *__restrict__ t) {
    const FrameInst *in;
    if (!frame_shot_instance(t, in)) return;
    shot_cell_scan_body(in->sw.grid, in->sw.cell_count, in->sw.cell_start, in->sw.cell_fill);
}

__global__ void __launch_bounds__(256) frame_shot_scatter_kernel(const FrameTable *__restrict__ t) {
    const FrameInst *in;
    if (!frame_shot_instance(t, in)) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < in->n; i += gridDim.x * blockDim.x)
        shot_scatter_one(in->pc, i, in->sw.cell_of, in->sw.cell_fill, in->sw.normals_sorted);   // staging: ordered next
}

__global__ void __launch_bounds__(256) frame_shot_order_kernel(const FrameTable *__restrict__ t) {
    const FrameInst *in;
    if (!frame_shot_instance(t, in)) return;
    shot_cell_order_body(in->sw.grid, in->sw.cell_start, in->sw.normals_sorted, in->sw.sorted, blockIdx.x, gridDim.x);
}

// The two per-point kernels take ONE flat list of all instances' points (shot_base = prefix sums of the clouds' sizes): warp w
// of the grid owns the flat positions w, w + n_warps, ...  A frame's clouds differ in size by 2-3x; with one grid row per
// instance the largest cloud set the time of the launch.
__device__ __forceinline__ int frame_shot_locate(const FrameTable *__restrict__ t, int g, int &s) {
    int i = 0;
    while (i + 1 < t->n_inst && g >= t->shot_base[i + 1]) ++i;
    s = g - t->shot_base[i];
    return i;
}

// normals; points with non-finite coordinates (never in the grid) get their NaN rows here
__global__ void __launch_bounds__(kShotWarps * 32) frame_shot_normals_kernel(const FrameTable *__restrict__ t) {
    const int total = t->shot_base[t->n_inst];
    const int lane = lane_id();
    const float nanv = CUDART_NAN_F;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < total; g += (gridDim.x * blockDim.x) >> 5) {
        int s;
        const FrameInst &in = t->inst[frame_shot_locate(t, g, s)];
        // position s serves twice: as an ORIGINAL index for the NaN rows of points outside the grid ...
        if (in.sw.cell_of[s] < 0) {
            if (lane < 3) in.normals[3 * s + lane] = nanv;
            for (int j = lane; j < CPPF_SHOT_DIM; j += 32) in.shot_desc[static_cast<size_t>(s) * CPPF_SHOT_DIM + j] = nanv;
        }
        // ... and as a SORTED position for the normal of the point stored there
        const ShotGrid gr = *in.sw.grid;
        if (s >= in.sw.cell_start[gr.cells]) continue;
        const float nr2 = static_cast<float>(static_cast<double>(in.normal_r) * static_cast<double>(in.normal_r));
        shot_normal_point(gr, in.sw.cell_start, in.sw.sorted, nr2, in.normals, in.sw.normals_sorted, s, lane);
    }
}

// LRF stage 1 (warp per flat position) and stage 2 (thread per flat position)
__global__ void __launch_bounds__(kShotWarps * 32, 4) frame_shot_lrf_cov_kernel(const FrameTable *__restrict__ t) {
    __shared__ int s_list[kShotWarps][kShotListCap];
    const int total = t->shot_base[t->n_inst];
    const int lane = lane_id();
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < total; g += (gridDim.x * blockDim.x) >> 5) {
        int s;
        const FrameInst &in = t->inst[frame_shot_locate(t, g, s)];
        const ShotGrid gr = *in.sw.grid;
        if (s >= in.sw.cell_start[gr.cells]) continue;
        shot_lrf_cov_point(gr, in.sw.cell_start, in.sw.sorted, static_cast<double>(in.shot_r), in.sw.lrf, s_list[threadIdx.x >> 5], s, lane);
    }
}

__global__ void __launch_bounds__(64) frame_shot_lrf_solve_kernel(const FrameTable *__restrict__ t) {
    const int total = t->shot_base[t->n_inst];
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
        int s;
        const FrameInst &in = t->inst[frame_shot_locate(t, g, s)];
        if (s >= in.sw.cell_start[in.sw.grid->cells]) continue;
        shot_lrf_solve_one(in.sw.lrf + static_cast<size_t>(s) * kLrfStride);
    }
}

__global__ void __launch_bounds__(kShotWarps * 32, 4) frame_shot_descriptor_kernel(const FrameTable *__restrict__ t) {
    __shared__ unsigned int s_hist[kShotWarps][CPPF_SHOT_DIM];
    __shared__ int s_list[kShotWarps][kShotListCap];
    const int total = t->shot_base[t->n_inst];
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < total; g += (gridDim.x * blockDim.x) >> 5) {
        int s;
        const FrameInst &in = t->inst[frame_shot_locate(t, g, s)];
        const ShotGrid gr = *in.sw.grid;
        if (s >= in.sw.cell_start[gr.cells]) continue;
        shot_descriptor_point<float>(gr, in.sw.cell_start, in.sw.sorted, in.sw.normals_sorted, in.sw.lrf, static_cast<double>(in.shot_r),
                                     in.shot_desc, nullptr, s_hist[wib], s_list[wib], s, lane);
    }
}

// Plain stream-ordered launches: with programmatic dependent launch (common.cuh) along these kernels the stage was
// slower on B200 (0.294 against 0.277 ms per frame), unlike the vote chain's.
int frame_launch_shot(const FrameTable *t, int ni, int64_t n_cap, cudaStream_t s) {
    if (ni <= 0 || n_cap <= 0) return CPPF_OK;
    const int sms = device_info().sm_count;
    // per instance: enough CTAs for the capacity, the whole launch a few waves of the GPU
    auto per_inst = [&](int64_t items, int threads, int per_sm) {
        const int64_t need = (items + threads - 1) / threads;
        const int64_t cap = std::max<int64_t>(1, (static_cast<int64_t>(sms) * per_sm + ni - 1) / ni);
        return static_cast<int>(std::max<int64_t>(1, std::min(need, cap)));
    };
    frame_shot_setup_kernel<<<dim3(1, ni), 1024, 0, s>>>(t);
    CPPF_LAUNCH_CHECK();
    frame_shot_count_kernel<<<dim3(per_inst(n_cap, 256, 8), ni), 256, 0, s>>>(t);
    CPPF_LAUNCH_CHECK();
    frame_shot_scan_kernel<<<dim3(1, ni), 1024, 0, s>>>(t);
    CPPF_LAUNCH_CHECK();
    frame_shot_scatter_kernel<<<dim3(per_inst(n_cap, 256, 8), ni), 256, 0, s>>>(t);
    CPPF_LAUNCH_CHECK();
    frame_shot_order_kernel<<<dim3(per_inst(n_cap * 8, 256, 8), ni), 256, 0, s>>>(t);
    CPPF_LAUNCH_CHECK();
    // flat over all instances' points: a whole number of waves, capped by the work the capacity allows
    const int64_t warps_cap = n_cap * ni;
    frame_shot_normals_kernel<<<dim3(grid_for(warps_cap * 32, kShotWarps * 32, 8)), kShotWarps * 32, 0, s>>>(t);
    CPPF_LAUNCH_CHECK();
    frame_shot_lrf_cov_kernel<<<dim3(grid_for(warps_cap * 32, kShotWarps * 32, 4)), kShotWarps * 32, 0, s>>>(t);
    CPPF_LAUNCH_CHECK();
    frame_shot_lrf_solve_kernel<<<dim3(static_cast<int>(std::min<int64_t>(div_up(warps_cap, 64), sms * 32))), 64, 0, s>>>(t);
    CPPF_LAUNCH_CHECK();
    frame_shot_descriptor_kernel<<<dim3(grid_for(warps_cap * 32, kShotWarps * 32, 4)), kShotWarps * 32, 0, s>>>(t);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

}  // namespace cppf

using namespace cppf;

CPPF_API int64_t cppf_shot_workspace_bytes(int64_t n) { return static_cast<int64_t>(shot_carve(nullptr, n < 1 ? 1 : n, nullptr)); }

static int shot_run(const float *pc, int64_t n, float normal_r, float shot_r, float *desc, float *normals, float *rf_out,
                    int fast_math, const float *normals_in, void *ws, int64_t ws_bytes, void *stream) {
    if (n == 0) return CPPF_OK;
    if (!pc || !ws || n < 0 || (desc && !(shot_r > 0.0f))) return CPPF_ERR_INVALID_ARGUMENT;
    if (!normals_in && (!normals || !(normal_r > 0.0f))) return CPPF_ERR_INVALID_ARGUMENT;
    if (n > (1ll << 30)) return CPPF_ERR_UNSUPPORTED;
    if (n == 0) return CPPF_OK;
    if (ws_bytes < cppf_shot_workspace_bytes(n)) return CPPF_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ShotWorkspace w;
    shot_carve(ws, n, &w);
    // one grid serves both searches: its cell edge is the larger of the two radii
    const float cell_r = normals_in ? shot_r : (desc ? fmaxf(normal_r, shot_r) : normal_r);
    int rc = shot_build_grid(pc, n, cell_r, w, s);
    if (rc) return rc;
    if (normals_in) {
        shot_gather_normals_kernel<<<div_up(n, 256), 256, 0, s>>>(w.grid, w.cell_start, w.sorted, normals_in, w.normals_sorted);
        CPPF_LAUNCH_CHECK();
        if (normals && normals != normals_in)
            CPPF_CUDA_TRY(cudaMemcpyAsync(normals, normals_in, sizeof(float) * 3 * static_cast<size_t>(n), cudaMemcpyDeviceToDevice, s));
        shot_nan_fill_kernel<<<static_cast<int>(n), 64, 0, s>>>(w.cell_of, static_cast<int>(n), nullptr, desc);
        CPPF_LAUNCH_CHECK();
    } else {
        shot_nan_fill_kernel<<<static_cast<int>(n), 64, 0, s>>>(w.cell_of, static_cast<int>(n), normals, desc);
        CPPF_LAUNCH_CHECK();
        const int blocks = grid_for(n * 32, kShotWarps * 32, 8);
        const float nr2 = static_cast<float>(static_cast<double>(normal_r) * static_cast<double>(normal_r));
        shot_normals_kernel<<<blocks, kShotWarps * 32, 0, s>>>(w.grid, w.cell_start, w.sorted, static_cast<int>(n), nr2, normals,
                                                              w.normals_sorted);
        CPPF_LAUNCH_CHECK();
    }
    if (!desc) return CPPF_OK;
    shot_lrf_cov_kernel<<<grid_for(n * 32, kShotWarps * 32, 4), kShotWarps * 32, 0, s>>>(w.grid, w.cell_start, w.sorted,
                                                                                        static_cast<double>(shot_r), w.lrf);
    CPPF_LAUNCH_CHECK();
    shot_lrf_solve_kernel<<<div_up(n, 64), 64, 0, s>>>(w.grid, w.cell_start, w.lrf);
    CPPF_LAUNCH_CHECK();
    const int dblocks = grid_for(n * 32, kShotWarps * 32, 4);
    if (fast_math)
        shot_descriptor_kernel<float><<<dblocks, kShotWarps * 32, 0, s>>>(w.grid, w.cell_start, w.sorted, w.normals_sorted, w.lrf,
                                                                         static_cast<double>(shot_r), desc, rf_out);
    else
        shot_descriptor_kernel<double><<<dblocks, kShotWarps * 32, 0, s>>>(w.grid, w.cell_start, w.sorted, w.normals_sorted, w.lrf,
                                                                          static_cast<double>(shot_r), desc, rf_out);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_shot_compute(const float *pc, int64_t n, float normal_r, float shot_r, float *desc, float *normals,
                               void *ws, int64_t ws_bytes, void *stream) {
    if (n == 0) return CPPF_OK;
    if (!desc) return CPPF_ERR_INVALID_ARGUMENT;
    // float interpolation weights: indistinguishable from PCL's double ones at float32 output precision
    // (tests/test_gpu_shot.py compares both against the oracle), and faster
    return shot_run(pc, n, normal_r, shot_r, desc, normals, nullptr, 1, nullptr, ws, ws_bytes, stream);
}

// Same with the interpolation weights in float (fast_math != 0), the local reference frames exposed
// (rf_out [n,9], rows x,y,z; may be NULL) and, optionally, caller-supplied normals (normals_in [n,3]:
// the normal estimation is skipped, e.g. to reuse normals or to test the descriptor stage in isolation).
CPPF_API int cppf_shot_compute_ex(const float *pc, int64_t n, float normal_r, float shot_r, float *desc, float *normals,
                                  float *rf_out, int fast_math, const float *normals_in, void *ws, int64_t ws_bytes,
                                  void *stream) {
    if (n == 0) return CPPF_OK;
    if (!desc) return CPPF_ERR_INVALID_ARGUMENT;
    return shot_run(pc, n, normal_r, shot_r, desc, normals, rf_out, fast_math, normals_in, ws, ws_bytes, stream);
}

CPPF_API int cppf_estimate_normal(const float *pc, int64_t n, float normal_r, float *normals, void *ws, int64_t ws_bytes,
                                  void *stream) {
    return shot_run(pc, n, normal_r, 0.0f, nullptr, normals, nullptr, 0, nullptr, ws, ws_bytes, stream);
}

CPPF_API int cppf_shot_compute_color(const float *, const float *, int64_t, float, float, float *, void *) {
    return CPPF_ERR_UNSUPPORTED;  // SHOT1344 (shot.cpp:102-161) has no caller anywhere in the reference
}
