/*
 * cppf_oracle.c -- CPU restatement of the CPPF++ voting path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA kernels in cppf2_b200/csrc/ and the CPU baseline that
 * bench.py times next to them.  It is never linked into, imported by or called from the product
 * path (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs).
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit (integer grids) or within
 * the stated tolerance (float stages) against tests/golden/*.npz, which oracle/make_golden.py
 * minted by running the unmodified reference functions on torch-CPU in the build container.
 *
 * The reference computes this stage with eager torch-CPU / numpy ops, so every operation rounds
 * to its dtype; the few places where torch's CPU kernels fuse (FMA chain inside torch.norm, FMA in
 * torch.cross) are restated with explicit fmaf().  Build with -ffp-contract=off (see Makefile) so
 * that the compiler adds no contractions of its own.
 *
 * Reference files followed (paths relative to the reference root):
 *   train_dino.py:171-215  vote_center          -> oracle_grid_geometry, oracle_vote_center, oracle_grid_argmax
 *   train_dino.py:218-239  vote_rotation        -> oracle_vote_rotation
 *   dataset.py:118-135     generate_target_pairs-> oracle_generate_targets
 *   eval.py:37-51          get_topk_dir         -> oracle_sphere_hist
 *   eval.py:225-235        decode               -> oracle_decode_pairs
 *   utils/util.py:191-207  fibonacci_sphere     -> oracle_fibonacci_sphere
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

/* torch.norm(v, dim=-1) over 3 contiguous floats on the CPU path == sqrt(fma(z,z,fma(y,y,x*x))) */
static inline float norm3_torch(float x, float y, float z) { return sqrtf(fmaf(z, z, fmaf(y, y, x * x))); }

/* np.linalg.norm(v, axis=-1) on float32: sqrt((x*x + y*y) + z*z), every op rounded, no FMA */
static inline float norm3_numpy(float x, float y, float z) { return sqrtf((x * x + y * y) + z * z); }

EXPORT int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

EXPORT void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Grid geometry: corners = [pc.min(0), pc.max(0)]; grid_res = trunc((hi-lo)/res32) + 1
 * (train_dino.py:172-173).  `res` is a Python float in the reference; tensor/float on torch-CPU is a
 * true division by float32(res).
 * ---------------------------------------------------------------------------------------------- */
EXPORT int oracle_grid_geometry(const float *pc, int64_t n, float res, float lo[3], float hi[3], int64_t grid_res[3]) {
    if (n <= 0) return 1;
    for (int k = 0; k < 3; ++k) { lo[k] = pc[k]; hi[k] = pc[k]; }
    for (int64_t i = 1; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            float v = pc[3 * i + k];
            if (v < lo[k]) lo[k] = v;
            if (v > hi[k]) hi[k] = v;
        }
    for (int k = 0; k < 3; ++k) grid_res[k] = (int64_t)((hi[k] - lo[k]) / res) + 1;
    return 0;
}

/* Pair frame shared by vote_center (train_dino.py:178-192) and vote_rotation (:219-231).
 * Returns 0 when the pair is masked out by |ab| > 1e-7.  `clamp_co` selects vote_rotation's
 * clamp_min(norm(co),1e-7) (train_dino.py:229) over vote_center's bare division (:191). */
static inline int pair_frame(const float *a, const float *b, int clamp_co, float ab[3], float x[3]) {
    ab[0] = a[0] - b[0]; ab[1] = a[1] - b[1]; ab[2] = a[2] - b[2];
    float n = norm3_torch(ab[0], ab[1], ab[2]);
    if (!(n > 1e-7f)) return 0;
    float d = n < 1e-7f ? 1e-7f : n;                    /* clamp_min(norm, 1e-7) */
    ab[0] /= d; ab[1] /= d; ab[2] /= d;
    float co[3] = {0.0f, -ab[2], ab[1]};
    float cn = norm3_torch(co[0], co[1], co[2]);
    if (cn < 1e-7f) {                                   /* ab parallel to the x axis */
        co[0] = -ab[1]; co[1] = ab[0]; co[2] = 0.0f;
        cn = norm3_torch(co[0], co[1], co[2]);
    }
    if (clamp_co && cn < 1e-7f) cn = 1e-7f;
    x[0] = co[0] / cn; x[1] = co[1] / cn; x[2] = co[2] / cn;
    return 1;
}

/* torch.cross(x, ab) on CPU: fma(x1, ab2, -(x2*ab1)) and cyclic */
static inline void cross_torch(const float x[3], const float ab[3], float y[3]) {
    y[0] = fmaf(x[1], ab[2], -(x[2] * ab[1]));
    y[1] = fmaf(x[2], ab[0], -(x[0] * ab[2]));
    y[2] = fmaf(x[0], ab[1], -(x[1] * ab[0]));
}

/* ------------------------------------------------------------------------------------------------
 * vote_center (train_dino.py:171-215), integer grid.  idx is [T, idx_stride] int64 with the pair in
 * columns 0,1.  cos_tab/sin_tab hold torch.cos/sin of arange(R)/R*2*pi computed by the caller on
 * torch-CPU (they are inputs on purpose: Sleef differs from libm in a few ulp).
 * grid is int64 [gx*gy*gz], C order, accumulated into (caller zeroes it).  Returns the number of
 * votes that landed inside the grid.
 * ---------------------------------------------------------------------------------------------- */
EXPORT int64_t oracle_vote_center(const float *pc, int64_t n, const int64_t *idx, int64_t idx_stride, const float *tr,
                                  int64_t T, float res, const float *cos_tab, const float *sin_tab, int R,
                                  const float lo[3], const int64_t grid_res[3], int64_t *grid) {
    (void)n;
    int64_t landed = 0;
    const int64_t gy = grid_res[1], gz = grid_res[2];
#pragma omp parallel for schedule(static) reduction(+ : landed)
    for (int64_t t = 0; t < T; ++t) {
        const float *a = pc + 3 * idx[t * idx_stride + 0];
        const float *b = pc + 3 * idx[t * idx_stride + 1];
        const float proj_len = tr[2 * t + 0], odist = tr[2 * t + 1];
        float ab[3], x[3], y[3], c[3];
        if (!(odist > res)) continue;                                   /* :182 */
        if (!pair_frame(a, b, 0, ab, x)) continue;
        for (int k = 0; k < 3; ++k) c[k] = a[k] - ab[k] * proj_len;      /* :186 */
        for (int k = 0; k < 3; ++k) x[k] = x[k] * odist;                 /* :191 (divide, then multiply) */
        cross_torch(x, ab, y);                                          /* :192 */
        for (int r = 0; r < R; ++r) {
            int64_t cell[3];
            int ok = 1;
            for (int k = 0; k < 3; ++k) {
                float off = cos_tab[r] * x[k] + sin_tab[r] * y[k];      /* :196 */
                float g = ((c[k] + off) - lo[k]) / res;                 /* :197 */
                cell[k] = (int64_t)(g + 0.5f);                          /* :198, trunc toward zero */
                ok &= (cell[k] > 0) & (cell[k] < grid_res[k]);          /* :200, cell 0 is excluded */
            }
            if (!ok) continue;
            int64_t lin = cell[0] * gy * gz + cell[1] * gz + cell[2];   /* :203 */
#pragma omp atomic
            grid[lin] += 1;
            ++landed;
        }
    }
    return landed;
}

/* np.argmax(grid) -- first maximum in C order (train_dino.py:212); world = lo + cell*res in f64 (:213) */
EXPORT int64_t oracle_grid_argmax(const int64_t *grid, const int64_t grid_res[3], const float lo[3], double res,
                                  double world[3]) {
    int64_t G = grid_res[0] * grid_res[1] * grid_res[2], best = 0;
    for (int64_t i = 1; i < G; ++i)
        if (grid[i] > grid[best]) best = i;
    int64_t cz = best % grid_res[2], cy = (best / grid_res[2]) % grid_res[1], cx = best / (grid_res[2] * grid_res[1]);
    world[0] = (double)lo[0] + (double)cx * res;
    world[1] = (double)lo[1] + (double)cy * res;
    world[2] = (double)lo[2] + (double)cz * res;
    return best;
}

/* ------------------------------------------------------------------------------------------------
 * generate_target_pairs (dataset.py:118-135).  pairs [T,2,3] float32.  dtype flow of the reference
 * when fed float32 pairs: pdist and pdist_unit stay float32 (numpy norm, +1e-7 as a weak scalar),
 * everything that touches `center` (float64) or the integer axes is float64; results cast to f32.
 * axes are given in the POSITIONAL order of the signature (up, right, front) -- the eval.py call
 * sites pass (cfg.up, cfg.front, cfg.right), the caller of this function does the same.
 * ---------------------------------------------------------------------------------------------- */
EXPORT void oracle_generate_targets(const float *pairs, int64_t T, const double axes[9], const double center[3],
                                    float *tr_out, float *rot_out) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t) {
        const float *a = pairs + 6 * t, *b = a + 3;
        float pd[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
        float nrm = norm3_numpy(pd[0], pd[1], pd[2]) + 1e-7f;
        float u[3] = {pd[0] / nrm, pd[1] / nrm, pd[2] / nrm};
        double am[3] = {(double)a[0] - center[0], (double)a[1] - center[1], (double)a[2] - center[2]};
        double proj = (am[0] * (double)u[0] + am[1] * (double)u[1]) + am[2] * (double)u[2];
        double oc[3] = {am[0] - proj * (double)u[0], am[1] - proj * (double)u[1], am[2] - proj * (double)u[2]};
        double dist = sqrt((oc[0] * oc[0] + oc[1] * oc[1]) + oc[2] * oc[2]);
        tr_out[2 * t + 0] = (float)proj;
        tr_out[2 * t + 1] = (float)dist;
        if (rot_out)
            for (int k = 0; k < 3; ++k) {
                const double *ax = axes + 3 * k;
                double d = ((double)u[0] * ax[0] + (double)u[1] * ax[1]) + (double)u[2] * ax[2];
                rot_out[3 * t + k] = (float)acos(d);
            }
    }
}

/* ------------------------------------------------------------------------------------------------
 * decode (eval.py:228-235) with the multinomial draws injected: bins [T,6] uint8 in 0..num_bins-1.
 * pred = bin/(num_bins-1) - 0.5 (torch f32); pair scale = numpy-norm(real pair) / max(torch-norm(pred
 * pair), 1e-7); scaled = pred*scale.  Outputs: pred [T,2,3], scaled [T,2,3], pair_scale [T].
 * ---------------------------------------------------------------------------------------------- */
EXPORT void oracle_decode_pairs(const float *pc, const int64_t *idx, int64_t idx_stride, const uint8_t *bins, int64_t T,
                                int num_bins, float *pred, float *scaled, float *pair_scale) {
    const float denom = (float)(num_bins - 1);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t) {
        float p[6];
        for (int k = 0; k < 6; ++k) p[k] = (float)bins[6 * t + k] / denom - 0.5f;
        const float *a = pc + 3 * idx[t * idx_stride], *b = pc + 3 * idx[t * idx_stride + 1];
        float real = norm3_numpy(b[0] - a[0], b[1] - a[1], b[2] - a[2]);        /* eval.py:233 */
        float pn = norm3_torch(p[3] - p[0], p[4] - p[1], p[5] - p[2]);
        float s = real / (pn < 1e-7f ? 1e-7f : pn);
        pair_scale[t] = s;
        for (int k = 0; k < 6; ++k) {
            if (pred) pred[6 * t + k] = p[k];
            scaled[6 * t + k] = p[k] * s;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * vote_rotation (train_dino.py:218-239).  Writes candidates for every pair (masked pairs get
 * zeros and mask=0; the Python wrapper compacts them the way the reference's boolean indexing does).
 * tan(theta) is evaluated in double and rounded to float: torch-CPU's Sleef tanf is within 1 ulp of
 * that, the CUDA kernel does the same, so kernel and oracle agree bit-for-bit on this stage while
 * the reference agrees to ~1e-7 relative (tolerance stated in tests/test_oracle_golden.py).
 * ---------------------------------------------------------------------------------------------- */
static inline int rotation_candidates(const float *a, const float *b, float theta, const float *cos_tab,
                                      const float *sin_tab, int R, float *up /* [R,3] or NULL */,
                                      float ab[3], float x[3], float y[3], float *tan_out) {
    if (!pair_frame(a, b, 1, ab, x)) return 0;
    cross_torch(x, ab, y);
    float tn = (float)tan((double)theta);
    *tan_out = tn;
    if (up) {
        float sg = tn > 0.0f ? 1.0f : -1.0f;                              /* tan == 0 -> -1 (:235) */
        for (int r = 0; r < R; ++r) {
            float v[3];
            for (int k = 0; k < 3; ++k) {
                float off = cos_tab[r] * x[k] + sin_tab[r] * y[k];
                v[k] = tn * off + sg * ab[k];
            }
            float nv = norm3_torch(v[0], v[1], v[2]);
            if (nv < 1e-7f) nv = 1e-7f;
            up[3 * r + 0] = v[0] / nv; up[3 * r + 1] = v[1] / nv; up[3 * r + 2] = v[2] / nv;
        }
    }
    return 1;
}

EXPORT int64_t oracle_vote_rotation(const float *pc, const int64_t *idx, int64_t idx_stride, const float *theta, int64_t M,
                                    const float *cos_tab, const float *sin_tab, int R, float *up /* [M,R,3] */,
                                    uint8_t *mask /* [M] */) {
    int64_t kept = 0;
#pragma omp parallel for schedule(static) reduction(+ : kept)
    for (int64_t m = 0; m < M; ++m) {
        float ab[3], x[3], y[3], tn;
        float *dst = up + (size_t)m * R * 3;
        int ok = rotation_candidates(pc + 3 * idx[m * idx_stride], pc + 3 * idx[m * idx_stride + 1], theta[m], cos_tab,
                                     sin_tab, R, dst, ab, x, y, &tn);
        mask[m] = (uint8_t)ok;
        if (!ok) memset(dst, 0, sizeof(float) * 3 * (size_t)R);
        kept += ok;
    }
    return kept;
}

/* ------------------------------------------------------------------------------------------------
 * get_topk_dir's histogram (eval.py:37-46), brute force over all S sphere points:
 *   counts[s] = sum_rows [ dot(pred_row, sphere_s) > float32(cos_thr) ] / wt_row
 * The reference forms the dots with a float32 GEMM (K=3; accumulation order is the BLAS's), divides
 * by the float64 weights, sums every 100k-row chunk in float64 and rounds the running total to
 * float32 per chunk.  Restated: dot = fma(p2,s2,fma(p1,s1,p0*s0)); float64 accumulation; the float32
 * rounding of the total is left to the caller.  `wt` may be NULL (unit weights) and is per row.
 * Rows of `pred` whose row_valid flag is 0 are skipped (the reference never materialises them).
 * ---------------------------------------------------------------------------------------------- */
EXPORT void oracle_sphere_hist(const float *pred, int64_t rows, const double *wt, const uint8_t *row_valid,
                               const float *sphere, int S, float cos_thr, double *counts /* [S], accumulated */) {
#pragma omp parallel
    {
        double *local = (double *)calloc((size_t)S, sizeof(double));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < rows; ++i) {
            if (row_valid && !row_valid[i]) continue;
            const float *p = pred + 3 * i;
            double w = wt ? 1.0 / wt[i] : 1.0;
            for (int s = 0; s < S; ++s) {
                const float *q = sphere + 3 * s;
                float d = fmaf(p[2], q[2], fmaf(p[1], q[1], p[0] * q[0]));
                if (d > cos_thr) local[s] += w;
            }
        }
#pragma omp critical
        for (int s = 0; s < S; ++s) counts[s] += local[s];
        free(local);
    }
}

/* Fused vote_rotation + get_topk_dir for kept pairs, never materialising [M,R,3]: what the CUDA
 * pipeline computes.  wt_pair is per PAIR (expanded x R by the reference, eval.py:283). */
EXPORT void oracle_rotation_hist(const float *pc, const int64_t *idx, int64_t idx_stride, const float *theta,
                                 const double *wt_pair, const uint8_t *pair_keep, int64_t M, const float *cos_tab,
                                 const float *sin_tab, int R, const float *sphere, int S, float cos_thr,
                                 double *counts) {
#pragma omp parallel
    {
        double *local = (double *)calloc((size_t)S, sizeof(double));
        float *up = (float *)malloc(sizeof(float) * 3 * (size_t)R);
#pragma omp for schedule(dynamic, 64)
        for (int64_t m = 0; m < M; ++m) {
            if (pair_keep && !pair_keep[m]) continue;
            float ab[3], x[3], y[3], tn;
            if (!rotation_candidates(pc + 3 * idx[m * idx_stride], pc + 3 * idx[m * idx_stride + 1], theta[m], cos_tab,
                                     sin_tab, R, up, ab, x, y, &tn))
                continue;
            double w = wt_pair ? 1.0 / wt_pair[m] : 1.0;
            for (int r = 0; r < R; ++r) {
                const float *p = up + 3 * r;
                for (int s = 0; s < S; ++s) {
                    const float *q = sphere + 3 * s;
                    float d = fmaf(p[2], q[2], fmaf(p[1], q[1], p[0] * q[0]));
                    if (d > cos_thr) local[s] += w;
                }
            }
        }
#pragma omp critical
        for (int s = 0; s < S; ++s) counts[s] += local[s];
        free(local);
        free(up);
    }
}

/* fibonacci_sphere (utils/util.py:191-207): float64 math, cast to float32 by the caller (eval.py:80) */
EXPORT void oracle_fibonacci_sphere(int samples, float *out /* [samples,3] */) {
    const double phi = M_PI * (3.0 - sqrt(5.0));
    for (int i = 0; i < samples; ++i) {
        double y = 1.0 - ((double)i / (double)(samples - 1)) * 2.0;
        double radius = sqrt(1.0 - y * y);
        double th = phi * (double)i;
        out[3 * i + 0] = (float)(cos(th) * radius);
        out[3 * i + 1] = (float)y;
        out[3 * i + 2] = (float)(sin(th) * radius);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Back-vote error (eval.py:252-256): float32 L2 norm over the 2 columns of targets_tr - targets_tr_back,
 * numpy order sqrt(d0*d0 + d1*d1).
 * ---------------------------------------------------------------------------------------------- */
EXPORT void oracle_backvote_errors(const float *tr, const float *tr_back, int64_t T, float *errs) {
    for (int64_t t = 0; t < T; ++t) {
        float d0 = tr[2 * t] - tr_back[2 * t], d1 = tr[2 * t + 1] - tr_back[2 * t + 1];
        errs[t] = sqrtf(d0 * d0 + d1 * d1);
    }
}
