/*
 * cppf_b200.h -- C ABI of libcppf_b200.so, the B200-native (sm_100a) pose-voting hot path of CPPF++.
 *
 * Every entry point replaces one operator of the reference (cited per function as file:line relative
 * to the reference root).  Conventions, identical for all calls:
 *   - plain C types only; pointers are DEVICE pointers unless the parameter name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; every call is stream-ordered, allocates nothing and
 *     never synchronises the host (workspaces are passed in; `cppf_*_workspace_bytes` size them);
 *   - the return value is 0 (CPPF_OK) or a CPPF_ERR_* code; cppf_error_string() names it.  Launch
 *     errors are reported by the call that made the launch; data-dependent conditions that only the
 *     device can see (grid larger than the caller's buffer) are written to the `status` word the
 *     caller passes and are sticky across calls;
 *   - thread-safe as long as two host threads do not share one workspace; no global mutable state.
 *   - indices are int64 (the reference's dtype) or int32, selected by `idx_is_i64`, with a row stride
 *     in elements so that `point_idxs_all[:, :2]` views (stride 5) can be passed without a copy.
 */
#ifndef CPPF_B200_H
#define CPPF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPPF_OK 0
#define CPPF_ERR_INVALID_ARGUMENT 1
#define CPPF_ERR_CUDA 2            /* a CUDA runtime call or kernel launch failed (sticky error logged) */
#define CPPF_ERR_WORKSPACE 3       /* workspace too small for the requested sizes */
#define CPPF_ERR_UNSUPPORTED 4     /* e.g. compute_color (SHOT1344): no caller in the reference */
#define CPPF_ERR_NO_DEVICE 5

/* bits of the device-side status word */
#define CPPF_STATUS_GRID_OVERFLOW 1u   /* gx*gy*gz exceeds the grid buffer handed to cppf_vote_center */
#define CPPF_STATUS_GRID_GUARD 2u      /* extent/res > 1000 on some axis: the reference skips the instance (eval.py:200) */
#define CPPF_STATUS_EMPTY 4u           /* no tuple survived a stage (empty cloud, no kept pair) */
#define CPPF_STATUS_REFINED 8u         /* R and t went through the online refinement (eval.py:319-355) */

#define CPPF_SHOT_DIM 352
#define CPPF_NUM_BINS 32

/* Geometry of the centre-vote grid, produced on the device by cppf_cloud_bounds and consumed by the
 * vote kernels without a host round trip (train_dino.py:172-173). */
typedef struct cppf_grid_geom {
    float lo[3];            /* pc.min(0) */
    float hi[3];            /* pc.max(0) */
    float res;              /* float32(res) */
    uint32_t flags;         /* CPPF_STATUS_* raised while building the geometry */
    int64_t grid_res[3];    /* trunc((hi-lo)/res) + 1 */
    int64_t cells;          /* gx*gy*gz */
} cppf_grid_geom;

/* Result of the centre vote: first maximum in C order and its world position (train_dino.py:212-213). */
typedef struct cppf_center {
    double world[3];        /* float64(lo) + cell*res */
    int64_t cell[3];
    int64_t linear;         /* flat C-order index of the arg-max cell */
    uint32_t votes;         /* count in that cell */
    uint32_t status;        /* CPPF_STATUS_* of the grid stage: geometry flags, overflow of the caller's grid buffer */
    int64_t cells;          /* gx*gy*gz of the grid that was (or would have been) voted */
} cppf_center;

/* Per-instance pose record written by cppf_pose_finalize (eval.py:284-313, :358-363). */
typedef struct cppf_pose {
    double R[9];            /* row-major R_est */
    double t[3];            /* T_est */
    float scale[3];         /* lower median of the kept per-tuple scale predictions */
    float scale_norm;       /* ||scale||  (RT[:3,:3] = R*scale_norm, scales = scale/scale_norm) */
    double loss;            /* mean clip(|canon - pred|, 0, 0.1) over kept pairs */
    int32_t bin_up, bin_right;
    float count_up, count_right;
    int64_t kept;           /* pairs surviving the back-vote filter */
    uint32_t status;        /* CPPF_STATUS_* of every stage: extent guard (eval.py:200), grid overflow, no kept pair, refined */
    uint32_t grid_cells;    /* gx*gy*gz, saturated at 2^32-1: with CPPF_STATUS_GRID_OVERFLOW the grid buffer the caller has to
                               provide before repeating the call */
} cppf_pose;

const char *cppf_error_string(int code);
int cppf_version(void);
/* Number of SMs / bytes of L2 of the current device, as the launch heuristics see them. */
int cppf_device_info(int *sm_count, int64_t *l2_bytes, int *cc_major, int *cc_minor);

/* ---- centre voting ------------------------------------------------------------------------------
 * replaces vote_center, train_dino.py:171-215 (integer Hough grid, bit-exact vs torch-CPU).          */

/* corners / grid_res of `pc [n,3]` -> *geom (device).  train_dino.py:172-173; guard eval.py:200. */
int cppf_cloud_bounds(const float *pc, int64_t n, float res, cppf_grid_geom *geom, void *stream);

/* Zeroes the grid and accumulates the T*R votes.  cos_tab/sin_tab [R] are the caller's torch-CPU
 * evaluation of cos/sin(arange(R)/R*2*pi) (train_dino.py:195-196; inputs because Sleef, libm and CUDA
 * differ in the last ulp).  grid is uint32 [grid_capacity]; the result occupies its first geom->cells
 * words, the rest is scratch: when capacity allows, up to 32 copies of the grid take the votes (spreading
 * the hot cache lines of the vote peak over L2 slices) and are folded into the first.  cells_hint > 0
 * (gx*gy*gz, when the caller knows it) lets grids that fit one SM's shared memory be privatised there.
 * `status` (uint32, device) receives CPPF_STATUS_GRID_OVERFLOW when geom->cells > grid_capacity (or >
 * cells_hint).  `accumulate` != 0 skips the zeroing and adds into the existing grid (chunked voting). */
int cppf_vote_center(const float *pc, int64_t n, const void *idx, int idx_is_i64, int64_t idx_stride,
                     const float *preds_tr, int64_t T, const float *cos_tab, const float *sin_tab, int R,
                     const cppf_grid_geom *geom, uint32_t *grid, int64_t grid_capacity, int64_t cells_hint,
                     int accumulate, uint32_t *status, void *stream);
/* Same with the strategy explicit (mode 0: L2 copies, replicas_max of them; mode 1: shared memory). */
int cppf_vote_center_ex(const float *pc, int64_t n, const void *idx, int idx_is_i64, int64_t idx_stride,
                        const float *preds_tr, int64_t T, const float *cos_tab, const float *sin_tab, int R,
                        const cppf_grid_geom *geom, uint32_t *grid, int64_t grid_capacity, int accumulate,
                        uint32_t *status, int mode, int replicas_max, int64_t smem_cells, void *stream);
int64_t cppf_vote_center_smem_cells(void);

/* First-maximum arg-max of the grid and its world position -> *center (device).  train_dino.py:212-213.
 * grid_capacity = words behind `grid`: a geometry with more cells than that was never voted (CPPF_STATUS_GRID_OVERFLOW)
 * and is not scanned.  `status` (device, may be NULL) = the word the vote call wrote; it is folded, with the geometry
 * flags, into center->status so that the pose record carries it. */
int cppf_grid_argmax(const uint32_t *grid, int64_t grid_capacity, const cppf_grid_geom *geom, double res,
                     const uint32_t *status, cppf_center *center, void *stream);

/* uint32 -> int64 widening of the first min(geom->cells, grid_capacity) cells (the reference returns an int64 numpy
 * grid, train_dino.py:204-206). */
int cppf_grid_to_i64(const uint32_t *grid, int64_t grid_capacity, const cppf_grid_geom *geom, int64_t *grid_i64,
                     void *stream);

/* ---- decode + vote targets ----------------------------------------------------------------------
 * replaces eval.py:225-235 (softmax, multinomial, pair scale) and generate_target_pairs,
 * dataset.py:118-135 (float64 where the reference is float64).                                        */

/* Tuple sampling, eval.py:207 (np.random.randint(0, N, (T, K)), with replacement): idx i32 [T,arity] drawn on the
 * device from a counter-based generator keyed by `seed`.  Host-sampled indices stay injectable in every call that
 * takes `idx`; this is the default of the frame driver when the caller supplies none. */
int cppf_sample_tuples(int64_t n, int64_t T, int arity, uint64_t seed, int32_t *idx, void *stream);

/* logits [T,6,num_bins] f32 -> bins uint8 [T,6].  One draw per (tuple, coordinate) by inverse CDF of
 * softmax(logits): with u01 [T,6] given the draw is the first bin whose cumulative probability exceeds
 * u; with u01 == NULL uniforms come from a counter-based generator keyed by (seed, tuple, coordinate). */
int cppf_sample_bins(const float *logits, int64_t T, int num_bins, const float *u01, uint64_t seed,
                     uint8_t *bins, void *stream);

/* bins -> pred_pairs (bin/(num_bins-1)-0.5), per-pair metric scale, scaled pairs, and the vote
 * targets of the scaled pairs w.r.t. the origin.  axes_host = {up, right, front} in the POSITIONAL
 * order of dataset.py:118 (the eval.py call sites pass cfg.up, cfg.front, cfg.right).
 * Outputs (any may be NULL): targets_tr f32 [T,2], targets_rot f32 [T,3], pair_scale f32 [T],
 * pred_pairs_scaled f32 [T,2,3]. */
int cppf_decode_targets(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride, const uint8_t *bins,
                        int64_t T, int num_bins, const double *axes_host, float *targets_tr, float *targets_rot,
                        float *pair_scale, float *pred_pairs_scaled, void *stream);

/* generate_target_pairs on explicit float32 pairs [T,2,3] with a float64 centre (device pointer, 3
 * doubles, or NULL for the origin).  dataset.py:118-135. */
int cppf_generate_targets(const float *pairs, int64_t T, const double *axes_host, const double *center,
                          float *targets_tr, float *targets_rot, void *stream);

/* ---- back-vote filter ---------------------------------------------------------------------------
 * replaces eval.py:251-275.  ws sized by cppf_backvote_workspace_bytes(T, n).                         */
int64_t cppf_backvote_workspace_bytes(int64_t T, int64_t n);

/* errs[t] = || targets_tr[t] - targets(real pair t, centre) ||  (float32, numpy order), then the exact
 * order statistics `rank_lo` and `rank_lo+1`, threshold = lerp (numpy 'linear' percentile, float32),
 * keep[t] = errs[t] < threshold, imp[p] = occurrences of point p among kept pair endpoints.
 * rank_lo / gamma come from the host (functions of T and backproj_ratio only).
 * Outputs: errs f32 [T], keep uint8 [T], kept_list int32 [T] (+ count in summary), imp int32 [n],
 * summary: {threshold f32, s_lo f32, s_hi f32, imp_max i32, kept i64} as cppf_backvote_summary. */
typedef struct cppf_backvote_summary {
    float threshold, s_lo, s_hi;
    int32_t imp_max;
    int64_t kept;
} cppf_backvote_summary;

int cppf_backvote_filter(const float *pc, int64_t n, const void *idx, int idx_is_i64, int64_t idx_stride,
                         const float *targets_tr, int64_t T, const double *axes_host, const cppf_center *center,
                         int64_t rank_lo, float gamma, float *errs, uint8_t *keep, int32_t *kept_list, int32_t *imp,
                         cppf_backvote_summary *summary, void *ws, int64_t ws_bytes, void *stream);

/* The three stages of cppf_backvote_filter, separately callable so that a tuple-sharded run can
 * all-gather `errs` before the selection and all-reduce `imp` before the maximum. */
int cppf_backvote_errors(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride, const float *targets_tr,
                         int64_t T, const cppf_center *center, float *errs, void *stream);
int cppf_backvote_select(const float *errs, int64_t T, int64_t rank_lo, float gamma, cppf_backvote_summary *summary,
                         void *ws, int64_t ws_bytes, void *stream);
int cppf_backvote_mask(const float *errs, const void *idx, int idx_is_i64, int64_t idx_stride, int64_t T, int64_t n,
                       cppf_backvote_summary *summary, uint8_t *keep, int32_t *kept_list, int32_t *imp,
                       int zero_outputs, void *stream);
int cppf_backvote_imp_max(const int32_t *imp, int64_t n, cppf_backvote_summary *summary, void *stream);

/* ---- rotation voting ----------------------------------------------------------------------------
 * replaces vote_rotation (train_dino.py:218-239) + get_topk_dir (eval.py:37-51) + fibonacci_sphere
 * (utils/util.py:191-207).                                                                            */

/* Materialising form, for the drop-in shim: up f32 [M,R,3] (rows of masked pairs zeroed), mask u8 [M]. */
int cppf_vote_rotation(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride, const float *preds_rot,
                       int64_t M, const float *cos_tab, const float *sin_tab, int R, float *up, uint8_t *mask,
                       void *stream);

/* counts[s] += sum over rows of [dot(pred_row, sphere_s) > cos_thr] / wt_row  (float64 bins).
 * pred f32 [rows,3]; wt f64 [rows] or NULL; band = half-width (in sphere indices) of the latitude band
 * searched around each row (cppf_sphere_band computes a safe value; band >= S means brute force). */
int cppf_sphere_hist(const float *pred, int64_t rows, const double *wt, const float *sphere, int S, float cos_thr,
                     int band, double *counts, void *stream);
int cppf_sphere_band(int S, float cos_thr);

/* Fused form used by the pipeline: for every kept tuple in kept_list (or all M when NULL) form the R
 * candidate directions for n_theta angle columns of `theta` [M, theta_stride] and vote them into
 * counts f64 [n_theta, S]; pair weight = imp[i]/imp_max + imp[j]/imp_max + margin when imp != NULL. */
int cppf_rotation_hist(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride, const float *theta,
                       int64_t theta_stride, const int *theta_cols_host, int n_theta, const int32_t *kept_list,
                       const int64_t *kept_count, int64_t M, const int32_t *imp, const cppf_backvote_summary *summary,
                       double margin, const float *cos_tab, const float *sin_tab, int R, const float *sphere, int S,
                       float cos_thr, int band, const void *lut, int lut_g, double *counts, void *stream);
/* Optional accelerator of the sphere test: a cube-map lookup table (6 x G x G cells, each listing the <= 4 lattice
 * points a direction inside the cell can hit -- a conservative superset, so the reference test eval.py:45 gives the
 * same hits as against all S points).  Built on the HOST from the host copy of `sphere`; the caller uploads the
 * cppf_sphere_lut_bytes(G) bytes and passes the device pointer as `lut` with `lut_g` = G (lut = NULL: latitude band).
 * cppf_sphere_lut_build returns CPPF_ERR_UNSUPPORTED when some cell needs more than 4 entries (use a finer G). */
int64_t cppf_sphere_lut_bytes(int G);
int cppf_sphere_lut_build(const float *sphere_host, int S, float cos_thr, int G, void *lut_host);
/* Same over the kept TUPLES whose id is congruent to `part` modulo `n_parts` (a partition by tuple id: the order of
 * kept_list is that of an atomic compaction and differs between runs and ranks); the [n_theta,S] bins of the parts
 * add up to the whole histogram. */
int cppf_rotation_hist_part(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride, const float *theta,
                            int64_t theta_stride, const int *theta_cols_host, int n_theta, const int32_t *kept_list,
                            const int64_t *kept_count, int64_t M, const int32_t *imp,
                            const cppf_backvote_summary *summary, double margin, const float *cos_tab,
                            const float *sin_tab, int R, const float *sphere, int S, float cos_thr, int band,
                            const void *lut, int lut_g, double *counts, int part, int n_parts, void *stream);

/* ---- pose assembly ------------------------------------------------------------------------------
 * replaces eval.py:284-313 (top-1 directions, Gram-Schmidt, scale median) and :358-363 (branch loss). */
int64_t cppf_pose_workspace_bytes(int64_t T);
int cppf_pose_finalize(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride, const uint8_t *bins,
                       int num_bins, const float *pred_scales, const int32_t *kept_list,
                       const cppf_backvote_summary *summary, const double *counts /* [2,S] */, const float *sphere,
                       int S, const cppf_center *center, int up_loc, int right_loc, int loss_y_only,
                       const float *scale_override /* 3 floats or NULL: eval.py:308 reuses the DINO scale */,
                       cppf_pose *pose, void *ws, int64_t ws_bytes, void *stream);

/* The scale median when the kept tuples are spread over several GPUs (tuple-sharded runs): torch.median(pred_scales[mask], 0)
 * (eval.py:309) as a two-pass, 16-bit-digit radix selection per axis.  Per pass every rank calls cppf_scale_median_hist on its
 * own kept tuples (hist u32 [3][65536], zeroed by the call; kept_count = device count of kept_list entries, kept_max its
 * host-side bound, sizing the launch), the ranks sum the histograms (integer all-reduce: exact), and every rank calls
 * cppf_scale_median_pick with the global kept count (device int32).  Pass 1 writes the three medians to scale_out (device),
 * which then goes to cppf_pose_finalize as scale_override.  With one rank the result equals cppf_pose_finalize's own median. */
typedef struct cppf_scale_select {
    uint32_t prefix[3];     /* key bits decided so far, per axis */
    uint32_t pad;
    uint64_t k[3];          /* rank still to resolve inside the prefix */
} cppf_scale_select;
int cppf_scale_median_hist(const float *pred_scales, const int32_t *kept_list, const int64_t *kept_count, int64_t kept_max,
                           int pass, const cppf_scale_select *state, uint32_t *hist, void *stream);
int cppf_scale_median_pick(const uint32_t *hist, const int32_t *kept_total, int pass, cppf_scale_select *state,
                           float *scale_out, void *stream);

/* Same with the reference's online refinement (eval.py:319-355, `opt=True`, its default) between the pose assembly and
 * the branch loss when refine_iters > 0: refine_iters Adam steps (reference: 100, lr 1e-2) on (t, quaternion) minimising
 * the L1 distance between the canonicalised kept pairs and the scaled predictions, in one single-CTA kernel.  R and t of
 * *pose are then the refined float32 values and CPPF_STATUS_REFINED is raised.  T = number of tuples (sizes the scratch:
 * ws_bytes >= cppf_pose_workspace_bytes(T)).  lietorch (SO3) is not part of the reference tree: its semantics are restated,
 * parity unpinned (oracle/refine_torch.py). */
int cppf_pose_finalize_refine(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride, const uint8_t *bins,
                              int num_bins, const float *pred_scales, const int32_t *kept_list,
                              const cppf_backvote_summary *summary, const double *counts, const float *sphere, int S,
                              const cppf_center *center, int up_loc, int right_loc, int loss_y_only,
                              const float *scale_override, int refine_iters, float refine_lr, int64_t T, cppf_pose *pose,
                              void *ws, int64_t ws_bytes, void *stream);

/* ---- instance cloud preparation (the step in front of SHOT) -----------------------
 * replaces backproject (utils/util.py:2586-2607, with the callers' un-flip of x and y, eval.py:185-189), downsample
 * (utils/util.py:39-46: Open3D voxel_down_sample_and_trace + np.random.choice) and the 50 000-point cap
 * (eval.py:195-198).  ws sized by cppf_cloud_workspace_bytes(n) with n = H*W resp. the number of points.  */
int64_t cppf_cloud_workspace_bytes(int64_t n);

/* depth [H,W] uint16 (depth_is_u16) or float32, divided by depth_div (1000 for REAL275 millimetres; 1 = as is);
 * mask u8 [H,W]; kinv_host = the 9 doubles of np.linalg.inv(intrinsics), row-major.  Outputs in row-major pixel
 * order (np.where): pc f32 [H*W,3] capacity, pix i32 [H*W] = row*W + col of each point, *count (device). */
int cppf_backproject(const void *depth, int depth_is_u16, double depth_div, const uint8_t *mask, int H, int W,
                     const double *kinv_host, float *pc, int32_t *pix, int64_t *count, void *ws, int64_t ws_bytes,
                     void *stream);

/* One member per occupied voxel of size `res` (Open3D grid: origin = min bound - res/2, float64).  The member with
 * the smallest (priority, index) wins: priority = prio[i] (f32 in [0,1), injectable) or a counter-based draw keyed
 * by `seed`.  Outputs in ascending index order: pc_out f32 [n,3] capacity, kept_idx i32 [n], *count (device);
 * side_in/side_out (i32 [n], optional) carry a per-point payload (the pixel index) along. */
int cppf_voxel_downsample(const float *pc, int64_t n, double res, const float *prio, uint64_t seed, float *pc_out,
                          int32_t *kept_idx, int64_t *count, const int32_t *side_in, int32_t *side_out, void *ws,
                          int64_t ws_bytes, void *stream);

/* pc_out[j] = pc[idx[j]], j < m (the 50 000-point cap: idx from cppf_sample_tuples(n, 50000, 1, ...)). */
int cppf_gather_points(const float *pc, const int32_t *idx, int64_t m, float *pc_out, const int32_t *side_in,
                       int32_t *side_out, void *stream);

/* ---- key-point descriptor sampling ------------------------------------------------
 * replaces interpolate_features, dataset.py:40-59: bilinear grid_sample (align_corners=False, zero padding) of the
 * patch-token map desc f32 [C,h,w] (element strides given, so the ViT's [h*w,C] layout is read in place) at the
 * key-points pts f32 [n,2] = (x, y) in pixels of the (cropped) image, `stride` = image pixels per token, followed by
 * F.normalize over the channels when `normalize`.  out f32 [n,C]. */
int cppf_interpolate_features(const float *desc, int C, int h, int w, int64_t stride_c, int64_t stride_h, int64_t stride_w,
                              const float *pts, int64_t n, float stride, int normalize, float *out, void *stream);

/* ---- the whole vote chain in one call ------------------------------------------------------------
 * The body of the instance loop after the heads, eval.py:230-313 and :358-363: cppf_decode_targets ->
 * cppf_cloud_bounds -> cppf_vote_center -> cppf_grid_argmax -> cppf_backvote_filter -> cppf_rotation_hist
 * (angle columns 0 and 2) -> cppf_pose_finalize, enqueued on `stream` from one host call.  Every pointer in both
 * structs is a DEVICE pointer to a caller-owned buffer except `axes` (host values).                   */
typedef struct cppf_vote_params {
    double res;                 /* cfg.res (eval.py:172): the Python float; the grid geometry uses float32(res), the voted
                                   centre lo + cell*res the float64 value (train_dino.py:213) */
    int num_rots, num_bins;     /* 180, 32 */
    int sphere_bins;            /* S = fibonacci lattice size (eval.py:79) */
    float cos_thr;              /* float32(cos(2*angle_tol)) */
    int band;                   /* cppf_sphere_band(S, cos_thr) */
    int lut_g;                  /* cube-map resolution of `lut` (ignored when lut == NULL) */
    int up_loc, right_loc, loss_y_only;
    const void *lut;            /* cppf_sphere_lut_build table or NULL */
    const float *cos_tab, *sin_tab;   /* [num_rots], torch-CPU values */
    const float *sphere;        /* [S,3] */
    double axes[9];             /* positional (up, right, front) of dataset.py:118 as passed at eval.py:237-240 */
    double imp_margin;          /* eval.py:275: 0.01 */
    int64_t rank_lo;            /* floor(ratio*(T-1)) */
    float gamma;                /* fractional part of ratio*(T-1) */
    int refine_iters;           /* 0 = opt=False; the reference's opt=True runs 100 (eval.py:326) */
    float refine_lr;            /* 1e-2 (eval.py:324) */
    int pad;
} cppf_vote_params;

typedef struct cppf_vote_buffers {
    uint32_t *grid;             /* [grid_capacity] */
    int64_t grid_capacity;
    cppf_grid_geom *geom;
    cppf_center *center;
    cppf_backvote_summary *summary;
    uint32_t *status;
    float *targets_tr;          /* [T,2] */
    float *targets_rot;         /* [T,3] */
    float *errs;                /* [T] */
    uint8_t *keep;              /* [T] */
    int32_t *kept_list;         /* [T] */
    int32_t *imp;               /* [n] */
    double *counts;             /* [2,S] */
    void *ws_backvote;
    int64_t ws_backvote_bytes;  /* cppf_backvote_workspace_bytes(T, n) */
    void *ws_pose;
    int64_t ws_pose_bytes;      /* cppf_pose_workspace_bytes(T) */
} cppf_vote_buffers;

/* bins u8 [T,6] are the multinomial draws (cppf_sample_bins / cppf_heads_forward_sampled / injected). */
int cppf_vote_chain(const float *pc, int64_t n, const void *idx, int idx_is_i64, int64_t idx_stride, int64_t T,
                    const uint8_t *bins, const float *pred_scales, const float *scale_override, int64_t cells_hint,
                    const cppf_vote_params *params, const cppf_vote_buffers *buffers, cppf_pose *pose_out, void *stream);

/* ---- one instance in one call ------------------------------------------------------------------------
 * The body of the instance loop, eval.py:203-372, for one detection: SHOT-352 + normals, then for the DINO branch
 * (when heads_dino and dino_desc are given) and the SHOT branch (when heads_shot is given): heads with the decode
 * fused in (cppf_heads_forward_sampled, precision 1) and cppf_vote_chain; the SHOT branch reuses the DINO branch's
 * scale (eval.py:308-310).  All pointers are device pointers to caller-owned buffers. */
struct cppf_heads;
typedef struct cppf_instance_io {
    const float *pc;            /* [n,3] */
    int64_t n;
    const void *idx;            /* [T, >= 5] tuple indices */
    int idx_is_i64;
    int pad0;
    int64_t idx_stride;
    int64_t T;
    const float *dino_desc;     /* [n,1024] or NULL */
    const struct cppf_heads *heads_dino, *heads_shot;   /* either may be NULL */
    float normal_r, shot_r;     /* 10 * res (eval.py:210) */
    float *shot_desc;           /* out [n,352] */
    float *normals;             /* out [n,3] */
    void *ws_shot;
    int64_t ws_shot_bytes;      /* cppf_shot_workspace_bytes(n) */
    uint8_t *bins;              /* out [2][T,6]: DINO draws, then SHOT draws */
    float *scales;              /* out [2][T,3] */
    void *ws_heads;
    int64_t ws_heads_bytes;     /* max over the branches of cppf_heads_workspace_bytes(h, T, n, 1) */
    uint64_t seed_dino, seed_shot;
    int64_t cells_hint;
    cppf_pose *pose_dino, *pose_shot;
    int32_t *idx_draw;          /* cppf_frame_pose only: with idx == NULL the tuple indices are drawn on the device (eval.py:207,
                                   cppf_sample_tuples semantics, seed_idx) into this int32 [T,5] buffer */
    uint64_t seed_idx;
} cppf_instance_io;

int cppf_instance_pose(const cppf_instance_io *io, const cppf_vote_params *params, const cppf_vote_buffers *buffers,
                       void *stream);

/* ---- one FRAME in one call -------------------------------------------------------------------------
 * The instance loop of eval.py:153-372 over every detection of a frame, batched: each stage (tuple sampling, SHOT, the
 * heads of both branches, decode, centre votes, back-vote filter, rotation votes, pose) is launched ONCE for all instances /
 * (instance, branch) jobs -- about 25 launches per frame whose grids carry a job dimension, instead of ~50 per instance.
 * Results are identical to cppf_instance_pose per instance (same kernels' arithmetic, same seeds).  Everything that changes
 * from frame to frame travels in one table (table_host -> table_dev, one small copy); launch dimensions depend on the
 * capacities only, so the kernel sequence (mode = CPPF_FRAME_LAUNCH) can be captured in a CUDA graph once and replayed
 * after refreshing the table (mode = CPPF_FRAME_FILL | CPPF_FRAME_COPY: fills table_host and queues its copy, no kernel).
 * table_host must stay untouched until its copy has executed: callers with several frames in flight rotate pinned buffers.
 * Requirements: bf16 tensor-core heads (cppf_heads_has_tc), T <= 2^17 per instance, 5-point tuples; ws_heads of an instance
 * holds the per-point tables of BOTH branches (cppf_frame_heads_workspace_bytes).  Job 2i is the DINO branch of instance i,
 * job 2i+1 its SHOT branch. */
#define CPPF_FRAME_MAX_INSTANCES 16
#define CPPF_FRAME_FILL 1        /* build the frame's table in table_host (host work only) */
#define CPPF_FRAME_COPY 2        /* queue the copy table_host -> table_dev */
#define CPPF_FRAME_LAUNCH 4      /* queue the kernels; reads the table on the device only */
#define CPPF_FRAME_ALL 7

typedef struct cppf_frame {
    int n_instances;                    /* 0 .. CPPF_FRAME_MAX_INSTANCES (used by CPPF_FRAME_FILL) */
    int mode;                           /* bit set of CPPF_FRAME_FILL / COPY / LAUNCH */
    const cppf_instance_io *io;         /* [n_instances], host */
    const cppf_vote_params *params;     /* [n_instances], host: the category configuration of each instance; the members that
                                           size kernels (num_rots, num_bins, sphere_bins, tables, lut) are taken from params[0] and
                                           from `shared` below and must agree across instances */
    const cppf_vote_buffers *buffers;   /* [2 * n_instances], host */
    const cppf_vote_params *shared;     /* the frame-wide members (tables, lattice, lut, thresholds, refine_iters > 0) */
    const struct cppf_heads *heads_dino_any, *heads_shot_any;   /* any model of each branch (the architecture sizes the heads
                                           launches); NULL: that branch is never run */
    void *table_host;                   /* pinned host scratch of cppf_frame_table_bytes() */
    void *table_dev;                    /* device scratch of the same size */
    int capacity_instances;             /* grids are sized for this many instances (0 with CPPF_FRAME_FILL: this frame's counts), ... */
    int replicas_max;                   /* centre-vote grid copies for L2-voted grids (0: default) */
    int64_t capacity_tuples;            /* ... this many tuples per instance ... */
    int64_t capacity_points;            /* ... and this many points per instance */
    void *const *stage_events;          /* optional (CPPF_FRAME_LAUNCH): CPPF_FRAME_STAGE_EVENTS cudaEvent_t handles recorded on the
                                           stream at the stage boundaries -- begin, after tuple sampling, SHOT, heads, centre vote,
                                           back-vote filter, rotation vote, pose -- for per-stage timing on the launching stream */
} cppf_frame;
#define CPPF_FRAME_STAGE_EVENTS 8

int64_t cppf_frame_table_bytes(void);
int64_t cppf_frame_heads_workspace_bytes(const struct cppf_heads *heads_dino, const struct cppf_heads *heads_shot, int64_t n);
int cppf_frame_pose(const cppf_frame *frame, void *stream);

/* ---- SHOT descriptor ----------------------------------------------------------------------------
 * replaces shot.compute / shot.estimate_normal, src_shot/shot.cpp:12-42, :45-100 (PCL NormalEstimation +
 * SHOTEstimation<SHOT352>, radius search, viewpoint at the origin, NaN rows for invalid points).       */
int64_t cppf_shot_workspace_bytes(int64_t n);
/* pc [n,3] f32 -> normals [n,3] f32 and (when desc != NULL) desc [n,352] f32. */
int cppf_shot_compute(const float *pc, int64_t n, float normal_r, float shot_r, float *desc, float *normals,
                      void *ws, int64_t ws_bytes, void *stream);
int cppf_estimate_normal(const float *pc, int64_t n, float normal_r, float *normals, void *ws, int64_t ws_bytes,
                         void *stream);
/* Same as cppf_shot_compute with three extras: rf_out [n,9] (rows x,y,z of the SHOT local reference
 * frame, NaN when invalid; may be NULL); fast_math != 0 evaluates the interpolation weights (acos, atan2)
 * in float instead of PCL's double (descriptor moves by ~1e-7: the interpolation is continuous); and
 * normals_in [n,3] (may be NULL) supplies the normals instead of estimating them. */
int cppf_shot_compute_ex(const float *pc, int64_t n, float normal_r, float shot_r, float *desc, float *normals,
                         float *rf_out, int fast_math, const float *normals_in, void *ws, int64_t ws_bytes, void *stream);
/* SHOT1344 (shot.cpp:102-161) has no caller in the reference: always CPPF_ERR_UNSUPPORTED. */
int cppf_shot_compute_color(const float *pc, const float *rgb, int64_t n, float normal_r, float shot_r, float *desc,
                            void *stream);

/* ---- learned heads ------------------------------------------------------------------------------
 * replaces BeyondCPPF.forward, train_shot.py:117-122 and train_dino.py:128-133 (ResLayer stacks,
 * train_shot.py:19-43), including prepare_tuple_inputs (train_shot.py:75-83, train_dino.py:91-97).
 * Weights are packed once by cppf_heads_pack (host arrays in state_dict order) into a device blob.     */
typedef struct cppf_heads cppf_heads;   /* opaque, device-resident packed weights */

/* branch: 0 = SHOT, 1 = DINO.  weights_host: concatenated float32 [out,in] weight then [out] bias of
 * every nn.Linear in the order of cppf2_b200.heads_spec.linear_shapes(); n_floats is checked. */
int cppf_heads_create(int branch, int num_more, const float *weights_host, int64_t n_floats, cppf_heads **out);
int cppf_heads_destroy(cppf_heads *h);
/* 1 when the bf16 tensor-core weights were packed (precision 1 available), else 0. */
int cppf_heads_has_tc(const cppf_heads *h);
int64_t cppf_heads_workspace_bytes(const cppf_heads *h, int64_t T, int64_t n, int precision);

/* precision: 0 = fp32 CUDA-core reference path, 1 = bf16 tensor-core (tcgen05) path.
 * SHOT branch: feat = shot descriptors [n,352] f32, normal [n,3] f32.
 * DINO branch: feat = descriptors [n,1024] f32, normal = NULL.
 * Outputs: logits f32 [T,6,32], scale f32 [T,3]. */
int cppf_heads_forward(const cppf_heads *h, int precision, const float *pc, int64_t n, const void *idx, int idx_is_i64,
                       int64_t idx_stride, int64_t T, const float *feat, const float *normal, float *logits,
                       float *scale, void *ws, int64_t ws_bytes, void *stream);

/* The same forward with the decode of eval.py:225-229 fused into the logits epilogue (precision 1 only, else
 * CPPF_ERR_UNSUPPORTED): softmax over the 32 bins of each of the 6 coordinates and one inverse-CDF draw per
 * (tuple, coordinate), uniforms from u01 f32 [T,6] or, when NULL, the counter-based generator of cppf_sample_bins
 * keyed by `seed`.  Writes bins u8 [T,6] and scale f32 [T,3]; the [T,6,32] logits never reach HBM. */
int cppf_heads_forward_sampled(const cppf_heads *h, int precision, const float *pc, int64_t n, const void *idx,
                               int idx_is_i64, int64_t idx_stride, int64_t T, const float *feat, const float *normal,
                               const float *u01, uint64_t seed, uint8_t *bins, float *scale, void *ws, int64_t ws_bytes,
                               void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CPPF_B200_H */
