// One host call for the whole per-(instance, branch) vote chain -- the body of the reference's instance loop after the
// heads (eval.py:230-313, 358-363): decode -> vote targets -> centre vote -> arg-max -> back-vote filter -> rotation
// votes -> pose record.  Host-side glue only: it enqueues the kernels of the stage entry points on one stream; nothing
// returns to the host in between.  The stages stay individually callable (tuple-sharded runs interleave collectives).
#include "common.cuh"

using namespace cppf;

#define CPPF_TRY(expr)              \
    do {                            \
        const int _rc = (expr);     \
        if (_rc != CPPF_OK) return _rc; \
    } while (0)

CPPF_API int cppf_vote_chain(const float *pc, int64_t n, const void *idx, int idx_is_i64, int64_t idx_stride, int64_t T,
                             const uint8_t *bins, const float *pred_scales, const float *scale_override, int64_t cells_hint,
                             const cppf_vote_params *p, const cppf_vote_buffers *b, cppf_pose *pose_out, void *stream) {
    if (!pc || !idx || !bins || !p || !b || !pose_out || n <= 0 || T <= 0) return CPPF_ERR_INVALID_ARGUMENT;
    if (!pred_scales && !scale_override) return CPPF_ERR_INVALID_ARGUMENT;
    if (!b->grid || !b->geom || !b->center || !b->summary || !b->status || !b->targets_tr || !b->targets_rot || !b->errs ||
        !b->keep || !b->kept_list || !b->imp || !b->counts || !b->ws_backvote || !b->ws_pose)
        return CPPF_ERR_INVALID_ARGUMENT;
    if (!p->cos_tab || !p->sin_tab || !p->sphere || p->num_rots <= 0 || p->sphere_bins <= 0) return CPPF_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // decode (eval.py:230-240)
    CPPF_TRY(cppf_decode_targets(pc, idx, idx_is_i64, idx_stride, bins, T, p->num_bins, p->axes, b->targets_tr, b->targets_rot,
                                 nullptr, nullptr, stream));
    // centre vote + arg-max (train_dino.py:171-215)
    CPPF_TRY(cppf_cloud_bounds(pc, n, static_cast<float>(p->res), b->geom, stream));
    CPPF_CUDA_TRY(cudaMemsetAsync(b->status, 0, sizeof(uint32_t), s));
    CPPF_TRY(cppf_vote_center(pc, n, idx, idx_is_i64, idx_stride, b->targets_tr, T, p->cos_tab, p->sin_tab, p->num_rots, b->geom,
                              b->grid, b->grid_capacity, cells_hint, 0, b->status, stream));
    CPPF_TRY(cppf_grid_argmax(b->grid, b->grid_capacity, b->geom, p->res, b->status, b->center, stream));
    // back-vote filter (eval.py:251-275)
    CPPF_TRY(cppf_backvote_filter(pc, n, idx, idx_is_i64, idx_stride, b->targets_tr, T, p->axes, b->center, p->rank_lo, p->gamma,
                                  b->errs, b->keep, b->kept_list, b->imp, b->summary, b->ws_backvote, b->ws_backvote_bytes, stream));
    // rotation votes for the angle to `up` (column 0) and to `right` (column 2) (eval.py:277-293)
    CPPF_CUDA_TRY(cudaMemsetAsync(b->counts, 0, sizeof(double) * 2 * p->sphere_bins, s));
    const int cols[2] = {0, 2};
    CPPF_TRY(cppf_rotation_hist(pc, idx, idx_is_i64, idx_stride, b->targets_rot, 3, cols, 2, b->kept_list, &b->summary->kept, T,
                                b->imp, b->summary, p->imp_margin, p->cos_tab, p->sin_tab, p->num_rots, p->sphere, p->sphere_bins,
                                p->cos_thr, p->band, p->lut, p->lut_g, b->counts, stream));
    // pose assembly (eval.py:284-313), optional refinement (eval.py:319-355), branch loss (eval.py:358-363)
    return cppf_pose_finalize_refine(pc, idx, idx_is_i64, idx_stride, bins, p->num_bins, pred_scales, b->kept_list, b->summary, b->counts,
                                     p->sphere, p->sphere_bins, b->center, p->up_loc, p->right_loc, p->loss_y_only, scale_override,
                                     p->refine_iters, p->refine_lr, T, pose_out, b->ws_pose, b->ws_pose_bytes, stream);
}

// One host call per INSTANCE (eval.py:203-372 for one detection): SHOT-352 + normals, then per branch the heads with the decode
// fused in and the vote chain.  The DINO branch runs first; the SHOT branch reuses its scale (eval.py:308-310).  Host glue
// only -- about 60 kernel launches leave from one C call instead of ~10 foreign-function calls of the Python driver.
CPPF_API int cppf_instance_pose(const cppf_instance_io *io, const cppf_vote_params *p, const cppf_vote_buffers *b, void *stream) {
    if (!io || !p || !b || !io->pc || !io->idx || io->n <= 0 || io->T <= 0) return CPPF_ERR_INVALID_ARGUMENT;
    const bool run_dino = io->heads_dino != nullptr && io->dino_desc != nullptr;
    const bool run_shot = io->heads_shot != nullptr;
    if (!run_dino && !run_shot) return CPPF_ERR_INVALID_ARGUMENT;
    if (!io->bins || !io->scales || !io->ws_heads) return CPPF_ERR_INVALID_ARGUMENT;
    if (run_dino && !io->pose_dino) return CPPF_ERR_INVALID_ARGUMENT;
    if (run_shot && (!io->pose_shot || !io->shot_desc || !io->normals || !io->ws_shot)) return CPPF_ERR_INVALID_ARGUMENT;
    uint8_t *bins_dino = io->bins, *bins_shot = io->bins + io->T * 6;
    float *scale_dino = io->scales, *scale_shot = io->scales + io->T * 3;
    if (run_shot)   // eval.py:210; queued first so that its small kernels overlap the previous instance's tail
        CPPF_TRY(cppf_shot_compute_ex(io->pc, io->n, io->normal_r, io->shot_r, io->shot_desc, io->normals, nullptr, 1, nullptr, io->ws_shot,
                                      io->ws_shot_bytes, stream));
    if (run_dino) {
        CPPF_TRY(cppf_heads_forward_sampled(io->heads_dino, 1, io->pc, io->n, io->idx, io->idx_is_i64, io->idx_stride, io->T, io->dino_desc,
                                            nullptr, nullptr, io->seed_dino, bins_dino, scale_dino, io->ws_heads, io->ws_heads_bytes, stream));
        CPPF_TRY(cppf_vote_chain(io->pc, io->n, io->idx, io->idx_is_i64, io->idx_stride, io->T, bins_dino, scale_dino, nullptr,
                                 io->cells_hint, p, b, io->pose_dino, stream));
    }
    if (run_shot) {
        CPPF_TRY(cppf_heads_forward_sampled(io->heads_shot, 1, io->pc, io->n, io->idx, io->idx_is_i64, io->idx_stride, io->T, io->shot_desc,
                                            io->normals, nullptr, io->seed_shot, bins_shot, scale_shot, io->ws_heads, io->ws_heads_bytes, stream));
        CPPF_TRY(cppf_vote_chain(io->pc, io->n, io->idx, io->idx_is_i64, io->idx_stride, io->T, bins_shot, scale_shot,
                                 run_dino ? io->pose_dino->scale : nullptr, io->cells_hint, p, b, io->pose_shot, stream));
    }
    return CPPF_OK;
}
