// cppf_frame_pose: the instance loop of eval.py:153-372 over every detection of a frame as ~25 batched launches
// (frame.cuh).  Host glue only: fills the per-frame table in pinned memory, copies it to the device, and enqueues the
// batched stages in order; nothing returns to the host in between.
//
//   tuple sampling (1) -> SHOT-352 + normals of all clouds (7) -> heads: per-point programs DINO, SHOT, per-tuple programs
//   DINO, SHOT (4) -> bounds + zero (1) -> decode + targets + grid zero (1) -> centre votes (1) -> fold + arg-max (1) ->
//   back-vote errors (1), selection (1), mask (1) -> rotation votes (1) -> directions + scale median (1) [-> refinement
//   (1)] -> loss (1)
#include "frame.cuh"

#include <cstdlib>
#include <cstring>

using namespace cppf;

extern "C" const void *cppf_heads_tc_state(const cppf_heads *h);
extern "C" int cppf_heads_branch(const cppf_heads *h);
extern "C" int cppf_heads_tc_fill_job(const void *state, int kind, const float *pc, int64_t n, const void *idx, int idx_is_i64,
                                      int64_t idx_stride, int64_t T, const float *feat, const float *normal, void *point_feat,
                                      float *scale, unsigned char *bins, unsigned long long seed, void *args_out);
extern "C" int64_t cppf_heads_tc_point_bytes(const void *state, int64_t n);

CPPF_API int64_t cppf_frame_table_bytes(void) { return static_cast<int64_t>((sizeof(FrameTable) + 255) / 256 * 256); }

CPPF_API int64_t cppf_frame_heads_workspace_bytes(const cppf_heads *heads_dino, const cppf_heads *heads_shot, int64_t n) {
    if (n < 0) return 0;
    return cppf_heads_tc_point_bytes(cppf_heads_tc_state(heads_dino), n) + cppf_heads_tc_point_bytes(cppf_heads_tc_state(heads_shot), n) + 256;
}

static FrameShared shared_of(const cppf_vote_params *p, int replicas_max) {
    FrameShared sh{};
    sh.R = p->num_rots;
    sh.S = p->sphere_bins;
    sh.num_bins = p->num_bins;
    sh.band = p->band;
    sh.lut_g = p->lut_g;
    sh.replicas_max = replicas_max > 0 ? replicas_max : 1;      // measured on B200 with the run-length aggregated vote: 1 / 2 / 4 / 8 copies -> centre stage 0.413 / 0.423 / 0.437 / 0.482 ms per frame (before the aggregation 4 copies won: 0.630 / 0.584 / 0.573 / 0.615)
    static const int vote_lanes = [] {
        const char *e = getenv("CPPF_VOTE_LANES");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    sh.vote_lanes = vote_lanes;
    sh.rot_fast = (rotation_fast_enabled() && p->cos_thr > 0.0f) ? 1 : 0;      // the cheap test's band assumes a positive threshold
    sh.cos_thr = p->cos_thr;
    sh.lut_cells = p->lut ? reinterpret_cast<const uint2 *>(static_cast<const unsigned char *>(p->lut) + 16) : nullptr;
    sh.cos_tab = p->cos_tab;
    sh.sin_tab = p->sin_tab;
    sh.sphere = p->sphere;
    return sh;
}

// host: the frame's table.  Job 2i = DINO branch of instance i, 2i+1 = SHOT branch; jobs of absent branches are left out of
// the compacted job list (t->job[0 .. n_jobs)).
static int fill_table(const cppf_frame *f, FrameTable *t) {
    memset(t, 0, sizeof(FrameTable));
    if (f->n_instances < 0 || f->n_instances > kFrameMaxInst) return CPPF_ERR_INVALID_ARGUMENT;
    if (f->n_instances > 0 && (!f->io || !f->params || !f->buffers)) return CPPF_ERR_INVALID_ARGUMENT;
    t->n_inst = f->n_instances;
    int nj = 0;
    for (int i = 0; i < f->n_instances; ++i) {
        const cppf_instance_io &io = f->io[i];
        const cppf_vote_params &p = f->params[i];
        if (!io.pc || io.n <= 0 || io.T <= 0 || io.n > 0x7fffffffll || io.T > (1 << 17)) return CPPF_ERR_INVALID_ARGUMENT;
        if (!io.idx && !io.idx_draw) return CPPF_ERR_INVALID_ARGUMENT;
        if (io.idx && io.idx_stride < 5) return CPPF_ERR_INVALID_ARGUMENT;
        const bool run_dino = io.heads_dino != nullptr && io.dino_desc != nullptr;
        const bool run_shot = io.heads_shot != nullptr;
        if (!run_dino && !run_shot) return CPPF_ERR_INVALID_ARGUMENT;
        if (!io.bins || !io.scales || !io.ws_heads) return CPPF_ERR_INVALID_ARGUMENT;
        if (run_shot && (!io.shot_desc || !io.normals || !io.ws_shot || !io.pose_shot)) return CPPF_ERR_INVALID_ARGUMENT;
        if (run_dino && !io.pose_dino) return CPPF_ERR_INVALID_ARGUMENT;
        if (run_shot && io.ws_shot_bytes < cppf_shot_workspace_bytes(io.n)) return CPPF_ERR_WORKSPACE;
        if (io.ws_heads_bytes < cppf_frame_heads_workspace_bytes(run_dino ? io.heads_dino : nullptr, run_shot ? io.heads_shot : nullptr, io.n))
            return CPPF_ERR_WORKSPACE;
        if ((run_dino && !cppf_heads_tc_state(io.heads_dino)) || (run_shot && !cppf_heads_tc_state(io.heads_shot))) return CPPF_ERR_UNSUPPORTED;
        FrameInst &in = t->inst[i];
        in.pc = io.pc;
        in.n = static_cast<int>(io.n);
        in.T = static_cast<int>(io.T);
        if (io.idx) {
            in.idx = IdxView{io.idx, io.idx_stride, io.idx_is_i64};
            in.idx_draw = nullptr;
        } else {
            in.idx = IdxView{io.idx_draw, 5, 0};
            in.idx_draw = io.idx_draw;
            in.seed_idx = io.seed_idx;
        }
        in.normal_r = io.normal_r;
        in.shot_r = io.shot_r;
        in.shot_desc = run_shot ? io.shot_desc : nullptr;
        in.normals = run_shot ? io.normals : nullptr;
        if (run_shot) shot_carve(io.ws_shot, io.n, &in.sw);
        // heads: ws_heads = [DINO per-point table | SHOT per-point features]
        unsigned char *ws = static_cast<unsigned char *>(io.ws_heads);
        void *pf_dino = ws;
        void *pf_shot = ws + (run_dino ? cppf_heads_tc_point_bytes(cppf_heads_tc_state(io.heads_dino), io.n) : 0);
        uint8_t *bins[2] = {io.bins + io.T * 6, io.bins};            // [branch]: SHOT draws second, like cppf_instance_pose
        float *scales[2] = {io.scales + io.T * 3, io.scales};
        const int dino_job = run_dino ? nj : -1;
        for (int b = 1; b >= 0; --b) {                                  // DINO (1) first, then SHOT (0)
            if (b == 1 ? !run_dino : !run_shot) continue;
            const cppf_heads *h = b == 1 ? io.heads_dino : io.heads_shot;
            if (cppf_heads_branch(h) != b) return CPPF_ERR_INVALID_ARGUMENT;
            const void *st = cppf_heads_tc_state(h);
            void *pf = b == 1 ? pf_dino : pf_shot;
            for (int kind = 0; kind < 2; ++kind) {
                tc::MultiArgs &m = t->heads[kind][b];
                int rc = cppf_heads_tc_fill_job(st, kind, io.pc, io.n, in.idx.ptr, in.idx.is_i64, in.idx.stride, io.T,
                                                b == 1 ? io.dino_desc : io.shot_desc, b == 1 ? nullptr : io.normals, pf, scales[b],
                                                bins[b], b == 1 ? io.seed_dino : io.seed_shot, &m.job[m.n_jobs]);
                if (rc) return rc;
                ++m.n_jobs;
            }
            const cppf_vote_buffers &vb = f->buffers[2 * i + (b == 1 ? 0 : 1)];
            if (!vb.grid || !vb.geom || !vb.center || !vb.summary || !vb.status || !vb.targets_tr || !vb.targets_rot || !vb.errs ||
                !vb.keep || !vb.kept_list || !vb.imp || !vb.counts || !vb.ws_pose)
                return CPPF_ERR_INVALID_ARGUMENT;
            if (vb.ws_pose_bytes < cppf_pose_workspace_bytes(p.refine_iters > 0 ? io.T : 0)) return CPPF_ERR_WORKSPACE;
            FrameJob &j = t->job[nj];
            j.inst = i;
            j.branch = b;
            // eval.py:308-310: the SHOT branch keeps the DINO branch's scale when both run
            j.scale_from = (b == 0 && dino_job >= 0) ? dino_job : nj;
            j.up_loc = p.up_loc;
            j.right_loc = p.right_loc;
            j.loss_y_only = p.loss_y_only;
            j.refine_iters = p.refine_iters;
            j.res = static_cast<float>(p.res);
            j.gamma = p.gamma;
            j.refine_lr = p.refine_lr;
            j.res64 = p.res;
            j.imp_margin = p.imp_margin;
            for (int k = 0; k < 9; ++k) j.axes[k] = p.axes[k];
            j.rank_lo = p.rank_lo;
            j.bins = bins[b];
            j.scales = scales[b];
            j.grid = vb.grid;
            j.grid_capacity = vb.grid_capacity;
            j.geom = vb.geom;
            j.center = vb.center;
            j.summary = vb.summary;
            j.status = vb.status;
            j.targets_tr = vb.targets_tr;
            j.targets_rot = vb.targets_rot;
            j.errs = vb.errs;
            j.keep = vb.keep;
            j.kept_list = vb.kept_list;
            j.imp = vb.imp;
            j.counts = vb.counts;
            j.ws_pose = vb.ws_pose;
            j.pose = b == 1 ? io.pose_dino : io.pose_shot;
            if (p.refine_iters > 0) t->any_refine = 1;
            ++nj;
        }
    }
    t->n_jobs = nj;
    for (int i = 0; i < f->n_instances; ++i) t->shot_base[i + 1] = t->shot_base[i] + (t->inst[i].shot_desc ? t->inst[i].n : 0);
    for (int i = f->n_instances; i < kFrameMaxInst; ++i) t->shot_base[i + 1] = t->shot_base[i];
    return CPPF_OK;
}

CPPF_API int cppf_frame_pose(const cppf_frame *f, void *stream) {
    if (!f || !f->table_host || !f->table_dev || !f->shared) return CPPF_ERR_INVALID_ARGUMENT;
    if (f->mode <= 0 || f->mode > CPPF_FRAME_ALL) return CPPF_ERR_INVALID_ARGUMENT;
    const cppf_vote_params *p = f->shared;
    if (!p->cos_tab || !p->sin_tab || !p->sphere || p->num_rots <= 0 || p->sphere_bins <= 0) return CPPF_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    FrameTable *host = static_cast<FrameTable *>(f->table_host);
    const bool filled = (f->mode & CPPF_FRAME_FILL) != 0;
    if (filled) {
        const int rc = fill_table(f, host);
        if (rc) return rc;
        if (f->capacity_instances > 0 && host->n_inst > f->capacity_instances) return CPPF_ERR_INVALID_ARGUMENT;
        for (int i = 0; i < host->n_inst; ++i)
            if ((f->capacity_tuples > 0 && host->inst[i].T > f->capacity_tuples) || (f->capacity_points > 0 && host->inst[i].n > f->capacity_points))
                return CPPF_ERR_INVALID_ARGUMENT;
    }
    if (f->mode & CPPF_FRAME_COPY)
        CPPF_CUDA_TRY(cudaMemcpyAsync(f->table_dev, f->table_host, sizeof(FrameTable), cudaMemcpyHostToDevice, s));
    if (!(f->mode & CPPF_FRAME_LAUNCH)) return CPPF_OK;
    // launch sizing: the capacities when given (the sequence is then the same for every frame), else this frame's sizes
    int ni = f->capacity_instances;
    int64_t T_cap = f->capacity_tuples, n_cap = f->capacity_points;
    if (!filled) {
        if (ni <= 0 || T_cap <= 0 || n_cap <= 0) return CPPF_ERR_INVALID_ARGUMENT;
    } else {
        if (ni <= 0) ni = host->n_inst;
        if (T_cap <= 0)
            for (int i = 0; i < host->n_inst; ++i) T_cap = std::max<int64_t>(T_cap, host->inst[i].T);
        if (n_cap <= 0)
            for (int i = 0; i < host->n_inst; ++i) n_cap = std::max<int64_t>(n_cap, host->inst[i].n);
    }
    if (ni <= 0) return CPPF_OK;
    if (ni > kFrameMaxInst || T_cap > (1 << 17)) return CPPF_ERR_INVALID_ARGUMENT;
    const int nj = 2 * ni;
    const bool any_refine = p->refine_iters > 0 || (filled && host->any_refine != 0);
    const FrameTable *t = static_cast<const FrameTable *>(f->table_dev);
    const FrameShared sh = shared_of(p, f->replicas_max);
    const void *tc_states[2] = {cppf_heads_tc_state(f->heads_shot_any), cppf_heads_tc_state(f->heads_dino_any)};
    int rc, stage = 0;
    // CPPF_FRAME_SYNC=1 (debugging): synchronise after every stage and name the one whose kernels failed
    static const int sync_stages = [] {
        const char *e = getenv("CPPF_FRAME_SYNC");
        return e ? atoi(e) : 0;             // 1: synchronise per stage; 2: also report every completed stage
    }();
    static const char *const stage_name[] = {"begin", "tuple sampling", "SHOT", "heads", "centre vote", "back-vote", "rotation", "pose"};
    auto mark = [&]() -> int {
        if (sync_stages) {
            const cudaError_t e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) {
                fprintf(stderr, "[cppf_b200] cppf_frame_pose: stage '%s' failed: %s\n", stage_name[stage < 8 ? stage : 7], cudaGetErrorString(e));
                return CPPF_ERR_CUDA;
            }
            if (sync_stages > 1) fprintf(stderr, "[cppf_b200] cppf_frame_pose: stage '%s' done\n", stage_name[stage < 8 ? stage : 7]);
        }
        if (f->stage_events && f->stage_events[stage]) CPPF_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(f->stage_events[stage]), s));
        ++stage;
        return CPPF_OK;
    };
    if ((rc = mark())) return rc;
    if ((rc = frame_launch_sample_tuples(t, ni, T_cap, s)) || (rc = mark())) return rc;
    if ((tc_states[0] && (rc = frame_launch_shot(t, ni, n_cap, s))) || (rc = mark())) return rc;
    if ((rc = frame_launch_heads(t, tc_states, ni, n_cap, T_cap, s)) || (rc = mark())) return rc;
    if ((rc = frame_launch_center(t, nj, T_cap, sh, s)) || (rc = mark())) return rc;
    if ((rc = frame_launch_backvote(t, nj, T_cap, s)) || (rc = mark())) return rc;
    if ((rc = frame_launch_rotation(t, nj, T_cap, sh, s)) || (rc = mark())) return rc;
    if ((rc = frame_launch_pose(t, nj, T_cap, any_refine ? 1 : 0, sh, s))) return rc;
    return mark();
}
