// The batched ("whole frame") form of the hot path: every kernel of the per-instance loop (eval.py:153-372) launched ONCE
// per frame over all instances / (instance, branch) jobs, instead of once per job.
//
// A frame of 6 instances x 2 branches was ~290 launches of kernels that each filled a fraction of the GPU; here it is ~30
// launches whose grids carry a job dimension (blockIdx.y, or a flattened tile index for the tensor-core heads).  Everything
// that varies from frame to frame -- cloud pointers and sizes, category configuration, weights, seeds, output slots --
// lives in ONE device-resident table (FrameTable) that the host refreshes with a single small copy per frame; kernel
// parameters and launch dimensions depend only on capacities, so the launch sequence is identical for every frame and can
// be captured once in a CUDA graph and replayed.
#pragma once

#include "common.cuh"

#include <cuda_bf16.h>

namespace cppf {

constexpr int kFrameMaxInst = CPPF_FRAME_MAX_INSTANCES;
constexpr int kFrameMaxJobs = 2 * kFrameMaxInst;

// ---- SHOT search-grid workspace of one cloud (shot.cu) -------------------------------------------------------------
struct ShotGrid {        // device-resident search grid header
    float lo[3];
    float inv;           // 1 / cell edge
    int dim[3];
    int cells;
};

struct ShotWorkspace {
    cppf_grid_geom *bounds;
    ShotGrid *grid;
    int *cell_of, *cell_count, *cell_start, *cell_fill;
    float4 *sorted, *normals_sorted;
    double *lrf;         // [n][8] per sorted position: scatter matrix (6), weight sum, counts -> LRF axes (shot.cu, kLrf*)
};

size_t shot_carve(void *ws, int64_t n, ShotWorkspace *out);

// ---- one job of the tensor-core heads kernel (heads_tc.cu) ----------------------------------------------------------
namespace tc {
struct Args {
    int64_t rows;                       // tuples or points
    const float *x;                     // LoadRows source [rows][x_ld]
    int x_ld;
    const float *pc, *normal;           // tuple encoders
    const __nv_bfloat16 *point_feat;    // [n][gather_cols] bf16 (per-point program output)
    IdxView idx;
    int arity;
    const unsigned char *weights;       // slab stream of this program
    float *out0, *out1;
    __nv_bfloat16 *out_bf16;
    unsigned char *bins;                // non-null: out0 is not written; the logits epilogue draws one bin per (row, coord)
    const float *u01;                   // [rows][6] injected uniforms or nullptr (counter-based generator keyed by seed)
    unsigned long long seed;
    long long *prof;                    // optional [gridDim.x][64] cycle counters (cppf_debug_heads_tc_profile)
};

struct MultiArgs {                      // the jobs of one launch: same program (architecture), different rows / weights
    int n_jobs;
    int pad;
    Args job[kFrameMaxInst];
};
}  // namespace tc

// ---- the per-frame table --------------------------------------------------------------------------------------------
struct FrameInst {                      // one detection
    const float *pc;
    int n, T;
    IdxView idx;                        // [T, >= 5] tuple indices (the caller's, or idx_draw once drawn)
    int32_t *idx_draw;                  // non-null: indices are drawn on the device (eval.py:207) into this buffer
    unsigned long long seed_idx;
    float normal_r, shot_r;
    float *shot_desc, *normals;         // null: no SHOT branch for this instance
    ShotWorkspace sw;
};

struct FrameJob {                       // one (instance, branch)
    int inst, branch;                   // branch: 0 = SHOT, 1 = DINO
    int scale_from;                     // job whose kept scale predictions give the median (eval.py:308-310: the SHOT branch
                                        // reuses the DINO branch's scale); its own index otherwise
    int up_loc, right_loc, loss_y_only, refine_iters;
    float res, gamma, refine_lr;
    double res64, imp_margin;
    double axes[9];
    long long rank_lo;
    const unsigned char *bins;          // [T,6] draws of this job
    const float *scales;                // [T,3] scale head output of this job
    uint32_t *grid;
    long long grid_capacity;
    cppf_grid_geom *geom;
    cppf_center *center;
    cppf_backvote_summary *summary;
    uint32_t *status;
    float *targets_tr, *targets_rot, *errs;
    uint8_t *keep;
    int32_t *kept_list, *imp;
    double *counts;
    void *ws_pose;
    cppf_pose *pose;
};

struct FrameTable {
    int n_inst, n_jobs;
    int any_refine, pad;
    int shot_base[kFrameMaxInst + 1];   // prefix sums of the point counts of the instances that run SHOT: the per-point SHOT
                                        // kernels spread one flat list of points over the grid
    FrameInst inst[kFrameMaxInst];
    FrameJob job[kFrameMaxJobs];
    tc::MultiArgs heads[2][2];          // [kind: 0 per-point program, 1 per-tuple program][branch]
};

struct FrameShared {                    // identical for every job of a frame; passed by value
    int R, S, num_bins, band, lut_g, replicas_max;
    int vote_lanes;                     // centre vote: lane-per-tuple form (vote_center.cu; CPPF_VOTE_LANES=0: warp-per-tuple)
    int rot_fast;                       // rotation vote: cheap lattice test with an exact slow path (rotation.cu; CPPF_ROT_FAST=0: exact only)
    float cos_thr;
    const uint2 *lut_cells;
    const float *cos_tab, *sin_tab, *sphere;
};

// Launchers of the batched stages, each in the file that owns the single-job kernels.  `t` is the DEVICE table; `ni` / `nj`
// are the instance / job counts the grids are sized for (the capacities when the sequence is graph-captured: inactive slots
// exit at once); T_cap / n_cap bound the per-job tuple and point counts.
int frame_launch_sample_tuples(const FrameTable *t, int ni, int64_t T_cap, cudaStream_t s);                       // targets.cu
int frame_launch_shot(const FrameTable *t, int ni, int64_t n_cap, cudaStream_t s);                                // shot.cu
int frame_launch_heads(const FrameTable *t, const void *const *tc_states, int ni, int64_t n_cap, int64_t T_cap, cudaStream_t s);   // heads_tc.cu


int frame_launch_center(const FrameTable *t, int nj, int64_t T_cap, const FrameShared &sh, cudaStream_t s);        // vote_center.cu
int frame_launch_backvote(const FrameTable *t, int nj, int64_t T_cap, cudaStream_t s);                            // backvote.cu
bool rotation_fast_enabled();                                                                                     // rotation.cu
int frame_launch_rotation(const FrameTable *t, int nj, int64_t T_cap, const FrameShared &sh, cudaStream_t s);     // rotation.cu
int frame_launch_pose(const FrameTable *t, int nj, int64_t T_cap, int any_refine, const FrameShared &sh, cudaStream_t s);  // pose.cu

}  // namespace cppf
