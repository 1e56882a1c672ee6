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
#include "frame.cuh"

namespace cppf {

// CPPF_ROT_FAST=0: every candidate is normalised and tested with the reference's arithmetic (A/B and fallback; lut_vote_fast)
bool rotation_fast_enabled() {
    static const bool on = [] {
        const char *e = getenv("CPPF_ROT_FAST");
        return !(e && e[0] == '0');
    }();
    return on;
}

#ifndef CPPF_ROT_MINB
#define CPPF_ROT_MINB 4      // resident CTAs per SM the frame kernel's registers are capped for (5: 48 registers with spills, measured in session 3)
#endif

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

// the candidate direction before normalisation (train_dino.py:231-233)
__device__ __forceinline__ void rotation_direction(const RotFrame &f, float cr, float sr, float v[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float off = __fadd_rn(__fmul_rn(cr, f.x[k]), __fmul_rn(sr, f.y[k]));
        v[k] = __fadd_rn(__fmul_rn(f.tn, off), __fmul_rn(f.sg, f.ab[k]));
    }
}

// up = v / max(|v|, 1e-7) as torch computes it (train_dino.py:234-235)
__device__ __forceinline__ void rotation_normalise(const float v[3], float up[3]) {
    float nv = norm3_torch(v[0], v[1], v[2]);
    nv = nv < 1e-7f ? 1e-7f : nv;
    // three correctly rounded quotients by one divisor: one reciprocal + FMA corrections (div_by) instead of three IEEE
    // division sequences; |v[k]| <= nv, so every quotient is a normal number in [-1, 1] (or zero)
    const float inv = __frcp_rn(nv);
    up[0] = div_by(v[0], nv, inv);
    up[1] = div_by(v[1], nv, inv);
    up[2] = div_by(v[2], nv, inv);
}

__device__ __forceinline__ void rotation_candidate(const RotFrame &f, float cr, float sr, float up[3]) {
    float v[3];
    rotation_direction(f, cr, sr, v);
    rotation_normalise(v, up);
}

// Tests direction p against the lattice band and adds w to every bin it hits.
template <class Add>
__device__ __forceinline__ void band_vote(const float p[3], const float *__restrict__ s_sphere, int S,
                                          float cos_thr, int band, float half_sm1, Add &&add) {
    int lo = 0, hi = S - 1;
    if (band < S) {
        const int ic = __float2int_rn((1.0f - p[1]) * half_sm1);
        lo = max(ic - band, 0);
        hi = min(ic + band, S - 1);
    }
    for (int i = lo; i <= hi; ++i) {
        const float d = __fmaf_rn(p[2], s_sphere[3 * i + 2], __fmaf_rn(p[1], s_sphere[3 * i + 1], __fmul_rn(p[0], s_sphere[3 * i])));
        if (d > cos_thr) add(i);
    }
}

// Cube-map lookup of the lattice points a direction can possibly hit.  The table (cppf_sphere_lut_build) lists, for
// every cell of a 6 x G x G cube map, up to kLutCap lattice points within (cell radius + tolerance + margin) of the
// cell centre -- a superset of the points whose test can succeed -- so the exact reference test below runs on 0-4
// points instead of the 2*band+1 of the latitude band; the set of hits, hence the histogram, is the same.
constexpr int kLutCap = 4;
constexpr uint32_t kLutMagic = 0x4c555431u;   // "LUT1"

__device__ __forceinline__ int cube_cell(const float p[3], int G) {
    const float ax = fabsf(p[0]), ay = fabsf(p[1]), az = fabsf(p[2]);
    const int axis = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
    const float m = axis == 0 ? p[0] : (axis == 1 ? p[1] : p[2]);
    const float a = axis == 0 ? p[1] : (axis == 1 ? p[2] : p[0]);
    const float b = axis == 0 ? p[2] : (axis == 1 ? p[0] : p[1]);
    const float inv = 1.0f / fabsf(m);
    const float half_g = 0.5f * static_cast<float>(G);
    int iu = static_cast<int>((a * inv + 1.0f) * half_g);     // NaN / inf saturate and are clamped: such rows hit nothing anyway
    int iv = static_cast<int>((b * inv + 1.0f) * half_g);
    iu = min(max(iu, 0), G - 1);
    iv = min(max(iv, 0), G - 1);
    return ((2 * axis + (m < 0.0f ? 1 : 0)) * G + iv) * G + iu;
}

template <class Add>
__device__ __forceinline__ void lut_vote(const float p[3], const float *__restrict__ s_sphere, float cos_thr,
                                         const uint2 *__restrict__ lut_cells, int G, Add &&add) {
    const uint2 e = __ldg(lut_cells + cube_cell(p, G));
    const uint32_t word[2] = {e.x, e.y};
#pragma unroll
    for (int k = 0; k < kLutCap; ++k) {
        const uint32_t i = (word[k >> 1] >> (16 * (k & 1))) & 0xffffu;
        if (i == 0xffffu) break;     // entries are packed front to back
        const float d = __fmaf_rn(p[2], s_sphere[3 * i + 2], __fmaf_rn(p[1], s_sphere[3 * i + 1], __fmul_rn(p[0], s_sphere[3 * i])));
        if (d > cos_thr) add(static_cast<int>(i));
    }
}

// The same hit set without normalising the common candidate (session 3).  The reference's test is dot(v / |v|, s) > cos_thr in
// float32 with correctly rounded quotients -- ~30 of a candidate's ~160 instructions -- but its OUTCOME is decided by the
// unnormalised dot product for every (candidate, lattice point) pair that is not within rounding of the threshold:
// d' = dot(v, s) against cos_thr * |v|.  The two evaluations differ by < 1e-6 relative (|v| from MUFU.RSQ: 3e-7; the rounded
// quotients and the FMA chains: 4e-7), so outside a band of +-4e-6 around the threshold the cheap test is the reference's
// answer, and inside it (and for |v| near the 1e-7 clamp, infinities and NaNs) the exact sequence decides (the launchers use this
// form only for cos_thr > 0: the band's two bounds are ordered for a positive threshold).  The cube-map cell is
// taken from v directly (its formula is scale invariant; the table's 1e-3 rad margin covers the rounding).
template <class Add>
__device__ __forceinline__ void lut_vote_fast(const float v[3], const float *__restrict__ s_sphere, float cos_thr,
                                              const uint2 *__restrict__ lut_cells, int G, Add &&add) {
    const float n2 = __fmaf_rn(v[2], v[2], __fmaf_rn(v[1], v[1], __fmul_rn(v[0], v[0])));
    if (!(n2 > 1.1e-14f && n2 < 1e30f)) {           // rare: the exact path as it is
        float p[3];
        rotation_normalise(v, p);
        lut_vote(p, s_sphere, cos_thr, lut_cells, G, add);
        return;
    }
    const float t = cos_thr * (n2 * rsqrtf(n2));
    const float t_lo = t * 0.999996f, t_hi = t * 1.000004f;
    const uint2 e = __ldg(lut_cells + cube_cell(v, G));
    const uint32_t word[2] = {e.x, e.y};
#pragma unroll
    for (int k = 0; k < kLutCap; ++k) {
        const uint32_t i = (word[k >> 1] >> (16 * (k & 1))) & 0xffffu;
        if (i == 0xffffu) break;     // entries are packed front to back
        const float dv = __fmaf_rn(v[2], s_sphere[3 * i + 2], __fmaf_rn(v[1], s_sphere[3 * i + 1], __fmul_rn(v[0], s_sphere[3 * i])));
        if (dv > t_lo) {
            bool hit = dv > t_hi;
            if (!hit) {              // within rounding of the threshold: the reference's arithmetic decides
                float p[3];
                rotation_normalise(v, p);
                const float d = __fmaf_rn(p[2], s_sphere[3 * i + 2], __fmaf_rn(p[1], s_sphere[3 * i + 1], __fmul_rn(p[0], s_sphere[3 * i])));
                hit = d > cos_thr;
            }
            if (hit) add(static_cast<int>(i));
        }
    }
}

// A weighted bin as two 32-bit limbs of a 32.32 fixed-point sum: shared-memory atomicAdd on a double (and on a 64-bit
// integer) is a compare-and-swap loop -- a quarter of rotation_hist_kernel's stall samples (ncu) -- while the 32-bit integer
// add is one native ATOMS.ADD.  The low limb takes the fraction, a wrap of it carries into the high limb together with the
// integer part; the sum is exact, so it does not depend on the order the lanes arrive in.
__device__ __forceinline__ void fixed_add(unsigned long long *bin, unsigned long long v) {
    uint32_t *limb = reinterpret_cast<uint32_t *>(bin);          // little endian: [0] fraction, [1] integer part
    const uint32_t lo = static_cast<uint32_t>(v), hi = static_cast<uint32_t>(v >> 32);
    const uint32_t old = atomicAdd(&limb[0], lo);
    const uint32_t carry = (old + lo) < old ? 1u : 0u;
    if (hi + carry) atomicAdd(&limb[1], hi + carry);
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
        band_vote(p, s_sphere, S, cos_thr, band, half_sm1, [&](int i) { atomicAdd(&s_bins[i], w); });
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

__device__ __forceinline__ void rotation_hist_body(
    const float *__restrict__ pc, const IdxView &idx, const float *__restrict__ theta, int64_t theta_stride, const ThetaCols &cols,
    const int32_t *__restrict__ kept_list, const int64_t *__restrict__ kept_count, int64_t M,
    const int32_t *__restrict__ imp, const cppf_backvote_summary *__restrict__ summary, double margin,
    const float *__restrict__ cos_tab, const float *__restrict__ sin_tab, int R, const float *__restrict__ sphere, int S,
    float cos_thr, int band, const uint2 *__restrict__ lut_cells, int lut_g, double *__restrict__ counts, int part, int n_parts,
    int bid, int nblk, bool fast_test) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *s_bins = reinterpret_cast<unsigned long long *>(smem_raw);     // [n_theta][S], 32.32 fixed point
    float *s_sphere = reinterpret_cast<float *>(s_bins + cols.n * S);        // [S][3]
    float *s_cos = s_sphere + 3 * S;                                         // [R]
    float *s_sin = s_cos + R;
    for (int i = threadIdx.x; i < cols.n * S; i += blockDim.x) s_bins[i] = 0ull;
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
    const int64_t warp = (static_cast<int64_t>(bid) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(nblk) * blockDim.x) >> 5;
    // a partitioned run (cppf_rotation_hist_part) votes the TUPLES congruent to `part` modulo `n_parts`: the partition is by
    // tuple id, not by position in kept_list, whose order (atomic compaction) differs from run to run and rank to rank
    for (int64_t it = warp; it < n_items; it += n_warps) {
        const int64_t m = kept_list ? static_cast<int64_t>(kept_list[it]) : it;
        if (n_parts > 1 && m % n_parts != part) continue;      // warp-uniform
        const int64_t ia = idx.at(m, 0), ib = idx.at(m, 1);
        const float a[3] = {pc[3 * ia], pc[3 * ia + 1], pc[3 * ia + 2]};
        const float b[3] = {pc[3 * ib], pc[3 * ib + 1], pc[3 * ib + 2]};
        double w = 1.0;
        if (imp) {  // imp_wt = imp / imp.max(); pair weight = imp_wt[i] + imp_wt[j] + margin (eval.py:274-275)
            const double wi = __ddiv_rn(static_cast<double>(imp[ia]), imp_max);
            const double wj = __ddiv_rn(static_cast<double>(imp[ib]), imp_max);
            w = __drcp_rn(__dadd_rn(__dadd_rn(wi, wj), margin));
        }
        // the pair weight 1 / (imp_i + imp_j + margin) lies in (0, 1 / margin]: rounded once to 2^-32 (relative 5e-10 at
        // the smallest weight the reference can produce, 1 / 2.01); weights beyond 2^31 saturate
        const unsigned long long wv = __double2ull_rn(fmin(w, 2147483648.0) * 4294967296.0);
        // the pair frame (ab, x, y) is the same for every angle column; the float64 tangents of the columns are evaluated side
        // by side, lane c taking column c (the warp runs the tan() sequence once instead of once per column)
        RotFrame f;
        if (!pair_frame(a, b, true, f.ab, f.x)) continue;                     // warp-uniform (train_dino.py:222 mask)
        cross_torch(f.x, f.ab, f.y);
        const int my_col = (lane == 1 && cols.n > 1) ? cols.col[1] : ((lane == 2 && cols.n > 2) ? cols.col[2] : cols.col[0]);
        const float tn_lane = static_cast<float>(tan(static_cast<double>(theta[m * theta_stride + my_col])));
        for (int c = 0; c < cols.n; ++c) {
            f.tn = __shfl_sync(0xffffffffu, tn_lane, c);
            f.sg = f.tn > 0.0f ? 1.0f : -1.0f;  // torch.where(tan > 0, 1., -1.)
            unsigned long long *bins = s_bins + c * S;
            auto add = [&](int i) { fixed_add(&bins[i], wv); };
            if (lut_cells && fast_test) {
                for (int r = lane; r < R; r += 32) {     // (unroll 2 measured slower: 0.191 against 0.171 ms per frame)
                    float v[3];
                    rotation_direction(f, s_cos[r], s_sin[r], v);
                    lut_vote_fast(v, s_sphere, cos_thr, lut_cells, lut_g, add);
                }
            } else {
                for (int r = lane; r < R; r += 32) {
                    float p[3];
                    rotation_candidate(f, s_cos[r], s_sin[r], p);
                    if (lut_cells) lut_vote(p, s_sphere, cos_thr, lut_cells, lut_g, add);
                    else band_vote(p, s_sphere, S, cos_thr, band, half_sm1, add);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cols.n * S; i += blockDim.x)
        if (s_bins[i] != 0ull) atomicAdd(&counts[i], static_cast<double>(s_bins[i]) * (1.0 / 4294967296.0));
}

__global__ void __launch_bounds__(256) rotation_hist_kernel(
    const float *__restrict__ pc, IdxView idx, const float *__restrict__ theta, int64_t theta_stride, ThetaCols cols,
    const int32_t *__restrict__ kept_list, const int64_t *__restrict__ kept_count, int64_t M,
    const int32_t *__restrict__ imp, const cppf_backvote_summary *__restrict__ summary, double margin,
    const float *__restrict__ cos_tab, const float *__restrict__ sin_tab, int R, const float *__restrict__ sphere, int S,
    float cos_thr, int band, const uint2 *__restrict__ lut_cells, int lut_g, double *__restrict__ counts, int part, int n_parts,
    int fast_test) {
    rotation_hist_body(pc, idx, theta, theta_stride, cols, kept_list, kept_count, M, imp, summary, margin, cos_tab, sin_tab, R, sphere,
                       S, cos_thr, band, lut_cells, lut_g, counts, part, n_parts, blockIdx.x, gridDim.x, fast_test != 0);
}

// Batched frame path (frame.cuh): the rotation votes of every job (angle columns 0 and 2: `up` and `right`, eval.py:277-293)
// in one launch; blockIdx.y = job.
__global__ void __launch_bounds__(256, CPPF_ROT_MINB) frame_rotation_hist_kernel(const FrameTable *__restrict__ t, FrameShared sh) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.y) >= t->n_jobs) return;
    const FrameJob &j = t->job[blockIdx.y];
    const FrameInst &in = t->inst[j.inst];
    ThetaCols cols;
    cols.n = 2;
    cols.col[0] = 0;
    cols.col[1] = 2;
    cols.col[2] = 0;
    // CTAs beyond what the kept count can use leave at once (they would only zero and flush empty bins)
    const int64_t kept = j.summary->kept;
    const int64_t useful = (kept + 7) / 8 + 1;            // 8 warps per CTA, at least one kept tuple per warp
    const int nblk = static_cast<int>(useful < static_cast<int64_t>(gridDim.x) ? useful : gridDim.x);
    if (static_cast<int>(blockIdx.x) >= nblk) return;
    rotation_hist_body(in.pc, in.idx, j.targets_rot, 3, cols, j.kept_list, &j.summary->kept, in.T, j.imp, j.summary, j.imp_margin,
                       sh.cos_tab, sh.sin_tab, sh.R, sh.sphere, sh.S, sh.cos_thr, sh.band, sh.lut_cells, sh.lut_g, j.counts, 0, 1,
                       blockIdx.x, nblk, sh.rot_fast != 0);
}

int frame_launch_rotation(const FrameTable *t, int nj, int64_t T_cap, const FrameShared &sh, cudaStream_t s) {
    if (nj <= 0) return CPPF_OK;
    const size_t smem = 2 * static_cast<size_t>(sh.S) * sizeof(double) + 3 * static_cast<size_t>(sh.S) * sizeof(float) +
                        2 * static_cast<size_t>(sh.R) * sizeof(float);
    if (smem > static_cast<size_t>(device_info().max_smem_optin)) return CPPF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        CPPF_CUDA_TRY(cudaFuncSetAttribute(frame_rotation_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    // per job as many CTAs as the single-job launch (kept ~ T/10, one warp per kept tuple at a time), the whole launch a few
    // CTAs per SM
    const int64_t guess = T_cap / 8 + 1;
    const int per_job = std::max(1, std::min<int>(div_up(guess * 32, 256), std::max(8, device_info().sm_count * 4 / nj)));
    CPPF_CUDA_TRY(launch_frame_kernel(frame_rotation_hist_kernel, dim3(per_job, nj), dim3(256), smem, s, t, sh));
    return CPPF_OK;
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

// Host-side builder of the cube-map lookup table.  Layout: uint32 header[4] = {magic, G, cap, S}, then 6*G*G cells of
// `cap` uint16 lattice indices (0xffff = empty, packed front to back).  A lattice point s belongs to a cell when
// angle(cell centre, s) <= cell radius + acos(cos_thr) + margin, the margin (1e-3 rad) covering float32 rounding of the
// device-side cell assignment, of the dot product and of |p| != 1.  Returns CPPF_ERR_UNSUPPORTED when a cell would
// need more than `cap` entries (the caller retries with a finer G or falls back to the latitude band).
CPPF_API int64_t cppf_sphere_lut_bytes(int G) { return G < 1 ? 0 : 16 + static_cast<int64_t>(6) * G * G * kLutCap * 2; }

CPPF_API int cppf_sphere_lut_build(const float *sphere_host, int S, float cos_thr, int G, void *lut_host) {
    if (!sphere_host || !lut_host || S < 1 || S >= 0xffff || G < 1 || G > 1024) return CPPF_ERR_INVALID_ARGUMENT;
    uint32_t *hdr = static_cast<uint32_t *>(lut_host);
    hdr[0] = kLutMagic;
    hdr[1] = static_cast<uint32_t>(G);
    hdr[2] = kLutCap;
    hdr[3] = static_cast<uint32_t>(S);
    uint16_t *cells = reinterpret_cast<uint16_t *>(hdr + 4);
    const double tol = acos(fmin(1.0, fmax(-1.0, static_cast<double>(cos_thr)))) + 1e-3;
    auto dir_of = [](int axis, int neg, double u, double v, double out[3]) {
        double q[3];
        q[axis] = neg ? -1.0 : 1.0;
        q[(axis + 1) % 3] = u;
        q[(axis + 2) % 3] = v;
        const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
        out[0] = q[0] / n;
        out[1] = q[1] / n;
        out[2] = q[2] / n;
    };
    for (int face = 0; face < 6; ++face)
        for (int iv = 0; iv < G; ++iv)
            for (int iu = 0; iu < G; ++iu) {
                const int axis = face >> 1, neg = face & 1;
                const double u0 = -1.0 + 2.0 * iu / G, u1 = -1.0 + 2.0 * (iu + 1) / G;
                const double v0 = -1.0 + 2.0 * iv / G, v1 = -1.0 + 2.0 * (iv + 1) / G;
                double c[3], q[3];
                dir_of(axis, neg, 0.5 * (u0 + u1), 0.5 * (v0 + v1), c);
                double min_dot = 1.0;
                const double cu[4] = {u0, u1, u0, u1}, cv[4] = {v0, v0, v1, v1};
                for (int k = 0; k < 4; ++k) {
                    dir_of(axis, neg, cu[k], cv[k], q);
                    min_dot = fmin(min_dot, c[0] * q[0] + c[1] * q[1] + c[2] * q[2]);
                }
                const double reach = acos(fmin(1.0, fmax(-1.0, min_dot))) + tol;
                const double cos_reach = reach >= 3.141592653589793 ? -2.0 : cos(reach);
                uint16_t *cell = cells + (static_cast<size_t>(face * G + iv) * G + iu) * kLutCap;
                int n = 0;
                for (int k = 0; k < kLutCap; ++k) cell[k] = 0xffff;
                for (int i = 0; i < S; ++i) {
                    const double sx = sphere_host[3 * i], sy = sphere_host[3 * i + 1], sz = sphere_host[3 * i + 2];
                    const double sn = sqrt(sx * sx + sy * sy + sz * sz);
                    if (sn > 0.0 && (c[0] * sx + c[1] * sy + c[2] * sz) / sn < cos_reach) continue;
                    if (n == kLutCap) return CPPF_ERR_UNSUPPORTED;
                    cell[n++] = static_cast<uint16_t>(i);
                }
            }
    return CPPF_OK;
}

CPPF_API int cppf_rotation_hist_part(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride, const float *theta,
                                     int64_t theta_stride, const int *theta_cols_host, int n_theta, const int32_t *kept_list,
                                     const int64_t *kept_count, int64_t M, const int32_t *imp,
                                     const cppf_backvote_summary *summary, double margin, const float *cos_tab,
                                     const float *sin_tab, int R, const float *sphere, int S, float cos_thr, int band,
                                     const void *lut, int lut_g, double *counts, int part, int n_parts, void *stream) {
    if (!pc || !idx || !theta || !theta_cols_host || !cos_tab || !sin_tab || !sphere || !counts)
        return CPPF_ERR_INVALID_ARGUMENT;
    if (n_parts < 1 || part < 0 || part >= n_parts) return CPPF_ERR_INVALID_ARGUMENT;
    if (n_theta < 1 || n_theta > kMaxTheta || M < 0 || R <= 0 || S < 1 || idx_stride < 2) return CPPF_ERR_INVALID_ARGUMENT;
    if (imp && !summary) return CPPF_ERR_INVALID_ARGUMENT;
    if (lut && lut_g < 1) return CPPF_ERR_INVALID_ARGUMENT;
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
    // one warp per kept tuple at a time; the kept count is only known on the device, so size for M/8 (ratio 0.1).  With the
    // lookup table a tuple is cheap and the per-CTA bin flush dominates: two CTAs per SM, several tuples per warp.
    const int64_t guess = (kept_list ? (M / 8 + 1) : M) + 1;
    const uint2 *cells = lut ? reinterpret_cast<const uint2 *>(static_cast<const unsigned char *>(lut) + 16) : nullptr;
    // few kept tuples (one frame's T = 50 000 -> ~6 000): the per-CTA bin flush dominates, two CTAs per SM; many (the tuple
    // sweep's 10^5 .. 10^6): the kernel is issue-bound and wants all the warps the registers allow (four CTAs per SM)
    const int per_sm = (lut && guess < 32768) ? 2 : 4;
    rotation_hist_kernel<<<grid_for(guess * 32, 256, per_sm), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        pc, iv, theta, theta_stride, cols, kept_list, kept_count, M, imp, summary, margin, cos_tab, sin_tab, R, sphere, S,
        cos_thr, band, cells, lut_g, counts, part, n_parts, (rotation_fast_enabled() && cos_thr > 0.0f) ? 1 : 0);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_rotation_hist(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride, const float *theta,
                                int64_t theta_stride, const int *theta_cols_host, int n_theta, const int32_t *kept_list,
                                const int64_t *kept_count, int64_t M, const int32_t *imp,
                                const cppf_backvote_summary *summary, double margin, const float *cos_tab,
                                const float *sin_tab, int R, const float *sphere, int S, float cos_thr, int band,
                                const void *lut, int lut_g, double *counts, void *stream) {
    return cppf_rotation_hist_part(pc, idx, idx_is_i64, idx_stride, theta, theta_stride, theta_cols_host, n_theta, kept_list,
                                   kept_count, M, imp, summary, margin, cos_tab, sin_tab, R, sphere, S, cos_thr, band, lut, lut_g,
                                   counts, 0, 1, stream);
}
