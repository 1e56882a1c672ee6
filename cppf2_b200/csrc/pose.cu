// Pose assembly on B200 -- replaces eval.py:284-313 (top-1 directions, Gram-Schmidt, third axis by
// cross product, lower-median scale of the kept tuples) and eval.py:358-363 (branch selection loss).
// Three tiny stream-ordered kernels; nothing returns to the host until the caller reads cppf_pose.
#include "common.cuh"
#include "frame.cuh"

namespace cppf {

struct PoseScratch {  // head of the workspace, zeroed by the host wrapper
    double loss_sum;
    unsigned long long loss_cnt;
    unsigned int ticket;
    unsigned int pad;
};

// ---- directions + rotation matrix -------------------------------------------------------------------
__device__ __forceinline__ void pose_directions_body(const double *__restrict__ counts, const float *__restrict__ sphere, int S,
                                                     const cppf_center *__restrict__ center,
                                                     const cppf_backvote_summary *__restrict__ summary, int up_loc,
                                                     int right_loc, cppf_pose *__restrict__ pose) {
    // first arg-max of float32(counts) per angle column: the reference accumulates in float32 (eval.py:39)
    __shared__ float s_val[2][32];
    __shared__ int s_idx[2][32];
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
        const int n_w = static_cast<int>(blockDim.x >> 5);
        for (int w = 1; w < n_w; ++w)
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
    // every stage's flags end up in the one record the caller reads: the grid stage's (extent guard eval.py:200, overflow of
    // the grid buffer, empty cloud) arrive through the centre
    pose->status = ((summary && summary->kept == 0) ? CPPF_STATUS_EMPTY : 0u) | center->status;
    pose->grid_cells = center->cells > 0xffffffffll ? 0xffffffffu : static_cast<uint32_t>(center->cells);
}

__global__ void __launch_bounds__(256) pose_directions_kernel(const double *__restrict__ counts,
                                                              const float *__restrict__ sphere, int S,
                                                              const cppf_center *__restrict__ center,
                                                              const cppf_backvote_summary *__restrict__ summary,
                                                              int up_loc, int right_loc, cppf_pose *__restrict__ pose) {
    pose_directions_body(counts, sphere, S, center, summary, up_loc, right_loc, pose);
}

// ---- lower median of the kept scale predictions, one CTA per axis ------------------------------------
constexpr int kScaleRegKeys = 8;      // keys a thread keeps in registers across the four passes (kept <= 8 * 1024: one frame's ~5 000)

__device__ __forceinline__ void scale_median_body(const float *__restrict__ pred_scales, const int32_t *__restrict__ kept_list,
                                                  const cppf_backvote_summary *__restrict__ summary,
                                                  const float *__restrict__ scale_override, cppf_pose *__restrict__ pose,
                                                  int axis) {
    __shared__ uint32_t s_hist[256];
    __shared__ unsigned long long s_cum[256];
    __shared__ unsigned long long s_warp_tot[8];
    __shared__ int s_digit;
    const int64_t M = summary->kept;
    const int tid = threadIdx.x, lane = lane_id();
    if (scale_override || M <= 0) {
        if (tid == 0) pose->scale[axis] = scale_override ? scale_override[axis] : 0.0f;
        return;
    }
    // The keys are gathered once (two dependent L2 reads each) and stay in registers when the kept set is a frame's size;
    // thread tid owns the tuples u * blockDim.x + tid, so every warp iterates the same u (full-width ballots).
    const bool in_regs = M <= static_cast<int64_t>(kScaleRegKeys) * blockDim.x;
    uint32_t reg_key[kScaleRegKeys];
    auto load_key = [&](int64_t i) { return float_to_key(pred_scales[3 * static_cast<int64_t>(kept_list[i]) + axis]); };
    if (in_regs) {
#pragma unroll
        for (int u = 0; u < kScaleRegKeys; ++u) {
            const int64_t i = static_cast<int64_t>(u) * blockDim.x + tid;
            reg_key[u] = i < M ? load_key(i) : 0u;
        }
    }
    // the selection state lives in registers: every thread derives the same values from the histogram
    uint32_t prefix = 0u;
    unsigned long long k = static_cast<unsigned long long>((M - 1) / 2);  // torch.median: lower middle element
    for (int pass = 0; pass < 4; ++pass) {
        if (tid < 256) s_hist[tid] = 0u;
        if (tid == 0) s_digit = 256;
        __syncthreads();
        const int shift = 24 - 8 * pass;
        const uint32_t decided = pass == 0 ? 0u : ~((1u << (shift + 8)) - 1u);
        // scale predictions share their leading digits, so lanes with equal digits elect one leader that adds the group's size
        // instead of serialising on one shared-memory address (warp_hist_add)
        if (in_regs) {
#pragma unroll
            for (int u = 0; u < kScaleRegKeys; ++u) {
                const int64_t i = static_cast<int64_t>(u) * blockDim.x + tid;
                if (static_cast<int64_t>(u) * blockDim.x < M) {          // warp-uniform (and CTA-uniform)
                    const uint32_t key = reg_key[u];
                    warp_hist_add(s_hist, (key >> shift) & 0xffu, i < M && (key & decided) == (prefix & decided));
                }
            }
        } else {
            for (int64_t base = tid - lane; base < M; base += blockDim.x) {
                const int64_t i = base + lane;
                const bool live = i < M;
                const uint32_t key = live ? load_key(i) : 0u;
                warp_hist_add(s_hist, (key >> shift) & 0xffu, live && (key & decided) == (prefix & decided));
            }
        }
        __syncthreads();
        // the first digit whose cumulative count exceeds the remaining rank: inclusive scan of the 256 bins by warp shuffles
        // (a one-thread walk over the bins is 256 dependent shared-memory reads: ~4 us of a pass)
        unsigned long long incl = 0ull;
        if (tid < 256) {
            incl = s_hist[tid];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (lane == 31) s_warp_tot[tid >> 5] = incl;
        }
        __syncthreads();
        if (tid < 256) {
            for (int w = 0; w < (tid >> 5); ++w) incl += s_warp_tot[w];
            s_cum[tid] = incl;
            if (incl > k) atomicMin(&s_digit, tid);
        }
        __syncthreads();
        int d = s_digit;
        if (d == 256) d = 255;
        const unsigned long long cum = d > 0 ? s_cum[d - 1] : 0ull;
        prefix |= static_cast<uint32_t>(d) << shift;
        k -= cum;
        __syncthreads();                                  // s_cum / s_digit / s_hist are rewritten by the next pass
    }
    if (tid == 0) pose->scale[axis] = key_to_float(prefix);
}

__global__ void __launch_bounds__(1024) scale_median_kernel(const float *__restrict__ pred_scales,
                                                            const int32_t *__restrict__ kept_list,
                                                            const cppf_backvote_summary *__restrict__ summary,
                                                            const float *__restrict__ scale_override,
                                                            cppf_pose *__restrict__ pose) {
    scale_median_body(pred_scales, kept_list, summary, scale_override, pose, blockIdx.x);
}

// ---- the same lower median when the kept tuples are spread over several GPUs (tuple-sharded runs) -----------
// torch.median(pred_scales[mask], 0) (eval.py:309) is an exact order statistic per axis, so it is found the way the
// back-vote percentile is: a most-significant-digit radix selection over the order-preserving uint32 image of the float32
// values -- here two passes of 16 bits, so that a run costs two histogram exchanges.  Each rank histograms its own kept
// tuples (cppf_scale_median_hist), the ranks all-reduce the 3 x 65536 counters (integer sum: exact, order-independent), and
// every rank picks the same digit (cppf_scale_median_pick).  With one rank this is the single-GPU median, bit for bit.
constexpr int kScaleDigits = 1 << 16;

__global__ void __launch_bounds__(256) scale_hist_kernel(const float *__restrict__ pred_scales,
                                                         const int32_t *__restrict__ kept_list,
                                                         const int64_t *__restrict__ kept_count, int pass,
                                                         const cppf_scale_select *__restrict__ st, uint32_t *__restrict__ hist) {
    const int64_t M = *kept_count;
    const uint32_t p0 = st->prefix[0], p1 = st->prefix[1], p2 = st->prefix[2];
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int lane = lane_id();
    // whole warps iterate together: scale predictions share their leading digits, so lanes with equal digits elect one
    // leader that adds the group's size (one L2 atomic per distinct digit per warp instead of 32 on the same address)
    for (int64_t base = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x - lane; base < M; base += stride) {
        const int64_t i = base + lane;
        const bool live = i < M;
        const float *row = pred_scales + 3 * static_cast<int64_t>(live ? kept_list[i] : 0);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const uint32_t key = live ? float_to_key(row[a]) : 0u;
            const bool on = live && (pass == 0 || (key >> 16) == ((a == 0 ? p0 : (a == 1 ? p1 : p2)) >> 16));
            const uint32_t digit = pass == 0 ? (key >> 16) : (key & 0xffffu);
            warp_hist_add(hist + a * kScaleDigits, digit, on);
        }
    }
}

// one CTA per axis: the first digit whose cumulative count exceeds the remaining rank
__global__ void __launch_bounds__(1024) scale_pick_kernel(const uint32_t *__restrict__ hist, const int32_t *__restrict__ kept_total,
                                                          int pass, cppf_scale_select *__restrict__ st, float *__restrict__ scale_out) {
    __shared__ unsigned long long s_part[1024];
    __shared__ int s_owner;
    const int axis = blockIdx.x, tid = threadIdx.x;
    const int64_t M = *kept_total;
    if (M <= 0) {
        if (tid == 0) {
            st->prefix[axis] = 0u;
            st->k[axis] = 0ull;
            if (pass == 1 && scale_out) scale_out[axis] = 0.0f;
        }
        return;
    }
    const unsigned long long k = pass == 0 ? static_cast<unsigned long long>((M - 1) / 2) : st->k[axis];   // lower middle element
    const uint32_t *h = hist + axis * kScaleDigits + tid * 64;
    unsigned long long mine = 0;
    for (int j = 0; j < 64; ++j) mine += h[j];
    s_part[tid] = mine;
    if (tid == 0) s_owner = 1023;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {            // inclusive scan of the 1024 partial sums
        const unsigned long long add = tid >= o ? s_part[tid - o] : 0ull;
        __syncthreads();
        s_part[tid] += add;
        __syncthreads();
    }
    if (s_part[tid] > k) atomicMin(&s_owner, tid);
    __syncthreads();
    if (tid != s_owner) return;
    unsigned long long cum = tid > 0 ? s_part[tid - 1] : 0ull;
    int d = 0;
    for (; d < 63; ++d) {
        if (cum + h[d] > k) break;
        cum += h[d];
    }
    const uint32_t digit = static_cast<uint32_t>(tid * 64 + d);
    if (pass == 0) {
        st->prefix[axis] = digit << 16;
        st->k[axis] = k - cum;
    } else {
        const uint32_t key = (st->prefix[axis] & 0xffff0000u) | digit;
        st->prefix[axis] = key;
        if (scale_out) scale_out[axis] = key_to_float(key);
    }
}

// ---- online pose refinement (eval.py:319-355, `opt=True`) ---------------------------------------------
// 100 Adam steps (torch.optim.Adam defaults, lr 1e-2) on the translation (3) and a raw quaternion (x, y, z, w), started at
// (T_est, identity), minimising mean |((pc - t) @ (Q(q) R_est))[pair] - pred_pairs_scaled| over the kept pairs (the y
// column only for the symmetric categories).  lietorch semantics restated in oracle/refine_torch.py: Q is the rotation
// of the NORMALISED quaternion, its gradient is the left-perturbation tangent sum_i Q[:,i] x G[:,i] stored in slots
// x, y, z (slot w is zero, so w never moves), scaled by pi/180 before the optimiser step.
// The reference launches ~15 kernels per step; here one CTA carries the whole loop: the 2M (point, target) rows are
// compacted once into a scratch the L1 keeps, each step is one pass over them with 12 float64 sums (9 of d_j*sign_k,
// 3 of sign_k) reduced through shuffles + shared memory, and every thread repeats the scalar update redundantly, which
// saves the broadcast.  Per-element arithmetic is float32 like torch's; the sums are float64, so the result does not
// depend on the reduction order beyond one float32 rounding (the reference's own order is unspecified: cuBLAS + atomics).
struct __align__(8) RefineRow {
    float p[3];
    float y[3];
};

__device__ __forceinline__ void pose_refine_body(const float *__restrict__ pc, const IdxView &idx,
                                                 const uint8_t *__restrict__ bins, int num_bins,
                                                 const int32_t *__restrict__ kept_list,
                                                 const cppf_backvote_summary *__restrict__ summary, int loss_y_only, int iters,
                                                 float lr, RefineRow *__restrict__ rows, cppf_pose *__restrict__ pose) {
    __shared__ double s_part[16][12];
    __shared__ double s_tot[12];
    __shared__ float s_pose[12];                                // rot (9) and t (3) of the current step
    const int64_t M = summary->kept;
    if (M <= 0 || iters <= 0) return;
    const int64_t n_rows = 2 * M;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // (1) compact: point and scaled prediction of both ends of every kept pair (eval.py:231-235 for the scaling)
    const float denom = static_cast<float>(num_bins - 1);
    for (int64_t i = tid; i < M; i += blockDim.x) {
        const int64_t m = kept_list[i];
        const int64_t ia = idx.at(m, 0), ib = idx.at(m, 1);
        float p[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) p[k] = __fsub_rn(__fdiv_rn(static_cast<float>(bins[6 * m + k]), denom), 0.5f);
        const float a[3] = {pc[3 * ia], pc[3 * ia + 1], pc[3 * ia + 2]};
        const float b[3] = {pc[3 * ib], pc[3 * ib + 1], pc[3 * ib + 2]};
        const float real = norm3_numpy(__fsub_rn(b[0], a[0]), __fsub_rn(b[1], a[1]), __fsub_rn(b[2], a[2]));
        const float pn = norm3_torch(__fsub_rn(p[3], p[0]), __fsub_rn(p[4], p[1]), __fsub_rn(p[5], p[2]));
        const float sc = __fdiv_rn(real, pn < 1e-7f ? 1e-7f : pn);
        RefineRow ra, rb;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            ra.p[k] = a[k];
            rb.p[k] = b[k];
            ra.y[k] = __fmul_rn(p[k], sc);
            rb.y[k] = __fmul_rn(p[3 + k], sc);
        }
        rows[2 * i] = ra;
        rows[2 * i + 1] = rb;
    }
    __syncthreads();
    // (2) the optimiser state, replicated in every thread
    float R0[9], t[3], q[4] = {0.f, 0.f, 0.f, 1.f};
#pragma unroll
    for (int i = 0; i < 9; ++i) R0[i] = static_cast<float>(pose->R[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = static_cast<float>(pose->t[i]);
    float m_t[3] = {0.f, 0.f, 0.f}, v_t[3] = {0.f, 0.f, 0.f}, m_q[3] = {0.f, 0.f, 0.f}, v_q[3] = {0.f, 0.f, 0.f};
    const double beta1 = 0.9, beta2 = 0.999;
    double b1_pow = 1.0, b2_pow = 1.0;
    const float inv_cnt = 1.0f / (static_cast<float>(n_rows) * (loss_y_only ? 1.0f : 3.0f));
    float Q[9], rot[9];
    auto quat_matrix = [&](const float *qq, float *out) {      // column i = v + w (2 u x v) + u x (2 u x v), v = e_i
        const float n = sqrtf(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3]);
        const float u[3] = {qq[0] / n, qq[1] / n, qq[2] / n}, w = qq[3] / n;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float v[3] = {i == 0 ? 1.f : 0.f, i == 1 ? 1.f : 0.f, i == 2 ? 1.f : 0.f};
            float uv[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
#pragma unroll
            for (int k = 0; k < 3; ++k) uv[k] += uv[k];
            const float c2[3] = {u[1] * uv[2] - u[2] * uv[1], u[2] * uv[0] - u[0] * uv[2], u[0] * uv[1] - u[1] * uv[0]};
#pragma unroll
            for (int k = 0; k < 3; ++k) out[3 * k + i] = v[k] + w * uv[k] + c2[k];
        }
    };
    auto matmul3 = [](const float *A, const float *B, float *C) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    };
    // warp 0 owns the optimiser state and publishes (rot, t) through shared memory; the other warps only sweep rows.
    // Measured 6.7 us per step at 2M = 10 000 rows: ~3.7 us the row sweep (issue-bound: ~90 instructions per row, 16 warps on
    // 4 schedulers), ~2.3 us the dependent scalar chain (IEEE sqrt / divisions of Adam and of the quaternion), the rest the
    // two reductions.  More rows in flight per thread and float32 partial sums did not help (the latter also made the result
    // depend on the kept list's order).
    if (wid == 0) {
        quat_matrix(q, Q);
        matmul3(Q, R0, rot);                                   // rot = Q @ R_est
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) s_pose[k] = rot[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) s_pose[9 + k] = t[k];
        }
    }
    __syncthreads();
    for (int it = 1; it <= iters; ++it) {
#pragma unroll
        for (int k = 0; k < 9; ++k) rot[k] = s_pose[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = s_pose[9 + k];
        // float64 from the first add on: the kept list comes out of an atomic compaction in no particular order, and float64
        // sums of float32 terms make the result independent of it (float32 partials were seen to differ run to run)
        double acc[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) acc[k] = 0.0;
        for (int64_t i = tid; i < n_rows; i += blockDim.x) {
            const RefineRow r = rows[i];
            const float d[3] = {r.p[0] - t[0], r.p[1] - t[1], r.p[2] - t[2]};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (loss_y_only && k != 1) continue;
                const float c = d[0] * rot[k] + d[1] * rot[3 + k] + d[2] * rot[6 + k];
                const float e = c - r.y[k];
                const float sg = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);      // torch: d|x|/dx = sign(x), 0 at 0
                acc[9 + k] += static_cast<double>(sg);
#pragma unroll
                for (int j = 0; j < 3; ++j) acc[3 * j + k] += static_cast<double>(d[j] * sg);
            }
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) acc[k] = warp_sum(acc[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) s_part[wid][k] = acc[k];
        }
        __syncthreads();
        if (wid == 0) {
            if (lane < 12) {                                    // one fixed-order sum of the warps' partials per component
                const int n_w = blockDim.x >> 5;
                double v = 0.0;
                for (int w = 0; w < n_w; ++w) v += s_part[w][lane];
                s_tot[lane] = v;
            }
            __syncwarp();
            // gradients (float32 like the reference's autograd): g_rot = (pc - t)^T @ (sign / cnt), g_t = -sum_rows (sign / cnt) @ rot^T
            float g_rot[9], g_t[3], G[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) g_rot[k] = static_cast<float>(s_tot[k]) * inv_cnt;
#pragma unroll
            for (int j = 0; j < 3; ++j)
                g_t[j] = -(rot[3 * j] * static_cast<float>(s_tot[9]) + rot[3 * j + 1] * static_cast<float>(s_tot[10]) +
                           rot[3 * j + 2] * static_cast<float>(s_tot[11])) * inv_cnt;
            // dL/dQ = g_rot @ R_est^T
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) G[3 * i + j] = g_rot[3 * i] * R0[3 * j] + g_rot[3 * i + 1] * R0[3 * j + 1] + g_rot[3 * i + 2] * R0[3 * j + 2];
            float g_q[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 3; ++i) {                           // sum_i Q[:,i] x G[:,i]
                const float a0 = Q[i], a1 = Q[3 + i], a2 = Q[6 + i], b0 = G[i], b1 = G[3 + i], b2 = G[6 + i];
                g_q[0] += a1 * b2 - a2 * b1;
                g_q[1] += a2 * b0 - a0 * b2;
                g_q[2] += a0 * b1 - a1 * b0;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) g_q[k] = g_q[k] / 180.f * 3.14159274f;      // delta_rot.grad / 180 * np.pi
            // torch.optim.Adam, single-tensor path, float32 parameters
            b1_pow *= beta1;
            b2_pow *= beta2;
            const float step_size = static_cast<float>(static_cast<double>(lr) / (1.0 - b1_pow));
            const float bc2_sqrt = static_cast<float>(sqrt(1.0 - b2_pow));
            auto adam = [&](float &p, float &m, float &v, float g) {
                m = m + (g - m) * 0.1f;                              // exp_avg.lerp_(grad, 1 - beta1)
                v = v * 0.999f + 0.001f * g * g;                     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                const float den = sqrtf(v) / bc2_sqrt + 1e-8f;
                p = p - step_size * (m / den);
            };
#pragma unroll
            for (int k = 0; k < 3; ++k) adam(t[k], m_t[k], v_t[k], g_t[k]);
#pragma unroll
            for (int k = 0; k < 3; ++k) adam(q[k], m_q[k], v_q[k], g_q[k]);
            // q[3] has a zero gradient in every step: exp_avg = exp_avg_sq = 0, update 0 / (0 + eps) = 0
            quat_matrix(q, Q);
            matmul3(Q, R0, rot);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 9; ++k) s_pose[k] = rot[k];
#pragma unroll
                for (int k = 0; k < 3; ++k) s_pose[9 + k] = t[k];
            }
        }
        __syncthreads();
    }
    if (tid == 0) {                                             // warp 0 holds the final (rot, t)
        for (int i = 0; i < 9; ++i) pose->R[i] = static_cast<double>(rot[i]);
        for (int i = 0; i < 3; ++i) pose->t[i] = static_cast<double>(t[i]);
        pose->status |= CPPF_STATUS_REFINED;
    }
}

__global__ void __launch_bounds__(512, 1) pose_refine_kernel(const float *__restrict__ pc, IdxView idx,
                                                              const uint8_t *__restrict__ bins, int num_bins,
                                                              const int32_t *__restrict__ kept_list,
                                                              const cppf_backvote_summary *__restrict__ summary,
                                                              int loss_y_only, int iters, float lr,
                                                              RefineRow *__restrict__ rows, cppf_pose *__restrict__ pose) {
    pose_refine_body(pc, idx, bins, num_bins, kept_list, summary, loss_y_only, iters, lr, rows, pose);
}

// ---- branch loss: mean clip(|canon(pc[pair]) - pred_pairs|, 0, 0.1) over kept pairs (eval.py:358-363) -
__device__ __forceinline__ void pose_loss_body(const float *__restrict__ pc, const IdxView &idx,
                                               const uint8_t *__restrict__ bins, int num_bins,
                                               const int32_t *__restrict__ kept_list,
                                               const cppf_backvote_summary *__restrict__ summary, int loss_y_only,
                                               cppf_pose *__restrict__ pose, PoseScratch *__restrict__ scratch, int bid, int nblk) {
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
    const int64_t stride = static_cast<int64_t>(nblk) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(bid) * blockDim.x + threadIdx.x; i < 2 * M; i += stride) {
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
        s_last = (atomicAdd(&scratch->ticket, 1u) == static_cast<unsigned int>(nblk) - 1u);
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

__global__ void __launch_bounds__(256) pose_loss_kernel(const float *__restrict__ pc, IdxView idx,
                                                        const uint8_t *__restrict__ bins, int num_bins,
                                                        const int32_t *__restrict__ kept_list,
                                                        const cppf_backvote_summary *__restrict__ summary,
                                                        int loss_y_only, cppf_pose *__restrict__ pose,
                                                        PoseScratch *__restrict__ scratch) {
    pose_loss_body(pc, idx, bins, num_bins, kept_list, summary, loss_y_only, pose, scratch, blockIdx.x, gridDim.x);
}

// =====================================================================================================================
// Batched frame path (frame.cuh): pose assembly of every job in two launches (three with the refinement).
// =====================================================================================================================
// blockIdx.x = 0: top-1 directions + rotation matrix of the job; 1..3: lower median of the kept scale predictions, one axis
// each.  The SHOT branch of an instance takes the median over the DINO branch's kept tuples (scale_from; eval.py:308-310
// reuses the DINO scale) -- recomputed here rather than read from the other job's record, so that the jobs of one launch
// stay independent.
__global__ void __launch_bounds__(1024) frame_pose_dirs_scale_kernel(const FrameTable *__restrict__ t, FrameShared sh) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.y) >= t->n_jobs) return;
    const FrameJob &j = t->job[blockIdx.y];
    if (blockIdx.x == 0) {
        pose_directions_body(j.counts, sh.sphere, sh.S, j.center, j.summary, j.up_loc, j.right_loc, j.pose);
    } else {
        const FrameJob &src = t->job[j.scale_from];
        scale_median_body(src.scales, src.kept_list, src.summary, nullptr, j.pose, static_cast<int>(blockIdx.x) - 1);
    }
}

__global__ void __launch_bounds__(512, 1) frame_pose_refine_kernel(const FrameTable *__restrict__ t, FrameShared sh) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.x) >= t->n_jobs) return;
    const FrameJob &j = t->job[blockIdx.x];
    if (j.refine_iters <= 0) return;
    const FrameInst &in = t->inst[j.inst];
    RefineRow *rows = reinterpret_cast<RefineRow *>(static_cast<unsigned char *>(j.ws_pose) + 256);
    pose_refine_body(in.pc, in.idx, j.bins, sh.num_bins, j.kept_list, j.summary, j.loss_y_only, j.refine_iters, j.refine_lr, rows, j.pose);
}

__global__ void __launch_bounds__(256) frame_pose_loss_kernel(const FrameTable *__restrict__ t, FrameShared sh) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.y) >= t->n_jobs) return;
    const FrameJob &j = t->job[blockIdx.y];
    const FrameInst &in = t->inst[j.inst];
    pose_loss_body(in.pc, in.idx, j.bins, sh.num_bins, j.kept_list, j.summary, j.loss_y_only, j.pose,
                   static_cast<PoseScratch *>(j.ws_pose), blockIdx.x, gridDim.x);
}

int frame_launch_pose(const FrameTable *t, int nj, int64_t T_cap, int any_refine, const FrameShared &sh, cudaStream_t s) {
    (void)T_cap;
    if (nj <= 0) return CPPF_OK;
    CPPF_CUDA_TRY(launch_frame_kernel(frame_pose_dirs_scale_kernel, dim3(4, nj), dim3(1024), 0, s, t, sh));
    if (any_refine) {
        CPPF_CUDA_TRY(launch_frame_kernel(frame_pose_refine_kernel, dim3(nj), dim3(512), 0, s, t, sh));
    }
    const int per_job = std::max(4, device_info().sm_count * 2 / nj);
    CPPF_CUDA_TRY(launch_frame_kernel(frame_pose_loss_kernel, dim3(per_job, nj), dim3(256), 0, s, t, sh));
    return CPPF_OK;
}

}  // namespace cppf

using namespace cppf;

CPPF_API int64_t cppf_pose_workspace_bytes(int64_t T) {
    // PoseScratch + the refinement's compacted rows (2 per kept pair, at most T pairs)
    return 256 + static_cast<int64_t>(sizeof(RefineRow)) * 2 * (T > 0 ? T : 0);
}

CPPF_API int cppf_scale_median_hist(const float *pred_scales, const int32_t *kept_list, const int64_t *kept_count, int64_t kept_max,
                                    int pass, const cppf_scale_select *state, uint32_t *hist, void *stream) {
    if (!pred_scales || !kept_list || !kept_count || !state || !hist || pass < 0 || pass > 1 || kept_max < 0) return CPPF_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CPPF_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * 3 * kScaleDigits, s));
    if (kept_max == 0) return CPPF_OK;
    scale_hist_kernel<<<grid_for(kept_max, 256, 4), 256, 0, s>>>(pred_scales, kept_list, kept_count, pass, state, hist);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_scale_median_pick(const uint32_t *hist, const int32_t *kept_total, int pass, cppf_scale_select *state,
                                    float *scale_out, void *stream) {
    if (!hist || !kept_total || !state || pass < 0 || pass > 1 || (pass == 1 && !scale_out)) return CPPF_ERR_INVALID_ARGUMENT;
    scale_pick_kernel<<<3, 1024, 0, static_cast<cudaStream_t>(stream)>>>(hist, kept_total, pass, state, scale_out);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_pose_finalize(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride,
                                const uint8_t *bins, int num_bins, const float *pred_scales, const int32_t *kept_list,
                                const cppf_backvote_summary *summary, const double *counts, const float *sphere, int S,
                                const cppf_center *center, int up_loc, int right_loc, int loss_y_only,
                                const float *scale_override, cppf_pose *pose, void *ws, int64_t ws_bytes, void *stream) {
    return cppf_pose_finalize_refine(pc, idx, idx_is_i64, idx_stride, bins, num_bins, pred_scales, kept_list, summary, counts, sphere, S,
                                     center, up_loc, right_loc, loss_y_only, scale_override, 0, 0.0f, 0, pose, ws, ws_bytes, stream);
}

CPPF_API int cppf_pose_finalize_refine(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride,
                                       const uint8_t *bins, int num_bins, const float *pred_scales, const int32_t *kept_list,
                                       const cppf_backvote_summary *summary, const double *counts, const float *sphere, int S,
                                       const cppf_center *center, int up_loc, int right_loc, int loss_y_only,
                                       const float *scale_override, int refine_iters, float refine_lr, int64_t T, cppf_pose *pose,
                                       void *ws, int64_t ws_bytes, void *stream) {
    if (!pc || !idx || !bins || !kept_list || !summary || !counts || !sphere || !center || !pose || !ws)
        return CPPF_ERR_INVALID_ARGUMENT;
    if (!pred_scales && !scale_override) return CPPF_ERR_INVALID_ARGUMENT;
    if (up_loc < 0 || up_loc > 2 || right_loc < 0 || right_loc > 2 || up_loc == right_loc || S < 1 || idx_stride < 2)
        return CPPF_ERR_INVALID_ARGUMENT;
    if (refine_iters < 0 || (refine_iters > 0 && !(refine_lr > 0.0f))) return CPPF_ERR_INVALID_ARGUMENT;
    if (ws_bytes < cppf_pose_workspace_bytes(refine_iters > 0 ? T : 0)) return CPPF_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PoseScratch *scratch = static_cast<PoseScratch *>(ws);
    CPPF_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(PoseScratch), s));
    pose_directions_kernel<<<1, 256, 0, s>>>(counts, sphere, S, center, summary, up_loc, right_loc, pose);
    CPPF_LAUNCH_CHECK();
    scale_median_kernel<<<3, 1024, 0, s>>>(pred_scales, kept_list, summary, scale_override, pose);
    CPPF_LAUNCH_CHECK();
    IdxView iv{idx, idx_stride, idx_is_i64};
    if (refine_iters > 0) {      // eval.py:319-355: after R_est / scale are assembled, before the branch loss
        RefineRow *rows = reinterpret_cast<RefineRow *>(static_cast<unsigned char *>(ws) + 256);
        pose_refine_kernel<<<1, 512, 0, s>>>(pc, iv, bins, num_bins, kept_list, summary, loss_y_only, refine_iters, refine_lr, rows, pose);
        CPPF_LAUNCH_CHECK();
    }
    pose_loss_kernel<<<device_info().sm_count, 256, 0, s>>>(pc, iv, bins, num_bins, kept_list, summary, loss_y_only, pose,
                                                           scratch);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}
