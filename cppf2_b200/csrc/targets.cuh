// Decode + vote-target arithmetic shared by targets.cu (single-job kernels) and vote_center.cu (the batched frame path):
// eval.py:225-235 and generate_target_pairs (dataset.py:118-135).  Explicit _rn intrinsics only (compiled with -fmad=false).
#pragma once

#include "common.cuh"

namespace cppf {

struct Axes {
    double v[9];  // rows: positional (up, right, front) of dataset.py:118
};

__device__ __forceinline__ void targets_of_pair(const float a[3], const float b[3], const double center[3],
                                                const Axes &ax, float tr[2], float rot[3], bool want_rot) {
    const float pd0 = __fsub_rn(a[0], b[0]), pd1 = __fsub_rn(a[1], b[1]), pd2 = __fsub_rn(a[2], b[2]);
    const float nrm = __fadd_rn(norm3_numpy(pd0, pd1, pd2), 1e-7f);
    const double u0 = static_cast<double>(__fdiv_rn(pd0, nrm)), u1 = static_cast<double>(__fdiv_rn(pd1, nrm)),
                 u2 = static_cast<double>(__fdiv_rn(pd2, nrm));
    const double am0 = __dsub_rn(static_cast<double>(a[0]), center[0]), am1 = __dsub_rn(static_cast<double>(a[1]), center[1]),
                 am2 = __dsub_rn(static_cast<double>(a[2]), center[2]);
    const double proj = __dadd_rn(__dadd_rn(__dmul_rn(am0, u0), __dmul_rn(am1, u1)), __dmul_rn(am2, u2));
    const double oc0 = __dsub_rn(am0, __dmul_rn(proj, u0)), oc1 = __dsub_rn(am1, __dmul_rn(proj, u1)),
                 oc2 = __dsub_rn(am2, __dmul_rn(proj, u2));
    const double dist = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(oc0, oc0), __dmul_rn(oc1, oc1)), __dmul_rn(oc2, oc2)));
    tr[0] = static_cast<float>(proj);
    tr[1] = static_cast<float>(dist);
    if (want_rot) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double d = __dadd_rn(__dadd_rn(__dmul_rn(u0, ax.v[3 * k]), __dmul_rn(u1, ax.v[3 * k + 1])),
                                       __dmul_rn(u2, ax.v[3 * k + 2]));
            rot[k] = static_cast<float>(acos(d));
        }
    }
}

// decode (eval.py:230-235) + generate_target_pairs (dataset.py:118-135) of the tuples this CTA owns; `bid` / `nblk` are the
// CTA's index and count along the tuple dimension (blockIdx.x / gridDim.x in the single-job kernel)
__device__ __forceinline__ void decode_targets_body(const float *__restrict__ pc, const IdxView &idx,
                                                    const uint8_t *__restrict__ bins, int64_t T, int num_bins,
                                                    const Axes &ax, float *__restrict__ targets_tr,
                                                    float *__restrict__ targets_rot, float *__restrict__ pair_scale,
                                                    float *__restrict__ scaled_out, int bid, int nblk) {
    const float denom = static_cast<float>(num_bins - 1);
    const double zero[3] = {0.0, 0.0, 0.0};
    const int64_t stride = static_cast<int64_t>(nblk) * blockDim.x;
    for (int64_t t = static_cast<int64_t>(bid) * blockDim.x + threadIdx.x; t < T; t += stride) {
        float p[6];
#pragma unroll
        for (int k = 0; k < 6; ++k)  // bin/(num_bins-1) - 0.5  (eval.py:230)
            p[k] = __fsub_rn(__fdiv_rn(static_cast<float>(bins[6 * t + k]), denom), 0.5f);
        const int64_t ia = idx.at(t, 0), ib = idx.at(t, 1);
        const float a[3] = {pc[3 * ia], pc[3 * ia + 1], pc[3 * ia + 2]};
        const float b[3] = {pc[3 * ib], pc[3 * ib + 1], pc[3 * ib + 2]};
        // eval.py:233: numpy norm of (input_pairs[:,1]-input_pairs[:,0]) / clamp_min(torch norm of pred pair)
        const float real = norm3_numpy(__fsub_rn(b[0], a[0]), __fsub_rn(b[1], a[1]), __fsub_rn(b[2], a[2]));
        const float pn = norm3_torch(__fsub_rn(p[3], p[0]), __fsub_rn(p[4], p[1]), __fsub_rn(p[5], p[2]));
        const float s = __fdiv_rn(real, pn < 1e-7f ? 1e-7f : pn);
        float q[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) q[k] = __fmul_rn(p[k], s);
        if (pair_scale) pair_scale[t] = s;
        if (scaled_out) {
#pragma unroll
            for (int k = 0; k < 6; ++k) scaled_out[6 * t + k] = q[k];
        }
        float tr[2], rot[3];
        targets_of_pair(q, q + 3, zero, ax, tr, rot, targets_rot != nullptr);
        if (targets_tr) reinterpret_cast<float2 *>(targets_tr)[t] = make_float2(tr[0], tr[1]);
        if (targets_rot) {
            targets_rot[3 * t] = rot[0];
            targets_rot[3 * t + 1] = rot[1];
            targets_rot[3 * t + 2] = rot[2];
        }
    }
}

}  // namespace cppf
