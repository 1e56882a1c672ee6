// Shared description of the BeyondCPPF heads (reference train_shot.py:46-73, train_dino.py:58-85) for
// the fp32 and the tcgen05 kernels: layer inventory, flat-weight parsing, and the tuple-input encoder
// (prepare_tuple_inputs, train_shot.py:75-83 / train_dino.py:91-97).
#pragma once

#include <vector>

#include "common.cuh"

namespace cppf {

constexpr int kMaxResLayers = 8;

// One ResLayer (train_shot.py:19-43, bn/dropout disabled at inference):
//   h = relu(fc1 x + b1);  y = fc2 h + b2 + (fc0 x + b0  if din != dout else  x)
struct ResLayerDesc {
    int din, dout;
    int has_fc0;
    // offsets (in floats) into the caller's flat weight array
    int64_t w1, b1, w2, b2, w0, b0;
};

struct StackDesc {
    int n_layers;
    ResLayerDesc layer[kMaxResLayers];
    int in_dim() const { return layer[0].din; }
    int out_dim() const { return layer[n_layers - 1].dout; }
};

struct LinearDesc {
    int din, dout;
    int64_t w, b;
};

struct HeadsModel {
    int branch;      // 0 = SHOT, 1 = DINO
    int num_more;    // tuple arity = num_more + 2
    int arity, n_pairs;
    StackDesc shot_encoder;   // SHOT only (per point)
    StackDesc tuple_encoder, logit_encoder, scale_encoder;
    LinearDesc desc_transform, desc_pair_transform;   // DINO only
    int64_t n_floats;
};

// Mirrors cppf2_b200.heads_spec.linear_shapes(): every nn.Linear as weight [out,in] then bias [out], stacks
// in module order, fc1, fc2, (fc0) per ResLayer.
inline HeadsModel describe_heads(int branch, int num_more) {
    HeadsModel m{};
    m.branch = branch;
    m.num_more = num_more;
    m.arity = num_more + 2;
    m.n_pairs = m.arity * (m.arity - 1) / 2;
    int64_t off = 0;
    auto stack = [&](StackDesc &s, std::vector<int> dims) {
        s.n_layers = static_cast<int>(dims.size()) - 1;
        for (int i = 0; i < s.n_layers; ++i) {
            ResLayerDesc &l = s.layer[i];
            l.din = dims[i];
            l.dout = dims[i + 1];
            l.has_fc0 = l.din != l.dout;
            l.w1 = off; off += static_cast<int64_t>(l.dout) * l.din;
            l.b1 = off; off += l.dout;
            l.w2 = off; off += static_cast<int64_t>(l.dout) * l.dout;
            l.b2 = off; off += l.dout;
            if (l.has_fc0) {
                l.w0 = off; off += static_cast<int64_t>(l.dout) * l.din;
                l.b0 = off; off += l.dout;
            } else {
                l.w0 = l.b0 = -1;
            }
        }
    };
    int tuple_in;
    if (branch == 0) {
        stack(m.shot_encoder, {CPPF_SHOT_DIM, 128, 128, 128, 128, 128, 64});
        tuple_in = m.n_pairs * 4 + m.arity * 64;
    } else {
        tuple_in = m.n_pairs * 3 + 256;
    }
    stack(m.tuple_encoder, {tuple_in, 128, 128, 128, 128, 128, 256});
    stack(m.logit_encoder, {256, 256, 256, 192});
    stack(m.scale_encoder, {256, 128, 64, 3});
    if (branch == 1) {
        m.desc_transform = {1024, 256, off, 0};
        off += 256 * 1024;
        m.desc_transform.b = off;
        off += 256;
        m.desc_pair_transform = {256 * m.arity, 256, off, 0};
        off += static_cast<int64_t>(256) * 256 * m.arity;
        m.desc_pair_transform.b = off;
        off += 256;
    }
    m.n_floats = off;
    return m;
}

// ---- prepare_tuple_inputs -------------------------------------------------------------------------------
// Geometric part of the tuple encoding, float32 exactly as torch computes it on either device:
//   coords : for (i,j) in combinations(arity,2), lexicographic: p_i - p_j                (3 each)
//   normals: max(n_i . n_j, -(n_i) . n_j) with the dot as a plain left-to-right sum        (SHOT branch)
// `write(col, value)` receives column indices in the reference's concatenation order:
//   SHOT: [coords 3*P | normals P | feats arity*64];  DINO: [coords 3*P | pair 256]
template <typename W>
__device__ __forceinline__ void encode_tuple_geometry(const float *__restrict__ pc, const float *__restrict__ normal,
                                                      const int64_t *pt, int arity, bool with_normals, W &&write) {
    int col = 0;
    for (int i = 0; i < arity; ++i)
        for (int j = i + 1; j < arity; ++j) {
            const float *a = pc + 3 * pt[i], *b = pc + 3 * pt[j];
            write(col++, __fsub_rn(a[0], b[0]));
            write(col++, __fsub_rn(a[1], b[1]));
            write(col++, __fsub_rn(a[2], b[2]));
        }
    if (!with_normals) return;
    for (int i = 0; i < arity; ++i)
        for (int j = i + 1; j < arity; ++j) {
            float a[3], b[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {   // NaN normals of points with < 3 neighbours count as zeros (eval.py:216)
                const float av = normal[3 * pt[i] + k], bv = normal[3 * pt[j] + k];
                a[k] = (av == av) ? av : 0.0f;
                b[k] = (bv == bv) ? bv : 0.0f;
            }
            // torch.sum(n_i * n_j, -1) over 3 elements: ((x + y) + z); the second operand negates n_i first
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2]));
            const float e = __fadd_rn(__fadd_rn(__fmul_rn(-a[0], b[0]), __fmul_rn(-a[1], b[1])), __fmul_rn(-a[2], b[2]));
            write(col++, fmaxf(d, e));
        }
}

}  // namespace cppf
