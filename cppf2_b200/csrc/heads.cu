// BeyondCPPF heads on B200 -- replaces BeyondCPPF.forward (reference train_shot.py:117-122,
// train_dino.py:128-133): prepare_tuple_inputs + the ResLayer stacks.
//
// This translation unit owns the model object (cppf_heads: weights packed once on the device) and the
// float32 path (precision 0): a generic "ResLayer chain" kernel in which one CTA carries a tile of 32 rows
// (points or tuples) through a whole stack with the activations resident in shared memory -- nothing but
// the stack's input and output touches HBM, where the reference launches 3-5 kernels per nn.Linear and
// materialises every intermediate.  The tuple input [T,360] / [T,286] of the reference is
// never materialised either: the chain's prologue gathers points, normals and per-point features straight
// into the tile.  precision 1 (bf16 tcgen05, heads_tc.cu) shares the packing and the plans.
#include "heads_common.cuh"

#include <cstdlib>
#include <cstring>

namespace cppf {

constexpr int kTM = 32;           // rows per CTA
constexpr int kW0 = 368;          // width of buffer 0 (widest stack input: 360 -> padded)
constexpr int kW1 = 256;          // width of buffers 1, 2
constexpr int kChainThreads = 256;

struct F32Layer {                 // one ResLayer, weights transposed to [K][N] (K padded to 4) for coalesced reads
    int din, dout, din_pad, dout_pad;
    int has_fc0;
    const float *w1t, *b1;        // [din_pad][dout], [dout]
    const float *w2t;             // [dout_pad][dout]
    const float *w0t;             // [din_pad][dout] or null
    const float *b20;             // b2 (+ b0)
};

struct F32Chain {
    int n_layers;
    F32Layer layer[kMaxResLayers];
};

struct F32Linear {                // plain nn.Linear evaluated over `chunks` K-chunks of 256
    int chunks, dout;
    const float *wt;              // [chunks*256][dout]
    const float *b;
};

enum Prologue { kRowsFromGlobal = 0, kShotTuple = 1, kDinoTuple = 2 };

struct ChainArgs {
    int prologue;
    int64_t M;                    // rows (points or tuples)
    // kRowsFromGlobal
    const float *x;               // [M][x_dim]
    int x_dim;
    // tuple prologues
    const float *pc, *normal, *feat;   // feat: SHOT [n,64] encoded descriptors / DINO [n,256] transformed descriptors
    IdxView idx;
    int arity;
    F32Linear pair;               // DINO desc_pair_transform
    float *out;                   // [M][out_dim]
    int out_dim;
};

// dst[r][n] = act( sum_k src1[r][k] W1T[k][n] (+ sum_k src2[r][k] W2T[k][n]) + bias[n] (+ res[r][n]) )
template <int ROWS>
__device__ __forceinline__ void dense_step(const float *__restrict__ src1, int ld1, int K1, const float *__restrict__ w1t,
                                           const float *__restrict__ src2, int ld2, int K2, const float *__restrict__ w2t,
                                           const float *__restrict__ bias, const float *__restrict__ res, int ldr, bool relu,
                                           float *__restrict__ dst, int ldd, int N, int N_pad, int col, int row0) {
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.0f;
    for (int k = 0; k < K1; k += 4) {
        const float w0 = w1t[(k + 0) * N + col], w1 = w1t[(k + 1) * N + col], w2 = w1t[(k + 2) * N + col],
                    w3 = w1t[(k + 3) * N + col];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float4 a = *reinterpret_cast<const float4 *>(src1 + (row0 + r) * ld1 + k);
            acc[r] = fmaf(a.x, w0, acc[r]);
            acc[r] = fmaf(a.y, w1, acc[r]);
            acc[r] = fmaf(a.z, w2, acc[r]);
            acc[r] = fmaf(a.w, w3, acc[r]);
        }
    }
    if (src2) {
        for (int k = 0; k < K2; k += 4) {
            const float w0 = w2t[(k + 0) * N + col], w1 = w2t[(k + 1) * N + col], w2 = w2t[(k + 2) * N + col],
                        w3 = w2t[(k + 3) * N + col];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const float4 a = *reinterpret_cast<const float4 *>(src2 + (row0 + r) * ld2 + k);
                acc[r] = fmaf(a.x, w0, acc[r]);
                acc[r] = fmaf(a.y, w1, acc[r]);
                acc[r] = fmaf(a.z, w2, acc[r]);
                acc[r] = fmaf(a.w, w3, acc[r]);
            }
        }
    }
    const float b = bias ? bias[col] : 0.0f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        float v = acc[r] + b;
        if (res) v += res[(row0 + r) * ldr + col];
        if (relu) v = fmaxf(v, 0.0f);
        dst[(row0 + r) * ldd + col] = v;
    }
    (void)N_pad;
}

// Distributes the kTM x N outputs over the CTA: one column per thread, ROWS rows each.
__device__ __forceinline__ void run_dense(const float *src1, int ld1, int K1, const float *w1t, const float *src2, int ld2,
                                          int K2, const float *w2t, const float *bias, const float *res, int ldr, bool relu,
                                          float *dst, int ldd, int N) {
    const int N_pad = (N + 3) & ~3;
    int groups = kChainThreads / N;
    if (groups < 1) groups = 1;
    if (groups > kTM) groups = kTM;
    // largest power-of-two row count per thread that still covers the tile
    int rows = kTM;
    while (rows > 1 && (kTM / (rows / 2)) <= groups) rows >>= 1;
    const int used_groups = kTM / rows;
    const int t = threadIdx.x;
    const int col = t % N, grp = t / N;
    if (t < N * used_groups && grp < used_groups) {
        const int row0 = grp * rows;
        switch (rows) {
            case 32: dense_step<32>(src1, ld1, K1, w1t, src2, ld2, K2, w2t, bias, res, ldr, relu, dst, ldd, N, N_pad, col, row0); break;
            case 16: dense_step<16>(src1, ld1, K1, w1t, src2, ld2, K2, w2t, bias, res, ldr, relu, dst, ldd, N, N_pad, col, row0); break;
            case 8: dense_step<8>(src1, ld1, K1, w1t, src2, ld2, K2, w2t, bias, res, ldr, relu, dst, ldd, N, N_pad, col, row0); break;
            case 4: dense_step<4>(src1, ld1, K1, w1t, src2, ld2, K2, w2t, bias, res, ldr, relu, dst, ldd, N, N_pad, col, row0); break;
            case 2: dense_step<2>(src1, ld1, K1, w1t, src2, ld2, K2, w2t, bias, res, ldr, relu, dst, ldd, N, N_pad, col, row0); break;
            default: dense_step<1>(src1, ld1, K1, w1t, src2, ld2, K2, w2t, bias, res, ldr, relu, dst, ldd, N, N_pad, col, row0); break;
        }
    }
    // zero the K-padding columns [N, N_pad) so that the next layer's 4-wide k loop reads zeros
    if (N_pad > N)
        for (int i = t; i < kTM * (N_pad - N); i += kChainThreads) dst[(i / (N_pad - N)) * ldd + N + i % (N_pad - N)] = 0.0f;
}

__global__ void __launch_bounds__(kChainThreads) chain_f32_kernel(F32Chain chain, ChainArgs a) {
    extern __shared__ __align__(16) float smem[];
    float *buf[3] = {smem, smem + kTM * kW0, smem + kTM * kW0 + kTM * kW1};
    const int ld[3] = {kW0, kW1, kW1};
    const int t = threadIdx.x;
    const int64_t n_tiles = (a.M + kTM - 1) / kTM;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row_base = tile * kTM;
        const int valid = static_cast<int>(min(static_cast<int64_t>(kTM), a.M - row_base));
        // ---- prologue: build the stack input in buffer 0 ------------------------------------------------
        const int in_dim = chain.layer[0].din, in_pad = chain.layer[0].din_pad;
        if (a.prologue == kRowsFromGlobal) {
            for (int i = t; i < kTM * in_pad; i += kChainThreads) {
                const int r = i / in_pad, c = i % in_pad;
                float v = (r < valid && c < in_dim) ? a.x[(row_base + r) * a.x_dim + c] : 0.0f;
                buf[0][r * kW0 + c] = (v == v) ? v : 0.0f;   // NaN rows of invalid SHOT points count as zeros (eval.py:215)
            }
        } else {
            const int P = a.arity * (a.arity - 1) / 2;
            const bool shot = a.prologue == kShotTuple;
            const int geo = shot ? 4 * P : 3 * P;          // columns of geometry
            const int geo_off = shot ? 0 : 256;           // DINO tile layout: [pair 256 | coords 3P]
            if (t < kTM) {
                const int r = t;
                float *row = buf[0] + r * kW0;
                if (r < valid) {
                    int64_t pt[8];
                    for (int k = 0; k < a.arity; ++k) pt[k] = a.idx.at(row_base + r, k);
                    encode_tuple_geometry(a.pc, a.normal, pt, a.arity, shot, [&](int col, float v) { row[geo_off + col] = v; });
                } else {
                    for (int c = 0; c < geo; ++c) row[geo_off + c] = 0.0f;
                }
                for (int c = geo_off + geo + (shot ? a.arity * 64 : 0); c < in_pad; ++c) row[c] = 0.0f;
            }
            if (shot) {  // gather the encoded descriptors of the tuple's points: [geo | feats arity*64]
                const int fw = a.arity * 64;
                for (int i = t; i < kTM * fw; i += kChainThreads) {
                    const int r = i / fw, c = i % fw;
                    float v = 0.0f;
                    if (r < valid) v = a.feat[a.idx.at(row_base + r, c >> 6) * 64 + (c & 63)];
                    buf[0][r * kW0 + geo + c] = v;
                }
            } else {     // desc_pair_transform over the gathered, already transformed descriptors (train_dino.py:95-96)
                float acc[kTM];
#pragma unroll
                for (int r = 0; r < kTM; ++r) acc[r] = 0.0f;
                for (int c = 0; c < a.pair.chunks; ++c) {
                    __syncthreads();
                    for (int i = t; i < kTM * 256; i += kChainThreads) {
                        const int r = i >> 8, k = i & 255;
                        buf[1][r * kW1 + k] = r < valid ? a.feat[a.idx.at(row_base + r, c) * 256 + k] : 0.0f;
                    }
                    __syncthreads();
                    const float *wt = a.pair.wt + static_cast<int64_t>(c) * 256 * 256;
                    for (int k = 0; k < 256; k += 4) {
                        const float w0 = wt[(k + 0) * 256 + t], w1 = wt[(k + 1) * 256 + t], w2 = wt[(k + 2) * 256 + t],
                                    w3 = wt[(k + 3) * 256 + t];
#pragma unroll
                        for (int r = 0; r < kTM; ++r) {
                            const float4 x = *reinterpret_cast<const float4 *>(buf[1] + r * kW1 + k);
                            acc[r] = fmaf(x.x, w0, acc[r]);
                            acc[r] = fmaf(x.y, w1, acc[r]);
                            acc[r] = fmaf(x.z, w2, acc[r]);
                            acc[r] = fmaf(x.w, w3, acc[r]);
                        }
                    }
                }
                const float b = a.pair.b[t];
#pragma unroll
                for (int r = 0; r < kTM; ++r) buf[0][r * kW0 + t] = acc[r] + b;
            }
        }
        __syncthreads();
        // ---- the ResLayers: x in buf[xi], h in buf[hi], y in buf[yi] -------------------------------------
        int xi = 0, hi = 1, yi = 2;
        for (int l = 0; l < chain.n_layers; ++l) {
            const F32Layer &L = chain.layer[l];
            run_dense(buf[xi], ld[xi], L.din_pad, L.w1t, nullptr, 0, 0, nullptr, L.b1, nullptr, 0, true, buf[hi], ld[hi], L.dout);
            __syncthreads();
            if (L.has_fc0)
                run_dense(buf[hi], ld[hi], L.dout_pad, L.w2t, buf[xi], ld[xi], L.din_pad, L.w0t, L.b20, nullptr, 0, false, buf[yi],
                          ld[yi], L.dout);
            else
                run_dense(buf[hi], ld[hi], L.dout_pad, L.w2t, nullptr, 0, 0, nullptr, L.b20, buf[xi], ld[xi], false, buf[yi], ld[yi],
                          L.dout);
            __syncthreads();
            // rotate so that the next x is this y; buffer 0 is the only 368-wide one but later inputs are <= 256
            const int nx = yi;
            yi = xi;
            xi = nx;
        }
        // ---- epilogue -----------------------------------------------------------------------------------
        const int od = a.out_dim;
        for (int i = t; i < valid * od; i += kChainThreads) {
            const int r = i / od, c = i % od;
            a.out[(row_base + r) * od + c] = buf[xi][r * ld[xi] + c];
        }
        __syncthreads();
    }
}

// plain Linear over K = chunks*256 input columns streamed from global: desc_transform (train_dino.py:80,95)
__global__ void __launch_bounds__(kChainThreads) linear_f32_kernel(F32Linear lin, const float *__restrict__ x, int64_t M,
                                                                  float *__restrict__ out) {
    __shared__ __align__(16) float s_x[kTM * 256];
    const int t = threadIdx.x;
    const int64_t n_tiles = (M + kTM - 1) / kTM;
    const int K = lin.chunks * 256;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row_base = tile * kTM;
        const int valid = static_cast<int>(min(static_cast<int64_t>(kTM), M - row_base));
        float acc[kTM];
#pragma unroll
        for (int r = 0; r < kTM; ++r) acc[r] = 0.0f;
        for (int c = 0; c < lin.chunks; ++c) {
            __syncthreads();
            for (int i = t; i < kTM * 256; i += kChainThreads) {
                const int r = i >> 8, k = i & 255;
                const float v = r < valid ? x[(row_base + r) * K + c * 256 + k] : 0.0f;
                s_x[i] = (v == v) ? v : 0.0f;
            }
            __syncthreads();
            const float *wt = lin.wt + static_cast<int64_t>(c) * 256 * lin.dout;
            if (t < lin.dout)
                for (int k = 0; k < 256; k += 4) {
                    const float w0 = wt[(k + 0) * lin.dout + t], w1 = wt[(k + 1) * lin.dout + t], w2 = wt[(k + 2) * lin.dout + t],
                                w3 = wt[(k + 3) * lin.dout + t];
#pragma unroll
                    for (int r = 0; r < kTM; ++r) {
                        const float4 v = *reinterpret_cast<const float4 *>(s_x + r * 256 + k);
                        acc[r] = fmaf(v.x, w0, acc[r]);
                        acc[r] = fmaf(v.y, w1, acc[r]);
                        acc[r] = fmaf(v.z, w2, acc[r]);
                        acc[r] = fmaf(v.w, w3, acc[r]);
                    }
                }
        }
        if (t < lin.dout) {
            const float b = lin.b[t];
            for (int r = 0; r < valid; ++r) out[(row_base + r) * lin.dout + t] = acc[r] + b;
        }
    }
}

}  // namespace cppf

using namespace cppf;

// ---------------------------------------------------------------------------------------------------------
// model object
// ---------------------------------------------------------------------------------------------------------
struct cppf_heads {
    HeadsModel model;
    float *dev_f32;          // packed float32 weights (transposed, padded)
    size_t dev_f32_bytes;
    F32Chain shot_encoder, tuple_encoder, logit_encoder, scale_encoder;
    F32Linear desc_transform, desc_pair;
    void *tc;                // heads_tc.cu state (bf16 images), or null
    float *host_weights;     // the caller's flat array (kept for the tensor-core packer)
};

namespace {

int pad4(int v) { return (v + 3) & ~3; }

// Appends W^T ([din_pad][dout], zero padded) of a [dout][din] row-major weight; `perm` maps packed k -> source column.
size_t push_transposed(std::vector<float> &blob, const float *w, int dout, int din, const std::vector<int> *perm = nullptr) {
    const size_t off = blob.size();
    const int kp = pad4(din);
    blob.resize(off + static_cast<size_t>(kp) * dout, 0.0f);
    for (int k = 0; k < din; ++k) {
        const int src = perm ? (*perm)[k] : k;
        for (int n = 0; n < dout; ++n) blob[off + static_cast<size_t>(k) * dout + n] = w[static_cast<size_t>(n) * din + src];
    }
    // keep every array 16-byte aligned for the float4 loads of the activations (weights are scalar loads, but cheap)
    while (blob.size() % 4) blob.push_back(0.0f);
    return off;
}

size_t push_vector(std::vector<float> &blob, const float *a, const float *b, int n) {
    const size_t off = blob.size();
    for (int i = 0; i < n; ++i) blob.push_back(a[i] + (b ? b[i] : 0.0f));
    while (blob.size() % 4) blob.push_back(0.0f);
    return off;
}

struct ChainOffsets {
    size_t w1t[kMaxResLayers], b1[kMaxResLayers], w2t[kMaxResLayers], w0t[kMaxResLayers], b20[kMaxResLayers];
};

void pack_chain(std::vector<float> &blob, const float *w, const StackDesc &s, ChainOffsets &o, const std::vector<int> *perm0) {
    for (int l = 0; l < s.n_layers; ++l) {
        const ResLayerDesc &L = s.layer[l];
        const std::vector<int> *perm = l == 0 ? perm0 : nullptr;
        o.w1t[l] = push_transposed(blob, w + L.w1, L.dout, L.din, perm);
        o.b1[l] = push_vector(blob, w + L.b1, nullptr, L.dout);
        o.w2t[l] = push_transposed(blob, w + L.w2, L.dout, L.dout);
        o.w0t[l] = L.has_fc0 ? push_transposed(blob, w + L.w0, L.dout, L.din, perm) : 0;
        o.b20[l] = push_vector(blob, w + L.b2, L.has_fc0 ? w + L.b0 : nullptr, L.dout);
    }
}

void bind_chain(F32Chain &c, const StackDesc &s, const ChainOffsets &o, const float *dev) {
    c.n_layers = s.n_layers;
    for (int l = 0; l < s.n_layers; ++l) {
        const ResLayerDesc &L = s.layer[l];
        F32Layer &F = c.layer[l];
        F.din = L.din;
        F.dout = L.dout;
        F.din_pad = pad4(L.din);
        F.dout_pad = pad4(L.dout);
        F.has_fc0 = L.has_fc0;
        F.w1t = dev + o.w1t[l];
        F.b1 = dev + o.b1[l];
        F.w2t = dev + o.w2t[l];
        F.w0t = L.has_fc0 ? dev + o.w0t[l] : nullptr;
        F.b20 = dev + o.b20[l];
    }
}

}  // namespace

// heads_tc.cu
extern "C" int cppf_heads_tc_create(const HeadsModel *model, const float *weights_host, void **state);
extern "C" void cppf_heads_tc_destroy(void *state);
extern "C" int64_t cppf_heads_tc_workspace_bytes(const void *state, int64_t T, int64_t n);
extern "C" int cppf_heads_tc_forward(const void *state, const float *pc, int64_t n, const void *idx, int idx_is_i64,
                                     int64_t idx_stride, int64_t T, const float *feat, const float *normal, float *logits,
                                     float *scale, unsigned char *bins, const float *u01, unsigned long long seed, void *ws,
                                     int64_t ws_bytes, void *stream);

CPPF_API int cppf_heads_create(int branch, int num_more, const float *weights_host, int64_t n_floats, cppf_heads **out) {
    if ((branch != 0 && branch != 1) || num_more < 0 || num_more > 6 || !weights_host || !out) return CPPF_ERR_INVALID_ARGUMENT;
    HeadsModel m = describe_heads(branch, num_more);
    if (n_floats != m.n_floats) {
        fprintf(stderr, "[cppf_b200] cppf_heads_create: expected %lld floats for branch %d, got %lld\n",
                static_cast<long long>(m.n_floats), branch, static_cast<long long>(n_floats));
        return CPPF_ERR_INVALID_ARGUMENT;
    }
    if (m.tuple_encoder.in_dim() > kW0 || m.arity > 8) return CPPF_ERR_UNSUPPORTED;
    cppf_heads *h = new cppf_heads();
    h->model = m;
    h->tc = nullptr;
    h->host_weights = static_cast<float *>(malloc(sizeof(float) * n_floats));
    memcpy(h->host_weights, weights_host, sizeof(float) * n_floats);
    std::vector<float> blob;
    ChainOffsets o_shot{}, o_tuple{}, o_logit{}, o_scale{};
    size_t o_dt = 0, o_dtb = 0, o_dp = 0, o_dpb = 0;
    std::vector<int> perm;  // DINO tile layout [pair 256 | coords 3P]  <-  reference [coords 3P | pair 256]
    if (branch == 0) {
        pack_chain(blob, weights_host, m.shot_encoder, o_shot, nullptr);
    } else {
        const int geo = 3 * m.n_pairs;
        for (int k = 0; k < 256; ++k) perm.push_back(geo + k);
        for (int k = 0; k < geo; ++k) perm.push_back(k);
        o_dt = push_transposed(blob, weights_host + m.desc_transform.w, 256, 1024);
        o_dtb = push_vector(blob, weights_host + m.desc_transform.b, nullptr, 256);
        o_dp = push_transposed(blob, weights_host + m.desc_pair_transform.w, 256, 256 * m.arity);
        o_dpb = push_vector(blob, weights_host + m.desc_pair_transform.b, nullptr, 256);
    }
    pack_chain(blob, weights_host, m.tuple_encoder, o_tuple, branch == 1 ? &perm : nullptr);
    pack_chain(blob, weights_host, m.logit_encoder, o_logit, nullptr);
    pack_chain(blob, weights_host, m.scale_encoder, o_scale, nullptr);
    h->dev_f32_bytes = blob.size() * sizeof(float);
    if (cudaMalloc(&h->dev_f32, h->dev_f32_bytes) != cudaSuccess ||
        cudaMemcpy(h->dev_f32, blob.data(), h->dev_f32_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        fprintf(stderr, "[cppf_b200] cppf_heads_create: device allocation/copy failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        free(h->host_weights);
        delete h;
        return CPPF_ERR_CUDA;
    }
    if (branch == 0) bind_chain(h->shot_encoder, m.shot_encoder, o_shot, h->dev_f32);
    bind_chain(h->tuple_encoder, m.tuple_encoder, o_tuple, h->dev_f32);
    bind_chain(h->logit_encoder, m.logit_encoder, o_logit, h->dev_f32);
    bind_chain(h->scale_encoder, m.scale_encoder, o_scale, h->dev_f32);
    if (branch == 1) {
        h->desc_transform = {4, 256, h->dev_f32 + o_dt, h->dev_f32 + o_dtb};
        h->desc_pair = {m.arity, 256, h->dev_f32 + o_dp, h->dev_f32 + o_dpb};
    }
    const int rc = cppf_heads_tc_create(&h->model, h->host_weights, &h->tc);
    if (rc != CPPF_OK && rc != CPPF_ERR_UNSUPPORTED) {
        cudaFree(h->dev_f32);
        free(h->host_weights);
        delete h;
        return rc;
    }
    const size_t smem = sizeof(float) * kTM * (kW0 + 2 * kW1);
    CPPF_CUDA_TRY(cudaFuncSetAttribute(chain_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    *out = h;
    return CPPF_OK;
}

CPPF_API int cppf_heads_destroy(cppf_heads *h) {
    if (!h) return CPPF_OK;
    if (h->tc) cppf_heads_tc_destroy(h->tc);
    cudaFree(h->dev_f32);
    free(h->host_weights);
    delete h;
    return CPPF_OK;
}

CPPF_API int cppf_heads_has_tc(const cppf_heads *h) { return (h && h->tc) ? 1 : 0; }

// internal (frame.cu): the tensor-core state and branch of a model
extern "C" const void *cppf_heads_tc_state(const cppf_heads *h) { return h ? h->tc : nullptr; }
extern "C" int cppf_heads_branch(const cppf_heads *h) { return h ? h->model.branch : -1; }

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

CPPF_API int64_t cppf_heads_workspace_bytes(const cppf_heads *h, int64_t T, int64_t n, int precision) {
    if (!h || T < 0 || n < 0) return 0;
    if (precision == 1 && h->tc) return cppf_heads_tc_workspace_bytes(h->tc, T, n);
    const size_t per_point = h->model.branch == 0 ? 64 : 256;
    return static_cast<int64_t>(align256(sizeof(float) * per_point * static_cast<size_t>(n)) +
                                align256(sizeof(float) * 256 * static_cast<size_t>(T)) + 256);
}

CPPF_API int cppf_heads_forward_sampled(const cppf_heads *h, int precision, const float *pc, int64_t n, const void *idx,
                                        int idx_is_i64, int64_t idx_stride, int64_t T, const float *feat, const float *normal,
                                        const float *u01, uint64_t seed, uint8_t *bins, float *scale, void *ws, int64_t ws_bytes,
                                        void *stream) {
    if (!h || !pc || !feat || !bins || !scale || !ws || n <= 0 || T < 0) return CPPF_ERR_INVALID_ARGUMENT;
    if (T > 0 && !idx) return CPPF_ERR_INVALID_ARGUMENT;
    if (h->model.branch == 0 && !normal) return CPPF_ERR_INVALID_ARGUMENT;
    if (idx_stride < h->model.arity) return CPPF_ERR_INVALID_ARGUMENT;
    // the draw is an epilogue of the tensor-core kernel only; the float32 path keeps the two-call form
    if (precision != 1 || !h->tc) return CPPF_ERR_UNSUPPORTED;
    return cppf_heads_tc_forward(h->tc, pc, n, idx, idx_is_i64, idx_stride, T, feat, normal, nullptr, scale, bins, u01, seed, ws,
                                 ws_bytes, stream);
}

CPPF_API int cppf_heads_forward(const cppf_heads *h, int precision, const float *pc, int64_t n, const void *idx, int idx_is_i64,
                                int64_t idx_stride, int64_t T, const float *feat, const float *normal, float *logits,
                                float *scale, void *ws, int64_t ws_bytes, void *stream) {
    if (!h || !pc || !feat || !logits || !scale || !ws || n <= 0 || T < 0) return CPPF_ERR_INVALID_ARGUMENT;
    if (T > 0 && !idx) return CPPF_ERR_INVALID_ARGUMENT;
    if (h->model.branch == 0 && !normal) return CPPF_ERR_INVALID_ARGUMENT;
    if (idx_stride < h->model.arity) return CPPF_ERR_INVALID_ARGUMENT;
    if (precision == 1) {
        if (!h->tc) return CPPF_ERR_UNSUPPORTED;
        return cppf_heads_tc_forward(h->tc, pc, n, idx, idx_is_i64, idx_stride, T, feat, normal, logits, scale, nullptr, nullptr, 0,
                                     ws, ws_bytes, stream);
    }
    if (precision != 0) return CPPF_ERR_INVALID_ARGUMENT;
    if (ws_bytes < cppf_heads_workspace_bytes(h, T, n, 0)) return CPPF_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t smem = sizeof(float) * kTM * (kW0 + 2 * kW1);
    const int per_point = h->model.branch == 0 ? 64 : 256;
    float *point_feat = static_cast<float *>(ws);
    float *tuple_feat = reinterpret_cast<float *>(static_cast<unsigned char *>(ws) + align256(sizeof(float) * per_point * static_cast<size_t>(n)));
    const int sms = device_info().sm_count;
    auto tiles = [&](int64_t rows) { return static_cast<int>(std::min<int64_t>((rows + kTM - 1) / kTM, static_cast<int64_t>(sms) * 2)); };

    // per point: shot_encoder (train_shot.py:118) or the hoisted desc_transform (train_dino.py:95)
    if (h->model.branch == 0) {
        ChainArgs a{};
        a.prologue = kRowsFromGlobal;
        a.M = n;
        a.x = feat;
        a.x_dim = CPPF_SHOT_DIM;
        a.out = point_feat;
        a.out_dim = 64;
        chain_f32_kernel<<<tiles(n), kChainThreads, smem, s>>>(h->shot_encoder, a);
    } else {
        linear_f32_kernel<<<tiles(n), kChainThreads, 0, s>>>(h->desc_transform, feat, n, point_feat);
    }
    CPPF_LAUNCH_CHECK();
    if (T == 0) return CPPF_OK;
    // per tuple: gather-encode + tuple_encoder
    {
        ChainArgs a{};
        a.prologue = h->model.branch == 0 ? kShotTuple : kDinoTuple;
        a.M = T;
        a.pc = pc;
        a.normal = normal;
        a.feat = point_feat;
        a.idx = IdxView{idx, idx_stride, idx_is_i64};
        a.arity = h->model.arity;
        a.pair = h->desc_pair;
        a.out = tuple_feat;
        a.out_dim = 256;
        chain_f32_kernel<<<tiles(T), kChainThreads, smem, s>>>(h->tuple_encoder, a);
        CPPF_LAUNCH_CHECK();
    }
    for (int head = 0; head < 2; ++head) {
        ChainArgs a{};
        a.prologue = kRowsFromGlobal;
        a.M = T;
        a.x = tuple_feat;
        a.x_dim = 256;
        a.out = head == 0 ? scale : logits;
        a.out_dim = head == 0 ? 3 : 192;
        chain_f32_kernel<<<tiles(T), kChainThreads, smem, s>>>(head == 0 ? h->scale_encoder : h->logit_encoder, a);
        CPPF_LAUNCH_CHECK();
    }
    return CPPF_OK;
}
