// bf16 tensor-core path of the BeyondCPPF heads: tcgen05.mma with TMEM accumulators (sm_100a).
//
// One CTA carries a tile of 128 rows (tuples or points) through a whole *program* -- tuple encoding,
// tuple_encoder, logit_encoder and scale_encoder (train_shot.py:117-122, train_dino.py:128-133) -- with
//   * activations resident in shared memory as bf16, written by the epilogue directly in the UMMA
//     K-major no-swizzle layout (8x16-byte core matrices; plane kc = 8 columns x 128 rows = 2 KB), so
//     the output of one layer is the A operand of the next without any reshuffle;
//   * accumulators in TMEM (128 lanes x N<=256 fp32 columns), read back with tcgen05.ld 32x32b;
//   * weights pre-packed on the host as the exact shared-memory image of the B operand and streamed
//     from L2 in 64-column slabs with cp.async.bulk (the TMA engine) through a 2-stage mbarrier ring.
// Warp roles: warps 0-7 prologue/epilogue (thread <-> TMEM lane/row), warp 8 weight producer, warp 9 MMA
// issuer (one elected thread).  A ResLayer y = fc2(relu(fc1 x)) + (fc0 x | x) is two phases:
//   phase 1: D = X W1^T                      epilogue: H = relu(D + b1)                 (bf16 -> smem)
//   phase 2: D = H W2^T (+ X W0^T)           epilogue: X = D + b2 (+ b0) (+ X)          (bf16 -> smem, in place)
// so HBM sees only the program's inputs (points, normals, per-point features, tuple indices) and its
// outputs (logits, scales).
#include "heads_common.cuh"

#include <cuda_bf16.h>
#include <cstring>
#include <vector>

namespace cppf {
namespace tc {

constexpr int kRows = 128;                 // rows per tile == TMEM lanes == UMMA M
constexpr int kPlane = kRows * 16;         // bytes of one 8-column plane of an activation buffer
constexpr int kXCols = 368;                // widest stack input (SHOT tuple encoding 360 -> 368)
constexpr int kHCols = 256;
constexpr int kSlabCols = 64;              // K columns per weight slab
constexpr int kRing = 2;
constexpr int kSlabBytesMax = 256 * kSlabCols * 2;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kTmemCols = 256;
constexpr int kMaxPhases = 40;

constexpr int kSmemX = 0;
constexpr int kSmemH = kSmemX + (kXCols / 8) * kPlane;            //  94208
constexpr int kSmemRing = kSmemH + (kHCols / 8) * kPlane;         // 159744
constexpr int kSmemBar = kSmemRing + kRing * kSlabBytesMax;       // 225280
constexpr int kSmemTotal = kSmemBar + 128;

// ---- program description (built on the host, read by every role) ---------------------------------------
enum Action : int {
    kActLoadRows = 0,     // rows of a float32 global matrix -> bf16 buffer (optionally a 256-column chunk)
    kActEncodeShot = 1,   // SHOT tuple encoding: [coords 30 | normals 10 | feats 5x64 | pad] -> X
    kActGather = 2,       // gathered per-point bf16 rows (chunk c of the tuple) -> buffer
    kActHidden = 3,       // H = relu(D + b)
    kActOut = 4,          // X = D + b (+ X)
    kActFinal = 5,        // global float32 / bf16 output = D + b
    kActPairOut = 6,      // X[0:256] = D + b, then DINO coords -> X[256:288]
};

struct Part {             // one A operand x one weight matrix, accumulated into the phase's D
    int a_buf;            // 0 = X, 1 = H
    int k_cols;           // multiple of 16
    int n_slabs;
};

struct Phase {
    int action;           // what warps 0-7 do before the phase's MMAs
    int wait_done;        // the action first waits for the previous phase's MMAs
    int has_mma;
    int acc_prev;         // keep accumulating into the previous phase's D (K-chunked plain Linear)
    int n;                // UMMA N of this phase's accumulators (multiple of 16)
    int n_parts;
    Part part[2];
    // action parameters
    int dst_buf;          // kActLoadRows / kActGather destination
    int chunk;            // kActLoadRows: column offset / 256; kActGather: tuple slot
    int cols;             // kActLoadRows: valid columns to read; epilogues: real output columns
    int residual;         // kActOut: add the previous X
    int store_feat;       // kActOut: also spill X (bf16) to the feature scratch
    int reload_feat;      // kActFinal: afterwards reload X from the feature scratch
    int out_sel;          // kActFinal: 0 = out0 (float32), 1 = out1 (float32), 2 = point features (bf16)
    int out_ld;           // leading dimension of that output
    int64_t bias_off;     // floats, into the bias blob
};

struct Program {
    int n_phases;
    int gather_cols;      // per-point feature width gathered per tuple slot (64 SHOT, 256 DINO)
    Phase phase[kMaxPhases];
};

struct Args {
    int64_t rows;                       // tuples or points
    const float *x;                     // kActLoadRows source [rows][x_ld]
    int x_ld;
    const float *pc, *normal;           // tuple encoders
    const __nv_bfloat16 *point_feat;    // [n][gather_cols] bf16 (per-point program output)
    IdxView idx;
    int arity;
    const unsigned char *weights;       // slab stream of this program
    const float *bias;
    __nv_bfloat16 *feat_scratch;        // [rows][256] bf16
    float *out0, *out1;
    __nv_bfloat16 *out_bf16;
};

// ---- PTX wrappers ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8x16 B core matrices,
// LBO = byte distance between the two 8-column halves of a K=16 step, SBO = between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3fffu) | (static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: bf16 x bf16 -> f32, both operands K-major, M = 128
__device__ __forceinline__ uint32_t instr_desc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(kRows >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint4 pack8(const float v[8]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t *>(&a);
    o.y = *reinterpret_cast<uint32_t *>(&b);
    o.z = *reinterpret_cast<uint32_t *>(&c);
    o.w = *reinterpret_cast<uint32_t *>(&d);
    return o;
}
__device__ __forceinline__ void unpack8(const uint4 &p, float v[8]) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
// element (row, col) of an activation buffer: plane col/8, 16 bytes per row
__device__ __forceinline__ unsigned char *act_chunk(unsigned char *buf, int row, int col8) { return buf + col8 * kPlane + row * 16; }

__device__ __forceinline__ void put_elem(unsigned char *buf, int row, int col, float v) {
    reinterpret_cast<__nv_bfloat16 *>(act_chunk(buf, row, col >> 3))[col & 7] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(kThreads, 1) chain_tc_kernel(const Program *__restrict__ prog_g, Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ Program prog;
    __shared__ uint32_t s_tmem_base;
    unsigned char *bufX = smem + kSmemX, *bufH = smem + kSmemH;
    const uint32_t ring0 = smem_u32(smem + kSmemRing);
    const uint32_t bar0 = smem_u32(smem + kSmemBar);
    const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * kRing, bar_act = bar0 + 16 * kRing, bar_done = bar_act + 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < static_cast<int>(sizeof(Program) / 4); i += kThreads)
        reinterpret_cast<uint32_t *>(&prog)[i] = reinterpret_cast<const uint32_t *>(prog_g)[i];
    if (tid == 0) {
        for (int s = 0; s < kRing; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_act, kEpiWarps * 32);
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // TMEM: 256 columns x 128 lanes of fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;
    const int64_t n_tiles = (a.rows + kRows - 1) / kRows;

    if (warp == kEpiWarps) {
        // =============================== weight producer ===============================================
        if (lane == 0) {
            uint32_t slab_seq = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const unsigned char *src = a.weights;
                for (int p = 0; p < prog.n_phases; ++p) {
                    const Phase &ph = prog.phase[p];
                    if (!ph.has_mma) continue;
                    for (int q = 0; q < ph.n_parts; ++q) {
                        int k_left = ph.part[q].k_cols;
                        for (int s = 0; s < ph.part[q].n_slabs; ++s, ++slab_seq) {
                            const int cols = k_left < kSlabCols ? k_left : kSlabCols;
                            const uint32_t bytes = static_cast<uint32_t>(ph.n) * 2u * cols;
                            const uint32_t stage = slab_seq % kRing, round = slab_seq / kRing;
                            mbar_wait(bar_empty + 8 * stage, (round & 1u) ^ 1u);
                            mbar_expect_tx(bar_full + 8 * stage, bytes);
                            bulk_load(ring0 + stage * kSlabBytesMax, src, bytes, bar_full + 8 * stage);
                            src += bytes;
                            k_left -= cols;
                        }
                    }
                }
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // =============================== MMA issuer =====================================================
        if (lane == 0) {
            uint32_t slab_seq = 0, act_seq = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int p = 0; p < prog.n_phases; ++p) {
                    const Phase &ph = prog.phase[p];
                    if (!ph.has_mma) continue;
                    mbar_wait(bar_act, act_seq & 1u);   // the phase's A operand is in shared memory, D is free
                    ++act_seq;
                    tc_fence_after();
                    const uint32_t idesc = instr_desc(ph.n);
                    uint32_t accumulate = ph.acc_prev ? 1u : 0u;
                    for (int q = 0; q < ph.n_parts; ++q) {
                        const uint32_t a_base = smem_u32(ph.part[q].a_buf == 0 ? bufX : bufH);
                        int k_done = 0;
                        for (int s = 0; s < ph.part[q].n_slabs; ++s, ++slab_seq) {
                            const int cols = ph.part[q].k_cols - k_done < kSlabCols ? ph.part[q].k_cols - k_done : kSlabCols;
                            const uint32_t stage = slab_seq % kRing, round = slab_seq / kRing;
                            mbar_wait(bar_full + 8 * stage, round & 1u);
                            tc_fence_after();
                            const uint32_t b_base = ring0 + stage * kSlabBytesMax;
                            const uint32_t b_plane = static_cast<uint32_t>(ph.n) * 16u;
                            for (int k = 0; k < cols; k += 16) {
                                const uint64_t ad = smem_desc(a_base + ((k_done + k) >> 3) * kPlane, kPlane, 128);
                                const uint64_t bd = smem_desc(b_base + (k >> 3) * b_plane, b_plane, 128);
                                umma(tmem, ad, bd, idesc, accumulate);
                                accumulate = 1;
                            }
                            umma_commit(bar_empty + 8 * stage);   // slab consumed once these MMAs retire
                            k_done += cols;
                        }
                    }
                    umma_commit(bar_done);
                }
            }
        }
    } else {
        // =============================== prologue / epilogue warps ======================================
        const int row = (warp & 3) * 32 + lane;        // TMEM lane == tile row
        const int half = warp >> 2;                    // column half handled by this warpgroup
        const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        uint32_t done_seq = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t row_base = tile * kRows;
            const int64_t grow = row_base + row;
            const bool live = grow < a.rows;
            for (int p = 0; p < prog.n_phases; ++p) {
                const Phase &ph = prog.phase[p];
                if (ph.wait_done) {
                    mbar_wait(bar_done, done_seq & 1u);
                    ++done_seq;
                    tc_fence_after();
                }
                const Phase *prev = p > 0 ? &prog.phase[p - 1] : nullptr;
                const float *bias = prev ? a.bias + prev->bias_off : nullptr;
                const int n_prev = prev ? prev->n : 0;
                switch (ph.action) {
                    case kActLoadRows: {
                        unsigned char *dst = ph.dst_buf == 0 ? bufX : bufH;
                        const int width = ph.part[0].k_cols;      // columns the MMA will read (multiple of 16)
                        // thread <-> (row, 8-column chunk): 128 rows x width/8 chunks over 256 threads
                        for (int i = tid; i < kRows * (width >> 3); i += kEpiWarps * 32) {
                            const int r = i & (kRows - 1), c8 = i >> 7;
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int col = c8 * 8 + j;
                                float x = 0.0f;
                                if (row_base + r < a.rows && col < ph.cols) x = a.x[(row_base + r) * a.x_ld + ph.chunk * 256 + col];
                                v[j] = (x == x) ? x : 0.0f;   // NaN rows of invalid SHOT points count as zeros (eval.py:215)
                            }
                            *reinterpret_cast<uint4 *>(act_chunk(dst, r, c8)) = pack8(v);
                        }
                        break;
                    }
                    case kActEncodeShot: {
                        const int P = a.arity * (a.arity - 1) / 2;
                        if (tid < kRows) {
                            int64_t pt[8];
                            for (int k = 0; k < a.arity; ++k) pt[k] = live ? a.idx.at(grow, k) : 0;
                            if (live)
                                encode_tuple_geometry(a.pc, a.normal, pt, a.arity, true, [&](int col, float v) { put_elem(bufX, row, col, v); });
                            else
                                for (int c = 0; c < 4 * P; ++c) put_elem(bufX, row, c, 0.0f);
                            for (int c = 4 * P + a.arity * 64; c < ph.part[0].k_cols; ++c) put_elem(bufX, row, c, 0.0f);
                        }
                        // gathered per-point features: slot s occupies columns 4P + 64 s .. + 63 (8 chunks of 8)
                        for (int i = tid; i < kRows * a.arity * 8; i += kEpiWarps * 32) {
                            const int r = i & (kRows - 1), c = i >> 7, slot = c >> 3, sub = c & 7;
                            uint4 v = make_uint4(0, 0, 0, 0);
                            if (row_base + r < a.rows)
                                v = *reinterpret_cast<const uint4 *>(a.point_feat + a.idx.at(row_base + r, slot) * 64 + sub * 8);
                            // 4P = 40 columns of geometry: the features start at column 40 = chunk 5
                            *reinterpret_cast<uint4 *>(act_chunk(bufX, r, (4 * P) / 8 + c)) = v;
                        }
                        break;
                    }
                    case kActGather: {
                        unsigned char *dst = ph.dst_buf == 0 ? bufX : bufH;
                        for (int i = tid; i < kRows * 32; i += kEpiWarps * 32) {
                            const int r = i & (kRows - 1), c8 = i >> 7;
                            uint4 v = make_uint4(0, 0, 0, 0);
                            if (row_base + r < a.rows)
                                v = *reinterpret_cast<const uint4 *>(a.point_feat + a.idx.at(row_base + r, ph.chunk) * 256 + c8 * 8);
                            *reinterpret_cast<uint4 *>(act_chunk(dst, r, c8)) = v;
                        }
                        break;
                    }
                    case kActHidden:
                    case kActOut:
                    case kActPairOut:
                    case kActFinal: {
                        // this warpgroup's half of the previous phase's accumulator columns, 8 at a time
                        const int c_begin = half * (n_prev / 2), c_end = c_begin + n_prev / 2;
                        for (int c = c_begin; c < c_end; c += 8) {
                            float v[8];
                            tmem_ld8(t_lane + static_cast<uint32_t>(c), v);
                            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + c));
                            const float4 b1 = __ldg(reinterpret_cast<const float4 *>(bias + c + 4));
                            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                            if (ph.action == kActHidden) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
                                *reinterpret_cast<uint4 *>(act_chunk(bufH, row, c >> 3)) = pack8(v);
                            } else if (ph.action == kActOut || ph.action == kActPairOut) {
                                uint4 *slot = reinterpret_cast<uint4 *>(act_chunk(bufX, row, c >> 3));
                                if (ph.residual) {
                                    float x[8];
                                    unpack8(*slot, x);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[j] += x[j];
                                }
                                const uint4 packed = pack8(v);
                                *slot = packed;
                                if (ph.store_feat && live) __stcg(reinterpret_cast<uint4 *>(a.feat_scratch + grow * 256 + c), packed);
                            } else if (live) {
                                if (ph.out_sel == 2) {
                                    if (c < ph.cols) *reinterpret_cast<uint4 *>(a.out_bf16 + grow * ph.out_ld + c) = pack8(v);
                                } else {
                                    float *out = (ph.out_sel == 0 ? a.out0 : a.out1) + grow * ph.out_ld;
#pragma unroll
                                    for (int j = 0; j < 8; ++j)
                                        if (c + j < ph.cols) out[c + j] = v[j];
                                }
                            }
                        }
                        if (ph.action == kActPairOut && tid < kRows) {   // DINO coords after the 256 pair columns
                            int64_t pt[8];
                            for (int k = 0; k < a.arity; ++k) pt[k] = live ? a.idx.at(grow, k) : 0;
                            const int P = a.arity * (a.arity - 1) / 2;
                            if (live)
                                encode_tuple_geometry(a.pc, a.normal, pt, a.arity, false, [&](int col, float v) { put_elem(bufX, row, 256 + col, v); });
                            else
                                for (int c = 0; c < 3 * P; ++c) put_elem(bufX, row, 256 + c, 0.0f);
                            for (int c = 256 + 3 * P; c < ph.part[0].k_cols; ++c) put_elem(bufX, row, c, 0.0f);
                        }
                        if (ph.action == kActFinal && ph.reload_feat) {
                            // all epilogue threads must be done reading D / writing outputs before X is refilled: the
                            // refill only touches X, which no in-flight MMA reads (the phase waited for them)
                            for (int i = tid; i < kRows * 32; i += kEpiWarps * 32) {
                                const int r = i & (kRows - 1), c8 = i >> 7;
                                uint4 v = make_uint4(0, 0, 0, 0);
                                if (row_base + r < a.rows) v = __ldcg(reinterpret_cast<const uint4 *>(a.feat_scratch + (row_base + r) * 256 + c8 * 8));
                                *reinterpret_cast<uint4 *>(act_chunk(bufX, r, c8)) = v;
                            }
                        }
                        break;
                    }
                    default: break;
                }
                if (ph.has_mma) {
                    fence_async_smem();      // generic-proxy writes -> visible to the tensor core's async proxy
                    tc_fence_before();
                    mbar_arrive(bar_act);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
}

// ---------------------------------------------------------------------------------------------------------
// host side: programs + weight images
// ---------------------------------------------------------------------------------------------------------
struct Builder {
    Program prog{};
    std::vector<uint16_t> stream;     // bf16 bits of the slab stream
    std::vector<float> bias;
    const float *w;

    static uint16_t bf16_bits(float f) {
        uint32_t u;
        memcpy(&u, &f, 4);
        if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;               // NaN
        const uint32_t lsb = (u >> 16) & 1u;
        return static_cast<uint16_t>((u + 0x7fffu + lsb) >> 16);            // round to nearest even
    }

    // appends the UMMA image of W [n_real][k_real] (row-major, source column = perm ? perm[k] : k) padded to n_pad x k_pad
    void push_weight(const float *W, int n_real, int k_real, int n_pad, int k_pad, const std::vector<int> *perm = nullptr) {
        const size_t off = stream.size();
        stream.resize(off + static_cast<size_t>(n_pad) * k_pad, 0);
        for (int k = 0; k < k_real; ++k) {
            const int src = perm ? (*perm)[k] : k;
            for (int n = 0; n < n_real; ++n)
                stream[off + static_cast<size_t>(k >> 3) * n_pad * 8 + static_cast<size_t>(n) * 8 + (k & 7)] =
                    bf16_bits(W[static_cast<size_t>(n) * k_real + src]);
        }
    }
    // same for an explicit list of source columns of W [n_real][ld]
    void push_weight_cols(const float *W, int n_real, int ld, int n_pad, int k_pad, const std::vector<int> &cols) {
        const size_t off = stream.size();
        stream.resize(off + static_cast<size_t>(n_pad) * k_pad, 0);
        for (size_t k = 0; k < cols.size(); ++k)
            for (int n = 0; n < n_real; ++n)
                stream[off + (k >> 3) * static_cast<size_t>(n_pad) * 8 + static_cast<size_t>(n) * 8 + (k & 7)] =
                    bf16_bits(W[static_cast<size_t>(n) * ld + cols[k]]);
    }
    int64_t push_bias(const float *a, const float *b, int n_real, int n_pad) {
        const int64_t off = static_cast<int64_t>(bias.size());
        for (int i = 0; i < n_pad; ++i) bias.push_back(i < n_real ? a[i] + (b ? b[i] : 0.0f) : 0.0f);
        return off;
    }
    static int pad16(int v) { return (v + 15) & ~15; }
    static int slabs(int k) { return (k + kSlabCols - 1) / kSlabCols; }

    Phase &add(int action, int wait_done) {
        Phase &ph = prog.phase[prog.n_phases++];
        ph = Phase{};
        ph.action = action;
        ph.wait_done = wait_done;
        return ph;
    }

    // Appends the two MMA phases of a ResLayer to the phase whose action produced X (`first`), returns the
    // phase that must carry the layer's output epilogue.
    void res_layer(const ResLayerDesc &L, Phase *first, int k_in_pad, const std::vector<int> *perm) {
        const int n_pad = pad16(L.dout);
        // phase A (already created by the caller): D = X W1^T
        first->has_mma = 1;
        first->n = n_pad;
        first->n_parts = 1;
        first->part[0] = Part{0, k_in_pad, slabs(k_in_pad)};
        push_weight(w + L.w1, L.dout, L.din, n_pad, k_in_pad, perm);
        first->bias_off = push_bias(w + L.b1, nullptr, L.dout, n_pad);
        // phase B: H = relu(D + b1); D = H W2^T (+ X W0^T)
        Phase &hb = add(kActHidden, 1);
        hb.has_mma = 1;
        hb.n = n_pad;
        hb.n_parts = L.has_fc0 ? 2 : 1;
        hb.part[0] = Part{1, n_pad, slabs(n_pad)};
        push_weight(w + L.w2, L.dout, L.dout, n_pad, n_pad);
        if (L.has_fc0) {
            hb.part[1] = Part{0, k_in_pad, slabs(k_in_pad)};
            push_weight(w + L.w0, L.dout, L.din, n_pad, k_in_pad, perm);
        }
        hb.bias_off = push_bias(w + L.b2, L.has_fc0 ? w + L.b0 : nullptr, L.dout, n_pad);
    }

    // Runs a stack whose input X is produced by `*cur`'s action; on return `*cur` is the phase whose action
    // must write the stack's output (its fields action/residual/... are set by the caller).
    Phase *stack(const StackDesc &s, Phase *cur, int k_in_pad, const std::vector<int> *perm0) {
        for (int l = 0; l < s.n_layers; ++l) {
            const ResLayerDesc &L = s.layer[l];
            res_layer(L, cur, l == 0 ? k_in_pad : pad16(L.din), l == 0 ? perm0 : nullptr);
            Phase &out = add(kActOut, 1);
            out.residual = !L.has_fc0;
            out.cols = L.dout;
            cur = &out;
        }
        return cur;
    }
};

struct State {
    HeadsModel model;
    Program *d_point_prog, *d_tuple_prog;
    unsigned char *d_point_w, *d_tuple_w;
    float *d_point_b, *d_tuple_b;
    int point_cols;
};

static int upload(const Builder &b, Program **d_prog, unsigned char **d_w, float **d_b) {
    CPPF_CUDA_TRY(cudaMalloc(d_prog, sizeof(Program)));
    CPPF_CUDA_TRY(cudaMemcpy(*d_prog, &b.prog, sizeof(Program), cudaMemcpyHostToDevice));
    CPPF_CUDA_TRY(cudaMalloc(d_w, b.stream.size() * 2 + 16));
    CPPF_CUDA_TRY(cudaMemcpy(*d_w, b.stream.data(), b.stream.size() * 2, cudaMemcpyHostToDevice));
    CPPF_CUDA_TRY(cudaMalloc(d_b, b.bias.size() * 4 + 16));
    CPPF_CUDA_TRY(cudaMemcpy(*d_b, b.bias.data(), b.bias.size() * 4, cudaMemcpyHostToDevice));
    return CPPF_OK;
}

}  // namespace tc
}  // namespace cppf

using namespace cppf;
using namespace cppf::tc;

extern "C" int cppf_heads_tc_create(const HeadsModel *model, const float *w, void **state) {
    *state = nullptr;
    const HeadsModel &m = *model;
    if (m.arity != 5) return CPPF_ERR_UNSUPPORTED;   // the tile layouts below assume 10 pairs (4P = 40, 3P = 30)
    State *st = new State{};
    st->model = m;
    // ---- per-point program -------------------------------------------------------------------------------
    Builder pb;
    pb.w = w;
    if (m.branch == 0) {   // shot_encoder: [n,352] -> [n,64] bf16
        Phase *cur = &pb.add(kActLoadRows, 0);
        cur->dst_buf = 0;
        cur->cols = CPPF_SHOT_DIM;
        cur = pb.stack(m.shot_encoder, cur, Builder::pad16(CPPF_SHOT_DIM), nullptr);
        cur->action = kActFinal;
        cur->out_sel = 2;
        cur->out_ld = 64;
        cur->cols = 64;
        st->point_cols = 64;
    } else {               // desc_transform (train_dino.py:80,95), hoisted per point: Linear 1024 -> 256 as four K-chunks
        const LinearDesc &L = m.desc_transform;
        for (int c = 0; c < 4; ++c) {
            Phase &ph = pb.add(kActLoadRows, c > 0);     // X is refilled only after the previous chunk's MMAs retired
            ph.dst_buf = 0;
            ph.chunk = c;
            ph.cols = 256;
            ph.has_mma = 1;
            ph.acc_prev = c > 0;
            ph.n = 256;
            ph.n_parts = 1;
            ph.part[0] = Part{0, 256, Builder::slabs(256)};
            std::vector<int> cols(256);
            for (int k = 0; k < 256; ++k) cols[k] = c * 256 + k;
            pb.push_weight_cols(w + L.w, 256, L.din, 256, 256, cols);
            ph.bias_off = pb.push_bias(w + L.b, nullptr, 256, 256);
        }
        Phase &fin = pb.add(kActFinal, 1);
        fin.out_sel = 2;
        fin.out_ld = 256;
        fin.cols = 256;
        st->point_cols = 256;
    }
    // ---- per-tuple program -------------------------------------------------------------------------------
    Builder tb;
    tb.w = w;
    Phase *cur;
    if (m.branch == 0) {
        cur = &tb.add(kActEncodeShot, 0);
        cur = tb.stack(m.tuple_encoder, cur, Builder::pad16(m.tuple_encoder.in_dim()), nullptr);
    } else {
        // desc_pair_transform over the 5 gathered (already transformed) descriptors (train_dino.py:95-96), then
        // the tile is [pair 256 | coords 30 | pad 2]: tuple_encoder.0's input columns are permuted to match
        const LinearDesc &L = m.desc_pair_transform;
        for (int c = 0; c < m.arity; ++c) {
            Phase &ph = tb.add(kActGather, c > 0);
            ph.dst_buf = 0;
            ph.chunk = c;
            ph.has_mma = 1;
            ph.acc_prev = c > 0;
            ph.n = 256;
            ph.n_parts = 1;
            ph.part[0] = Part{0, 256, Builder::slabs(256)};
            std::vector<int> cols(256);
            for (int k = 0; k < 256; ++k) cols[k] = c * 256 + k;
            tb.push_weight_cols(w + L.w, 256, L.din, 256, 256, cols);
            ph.bias_off = tb.push_bias(w + L.b, nullptr, 256, 256);
        }
        cur = &tb.add(kActPairOut, 1);
        const int geo = 3 * m.n_pairs;
        std::vector<int> perm;
        for (int k = 0; k < 256; ++k) perm.push_back(geo + k);
        for (int k = 0; k < geo; ++k) perm.push_back(k);
        cur = tb.stack(m.tuple_encoder, cur, Builder::pad16(256 + geo), &perm);
    }
    cur->store_feat = 1;                         // feat = tuple_encoder output, needed by both heads
    cur = tb.stack(m.logit_encoder, cur, 256, nullptr);
    cur->action = kActFinal;                     // logits [T,192] float32
    cur->out_sel = 0;
    cur->out_ld = 192;
    cur->cols = 192;
    cur->reload_feat = 1;
    cur = tb.stack(m.scale_encoder, cur, 256, nullptr);
    cur->action = kActFinal;                     // scale [T,3] float32
    cur->out_sel = 1;
    cur->out_ld = 3;
    cur->cols = 3;
    if (pb.prog.n_phases > kMaxPhases || tb.prog.n_phases > kMaxPhases) {
        delete st;
        return CPPF_ERR_UNSUPPORTED;
    }
    pb.prog.gather_cols = tb.prog.gather_cols = st->point_cols;
    int rc = upload(pb, &st->d_point_prog, &st->d_point_w, &st->d_point_b);
    if (rc == CPPF_OK) rc = upload(tb, &st->d_tuple_prog, &st->d_tuple_w, &st->d_tuple_b);
    if (rc != CPPF_OK) {
        delete st;
        return rc;
    }
    CPPF_CUDA_TRY(cudaFuncSetAttribute(chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    *state = st;
    return CPPF_OK;
}

extern "C" void cppf_heads_tc_destroy(void *state) {
    State *st = static_cast<State *>(state);
    if (!st) return;
    cudaFree(st->d_point_prog);
    cudaFree(st->d_tuple_prog);
    cudaFree(st->d_point_w);
    cudaFree(st->d_tuple_w);
    cudaFree(st->d_point_b);
    cudaFree(st->d_tuple_b);
    delete st;
}

static size_t tc_align(size_t x) { return (x + 255) / 256 * 256; }

extern "C" int64_t cppf_heads_tc_workspace_bytes(const void *state, int64_t T, int64_t n) {
    const State *st = static_cast<const State *>(state);
    if (!st) return 0;
    return static_cast<int64_t>(tc_align(2 * static_cast<size_t>(st->point_cols) * n) + tc_align(2 * 256 * static_cast<size_t>(T)) + 256);
}

extern "C" int cppf_heads_tc_forward(const void *state, const float *pc, int64_t n, const void *idx, int idx_is_i64,
                                     int64_t idx_stride, int64_t T, const float *feat, const float *normal, float *logits,
                                     float *scale, void *ws, int64_t ws_bytes, void *stream) {
    const State *st = static_cast<const State *>(state);
    if (!st) return CPPF_ERR_UNSUPPORTED;
    if (ws_bytes < cppf_heads_tc_workspace_bytes(state, T, n)) return CPPF_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    __nv_bfloat16 *point_feat = static_cast<__nv_bfloat16 *>(ws);
    __nv_bfloat16 *feat_scratch = reinterpret_cast<__nv_bfloat16 *>(static_cast<unsigned char *>(ws) + tc_align(2 * static_cast<size_t>(st->point_cols) * n));
    const int sms = device_info().sm_count;
    {
        Args a{};
        a.rows = n;
        a.x = feat;
        a.x_ld = st->model.branch == 0 ? CPPF_SHOT_DIM : 1024;
        a.arity = st->model.arity;
        a.weights = st->d_point_w;
        a.bias = st->d_point_b;
        a.out_bf16 = point_feat;
        const int blocks = static_cast<int>(std::min<int64_t>((n + kRows - 1) / kRows, sms));
        chain_tc_kernel<<<blocks, kThreads, kSmemTotal, s>>>(st->d_point_prog, a);
        CPPF_LAUNCH_CHECK();
    }
    if (T == 0) return CPPF_OK;
    {
        Args a{};
        a.rows = T;
        a.pc = pc;
        a.normal = normal;
        a.point_feat = point_feat;
        a.idx = IdxView{idx, idx_stride, idx_is_i64};
        a.arity = st->model.arity;
        a.weights = st->d_tuple_w;
        a.bias = st->d_tuple_b;
        a.feat_scratch = feat_scratch;
        a.out0 = logits;
        a.out1 = scale;
        const int blocks = static_cast<int>(std::min<int64_t>((T + kRows - 1) / kRows, sms));
        chain_tc_kernel<<<blocks, kThreads, kSmemTotal, s>>>(st->d_tuple_prog, a);
        CPPF_LAUNCH_CHECK();
    }
    return CPPF_OK;
}
