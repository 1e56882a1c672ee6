// bf16 tensor-core path of the BeyondCPPF heads: tcgen05.mma with TMEM accumulators (sm_100a), version 2.
//
// One CTA carries TWO independent 128-row tiles ("slots") through a *program* -- tuple encoding, tuple_encoder,
// logit_encoder and scale_encoder (train_shot.py:117-122, train_dino.py:128-133) -- so that the tensor pipe
// works on one slot while the epilogue warps of the other turn its accumulators into the next layer's operand:
//   * layer inputs X stay in shared memory as bf16 in the UMMA K-major no-swizzle layout (8x16-byte core
//     matrices; plane kc = 8 columns x 128 rows = 2 KB), written by the epilogue in place;
//   * accumulators D live in TMEM (128 lanes x <=256 fp32 columns per slot) and are read with tcgen05.ld;
//   * the hidden activation H = relu(fc1 x + b1) never touches shared memory: the epilogue packs it to bf16 and
//     writes it back to TMEM (tcgen05.st), where it is the A operand of the second GEMM (tcgen05.mma with A in
//     TMEM), so a ResLayer y = fc2(H) + (fc0 x | x) costs one shared-memory activation buffer per slot;
//   * every accumulator is at most 128 columns wide (256-wide layers are computed as two N-chunks), which is
//     what lets two slots share the 512 TMEM columns; inputs wider than 256 columns (the 360/352/286-wide
//     first layers) are fed as two K-chunks through the same X buffer;
//   * weights are pre-packed on the host as the exact shared-memory image of the B operand and streamed from L2
//     in 16 KB slabs with cp.async.bulk (the TMA engine) through a 5-stage mbarrier ring.
// Warp roles: warps 0-7 epilogue of slot 0, warps 8-15 epilogue of slot 1 (thread <-> TMEM lane/row, two
// threads per row splitting the columns), warp 16 weight producer, warp 17 MMA issuer (one elected thread).
// The issue order is static (phase-major, slot-minor), so producer and issuer agree without communication.
// HBM sees only the program's inputs (points, normals, per-point features, tuple indices) and its outputs.
#include "heads_common.cuh"
#include "frame.cuh"

#include <cuda_bf16.h>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef CPPF_TC_EXP
#define CPPF_TC_EXP 0      // timing experiments only (tools/heads_profile.py; results are garbage): bit 2 no Final stores, 3 no
#endif                     // TMEM loads, 5 no MMA issue, 6 slot 1 ignores the weight ring (no wait, no release), 7 no
                           // fence.proxy.async, 8 nobody waits for weights and the producer does not run; 9 / 10 (valid
                           // results) the epilogue's / the issuers' critical waits spin on test_wait instead of try_wait

namespace cppf {
namespace tc {

constexpr int kRows = 128;                 // rows per tile == TMEM lanes == UMMA M
constexpr int kPlane = kRows * 16;         // bytes of one 8-column plane of an activation buffer
constexpr int kXCols = 256;
constexpr int kXBytes = (kXCols / 8) * kPlane;        // 65536
constexpr int kSlots = 2;
constexpr int kStages = 5;
constexpr int kSlabBytes = 16384;          // N<=128: 64 K-columns, N=256: 32 K-columns
constexpr int kSlotWarps = 8;
constexpr int kEpiWarps = kSlots * kSlotWarps;
constexpr int kThreads = (kEpiWarps + 1 + kSlots) * 32;   // 608: epilogue warps, weight producer, one MMA issuer per slot
constexpr int kTmemCols = 512;
constexpr int kSlotTmem = 256;             // TMEM columns per slot: D at +0 (and +128), H at +128
constexpr int kHTmem = 128;
constexpr int kMaxPhases = 48;

constexpr int kSmemX = 0;
constexpr int kSmemRing = kSmemX + kSlots * kXBytes;              // 131072
constexpr int kIdxStride = 5;              // staged tuple indices per row (the tile layouts assume 5-point tuples)
constexpr int kSmemIdx = kSmemRing + kStages * kSlabBytes;        // 212992: int32 [kSlots][128][kIdxStride]
constexpr int kSmemOnes = kSmemIdx + kSlots * kRows * kIdxStride * 4;   // 128 x 16 bf16 operand: columns 0,1 = 1, rest 0
constexpr int kSmemBar = kSmemOnes + 2 * kPlane;
constexpr int kSmemTotal = kSmemBar + 256;

// ---- program description (built on the host, read by every role) ---------------------------------------
enum Action : int {
    kActLoadRows = 0,     // float32 rows of a global matrix (columns src_col .. +cols) -> X[0 : width)
    kActEncodeShotA = 1,  // SHOT tuple encoding, chunk A: gathered features of tuple slots 0..3 -> X[0:256)
    kActEncodeShotB = 2,  // chunk B: [features of slot 4 (64) | coords 30 | normals 10 | pad 8] -> X[0:112)
    kActGather = 3,       // gathered 256-wide per-point bf16 rows of tuple slot `src_col` -> X[0:256)
    kActCoordsB = 4,      // DINO chunk B: [coords 30 | pad 2] -> X[0:32)
    kActHiddenT = 5,      // H[h_col : +n) = relu(D + b)      (bf16 -> TMEM)
    kActHiddenS = 6,      // X[dst_col : +n) = relu(D + b)    (bf16 -> shared memory; first layers)
    kActOut = 7,          // X[dst_col : +n) = D + b (+ X[res_col : +n))
    kActFinal = 8,        // global output = D + b
    kActOutT = 9,         // Y[dst_col : +n) = D + b          (bf16 -> TMEM, no relu: a layer output that stays in tensor memory)
    kActGatherSum = 10,   // X[0:256) = sum over the tuple's points k of the per-point row block G[idx_k][256 k : 256 k + 256)
};

struct Part {             // one A operand x one weight matrix, accumulated into D[d_col : d_col + n)   (host only)
    int a_tmem;           // 0: A = X in shared memory, 1: A = H in TMEM, 2: A = the constant ones operand (bias part:
                          //    B = [hi(bias) | lo(bias) | 0 ...], so that D += bias to ~16 mantissa bits)
    int a_col;            // first column of the operand (bf16 elements)
    int k_cols;           // multiple of 16
    int d_col;            // accumulator column offset inside the slot
    int n;                // UMMA N (16, 64, 128 or 256)
    int init;             // 1: the first MMA overwrites D, 0: accumulates onto it
    int64_t w_off;        // bytes, into the program's slab stream
};

// One slab of the weight stream = one TMA copy = up to four MMAs per slot.  Built on the host so that the single
// issuing warp does nothing per slab but test two barriers and fire the MMAs.
struct Slab {             // 16 bytes: the issuer reads a slab with ONE 128-bit uniform constant load
    uint32_t a_lo;        // A in shared memory: low descriptor word relative to the slot's X buffer; in TMEM: column offset
    uint32_t idesc;
    uint32_t nd;          // (n << 16) | d_col.  n = UMMA N: B-descriptor LBO = 16 n bytes (the field value is n), K-step = 32 n bytes
    uint32_t misc;        // bytes16 | n_mma << 16 | flags << 24   (slab size / 16, 1..4 MMAs)
    __host__ __device__ uint32_t bytes() const { return (misc & 0xffffu) << 4; }
};
enum : uint8_t {
    kSlabATmem = 1,       // A operand in TMEM
    kSlabFirst = 2,       // first slab of a phase: wait for the slot's operand
    kSlabLast = 4,        // last slab of a phase: signal the slot's epilogue
    kSlabAcc = 8,         // the first MMA accumulates (else it overwrites D)
    kSlabOnes = 16,       // A = the constant ones operand (bias rows as B)
};
constexpr int kMaxSlabs = 256;

struct Phase {
    int action;           // what the slot's epilogue warps do before the phase's MMAs
    int wait_done;        // the action first waits for the slot's previous MMAs
    int n_parts;          // 0: no MMA follows the action
    // action parameters
    int src_col;          // LoadRows: first source column; Gather: tuple slot
    int cols;             // LoadRows: valid source columns; Final: valid output columns of this chunk
    int width;            // LoadRows: columns written (zero padded), multiple of 32
    int d_col, n;         // accumulator chunk the epilogue reads
    int dst_col;          // X / H / global column the chunk is written to
    int residual, res_col;     // residual: 0 none, 1 = X[res_col : +n) in shared memory, 2 = bf16 in tensor memory (H/Y column res_col)
    int out_sel;          // Final: 0 = out0 (float32), 1 = out1 (float32), 2 = point features (bf16)
    int out_ld;
};

struct Program {
    int n_phases;
    int n_slabs;
    int gather_cols;
    Phase phase[kMaxPhases];
    Slab slab[kMaxSlabs];
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
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {   // non-blocking
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool elect_one() {   // one lane of the (converged) warp
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok));
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin)
        if (spin > (1u << 22)) __trap();      // >> any legitimate wait (a whole launch is < 1 ms): a deadlock surfaces within seconds
}
// Spinning on the non-blocking test instead of try_wait (which may suspend the warp for a HW-defined time before it looks
// again): for the waits that sit on the MMA <-> epilogue critical path of a slot.
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_test(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 16-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
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
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand in TMEM: element (row m, k) at lane m, 32-bit column a_tmem + k/2 (tools/probes/probe_ts_mma.cu)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 8 consecutive 32-bit columns (16 packed bf16) of this thread's lane; no wait (tmem_ld_wait)
__device__ __forceinline__ void tmem_ld8_raw(uint32_t taddr, uint32_t r[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t r[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t r[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
// relu fused into the conversion (low half <- a, high half <- b)
__device__ __forceinline__ uint32_t pack2_relu(float a, float b) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
__device__ __forceinline__ uint4 pack8(const float v[8]) {
    return make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
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
// 16-byte chunk (row, 8-column group col8) of an activation buffer
__device__ __forceinline__ unsigned char *act_chunk(unsigned char *buf, int row, int col8) { return buf + col8 * kPlane + row * 16; }
__device__ __forceinline__ void put_elem(unsigned char *buf, int row, int col, float v) {
    reinterpret_cast<__nv_bfloat16 *>(act_chunk(buf, row, col >> 3))[col & 7] = __float2bfloat16_rn(v);
}

// Geometry columns of the 5-point tuple encoding (prepare_tuple_inputs, train_shot.py:75-83 / train_dino.py:91-97) for
// one row, written as whole 16-byte chunks: 30 pair differences (+ 10 |n_i . n_j| for the SHOT branch), same float32
// operations as encode_tuple_geometry.  All 15 (30) point loads are issued before the first use; two threads share a
// row (`part` selects which chunks a thread stores).
__device__ __forceinline__ void tuple_geometry_chunks(const float *__restrict__ pc, const float *__restrict__ normal,
                                                      const int *__restrict__ idx_row, bool live, bool with_normals, int part,
                                                      unsigned char *X, int row, int first_chunk) {
    float pt[5][3], nr[5][3];
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) pt[k][c] = live ? __ldg(pc + 3 * static_cast<int64_t>(idx_row[k]) + c) : 0.0f;
    const bool need_normals = with_normals && part == 1;
    if (need_normals) {
#pragma unroll
        for (int k = 0; k < 5; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) {   // NaN normals of points with < 3 neighbours count as zeros (eval.py:216)
                const float v = live ? __ldg(normal + 3 * static_cast<int64_t>(idx_row[k]) + c) : 0.0f;
                nr[k][c] = (v == v) ? v : 0.0f;
            }
    }
    float v[48];
    int col = 0, pair = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = i + 1; j < 5; ++j) {
#pragma unroll
            for (int c = 0; c < 3; ++c) v[col++] = __fsub_rn(pt[i][c], pt[j][c]);
            if (need_normals) {
                const float *a = nr[i], *b = nr[j];
                const float d = __fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2]));
                const float e = __fadd_rn(__fadd_rn(__fmul_rn(-a[0], b[0]), __fmul_rn(-a[1], b[1])), __fmul_rn(-a[2], b[2]));
                v[30 + pair] = fmaxf(d, e);
            }
            ++pair;
        }
#pragma unroll
    for (int c = with_normals ? 40 : 30; c < 48; ++c) v[c] = 0.0f;
    // constant register indices only (a `part`-dependent index would send v[] to local memory)
    if (with_normals) {       // 40 columns + 8 zeros = 6 chunks: part 0 stores chunks 0..2 (coords 0..23), part 1 chunks 3..5
        if (part == 0) {
#pragma unroll
            for (int j = 0; j < 3; ++j) *reinterpret_cast<uint4 *>(act_chunk(X, row, first_chunk + j)) = pack8(v + 8 * j);
        } else {
#pragma unroll
            for (int j = 3; j < 6; ++j) *reinterpret_cast<uint4 *>(act_chunk(X, row, first_chunk + j)) = pack8(v + 8 * j);
        }
    } else {                  // 30 columns + 2 zeros = 4 chunks: two each
        if (part == 0) {
#pragma unroll
            for (int j = 0; j < 2; ++j) *reinterpret_cast<uint4 *>(act_chunk(X, row, first_chunk + j)) = pack8(v + 8 * j);
        } else {
#pragma unroll
            for (int j = 2; j < 4; ++j) *reinterpret_cast<uint4 *>(act_chunk(X, row, first_chunk + j)) = pack8(v + 8 * j);
        }
    }
}

// Work assignment.  The unit is a PAIR of 128-row tiles of one job (both slots of a CTA share every weight slab of the
// round, so they must run the same job's weights); round r gives CTA b the pair r * gridDim.x + b of the concatenation of
// all jobs' pairs.  A launch carries one job (`one`, kernel-parameter space) or the jobs of a frame (`multi`, device
// memory: same program, different rows / weights / outputs).  Every role evaluates this identically.
//
// The CTAs share the pairs in full rounds: with P pairs and G CTAs launched, ceil(P / G) rounds are unavoidable, so only
// G' = ceil(P / rounds) CTAs take part (391 tiles = 196 pairs on 148 SMs: 98 CTAs x 2 rounds, not 148 + 48; a frame's
// 1176 pairs: 147 CTAs x 8 rounds) and the rest leave their SMs to whatever else is queued.  G' is computed here, on the
// device, from the table: the host may launch for a capacity (CUDA-graph replay) without unbalancing the last round.
__device__ __forceinline__ int cta_share(const MultiArgs *__restrict__ multi, const Args &one, bool &single) {
    const int n_jobs = multi ? multi->n_jobs : 1;
    int64_t pairs = 0, tiles = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const int64_t rows = multi ? multi->job[j].rows : one.rows;
        const int64_t t = (rows + kRows - 1) / kRows;
        tiles += t;
        pairs += (t + 1) >> 1;
    }
    const int64_t g = gridDim.x;
    // few tiles (the per-point programs: ~20 tiles per cloud): one tile per CTA on as many SMs as there are tiles finishes
    // in one tile time; pairing them would halve the SMs in use and serialise two tiles per CTA
    single = tiles <= g;
    const int64_t total = single ? tiles : pairs;
    if (total <= 0) return 1;
    const int64_t rounds = (total + g - 1) / g;
    return static_cast<int>((total + rounds - 1) / rounds);
}

__device__ __forceinline__ bool pair_of(const MultiArgs *__restrict__ multi, const Args &one, bool single, int64_t q, int &job,
                                        int64_t &tile0, int64_t &n_tiles) {
    const int n_jobs = multi ? multi->n_jobs : 1;
    for (int j = 0; j < n_jobs; ++j) {
        const int64_t rows = multi ? multi->job[j].rows : one.rows;
        const int64_t tiles = (rows + kRows - 1) / kRows;
        const int64_t units = single ? tiles : (tiles + 1) >> 1;
        if (q < units) {
            job = j;
            tile0 = single ? q : 2 * q;
            n_tiles = single ? q + 1 : tiles;      // single-tile units: the second slot never has a tile
            return true;
        }
        q -= units;
    }
    return false;
}

template <int NB>
__device__ __forceinline__ void epilogue_batch(const Phase &ph, const Args &a, unsigned char *X, uint32_t t_slot_lane, int row,
                                               int64_t grow, bool live, int c /* column inside the chunk */) {
    float v[NB];     // the bias is already in the accumulator (ones-operand MMA)
    if (CPPF_TC_EXP & 8) {
#pragma unroll
        for (int j = 0; j < NB; ++j) v[j] = static_cast<float>(row + j);
    } else if (NB == 32) tmem_ld32(t_slot_lane + static_cast<uint32_t>(ph.d_col + c), v);
    else tmem_ld8(t_slot_lane + static_cast<uint32_t>(ph.d_col + c), v);
    if (NB == 32 && ph.residual == 2) {
        // the layer input lives in tensor memory (bf16 pairs, H/Y column space): identity residual read back from there
#pragma unroll
        for (int h = 0; h < 2; ++h) {      // two loads of 8 columns: 8 fewer live registers than one of 16 (no spills)
            uint32_t y[8];
            tmem_ld8_raw(t_slot_lane + kHTmem + static_cast<uint32_t>((ph.res_col + c) >> 1) + 8u * h, y);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[16 * h + 2 * j] += __uint_as_float(y[j] << 16);
                v[16 * h + 2 * j + 1] += __uint_as_float(y[j] & 0xffff0000u);
            }
        }
    }
    if (ph.action == kActHiddenT || ph.action == kActOutT) {
        uint32_t p[NB / 2];
        if (ph.action == kActHiddenT) {
#pragma unroll
            for (int j = 0; j < NB / 2; ++j) p[j] = pack2_relu(v[2 * j], v[2 * j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < NB / 2; ++j) p[j] = pack2(v[2 * j], v[2 * j + 1]);
        }
        const uint32_t h_addr = t_slot_lane + kHTmem + static_cast<uint32_t>((ph.dst_col + c) >> 1);
        if (NB == 32) tmem_st16(h_addr, p);
        else tmem_st4(h_addr, p);
    } else if (ph.action == kActHiddenS) {
#pragma unroll
        for (int j = 0; j < NB / 8; ++j)
            *reinterpret_cast<uint4 *>(act_chunk(X, row, (ph.dst_col + c) / 8 + j)) =
                make_uint4(pack2_relu(v[8 * j], v[8 * j + 1]), pack2_relu(v[8 * j + 2], v[8 * j + 3]), pack2_relu(v[8 * j + 4], v[8 * j + 5]),
                           pack2_relu(v[8 * j + 6], v[8 * j + 7]));
    } else if (ph.action == kActOut) {
        if (ph.residual == 1) {
#pragma unroll
            for (int j = 0; j < NB / 8; ++j) {
                float x[8];
                unpack8(*reinterpret_cast<const uint4 *>(act_chunk(X, row, (ph.res_col + c) / 8 + j)), x);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[8 * j + k] += x[k];
            }
        }
#pragma unroll
        for (int j = 0; j < NB / 8; ++j) {
            const uint4 packed = pack8(v + 8 * j);
            *reinterpret_cast<uint4 *>(act_chunk(X, row, (ph.dst_col + c) / 8 + j)) = packed;
        }
    } else if (NB == 32 && ph.out_sel == 0 && a.bins != nullptr) {
        // kActFinal of the logits with the decode fused in (eval.py:225-229): the 32 columns of this batch are the 32
        // bins of coordinate (dst_col + c) / 32 of this row; softmax numerators, running sum, one inverse-CDF draw
        if (live) {
            float m = v[0];
#pragma unroll
            for (int j = 1; j < NB; ++j) m = fmaxf(m, v[j]);
            // exp(v - m) = 2^(v log2e - m log2e): one FFMA (the product is exact inside it; the rounding of m log2e is common to
            // the 32 bins and cancels in the CDF) and one MUFU.EX2 per bin instead of expf's nine instructions
            constexpr float kLog2e = 1.4426950408889634f;
            const float shift = -m * kLog2e;
            float run = 0.0f;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                run += ex2_approx(fmaf(v[j], kLog2e, shift));
                v[j] = run;
            }
            const int64_t r6 = grow * 6 + ((ph.dst_col + c) >> 5);
            const float u = a.u01 ? __ldg(a.u01 + r6) : uniform_from_counter(a.seed, static_cast<uint64_t>(r6));
            const float cut = u * run;
            int b = 0;
#pragma unroll
            for (int j = 0; j < NB; ++j) b += v[j] <= cut ? 1 : 0;
            a.bins[r6] = static_cast<unsigned char>(b < 32 ? b : 31);
        }
    } else if (live && !((CPPF_TC_EXP & 4) && ph.out_sel != 2)) {   // kActFinal
        if (ph.out_sel == 2) {
#pragma unroll
            for (int j = 0; j < NB / 8; ++j)
                if (c + 8 * j < ph.cols)
                    *reinterpret_cast<uint4 *>(a.out_bf16 + grow * ph.out_ld + ph.dst_col + c + 8 * j) = pack8(v + 8 * j);
        } else {
            float *out = (ph.out_sel == 0 ? a.out0 : a.out1) + grow * ph.out_ld + ph.dst_col + c;
            if ((ph.out_ld & 3) == 0 && c + NB <= ph.cols) {
#pragma unroll
                for (int j = 0; j < NB / 4; ++j) reinterpret_cast<float4 *>(out)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if (c + j < ph.cols) out[j] = v[j];
            }
        }
    }
}

// The program lives in kernel-parameter (constant) space: every field the MMA issuer and the producer read is a
// uniform-datapath load indexed by warp-uniform counters, which is what lets tcgen05.mma / cp.async.bulk take their
// descriptors from uniform registers without per-lane "waterfall" loops (tools/probes/probe_mma_rate.cu: 64 cycles
// per M128 N128 K16 MMA when issued that way, 160-400 when issued from lane-divergent code).
// clock64() is read only by the profiling instantiation (cppf_debug_heads_tc_profile); the production one carries no counters
template <bool kProf>
__device__ __forceinline__ long long prof_clock() {
    if constexpr (kProf) return clock64();
    else return 0;
}

template <bool kProf>
__global__ void __launch_bounds__(kThreads, 1) chain_tc_kernel(const __grid_constant__ Program prog, const __grid_constant__ Args one,
                                                               const MultiArgs *__restrict__ multi) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem_base;
    const uint32_t ring0 = smem_u32(smem + kSmemRing);
    const uint32_t bar0 = smem_u32(smem + kSmemBar);
    const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * kStages, bar_act = bar0 + 16 * kStages, bar_done = bar_act + 8 * kSlots;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, (CPPF_TC_EXP & 64) ? 1 : kSlots);     // a slab is shared: both slots' MMAs must retire before it is refilled
        }
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(bar_act + 8 * s, kEpiWarps);      // every epilogue warp arrives for whichever slot it has just served
            mbar_init(bar_done + 8 * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 2 * kRows) {   // ones operand, K-major core-matrix layout like X: plane 0 = columns 0..7, plane 1 = columns 8..15
        const uint32_t one2 = 0x3f803f80u;   // bf16 (1, 1)
        *reinterpret_cast<uint4 *>(smem + kSmemOnes + tid * 16) = tid < kRows ? make_uint4(one2, 0, 0, 0) : make_uint4(0, 0, 0, 0);
    }
    fence_async_smem();
    if (warp == 0) {  // the whole TMEM: two slots x 256 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;
    if (tmem != 0) __trap();     // one CTA per SM owning all 512 columns: the allocation starts at lane 0, column 0
    pdl_enter();                 // frame path (programmatic dependent launch): barriers and tensor memory are set up; everything
                                 // below reads what the previous kernel of the frame wrote
    bool single;
    const int n_ctas = cta_share(multi, one, single);               // CTAs that take part (see pair_of)
    const int64_t q0 = static_cast<int>(blockIdx.x) < n_ctas ? static_cast<int64_t>(blockIdx.x) : (1ll << 60);   // others: no pair

    if (warp == kEpiWarps && !(CPPF_TC_EXP & 256)) {
        // =============================== weight producer ===============================================
        // One pass over the slab stream per round; every slab serves both slots of the round.  The whole warp runs
        // the (uniform) loop; the elected lane issues the copies.
        const bool leader = elect_one();
        uint32_t seq = 0;
        long long t_wait = 0;
        const long long t_begin = prof_clock<kProf>();
        for (int round = 0;; ++round) {
            int job;
            int64_t tile0, n_tiles;
            if (!pair_of(multi, one, single, static_cast<int64_t>(round) * n_ctas + q0, job, tile0, n_tiles)) break;
            const unsigned char *src = multi ? multi->job[job].weights : one.weights;
            for (int i = 0; i < prog.n_slabs; ++i, ++seq) {
                const uint32_t bytes = prog.slab[i].bytes();
                const uint32_t stage = seq % kStages, turn = seq / kStages;
                const long long t0 = prof_clock<kProf>();
                mbar_wait(bar_empty + 8 * stage, (turn & 1u) ^ 1u);
                t_wait += prof_clock<kProf>() - t0;
                if (leader) {
                    mbar_expect_tx(bar_full + 8 * stage, bytes);
                    bulk_load(ring0 + stage * kSlabBytes, src, bytes, bar_full + 8 * stage);
                }
                __syncwarp();
                src += bytes;
            }
        }
        if (kProf && one.prof && leader) {
            one.prof[blockIdx.x * 64 + 4] = prof_clock<kProf>() - t_begin;
            one.prof[blockIdx.x * 64 + 5] = t_wait;
        }
    } else if (warp == kEpiWarps) {
        // (experiment 8: no producer)
    } else if (warp > kEpiWarps) {
        // =============================== MMA issuers (one warp per slot) ================================
        // Each slot has its own issuing warp walking the slab list in order with blocking waits, exactly like a GEMM
        // main loop; the two warps interleave on the tensor pipe by themselves, and a slab is released to the producer
        // when both have retired their MMAs on it (bar_empty counts two).  The warp runs converged on warp-uniform
        // state (loop counters, kernel-parameter loads), so every descriptor lives in uniform registers; only the
        // tcgen05 instructions are predicated on the elected lane.
        const int s = warp - (kEpiWarps + 1);
        const bool leader = elect_one();
        long long t_act = 0, t_full = 0, t_issue = 0, n_steps = 0;
        const long long t_begin = prof_clock<kProf>();
        constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);    // SBO = 128 B, descriptor version bit 46
        const uint32_t ones_lo = (smem_u32(smem + kSmemOnes) >> 4) | (static_cast<uint32_t>(kPlane >> 4) << 16);
        const uint32_t x_lo = smem_u32(smem + kSmemX + s * kXBytes) >> 4;
        const uint32_t ring_lo = ring0 >> 4;
        const uint32_t t_slot = static_cast<uint32_t>(s * kSlotTmem);     // TMEM base is 0: the CTA owns all 512 columns (checked above)
        uint32_t stage = 0, turn = 0, act_par = 0;
        for (int round = 0;; ++round) {
            int job;
            int64_t tile0, n_tiles;
            if (!pair_of(multi, one, single, static_cast<int64_t>(round) * n_ctas + q0, job, tile0, n_tiles)) break;
            if (tile0 + s >= n_tiles) {
                // No tile for this slot in this round (a job's odd last tile, or single-tile units).  The slot still walks
                // the weight ring slab by slab -- wait for the fill, release it -- so that it stays within one ring turn of
                // the producer: mbarrier waits are by phase PARITY, and a warp that skipped a whole round (~30 turns) would
                // pass or block on the wrong phases in its next round.
                for (int i = 0; i < prog.n_slabs; ++i) {
                    if ((CPPF_TC_EXP & 256) || ((CPPF_TC_EXP & 64) && s == 1)) break;
                    mbar_wait(bar_full + 8 * stage, turn);
                    if (leader) mbar_arrive(bar_empty + 8 * stage);
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        turn ^= 1u;
                    }
                }
                continue;
            }
            for (int i = 0; i < prog.n_slabs; ++i) {
                const uint4 w = *reinterpret_cast<const uint4 *>(&prog.slab[i]);      // a_lo, idesc, nd, misc
                const uint32_t flags = w.w >> 24;
                long long t0 = prof_clock<kProf>();
                if (flags & kSlabFirst) {        // the slot's A operand is ready and its accumulator is free
                    if (CPPF_TC_EXP & 1024) mbar_spin(bar_act + 8 * s, act_par);
                    else mbar_wait(bar_act + 8 * s, act_par);
                    act_par ^= 1u;
                    tc_fence_after();            // the epilogue's tcgen05.ld/st of this slot are ordered before the MMAs below
                }
                long long t1 = prof_clock<kProf>();
                // weights arrive through the async proxy (cp.async.bulk, complete_tx on this barrier): no tcgen05 fence needed
                const bool no_ring = (CPPF_TC_EXP & 256) || ((CPPF_TC_EXP & 64) && s == 1);
                if (!no_ring) mbar_wait(bar_full + 8 * stage, turn);
                long long t2 = prof_clock<kProf>();
                if (leader && !(CPPF_TC_EXP & 32)) {
                    // every value below is warp-uniform: descriptors stay in uniform registers, a handful of adds per MMA
                    const uint32_t idesc = w.y, n_mma = (w.w >> 16) & 0xffu;
                    const uint32_t d_addr = t_slot + (w.z & 0xffffu);
                    const uint32_t b_lo = (ring_lo + stage * static_cast<uint32_t>(kSlabBytes >> 4)) | (w.z & 0xffff0000u);   // LBO field = n
                    const uint32_t b_step = (w.z >> 16) << 1;                                  // 32 n bytes per K = 16 step
                    const uint32_t acc0 = (flags >> 3) & 1u;                                   // kSlabAcc
                    constexpr uint64_t kHi = static_cast<uint64_t>(kDescHi) << 32;
                    if (flags & kSlabATmem) {
                        const uint32_t a0 = t_slot + w.x;
                        umma_ts(d_addr, a0, kHi | b_lo, idesc, acc0);
                        if (n_mma > 1) {
                            umma_ts(d_addr, a0 + 8u, kHi | (b_lo + b_step), idesc, 1u);
                            if (n_mma > 2) {
                                umma_ts(d_addr, a0 + 16u, kHi | (b_lo + 2u * b_step), idesc, 1u);
                                if (n_mma > 3) umma_ts(d_addr, a0 + 24u, kHi | (b_lo + 3u * b_step), idesc, 1u);
                            }
                        }
                    } else {
                        constexpr uint32_t kAStep = static_cast<uint32_t>((2 * kPlane) >> 4);
                        const uint32_t a0 = ((flags & kSlabOnes) ? ones_lo : x_lo) + w.x;
                        umma_ss(d_addr, kHi | a0, kHi | b_lo, idesc, acc0);
                        if (n_mma > 1) {
                            umma_ss(d_addr, kHi | (a0 + kAStep), kHi | (b_lo + b_step), idesc, 1u);
                            if (n_mma > 2) {
                                umma_ss(d_addr, kHi | (a0 + 2u * kAStep), kHi | (b_lo + 2u * b_step), idesc, 1u);
                                if (n_mma > 3) umma_ss(d_addr, kHi | (a0 + 3u * kAStep), kHi | (b_lo + 3u * b_step), idesc, 1u);
                            }
                        }
                    }
                    if (!no_ring) umma_commit(bar_empty + 8 * stage);   // this slot is done with the slab once these MMAs retire
                    if (flags & kSlabLast) umma_commit(bar_done + 8 * s);
                }
                __syncwarp();
                if (++stage == kStages) {
                    stage = 0;
                    turn ^= 1u;
                }
                t_act += t1 - t0;
                t_full += t2 - t1;
                t_issue += prof_clock<kProf>() - t2;
                ++n_steps;
            }
        }
        if (kProf && one.prof && leader) {
            long long *o = one.prof + blockIdx.x * 64 + 40 + 8 * s;
            o[0] = prof_clock<kProf>() - t_begin;
            o[1] = t_act;
            o[2] = t_full;
            o[3] = t_issue;
            o[4] = n_steps;
        }
    } else {
        // =============================== prologue / epilogue warps (all 16 serve both slots) ==============
        // The 16 warps work on ONE slot at a time, alternating: action(slot 0, phase p), action(slot 1, phase p), action(slot 0,
        // p + 1) ...  While they convert slot 1's accumulator the tensor pipe runs the MMAs slot 0's arrive just released, and
        // vice versa (ping-pong).  With 8 dedicated warps per slot each slot's epilogue took about as long as both slots' MMAs
        // together and its warps then idled for the MMAs; with 16 warps on the ready slot an action takes half as long and no
        // epilogue warp ever waits while the other slot has work.  TMEM lane access fixes the row ownership: warp w reads lanes
        // 32 (w % 4) .. +31, so the four warps of a lane quadrant split a chunk's columns four ways (two ways for the narrow
        // chunks of the scale head).
        const int ew = warp, etid = tid;                 // 0..15, 0..511
        const int row = (ew & 3) * 32 + lane;            // TMEM lane == tile row
        const int quarter = ew >> 2;                     // which share of the chunk's columns this thread handles
        // (8 rows x 4 chunks) per warp pass for the row movers: conflict-free 16-byte shared stores, whole 32 B sectors
        const int mv_r = lane & 7, mv_c = lane >> 3;
        uint32_t done_seq[kSlots] = {0, 0};
        long long t_done = 0, t_actn[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t_arrive = 0;
        const long long t_begin = prof_clock<kProf>();
        for (int round = 0;; ++round) {
            int job;
            int64_t tile0, n_tiles;
            if (!pair_of(multi, one, single, static_cast<int64_t>(round) * n_ctas + q0, job, tile0, n_tiles)) break;
            const Args &a = multi ? multi->job[job] : one;
            const int n_active = tile0 + 1 < n_tiles ? 2 : 1;
            if (a.idx.ptr != nullptr) {
                // the tiles' tuple indices, once: every gather below reads them from shared memory
                for (int sl = 0; sl < n_active; ++sl) {
                    int *idx_sl = reinterpret_cast<int *>(smem + kSmemIdx) + sl * kRows * kIdxStride;
                    const int64_t rb = (tile0 + sl) * kRows;
                    for (int e = etid; e < kRows * a.arity; e += kEpiWarps * 32) {
                        const int r = e / a.arity, k = e - r * a.arity;
                        idx_sl[r * kIdxStride + k] = rb + r < a.rows ? static_cast<int>(a.idx.at(rb + r, k)) : 0;
                    }
                }
                named_bar(1, kEpiWarps * 32);
            }
            for (int p = 0; p < prog.n_phases; ++p) {
                const Phase &ph = prog.phase[p];
                for (int slot = 0; slot < n_active; ++slot) {
                    unsigned char *X = smem + kSmemX + slot * kXBytes;
                    const uint32_t x_u32 = smem_u32(X);
                    const int *idx_s = reinterpret_cast<const int *>(smem + kSmemIdx) + slot * kRows * kIdxStride;
                    const uint32_t t_slot_lane = tmem + static_cast<uint32_t>(slot * kSlotTmem) + (static_cast<uint32_t>((ew & 3) * 32) << 16);
                    const int64_t row_base = (tile0 + slot) * kRows;
                    const int64_t grow = row_base + row;
                    const bool live = grow < a.rows;
                    long long t0 = prof_clock<kProf>();
                    if (ph.wait_done) {
                        if (CPPF_TC_EXP & 512) mbar_spin(bar_done + 8 * slot, done_seq[slot] & 1u);
                        else mbar_wait(bar_done + 8 * slot, done_seq[slot] & 1u);
                        ++done_seq[slot];
                        tc_fence_after();
                    }
                    long long t1 = prof_clock<kProf>();
                    t_done += t1 - t0;
                    switch (ph.action) {
                        case kActLoadRows: {
                            const int cb_n = ph.width >> 5;                  // blocks of 4 chunks
                            for (int it0 = ew; it0 < 16 * cb_n; it0 += 4 * kEpiWarps) {
                                float4 lo[4], hi[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {               // four independent row segments in flight
                                    const int it = it0 + u * kEpiWarps;
                                    const int r = (it & 15) * 8 + mv_r, c8 = (it >> 4) * 4 + mv_c;
                                    const int64_t gr = row_base + r;
                                    lo[u] = hi[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                                    if (it < 16 * cb_n && gr < a.rows && c8 * 8 < ph.cols) {      // cols is a multiple of 8
                                        const float4 *src = reinterpret_cast<const float4 *>(a.x + gr * a.x_ld + ph.src_col + c8 * 8);
                                        lo[u] = __ldg(src);
                                        hi[u] = __ldg(src + 1);
                                    }
                                }
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int it = it0 + u * kEpiWarps;
                                    if (it >= 16 * cb_n) break;
                                    const int r = (it & 15) * 8 + mv_r, c8 = (it >> 4) * 4 + mv_c;
                                    float v[8] = {lo[u].x, lo[u].y, lo[u].z, lo[u].w, hi[u].x, hi[u].y, hi[u].z, hi[u].w};
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[j] = (v[j] == v[j]) ? v[j] : 0.0f;   // NaN rows of invalid SHOT points count as zeros (eval.py:215)
                                    *reinterpret_cast<uint4 *>(act_chunk(X, r, c8)) = pack8(v);
                                }
                            }
                            break;
                        }
                        case kActEncodeShotA: {
                            // features (64 bf16 = one 128-byte line per point) of tuple slots 0..3 -> X[0:256): a warp pass copies the
                            // whole lines of 4 rows (see kActGatherSum for why not 8 half lines); asynchronous 16-byte copies
                            // straight into the operand layout, all in flight at once
                            for (int it = ew; it < 32 * 4; it += kEpiWarps) {
                                const int r = (it & 31) * 4 + (lane >> 3), k = it >> 5, c8 = k * 8 + (lane & 7);
                                const bool ok = row_base + r < a.rows;
                                cp_async16(x_u32 + c8 * kPlane + r * 16, a.point_feat + static_cast<int64_t>(idx_s[r * kIdxStride + k]) * 64 + (lane & 7) * 8,
                                           ok ? 16u : 0u);
                            }
                            cp_async_wait_all();
                            break;
                        }
                        case kActGather: {
                            // one 256-wide per-point row (tuple slot src_col) -> X[0:256)
                            for (int it = ew; it < 32 * 4; it += kEpiWarps) {
                                const int r = (it & 31) * 4 + (lane >> 3), c8 = (it >> 5) * 8 + (lane & 7);
                                const bool ok = row_base + r < a.rows;
                                cp_async16(x_u32 + c8 * kPlane + r * 16,
                                           a.point_feat + static_cast<int64_t>(idx_s[r * kIdxStride + ph.src_col]) * prog.gather_cols + c8 * 8, ok ? 16u : 0u);
                            }
                            cp_async_wait_all();
                            break;
                        }
                        case kActEncodeShotB: {
                            for (int it = ew; it < 32; it += kEpiWarps) {   // features of tuple slot 4 -> columns 0..63, whole lines
                                const int r = it * 4 + (lane >> 3), c8 = lane & 7;
                                const bool ok = row_base + r < a.rows;
                                cp_async16(x_u32 + c8 * kPlane + r * 16, a.point_feat + static_cast<int64_t>(idx_s[r * kIdxStride + 4]) * 64 + c8 * 8, ok ? 16u : 0u);
                            }
                            if (etid < 2 * kRows) {   // geometry of row etid % 128 (two threads per row): coords -> columns 64..93, normals -> 94..103, zeros -> 104..111
                                const int r = etid & (kRows - 1);
                                tuple_geometry_chunks(a.pc, a.normal, idx_s + r * kIdxStride, row_base + r < a.rows, true, etid >> 7, X, r, 8);
                            }
                            cp_async_wait_all();
                            break;
                        }
                        case kActGatherSum: {
                            // desc_pair_transform by linearity (train_dino.py:95-96): W [256, 5*256] applied to the concatenation of
                            // the 5 transformed descriptors = sum_k W_k f(desc[idx_k]); the per-point program has stored
                            // G[n][256 k : 256 k + 256) = W_k f(desc_n) (+ bias in block 0) as bf16, so a tuple costs five 512-byte row
                            // gathers and no tensor work.  Sum in float32, one rounding to bf16 into the operand layout.
                            const int64_t ld = prog.gather_cols;
#pragma unroll 1
                            for (int it = ew; it < 32 * 4; it += kEpiWarps) {
                                // a warp pass = 4 rows x 128 contiguous bytes: whole 128-byte lines per L2 request (a 64-byte
                                // half-line mapping moves the same sectors with twice the requests); the shared-memory stores
                                // then land 2-way conflicted, which is the cheaper side.  (Two passes in flight at once -- ten
                                // loads per lane -- measured the same: the gather is not bound by the loads in flight.)
                                const int r = (it & 31) * 4 + (lane >> 3), c8 = (it >> 5) * 8 + (lane & 7);
                                uint4 g[5];
#pragma unroll
                                for (int k = 0; k < 5; ++k)          // five independent 16-byte loads in flight per lane
                                    g[k] = __ldg(reinterpret_cast<const uint4 *>(a.point_feat + static_cast<int64_t>(idx_s[r * kIdxStride + k]) * ld + k * 256 + c8 * 8));
                                float acc[8], v[8];
                                unpack8(g[0], acc);
#pragma unroll
                                for (int k = 1; k < 5; ++k) {
                                    unpack8(g[k], v);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) acc[j] += v[j];
                                }
                                *reinterpret_cast<uint4 *>(act_chunk(X, r, c8)) = row_base + r < a.rows ? pack8(acc) : make_uint4(0, 0, 0, 0);
                            }
                            break;
                        }
                        case kActCoordsB: {
                            if (etid < 2 * kRows) {   // coords -> columns 0..29, zeros -> 30..31
                                const int r = etid & (kRows - 1);
                                tuple_geometry_chunks(a.pc, a.normal, idx_s + r * kIdxStride, row_base + r < a.rows, false, etid >> 7, X, r, 0);
                            }
                            break;
                        }
                        case kActHiddenT:
                        case kActOutT:
                        case kActHiddenS:
                        case kActOut:
                        case kActFinal: {
                            // wide chunks (128 / 256 columns): four threads per row, 32 / 64 columns each; narrow chunks (16 / 64): the
                            // first two warps of the quadrant, 8 / 32 columns each
                            if (ph.n >= 128) {
                                const int per = ph.n >> 2;
                                for (int c = quarter * per; c < (quarter + 1) * per; c += 32) epilogue_batch<32>(ph, a, X, t_slot_lane, row, grow, live, c);
                            } else if (quarter < 2) {
                                const int per = ph.n >> 1;
                                if (per >= 32) epilogue_batch<32>(ph, a, X, t_slot_lane, row, grow, live, quarter * per);
                                else epilogue_batch<8>(ph, a, X, t_slot_lane, row, grow, live, quarter * per);
                            }
                            if (ph.action == kActHiddenT || ph.action == kActOutT) tmem_st_wait();
                            break;
                        }
                        default: break;
                    }
                    t0 = prof_clock<kProf>();
                    if (kProf) t_actn[ph.action] += t0 - t1;
                    if (ph.n_parts) {
                        if (ph.action != kActHiddenT && ph.action != kActOutT && !(CPPF_TC_EXP & 128)) fence_async_smem();   // generic-proxy writes to X -> visible to the tensor core's async proxy
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_act + 8 * slot);
                    }
                    t_arrive += prof_clock<kProf>() - t0;
                }
            }
        }
        if (kProf && one.prof && ew == 0 && lane == 0) {
            long long *o = one.prof + blockIdx.x * 64 + 8;
            o[0] = prof_clock<kProf>() - t_begin;
            o[1] = t_done;
            for (int k = 0; k < 10; ++k) o[2 + k] = t_actn[k];
            o[12] = t_arrive;
            o[13] = t_actn[10];
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
    std::vector<std::vector<Part>> parts = std::vector<std::vector<Part>>(kMaxPhases);   // per phase, in stream order
    const float *w;

    static uint16_t bf16_bits(float f) {
        uint32_t u;
        memcpy(&u, &f, 4);
        if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;               // NaN
        const uint32_t lsb = (u >> 16) & 1u;
        return static_cast<uint16_t>((u + 0x7fffu + lsb) >> 16);            // round to nearest even
    }
    static int pad16(int v) { return (v + 15) & ~15; }
    // accumulator chunk widths the epilogue supports
    static int chunk_pad(int v) { return v <= 16 ? 16 : (v <= 64 ? 64 : 128); }

    // Appends the UMMA image of rows [n0, n0+n_real) x the listed source columns of W [.][ld], padded to n_pad x k_pad;
    // returns its byte offset in the stream.  cols[k] < 0 is a zero column.
    int64_t push_weight(const float *W, int ld, int n0, int n_real, int n_pad, const std::vector<int> &cols, int k_pad) {
        const size_t off = stream.size();
        stream.resize(off + static_cast<size_t>(n_pad) * k_pad, 0);
        for (size_t k = 0; k < cols.size(); ++k) {
            if (cols[k] < 0) continue;
            for (int n = 0; n < n_real; ++n)
                stream[off + (k >> 3) * static_cast<size_t>(n_pad) * 8 + static_cast<size_t>(n) * 8 + (k & 7)] =
                    bf16_bits(W[static_cast<size_t>(n0 + n) * ld + cols[k]]);
        }
        return static_cast<int64_t>(off) * 2;
    }
    // Bias of the accumulator chunk D[d_col : d_col + n_pad) of phase `ph`: one more K-step of 16 whose B rows are
    // k = 0: bf16(bias), k = 1: bf16(bias - bf16(bias)) and whose A operand is the constant ones matrix.
    void add_bias(Phase &ph, int d_col, const float *a, const float *b, int n0, int n_real, int n_pad) {
        const size_t off = stream.size();
        stream.resize(off + static_cast<size_t>(n_pad) * 16, 0);
        for (int n = 0; n < n_real; ++n) {
            const float v = a[n0 + n] + (b ? b[n0 + n] : 0.0f);
            const uint16_t hi = bf16_bits(v);
            uint32_t hb = static_cast<uint32_t>(hi) << 16;
            float hf;
            memcpy(&hf, &hb, 4);
            stream[off + static_cast<size_t>(n) * 8 + 0] = hi;
            stream[off + static_cast<size_t>(n) * 8 + 1] = bf16_bits(v - hf);
        }
        add_part(ph, 2, 0, 16, d_col, n_pad, 0, static_cast<int64_t>(off) * 2);
    }
    static std::vector<int> iota(int n, int start = 0) {
        std::vector<int> v(n);
        for (int i = 0; i < n; ++i) v[i] = start + i;
        return v;
    }

    Phase &add(int action, int wait_done) {
        Phase &ph = prog.phase[prog.n_phases++];
        ph = Phase{};
        ph.action = action;
        ph.wait_done = wait_done;
        return ph;
    }
    void add_part(Phase &ph, int a_tmem, int a_col, int k_cols, int d_col, int n, int init, int64_t w_off) {
        parts[&ph - prog.phase].push_back(Part{a_tmem, a_col, k_cols, d_col, n, init, w_off});
        ++ph.n_parts;
    }

    // Phases -> the slab list the producer and the MMA issuer walk.  Returns false when it does not fit.
    bool finalize() {
        int64_t expect = 0;
        prog.n_slabs = 0;
        for (int p = 0; p < prog.n_phases; ++p) {
            const std::vector<Part> &ps = parts[p];
            for (size_t q = 0; q < ps.size(); ++q) {
                const Part &pt = ps[q];
                if (pt.w_off != expect) return false;               // the stream must be consumed in the order it was written
                const int slab_k = pt.n <= 128 ? 64 : 32;
                for (int k_done = 0; k_done < pt.k_cols; k_done += slab_k) {
                    if (prog.n_slabs >= kMaxSlabs) return false;
                    const int cols = pt.k_cols - k_done < slab_k ? pt.k_cols - k_done : slab_k;
                    Slab &sl = prog.slab[prog.n_slabs++];
                    sl = Slab{};
                    sl.nd = (static_cast<uint32_t>(pt.n) << 16) | static_cast<uint32_t>(pt.d_col);
                    sl.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(pt.n >> 3) << 17) | (static_cast<uint32_t>(kRows >> 4) << 24);
                    uint32_t flags = 0;
                    const int col = pt.a_col + k_done;
                    if (pt.a_tmem == 1) {
                        flags |= kSlabATmem;
                        sl.a_lo = static_cast<uint32_t>(kHTmem + (col >> 1));
                    } else if (pt.a_tmem == 2) {
                        flags |= kSlabOnes;
                    } else {
                        sl.a_lo = (static_cast<uint32_t>((col >> 3) * kPlane) >> 4) | (static_cast<uint32_t>(kPlane >> 4) << 16);
                    }
                    if (!(pt.init && k_done == 0)) flags |= kSlabAcc;
                    if (q == 0 && k_done == 0) flags |= kSlabFirst;
                    if (q + 1 == ps.size() && k_done + cols >= pt.k_cols) flags |= kSlabLast;
                    sl.misc = static_cast<uint32_t>(pt.n * 2 * cols / 16) | (static_cast<uint32_t>(cols / 16) << 16) | (flags << 24);
                    expect += static_cast<int64_t>(pt.n) * 2 * cols;
                }
            }
        }
        return expect == static_cast<int64_t>(stream.size()) * 2;
    }

    // ResLayer whose input (din <= 256 columns) sits in X at in_col; `cur` is the phase whose action produced it.
    // Output chunks go to X at out_col.  Returns the phase(s) that carry the output epilogue: the LAST one is
    // returned, earlier output chunks (256- and 192-wide layers) are complete Out phases unless `final_out`.
    struct OutSpec {
        int action = kActOut;       // kActOut or kActFinal
        int out_col = 0;
        int out_sel = 0, out_ld = 0;
    };
    // in_tmem: the layer input is a bf16 activation in tensor memory (column in_col of the H/Y space, i.e. Y = 128..255),
    // fed to fc1 and fc0 as the A operand from TMEM; din must then be a multiple of 16 (no zero padding in TMEM).
    Phase *res_layer(const ResLayerDesc &L, Phase *cur, int in_col, const OutSpec &os, int in_tmem = 0) {
        const int k_in = pad16(L.din);
        const std::vector<int> in_cols = iota(L.din);
        // hidden / output chunks of at most 128 columns
        std::vector<std::pair<int, int>> chunks;   // (first column, real width)
        for (int c = 0; c < L.dout; c += 128) chunks.push_back({c, L.dout - c < 128 ? L.dout - c : 128});
        int h_width = 0;
        for (auto &ch : chunks) {
            const int n_pad = chunk_pad(ch.second);
            add_part(*cur, in_tmem, in_col, k_in, 0, n_pad, 1, push_weight(w + L.w1, L.din, ch.first, ch.second, n_pad, in_cols, k_in));
            Phase &hid = add(kActHiddenT, 1);
            hid.d_col = 0;
            hid.n = n_pad;
            hid.dst_col = ch.first;
            add_bias(*cur, 0, w + L.b1, nullptr, ch.first, ch.second, n_pad);
            cur = &hid;
            h_width = ch.first + n_pad;
        }
        const std::vector<int> h_cols = iota(L.dout);   // hidden column j of H <-> fc2 input j (pad columns are zero in both)
        for (size_t i = 0; i < chunks.size(); ++i) {
            const auto &ch = chunks[i];
            const int n_pad = chunk_pad(ch.second);
            std::vector<int> hc(h_width, -1);
            for (int j = 0; j < L.dout; ++j) hc[j] = j;
            add_part(*cur, 1, 0, h_width, 0, n_pad, 1, push_weight(w + L.w2, L.dout, ch.first, ch.second, n_pad, hc, h_width));
            if (L.has_fc0) add_part(*cur, in_tmem, in_col, k_in, 0, n_pad, 0, push_weight(w + L.w0, L.din, ch.first, ch.second, n_pad, in_cols, k_in));
            add_bias(*cur, 0, w + L.b2, L.has_fc0 ? w + L.b0 : nullptr, ch.first, ch.second, n_pad);
            Phase &out = add(os.action, 1);
            out.d_col = 0;
            out.n = n_pad;
            out.cols = ch.second;
            out.dst_col = os.out_col + ch.first;
            out.residual = L.has_fc0 ? 0 : (in_tmem ? 2 : 1);
            out.res_col = in_col + ch.first;
            out.out_sel = os.out_sel;
            out.out_ld = os.out_ld;
            cur = &out;
        }
        return cur;
    }

    // First ResLayer of a stack whose input is wider than 256 columns (always din != dout, dout = 128): the input
    // arrives as two K-chunks through X[0:...).  `first` is the phase whose action produced chunk A (256 columns,
    // source columns cols_a of the layer input); chunk B (cols_b, padded to kb) is produced by `action_b`.
    // out_tmem: the layer output goes to tensor memory (Y = H/Y columns 128..255 = TMEM columns 192..255) as the A operand of
    // the next layer.  The accumulators then swap sides -- fc1 in D[128:256), the output sum in D[0:128) -- so that the
    // output epilogue never writes columns another warp is still reading.
    Phase *wide_first_layer(const ResLayerDesc &L, Phase *first, const std::vector<int> &cols_a, int action_b,
                            const std::vector<int> &cols_b, int kb, int out_col, Phase **phase_b, bool out_tmem = false) {
        const int n = 128;   // dout
        const int d_hid = out_tmem ? 128 : 0, d_out = out_tmem ? 0 : 128;
        add_part(*first, 0, 0, 256, d_hid, n, 1, push_weight(w + L.w1, L.din, 0, L.dout, n, cols_a, 256));
        add_part(*first, 0, 0, 256, d_out, n, 1, push_weight(w + L.w0, L.din, 0, L.dout, n, cols_a, 256));
        Phase &pb = add(action_b, 1);
        add_part(pb, 0, 0, kb, d_hid, n, 0, push_weight(w + L.w1, L.din, 0, L.dout, n, cols_b, kb));
        add_part(pb, 0, 0, kb, d_out, n, 0, push_weight(w + L.w0, L.din, 0, L.dout, n, cols_b, kb));
        add_bias(pb, d_hid, w + L.b1, nullptr, 0, L.dout, n);
        if (phase_b) *phase_b = &pb;
        Phase &hid = add(kActHiddenS, 1);     // H -> X[0:128) (chunk B has been consumed)
        hid.d_col = d_hid;
        hid.n = n;
        hid.dst_col = 0;
        add_part(hid, 0, 0, n, d_out, n, 0, push_weight(w + L.w2, L.dout, 0, L.dout, n, iota(L.dout), n));
        add_bias(hid, d_out, w + L.b2, w + L.b0, 0, L.dout, n);
        Phase &out = add(out_tmem ? kActOutT : kActOut, 1);
        out.d_col = d_out;
        out.n = n;
        out.cols = L.dout;
        out.dst_col = out_tmem ? 128 : out_col;
        return &out;
    }
};

struct State {
    HeadsModel model;
    Program point_prog, tuple_prog;     // passed by value as kernel parameters
    unsigned char *d_point_w, *d_tuple_w;
    int point_cols;
};

static int upload(const Builder &b, Program *prog, unsigned char **d_w) {
    *prog = b.prog;
    CPPF_CUDA_TRY(cudaMalloc(d_w, b.stream.size() * 2 + 16));
    CPPF_CUDA_TRY(cudaMemcpy(*d_w, b.stream.data(), b.stream.size() * 2, cudaMemcpyHostToDevice));
    return CPPF_OK;
}

}  // namespace tc
}  // namespace cppf

using namespace cppf;
using namespace cppf::tc;

// Appends a stack whose layers all take <= 256 input columns.  The layer outputs alternate so that a layer that
// widens 128 -> 256 finds its input in X[128:256) and can write its first output chunk to X[0:128) while the
// second GEMM still reads the input.
//
// tmem_chain: a 128-wide layer output that feeds a layer whose hidden activation is at most 128 wide stays in tensor memory
// (Y, next to H) as bf16 -- the next fc1 takes it as its A operand from there and the identity residual is read back from
// there -- so that a run of 128 -> 128 ResLayers never stores to shared memory: no generic-proxy stores competing with the
// MMAs' operand reads, no fence.proxy.async before the hand-off.  Same roundings (bf16 once per layer output) either way.
static bool tmem_chain_enabled() {
    const char *e = getenv("CPPF_TC_TMEM_CHAIN");
    return !(e && e[0] == '0');
}
static bool next_takes_tmem(const ResLayerDesc &L, const ResLayerDesc &nx) {
    return tmem_chain_enabled() && L.dout == 128 && nx.din == 128 && nx.dout <= 128;
}

static Phase *append_stack(Builder &b, const StackDesc &s, int first_layer, Phase *cur, int in_col, const Builder::OutSpec &last,
                           int in_tmem = 0) {
    for (int l = first_layer; l < s.n_layers; ++l) {
        const ResLayerDesc &L = s.layer[l];
        Builder::OutSpec os;
        if (l == s.n_layers - 1) {
            os = last;
        } else {
            const ResLayerDesc &nx = s.layer[l + 1];
            if (next_takes_tmem(L, nx)) {
                os.action = kActOutT;
                os.out_col = 128;          // Y
            } else {
                // the next layer widens beyond its input and has a projection: park this output in the upper half
                os.out_col = (nx.has_fc0 && nx.dout > 128 && L.dout <= 128) ? 128 : 0;
            }
        }
        cur = b.res_layer(L, cur, in_col, os, in_tmem);
        in_col = os.out_col;
        in_tmem = os.action == kActOutT ? 1 : 0;
    }
    return cur;
}

// Host only: the two programs and their weight streams (no CUDA call).
static int build_programs(const HeadsModel &m, const float *w, Builder &pb, Builder &tb, int *point_cols) {
    if (m.arity != 5) return CPPF_ERR_UNSUPPORTED;   // the tile layouts below assume 5-point tuples (10 pairs)
    struct { int point_cols; } stv{0}, *st = &stv;
    // ---- per-point program -------------------------------------------------------------------------------
    pb.w = w;
    if (m.branch == 0) {   // shot_encoder: [n,352] -> [n,64] bf16; first layer 352 = 256 + 96
        Phase *cur = &pb.add(kActLoadRows, 0);
        cur->src_col = 0;
        cur->cols = 256;
        cur->width = 256;
        Phase *pbb = nullptr;
        const bool pt_tmem = next_takes_tmem(m.shot_encoder.layer[0], m.shot_encoder.layer[1]);
        cur = pb.wide_first_layer(m.shot_encoder.layer[0], cur, Builder::iota(256), kActLoadRows, Builder::iota(96, 256), 96, 0, &pbb, pt_tmem);
        pbb->src_col = 256;
        pbb->cols = 96;
        pbb->width = 96;
        Builder::OutSpec last;
        last.action = kActFinal;
        last.out_sel = 2;
        last.out_ld = 64;
        cur = append_stack(pb, m.shot_encoder, 1, cur, pt_tmem ? 128 : 0, last, pt_tmem ? 1 : 0);
        st->point_cols = 64;
    } else {
        // desc_transform (train_dino.py:80,95) hoisted per point: Linear 1024 -> 256 as four K-chunks, output f kept in X as
        // bf16; then desc_pair_transform (train_dino.py:81,96) hoisted too: its weight [256, 5*256] splits into five
        // 256x256 blocks W_k, one per tuple slot, and G[n][256 k : 256 k + 256) = W_k f_n (+ bias in block 0) is stored per
        // point, so that the tuple program only gathers and sums (kActGatherSum): N << 5 T makes this ~18x fewer MACs.
        const LinearDesc &L = m.desc_transform;
        const LinearDesc &P = m.desc_pair_transform;
        for (int c = 0; c < 4; ++c) {
            Phase &ph = pb.add(kActLoadRows, c > 0);     // X is refilled only after the previous chunk's MMAs retired
            ph.src_col = c * 256;
            ph.cols = 256;
            ph.width = 256;
            pb.add_part(ph, 0, 0, 256, 0, 256, c == 0, pb.push_weight(w + L.w, L.din, 0, 256, 256, Builder::iota(256, c * 256), 256));
        }
        pb.add_bias(pb.prog.phase[pb.prog.n_phases - 1], 0, w + L.b, nullptr, 0, 256, 256);
        Phase *cur = &pb.add(kActOut, 1);                // X[0:256) = bf16(f)
        cur->d_col = 0;
        cur->n = 256;
        cur->cols = 256;
        cur->dst_col = 0;
        for (int k = 0; k < m.arity; ++k) {
            pb.add_part(*cur, 0, 0, 256, 0, 256, 1, pb.push_weight(w + P.w, P.din, 0, 256, 256, Builder::iota(256, k * 256), 256));
            if (k == 0) pb.add_bias(*cur, 0, w + P.b, nullptr, 0, 256, 256);
            cur = &pb.add(kActFinal, 1);                 // stores block k while (for k < 4) block k + 1 is not yet issued
            cur->d_col = 0;
            cur->n = 256;
            cur->cols = 256;
            cur->dst_col = k * 256;
            cur->out_sel = 2;
            cur->out_ld = 256 * m.arity;
        }
        st->point_cols = 256 * m.arity;
    }
    // ---- per-tuple program -------------------------------------------------------------------------------
    tb.w = w;
    Phase *cur;
    const ResLayerDesc &T0 = m.tuple_encoder.layer[0];
    const bool t0_tmem = next_takes_tmem(T0, m.tuple_encoder.layer[1]);
    if (m.branch == 0) {
        // tuple_encoder.0 input (train_shot.py:75-83) = [coords 30 | normals 10 | feats 5x64]; chunk A = feats of
        // slots 0..3 (source columns 40..295), chunk B = [feats of slot 4 (296..359) | coords+normals (0..39) | pad 8]
        cur = &tb.add(kActEncodeShotA, 0);
        std::vector<int> cols_b = Builder::iota(64, 296);
        for (int k = 0; k < 40; ++k) cols_b.push_back(k);
        cur = tb.wide_first_layer(T0, cur, Builder::iota(256, 40), kActEncodeShotB, cols_b, 112, 0, nullptr, t0_tmem);
    } else {
        // desc_pair_transform of the 5 gathered descriptors = sum of 5 per-point row blocks (kActGatherSum, see the
        // per-point program); its output is chunk A of tuple_encoder.0 ([coords 30 | pair 256])
        Phase &pair = tb.add(kActGatherSum, 0);
        cur = tb.wide_first_layer(T0, &pair, Builder::iota(256, 30), kActCoordsB, Builder::iota(30), 32, 0, nullptr, t0_tmem);
    }
    {
        // tuple_encoder output = the feature both heads read, left in X[0:256).  The scale head runs FIRST and keeps its
        // layer outputs in tensor memory (Y, the 64 spare TMEM columns next to H), so X still holds the feature when the
        // logit head starts: no spill of the feature to global memory, no reload.
        Builder::OutSpec feat;
        // layers 1..4 are 128 -> 128; layer 5 widens to 256: append_stack parks layer 4's output in X[128:256)
        cur = append_stack(tb, m.tuple_encoder, 1, cur, t0_tmem ? 128 : 0, feat, t0_tmem ? 1 : 0);
        const StackDesc &S = m.scale_encoder;
        bool scale_in_tmem = S.n_layers >= 2;
        for (int l = 1; l < S.n_layers; ++l) scale_in_tmem = scale_in_tmem && S.layer[l].has_fc0 && S.layer[l].din % 16 == 0 && S.layer[l].din <= 128;
        scale_in_tmem = scale_in_tmem && S.layer[0].has_fc0 && S.layer[0].dout <= 128;
        if (!scale_in_tmem) return CPPF_ERR_UNSUPPORTED;
        for (int l = 0; l < S.n_layers; ++l) {
            Builder::OutSpec os;
            if (l == S.n_layers - 1) {
                os.action = kActFinal;
                os.out_sel = 1;
                os.out_ld = 3;
            } else {
                os.action = kActOutT;
                os.out_col = 128;          // Y
            }
            cur = l == 0 ? tb.res_layer(S.layer[l], cur, 0, os, 0) : tb.res_layer(S.layer[l], cur, 128, os, 1);
        }
        Builder::OutSpec logits;
        logits.action = kActFinal;
        logits.out_sel = 0;
        logits.out_ld = 192;
        cur = append_stack(tb, m.logit_encoder, 0, cur, 0, logits);
    }
    if (pb.prog.n_phases > kMaxPhases || tb.prog.n_phases > kMaxPhases || !pb.finalize() || !tb.finalize()) return CPPF_ERR_UNSUPPORTED;
    pb.prog.gather_cols = tb.prog.gather_cols = st->point_cols;
    *point_cols = st->point_cols;
    return CPPF_OK;
}

extern "C" int cppf_heads_tc_create(const HeadsModel *model, const float *w, void **state) {
    *state = nullptr;
    const HeadsModel &m = *model;
    Builder pb, tb;
    int point_cols = 0;
    int rc = build_programs(m, w, pb, tb, &point_cols);
    if (rc != CPPF_OK) return rc;
    State *st = new State{};
    st->model = m;
    st->point_cols = point_cols;
    rc = upload(pb, &st->point_prog, &st->d_point_w);
    if (rc == CPPF_OK) rc = upload(tb, &st->tuple_prog, &st->d_tuple_w);
    if (rc != CPPF_OK) {
        delete st;
        return rc;
    }
    CPPF_CUDA_TRY(cudaFuncSetAttribute(chain_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    CPPF_CUDA_TRY(cudaFuncSetAttribute(chain_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    *state = st;
    return CPPF_OK;
}


// Debug hook (host only, no CUDA): prints the phases and MMA parts of the branch's two programs -- the data flow a change to
// the builder must keep consistent (which accumulator / activation region every phase reads and writes).
CPPF_API int cppf_debug_heads_tc_dump(int branch, int num_more) {
    const HeadsModel m = describe_heads(branch, num_more);
    std::vector<float> w(static_cast<size_t>(m.n_floats), 0.0f);
    Builder pb, tb;
    int point_cols = 0;
    const int rc = build_programs(m, w.data(), pb, tb, &point_cols);
    if (rc != CPPF_OK) return rc;
    static const char *names[] = {"LoadRows", "EncShotA", "EncShotB", "Gather", "CoordsB", "HiddenT", "HiddenS", "Out", "Final", "OutT", "GatherSum"};
    const Builder *bs[2] = {&pb, &tb};
    for (int k = 0; k < 2; ++k) {
        const Builder &b = *bs[k];
        printf("== %s program: %d phases, %d slabs, %zu weight bytes\n", k == 0 ? "point" : "tuple", b.prog.n_phases, b.prog.n_slabs,
               b.stream.size() * 2);
        for (int p = 0; p < b.prog.n_phases; ++p) {
            const Phase &ph = b.prog.phase[p];
            printf("  %2d %-9s wait=%d d_col=%3d n=%3d dst=%3d res=%d@%3d out_sel=%d |", p, names[ph.action], ph.wait_done, ph.d_col, ph.n,
                   ph.dst_col, ph.residual, ph.res_col, ph.out_sel);
            for (const Part &pt : b.parts[p])
                printf(" [A=%s%d K=%d -> D%d n=%d %s]", pt.a_tmem == 1 ? "T" : (pt.a_tmem == 2 ? "1" : "X"), pt.a_col, pt.k_cols, pt.d_col, pt.n,
                       pt.init ? "init" : "acc");
            printf("\n");
        }
    }
    return CPPF_OK;
}

static long long *g_tc_prof = nullptr;
// Debug hook: device buffer of [148][64] int64 cycle counters filled by the next launches (nullptr switches it off).
CPPF_API void cppf_debug_heads_tc_profile(void *dev_counters) { g_tc_prof = static_cast<long long *>(dev_counters); }

extern "C" void cppf_heads_tc_destroy(void *state) {
    State *st = static_cast<State *>(state);
    if (!st) return;
    cudaFree(st->d_point_w);
    cudaFree(st->d_tuple_w);
    delete st;
}

static size_t tc_align(size_t x) { return (x + 255) / 256 * 256; }

extern "C" int64_t cppf_heads_tc_workspace_bytes(const void *state, int64_t T, int64_t n) {
    const State *st = static_cast<const State *>(state);
    if (!st) return 0;
    (void)T;      // nothing per tuple lives in the workspace: the tuple feature stays in shared / tensor memory (no spill)
    return static_cast<int64_t>(tc_align(2 * static_cast<size_t>(st->point_cols) * n) + 256);
}

static int64_t pairs_of(int64_t rows) { return ((rows + kRows - 1) / kRows + 1) / 2; }

// At most one CTA per SM; how many of them take part is decided on the device (cta_share): the fewest that still finish
// in the minimal number of rounds.  CPPF_TC_GRID=full is gone with the host-side balancing it switched off.
static int blocks_for_pairs(int64_t pairs) {
    const int sms = device_info().sm_count;
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(2 * pairs, sms)));      // up to one CTA per tile when tiles are few
}

extern "C" int cppf_heads_tc_forward(const void *state, const float *pc, int64_t n, const void *idx, int idx_is_i64,
                                     int64_t idx_stride, int64_t T, const float *feat, const float *normal, float *logits,
                                     float *scale, unsigned char *bins, const float *u01, unsigned long long seed, void *ws,
                                     int64_t ws_bytes, void *stream) {
    const State *st = static_cast<const State *>(state);
    if (!st) return CPPF_ERR_UNSUPPORTED;
    if (!logits && !bins) return CPPF_ERR_INVALID_ARGUMENT;
    if (ws_bytes < cppf_heads_tc_workspace_bytes(state, T, n)) return CPPF_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    __nv_bfloat16 *point_feat = static_cast<__nv_bfloat16 *>(ws);
    {
        Args a{};
        a.rows = n;
        a.x = feat;
        a.x_ld = st->model.branch == 0 ? CPPF_SHOT_DIM : 1024;
        a.arity = st->model.arity;
        a.weights = st->d_point_w;
        a.out_bf16 = point_feat;
        a.prof = nullptr;
        chain_tc_kernel<false><<<blocks_for_pairs(pairs_of(n)), kThreads, kSmemTotal, s>>>(st->point_prog, a, nullptr);
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
        a.out0 = logits;
        a.out1 = scale;
        a.bins = bins;
        a.u01 = u01;
        a.seed = seed;
        a.prof = g_tc_prof;
        if (a.prof) chain_tc_kernel<true><<<blocks_for_pairs(pairs_of(T)), kThreads, kSmemTotal, s>>>(st->tuple_prog, a, nullptr);
        else chain_tc_kernel<false><<<blocks_for_pairs(pairs_of(T)), kThreads, kSmemTotal, s>>>(st->tuple_prog, a, nullptr);
        CPPF_LAUNCH_CHECK();
    }
    return CPPF_OK;
}

// ---- batched frame path (frame.cuh) ------------------------------------------------------------------------------------
// All instances' per-point programs of a branch in one launch, all their per-tuple programs in another: the program (layer
// inventory, slab list) is the branch's architecture and identical for every category; rows, inputs, outputs and the weight
// stream come from the frame table.  126 point tiles fill the GPU where six 33-CTA launches ran at a tenth of its rate, and
// the tuple tiles of all jobs form full rounds (6 x 391 tiles = 1176 pairs = 8 rounds of 147 CTAs).
extern "C" int cppf_heads_tc_fill_job(const void *state, int kind, const float *pc, int64_t n, const void *idx, int idx_is_i64,
                                      int64_t idx_stride, int64_t T, const float *feat, const float *normal, void *point_feat,
                                      float *scale, unsigned char *bins, unsigned long long seed, void *args_out) {
    const State *st = static_cast<const State *>(state);
    if (!st || !args_out) return CPPF_ERR_INVALID_ARGUMENT;
    Args a{};
    a.arity = st->model.arity;
    if (kind == 0) {
        a.rows = n;
        a.x = feat;
        a.x_ld = st->model.branch == 0 ? CPPF_SHOT_DIM : 1024;
        a.weights = st->d_point_w;
        a.out_bf16 = static_cast<__nv_bfloat16 *>(point_feat);
    } else {
        a.rows = T;
        a.pc = pc;
        a.normal = normal;
        a.point_feat = static_cast<const __nv_bfloat16 *>(point_feat);
        a.idx = IdxView{idx, idx_stride, idx_is_i64};
        a.weights = st->d_tuple_w;
        a.out1 = scale;
        a.bins = bins;
        a.seed = seed;
    }
    *static_cast<Args *>(args_out) = a;
    return CPPF_OK;
}

extern "C" int64_t cppf_heads_tc_point_bytes(const void *state, int64_t n) {
    const State *st = static_cast<const State *>(state);
    return st ? static_cast<int64_t>(tc_align(2 * static_cast<size_t>(st->point_cols) * n)) : 0;
}

namespace cppf {

// tc_states[branch] = any model of the branch (its two programs); rows_cap[kind] bounds the rows of a job, ni the jobs
int frame_launch_heads(const FrameTable *t, const void *const *tc_states, int ni, int64_t n_cap, int64_t T_cap, cudaStream_t s) {
    if (ni <= 0) return CPPF_OK;
    for (int kind = 0; kind < 2; ++kind)
        for (int branch = 1; branch >= 0; --branch) {        // DINO first: it does not wait for the SHOT descriptors
            const State *st = static_cast<const State *>(tc_states[branch]);
            if (!st) continue;
            const int64_t pairs = pairs_of(kind == 0 ? n_cap : T_cap) * ni;
            Args none{};
            CPPF_CUDA_TRY(launch_frame_kernel(chain_tc_kernel<false>, dim3(blocks_for_pairs(pairs)), dim3(kThreads), kSmemTotal, s,
                                              kind == 0 ? st->point_prog : st->tuple_prog, none, &t->heads[kind][branch]));
        }
    return CPPF_OK;
}

}  // namespace cppf
