// Centre Hough voting on B200 -- replaces vote_center (reference train_dino.py:171-215).
//
// Layout in HBM: cloud f32 [N,3] (replicated per GPU, ~50 KB), tuple indices int64/int32 [T,stride],
// vote targets f32 [T,2], grid uint32 [gx*gy*gz] (0.2-4 MB: always L2 resident on a 126 MB L2).
//
// Kernel shape: one warp owns CHUNK tuples at a time.  Lanes < CHUNK build the pair frame of one
// tuple each (c, x*odist, y -- ~100 flops incl. 2 sqrt and 6 IEEE divisions), then the warp walks
// the CHUNK tuples and its 32 lanes take the R=180 rotations of one tuple 32 at a time, so the
// cos/sin table reads are conflict-free shared-memory loads and every lane issues one fire-and-forget
// RED.ADD.U32 to L2 per vote.  The geometry (lo, grid_res) is read from the device-resident
// cppf_grid_geom, so no host round trip separates bounds -> vote -> arg-max.
//
// Arithmetic: every float op is an explicit _rn intrinsic following SURVEY.md Appendix B (torch-CPU
// semantics: FMA chain in norm, FMA in cross, true division by float32(res), trunc(x+0.5)).  The
// integer grid is bit-exact against the oracle and the reference golden vectors.
#include "common.cuh"

namespace cppf {

constexpr int kMaxReplicas = 32;

// Number of grid copies the vote is spread over (MODE 0): as many as fit the caller's buffer, at most
// `replicas_max`.  Computed on the device from the device-resident geometry, identically by the zero,
// vote and fold kernels.
__device__ __forceinline__ int replica_count(int64_t cells, int64_t capacity, int replicas_max) {
    if (cells <= 0 || replicas_max <= 1) return 1;
    const int64_t fit = capacity / cells;
    return static_cast<int>(fit < 1 ? 1 : (fit > replicas_max ? replicas_max : fit));
}

// ---------------------------------------------------------------------------------------------------
// bounds: single CTA, coalesced sweep, shared-memory tree reduction.  N <= 50 000 by eval.py:195.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) cloud_bounds_kernel(const float *__restrict__ pc, int64_t n, float res,
                                                            cppf_grid_geom *__restrict__ geom) {
    cloud_bounds_shared(pc, n, res, geom);
}

// ---------------------------------------------------------------------------------------------------
// zero the live part of the grid (size known only on the device)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_zero_body(uint32_t *__restrict__ grid, int64_t capacity,
                                               const cppf_grid_geom *__restrict__ geom, uint32_t *__restrict__ status,
                                               int replicas_max, int bid, int nblk) {
    int64_t cells = geom->cells;
    if (cells <= capacity) cells *= replica_count(cells, capacity, replicas_max);
    if (cells > capacity || geom->cells > 0x7fffffffll) {     // beyond the caller's buffer, or beyond the 32-bit cell index of the vote
        if (bid == 0 && threadIdx.x == 0 && status) atomicOr(status, CPPF_STATUS_GRID_OVERFLOW);
        cells = cells > capacity ? capacity : cells;
    }
    if (bid == 0 && threadIdx.x == 0 && status && geom->flags) atomicOr(status, geom->flags);
    int64_t vec = cells >> 2;
    uint4 *g4 = reinterpret_cast<uint4 *>(grid);
    int64_t stride = static_cast<int64_t>(nblk) * blockDim.x;
    int64_t tid = static_cast<int64_t>(bid) * blockDim.x + threadIdx.x;
    for (int64_t i = tid; i < vec; i += stride) g4[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t i = (vec << 2) + tid; i < cells; i += stride) grid[i] = 0u;
}

__global__ void __launch_bounds__(256) grid_zero_kernel(uint32_t *__restrict__ grid, int64_t capacity,
                                                        const cppf_grid_geom *__restrict__ geom,
                                                        uint32_t *__restrict__ status, int replicas_max) {
    grid_zero_body(grid, capacity, geom, status, replicas_max, blockIdx.x, gridDim.x);
}

// ---------------------------------------------------------------------------------------------------
// the vote kernel
// ---------------------------------------------------------------------------------------------------
constexpr int kVoteThreads = 256;
constexpr int kMaxRotsSmem = 1024;

// MODE 0: RED straight to the L2-resident grid; `replicas` copies of the grid (copy = block % replicas)
//         spread the hot cache lines of the vote peak over several L2 slices.
// MODE 1: the whole grid lives in this CTA's shared memory (cells*4 B <= ~220 KB), flushed at the end.
// NITER > 0: the rotation loop is unrolled for exactly NITER warp passes (R = 180 -> 6; no trip-count arithmetic between the
// passes); 0: any R.
template <int CHUNK, int MODE, int NITER = 0>
__device__ __forceinline__ void vote_center_body(
    const float *__restrict__ pc, const IdxView &idx, const float *__restrict__ preds_tr, int64_t T,
    const float *__restrict__ cos_tab, const float *__restrict__ sin_tab, int R,
    const cppf_grid_geom *__restrict__ geom, uint32_t *__restrict__ grid, int64_t capacity, int replicas_max,
    int64_t smem_cells, uint32_t *__restrict__ status, int bid, int nblk, float *__restrict__ s_cos, float *__restrict__ s_sin) {
    extern __shared__ __align__(16) uint32_t s_grid[];
    // the tables are padded to whole warps with NaN: a padding lane's quotient is NaN, its cell converts to 0, and cell 0 never
    // receives votes -- no `r < R` predicate (and no branch around the loads) in the vote loop
    const int R_pad = (R + 31) & ~31;
    for (int r = threadIdx.x; r < R_pad; r += blockDim.x) {
        s_cos[r] = r < R ? cos_tab[r] : __int_as_float(0x7fc00000);
        s_sin[r] = r < R ? sin_tab[r] : __int_as_float(0x7fc00000);
    }
    const int64_t cells = geom->cells;
    if (MODE == 1) {
        if (cells > smem_cells) {  // the caller's bound on the grid size was wrong: flag, never corrupt
            if (bid == 0 && threadIdx.x == 0 && status) atomicOr(status, CPPF_STATUS_GRID_OVERFLOW);
            return;
        }
        for (int64_t i = threadIdx.x; i < cells; i += blockDim.x) s_grid[i] = 0u;
    }
    __syncthreads();

    if (cells > capacity || cells > 0x7fffffffll) return;  // flagged by grid_zero_kernel
    if (MODE == 0) grid += (bid % replica_count(cells, capacity, replicas_max)) * cells;
    uint32_t *vote_base = MODE == 1 ? s_grid : grid;
    if (MODE == 0) asm volatile("" : "+l"(vote_base));      // one 64-bit base register: the vote's address is a single IMAD.WIDE
    const float res = geom->res;
    const float inv_res = __frcp_rn(res);      // correctly rounded reciprocal of the launch-wide divisor (see div_by)
    const float lo0 = geom->lo[0], lo1 = geom->lo[1], lo2 = geom->lo[2];
    const int g0 = static_cast<int>(geom->grid_res[0]), g1 = static_cast<int>(geom->grid_res[1]),
              g2 = static_cast<int>(geom->grid_res[2]);

    const int lane = lane_id();
    const float *cos_l = s_cos + lane, *sin_l = s_sin + lane;
    const uint32_t right_of_lane = ~((2u << lane) - 1u);       // the lanes to the right of this one
    const int64_t warp = (static_cast<int64_t>(bid) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(nblk) * blockDim.x) >> 5;

    for (int64_t base = warp * CHUNK; base < T; base += n_warps * CHUNK) {
        // ---- per-tuple frame, one tuple per lane (train_dino.py:176-192) -------------------------
        float c[3] = {0.f, 0.f, 0.f}, x[3] = {0.f, 0.f, 0.f}, y[3] = {0.f, 0.f, 0.f};
        bool ok = false;
        const int64_t t = base + lane;
        if (lane < CHUNK && t < T) {
            const float2 tr = reinterpret_cast<const float2 *>(preds_tr)[t];  // (proj_len, odist)
            const int64_t ia = idx.at(t, 0), ib = idx.at(t, 1);
            float a[3] = {pc[3 * ia], pc[3 * ia + 1], pc[3 * ia + 2]};
            float b[3] = {pc[3 * ib], pc[3 * ib + 1], pc[3 * ib + 2]};
            float ab[3];
            ok = (tr.y > res) && pair_frame(a, b, false, ab, x);  // :182
            if (ok) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    c[k] = __fsub_rn(a[k], __fmul_rn(ab[k], tr.x));  // :186
                    x[k] = __fmul_rn(x[k], tr.y);                    // :191
                }
                cross_torch(x, ab, y);  // :192
            }
        }
        const uint32_t ok_mask = __ballot_sync(0xffffffffu, ok);

        // ---- votes: lanes sweep the rotations of one tuple at a time -----------------------------
#pragma unroll 1
        for (int j = 0; j < CHUNK; ++j) {
            if (!((ok_mask >> j) & 1u)) continue;  // warp-uniform
            const float c0 = __shfl_sync(0xffffffffu, c[0], j), c1 = __shfl_sync(0xffffffffu, c[1], j),
                        c2 = __shfl_sync(0xffffffffu, c[2], j);
            const float x0 = __shfl_sync(0xffffffffu, x[0], j), x1 = __shfl_sync(0xffffffffu, x[1], j),
                        x2 = __shfl_sync(0xffffffffu, x[2], j);
            const float y0 = __shfl_sync(0xffffffffu, y[0], j), y1 = __shfl_sync(0xffffffffu, y[1], j),
                        y2 = __shfl_sync(0xffffffffu, y[2], j);
            auto vote32 = [&](int r0) {                   // warp-uniform: the run detection below is a full-warp operation
                const float cr = cos_l[r0], sr = sin_l[r0];
                // offset = cos*x + sin*y (mul, mul, add); g = ((c + offset) - lo) / res; cell = trunc(g + 0.5)
                const float o0 = __fadd_rn(__fmul_rn(cr, x0), __fmul_rn(sr, y0));
                const float o1 = __fadd_rn(__fmul_rn(cr, x1), __fmul_rn(sr, y1));
                const float o2 = __fadd_rn(__fmul_rn(cr, x2), __fmul_rn(sr, y2));
                const float q0 = div_by(__fsub_rn(__fadd_rn(c0, o0), lo0), res, inv_res);
                const float q1 = div_by(__fsub_rn(__fadd_rn(c1, o1), lo1), res, inv_res);
                const float q2 = div_by(__fsub_rn(__fadd_rn(c2, o2), lo2), res, inv_res);
                const int i0 = __float2int_rz(__fadd_rn(q0, 0.5f));
                const int i1 = __float2int_rz(__fadd_rn(q1, 0.5f));
                const int i2 = __float2int_rz(__fadd_rn(q2, 0.5f));
                // strictly inside (cell 0 never receives votes, :200); a padding lane's NaN converts to 0
                const bool valid = i0 > 0 && i1 > 0 && i2 > 0 && i0 < g0 && i1 < g1 && i2 < g2;
                // 32-bit cell index (grids beyond 2^31 cells are never voted, see above): the 64-bit form was 13 of the
                // ~64 instructions of a vote
                const int lin = valid ? (i0 * g1 + i1) * g2 + i2 : -1 - lane;
                // Consecutive rotations of a tuple are consecutive points of its circle: on the bench's grids (circle radius
                // of tens of cells, 180 votes) a cell receives ~2 consecutive votes on average, up to 5 for the 1 cm laptop
                // grid.  Lanes with the cell of their left neighbour stay silent and the first lane of each run adds the
                // run length: half the reductions reach L2, whose reduction rate bounds this kernel.  Integer sums: the grid
                // is unchanged.
                const int left = __shfl_up_sync(0xffffffffu, lin, 1);
                const bool head = lane == 0 || lin != left;
                const uint32_t heads = __ballot_sync(0xffffffffu, head);
                // distance to the next run head on the right = trailing zeros of the heads to the right, popc(~h & (h - 1)); with
                // no head there the formula gives 32, i.e. the run reaches the end of the warp
                const uint32_t hr = heads & right_of_lane;
                const uint32_t run = static_cast<uint32_t>(__popc(~hr & (hr - 1u)) - lane);
                if (valid && head) atomicAdd(vote_base + lin, run);      // ATOMS (MODE 1) / a reduction at L2 (result unused)
            };
            if (NITER > 0) {
#pragma unroll
                for (int k = 0; k < NITER; ++k) vote32(32 * k);
            } else {
#pragma unroll 2
                for (int r0 = 0; r0 < R_pad; r0 += 32) vote32(r0);
            }
        }
    }
    if (MODE == 1) {
        __syncthreads();
        // staggered flush: CTA b starts at its own offset so that the CTAs do not sweep the lines in lock step
        const int64_t start = (cells / nblk) * bid;
        for (int64_t i = threadIdx.x; i < cells; i += blockDim.x) {
            int64_t j = i + start;
            j = j >= cells ? j - cells : j;
            const uint32_t v = s_grid[j];
            if (v) atomicAdd(grid + j, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Lane-per-tuple form of the same vote (round 2): every lane owns ONE tuple and walks its circle rotation by rotation.
// cos/sin are warp-uniform shared-memory broadcasts, the pair frame never crosses lanes (no shuffles, no ballot), all 32
// lanes build frames (the warp-per-tuple form keeps 8 of 32 busy there), and the run-length aggregation is temporal:
// consecutive rotations that fall into one cell are counted in a register and leave the SM as one reduction, runs no
// longer cut at 32-rotation boundaries.  Same float operations per vote => the same integer grid.
// ---------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void vote_center_lanes_body(
    const float *__restrict__ pc, const IdxView &idx, const float *__restrict__ preds_tr, int64_t T,
    const float *__restrict__ cos_tab, const float *__restrict__ sin_tab, int R,
    const cppf_grid_geom *__restrict__ geom, uint32_t *__restrict__ grid, int64_t capacity, int replicas_max,
    int64_t smem_cells, uint32_t *__restrict__ status, int bid, int nblk, float2 *__restrict__ s_cs) {
    extern __shared__ __align__(16) uint32_t s_grid[];
    for (int r = threadIdx.x; r < R; r += blockDim.x) s_cs[r] = make_float2(cos_tab[r], sin_tab[r]);     // one LDS.64 per rotation
    const int64_t cells = geom->cells;
    if (MODE == 1) {
        if (cells > smem_cells) {  // the caller's bound on the grid size was wrong: flag, never corrupt
            if (bid == 0 && threadIdx.x == 0 && status) atomicOr(status, CPPF_STATUS_GRID_OVERFLOW);
            return;
        }
        for (int64_t i = threadIdx.x; i < cells; i += blockDim.x) s_grid[i] = 0u;
    }
    __syncthreads();

    if (cells > capacity || cells > 0x7fffffffll) return;  // flagged by grid_zero_kernel
    if (MODE == 0) grid += (bid % replica_count(cells, capacity, replicas_max)) * cells;
    uint32_t *vote_base = MODE == 1 ? s_grid : grid;
    if (MODE == 0) asm volatile("" : "+l"(vote_base));
    const float res = geom->res;
    const float inv_res = __frcp_rn(res);
    const float lo0 = geom->lo[0], lo1 = geom->lo[1], lo2 = geom->lo[2];
    const int g0 = static_cast<int>(geom->grid_res[0]), g1 = static_cast<int>(geom->grid_res[1]),
              g2 = static_cast<int>(geom->grid_res[2]);
    const int64_t first = static_cast<int64_t>(bid) * blockDim.x + threadIdx.x;
    const int64_t stride = static_cast<int64_t>(nblk) * blockDim.x;
    for (int64_t t = first; t - lane_id() < T; t += stride) {       // warp-uniform trip count
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, x0 = 0.f, x1 = 0.f, x2 = 0.f, y0 = 0.f, y1 = 0.f, y2 = 0.f;
        bool ok = false;
        if (t < T) {
            const float2 tr = reinterpret_cast<const float2 *>(preds_tr)[t];  // (proj_len, odist)
            const int64_t ia = idx.at(t, 0), ib = idx.at(t, 1);
            float a[3] = {pc[3 * ia], pc[3 * ia + 1], pc[3 * ia + 2]};
            float b[3] = {pc[3 * ib], pc[3 * ib + 1], pc[3 * ib + 2]};
            float ab[3], x[3], y[3];
            ok = (tr.y > res) && pair_frame(a, b, false, ab, x);  // :182
            if (ok) {
#pragma unroll
                for (int k = 0; k < 3; ++k) x[k] = __fmul_rn(x[k], tr.y);      // :191
                cross_torch(x, ab, y);                                          // :192
                c0 = __fsub_rn(a[0], __fmul_rn(ab[0], tr.x));                   // :186
                c1 = __fsub_rn(a[1], __fmul_rn(ab[1], tr.x));
                c2 = __fsub_rn(a[2], __fmul_rn(ab[2], tr.x));
                x0 = x[0], x1 = x[1], x2 = x[2];
                y0 = y[0], y1 = y[1], y2 = y[2];
            }
        }
        if (!__any_sync(0xffffffffu, ok)) continue;
        if (!ok) c0 = __int_as_float(0x7fc00000);      // masked pair (:182) or past the end: every quotient NaN -> cell 0 -> no vote
        int prev = -1;              // cell of the current run (-1: none), cnt its length so far
        uint32_t cnt = 0u;
#pragma unroll 4
        for (int r = 0; r < R; ++r) {
            const float2 cs = s_cs[r];                       // one address per warp: broadcast
            const float cr = cs.x, sr = cs.y;
            const float o0 = __fadd_rn(__fmul_rn(cr, x0), __fmul_rn(sr, y0));
            const float o1 = __fadd_rn(__fmul_rn(cr, x1), __fmul_rn(sr, y1));
            const float o2 = __fadd_rn(__fmul_rn(cr, x2), __fmul_rn(sr, y2));
            const float q0 = div_by(__fsub_rn(__fadd_rn(c0, o0), lo0), res, inv_res);
            const float q1 = div_by(__fsub_rn(__fadd_rn(c1, o1), lo1), res, inv_res);
            const float q2 = div_by(__fsub_rn(__fadd_rn(c2, o2), lo2), res, inv_res);
            const int i0 = __float2int_rz(__fadd_rn(q0, 0.5f));
            const int i1 = __float2int_rz(__fadd_rn(q1, 0.5f));
            const int i2 = __float2int_rz(__fadd_rn(q2, 0.5f));
            // strictly inside (cell 0 excluded, :200); a masked tuple's NaN centre converts to cell 0
            // (a predicate chain spelled out: the compiler's own choice for this expression was a chain of four SELs)
            int lin;
            asm("{\n\t.reg .pred p;\n\t"
                "setp.gt.s32 p, %1, 0;\n\t"
                "setp.lt.and.s32 p, %2, %3, p;\n\t"
                "setp.lt.and.s32 p, %4, %5, p;\n\t"
                "setp.lt.and.s32 p, %6, %7, p;\n\t"
                "selp.s32 %0, %8, -1, p;\n\t}"
                : "=r"(lin)
                : "r"(min(i0, min(i1, i2))), "r"(i0), "r"(g0), "r"(i1), "r"(g1), "r"(i2), "r"(g2), "r"((i0 * g1 + i1) * g2 + i2));
            const bool change = lin != prev;
            if (change && prev >= 0) atomicAdd(vote_base + prev, cnt);      // the finished run: one reduction (result unused)
            cnt = change ? 1u : cnt + 1u;
            prev = lin;
        }
        if (prev >= 0) atomicAdd(vote_base + prev, cnt);
    }
    if (MODE == 1) {
        __syncthreads();
        const int64_t start = (cells / nblk) * bid;
        for (int64_t i = threadIdx.x; i < cells; i += blockDim.x) {
            int64_t j = i + start;
            j = j >= cells ? j - cells : j;
            const uint32_t v = s_grid[j];
            if (v) atomicAdd(grid + j, v);
        }
    }
}

template <int CHUNK, int MODE, int THREADS, int NITER>
__global__ void __launch_bounds__(THREADS) vote_center_kernel(
    const float *__restrict__ pc, IdxView idx, const float *__restrict__ preds_tr, int64_t T,
    const float *__restrict__ cos_tab, const float *__restrict__ sin_tab, int R,
    const cppf_grid_geom *__restrict__ geom, uint32_t *__restrict__ grid, int64_t capacity, int replicas_max,
    int64_t smem_cells, uint32_t *__restrict__ status) {
    __shared__ float s_cos[kMaxRotsSmem], s_sin[kMaxRotsSmem];
    vote_center_body<CHUNK, MODE, NITER>(pc, idx, preds_tr, T, cos_tab, sin_tab, R, geom, grid, capacity, replicas_max, smem_cells,
                                         status, blockIdx.x, gridDim.x, s_cos, s_sin);
}

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) vote_center_lanes_kernel(
    const float *__restrict__ pc, IdxView idx, const float *__restrict__ preds_tr, int64_t T,
    const float *__restrict__ cos_tab, const float *__restrict__ sin_tab, int R,
    const cppf_grid_geom *__restrict__ geom, uint32_t *__restrict__ grid, int64_t capacity, int replicas_max,
    int64_t smem_cells, uint32_t *__restrict__ status) {
    __shared__ float2 s_cs[kMaxRotsSmem];
    vote_center_lanes_body<MODE>(pc, idx, preds_tr, T, cos_tab, sin_tab, R, geom, grid, capacity, replicas_max, smem_cells, status,
                                 blockIdx.x, gridDim.x, s_cs);
}

// CPPF_VOTE_LANES=0 selects the warp-per-tuple form (A/B and fallback)
static bool vote_lanes_enabled() {
    static const bool on = [] {
        const char *e = getenv("CPPF_VOTE_LANES");
        return !(e && e[0] == '0');
    }();
    return on;
}

// sums replicas 1..K-1 into replica 0
__global__ void __launch_bounds__(256) grid_fold_kernel(uint32_t *__restrict__ grid, const cppf_grid_geom *__restrict__ geom,
                                                        int64_t capacity, int replicas_max) {
    const int64_t cells = geom->cells;
    if (cells > capacity) return;
    const int replicas = replica_count(cells, capacity, replicas_max);
    if (replicas <= 1) return;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < cells; i += stride) {
        uint32_t acc = grid[i];
        for (int k = 1; k < replicas; ++k) acc += grid[k * cells + i];
        grid[i] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------
// arg-max: first maximum in C order.  key = count<<32 | ~index so that atomicMax picks the largest
// count and, among equals, the smallest index.  The last block to finish converts to world space.
// ---------------------------------------------------------------------------------------------------
struct CenterScratch {  // lives in the tail of cppf_center (pad fields), zeroed by cudaMemsetAsync
    unsigned long long key;
    unsigned int ticket;
};

__global__ void __launch_bounds__(256) grid_argmax_kernel(const uint32_t *__restrict__ grid, int64_t capacity,
                                                          const cppf_grid_geom *__restrict__ geom, double res,
                                                          const uint32_t *__restrict__ status,
                                                          cppf_center *__restrict__ out,
                                                          unsigned long long *__restrict__ key,
                                                          unsigned int *__restrict__ ticket) {
    // a grid larger than the caller's buffer was never voted (grid_zero_kernel raised CPPF_STATUS_GRID_OVERFLOW): nothing to
    // scan, and nothing may be read past the buffer
    const bool overflow = geom->cells > capacity || geom->cells > 0x7fffffffll;
    const int64_t cells = overflow ? 0 : geom->cells;
    unsigned long long best = 0ull;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < cells; i += stride) {
        unsigned long long k = (static_cast<unsigned long long>(grid[i]) << 32) |
                               static_cast<unsigned long long>(0xffffffffu - static_cast<uint32_t>(i));
        best = k > best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    __shared__ unsigned long long s_best[8];
    __shared__ bool s_last;
    if (lane_id() == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) best = s_best[w] > best ? s_best[w] : best;
        atomicMax(key, best);
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(key);
        int64_t lin = static_cast<int64_t>(0xffffffffu - static_cast<uint32_t>(k & 0xffffffffull));
        if (cells <= 0) lin = 0;
        const int64_t g1 = geom->grid_res[1], g2 = geom->grid_res[2];
        int64_t cell[3] = {lin / (g1 * g2), (lin / g2) % g1, lin % g2};
        for (int a = 0; a < 3; ++a) {
            out->cell[a] = cell[a];
            // corners[0].numpy() (f32) + cand (int64) * res (python float) -> float64 (train_dino.py:213)
            out->world[a] = __dadd_rn(static_cast<double>(geom->lo[a]), __dmul_rn(static_cast<double>(cell[a]), res));
        }
        out->linear = lin;
        out->votes = static_cast<uint32_t>(k >> 32);
        out->cells = geom->cells;
        // the stage's status travels with the centre into the pose record (pose_directions_kernel): geometry flags
        // (extent guard, empty cloud), overflow of the caller's grid buffer, and whatever the vote kernels raised
        out->status = geom->flags | (overflow ? CPPF_STATUS_GRID_OVERFLOW : 0u) | (status ? *status : 0u);
    }
}

__global__ void __launch_bounds__(256) grid_widen_kernel(const uint32_t *__restrict__ grid, int64_t capacity,
                                                         const cppf_grid_geom *__restrict__ geom,
                                                         int64_t *__restrict__ out) {
    const int64_t cells = geom->cells < capacity ? geom->cells : capacity;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < cells; i += stride)
        out[i] = static_cast<int64_t>(grid[i]);
}

}  // namespace cppf

using namespace cppf;

CPPF_API int cppf_cloud_bounds(const float *pc, int64_t n, float res, cppf_grid_geom *geom, void *stream) {
    if (!pc || !geom || n < 0 || !(res > 0.0f)) return CPPF_ERR_INVALID_ARGUMENT;
    cloud_bounds_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(pc, n, res, geom);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

// mode: 0 = RED to L2 with up to `replicas_max` grid copies, 1 = grid privatised in shared memory
// (smem_cells = caller's upper bound of geom->cells).  Exported for the tuning tool as well.
CPPF_API int cppf_vote_center_ex(const float *pc, int64_t n, const void *idx, int idx_is_i64, int64_t idx_stride,
                                 const float *preds_tr, int64_t T, const float *cos_tab, const float *sin_tab, int R,
                                 const cppf_grid_geom *geom, uint32_t *grid, int64_t grid_capacity, int accumulate,
                                 uint32_t *status, int mode, int replicas_max, int64_t smem_cells, void *stream) {
    if (!pc || !cos_tab || !sin_tab || !geom || !grid) return CPPF_ERR_INVALID_ARGUMENT;
    if (T > 0 && (!preds_tr || !idx)) return CPPF_ERR_INVALID_ARGUMENT;
    if (n <= 0 || T < 0 || R <= 0 || R > kMaxRotsSmem || idx_stride < 2 || grid_capacity <= 0 || replicas_max < 0)
        return CPPF_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const DeviceInfo &dev = device_info();
    if (replicas_max > kMaxReplicas) replicas_max = kMaxReplicas;
    if (accumulate || mode == 1 || replicas_max < 1) replicas_max = 1;  // partial grids must add into one copy
    if (!accumulate) {
        int zb = grid_for(grid_capacity / 4 + 1, 256, 8);
        grid_zero_kernel<<<zb, 256, 0, s>>>(grid, grid_capacity, geom, status, replicas_max);
        CPPF_LAUNCH_CHECK();
    }
    if (T == 0) return CPPF_OK;
    IdxView iv{idx, idx_stride, idx_is_i64};
    const int warps_per_block = kVoteThreads / 32;
    const bool six = (R + 31) / 32 == 6;        // the reference's num_rots = 180: the unrolled instantiation
    if (mode == 1) {
        const size_t smem = static_cast<size_t>(smem_cells) * sizeof(uint32_t);
        const int smem_max = dev.max_smem_optin - 2 * kMaxRotsSmem * static_cast<int>(sizeof(float)) - 1024;
        if (smem_cells <= 0 || smem > static_cast<size_t>(smem_max)) return CPPF_ERR_UNSUPPORTED;
        // the opt-in is per device and must be taken once, whichever host thread gets here first
        CPPF_TRY_ONCE_PER_DEVICE(cudaFuncSetAttribute(vote_center_kernel<8, 1, 1024, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      smem_max));
        CPPF_TRY_ONCE_PER_DEVICE(cudaFuncSetAttribute(vote_center_kernel<8, 1, 1024, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      smem_max));
        int64_t warps = (T + 7) / 8;
        int64_t blocks = (warps + 31) / 32;
        const int64_t per_sm = smem > 100 * 1024 ? 1 : 2;
        if (blocks > dev.sm_count * per_sm) blocks = dev.sm_count * per_sm;
        if (vote_lanes_enabled()) {
            CPPF_TRY_ONCE_PER_DEVICE(cudaFuncSetAttribute(vote_center_lanes_kernel<1, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
            vote_center_lanes_kernel<1, 1024><<<static_cast<int>(std::min<int64_t>(blocks, (T + 1023) / 1024)), 1024, smem, s>>>(
                pc, iv, preds_tr, T, cos_tab, sin_tab, R, geom, grid, grid_capacity, 1, smem_cells, status);
        } else if (six)
            vote_center_kernel<8, 1, 1024, 6><<<static_cast<int>(blocks), 1024, smem, s>>>(
                pc, iv, preds_tr, T, cos_tab, sin_tab, R, geom, grid, grid_capacity, 1, smem_cells, status);
        else
            vote_center_kernel<8, 1, 1024, 0><<<static_cast<int>(blocks), 1024, smem, s>>>(
                pc, iv, preds_tr, T, cos_tab, sin_tab, R, geom, grid, grid_capacity, 1, smem_cells, status);
        CPPF_LAUNCH_CHECK();
        return CPPF_OK;
    }
    // Tuples per warp pass: small chunks when T is small so that every SM still gets >= ~32 warps.
    const int64_t warps_full = static_cast<int64_t>(dev.sm_count) * 8 * warps_per_block;
    // Votes straight to L2: the lane-per-tuple form wins at a frame's sizes (12 jobs x 50 000 tuples: 0.317 against 0.336 ms)
    // but loses at T = 2^22 on the 0.8 M-cell example grid (3.20 against 2.80 ms): a warp's 32 votes then land on 32 different
    // circles instead of 32 neighbouring points of one, and L2 sees more distinct sectors per request.
    if (vote_lanes_enabled() && T < (1 << 18)) {
        const int64_t blocks = std::min<int64_t>((T + kVoteThreads - 1) / kVoteThreads, static_cast<int64_t>(dev.sm_count) * 8);
        vote_center_lanes_kernel<0, kVoteThreads><<<static_cast<int>(blocks), kVoteThreads, 0, s>>>(
            pc, iv, preds_tr, T, cos_tab, sin_tab, R, geom, grid, grid_capacity, replicas_max, 0, status);
    } else if (T >= warps_full * 32) {
        int blocks = dev.sm_count * 8;
        if (six)
            vote_center_kernel<32, 0, kVoteThreads, 6><<<blocks, kVoteThreads, 0, s>>>(pc, iv, preds_tr, T, cos_tab, sin_tab, R, geom,
                                                                                       grid, grid_capacity, replicas_max, 0, status);
        else
            vote_center_kernel<32, 0, kVoteThreads, 0><<<blocks, kVoteThreads, 0, s>>>(pc, iv, preds_tr, T, cos_tab, sin_tab, R, geom,
                                                                                       grid, grid_capacity, replicas_max, 0, status);
    } else {
        int64_t warps = (T + 7) / 8;
        int64_t blocks = (warps + warps_per_block - 1) / warps_per_block;
        int64_t cap = static_cast<int64_t>(dev.sm_count) * 8;
        if (blocks > cap) blocks = cap;
        if (six)
            vote_center_kernel<8, 0, kVoteThreads, 6><<<static_cast<int>(blocks), kVoteThreads, 0, s>>>(
                pc, iv, preds_tr, T, cos_tab, sin_tab, R, geom, grid, grid_capacity, replicas_max, 0, status);
        else
            vote_center_kernel<8, 0, kVoteThreads, 0><<<static_cast<int>(blocks), kVoteThreads, 0, s>>>(
                pc, iv, preds_tr, T, cos_tab, sin_tab, R, geom, grid, grid_capacity, replicas_max, 0, status);
    }
    CPPF_LAUNCH_CHECK();
    if (replicas_max > 1) {
        grid_fold_kernel<<<dev.sm_count * 4, 256, 0, s>>>(grid, geom, grid_capacity, replicas_max);
        CPPF_LAUNCH_CHECK();
    }
    return CPPF_OK;
}

// Shared-memory budget of the privatised mode, in cells.
CPPF_API int64_t cppf_vote_center_smem_cells(void) {
    const DeviceInfo &dev = device_info();
    return (dev.max_smem_optin - 2 * kMaxRotsSmem * static_cast<int64_t>(sizeof(float)) - 1024) / 4;
}

CPPF_API int cppf_vote_center(const float *pc, int64_t n, const void *idx, int idx_is_i64, int64_t idx_stride,
                              const float *preds_tr, int64_t T, const float *cos_tab, const float *sin_tab, int R,
                              const cppf_grid_geom *geom, uint32_t *grid, int64_t grid_capacity, int64_t cells_hint,
                              int accumulate, uint32_t *status, void *stream) {
    // cells_hint > 0: the caller knows gx*gy*gz (it built the cloud on the host).  Grids that fit one SM's
    // shared memory are privatised there; everything else votes with RED into up to 32 L2-resident copies
    // of the grid (as many as grid_capacity holds), which spreads the hot lines of the vote peak.
    const bool smem_ok = cells_hint > 0 && cells_hint <= cppf_vote_center_smem_cells() && T >= 16384;
    return cppf_vote_center_ex(pc, n, idx, idx_is_i64, idx_stride, preds_tr, T, cos_tab, sin_tab, R, geom, grid,
                               grid_capacity, accumulate, status, smem_ok ? 1 : 0, kMaxReplicas, cells_hint, stream);
}

CPPF_API int cppf_grid_argmax(const uint32_t *grid, int64_t grid_capacity, const cppf_grid_geom *geom, double res,
                              const uint32_t *status, cppf_center *center, void *stream) {
    if (!grid || !geom || !center || grid_capacity <= 0) return CPPF_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CPPF_CUDA_TRY(cudaMemsetAsync(center, 0, sizeof(cppf_center), s));
    // scratch words live in the struct's tail (`votes` is written last by the finalising thread; `status`
    // doubles as the ticket, `linear` as the packed key until then)
    unsigned long long *key = reinterpret_cast<unsigned long long *>(&center->linear);
    unsigned int *ticket = reinterpret_cast<unsigned int *>(&center->status);
    int blocks = device_info().sm_count * 4;
    grid_argmax_kernel<<<blocks, 256, 0, s>>>(grid, grid_capacity, geom, res, status, center, key, ticket);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_grid_to_i64(const uint32_t *grid, int64_t grid_capacity, const cppf_grid_geom *geom, int64_t *grid_i64,
                              void *stream) {
    if (!grid || !geom || !grid_i64 || grid_capacity <= 0) return CPPF_ERR_INVALID_ARGUMENT;
    int blocks = device_info().sm_count * 4;
    grid_widen_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(grid, grid_capacity, geom, grid_i64);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

// =====================================================================================================================
// Batched frame path (frame.cuh): the centre-vote stage of EVERY (instance, branch) job of a frame in four launches.
// blockIdx.y selects the job; all sizes come from the device-resident table, so the launch dimensions do not depend on the
// frame's content.  Arithmetic = the bodies above, so grids are bit-identical to the single-job calls.
// =====================================================================================================================
#include "frame.cuh"
#include "targets.cuh"

namespace cppf {

// prep: bounds of the job's cloud -> geom, and every small accumulator of the chain zeroed (centre record incl. its key /
// ticket words, back-vote summary incl. kept and imp_max, status, sphere bins, importance counts, pose scratch)
__global__ void __launch_bounds__(1024) frame_prep_kernel(const FrameTable *__restrict__ t, FrameShared sh) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.x) >= t->n_jobs) return;
    const FrameJob &j = t->job[blockIdx.x];
    const FrameInst &in = t->inst[j.inst];
    cloud_bounds_shared(in.pc, in.n, j.res, j.geom);
    const int tid = threadIdx.x;
    uint32_t *c = reinterpret_cast<uint32_t *>(j.center);
    for (int i = tid; i < static_cast<int>(sizeof(cppf_center) / 4); i += 1024) c[i] = 0u;
    uint32_t *sm = reinterpret_cast<uint32_t *>(j.summary);
    for (int i = tid; i < static_cast<int>(sizeof(cppf_backvote_summary) / 4); i += 1024) sm[i] = 0u;
    if (tid == 0) *j.status = 0u;
    uint32_t *ws = reinterpret_cast<uint32_t *>(j.ws_pose);
    for (int i = tid; i < 64; i += 1024) ws[i] = 0u;                       // PoseScratch (first 256 bytes of ws_pose)
    for (int i = tid; i < 2 * sh.S; i += 1024) j.counts[i] = 0.0;
    for (int i = tid; i < in.n; i += 1024) j.imp[i] = 0;
}

// decode + targets of the job's tuples (eval.py:230-240), and the live part of its grid zeroed
__global__ void __launch_bounds__(256, 4) frame_decode_zero_kernel(const FrameTable *__restrict__ t, FrameShared sh) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.y) >= t->n_jobs) return;
    const FrameJob &j = t->job[blockIdx.y];
    const FrameInst &in = t->inst[j.inst];
    grid_zero_body(j.grid, j.grid_capacity, j.geom, j.status, sh.replicas_max, blockIdx.x, gridDim.x);
    Axes ax;
#pragma unroll
    for (int i = 0; i < 9; ++i) ax.v[i] = j.axes[i];
    decode_targets_body(in.pc, in.idx, j.bins, in.T, sh.num_bins, ax, j.targets_tr, j.targets_rot, nullptr, nullptr, blockIdx.x,
                        gridDim.x);
}

// Every job votes with RED into its L2-resident grid copies (the mode the per-job path picks for all but the smallest
// grids; at the frame's T <= 2^17 it is also the faster one for those: ncu 37 us against 48 us per job, and a launch whose
// CTAs each reserve a whole SM's shared memory would serialise the jobs).
__global__ void __launch_bounds__(kVoteThreads) frame_vote_center_kernel(const FrameTable *__restrict__ t, FrameShared sh) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.y) >= t->n_jobs) return;
    const FrameJob &j = t->job[blockIdx.y];
    const FrameInst &in = t->inst[j.inst];
    __shared__ float s_cos[kMaxRotsSmem], s_sin[kMaxRotsSmem];
    // as many CTAs as the single-job launch would use for this T (the grid is sized for the capacity)
    const int64_t warps = (in.T + 7) / 8;
    int64_t nblk = (warps + kVoteThreads / 32 - 1) / (kVoteThreads / 32);
    if (nblk > static_cast<int64_t>(gridDim.x)) nblk = gridDim.x;
    if (static_cast<int64_t>(blockIdx.x) >= nblk) return;
    if ((sh.R + 31) / 32 == 6)
        vote_center_body<8, 0, 6>(in.pc, in.idx, j.targets_tr, in.T, sh.cos_tab, sh.sin_tab, sh.R, j.geom, j.grid, j.grid_capacity,
                                  sh.replicas_max, 0, j.status, blockIdx.x, static_cast<int>(nblk), s_cos, s_sin);
    else
        vote_center_body<8, 0, 0>(in.pc, in.idx, j.targets_tr, in.T, sh.cos_tab, sh.sin_tab, sh.R, j.geom, j.grid, j.grid_capacity,
                                  sh.replicas_max, 0, j.status, blockIdx.x, static_cast<int>(nblk), s_cos, s_sin);
}

// lane-per-tuple form (vote_center_lanes_body): one CTA per 256 tuples of the job; 32 registers, all 64 warp slots of an SM
// (capping the registers at 32 for eight resident CTAs per SM instead of six measured slower: 0.322 against 0.317 ms per frame)
__global__ void __launch_bounds__(kVoteThreads) frame_vote_center_lanes_kernel(const FrameTable *__restrict__ t, FrameShared sh) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.y) >= t->n_jobs) return;
    const FrameJob &j = t->job[blockIdx.y];
    const FrameInst &in = t->inst[j.inst];
    __shared__ float2 s_cs[kMaxRotsSmem];
    int64_t nblk = (in.T + kVoteThreads - 1) / kVoteThreads;
    if (nblk > static_cast<int64_t>(gridDim.x)) nblk = gridDim.x;
    if (static_cast<int64_t>(blockIdx.x) >= nblk) return;
    vote_center_lanes_body<0>(in.pc, in.idx, j.targets_tr, in.T, sh.cos_tab, sh.sin_tab, sh.R, j.geom, j.grid, j.grid_capacity,
                              sh.replicas_max, 0, j.status, blockIdx.x, static_cast<int>(nblk), s_cs);
}

// replicas folded into copy 0 and the first-maximum arg-max in ONE pass (the fold kernel's sum feeds the key directly);
// the last CTA of a job converts to world space (train_dino.py:212-213) exactly like grid_argmax_kernel
__global__ void __launch_bounds__(256) frame_fold_argmax_kernel(const FrameTable *__restrict__ t, FrameShared sh) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.y) >= t->n_jobs) return;
    const FrameJob &j = t->job[blockIdx.y];
    const cppf_grid_geom *geom = j.geom;
    const bool overflow = geom->cells > j.grid_capacity || geom->cells > 0x7fffffffll;
    const int64_t cells = overflow ? 0 : geom->cells;
    const int replicas = replica_count(cells, j.grid_capacity, sh.replicas_max);
    uint32_t *grid = j.grid;
    unsigned long long best = 0ull;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < cells; i += stride) {
        uint32_t acc = grid[i];
        if (replicas > 1) {
            for (int k = 1; k < replicas; ++k) acc += grid[k * cells + i];
            grid[i] = acc;
        }
        const unsigned long long k = (static_cast<unsigned long long>(acc) << 32) |
                                     static_cast<unsigned long long>(0xffffffffu - static_cast<uint32_t>(i));
        best = k > best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    __shared__ unsigned long long s_best[8];
    __shared__ bool s_last;
    cppf_center *out = j.center;
    unsigned long long *key = reinterpret_cast<unsigned long long *>(&out->linear);     // zeroed by frame_prep_kernel
    unsigned int *ticket = reinterpret_cast<unsigned int *>(&out->status);
    if (lane_id() == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) best = s_best[w] > best ? s_best[w] : best;
        atomicMax(key, best);
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(key);
        int64_t lin = static_cast<int64_t>(0xffffffffu - static_cast<uint32_t>(k & 0xffffffffull));
        if (cells <= 0) lin = 0;
        const int64_t g1 = geom->grid_res[1], g2 = geom->grid_res[2];
        const int64_t cell[3] = {lin / (g1 * g2), (lin / g2) % g1, lin % g2};
        for (int a = 0; a < 3; ++a) {
            out->cell[a] = cell[a];
            out->world[a] = __dadd_rn(static_cast<double>(geom->lo[a]), __dmul_rn(static_cast<double>(cell[a]), j.res64));
        }
        out->linear = lin;
        out->votes = static_cast<uint32_t>(k >> 32);
        out->cells = geom->cells;
        out->status = geom->flags | (overflow ? CPPF_STATUS_GRID_OVERFLOW : 0u) | *j.status;
    }
}

int frame_launch_center(const FrameTable *t, int nj, int64_t T_cap, const FrameShared &sh, cudaStream_t s) {
    const DeviceInfo &dev = device_info();
    if (nj <= 0) return CPPF_OK;
    CPPF_CUDA_TRY(launch_frame_kernel(frame_prep_kernel, dim3(nj), dim3(1024), 0, s, t, sh));
    const int per_job = std::max(1, std::min<int>(div_up(T_cap, 256), (dev.sm_count * 8 + nj - 1) / nj));
    CPPF_CUDA_TRY(launch_frame_kernel(frame_decode_zero_kernel, dim3(per_job, nj), dim3(256), 0, s, t, sh));
    {
        const int64_t warps = (T_cap + 7) / 8;
        const int blocks = static_cast<int>(std::min<int64_t>((warps + 7) / 8, static_cast<int64_t>(dev.sm_count) * 8));
        if (sh.vote_lanes) {
            const int lane_blocks = static_cast<int>(std::min<int64_t>((T_cap + kVoteThreads - 1) / kVoteThreads, static_cast<int64_t>(dev.sm_count) * 8));
            CPPF_CUDA_TRY(launch_frame_kernel(frame_vote_center_lanes_kernel, dim3(lane_blocks, nj), dim3(kVoteThreads), 0, s, t, sh));
        } else {
            CPPF_CUDA_TRY(launch_frame_kernel(frame_vote_center_kernel, dim3(blocks, nj), dim3(kVoteThreads), 0, s, t, sh));
        }
    }
    CPPF_CUDA_TRY(launch_frame_kernel(frame_fold_argmax_kernel, dim3(std::max(1, dev.sm_count * 4 / std::max(1, nj / 2)), nj), dim3(256), 0, s, t, sh));
    return CPPF_OK;
}

}  // namespace cppf
