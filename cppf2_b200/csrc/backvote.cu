// Back-vote filter on B200 -- replaces eval.py:251-275.
//
//   errs[t]   = || targets_tr[t] - targets(real pair t, voted centre) ||_2          (float32, numpy order)
//   threshold = np.percentile(errs, ratio*100)  (numpy 'linear': float32 lerp of two order statistics)
//   keep[t]   = errs[t] < threshold;  imp[p] = #occurrences of p among kept pair endpoints
//
// The order statistics are exact: a 4 x 8-bit most-significant-digit radix selection over the
// order-preserving uint32 image of the float32 errors.  Each pass is one grid-wide histogram
// (shared-memory privatised, one global atomic per non-empty bin per CTA); the last CTA to finish a
// pass (ticket counter) scans the 256 bins and narrows the prefix, so the whole selection is four
// stream-ordered launches with no host involvement.  Up to 2^17 errors (a frame's jobs) the selection
// is ONE launch instead: a cluster of 8 CTAs per job, each with its slice of the keys in shared memory and
// the per-pass histograms summed through distributed shared memory (select_cluster_body).  The filter is
// split into errors / select / mask entry points so that a tuple-sharded run can all-gather the errors
// (4*T bytes) between them.
#include "common.cuh"
#include "frame.cuh"

#include <cooperative_groups.h>

namespace cppf {

struct Axes {
    double v[9];
};

struct SelectState {          // lives at the head of the workspace, zeroed by the host wrapper
    uint32_t hist[4][256];
    uint32_t prefix;          // key bits decided so far
    uint32_t min_gt;          // smallest key strictly greater than the selected one
    unsigned long long k;     // rank still to resolve inside the current prefix
    unsigned long long below; // #keys < prefix (final)
    unsigned long long equal; // #keys == prefix (final)
    uint32_t tickets[8];
};

__device__ __forceinline__ void back_targets(const float a[3], const float b[3], const double ctr[3], float tr[2]) {
    // generate_target_pairs(input_pairs, ..., center=T_est) -- dataset.py:118-128, translation part only
    const float pd0 = __fsub_rn(a[0], b[0]), pd1 = __fsub_rn(a[1], b[1]), pd2 = __fsub_rn(a[2], b[2]);
    const float nrm = __fadd_rn(norm3_numpy(pd0, pd1, pd2), 1e-7f);
    const double u0 = static_cast<double>(__fdiv_rn(pd0, nrm)), u1 = static_cast<double>(__fdiv_rn(pd1, nrm)),
                 u2 = static_cast<double>(__fdiv_rn(pd2, nrm));
    const double am0 = __dsub_rn(static_cast<double>(a[0]), ctr[0]), am1 = __dsub_rn(static_cast<double>(a[1]), ctr[1]),
                 am2 = __dsub_rn(static_cast<double>(a[2]), ctr[2]);
    const double proj = __dadd_rn(__dadd_rn(__dmul_rn(am0, u0), __dmul_rn(am1, u1)), __dmul_rn(am2, u2));
    const double oc0 = __dsub_rn(am0, __dmul_rn(proj, u0)), oc1 = __dsub_rn(am1, __dmul_rn(proj, u1)),
                 oc2 = __dsub_rn(am2, __dmul_rn(proj, u2));
    const double dist = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(oc0, oc0), __dmul_rn(oc1, oc1)), __dmul_rn(oc2, oc2)));
    tr[0] = static_cast<float>(proj);
    tr[1] = static_cast<float>(dist);
}

__device__ __forceinline__ void backvote_errors_body(const float *__restrict__ pc, const IdxView &idx,
                                                     const float *__restrict__ targets_tr, int64_t T,
                                                     const cppf_center *__restrict__ center, float *__restrict__ errs,
                                                     int bid, int nblk) {
    const double ctr[3] = {center->world[0], center->world[1], center->world[2]};
    const int64_t stride = static_cast<int64_t>(nblk) * blockDim.x;
    for (int64_t t = static_cast<int64_t>(bid) * blockDim.x + threadIdx.x; t < T; t += stride) {
        const int64_t ia = idx.at(t, 0), ib = idx.at(t, 1);
        const float a[3] = {pc[3 * ia], pc[3 * ia + 1], pc[3 * ia + 2]};
        const float b[3] = {pc[3 * ib], pc[3 * ib + 1], pc[3 * ib + 2]};
        float back[2];
        back_targets(a, b, ctr, back);
        const float2 tr = reinterpret_cast<const float2 *>(targets_tr)[t];
        const float d0 = __fsub_rn(tr.x, back[0]), d1 = __fsub_rn(tr.y, back[1]);
        errs[t] = __fsqrt_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)));  // np.linalg.norm(axis=-1), eval.py:256
    }
}

__global__ void __launch_bounds__(256) backvote_errors_kernel(const float *__restrict__ pc, IdxView idx,
                                                              const float *__restrict__ targets_tr, int64_t T,
                                                              const cppf_center *__restrict__ center,
                                                              float *__restrict__ errs) {
    backvote_errors_body(pc, idx, targets_tr, T, center, errs, blockIdx.x, gridDim.x);
}

// One radix pass: histogram digit `pass` (most significant first) of the keys that match the prefix
// found so far; the last CTA narrows the prefix.  pass == 4 instead finds min_gt and the threshold.
__global__ void __launch_bounds__(256) select_pass_kernel(const float *__restrict__ errs, int64_t T, int pass,
                                                          SelectState *__restrict__ st, int64_t rank_lo, float gamma,
                                                          cppf_backvote_summary *__restrict__ summary) {
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_min;
    __shared__ bool s_last;
    s_hist[threadIdx.x] = 0u;
    if (threadIdx.x == 0) s_min = 0xffffffffu;
    __syncthreads();
    const uint32_t prefix = st->prefix;
    const int shift = 24 - 8 * pass;
    const uint32_t decided = pass == 0 ? 0u : (pass >= 4 ? 0xffffffffu : ~((1u << (shift + 8)) - 1u));
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    uint32_t local_min = 0xffffffffu;
    // whole warps iterate together (warp_hist_add is a full-warp operation)
    for (int64_t base = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x - lane_id(); base < T; base += stride) {
        const int64_t t = base + lane_id();
        const bool valid = t < T;
        const uint32_t key = valid ? float_to_key(errs[t]) : 0u;
        if (pass < 4) {
            // back-vote errors share their leading digits (same exponent): one leader adds for the warp when all lanes agree
            warp_hist_add(s_hist, (key >> shift) & 0xffu, valid && (key & decided) == (prefix & decided));
        } else if (valid && key > prefix) {
            local_min = key < local_min ? key : local_min;
        }
    }
    if (pass < 4) {
        __syncthreads();
        const uint32_t c = s_hist[threadIdx.x];
        if (c) atomicAdd(&st->hist[pass][threadIdx.x], c);
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            uint32_t other = __shfl_xor_sync(0xffffffffu, local_min, o);
            local_min = other < local_min ? other : local_min;
        }
        if (lane_id() == 0) atomicMin(&s_min, local_min);
        __syncthreads();
        if (threadIdx.x == 0) atomicMin(&st->min_gt, s_min);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->tickets[pass], 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (pass < 4) {
        // the last CTA scans the 256 bins together: inclusive prefix sum, then the first bin whose cumulative
        // count exceeds the remaining rank is the digit (a one-thread scan costs ~200 dependent L2 reads)
        __shared__ unsigned long long s_cum[256];
        __shared__ int s_digit;
        const unsigned long long k = pass == 0 ? static_cast<unsigned long long>(rank_lo) : st->k;
        const unsigned long long below = pass == 0 ? 0ull : st->below;
        const unsigned long long mine = *reinterpret_cast<volatile uint32_t *>(&st->hist[pass][threadIdx.x]);
        s_cum[threadIdx.x] = mine;
        if (threadIdx.x == 0) s_digit = 256;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const unsigned long long add = threadIdx.x >= o ? s_cum[threadIdx.x - o] : 0ull;
            __syncthreads();
            s_cum[threadIdx.x] += add;
            __syncthreads();
        }
        if (s_cum[threadIdx.x] > k) atomicMin(&s_digit, static_cast<int>(threadIdx.x));
        __syncthreads();
        if (threadIdx.x != 0) return;
        int d = s_digit;
        if (d == 256) d = 255;  // rank beyond the data (T == 0); keeps the state well defined
        const unsigned long long cum = d > 0 ? s_cum[d - 1] : 0ull;
        st->prefix = prefix | (static_cast<uint32_t>(d) << shift);
        st->k = k - cum;
        st->below = below + cum;
        st->equal = s_cum[d] - cum;
    } else if (threadIdx.x == 0) {
        // both order statistics are known: s[rank_lo] = prefix; s[rank_lo+1] is the same value when the
        // run of equal keys extends past rank_lo, else the smallest larger key (numpy clips at T-1)
        const float s_lo = key_to_float(prefix);
        float s_hi = s_lo;
        const unsigned long long next_rank = static_cast<unsigned long long>(rank_lo) + 1ull;
        if (next_rank >= st->below + st->equal && next_rank < static_cast<unsigned long long>(T))
            s_hi = key_to_float(*reinterpret_cast<volatile uint32_t *>(&st->min_gt));
        // numpy _lerp in float32: a + (b-a)*t, replaced by b - (b-a)*(1-t) when t >= 0.5
        const float diff = __fsub_rn(s_hi, s_lo);
        float thr = __fadd_rn(s_lo, __fmul_rn(diff, gamma));
        if (gamma >= 0.5f) thr = __fsub_rn(s_hi, __fmul_rn(diff, __fsub_rn(1.0f, gamma)));
        summary->threshold = thr;
        summary->s_lo = s_lo;
        summary->s_hi = s_hi;
    }
}

// The same selection in ONE launch for the sizes the instance loop runs (T = 50 000): a single 1024-thread CTA makes the
// four radix passes and the min-greater pass over the errors (200 KB, L2-resident) with its state in shared memory.  Seven
// launches (two memsets + five passes of ~5 us each, mostly launch latency) become one; results are identical (integer counts).
constexpr int64_t kSelectSmallMax = 1 << 17;

// kCached: the CTA first copies the keys into its shared memory (T * 4 bytes: 200 KB at T = 50 000, B200 has 227 KB per CTA),
// so the errors are read from L2 once instead of five times -- a single CTA sweeping 200 KB is bound by L2 latency, and the
// sweeps were 60 of the kernel's 75 us at T = 50 000.
constexpr int64_t kSelectCachedMax = 55 * 1024;            // keys the shared-memory cache holds (220 KB)

template <bool kCached>
__device__ __forceinline__ void select_small_body(const float *__restrict__ errs, int64_t T, int64_t rank_lo, float gamma,
                                                  cppf_backvote_summary *__restrict__ summary) {
    __shared__ uint32_t s_hist[256];
    __shared__ unsigned long long s_cum[256];
    __shared__ uint32_t s_prefix, s_min;
    __shared__ unsigned long long s_k, s_below, s_equal;
    __shared__ int s_digit;
    const int tid = threadIdx.x, lane = lane_id();
    if (tid == 0) {
        s_prefix = 0u;
        s_k = static_cast<unsigned long long>(rank_lo);
        s_below = 0ull;
        s_equal = 0ull;
        s_min = 0xffffffffu;
    }
    // thread tid owns the tuples u * blockDim.x + tid: coalesced, and every warp iterates the same u (full-width ballots)
    extern __shared__ __align__(16) uint32_t s_keys[];
    const int per = static_cast<int>((T + blockDim.x - 1) / blockDim.x);
    if (kCached) {
        for (int64_t t = tid; t < T; t += blockDim.x) s_keys[t] = float_to_key(__ldg(errs + t));     // all loads in flight at once
        __syncthreads();
    }
    auto key_at = [&](int u, bool &in) -> uint32_t {
        const int64_t t = static_cast<int64_t>(u) * blockDim.x + tid;
        in = t < T;
        if (!in) return 0u;
        return kCached ? s_keys[t] : float_to_key(__ldg(errs + t));
    };
    for (int pass = 0; pass < 4; ++pass) {
        if (tid < 256) s_hist[tid] = 0u;
        if (tid == 0) s_digit = 256;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const int shift = 24 - 8 * pass;
        const uint32_t decided = pass == 0 ? 0u : ~((1u << (shift + 8)) - 1u);
        auto count = [&](int u) {
            bool in;
            const uint32_t key = key_at(u, in);
            const bool in_prefix = in && (key & decided) == (prefix & decided);
            warp_hist_add(s_hist, (key >> shift) & 0xffu, in_prefix);
        };
        for (int u = 0; u < per; ++u) count(u);
        __syncthreads();
        if (tid < 256) s_cum[tid] = s_hist[tid];
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {                                     // inclusive scan of the 256 bins
            const unsigned long long add = (tid < 256 && tid >= o) ? s_cum[tid - o] : 0ull;
            __syncthreads();
            if (tid < 256) s_cum[tid] += add;
            __syncthreads();
        }
        const unsigned long long k = s_k;
        if (tid < 256 && s_cum[tid] > k) atomicMin(&s_digit, tid);
        __syncthreads();
        if (tid == 0) {
            int d = s_digit;
            if (d == 256) d = 255;
            const unsigned long long cum = d > 0 ? s_cum[d - 1] : 0ull;
            s_prefix = prefix | (static_cast<uint32_t>(d) << shift);
            s_k = k - cum;
            s_below += cum;
            s_equal = s_cum[d] - cum;
        }
        __syncthreads();
    }
    const uint32_t prefix = s_prefix;
    uint32_t local_min = 0xffffffffu;
    for (int u = 0; u < per; ++u) {
        bool in;
        const uint32_t key = key_at(u, in);
        if (in && key > prefix && key < local_min) local_min = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t other = __shfl_xor_sync(0xffffffffu, local_min, o);
        local_min = other < local_min ? other : local_min;
    }
    if (lane == 0) atomicMin(&s_min, local_min);
    __syncthreads();
    if (tid != 0) return;
    // both order statistics are known (same rules as the last pass of select_pass_kernel)
    const float s_lo = key_to_float(prefix);
    float s_hi = s_lo;
    const unsigned long long next_rank = static_cast<unsigned long long>(rank_lo) + 1ull;
    if (next_rank >= s_below + s_equal && next_rank < static_cast<unsigned long long>(T)) s_hi = key_to_float(s_min);
    const float diff = __fsub_rn(s_hi, s_lo);
    float thr = __fadd_rn(s_lo, __fmul_rn(diff, gamma));
    if (gamma >= 0.5f) thr = __fsub_rn(s_hi, __fmul_rn(diff, __fsub_rn(1.0f, gamma)));
    summary->threshold = thr;
    summary->s_lo = s_lo;
    summary->s_hi = s_hi;
}

__global__ void __launch_bounds__(1024) select_small_kernel(const float *__restrict__ errs, int64_t T, int64_t rank_lo, float gamma,
                                                            cppf_backvote_summary *__restrict__ summary, int cached) {
    if (cached) select_small_body<true>(errs, T, rank_lo, gamma, summary);
    else select_small_body<false>(errs, T, rank_lo, gamma, summary);
}

// The same selection spread over a thread-block CLUSTER of kSelectCluster CTAs (round 2, session 3).  One CTA sweeping
// 50 000 keys five times is bound by its own issue rate (49 keys per thread and pass, ~25 instructions each: 54 us per frame on
// 12 of the 148 SMs, ncu).  Here every CTA of the cluster caches 1/8 of the keys in its shared memory and histograms only
// those; after one cluster barrier per pass every CTA sums the eight 256-bin histograms through distributed shared memory and
// narrows the prefix by itself (identical integer arithmetic in all eight, so there is nothing to broadcast).  The histograms of
// the four passes are separate arrays: a pass needs ONE cluster barrier, none to recycle a buffer.  Integer counts, hence the
// same order statistics and the same threshold bits as the single-CTA and the multi-launch forms.
constexpr int kSelectCluster = 8;
constexpr int kSelectClusterThreads = 512;

static size_t select_cluster_cache_bytes(int64_t T) {
    return static_cast<size_t>((T + kSelectCluster - 1) / kSelectCluster) * sizeof(uint32_t);
}

__device__ __forceinline__ void select_cluster_body(const float *__restrict__ errs, int64_t T, int64_t rank_lo, float gamma,
                                                    cppf_backvote_summary *__restrict__ summary) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(cluster.block_rank());
    __shared__ uint32_t s_hist[4][256];
    __shared__ unsigned long long s_cum[256];
    __shared__ unsigned long long s_warp_tot[8];
    __shared__ uint32_t s_min;
    __shared__ int s_digit;
    extern __shared__ __align__(16) uint32_t s_keys[];
    const int tid = threadIdx.x, lane = lane_id();
    const int64_t per_cta = (T + kSelectCluster - 1) / kSelectCluster;
    const int64_t first = static_cast<int64_t>(rank) * per_cta;
    const int n = static_cast<int>(max(static_cast<int64_t>(0), min(T, first + per_cta) - first));
    for (int i = tid; i < 4 * 256; i += blockDim.x) (&s_hist[0][0])[i] = 0u;
    if (tid == 0) s_min = 0xffffffffu;
    for (int i = tid; i < n; i += blockDim.x) s_keys[i] = float_to_key(__ldg(errs + first + i));       // all loads in flight at once
    __syncthreads();
    // the selection state lives in registers: every thread of every CTA derives the same values from the summed histograms
    uint32_t prefix = 0u;
    unsigned long long k = static_cast<unsigned long long>(rank_lo), below = 0ull, equal = 0ull;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const uint32_t decided = pass == 0 ? 0u : ~((1u << (shift + 8)) - 1u);
        // whole warps iterate together (warp_hist_add is a full-warp operation)
        for (int i0 = tid - lane; i0 < n; i0 += blockDim.x) {
            const int i = i0 + lane;
            const bool in = i < n;
            const uint32_t key = in ? s_keys[i] : 0u;
            warp_hist_add(s_hist[pass], (key >> shift) & 0xffu, in && (key & decided) == (prefix & decided));
        }
        if (tid == 0) s_digit = 256;
        cluster.sync();                                   // every CTA's histogram of this pass is complete and visible
        unsigned long long incl = 0ull;
        if (tid < 256) {
#pragma unroll
            for (int r = 0; r < kSelectCluster; ++r) incl += *cluster.map_shared_rank(&s_hist[pass][tid], r);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {            // inclusive scan of the 256 bins: warp scans, then the warp totals
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
        if (d == 256) d = 255;                            // rank beyond the data; keeps the state well defined
        const unsigned long long cum = d > 0 ? s_cum[d - 1] : 0ull;
        prefix |= static_cast<uint32_t>(d) << shift;
        equal = s_cum[d] - cum;
        k -= cum;
        below += cum;
        __syncthreads();                                  // s_cum / s_digit / s_warp_tot are rewritten by the next pass
    }
    uint32_t local_min = 0xffffffffu;
    for (int i = tid; i < n; i += blockDim.x) {
        const uint32_t key = s_keys[i];
        if (key > prefix && key < local_min) local_min = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t other = __shfl_xor_sync(0xffffffffu, local_min, o);
        local_min = other < local_min ? other : local_min;
    }
    if (lane == 0 && local_min != 0xffffffffu) atomicMin(&s_min, local_min);
    cluster.sync();
    if (rank == 0 && tid == 0) {
        uint32_t min_gt = 0xffffffffu;
        for (int r = 0; r < kSelectCluster; ++r) {
            const uint32_t v = *cluster.map_shared_rank(&s_min, r);
            min_gt = v < min_gt ? v : min_gt;
        }
        // both order statistics are known (same rules as the last pass of select_pass_kernel)
        const float s_lo = key_to_float(prefix);
        float s_hi = s_lo;
        const unsigned long long next_rank = static_cast<unsigned long long>(rank_lo) + 1ull;
        if (next_rank >= below + equal && next_rank < static_cast<unsigned long long>(T)) s_hi = key_to_float(min_gt);
        const float diff = __fsub_rn(s_hi, s_lo);
        float thr = __fadd_rn(s_lo, __fmul_rn(diff, gamma));
        if (gamma >= 0.5f) thr = __fsub_rn(s_hi, __fmul_rn(diff, __fsub_rn(1.0f, gamma)));
        summary->threshold = thr;
        summary->s_lo = s_lo;
        summary->s_hi = s_hi;
    }
    cluster.sync();                                       // nobody leaves while CTA 0 still reads its neighbours' minima
}

__global__ void __launch_bounds__(kSelectClusterThreads) select_cluster_kernel(const float *__restrict__ errs, int64_t T, int64_t rank_lo,
                                                                              float gamma, cppf_backvote_summary *__restrict__ summary) {
    select_cluster_body(errs, T, rank_lo, gamma, summary);
}

// cluster launch (kSelectCluster CTAs along x), with the frame path's programmatic-dependent-launch attribute when asked for
template <typename... KArgs, typename... Args>
static cudaError_t launch_select_cluster(void (*kernel)(KArgs...), int jobs, size_t smem, bool pdl, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kSelectCluster, jobs);
    cfg.blockDim = dim3(kSelectClusterThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2]{};
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kSelectCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && frame_pdl_enabled()) ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// CPPF_SELECT_CLUSTER=0 keeps the single-CTA form (A/B measurements)
static bool select_cluster_enabled() {
    static const bool on = [] {
        const char *e = getenv("CPPF_SELECT_CLUSTER");
        return !(e && e[0] == '0');
    }();
    return on;
}

// dynamic shared memory of the cached form for up to T keys (0: the keys do not fit, sweep L2)
static size_t select_cache_bytes(int64_t T) { return T <= kSelectCachedMax ? static_cast<size_t>(T) * sizeof(uint32_t) : 0; }

// imp_max != nullptr: the running maximum of the occurrence counts is kept as well (the count a point ends with is the value
// its last increment returns plus one, so the maximum over all increments is the final maximum; one atomicMax per warp) --
// the batched path then needs no separate max kernel.
__device__ __forceinline__ void backvote_mask_body(const float *__restrict__ errs, const IdxView &idx, int64_t T,
                                                   const cppf_backvote_summary *__restrict__ summary_in,
                                                   uint8_t *__restrict__ keep, int32_t *__restrict__ kept_list,
                                                   int32_t *__restrict__ imp, unsigned long long *__restrict__ kept_counter,
                                                   int32_t *__restrict__ imp_max, int bid, int nblk) {
    const float thr = summary_in->threshold;
    const int64_t stride = static_cast<int64_t>(nblk) * blockDim.x;
    const int64_t t0 = static_cast<int64_t>(bid) * blockDim.x + threadIdx.x;
    int seen_max = 0;
    // whole warps iterate together so that the ballot below is always full-width
    for (int64_t base = t0 - lane_id(); base < T; base += stride) {
        const int64_t t = base + lane_id();
        const bool k = t < T && errs[t] < thr;  // strict '<' (eval.py:257)
        if (t < T && keep) keep[t] = k ? 1 : 0;
        const uint32_t m = __ballot_sync(0xffffffffu, k);
        if (m == 0u) continue;
        unsigned long long slot = 0;
        if (lane_id() == 0) slot = atomicAdd(kept_counter, static_cast<unsigned long long>(__popc(m)));
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (k) {
            if (kept_list) kept_list[slot + __popc(m & ((1u << lane_id()) - 1u))] = static_cast<int32_t>(t);
            if (imp) {
                const int c0 = atomicAdd(&imp[idx.at(t, 0)], 1);  // scatter_add of the flattened endpoints (eval.py:260-266)
                const int c1 = atomicAdd(&imp[idx.at(t, 1)], 1);
                seen_max = max(seen_max, max(c0, c1) + 1);
            }
        }
    }
    if (imp && imp_max) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) seen_max = max(seen_max, __shfl_xor_sync(0xffffffffu, seen_max, o));
        if (lane_id() == 0 && seen_max > 0) atomicMax(imp_max, seen_max);
    }
}

__global__ void __launch_bounds__(256) backvote_mask_kernel(const float *__restrict__ errs, IdxView idx, int64_t T,
                                                            const cppf_backvote_summary *__restrict__ summary_in,
                                                            uint8_t *__restrict__ keep, int32_t *__restrict__ kept_list,
                                                            int32_t *__restrict__ imp,
                                                            unsigned long long *__restrict__ kept_counter) {
    backvote_mask_body(errs, idx, T, summary_in, keep, kept_list, imp, kept_counter, nullptr, blockIdx.x, gridDim.x);
}

__global__ void __launch_bounds__(1024) imp_max_kernel(const int32_t *__restrict__ imp, int64_t n,
                                                       cppf_backvote_summary *__restrict__ summary) {
    __shared__ int s_max[32];
    int m = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = max(m, imp[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane_id() == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = s_max[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) summary->imp_max = m;
    }
}

}  // namespace cppf

using namespace cppf;

CPPF_API int64_t cppf_backvote_workspace_bytes(int64_t T, int64_t n) {
    (void)T;
    (void)n;
    return static_cast<int64_t>((sizeof(SelectState) + 255) / 256 * 256);
}

CPPF_API int cppf_backvote_errors(const float *pc, const void *idx, int idx_is_i64, int64_t idx_stride,
                                  const float *targets_tr, int64_t T, const cppf_center *center, float *errs,
                                  void *stream) {
    if (!pc || !idx || !targets_tr || !center || !errs || T < 0 || idx_stride < 2) return CPPF_ERR_INVALID_ARGUMENT;
    if (T == 0) return CPPF_OK;
    IdxView iv{idx, idx_stride, idx_is_i64};
    backvote_errors_kernel<<<grid_for(T, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(pc, iv, targets_tr, T,
                                                                                                center, errs);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_backvote_select(const float *errs, int64_t T, int64_t rank_lo, float gamma,
                                  cppf_backvote_summary *summary, void *ws, int64_t ws_bytes, void *stream) {
    if (!errs || !summary || !ws || T <= 0 || rank_lo < 0 || rank_lo >= T) return CPPF_ERR_INVALID_ARGUMENT;
    if (ws_bytes < cppf_backvote_workspace_bytes(T, 0)) return CPPF_ERR_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (T <= kSelectSmallMax && select_cluster_enabled()) {
        const size_t cache = select_cluster_cache_bytes(T);
        if (cache > 48 * 1024)
            CPPF_TRY_ONCE_PER_DEVICE(cudaFuncSetAttribute(select_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          static_cast<int>(select_cluster_cache_bytes(kSelectSmallMax))));
        CPPF_CUDA_TRY(launch_select_cluster(select_cluster_kernel, 1, cache, false, s, errs, T, rank_lo, gamma, summary));
        return CPPF_OK;
    }
    if (T <= kSelectSmallMax) {
        const size_t cache = select_cache_bytes(T);
        if (cache > 48 * 1024)
            CPPF_TRY_ONCE_PER_DEVICE(cudaFuncSetAttribute(select_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          static_cast<int>(kSelectCachedMax * sizeof(uint32_t))));
        select_small_kernel<<<1, 1024, cache, s>>>(errs, T, rank_lo, gamma, summary, cache ? 1 : 0);
        CPPF_LAUNCH_CHECK();
        return CPPF_OK;
    }
    SelectState *st = static_cast<SelectState *>(ws);
    CPPF_CUDA_TRY(cudaMemsetAsync(st, 0, sizeof(SelectState), s));
    CPPF_CUDA_TRY(cudaMemsetAsync(&st->min_gt, 0xff, sizeof(uint32_t), s));
    const int blocks = grid_for(T, 256, 4);
    for (int pass = 0; pass <= 4; ++pass) {
        select_pass_kernel<<<blocks, 256, 0, s>>>(errs, T, pass, st, rank_lo, gamma, summary);
        CPPF_LAUNCH_CHECK();
    }
    return CPPF_OK;
}

CPPF_API int cppf_backvote_mask(const float *errs, const void *idx, int idx_is_i64, int64_t idx_stride, int64_t T,
                                int64_t n, cppf_backvote_summary *summary, uint8_t *keep, int32_t *kept_list,
                                int32_t *imp, int zero_outputs, void *stream) {
    if (!errs || !idx || !summary || T < 0 || idx_stride < 2) return CPPF_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (zero_outputs) {
        if (imp) CPPF_CUDA_TRY(cudaMemsetAsync(imp, 0, sizeof(int32_t) * static_cast<size_t>(n), s));
        CPPF_CUDA_TRY(cudaMemsetAsync(&summary->kept, 0, sizeof(int64_t), s));
    }
    if (T == 0) return CPPF_OK;
    IdxView iv{idx, idx_stride, idx_is_i64};
    backvote_mask_kernel<<<grid_for(T, 256, 8), 256, 0, s>>>(errs, iv, T, summary, keep, kept_list, imp,
                                                            reinterpret_cast<unsigned long long *>(&summary->kept));
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_backvote_imp_max(const int32_t *imp, int64_t n, cppf_backvote_summary *summary, void *stream) {
    if (!imp || !summary || n <= 0) return CPPF_ERR_INVALID_ARGUMENT;
    imp_max_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(imp, n, summary);
    CPPF_LAUNCH_CHECK();
    return CPPF_OK;
}

CPPF_API int cppf_backvote_filter(const float *pc, int64_t n, const void *idx, int idx_is_i64, int64_t idx_stride,
                                  const float *targets_tr, int64_t T, const double *axes_host,
                                  const cppf_center *center, int64_t rank_lo, float gamma, float *errs, uint8_t *keep,
                                  int32_t *kept_list, int32_t *imp, cppf_backvote_summary *summary, void *ws,
                                  int64_t ws_bytes, void *stream) {
    (void)axes_host;  // the translation targets do not depend on the axes
    int rc = cppf_backvote_errors(pc, idx, idx_is_i64, idx_stride, targets_tr, T, center, errs, stream);
    if (rc) return rc;
    rc = cppf_backvote_select(errs, T, rank_lo, gamma, summary, ws, ws_bytes, stream);
    if (rc) return rc;
    rc = cppf_backvote_mask(errs, idx, idx_is_i64, idx_stride, T, n, summary, keep, kept_list, imp, 1, stream);
    if (rc) return rc;
    return cppf_backvote_imp_max(imp, n, summary, stream);
}

// =====================================================================================================================
// Batched frame path (frame.cuh): the back-vote filter of every job in three launches (errors, exact selection, mask).
// =====================================================================================================================
namespace cppf {

__global__ void __launch_bounds__(256) frame_backvote_errors_kernel(const FrameTable *__restrict__ t) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.y) >= t->n_jobs) return;
    const FrameJob &j = t->job[blockIdx.y];
    const FrameInst &in = t->inst[j.inst];
    backvote_errors_body(in.pc, in.idx, j.targets_tr, in.T, j.center, j.errs, blockIdx.x, gridDim.x);
}

__global__ void __launch_bounds__(1024) frame_select_kernel(const FrameTable *__restrict__ t, int cached) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.x) >= t->n_jobs) return;
    const FrameJob &j = t->job[blockIdx.x];
    const FrameInst &in = t->inst[j.inst];
    if (in.T <= 0) return;
    if (cached) select_small_body<true>(j.errs, in.T, j.rank_lo, j.gamma, j.summary);
    else select_small_body<false>(j.errs, in.T, j.rank_lo, j.gamma, j.summary);
}

__global__ void __launch_bounds__(kSelectClusterThreads) frame_select_cluster_kernel(const FrameTable *__restrict__ t) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.y) >= t->n_jobs) return;          // the whole cluster of a job leaves together
    const FrameJob &j = t->job[blockIdx.y];
    const FrameInst &in = t->inst[j.inst];
    if (in.T <= 0) return;
    select_cluster_body(j.errs, in.T, j.rank_lo, j.gamma, j.summary);
}

__global__ void __launch_bounds__(256) frame_backvote_mask_kernel(const FrameTable *__restrict__ t) {
    pdl_enter();      // frame path: programmatic dependent launch (common.cuh)
    if (static_cast<int>(blockIdx.y) >= t->n_jobs) return;
    const FrameJob &j = t->job[blockIdx.y];
    const FrameInst &in = t->inst[j.inst];
    // imp [n], the kept counter and imp_max were zeroed by frame_prep_kernel
    backvote_mask_body(j.errs, in.idx, in.T, j.summary, j.keep, j.kept_list, j.imp,
                       reinterpret_cast<unsigned long long *>(&j.summary->kept), &j.summary->imp_max, blockIdx.x, gridDim.x);
}

int frame_launch_backvote(const FrameTable *t, int nj, int64_t T_cap, cudaStream_t s) {
    if (nj <= 0) return CPPF_OK;
    if (T_cap > kSelectSmallMax) return CPPF_ERR_UNSUPPORTED;      // the single-CTA selection; larger T: the per-job path
    const int per_job = std::max(1, std::min<int>(div_up(T_cap, 256), (device_info().sm_count * 8 + nj - 1) / nj));
    CPPF_CUDA_TRY(launch_frame_kernel(frame_backvote_errors_kernel, dim3(per_job, nj), dim3(256), 0, s, t));
    if (select_cluster_enabled()) {
        const size_t cache = select_cluster_cache_bytes(T_cap);
        if (cache > 48 * 1024)
            CPPF_TRY_ONCE_PER_DEVICE(cudaFuncSetAttribute(frame_select_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          static_cast<int>(select_cluster_cache_bytes(kSelectSmallMax))));
        CPPF_CUDA_TRY(launch_select_cluster(frame_select_cluster_kernel, nj, cache, true, s, t));
    } else {
        const size_t cache = select_cache_bytes(T_cap);
        if (cache > 48 * 1024)
            CPPF_TRY_ONCE_PER_DEVICE(cudaFuncSetAttribute(frame_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          static_cast<int>(kSelectCachedMax * sizeof(uint32_t))));
        CPPF_CUDA_TRY(launch_frame_kernel(frame_select_kernel, dim3(nj), dim3(1024), cache, s, t, cache ? 1 : 0));
    }
    CPPF_CUDA_TRY(launch_frame_kernel(frame_backvote_mask_kernel, dim3(per_job, nj), dim3(256), 0, s, t));
    return CPPF_OK;
}

}  // namespace cppf
