// Probe: sustained tcgen05.mma rate of ONE issuing thread per SM for the operand layouts heads_tc.cu can use
// (kind::f16, bf16, M = 128, cta_group::1).  Garbage operands: only the timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_mma_rate probe_mma_rate.cu && ./probe_mma_rate
// Modes: 0 SS no-swizzle (8x16 B core matrices, LBO = plane), 1 TS (A in TMEM) + no-swizzle B, 2 SS 128B-swizzle,
//        3 TS + 128B-swizzle B.   Reports cycles per MMA as seen by the issuing thread (issue only) and until the
//        commit barrier fires (execution).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3fffu) | (static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t desc_swz128(uint32_t addr) {   // K-major, 128 B rows, 8-row groups 1024 B apart
    return static_cast<uint64_t>((addr >> 4) & 0x3fffu) | (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t instr_desc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a),
                 "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

template <int mode, int commit_every>
__global__ void __launch_bounds__(128, 1) probe(int n, int iters, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];   // 128 KB of operands
    __shared__ __align__(8) unsigned long long bar[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 131072 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (warp == 1) {   // warp-uniform control flow; only the tcgen05 instructions are predicated on the elected lane
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        const uint32_t idesc = instr_desc(n);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 65536;
        const long long t0 = clock64();
        uint64_t bd[8], ad[8];
#pragma unroll
        for (uint32_t ks = 0; ks < 8; ++ks) {
            if (mode < 2) {
                bd[ks] = desc_noswz(b_base + ks * 2 * (n * 16), n * 16, 128);
                ad[ks] = desc_noswz(a_base + ks * 4096, 2048, 128);
            } else {
                bd[ks] = desc_swz128(b_base + (ks >> 2) * (n * 128) + (ks & 3) * 32);
                ad[ks] = desc_swz128(a_base + (ks >> 2) * 16384 + (ks & 3) * 32);
            }
        }
        for (int i = 0; i < iters; i += 8) {
            if (leader) {
#pragma unroll
                for (uint32_t ks = 0; ks < 8; ++ks) {
                    if (mode & 1) mma_ts(tmem, tmem + 256 + ks * 8, bd[ks], idesc, (i + ks) > 0);
                    else mma_ss(tmem, ad[ks], bd[ks], idesc, (i + ks) > 0);
                    if (commit_every && (ks & (commit_every - 1)) == commit_every - 1) commit(smem_u32(&bar[1]));
                }
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (leader) commit(smem_u32(&bar[0]));
        __syncwarp();
        wait(smem_u32(&bar[0]), 0);
        const long long t2 = clock64();
        if (leader) {
            out[blockIdx.x * 2] = t1 - t0;
            out[blockIdx.x * 2 + 1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// SS / TS MMA rate under competing shared-memory traffic: warps 2.. store (or load) 16 B per lane in a loop with `gap`
// dependent ALU operations between accesses, into a region the MMAs do not touch.
template <int mode>
__global__ void __launch_bounds__(256, 1) probe_traffic(int n, int iters, int gap, int loads, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];   // 128 KB of operands + 32 KB traffic region
    __shared__ __align__(8) unsigned long long bar[2];
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 163840 / 16; i += 256) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        s_stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (warp == 1) {
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        const uint32_t idesc = instr_desc(n);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 65536;
        uint64_t bd[8], ad[8];
#pragma unroll
        for (uint32_t ks = 0; ks < 8; ++ks) {
            bd[ks] = desc_noswz(b_base + ks * 2 * (n * 16), n * 16, 128);
            ad[ks] = desc_noswz(a_base + ks * 4096, 2048, 128);
        }
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
            if (leader) {
#pragma unroll
                for (uint32_t ks = 0; ks < 8; ++ks) {
                    if (mode & 1) mma_ts(tmem, tmem + 256 + ks * 8, bd[ks], idesc, (i + ks) > 0);
                    else mma_ss(tmem, ad[ks], bd[ks], idesc, (i + ks) > 0);
                }
            }
            __syncwarp();
        }
        if (leader) commit(smem_u32(&bar[0]));
        __syncwarp();
        wait(smem_u32(&bar[0]), 0);
        const long long t2 = clock64();
        if (leader) {
            out[blockIdx.x * 2] = t2 - t0;
            out[blockIdx.x * 2 + 1] = t2 - t0;
        }
        s_stop = 1;
    } else if (warp >= 2) {
        unsigned char *region = smem + 131072;
        uint4 v = make_uint4(tid, 1, 2, 3);
        uint32_t acc = tid;
        int it = 0;
        while (!s_stop) {
            unsigned char *p = region + ((it & 15) * 2048) + ((tid - 64) & 127) * 16;
            if (loads) {
                uint32_t r0, r1, r2, r3;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(p)));
                acc += r0;
            } else {
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
            for (int g = 0; g < gap; ++g) acc = acc * 1664525u + 1013904223u;
            ++it;
        }
        if (acc == 0x12345u) out[0] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// Round trip of one dependent step: `burst` MMAs -> commit -> (another warp) wait -> tcgen05.ld -> arrive -> issuer wait.
__global__ void __launch_bounds__(160, 1) roundtrip(int n, int burst, int iters, int with_epilogue, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 131072 / 16; i += 160) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (warp == 4) {
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        const uint32_t idesc = instr_desc(n);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 65536;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (leader) {
                for (int ks = 0; ks < burst; ++ks)
                    mma_ss(tmem, desc_noswz(a_base + (ks & 7) * 4096, 2048, 128), desc_noswz(b_base + (ks & 7) * 2 * (n * 16), n * 16, 128), idesc, ks > 0);
                commit(smem_u32(&bar[0]));
            }
            __syncwarp();
            if (with_epilogue) wait(smem_u32(&bar[1]), i & 1);
            else wait(smem_u32(&bar[0]), i & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (leader) out[blockIdx.x * 2] = out[blockIdx.x * 2 + 1] = clock64() - t0;
    } else if (with_epilogue) {
        for (int i = 0; i < iters; ++i) {
            wait(smem_u32(&bar[0]), i & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tmem + (static_cast<uint32_t>(warp * 32) << 16)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (r[0] == 0x12345678u) out[0] = 1;
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
    {
        long long *d;
        cudaMalloc(&d, 148 * 2 * 8);
        for (int mode = 0; mode < 2; ++mode)
            for (int loads = 0; loads < 2; ++loads)
                for (int gap : {0, 8, 32, 128}) {
                    auto k = mode == 0 ? probe_traffic<0> : probe_traffic<1>;
                    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840);
                    k<<<148, 256, 163840>>>(128, 2048, gap, loads, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long h[2];
                    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    printf("traffic probe %s N=128, 6 warps of %s with %3d ALU ops between: %.1f cyc/MMA [%s]\n", mode ? "TS" : "SS", loads ? "ld.shared.v4" : "st.shared.v4",
                           gap, h[0] / 2048.0, cudaGetErrorString(e));
                }
    }
    {
        long long *d;
        cudaMalloc(&d, 148 * 2 * 8);
        cudaFuncSetAttribute(roundtrip, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
        for (int ep : {0, 1})
            for (int burst : {1, 8, 16}) {
                roundtrip<<<148, 160, 131072>>>(128, burst, 512, ep, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2];
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("roundtrip epilogue=%d burst=%2d MMAs (N=128): %.0f cycles per step (MMA floor %d) [%s]\n", ep, burst, h[0] / 512.0, burst * 64,
                       cudaGetErrorString(e));
            }
    }

    long long *d;
    cudaMalloc(&d, 148 * 2 * 8);
    const char *names[4] = {"SS no-swizzle", "TS + no-swizzle B", "SS 128B-swizzle", "TS + 128B-swizzle B"};
    const int iters = 2048;
    for (int grid : {1, 148})
        for (int mode = 0; mode < 4; ++mode)
            for (int n : {64, 128, 256})
                for (int ce : {0, 4}) {
                    auto k = mode == 0 ? (ce ? probe<0, 4> : probe<0, 0>) : mode == 1 ? (ce ? probe<1, 4> : probe<1, 0>)
                           : mode == 2 ? (ce ? probe<2, 4> : probe<2, 0>) : (ce ? probe<3, 4> : probe<3, 0>);
                    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
                    k<<<grid, 128, 131072>>>(n, iters, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) {
                        printf("mode %d n %d: %s\n", mode, n, cudaGetErrorString(e));
                        return 1;
                    }
                    long long h[296];
                    cudaMemcpy(h, d, grid * 16, cudaMemcpyDeviceToHost);
                    double issue = 0, exec = 0;
                    for (int b = 0; b < grid; ++b) {
                        issue += h[2 * b];
                        exec += h[2 * b + 1];
                    }
                    printf("grid %3d  %-20s N=%3d commit_every=%d : issue %.1f cyc/MMA, execute %.1f cyc/MMA (floor %d)\n", grid, names[mode], n, ce,
                           issue / grid / iters, exec / grid / iters, n / 2);
                }
    return 0;
}
