// Probe: tcgen05.mma with the A operand in TMEM (kind::f16, bf16, M=128).  Verifies the TMEM layout assumed by
// heads_tc.cu: element (row m, k) of A lives in lane m, 32-bit column a_col + k/2, low half = even k; written with
// tcgen05.st.32x32b by the thread that owns lane m.   nvcc -gencode arch=compute_100a,code=sm_100a -o probe probe_ts_mma.cu
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3fffu) | (static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t instr_desc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

constexpr int N = 128, K = 64;

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16 *A /*[128][K]*/, const __nv_bfloat16 *Bimg /*UMMA image*/, float *D /*[128][N]*/) {
    __shared__ __align__(1024) unsigned char sB[N * K * 2];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < N * K * 2 / 16; i += 128) reinterpret_cast<uint4 *>(sB)[i] = reinterpret_cast<const uint4 *>(Bimg)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    // A -> TMEM columns [128, 128 + K/2): thread = row
    {
        uint32_t r[K / 2];
        for (int c = 0; c < K / 2; ++c) {
            const __nv_bfloat162 v = __halves2bfloat162(A[tid * K + 2 * c], A[tid * K + 2 * c + 1]);
            r[c] = *reinterpret_cast<const uint32_t *>(&v);
        }
        for (int c = 0; c < K / 2; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(t_lane + 128 + c), "r"(r[c]),
                         "r"(r[c + 1]), "r"(r[c + 2]), "r"(r[c + 3]), "r"(r[c + 4]), "r"(r[c + 5]), "r"(r[c + 6]), "r"(r[c + 7]));
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = instr_desc(N);
        for (int k = 0; k < K; k += 16) {
            const uint64_t bd = smem_desc(smem_u32(sB) + (k >> 3) * (N * 16), N * 16, 128);
            const uint32_t a_t = tmem + 128 + k / 2;
            const uint32_t acc = k > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem), "r"(a_t), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; c += 32) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(t_lane + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[tid * N + c + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

int main() {
    std::vector<__nv_bfloat16> A(128 * K), W(N * K), Bimg(N * K);
    srand(1);
    for (auto &v : A) v = __float2bfloat16((rand() % 17 - 8) / 8.0f);
    for (auto &v : W) v = __float2bfloat16((rand() % 13 - 6) / 4.0f);
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) Bimg[(k >> 3) * N * 8 + n * 8 + (k & 7)] = W[n * K + k];
    __nv_bfloat16 *dA, *dB;
    float *dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, Bimg.size() * 2); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bimg.data(), Bimg.size() * 2, cudaMemcpyHostToDevice);
    probe<<<1, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    std::vector<float> D(128 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float ref = 0;
            for (int k = 0; k < K; ++k) ref += __bfloat162float(A[m * K + k]) * __bfloat162float(W[n * K + k]);
            maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
        }
    printf("TS-MMA max abs err vs CPU: %g  (%s)\n", maxerr, maxerr < 1e-3 ? "LAYOUT OK" : "LAYOUT MISMATCH");
    return 0;
}
