// tc_probe_2cta.cu -- probe for tcgen05.mma.cta_group::2 (M = 256 across a CTA pair) in the configuration the lookahead
// kernels use: kind::f16, A operand in TENSOR MEMORY (each CTA's 128 rows, written with tcgen05.st), B operand in shared
// memory as the K-major no-swizzle image, FP32 accumulation.  Hypothesis under test: every CTA of the pair holds HALF of
// the B tile (N/2 columns, image img[k/8][N/2][8]) at the same shared-memory offset, the leader CTA issues the MMA and
// commits with a multicast arrive to both CTAs' mbarriers.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../aemcarl_b200/csrc/tc_common.cuh"
using namespace aemtc;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe2(const float *A, const __half *Wimg /* [2 ranks][K/8][N/2][8] */, float *D, int N, int K)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __half *w = reinterpret_cast<__half *>(smem_raw);
    __shared__ __align__(8) uint64_t bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_rank();
    if (tid == 0) { mbar_init(&bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    // this CTA's half of B: plain copies, then make them visible to the async proxy
    const int half_elems = (N / 2) * K;
    for (int i = tid; i < half_elems; i += 128) w[i] = Wimg[(size_t)rank * half_elems + i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int row = (int)rank * 128 + tid;
    for (int c0 = 0; c0 < K / 2; c0 += 8) {               // A: fp16 pairs, even k in the low half (single pass, no split)
        uint32_t v[8];
        for (int j = 0; j < 8; ++j) {
            const __half2 h = __floats2half2_rn(A[(size_t)row * K + 2 * (c0 + j)], A[(size_t)row * K + 2 * (c0 + j) + 1]);
            v[j] = *reinterpret_cast<const uint32_t *>(&h);
        }
        tmem_st8(tmem + lane_base + c0, v);
    }
    tmem_wait_st();
    tc_fence_before();
    cluster_sync();                                        // both CTAs: operands written, barriers initialised
    tc_fence_after();
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint32_t lbo = (uint32_t)(N / 2) * 16, sbo = 128;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t bdesc = make_smem_desc(smem_u32(w) + ks * 2 * lbo, lbo, sbo);
            const uint32_t acc = ks ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem + 64), "r"(tmem + ks * 8),
                         "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(&bar_mma)), "h"((uint16_t)3) : "memory");
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(tmem + lane_base + 64 + c0, v);
        tmem_wait_ld();
        for (int j = 0; j < 8; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

int main()
{
    const int cfgs[][2] = {{64, 64}, {32, 16}, {64, 32}};
    int bad = 0;
    for (auto &c : cfgs) {
        const int N = c[0], K = c[1];
        std::vector<float> A(256 * K), W(N * K), D(256 * N);
        std::vector<__half> img((size_t)N * K);
        srand(N * 100 + K);
        for (auto &x : A) x = __half2float(__float2half_rn((float)rand() / RAND_MAX * 2.f - 1.f));
        for (auto &x : W) x = __half2float(__float2half_rn(((float)rand() / RAND_MAX * 2.f - 1.f) * 0.25f));
        const int H = N / 2;
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) {
                const int r = n / H, nn = n % H;                  // rank r holds columns [r H, (r + 1) H)
                img[(size_t)r * H * K + ((size_t)(k / 8) * H + nn) * 8 + (k % 8)] = __float2half_rn(W[n * K + k]);
            }
        float *dA, *dD; __half *dW;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, img.size() * 2); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dW, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, D.size() * 4);
        const size_t smem = (size_t)H * K * 2 + 1024;
        cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe2<<<2, 128, smem>>>(dA, dW, dD, N, K);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0; int worst_m = 0, worst_n = 0;
        for (int m = 0; m < 256; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)W[n * K + k];
            const double err = fabs(ref - D[m * N + n]);
            if (err > maxerr) { maxerr = err; worst_m = m; worst_n = n; }
            maxref = fmax(maxref, fabs(ref));
        }
        printf("2-CTA TS-mode N=%3d K=%3d: max|err| = %.3e at (%d, %d) (max|ref| = %.3f)  D[0][0..3] = %.4f %.4f %.4f %.4f  D[128][0] = %.4f\n",
               N, K, maxerr, worst_m, worst_n, maxref, D[0], D[1], D[2], D[3], D[128 * N]);
        if (maxerr > 1e-3) bad++;
        cudaFree(dA); cudaFree(dW); cudaFree(dD);
    }
    printf(bad ? "PROBE2 FAILED (%d)\n" : "PROBE2 OK\n", bad);
    return 0;
}
