// tc_probe16.cu -- probe for kind::f16 TS-mode MMA with FP16 hi/lo split (3 passes), A packed 2 x fp16 per TMEM column.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../aemcarl_b200/csrc/tc_common.cuh"
using namespace aemtc;

__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc),
                 "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t pack_hi_lo(float a, float b, uint32_t &lo_out)
{
    __half ah = __float2half_rn(a), bh = __float2half_rn(b);
    __half al = __float2half_rn(a - __half2float(ah)), bl = __float2half_rn(b - __half2float(bh));
    lo_out = (uint32_t)__half_as_ushort(al) | ((uint32_t)__half_as_ushort(bl) << 16);
    return (uint32_t)__half_as_ushort(ah) | ((uint32_t)__half_as_ushort(bh) << 16);
}

// mode 0: 3 passes; mode 1: hi*hi only.  order: which half of the 32-bit column holds the even k (0 = low half)
__global__ void __launch_bounds__(128, 1) probe16(const float *A, const __half *Wimg, float *D, int N, int K, int mode, int order)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __half *w_hi = reinterpret_cast<__half *>(smem_raw);
    __half *w_lo = w_hi + (size_t)N * K;
    __shared__ __align__(8) uint64_t bar_w, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar_w, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t img_bytes = (uint32_t)N * K * 2;
    if (tid == 0) { mbar_expect_tx(&bar_w, 2 * img_bytes); bulk_g2s(w_hi, Wimg, img_bytes, &bar_w); bulk_g2s(w_lo, Wimg + (size_t)N * K, img_bytes, &bar_w); }
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int row = tid, KC = K / 2;       // columns per part
    for (int c0 = 0; c0 < KC; c0 += 8) {
        uint32_t hi[8], lo[8];
        for (int j = 0; j < 8; ++j) {
            float a = A[(size_t)row * K + 2 * (c0 + j)], b = A[(size_t)row * K + 2 * (c0 + j) + 1];
            if (order) { float t = a; a = b; b = t; }
            hi[j] = pack_hi_lo(a, b, lo[j]);
        }
        tmem_st8(tmem + lane_base + c0, hi);
        tmem_st8(tmem + lane_base + KC + c0, lo);
    }
    tmem_wait_st(); tc_fence_before(); __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        mbar_wait(&bar_w, 0);
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t d_col = 2 * KC, lbo = (uint32_t)N * 16, sbo = 128;
        int first = 1;
        for (int pass = 0; pass < (mode == 0 ? 3 : 1); ++pass) {
            const uint32_t a_col = (pass == 1) ? KC : 0;
            const __half *wimg = (pass == 2) ? w_lo : w_hi;
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint64_t bdesc = make_smem_desc(smem_u32(wimg) + ks * 2 * lbo, lbo, sbo);
                mma_f16_ts(tmem + d_col, tmem + a_col + ks * 8, bdesc, idesc, first ? 0u : 1u);
                first = 0;
            }
        }
        mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0); tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(tmem + lane_base + 2 * KC + c0, v); tmem_wait_ld();
        for (int j = 0; j < 8; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main()
{
    const int cfgs[][2] = {{64, 16}, {64, 64}, {112, 64}, {64, 112}, {192, 64}, {80, 64}, {64, 80}};
    int bad = 0;
    for (auto &c : cfgs) {
        const int N = c[0], K = c[1];
        std::vector<float> A(128 * K), W(N * K), D(128 * N);
        std::vector<__half> img(2 * N * K);
        srand(N * 1000 + K);
        for (int i = 0; i < 128 * K; ++i) { float s = (i % 7 == 0) ? 8.f : ((i % 5 == 0) ? 0.01f : 1.f); A[i] = ((float)rand() / RAND_MAX * 2.f - 1.f) * s; }
        for (auto &x : W) x = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.15f;
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) {
                const float w = W[n * K + k];
                const __half hi = __float2half_rn(w);
                const __half lo = __float2half_rn(w - __half2float(hi));
                const size_t at = ((size_t)(k / 8) * N + n) * 8 + (k % 8);       // img[k/8][n][k%8]
                img[at] = hi; img[(size_t)N * K + at] = lo;
            }
        float *dA, *dD; __half *dW;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, img.size() * 2); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dW, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
        for (int order = 0; order < 2; ++order)
        for (int mode = 0; mode < 2; ++mode) {
            cudaMemset(dD, 0, D.size() * 4);
            const size_t smem = (size_t)2 * N * K * 2 + 1024;
            cudaFuncSetAttribute(probe16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            probe16<<<1, 128, smem>>>(dA, dW, dD, N, K, mode, order);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0, maxref = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
                double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)W[n * K + k];
                maxerr = fmax(maxerr, fabs(ref - D[m * N + n])); maxref = fmax(maxref, fabs(ref));
            }
            printf("N=%3d K=%3d order=%d %s: max|err| = %.3e (max|ref| = %.3f)\n", N, K, order, mode == 0 ? "3xFP16" : "1xFP16", maxerr, maxref);
            if (order == 0 && mode == 0 && maxerr > 3e-5) bad++;
        }
        cudaFree(dA); cudaFree(dW); cudaFree(dD);
    }
    printf(bad ? "PROBE16 FAILED (%d)\n" : "PROBE16 OK\n", bad);
    return 0;
}
