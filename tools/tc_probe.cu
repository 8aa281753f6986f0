// tc_probe.cu -- standalone hardware probe for the tcgen05 building blocks used by lookahead_tc.cu (sm_100a).
//
// Computes D[128 x N] = A[128 x K] * W[N x K]^T with the 3xTF32 split (hi*hi + lo*hi + hi*lo):
//   A (hi and lo parts) is written by the row-owner threads into TENSOR MEMORY with tcgen05.st (TS mode),
//   W hi / lo images sit in shared memory in the canonical K-major no-swizzle layout img[k/4][n][4] and arrive by
//   one cp.async.bulk each, the accumulator lives in TMEM and is read back with tcgen05.ld.
// Prints the max error against an FP64 reference for several (N, K).  Build: nvcc -arch=sm_100a -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../aemcarl_b200/csrc/tc_common.cuh"

using namespace aemtc;

__global__ void __launch_bounds__(128, 1) probe_kernel(const float *__restrict__ A, const float *__restrict__ Wimg,
                                                       float *__restrict__ D, int N, int K, int mode)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float *w_hi = reinterpret_cast<float *>(smem_raw);
    float *w_lo = w_hi + (size_t)N * K;
    __shared__ __align__(8) uint64_t bar_w, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const uint32_t img_bytes = (uint32_t)N * K * 4;
    if (tid == 0) {
        mbar_expect_tx(&bar_w, 2 * img_bytes);
        bulk_g2s(w_hi, Wimg, img_bytes, &bar_w);
        bulk_g2s(w_lo, Wimg + (size_t)N * K, img_bytes, &bar_w);
    }
    // row-owner thread writes its A row (hi -> cols [0,K), lo -> cols [K,2K)) into TMEM
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int row = tid;
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a = A[(size_t)row * K + k0 + j];
            split_tf32(a, hi[j], lo[j]);
        }
        tmem_st8(tmem + lane_base + k0, hi);
        tmem_st8(tmem + lane_base + K + k0, lo);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        mbar_wait(&bar_w, 0);
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint32_t d_col = 2 * K;                // accumulator columns after the A operand
        const uint32_t lbo = (uint32_t)N * 16, sbo = 128;
        int first = 1;
        for (int pass = 0; pass < (mode == 0 ? 3 : 1); ++pass) {
            const uint32_t a_col = (pass == 1) ? K : 0;                  // pass 1: A_lo
            const float *wimg = (pass == 2) ? w_lo : w_hi;               // pass 2: W_lo
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint64_t bdesc = make_smem_desc(smem_u32(wimg) + ks * 2 * lbo, lbo, sbo);
                mma_tf32_ts(tmem + d_col, tmem + a_col + ks * 8, bdesc, idesc, first ? 0u : 1u);
                first = 0;
            }
        }
        mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(tmem + lane_base + 2 * K + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static float tf32_rna(float x)
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x1000u;
    u &= 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

int main()
{
    int dev = 0;
    cudaSetDevice(dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    printf("device %s sm_%d%d\n", prop.name, prop.major, prop.minor);
    const int cfgs[][2] = {{64, 16}, {64, 64}, {112, 64}, {64, 104}, {160, 56}, {64, 152}, {112, 56}, {208, 64}};
    int bad = 0;
    for (auto &c : cfgs) {
        const int N = c[0], K = c[1];
        std::vector<float> A(128 * K), W(N * K), img(2 * N * K), D(128 * N);
        srand(N * 1000 + K);
        for (auto &x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
        for (auto &x : W) x = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.2f;
        // canonical K-major no-swizzle image: img[k/4][n][k%4]
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) {
                const float w = W[n * K + k];
                const float hi = tf32_rna(w);
                img[((size_t)(k / 4) * N + n) * 4 + (k % 4)] = hi;
                img[(size_t)N * K + ((size_t)(k / 4) * N + n) * 4 + (k % 4)] = w - hi;
            }
        float *dA, *dW, *dD;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, img.size() * 4); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dW, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
        for (int mode = 0; mode < 2; ++mode) {
            cudaMemset(dD, 0, D.size() * 4);
            const size_t smem = (size_t)2 * N * K * 4 + 1024;
            cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            probe_kernel<<<1, 128, smem>>>(dA, dW, dD, N, K, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("N=%d K=%d mode=%d CUDA error: %s\n", N, K, mode, cudaGetErrorString(e)); return 2; }
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0, maxref = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)W[n * K + k];
                    maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
                    maxref = fmax(maxref, fabs(ref));
                }
            printf("N=%3d K=%3d %s: max|err| = %.3e (max|ref| = %.3f)\n", N, K, mode == 0 ? "3xTF32" : "1xTF32", maxerr,
                   maxref);
            if (mode == 0 && maxerr > 2e-6 * (1 + maxref) * 4) bad++;
            if (mode == 1 && maxerr > 5e-3) bad++;
        }
        cudaFree(dA); cudaFree(dW); cudaFree(dD);
    }
    printf(bad ? "PROBE FAILED (%d)\n" : "PROBE OK\n", bad);
    return bad ? 1 : 0;
}
