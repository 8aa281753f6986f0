// mma_rate.cu -- timing probe: how long do R back-to-back tcgen05.mma (kind::f16, M = 128, K = 16) take for a given N,
// operand source (A in TMEM / shared) and B shared-memory layout (no swizzle with different LBO/SBO, 128B swizzle)?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../aemcarl_b200/csrc/tc_common.cuh"
using namespace aemtc;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// layout: 0 = no swizzle img[k/8][n][8] (LBO = N*16, SBO = 128); 1 = no swizzle img[n/8][k/8][8][8] (LBO = 128, SBO = KB*128/8..)
//         2 = 128B swizzle (rows of 64 halves, 8-row atoms of 1024 B)
__global__ void __launch_bounds__(512, 1) rate(int N, int R, int ts, int layout, int pollers, long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    {   // zero the A operand columns
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (warp < 4) for (int c = 0; c < 64; c += 8) tmem_st8(tm + ((uint32_t)(warp * 32) << 16) + c, z);
        tmem_wait_st();
    }
    tc_fence_before(); __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint64_t bd, kstep;
        const uint32_t sb = smem_u32(smem);
        if (layout == 0) { bd = make_smem_desc(sb, (uint32_t)N * 16, 128); kstep = (uint64_t)((2u * N * 16) >> 4); }
        else if (layout == 1) { bd = make_smem_desc(sb, 128, 8 * 128); kstep = (uint64_t)(256 >> 4); }   // K = 64 per image
        else { bd = ((uint64_t)((sb & 0x3FFFFu) >> 4)) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61); kstep = 2; }
        const uint64_t ad = make_smem_desc(sb + 32768, 128 * 16, 128);      // A (SS mode): 128 x 64 no swizzle
        const long long t0 = clock64();
        for (int r = 0; r < R; ++r) {
            const int ks = r & 3;
            const uint32_t d = tm + 64;
            if (ts) mma_ts(d, tm + ks * 8, bd + ks * kstep, idesc, r ? 1u : 0u);
            else mma_ss(d, ad + ks * ((2u * 128 * 16) >> 4), bd + ks * kstep, idesc, r ? 1u : 0u);
        }
        const long long t1 = clock64();
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else if (pollers == 1 || (pollers == 2 && (tid & 31) == 0)) {
        mbar_wait(&bar, 0);       // every other thread (or one lane per warp) polls the completion barrier, as the kernels do
    }
    __syncthreads();
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main()
{
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char *lname[] = {"noswz[k/8][n][8]", "noswz[n/8][k/8][8][8]", "swizzle128B"};
    for (int grid : {148})
        for (int ts = 1; ts >= 1; --ts)
            for (int layout = 0; layout < 1; ++layout)
              for (int pollers = 0; pollers < 3; ++pollers)
              for (int nthr : {128, 512})
                for (int N : {64, 256})
                    for (int R : {12, 48}) {
                        printf("threads %d pollers %d ", nthr, pollers);
                        for (int rep = 0; rep < 2; ++rep) rate<<<grid, nthr, 64 * 1024>>>(N, R, ts, layout, pollers, d);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
                        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                        printf("grid %3d %s %-22s N=%3d R=%2d: issue %6lld total %6lld  (%.1f cyc/MMA, floor %d)\n", grid,
                               ts ? "TS" : "SS", lname[layout], N, R, h[0], h[1], (double)h[1] / R, N / 2);
                    }
    return 0;
}
