// bulk_latency_probe.cu -- how long does a cp.async.bulk of one weight chunk (12 KB, L2-resident) take from issue to the
// mbarrier completing, alone and with every SM streaming the same images (the situation of lookahead_h16.cu's weight ring:
// a group's four chunks are requested when the previous group's MMAs have completed).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/bulk_latency_probe tools/bulk_latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../aemcarl_b200/csrc/tc_common.cuh"
using namespace aemtc;

__global__ void probe(const uint8_t *src, long long *out, int chunks, int bytes, int reps)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[4];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r) {
        long long t0 = 0;
        if (threadIdx.x == 0) {
            t0 = clock64();
            for (int c = 0; c < chunks; ++c) {
                mbar_expect_tx(&bar[c], (uint32_t)bytes);
                bulk_g2s(smem + (size_t)c * bytes, src + ((size_t)(r % 8) * 4 + c) * bytes, (uint32_t)bytes, &bar[c]);
            }
        }
        long long first = 0, last = 0;
        for (int c = 0; c < chunks; ++c) {
            mbar_wait(&bar[c], phase);
            if (c == 0) first = clock64();
        }
        last = clock64();
        phase ^= 1u;
        if (threadIdx.x == 0 && blockIdx.x == 0) {
            out[2 * r] = first - t0;
            out[2 * r + 1] = last - t0;
        }
        __syncthreads();
    }
}

int main()
{
    uint8_t *src;
    long long *out;
    cudaMalloc(&src, 32 * 12288);
    cudaMemset(src, 1, 32 * 12288);
    cudaMallocManaged(&out, 4096);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 12288);
    for (int grid : {1, 296}) {
        for (int chunks : {1, 4}) {
            probe<<<grid, 32, 4 * 12288>>>(src, out, chunks, 12288, 12);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
            printf("grid %3d, %d x 12 KB per request: first chunk / all chunks after", grid, chunks);
            for (int r = 0; r < 12; ++r) printf(" %lld/%lld", out[2 * r], out[2 * r + 1]);
            printf(" cycles\n");
        }
    }
    return 0;
}
