// tmem_ld_probe.cu -- tensor-memory read / write throughput as the epilogues of lookahead_h16.cu see it: W warps of one CTA
// (and optionally a second CTA on the same SM) each read their 32 lanes x C columns with tcgen05.ld.32x32b.xN, R times.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_ld_probe tools/tmem_ld_probe.cu
// Prints bytes per cycle per SM for each shape / warp count.
#include <cstdio>
#include <cuda_runtime.h>
#include "../aemcarl_b200/csrc/tc_common.cuh"
using namespace aemtc;

template <int N>
__device__ __forceinline__ void ldx(uint32_t addr, uint32_t &acc)
{
    if constexpr (N == 16) {
        uint32_t v[16];
        tmem_ld16(addr, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= v[i];
    } else if constexpr (N == 32) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(addr)
            : "memory");
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
    } else {
        uint32_t v[8];
        tmem_ld8(addr, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) acc ^= v[i];
    }
}

// mode 0: loads only; mode 1: load + store back (x8 stores of half the columns, like the in-place conversion)
template <int N>
__global__ void probe(long long *out, int reps, int cols, int mode, int active_warps)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    // initialise the columns
    for (int c = 0; c < 256; c += 8) {
        uint32_t z[8] = {1, 2, 3, 4, 5, 6, 7, 8};
        if (warp < 4) tmem_st8(tm + c, z);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const long long t0 = clock64();
    if (warp < active_warps) {
        const int base = (warp >> 2) * (cols);       // warps 4..7 read the second block of columns (like the thread halves)
        for (int r = 0; r < reps; ++r) {
            for (int c = 0; c < cols; c += N) {
                ldx<N>(tm + base + c, acc);
                if (mode == 1) {
                    uint32_t z[8] = {acc, acc, acc, acc, acc, acc, acc, acc};
                    tmem_st8(tm + base + c, z);
                    if (N >= 16) tmem_st8(tm + base + c + 8, z);
                }
            }
        }
        if (mode == 1) tmem_wait_st();
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) out[100] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 256);
}

template <int N>
void run(long long *out, int warps, int cols, int mode, int ctas)
{
    const int reps = 64;
    // ctas > 1: force co-residency on one SM is not controllable; launch 2 x num_sms CTAs and report the slowest
    int dev_sms = 148;
    const int grid = ctas == 1 ? 1 : 2 * dev_sms;
    probe<N><<<grid, 256, ctas == 1 ? 0 : 100 * 1024>>>(out, reps, cols, mode, warps);
    cudaDeviceSynchronize();
    long long worst = 0;
    for (int i = 0; i < grid; ++i) worst = out[i] > worst ? out[i] : worst;
    const double bytes = (double)warps * 32 * cols * 4 * reps * (ctas == 1 ? 1 : 2);
    printf("x%-2d warps %d cols/warp %3d %s %s: %7lld cycles, %6.1f B/cycle/SM read\n", N, warps, cols,
           mode ? "ld+st" : "ld   ", ctas == 1 ? "1 CTA      " : "2 CTAs / SM", worst, bytes / (double)worst);
}

int main()
{
    long long *out;
    cudaMallocManaged(&out, 8192);
    cudaFuncSetAttribute(probe<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(probe<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(probe<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    run<16>(out, 8, 64, 0, 1);      // warm-up
    for (int warps : {1, 4, 8}) {
        run<8>(out, warps, 64, 0, 1);
        run<16>(out, warps, 64, 0, 1);
        run<32>(out, warps, 64, 0, 1);
    }
    run<16>(out, 8, 64, 1, 1);
    run<32>(out, 8, 64, 1, 1);
    run<16>(out, 8, 64, 0, 2);
    run<32>(out, 8, 64, 0, 2);
    run<16>(out, 8, 64, 1, 2);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
