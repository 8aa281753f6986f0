// mma_issue_probe.cu -- how long does ONE thread take to issue a run of tcgen05.mma (kind::f16, M = 128, K = 16,
// A operand in tensor memory, B operand through a no-swizzle shared-memory descriptor: the shapes lookahead_h16.cu
// uses), and when do they complete?  Answers whether the issuing thread is blocked for the duration of the MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_issue_probe tools/mma_issue_probe.cu
//   tools/bin/mma_issue_probe
// Prints, per (N, count): cycles from the first issue to the last issue returning, to the commit returning, and to
// the commit's mbarrier completing.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../aemcarl_b200/csrc/tc_common.cuh"
using namespace aemtc;

__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
        : "memory");
}

template <int N, int CNT, bool SS>
__device__ void run(uint32_t tm, uint32_t smem_b, uint64_t *bar, uint32_t &phase, long long *out)
{
    constexpr uint32_t LBO = (uint32_t)N * 16;
    constexpr uint64_t DHI = ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bd = DHI | (uint64_t)((smem_b & 0x3FFFFu) >> 4);
    const uint64_t ad = (((uint64_t)(2048 >> 4)) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46) |
                        (uint64_t)(((smem_b + 32768) & 0x3FFFFu) >> 4);
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
#pragma unroll
        for (int i = 0; i < CNT; ++i) {
            if (SS) mma_f16_ss(tm + 256, ad, bd, IDESC, i ? 1u : 0u);
            else mma_f16_ts(tm + 256, tm + 8 * (i & 7), bd, IDESC, i ? 1u : 0u);
        }
        t1 = clock64();
        mma_commit(bar);
        t2 = clock64();
    }
    __syncwarp();
    mbar_wait(bar, phase);
    phase ^= 1u;
    t3 = clock64();
    tc_fence_after();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = t3 - t0; }
}

__global__ void probe_kernel(long long *out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncwarp();
    tc_fence_after();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t tm = slot;
    const uint32_t sb = smem_u32(smem);
    uint32_t phase = 0;
    int k = 0;
#define RUN(N, C, SS) run<N, C, SS>(tm, sb, &bar, phase, out + 3 * k); ++k;
    // warm-up, then the sweep (each twice: the second is the warm number)
    RUN(64, 8, false)
    RUN(64, 1, false) RUN(64, 2, false) RUN(64, 4, false) RUN(64, 8, false) RUN(64, 16, false) RUN(64, 32, false)
    RUN(128, 1, false) RUN(128, 4, false) RUN(128, 12, false) RUN(128, 24, false)
    RUN(256, 1, false) RUN(256, 4, false) RUN(256, 12, false)
    RUN(64, 1, true) RUN(64, 8, true) RUN(64, 32, true)
    RUN(128, 12, true) RUN(256, 12, true)
    RUN(32, 32, false) RUN(16, 32, false)
    tc_fence_before();
    __syncwarp();
    tmem_dealloc(tm, 512);
}

int main()
{
    long long *out;
    cudaMallocManaged(&out, 4096);
    for (int i = 0; i < 512; ++i) out[i] = 0;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 36864);
    probe_kernel<<<1, 32, 65536 + 36864>>>(out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    const char *names[] = {"warm N64 x8", "N64 x1", "N64 x2", "N64 x4", "N64 x8", "N64 x16", "N64 x32",
                           "N128 x1", "N128 x4", "N128 x12", "N128 x24", "N256 x1", "N256 x4", "N256 x12",
                           "SS N64 x1", "SS N64 x8", "SS N64 x32", "SS N128 x12", "SS N256 x12", "N32 x32", "N16 x32"};
    for (int k = 0; k < 21; ++k)
        printf("%-14s issue returned %6lld   commit returned %6lld   complete %6lld cycles\n", names[k], out[3 * k],
               out[3 * k + 1], out[3 * k + 2]);
    return 0;
}
