// orca.cu -- per-agent ORCA half-plane linear program, one warp per human, sm_100a.
//
// The algorithm lives in orca_device.cuh (shared with the fused env_prepare kernel of env.cu); this file is the
// stand-alone kernel behind aem_orca.  Compiled with -fmad=false like every user of that header.
#include "orca_device.cuh"

namespace aem {

constexpr int ORCA_WARPS = 4;

__global__ void __launch_bounds__(ORCA_WARPS * 32) orca_kernel(int B, int n, const double *__restrict__ humans,
                                                               const double *__restrict__ robot, float time_step,
                                                               double safety, int robot_visible,
                                                               const int32_t *__restrict__ nact,
                                                               double *__restrict__ new_vel)
{
    __shared__ Line s_lines[ORCA_WARPS][32];
    __shared__ Line s_proj[ORCA_WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long agent = (long long)blockIdx.x * ORCA_WARPS + wib;
    if (agent >= (long long)B * n) return;
    const int b = (int)(agent / n), i = (int)(agent - (long long)b * n);
    // per-environment human count (the 'mixed' scenario rule, crowd_sim.py:337-386): rows >= na are padding
    const int na = nact ? min(max(nact[b], 0), n) : n;
    if (i >= na) {
        if (lane == 0) { new_vel[agent * 2] = 0.0; new_vel[agent * 2 + 1] = 0.0; }
        return;
    }
    const V2 res = orca_agent(humans + (size_t)b * n * 8, robot + (size_t)b * 9, i, na, time_step, safety, robot_visible,
                              s_lines[wib], s_proj[wib], lane);
    if (lane == 0) {
        new_vel[agent * 2] = (double)res.x;
        new_vel[agent * 2 + 1] = (double)res.y;
    }
}

cudaError_t launch_orca(int B, int n, const double *humans, const double *robot, const aem_env_params &p,
                        double *new_vel, cudaStream_t st, const int32_t *nact)
{
    const long long agents = (long long)B * n;
    if (agents <= 0) return cudaSuccess;
    const int blocks = (int)((agents + ORCA_WARPS - 1) / ORCA_WARPS);
    orca_kernel<<<blocks, ORCA_WARPS * 32, 0, st>>>(B, n, humans, robot, (float)p.time_step, p.orca_safety_space,
                                                    p.robot_visible, nact, new_vel);
    return cudaGetLastError();
}

}  // namespace aem
