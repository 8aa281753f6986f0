// env.cu -- batched CrowdSim.step pieces around the ORCA solve, FP64, sm_100a.
//
//   env_lookahead_kernel : crowd_sim/envs/crowd_sim.py:557-651,676-680 -- step(update=False) for EVERY action of the
//                          action space of one environment in one CTA: collision test along the relative segment
//                          (:570-591, utils.py:4-26), adaptive occupancy-probability reward (gridmap, :41-219),
//                          goal / timeout tests and the reward / info priority (:622-651), and the humans' next
//                          observable states (agent.py:63-74).
//   env_apply_kernel     : crowd_sim.py:653-669 -- step(update=True) for the selected action.
//
// Occupancy reward without the 600x600 grid: the reference rasterises, per human, a normalised probability blob over
// the disc of radius |v|*human_timestep and sums the grid over the robot disc of each candidate action.  Here each
// blob that can reach any candidate disc is rasterised once into a shared-memory tile (<= (2m/res+1)^2 doubles),
// normalised cell by cell exactly like crowd_sim.py:108-118, and every action sums tile /\ disc.  Index arithmetic
// (round-half-even, truncating LUT index, exclusive upper bounds, lower clamp at 0) follows the reference bit for bit.
// DEVIATION: blob cells whose index is outside [0, 600) are dropped (the reference wraps negative indices and is
// out-of-bounds above 599).
#include "orca_device.cuh"     // env_prepare = ORCA + env lookahead in one launch (this file is built with -fmad=false)

namespace aem {

constexpr double GRID_RES = 0.02;
constexpr int GRID_W = 600;
constexpr double GRID_OFF = 6.0;
constexpr double PI_D = 3.141592653589793;
constexpr int ENV_THREADS = 256;

__device__ __forceinline__ double lut_lookup(const double *__restrict__ lut, double raw)
{
    const double t = raw * 1000.0 + 5000.0;          // crowd_sim.py:32-38, int() truncates toward zero
    if (!(t > -1.0e9 && t < 1.0e9)) return 0.0;
    const int index = (int)t;
    if (index > 9999 || index < 0) return 0.0;
    return __ldg(lut + index);
}

struct Blob {
    int cx, cy, xlo, xhi, ylo, yhi, rw, rh;
    double m, heading;
};

__device__ __forceinline__ Blob blob_setup(double px, double py, double vx, double vy, double hts)
{
    Blob b;
    const double p_x = px + GRID_OFF, p_y = py + GRID_OFF;
    b.cx = (int)rint(p_x / GRID_RES);
    b.cy = (int)rint(p_y / GRID_RES);
    const double dx = vx * hts, dy = vy * hts;
    b.m = hypot(dx, dy);
    b.xlo = (int)rint((p_x - b.m) / GRID_RES);
    b.xhi = (int)rint((p_x + b.m) / GRID_RES);
    b.ylo = (int)rint((p_y - b.m) / GRID_RES);
    b.yhi = (int)rint((p_y + b.m) / GRID_RES);
    b.rw = b.xhi - b.xlo;
    b.rh = b.yhi - b.ylo;
    b.heading = atan2(dy, dx);
    return b;
}

__device__ __forceinline__ double int_hypot(int dx, int dy)
{
    return sqrt((double)(dx * dx + dy * dy));         // exact integer sum, correctly rounded sqrt
}

__device__ __forceinline__ double blob_weight(const Blob &b, int x, int y, double pos_var, double dir_var,
                                              const double *__restrict__ lut)
{
    double angle;
    if ((y - b.cy) == 0 && (x - b.cx) == 0) angle = b.heading;
    else angle = atan2((double)(y - b.cy), (double)(x - b.cx));
    double delta = angle - b.heading;
    if (delta <= -PI_D) delta += 2.0 * PI_D;
    else if (delta >= PI_D) delta = 2.0 * PI_D - delta;
    const double x0 = pos_var * (double)(y - b.cy) / ((double)b.rh + 0.000001);
    const double x1 = pos_var * (double)(x - b.cx) / ((double)b.rw + 0.000001);
    const double x2 = dir_var * delta / PI_D;
    return lut_lookup(lut, x0) * lut_lookup(lut, x1) * lut_lookup(lut, x2);
}

// crowd_sim/envs/utils/utils.py:4-26 with (x3, y3) = origin
__device__ __forceinline__ double point_to_segment_dist0(double x1, double y1, double x2, double y2)
{
    const double px = x2 - x1, py = y2 - y1;
    if (px == 0 && py == 0) return sqrt((0 - x1) * (0 - x1) + (0 - y1) * (0 - y1));
    double u = ((0 - x1) * px + (0 - y1) * py) / (px * px + py * py);
    if (u > 1) u = 1; else if (u < 0) u = 0;
    const double x = x1 + u * px, y = y1 + u * py;
    return sqrt((x - 0) * (x - 0) + (y - 0) * (y - 0));
}

__device__ __forceinline__ double block_sum(double v, double *red)
{
    // deterministic: warp xor-tree, then the 8 warp sums in order
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < ENV_THREADS / 32; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(ENV_THREADS) env_lookahead_kernel(
    int B, int n, int A, const double *__restrict__ robot, const double *__restrict__ humans,
    const double *__restrict__ new_vel, const double *__restrict__ global_time, const double *__restrict__ actions_all,
    int action_env_stride, const double *__restrict__ lut, const aem_env_params p, double amax, int tile_cap,
    double *__restrict__ rewards, int32_t *__restrict__ info, double *__restrict__ dmin_out,
    double *__restrict__ humans_next, const int32_t *__restrict__ nact, double *new_vel_out, int robot_visible)
{
    // fused mode (new_vel_out != nullptr, aem_env_prepare): the CTA first solves ORCA for its own humans, one warp per
    // human exactly like orca_kernel, and writes the new velocities; `new_vel` is not read
    __shared__ int s_near[AEM_MAX_HUMANS];
    __shared__ Blob s_blob[AEM_MAX_HUMANS];
    extern __shared__ __align__(16) double dsm[];
    double *tile = dsm;                         // [tile_cap]
    double *partial = tile + tile_cap;          // [ENV_THREADS]
    double *red = partial + ENV_THREADS;        // [8]
    double *prob = red + 8;                     // [AEM_MAX_ACTIONS]
    double *sh = prob + AEM_MAX_ACTIONS;        // [n][8] humans
    double *sr9 = sh + AEM_MAX_HUMANS * 8;      // [9] robot
    // ORCA's half-plane lists (8 warps x 32 lines x 2) live in the tile, which nobody touches before the ORCA phase is over
    Line(*s_lines)[32] = reinterpret_cast<Line(*)[32]>(tile);
    Line(*s_proj)[32] = s_lines + ENV_THREADS / 32;

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int nstride = n;                      // row pitch of the human arrays
    if (nact) n = min(max(nact[b], 0), n);      // per-environment human count ('mixed' rule): rows >= n are padding
    // shared action table (stride 0) or one private table per environment (arbitrary ActionXY per env, stride 2 A)
    const double *__restrict__ actions = actions_all + (size_t)b * action_env_stride;
    for (int i = tid; i < n * 8; i += ENV_THREADS) sh[i] = humans[(size_t)b * nstride * 8 + i];
    if (tid < 9) sr9[tid] = robot[(size_t)b * 9 + tid];
    if (tid < AEM_MAX_ACTIONS) prob[tid] = 0.0;
    if (new_vel_out) {
        const int lane = tid & 31, wib = tid >> 5;
        for (int i = wib; i < nstride; i += ENV_THREADS / 32) {
            double *o = new_vel_out + ((size_t)b * nstride + i) * 2;
            if (i >= n) {
                if (lane == 0) { o[0] = 0.0; o[1] = 0.0; }
                continue;
            }
            const V2 res = orca_agent(humans + (size_t)b * nstride * 8, robot + (size_t)b * 9, i, n, (float)p.time_step,
                                      p.orca_safety_space, robot_visible, s_lines[wib], s_proj[wib], lane);
            if (lane == 0) { o[0] = (double)res.x; o[1] = (double)res.y; }
            __syncwarp();
        }
    }
    __syncthreads();
    const double *nv = new_vel_out ? new_vel_out : new_vel;
    const double rx = sr9[0], ry = sr9[1], rrad = sr9[4];
    const double search_radius = rrad + (n > 0 ? sh[4] : 0.0);   // crowd_sim.py:169-173
    const int TPA = ENV_THREADS / A;                              // threads per action (>= 2 for A <= 128)

    // The disc test of crowd_sim.py:176-178, sqrt(dx^2 + dy^2) * res > search_radius, is monotone in the integer dx^2 + dy^2
    // (correctly rounded sqrt, multiplication by a positive constant): find the largest integer that still fails it with
    // the SAME FP64 expression once, then every cell of every candidate disc is an integer comparison.
    int disc_s;
    {
        const double q = search_radius / GRID_RES;
        const double q2 = q * q;
        int s0 = (q2 >= 0.0 && q2 < 720000.0) ? (int)q2 : (q2 >= 720000.0 ? 720000 : 0);
        if (search_radius != search_radius) s0 = 720001;        // NaN radius: the comparison is false for every cell
        while (s0 < 720001 && !(sqrt((double)(s0 + 1)) * GRID_RES > search_radius)) ++s0;
        while (s0 >= 0 && sqrt((double)s0) * GRID_RES > search_radius) --s0;
        disc_s = s0;                                            // (2 * 600^2 = 720000 bounds dx^2 + dy^2 on the grid)
    }
    // which humans' blobs can reach a candidate disc at all: one thread per human (two FP64 hypot each -- a quarter of the
    // kernel's instructions when all 256 threads evaluated the test for every human)
    if (tid < n) {
        const double hpx = sh[tid * 8], hpy = sh[tid * 8 + 1], hvx = sh[tid * 8 + 2], hvy = sh[tid * 8 + 3];
        const double m0 = hypot(hvx * p.human_timestep, hvy * p.human_timestep);
        s_near[tid] = !(hypot(hpx - rx, hpy - ry) > m0 + search_radius + amax + 0.1);   // else: cannot touch any disc (exact)
        if (s_near[tid]) s_blob[tid] = blob_setup(hpx, hpy, hvx, hvy, p.human_timestep);  // (FP64 hypot + atan2: once, not x 256)
    }
    __syncthreads();
    for (int h = 0; h < n; ++h) {
        if (!s_near[h]) continue;
        const Blob bl = s_blob[h];
        const int cells = bl.rw * bl.rh;
        if (cells <= 0) continue;
        if (cells > tile_cap) {               // faster human than the host sized the tile for: flag with NaN
            if (tid < A) prob[tid] = nan("");
            continue;
        }
        // rasterise the blob (crowd_sim.py:84-106) and its normaliser
        double local = 0.0;
        for (int idx = tid; idx < cells; idx += ENV_THREADS) {
            const int yy = idx / bl.rw, xx = idx - yy * bl.rw;
            const int x = bl.xlo + xx, y = bl.ylo + yy;
            double w = 0.0;
            if (!(int_hypot(x - bl.cx, y - bl.cy) * GRID_RES > bl.m))
                w = blob_weight(bl, x, y, p.position_variance, p.direction_variance, lut);
            tile[idx] = w;
            local += w;
        }
        const double S = block_sum(local, red);
        if (S == 0.0) { __syncthreads(); continue; }      // crowd_sim.py:113-116
        for (int idx = tid; idx < cells; idx += ENV_THREADS) tile[idx] = tile[idx] / S;
        __syncthreads();
        // every action sums tile /\ disc (crowd_sim.py:122-132, 159-185)
        double psum = 0.0;
        if (tid < A * TPA) {
            const int a = tid / TPA, j = tid - a * TPA;
            const double agent_x = rx + actions[2 * a] * p.agent_timestep + GRID_OFF;
            const double agent_y = ry + actions[2 * a + 1] * p.agent_timestep + GRID_OFF;
            const int acx = (int)rint(agent_x / GRID_RES), acy = (int)rint(agent_y / GRID_RES);
            const double xlr = (agent_x - search_radius) < 0 ? 0.0 : (agent_x - search_radius);
            const double ylr = (agent_y - search_radius) < 0 ? 0.0 : (agent_y - search_radius);
            int x0 = (int)rint(xlr / GRID_RES), x1 = (int)rint((agent_x + search_radius) / GRID_RES);
            int y0 = (int)rint(ylr / GRID_RES), y1 = (int)rint((agent_y + search_radius) / GRID_RES);
            x0 = max(max(x0, bl.xlo), 0); x1 = min(min(x1, bl.xhi), GRID_W);
            y0 = max(max(y0, bl.ylo), 0); y1 = min(min(y1, bl.yhi), GRID_W);
            for (int y = y0 + j; y < y1; y += TPA) {
                // cells of this row inside the disc: (x - acx)^2 + dy^2 <= disc_s  (== !(int_hypot(..) * GRID_RES > search_radius)),
                // i.e. the chord |x - acx| <= floor(sqrt(disc_s - dy^2)): walk the chord only, in the reference's order
                const int rem = disc_s - (y - acy) * (y - acy);
                if (rem < 0) continue;
                int w = (int)sqrt((double)rem);
                while (w * w > rem) --w;
                while ((w + 1) * (w + 1) <= rem) ++w;
                const int xl = max(x0, acx - w), xh = min(x1 - 1, acx + w);
                const double *trow = tile + (y - bl.ylo) * bl.rw - bl.xlo;
                for (int x = xl; x <= xh; ++x) psum += trow[x];
            }
        }
        partial[tid] = psum;
        __syncthreads();
        if (tid < A) {
            double s = 0.0;
            for (int j = 0; j < TPA; ++j) s += partial[tid * TPA + j];
            prob[tid] += s;
        }
        __syncthreads();
    }
    __syncthreads();

    if (tid < A) {
        const int a = tid;
        const double ax = actions[2 * a], ay = actions[2 * a + 1];
        double dmin = INFINITY;
        bool collision = false;
        for (int i = 0; i < n; ++i) {                      // crowd_sim.py:570-591
            const double px = sh[i * 8] - rx, py = sh[i * 8 + 1] - ry;
            const double vx = sh[i * 8 + 2] - ax, vy = sh[i * 8 + 3] - ay;
            const double ex = px + vx * p.time_step, ey = py + vy * p.time_step;
            const double cd = point_to_segment_dist0(px, py, ex, ey) - sh[i * 8 + 4] - rrad;
            if (cd < 0) { collision = true; break; }
            else if (cd < dmin) dmin = cd;
        }
        const double gx = rx + ax * p.time_step - sr9[5], gy = ry + ay * p.time_step - sr9[6];
        const bool reaching_goal = sqrt(gx * gx + gy * gy) < rrad;      // crowd_sim.py:622-623
        double reward;
        int code;
        if (global_time[b] >= p.time_limit - 1) { reward = 0; code = AEM_INFO_TIMEOUT; }
        else if (collision) { reward = p.collision_penalty; code = AEM_INFO_COLLISION; }
        else if (reaching_goal) { reward = p.success_reward; code = AEM_INFO_REACHGOAL; }
        else {
            reward = -0.25 * prob[a] * p.reward_increment;
            code = (dmin < p.discomfort_dist) ? AEM_INFO_DANGER : AEM_INFO_NOTHING;
        }
        rewards[(size_t)b * A + a] = reward;
        if (info) info[(size_t)b * A + a] = code;
        if (dmin_out) dmin_out[(size_t)b * A + a] = dmin;
    }
    if (humans_next && tid >= n && tid < nstride) {        // padding rows: defined values
        double *o = humans_next + ((size_t)b * nstride + tid) * 5;
        o[0] = o[1] = o[2] = o[3] = o[4] = 0.0;
    }
    if (humans_next && tid < n) {                          // agent.py:63-74
        const double vx = nv[((size_t)b * nstride + tid) * 2], vy = nv[((size_t)b * nstride + tid) * 2 + 1];
        double *o = humans_next + ((size_t)b * nstride + tid) * 5;
        o[0] = sh[tid * 8] + vx * p.time_step;
        o[1] = sh[tid * 8 + 1] + vy * p.time_step;
        o[2] = vx; o[3] = vy; o[4] = sh[tid * 8 + 4];
    }
}

__global__ void env_apply_kernel(int B, int n, int A, double *__restrict__ robot, double *__restrict__ humans,
                                 const double *__restrict__ new_vel, double *__restrict__ global_time,
                                 const double *__restrict__ actions_all, int action_env_stride,
                                 const int32_t *__restrict__ chosen,
                                 const double *__restrict__ rewards, const int32_t *__restrict__ info,
                                 const double *__restrict__ dmin, double time_step, double *__restrict__ out_reward,
                                 int32_t *__restrict__ out_done, int32_t *__restrict__ out_info,
                                 double *__restrict__ out_dmin, const int32_t *__restrict__ nact)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int na = nact ? min(max(nact[b], 0), n) : n;
    int a = chosen ? chosen[b] : 0;
    if (a < 0 || a >= A) a = 0;
    const double *__restrict__ actions = actions_all + (size_t)b * action_env_stride;
    const double ax = actions[2 * a], ay = actions[2 * a + 1];
    if (rewards && out_reward) out_reward[b] = rewards[(size_t)b * A + a];
    if (info) {
        const int code = info[(size_t)b * A + a];
        if (out_info) out_info[b] = code;
        if (out_done) out_done[b] = (code >= AEM_INFO_REACHGOAL) ? 1 : 0;
    }
    if (dmin && out_dmin) out_dmin[b] = dmin[(size_t)b * A + a];
    double *rb = robot + (size_t)b * 9;                    // agent.py:122-131
    rb[0] = rb[0] + ax * time_step;
    rb[1] = rb[1] + ay * time_step;
    rb[2] = ax; rb[3] = ay;
    for (int i = 0; i < na; ++i) {
        double *hh = humans + ((size_t)b * n + i) * 8;
        const double vx = new_vel[((size_t)b * n + i) * 2], vy = new_vel[((size_t)b * n + i) * 2 + 1];
        hh[0] = hh[0] + vx * time_step;
        hh[1] = hh[1] + vy * time_step;
        hh[2] = vx; hh[3] = vy;
    }
    global_time[b] += time_step;
}

static int g_tile_cap[64] = {0};     // per device: cudaFuncSetAttribute is a per-device setting

cudaError_t launch_env_lookahead(int B, int n, int A, const double *robot, const double *humans,
                                    const double *new_vel, const double *global_time, const double *actions,
                                    int action_env_stride, const double *lut, const aem_env_params &p, double amax, double *rewards,
                                    int32_t *info, double *dmin, double *humans_next, cudaStream_t st, const int32_t *nact,
                                    double *new_vel_out)
{
    if (B <= 0) return cudaSuccess;
    const int side = 2 * (int)ceil(1.5 * p.human_timestep / GRID_RES) + 3;
    const int orca_doubles = 2 * (ENV_THREADS / 32) * 32 * (int)sizeof(Line) / 8;      // the fused ORCA phase borrows the tile
    const int tile_cap = side * side > orca_doubles ? side * side : orca_doubles;
    const size_t smem = (size_t)(tile_cap + ENV_THREADS + 8 + AEM_MAX_ACTIONS + AEM_MAX_HUMANS * 8 + 9 + 1) * 8;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    int dev = 0;
    cudaGetDevice(&dev);
    if (tile_cap > g_tile_cap[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(env_lookahead_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
        g_tile_cap[dev & 63] = tile_cap;
    }
    env_lookahead_kernel<<<B, ENV_THREADS, smem, st>>>(B, n, A, robot, humans, new_vel, global_time, actions,
                                                       action_env_stride, lut, p,
                                                       amax, tile_cap, rewards, info, dmin, humans_next, nact, new_vel_out,
                                                       p.robot_visible);
    return cudaGetLastError();
}

cudaError_t launch_env_apply(int B, int n, int A, double *robot, double *humans, const double *new_vel,
                             double *global_time, const double *actions, int action_env_stride,
                             const int32_t *chosen,
                             const double *rewards, const int32_t *info, const double *dmin, double time_step,
                             double *out_reward, int32_t *out_done, int32_t *out_info, double *out_dmin,
                             cudaStream_t st, const int32_t *nact)
{
    if (B <= 0) return cudaSuccess;
    const int threads = 128;
    env_apply_kernel<<<(B + threads - 1) / threads, threads, 0, st>>>(B, n, A, robot, humans, new_vel, global_time,
                                                                      actions, action_env_stride, chosen, rewards, info,
                                                                      dmin, time_step,
                                                                      out_reward, out_done, out_info, out_dmin, nact);
    return cudaGetLastError();
}

// Explorer.run_k_episodes (explorer.py:44-80) keeps one env busy with consecutive cases; here every finished env of
// the batch is re-seeded from a pool of pre-generated scenarios (CrowdSim.reset, crowd_sim.py:486-552 runs on the
// host with numpy's MT19937 stream) and the episode outcome is counted.
__global__ void env_reset_kernel(int B, int n, double *__restrict__ robot, double *__restrict__ humans,
                                 double *__restrict__ global_time, const int32_t *__restrict__ done,
                                 const int32_t *__restrict__ info, const double *__restrict__ pool_robot,
                                 const double *__restrict__ pool_humans, int P, int32_t *__restrict__ case_idx,
                                 int case_stride, unsigned long long *__restrict__ outcome_counts,
                                 int32_t *__restrict__ nact, const int32_t *__restrict__ pool_nact)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || !done[b]) return;
    const int c = case_idx[b];
    const int src = ((c % P) + P) % P;
    for (int k = 0; k < 9; ++k) robot[(size_t)b * 9 + k] = pool_robot[(size_t)src * 9 + k];
    for (int k = 0; k < n * 8; ++k) humans[(size_t)b * n * 8 + k] = pool_humans[(size_t)src * n * 8 + k];
    if (nact && pool_nact) nact[b] = pool_nact[src];      // the new scenario's human count ('mixed' rule)
    global_time[b] = 0.0;
    case_idx[b] = c + case_stride;
    if (outcome_counts && info) {
        const int code = info[b];
        if (code >= 0 && code < 5) atomicAdd(outcome_counts + code, 1ULL);
    }
}

cudaError_t launch_env_reset(int B, int n, double *robot, double *humans, double *global_time, const int32_t *done,
                             const int32_t *info, const double *pool_robot, const double *pool_humans, int P,
                             int32_t *case_idx, int case_stride, unsigned long long *outcome_counts, cudaStream_t st,
                             int32_t *nact, const int32_t *pool_nact)
{
    if (B <= 0) return cudaSuccess;
    const int threads = 128;
    env_reset_kernel<<<(B + threads - 1) / threads, threads, 0, st>>>(B, n, robot, humans, global_time, done, info,
                                                                      pool_robot, pool_humans, P, case_idx,
                                                                      case_stride, outcome_counts, nact, pool_nact);
    return cudaGetLastError();
}

// The end of a decision step in one launch, one warp per environment: the argmax of select_kernel (multi_human_rl.py:
// 321-322, 368-372), step(update=True) of env_apply_kernel for the action it chose, and -- when a scenario pool is given --
// the re-seeding of env_reset_kernel for the environments that finished.  Same arithmetic, same order per environment.
__global__ void __launch_bounds__(256) env_finish_kernel(
    int B, int n, int A, const double *__restrict__ values, double *__restrict__ robot, double *__restrict__ humans,
    const double *__restrict__ new_vel, double *__restrict__ global_time, const double *__restrict__ actions_all,
    int action_env_stride, const double *__restrict__ rewards, const int32_t *__restrict__ info,
    const double *__restrict__ dmin, double time_step, int32_t *__restrict__ best, double *__restrict__ out_reward,
    int32_t *__restrict__ out_done, int32_t *__restrict__ out_info, double *__restrict__ out_dmin, int32_t *nact,
    const double *__restrict__ pool_robot, const double *__restrict__ pool_humans, int P, int32_t *__restrict__ case_idx,
    int case_stride, unsigned long long *__restrict__ outcome_counts, const int32_t *__restrict__ pool_nact)
{
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const unsigned FULLM = 0xffffffffu;
    // ---- select: first strict maximum, NaN never wins; the goal short-circuit picks action 0
    const double *v = values + (size_t)b * A;
    double bv = -INFINITY;
    int bi = -1;
    for (int a = lane; a < A; a += 32) {
        const double x = v[a];
        if (x > bv) { bv = x; bi = a; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(FULLM, bv, off);
        const int oi = __shfl_xor_sync(FULLM, bi, off);
        if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    double *rb = robot + (size_t)b * 9;
    {
        const double dx = rb[0] - rb[5], dy = rb[1] - rb[6];
        if (sqrt(dy * dy + dx * dx) < rb[4]) bi = 0;
    }
    if (lane == 0 && best) best[b] = bi;
    __syncwarp();                       // every lane has read the robot row before lane 0 advances it
    // ---- apply (crowd_sim.py:653-669, agent.py:122-131)
    const int na = nact ? min(max(nact[b], 0), n) : n;
    int a = bi;
    if (a < 0 || a >= A) a = 0;
    const double *__restrict__ actions = actions_all + (size_t)b * action_env_stride;
    const double ax = actions[2 * a], ay = actions[2 * a + 1];
    const int code = info[(size_t)b * A + a];
    const int done = (code >= AEM_INFO_REACHGOAL) ? 1 : 0;
    const bool reseed = pool_robot != nullptr && done;
    if (lane == 0) {
        if (out_reward) out_reward[b] = rewards[(size_t)b * A + a];
        if (out_info) out_info[b] = code;
        if (out_done) out_done[b] = done;
        if (dmin && out_dmin) out_dmin[b] = dmin[(size_t)b * A + a];
    }
    if (!reseed) {
        if (lane == 0) {
            rb[0] = rb[0] + ax * time_step;
            rb[1] = rb[1] + ay * time_step;
            rb[2] = ax; rb[3] = ay;
            global_time[b] += time_step;
        }
        for (int i = lane; i < na; i += 32) {
            double *hh = humans + ((size_t)b * n + i) * 8;
            const double vx = new_vel[((size_t)b * n + i) * 2], vy = new_vel[((size_t)b * n + i) * 2 + 1];
            hh[0] = hh[0] + vx * time_step;
            hh[1] = hh[1] + vy * time_step;
            hh[2] = vx; hh[3] = vy;
        }
        return;
    }
    // ---- the episode ended: the next scenario of this environment (the applied state would be overwritten)
    const int c = case_idx[b];
    const int src = ((c % P) + P) % P;
    __syncwarp();
    for (int k = lane; k < 9; k += 32) robot[(size_t)b * 9 + k] = pool_robot[(size_t)src * 9 + k];
    for (int k = lane; k < n * 8; k += 32) humans[(size_t)b * n * 8 + k] = pool_humans[(size_t)src * n * 8 + k];
    if (lane == 0) {
        if (nact && pool_nact) nact[b] = pool_nact[src];
        global_time[b] = 0.0;
        case_idx[b] = c + case_stride;
        if (outcome_counts && code >= 0 && code < 5) atomicAdd(outcome_counts + code, 1ULL);
    }
}

cudaError_t launch_env_finish(int B, int n, int A, const double *values, double *robot, double *humans, const double *new_vel,
                              double *global_time, const double *actions, int action_env_stride, const double *rewards,
                              const int32_t *info, const double *dmin, double time_step, int32_t *best, double *out_reward,
                              int32_t *out_done, int32_t *out_info, double *out_dmin, int32_t *nact, const double *pool_robot,
                              const double *pool_humans, int P, int32_t *case_idx, int case_stride,
                              unsigned long long *outcome_counts, const int32_t *pool_nact, cudaStream_t st)
{
    if (B <= 0) return cudaSuccess;
    const int threads = 256;
    const int blocks = (int)(((long long)B * 32 + threads - 1) / threads);
    env_finish_kernel<<<blocks, threads, 0, st>>>(B, n, A, values, robot, humans, new_vel, global_time, actions,
                                                  action_env_stride, rewards, info, dmin, time_step, best, out_reward,
                                                  out_done, out_info, out_dmin, nact, pool_robot, pool_humans, P, case_idx,
                                                  case_stride, outcome_counts, pool_nact);
    return cudaGetLastError();
}

}  // namespace aem
