// scenario.cu -- CrowdSim.reset's seeded scenario generation on the device:
//   crowd_sim/envs/crowd_sim.py:486-552 (reset: np.random.seed(counter_offset + case)), :390-411 (circle crossing),
//   :413-442 (square crossing), agent.py:39-45 (sample_random_attributes).
// One thread per test case.  Each thread owns a complete MT19937 generator in local memory (numpy's legacy
// RandomState(int seed) = init_genrand; random_sample() = genrand_res53 from two tempered words) and replays the
// reference's draw order, rejection sampling included.  The 624-word twist is evaluated lazily, one word per draw, in
// the same in-place order as the block regeneration, so the stream is identical word for word.
// Compiled with -fmad=false: every FP64 operation is rounded individually like the Python float arithmetic.  The only
// non-bit-exact ingredient is cos / sin of the circle rule (CUDA libdevice vs numpy's libm: <= 1 ulp apart); the square
// rule is pure arithmetic and reproduces the host generator bit for bit (tests/test_gpu_api.py).
// The reference's rejection loops never end when the humans do not fit (e.g. 18 randomised radii on the 4 m circle); a
// kernel must: after SCN_MAX_TRIES rejected draws of one human the case is given up and filled with NaN.
#include "aem_common.cuh"

namespace aem {

struct Mt19937 {
    uint32_t mt[624];
    int idx;     // next word to produce, 0..623 within the current block
    __device__ void seed(uint32_t s)
    {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 0;
    }
    __device__ uint32_t next()
    {
        const int i = idx, i1 = i == 623 ? 0 : i + 1, im = i + 397 < 624 ? i + 397 : i + 397 - 624;
        const uint32_t y = (mt[i] & 0x80000000u) | (mt[i1] & 0x7fffffffu);
        uint32_t v = mt[im] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        mt[i] = v;
        idx = i1;
        v ^= v >> 11;
        v ^= (v << 7) & 0x9d2c5680u;
        v ^= (v << 15) & 0xefc60000u;
        v ^= v >> 18;
        return v;
    }
    __device__ double random_sample()
    {
        const uint32_t a = next() >> 5, b = next() >> 6;
        return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    }
    __device__ double uniform(double lo, double hi) { return lo + (hi - lo) * random_sample(); }
};

constexpr int SCN_MAX_TRIES = 20000;

__device__ __forceinline__ double norm2(double dx, double dy) { return sqrt(dx * dx + dy * dy); }

// agents so far: robot + humans [0, m); layout of `humans` rows: px,py,vx,vy,radius,gx,gy,v_pref
// rule 2 = 'mixed' (crowd_sim.py:337-386): a static scene with probability 0.2, else a dynamic one; the number of humans
// (<= 5) is part of the draw and goes to n_active_out, rows beyond it are padding (parked at (0, -100), zero velocity).
__global__ void __launch_bounds__(64) scenario_kernel(int rule, uint32_t first_seed, int count, int n, aem_scenario_params p,
                                                      double *__restrict__ robot_out, double *__restrict__ humans_out,
                                                      int32_t *__restrict__ n_active_out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    Mt19937 rs;
    rs.seed(first_seed + (uint32_t)c);
    const double PI = 3.141592653589793;
    double *rb = robot_out + (size_t)c * 9, *hs = humans_out + (size_t)c * n * 8;
    const double rpx = 0.0, rpy = -p.circle_radius, rgx = 0.0, rgy = p.circle_radius;      // crowd_sim.py:512
    rb[0] = rpx; rb[1] = rpy; rb[2] = 0.0; rb[3] = 0.0; rb[4] = p.robot_radius; rb[5] = rgx; rb[6] = rgy;
    rb[7] = p.robot_v_pref; rb[8] = PI / 2;
    bool failed = false;
    int humans_here = n;            // humans of this case (the 'mixed' rule draws it)
    bool is_static = false;
    if (rule == 2) {
        is_static = rs.random_sample() < 0.2;
        double prob = rs.random_sample();
        const double ps[6] = {0.05, 0.2, 0.2, 0.3, 0.1, 0.15};     // static scenes: 0..5 humans
        const double pd[5] = {0.3, 0.3, 0.2, 0.1, 0.1};            // dynamic scenes: 1..5 humans
        humans_here = 5;                                           // (the reference keeps its configured 5 if no entry matches)
        for (int k = 0; k < (is_static ? 6 : 5); ++k) {
            const double pk = is_static ? ps[k] : pd[k];
            if (prob - pk <= 0) { humans_here = is_static ? k : k + 1; break; }
            prob -= pk;
        }
        for (int m = 0; m < n; ++m) {                              // padding rows first; real humans overwrite theirs
            double *h = hs + m * 8;
            h[0] = 0.0; h[1] = -100.0; h[2] = 0.0; h[3] = 0.0; h[4] = p.human_radius; h[5] = 0.0; h[6] = -100.0; h[7] = 0.0;
        }
        if (is_static && humans_here == 0) {                       // one dummy human parked at (0, -10), crowd_sim.py:355-358
            hs[0] = 0.0; hs[1] = -10.0; hs[4] = p.human_radius; hs[5] = 0.0; hs[6] = -10.0; hs[7] = p.human_v_pref;
            if (n_active_out) n_active_out[c] = 1;
            return;
        }
        if (n_active_out) n_active_out[c] = humans_here;
    }
    for (int m = 0; m < humans_here && !failed; ++m) {
        // kind of this human: 0 crosses the circle, 1 crosses the square, 2 stands still ('mixed': static scenes; in a
        // dynamic scene the first two cross the circle, the others the square)
        const int kind = rule < 2 ? rule : (is_static ? 2 : (m < 2 ? 0 : 1));
        double v_pref = p.human_v_pref, radius = p.human_radius;
        if (p.randomize_attributes && kind != 2) {
            v_pref = rs.uniform(0.5, 1.5);
            radius = rs.uniform(0.3, 0.5);
        }
        double px = 0.0, py = 0.0, gx = 0.0, gy = 0.0;
        if (kind == 2) {                                   // a static obstacle in a 4 x 8 box, goal = start
            const double sign = rs.random_sample() > 0.5 ? -1.0 : 1.0;
            for (int tries = 0;; ++tries) {
                if (tries >= SCN_MAX_TRIES) { failed = true; break; }
                px = rs.random_sample() * 4.0 * 0.5 * sign;
                py = (rs.random_sample() - 0.5) * 8.0;
                bool collide = false;
                for (int a = -1; a < m && !collide; ++a) {
                    const double apx = a < 0 ? rpx : hs[a * 8 + 0], apy = a < 0 ? rpy : hs[a * 8 + 1];
                    const double ar = a < 0 ? p.robot_radius : hs[a * 8 + 4];
                    collide = !(norm2(px - apx, py - apy) >= radius + ar + p.discomfort_dist);
                }
                if (!collide) break;
            }
            gx = px;
            gy = py;
        } else if (kind == 0) {                            // circle crossing: start on the circle, goal opposite
            for (int tries = 0;; ++tries) {
                if (tries >= SCN_MAX_TRIES) { failed = true; break; }
                const double angle = rs.random_sample() * PI * 2;
                const double px_noise = (rs.random_sample() - 0.5) * v_pref;
                const double py_noise = (rs.random_sample() - 0.5) * v_pref;
                px = p.circle_radius * cos(angle) + px_noise;
                py = p.circle_radius * sin(angle) + py_noise;
                bool collide = false;
                for (int a = -1; a < m && !collide; ++a) {
                    const double apx = a < 0 ? rpx : hs[a * 8 + 0], apy = a < 0 ? rpy : hs[a * 8 + 1];
                    const double agx = a < 0 ? rgx : hs[a * 8 + 5], agy = a < 0 ? rgy : hs[a * 8 + 6];
                    const double ar = a < 0 ? p.robot_radius : hs[a * 8 + 4];
                    const double min_dist = radius + ar + p.discomfort_dist;
                    collide = norm2(px - apx, py - apy) < min_dist || norm2(px - agx, py - agy) < min_dist;
                }
                if (!collide) break;
            }
            gx = -px;
            gy = -py;
        } else {                                           // square crossing: start on one side, goal on the other
            const double sign = rs.random_sample() > 0.5 ? -1.0 : 1.0;
            for (int tries = 0;; ++tries) {
                if (tries >= SCN_MAX_TRIES) { failed = true; break; }
                px = rs.random_sample() * p.square_width * 0.5 * sign;
                py = (rs.random_sample() - 0.5) * p.square_width;
                bool collide = false;
                for (int a = -1; a < m && !collide; ++a) {
                    const double apx = a < 0 ? rpx : hs[a * 8 + 0], apy = a < 0 ? rpy : hs[a * 8 + 1];
                    const double ar = a < 0 ? p.robot_radius : hs[a * 8 + 4];
                    collide = norm2(px - apx, py - apy) < radius + ar + p.discomfort_dist;
                }
                if (!collide) break;
            }
            for (int tries = 0;; ++tries) {
                if (tries >= SCN_MAX_TRIES) { failed = true; break; }
                gx = rs.random_sample() * p.square_width * 0.5 * -sign;
                gy = (rs.random_sample() - 0.5) * p.square_width;
                bool collide = false;
                for (int a = -1; a < m && !collide; ++a) {
                    const double agx = a < 0 ? rgx : hs[a * 8 + 5], agy = a < 0 ? rgy : hs[a * 8 + 6];
                    const double ar = a < 0 ? p.robot_radius : hs[a * 8 + 4];
                    collide = norm2(gx - agx, gy - agy) < radius + ar + p.discomfort_dist;
                }
                if (!collide) break;
            }
        }
        double *h = hs + m * 8;
        h[0] = px; h[1] = py; h[2] = 0.0; h[3] = 0.0; h[4] = radius; h[5] = gx; h[6] = gy; h[7] = v_pref;
    }
    if (failed) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        for (int i = 0; i < n * 8; ++i) hs[i] = nan;
        if (n_active_out) n_active_out[c] = 0;
    }
}

cudaError_t launch_scenarios(int rule, uint32_t first_seed, int count, int n, const aem_scenario_params &p, double *robot,
                             double *humans, cudaStream_t st, int32_t *n_active)
{
    scenario_kernel<<<(count + 63) / 64, 64, 0, st>>>(rule, first_seed, count, n, p, robot, humans, n_active);
    return cudaGetLastError();
}

}  // namespace aem
