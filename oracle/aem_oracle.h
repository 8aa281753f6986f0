/*
 * aem_oracle.h -- CPU restatement (ORACLE) of the AEMCARL lookahead hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker or
 * the CPU baseline -- never as a fallback of the CUDA path.
 *
 * Every function cites the reference file:line (relative to the upstream
 * repository SJWang2015/AEMCARL) whose arithmetic it restates.
 *
 * Parity status
 *   value network, rotate, action table, reward map, scenario generation:
 *     pinned against golden vectors produced by importing the unmodified
 *     reference (oracle/make_golden.py -> tests/golden/).
 *   ORCA: **parity unpinned**.  The arithmetic lives in the third-party
 *     Python-RVO2 / RVO2 v2.0 C++ library (pyrvo2==0.0.0,
 *     crowd_nav/requirements.txt:147) which is not vendored and not
 *     installable here; aemo_orca_* restates the published RVO2 algorithm
 *     (Agent::computeNeighbors / computeNewVelocity / linearProgram1-3) in
 *     FP32 and is pinned by our own known-answer + property tests only.
 */
#ifndef AEM_ORACLE_H
#define AEM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AEMO_NUM_ACTIONS 81
#define AEMO_MAX_HUMANS 32
#define AEMO_HID 50
#define AEMO_XDIM 13

/* ---- flat weight blob (live parameters of ValueNetworkTransformerAndGRU,
 * state_dict order of crowd_nav/common/qnetwork_factory.py:586-639 without the
 * dead `encoder_layer.*` template), all row-major [out][in] like nn.Linear ---- */
enum {
    AEMO_OFF_Z1W = 0,                         /* in_mlp.rnn_cell.convz.0.weight [100][63] */
    AEMO_OFF_Z1B = AEMO_OFF_Z1W + 100 * 63,   /* convz.0.bias [100] */
    AEMO_OFF_Z2W = AEMO_OFF_Z1B + 100,        /* convz.2.weight [50][100] */
    AEMO_OFF_Z2B = AEMO_OFF_Z2W + 50 * 100,   /* convz.2.bias [50] */
    AEMO_OFF_R1W = AEMO_OFF_Z2B + 50,         /* convr.0.weight [100][63] */
    AEMO_OFF_R1B = AEMO_OFF_R1W + 100 * 63,
    AEMO_OFF_R2W = AEMO_OFF_R1B + 100,        /* convr.2.weight [50][100] */
    AEMO_OFF_R2B = AEMO_OFF_R2W + 50 * 100,
    AEMO_OFF_QW = AEMO_OFF_R2B + 50,          /* convq.weight [50][63] */
    AEMO_OFF_QB = AEMO_OFF_QW + 50 * 63,
    AEMO_OFF_PW = AEMO_OFF_QB + 50,           /* in_mlp.ponder_linear.weight [1][50] */
    AEMO_OFF_PB = AEMO_OFF_PW + 50,           /* ponder_linear.bias [1] */
    AEMO_OFF_ENC = AEMO_OFF_PB + 1,           /* 3 x encoder layer, AEMO_ENC_SIZE each */
    /* inside one encoder layer: */
    AEMO_ENC_INW = 0,                         /* self_attn.in_proj_weight [150][50] */
    AEMO_ENC_INB = AEMO_ENC_INW + 150 * 50,   /* in_proj_bias [150] */
    AEMO_ENC_OW = AEMO_ENC_INB + 150,         /* out_proj.weight [50][50] */
    AEMO_ENC_OB = AEMO_ENC_OW + 50 * 50,
    AEMO_ENC_L1W = AEMO_ENC_OB + 50,          /* linear1.weight [150][50] */
    AEMO_ENC_L1B = AEMO_ENC_L1W + 150 * 50,
    AEMO_ENC_L2W = AEMO_ENC_L1B + 150,        /* linear2.weight [50][150] */
    AEMO_ENC_L2B = AEMO_ENC_L2W + 50 * 150,
    AEMO_ENC_N1G = AEMO_ENC_L2B + 50,         /* norm1.weight, norm1.bias */
    AEMO_ENC_N1B = AEMO_ENC_N1G + 50,
    AEMO_ENC_N2G = AEMO_ENC_N1B + 50,
    AEMO_ENC_N2B = AEMO_ENC_N2G + 50,
    AEMO_ENC_SIZE = AEMO_ENC_N2B + 50,        /* = 25600 */
    AEMO_OFF_A0W = AEMO_OFF_ENC + 3 * AEMO_ENC_SIZE, /* action_mlp.0.weight [100][56] */
    AEMO_OFF_A0B = AEMO_OFF_A0W + 100 * 56,
    AEMO_OFF_A2W = AEMO_OFF_A0B + 100,        /* action_mlp.2.weight [50][100] */
    AEMO_OFF_A2B = AEMO_OFF_A2W + 50 * 100,
    AEMO_OFF_A4W = AEMO_OFF_A2B + 50,         /* action_mlp.4.weight [1][50] */
    AEMO_OFF_A4B = AEMO_OFF_A4W + 50,
    AEMO_NUM_WEIGHTS = AEMO_OFF_A4B + 1       /* = 113752 */
};

/* Environment / reward parameters (crowd_sim/envs/crowd_sim.py:265-299, env.config). */
typedef struct {
    double time_step;        /* [env] time_step (0.2) */
    double time_limit;       /* [env] time_limit (20) */
    double success_reward;   /* [reward] success_reward (1) */
    double collision_penalty;/* [reward] collision_penalty (-0.25) */
    double discomfort_dist;  /* [reward] discomfort_dist (0.2) */
    double agent_timestep;   /* [reward] agent_timestep */
    double human_timestep;   /* [reward] human_timestep */
    double reward_increment; /* [reward] reward_increment */
    double position_variance;
    double direction_variance;
    int32_t robot_visible;   /* robot.visible (test.py:127) */
    int32_t pad_;
    double orca_safety_space;/* ORCA.safety_space of the humans (0) */
} aemo_env_params;

/* info codes written by the step functions (crowd_sim/envs/utils/info.py) */
enum { AEMO_INFO_NOTHING = 0, AEMO_INFO_DANGER = 1, AEMO_INFO_REACHGOAL = 2,
       AEMO_INFO_COLLISION = 3, AEMO_INFO_TIMEOUT = 4 };

int aemo_num_weights(void);

/* cadrl.py:129-154 build_action_space (holonomic): out[81][2] */
void aemo_build_action_space(double v_pref, int speed_samples, int rotation_samples, double *out);

/* cadrl.py:239-274 rotate (FP32); in: 14 floats (robot 9, human 5), out 13 */
void aemo_rotate(const float *in14, float *out13);

/* qnetwork_factory.py:641-687 + components.py:38-137: one (env, action) sample.
 * state [n][13] (already rotated); act_fixed/act_steps as in ATCBasic. */
void aemo_value_forward(const float *W, const float *state, int n, int act_fixed, int act_steps,
                        float *value, int *steps);

/* min |accum_p - 0.95| over tokens / executed steps of one sample (how close the ACT halting test was to flipping) */
float aemo_act_margin(const float *W, const float *state, int n);

/* multi_human_rl.py:336-372 over a batch of envs: values[B][A] (f64 = reward + gbar*V),
 * net_values[B][A] (f32 raw V), best[B] (first strict maximum), steps[B][A].
 * robot[B][9] f64, humans_next[B][n][5] f64, rewards[B][A] f64, actions[A][2] f64. */
void aemo_lookahead(const float *W, int B, int n, int A, const double *robot, const double *humans_next,
                    const double *rewards, const double *actions, double dt, double gamma_bar,
                    int act_fixed, int act_steps, double *values, float *net_values, int32_t *best,
                    uint8_t *steps, int nthreads);

/* batched flag-5 forward on already rotated states [S][n][13] with per-sample halting
 * (what Explorer.update_memory / Trainer see, explorer.py:137, trainer.py:52) */
void aemo_value_forward_batch(const float *W, int S, int n, const float *states, int act_fixed, int act_steps,
                              float *values, uint8_t *steps, int nthreads);

/* GRU-ACT only (components.py:107-137) on rows grouped in samples of T rows: x[S*T][13] -> out[S*T][50] */
void aemo_gru_act(const float *W, int S, int T, const float *x, int act_fixed, int act_steps, float *out,
                  uint8_t *steps);

/* orca.py:82-132 + RVO2 Agent.cpp (FP32): new velocity for every human of every env.
 * humans[B][n][8] f64 = px,py,vx,vy,radius,gx,gy,v_pref ; robot[B][9]; out vel[B][n][2] f64 */
void aemo_orca(int B, int n, const double *humans, const double *robot, const aemo_env_params *p,
               double *new_vel, int nthreads);

/* single-agent ORCA solve exposed for the rvo2-shaped stub and the property tests.
 * others[m][5] f32 = px,py,vx,vy,radius(already incl. +0.01); returns new velocity, optionally the
 * neighbour order and the half-planes used. */
void aemo_orca_agent(const float *self_pos, const float *self_vel, float self_radius, float max_speed,
                     const float *pref_vel, int m, const float *others, float neighbor_dist, int max_neighbors,
                     float time_horizon, float time_step, float *new_vel, int *nb_idx, int *num_nb, float *lines);

/* crowd_sim.py:41-132,159-219: occupancy probability of the robot disc at (rx, ry) (already the
 * agent_timestep look-ahead position) given humans' CURRENT states hum[n][5] = px,py,vx,vy,radius */
double aemo_occupancy(const double *hum, int n, double rx, double ry, double robot_radius,
                      double human_timestep, double pos_var, double dir_var);

/* crowd_sim.py:557-691 with update=False for all A actions of every env:
 * rewards[B][A], info[B][A] (codes), dmin[B][A], humans_next[B][n][5] (get_next_observable_state),
 * given new_vel from aemo_orca. global_time[B]. */
void aemo_env_lookahead(int B, int n, int A, const double *robot, const double *humans, const double *new_vel,
                        const double *global_time, const double *actions, const aemo_env_params *p,
                        double *rewards, int32_t *info, double *dmin, double *humans_next, int nthreads);

/* crowd_sim.py:653-669 update=True for the chosen action index per env (in place) */
void aemo_env_apply(int B, int n, int A, double *robot, double *humans, const double *new_vel,
                    double *global_time, const double *actions, const int32_t *chosen,
                    const aemo_env_params *p);

#ifdef __cplusplus
}
#endif
#endif
