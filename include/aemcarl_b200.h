/*
 * aemcarl_b200.h -- C ABI of libaemcarl_b200.so (sm_100a).
 *
 * Drop-in boundary for AEMCARL's one-step-lookahead hot path.  The reference has no FFI layer of its own
 * (it is pure Python); the entry points below are what a maintainer of SJWang2015/AEMCARL binds with
 * ctypes to replace, call for call:
 *
 *   aem_orca            <- crowd_sim/envs/policy/orca.py:82-132  (ORCA.predict -> rvo2.PyRVOSimulator.doStep,
 *                          called n times per env.step at crowd_sim/envs/crowd_sim.py:562-568)
 *   aem_env_lookahead   <- crowd_sim/envs/crowd_sim.py:554-651,676-680 (onestep_lookahead == step(update=False)
 *                          for every action of the action space; gridmap reward crowd_sim.py:41-219)
 *   aem_lookahead       <- crowd_nav/policy/multi_human_rl.py:336-372 (propagate cadrl.py:156-181, rotate
 *                          cadrl.py:239-274, ValueNetworkTransformerAndGRU.forward qnetwork_factory.py:641-687,
 *                          ATCBasic/GRUEXND components.py:38-137, r + gamma^(dt*v_pref) * V, strict-'>' argmax)
 *   aem_env_apply       <- crowd_sim/envs/crowd_sim.py:653-669 (step(update=True): advance robot/humans/time)
 *   aem_env_prepare     <- aem_orca + aem_env_lookahead in ONE launch (crowd_sim.py:554-651 calls ORCA at :562-568 itself)
 *   aem_env_finish      <- argmax of aem_lookahead + aem_env_apply + aem_env_reset_done in ONE launch (multi_human_rl.py:368-372,
 *                          crowd_sim.py:653-669, explorer.py:44-47)
 *   aem_env_reset_done  <- CrowdSim.reset for finished episodes of a batch (crowd_sim.py:486-552, explorer.py:44-47)
 *   aem_value_forward   <- model(state) on already rotated states (explorer.py:137, trainer.py:52,75)
 *   aem_gru_act         <- ATCBasic.forward components.py:107-137 (fused GRU-cell / adaptive-computation-time)
 *   aem_env_step_xy     <- CrowdSim.step for one arbitrary ActionXY per environment (imitation phase, train.py:171-204)
 *   aem_transform       <- MultiHumanRL.transform multi_human_rl.py:423-438 (replay-memory sample of the current state)
 *   aem_generate_scenarios <- CrowdSim.reset's seeded scenario draw, crowd_sim.py:486-552,390-442
 *   aem_generate_mixed_scenarios <- the 'mixed' rule of that draw (1..5 humans per case), crowd_sim.py:337-386
 *   aem_train_grad      <- Trainer.optimize_batch / optimize_epoch forward + MSE + backward (crowd_nav/utils/trainer.py:47-58,70-82)
 *   aem_train_step_adam <- one iteration of Trainer.optimize_batch with Adam (sample + forward/backward + update), trainer.py:66-86
 *   aem_sample_batch    <- next(iter(DataLoader(memory, batch_size, shuffle=True))), trainer.py:70-72 + memory.py
 *   aem_adam_step / aem_sgd_step / aem_rmsprop_step <- optimizer.step() of the three optimizers of trainer.py:24-34
 *   aem_decide_step_host<- Robot.act + CrowdSim.step through HOST buffers (explorer.py:51-53), the e2e call
 *
 * Conventions: plain pointers and sizes only; `*_d` = device pointer, `*_h` = host pointer; all device entry
 * points are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) and the
 * caller owns every buffer.  Return value 0 = ok, otherwise an AEM_ERR_* code with a message available from
 * aem_last_error().  One handle per (process, device); a handle is not thread-safe.
 *
 * Array layouts (row-major):
 *   robot   [B][9]   f64  px,py,vx,vy,radius,gx,gy,v_pref,theta   (FullState, state.py:1-18)
 *   humans  [B][n][8] f64 px,py,vx,vy,radius,gx,gy,v_pref
 *   new_vel [B][n][2] f64 ORCA velocity of every human (FP32 value widened, as rvo2 returns it)
 *   humans_next [B][n][5] f64 px,py,vx,vy,radius  (Agent.get_next_observable_state, agent.py:63-74)
 *   actions [A][2]   f64  ActionXY table of CADRL.build_action_space (cadrl.py:129-154), A <= 128
 *   rewards / values / dmin [B][A] f64 ; info [B][A] i32 ; net_values [B][A] f32 ; act_steps [B][A] u8
 */
#ifndef AEMCARL_B200_H
#define AEMCARL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AEM_VERSION 100          /* 0.1.0 */
#define AEM_MAX_HUMANS 32
#define AEM_MAX_ACTIONS 128
#define AEM_NUM_WEIGHTS 113752   /* live floats of the flag-5 network in state_dict order */

enum {
    AEM_OK = 0,
    AEM_ERR_INVALID = 1,   /* bad argument (sizes, null pointers, n > AEM_MAX_HUMANS ...) */
    AEM_ERR_CUDA = 2,      /* CUDA runtime error (message in aem_last_error) */
    AEM_ERR_STATE = 3      /* weights / action space / env params not set yet */
};

/* info codes (crowd_sim/envs/utils/info.py): Nothing, Danger, ReachGoal, Collision, Timeout */
enum { AEM_INFO_NOTHING = 0, AEM_INFO_DANGER = 1, AEM_INFO_REACHGOAL = 2, AEM_INFO_COLLISION = 3,
       AEM_INFO_TIMEOUT = 4 };

/* crowd_sim.py:265-299 CrowdSim.configure + robot.visible (test.py:127) + ORCA.safety_space */
typedef struct {
    double time_step;
    double time_limit;
    double success_reward;
    double collision_penalty;
    double discomfort_dist;
    double agent_timestep;
    double human_timestep;
    double reward_increment;
    double position_variance;
    double direction_variance;
    int32_t robot_visible;
    int32_t pad_;
    double orca_safety_space;
} aem_env_params;

/* CrowdSim.configure / Agent attributes that enter the scenario draw (env.config [sim], [humans], [robot]) */
typedef struct {
    double circle_radius;
    double square_width;
    double discomfort_dist;
    double robot_radius;
    double robot_v_pref;
    double human_radius;
    double human_v_pref;
    int32_t randomize_attributes;
    int32_t pad_;
} aem_scenario_params;

typedef struct aem_handle_s *aem_handle;

int aem_version(void);
const char *aem_last_error(void);

/* device = CUDA ordinal.  Allocates the weight copies, LUT and scratch on that device. */
int aem_create(int device, aem_handle *out);
int aem_destroy(aem_handle h);

/* weights_h: AEM_NUM_WEIGHTS floats in state_dict order (aemcarl_b200/weights.py WEIGHT_SPEC).  Synchronous. */
int aem_load_weights(aem_handle h, const float *weights_h, size_t count);
/* actions_h [A][2] f64.  Synchronous. */
int aem_set_action_space(aem_handle h, const double *actions_h, int A);
int aem_set_env_params(aem_handle h, const aem_env_params *p);
/* ATCBasic act_fixed / act_steps (components.py:119-123); default adaptive (act_fixed = 0). */
int aem_set_act_mode(aem_handle h, int act_fixed, int act_steps);

/* Arithmetic of the value-network GEMMs inside aem_lookahead / aem_value_forward / aem_gru_act:
 *   0 = FP32 FFMA (k-ascending accumulation, the order of the CPU oracle)            [default]
 *   1 = tcgen05 tensor cores, 3xTF32 split with FP32 accumulation in tensor memory (~1e-6 relative of mode 0)
 *   2 = tcgen05 tensor cores, FP16 hi/lo split (3 MMAs per product, FP32 accumulation), two tiles in flight per SM
 *       (the fastest; what bench.py measures)
 *   3 = same arithmetic as 2, four threads per token row, one tile per SM (experimental).
 * Every mode is held to the same parity bar (tests/test_gpu_parity.py). */
int aem_set_precision(aem_handle h, int precision);

/* ---- device-pointer entry points ---- */
int aem_orca(aem_handle h, int B, int n, const double *humans_d, const double *robot_d, double *new_vel_d,
             void *stream);

int aem_env_lookahead(aem_handle h, int B, int n, const double *robot_d, const double *humans_d,
                      const double *new_vel_d, const double *global_time_d, double *rewards_d, int32_t *info_d,
                      double *dmin_d, double *humans_next_d, void *stream);

/* values_d = rewards + gamma_bar * V (f64); best_d = first strict maximum (-1 if every value is NaN; 0 when the
 * robot already is at its goal, policy.py:45-49).  net_values_d / act_steps_d may be NULL; best_d may be NULL when
 * aem_env_finish takes the argmax (one launch less). */
int aem_lookahead(aem_handle h, int B, int n, const double *robot_d, const double *humans_next_d,
                  const double *rewards_d, double gamma_bar, double *values_d, float *net_values_d,
                  int32_t *best_d, uint8_t *act_steps_d, void *stream);

/* out_reward_d [B] f64, out_done_d [B] i32, out_info_d [B] i32, out_dmin_d [B] f64 may each be NULL. */
int aem_env_apply(aem_handle h, int B, int n, double *robot_d, double *humans_d, const double *new_vel_d,
                  double *global_time_d, const int32_t *chosen_d, const double *rewards_d, const int32_t *info_d,
                  const double *dmin_d, double *out_reward_d, int32_t *out_done_d, int32_t *out_info_d,
                  double *out_dmin_d, void *stream);

/* The decision step in three launches: aem_env_prepare, aem_lookahead (best_d = NULL), aem_env_finish.
 * aem_env_prepare = aem_orca followed by aem_env_lookahead, bit for bit, in one kernel: every environment's CTA solves
 * ORCA for its own humans first (new_vel_d [B][n][2] is an OUTPUT here). */
int aem_env_prepare(aem_handle h, int B, int n, const double *robot_d, const double *humans_d, const double *global_time_d,
                    double *new_vel_d, double *rewards_d, int32_t *info_d, double *dmin_d, double *humans_next_d, void *stream);

/* aem_env_finish = the argmax of aem_lookahead over values_d [B][A] (-> best_d, may be NULL), aem_env_apply for that action
 * and, when pool_robot_d != NULL, aem_env_reset_done for the environments that finished -- one kernel, one warp per
 * environment, the same results as the three calls.  out_* [B] may each be NULL; without a pool pass NULL / 0 for the last six. */
int aem_env_finish(aem_handle h, int B, int n, const double *values_d, double *robot_d, double *humans_d,
                   const double *new_vel_d, double *global_time_d, const double *rewards_d, const int32_t *info_d,
                   const double *dmin_d, int32_t *best_d, double *out_reward_d, int32_t *out_done_d, int32_t *out_info_d,
                   double *out_dmin_d, const double *pool_robot_d, const double *pool_humans_d, int P, int32_t *case_idx_d,
                   int case_stride, unsigned long long *outcome_counts_d, void *stream);

/* Vectorised Explorer reset (explorer.py:44-47 + CrowdSim.reset crowd_sim.py:486-552): every env with done_d[b] != 0
 * is re-seeded from scenario `case_idx_d[b] mod P` of a host-generated pool (pool_robot_d [P][9], pool_humans_d
 * [P][n][8]), its clock is zeroed, case_idx_d[b] += case_stride, and outcome_counts_d[info_d[b]] (5 x u64, may be
 * NULL) is incremented. */
int aem_env_reset_done(aem_handle h, int B, int n, double *robot_d, double *humans_d, double *global_time_d,
                       const int32_t *done_d, const int32_t *info_d, const double *pool_robot_d,
                       const double *pool_humans_d, int P, int32_t *case_idx_d, int case_stride,
                       unsigned long long *outcome_counts_d, void *stream);

/* CrowdSim.reset's scenario draw (crowd_sim.py:486-552 with generate_circle_crossing_human :390-411 /
 * generate_square_crossing_human :413-442, agent.py:39-45) for `count` consecutive cases on the device: case c is seeded like
 * np.random.seed(first_seed + c) (first_seed = the phase's counter offset + first case, crowd_sim.py:501-505) and replays the
 * legacy MT19937 stream in the reference's call order, rejection sampling included.  rule: 0 = circle_crossing,
 * 1 = square_crossing.  -> robot_d [count][9] f64, humans_d [count][n][8] f64.  Square rule: bit-exact; circle rule: cos / sin
 * come from libdevice (<= 1 ulp from numpy's). */
int aem_generate_scenarios(aem_handle h, int rule, uint32_t first_seed, int count, int n, const aem_scenario_params *p,
                           double *robot_d, double *humans_d, void *stream);

/* The 'mixed' rule of the same draw (crowd_sim.py:337-386): a static scene with probability 0.2 (0..5 standing humans in a
 * 4 x 8 box, goal = start; none -> one dummy human parked at (0, -10)), else a dynamic one (1..5 humans, the first two cross
 * the circle, the others the square).  -> robot_d [count][9], humans_d [count][5][8] (rows beyond a case's human count are
 * padding: parked at (0, -100), zero velocity) and n_active_d [count] i32, the per-environment counts aem_set_active_counts
 * takes.  Static and square-crossing humans bit-exact, circle-crossing ones within an ulp of cos / sin. */
int aem_generate_mixed_scenarios(aem_handle h, uint32_t first_seed, int count, const aem_scenario_params *p, double *robot_d,
                                 double *humans_d, int32_t *n_active_d, void *stream);

/* states_d [S][n][13] f32 (rotated joint states) -> values_d [S] f32, steps_d [S] u8 (may be NULL) */
int aem_value_forward(aem_handle h, int S, int n, const float *states_d, float *values_d, uint8_t *steps_d,
                      void *stream);

/* x_d [S*T][13] f32, rows grouped in samples of T tokens -> out_d [S*T][50] f32, steps_d [S] u8 (may be NULL) */
int aem_gru_act(aem_handle h, int S, int T, const float *x_d, float *out_d, uint8_t *steps_d, void *stream);

/* CrowdSim.step(action) (crowd_sim.py:557-691) with ONE arbitrary ActionXY per environment -- what the imitation phase
 * needs, where the robot executes the continuous ORCA velocity (train.py:171-204, explorer.py:51-53).
 * actions_d [B][2] f64; max_speed >= every |action| (bounds the occupancy-map culling); scratch_d: 3 B doubles.
 * Updates robot_d / humans_d / global_time_d in place like aem_env_apply; out_* [B]. */
int aem_env_step_xy(aem_handle h, int B, int n, double *robot_d, double *humans_d, const double *new_vel_d,
                    double *global_time_d, const double *actions_d, double max_speed, double *scratch_d,
                    double *out_reward_d, int32_t *out_done_d, int32_t *out_info_d, double *out_dmin_d, void *stream);

/* MultiHumanRL.transform (multi_human_rl.py:423-438): rotate() of the current joint state of every environment, the
 * sample Explorer stores in the replay memory.  robot_d [B][9] f64, humans_d [B][n][8] f64 -> states_d [B][n][13] f32 */
int aem_transform(aem_handle h, int B, int n, const double *robot_d, const double *humans_d, float *states_d, void *stream);

/* Value-network regression step (crowd_nav/utils/trainer.py:47-58,70-82: outputs = model(inputs); loss = MSELoss(outputs,
 * values); loss.backward()): forward of ValueNetworkTransformerAndGRU (qnetwork_factory.py:641-687) on B rotated states with
 * the reference's BATCH-GLOBAL halting (components.py:132-134), mean squared error against targets_d, and the gradient with
 * respect to every live parameter.  Dropout follows aem_set_train_dropout (default: inactive, the eval()-mode network every
 * other entry point evaluates); with aem_set_act_mode(act_fixed = 1, act_steps <= 3) exactly act_steps GRU steps run and h_K
 * is handed on (components.py:109-112), nothing flows into ponder_linear.
 * params_d [AEM_NUM_WEIGHTS] f32 (blob order; the caller's live parameters, not the handle's copy), states_d [B][n][13] f32,
 * targets_d [B] f32 -> grad_d [AEM_NUM_WEIGHTS] f32 (overwritten; deterministic summation order), loss_d [1] f32,
 * values_d [B] f32, steps_d [1] i32 = the batch's GRU step count (the last three may be NULL).  Three kernel launches. */
int aem_train_grad(aem_handle h, int B, int n, const float *params_d, const float *states_d, const float *targets_d,
                   float *grad_d, float *loss_d, float *values_d, int32_t *steps_d, void *stream);

/* Per-environment human counts: the 'mixed' scenario rule of CrowdSim.reset draws 1..5 humans per case
 * (crowd_sim/envs/crowd_sim.py:337-386).  n_active_d[B] (int32, device; 1 <= count <= n) makes aem_orca, aem_env_lookahead,
 * aem_lookahead (precision 0 or 2), aem_env_apply and aem_env_step_xy treat rows >= count of every [B][n][..] array as
 * padding: no ORCA agent, no collision / occupancy contribution, no token of the value network (attention keys, halting vote).
 * pool_n_active_d[P] (optional) are the counts of the scenario pool: aem_env_reset_done copies the new case's count.
 * The pointers are remembered until the next call; NULL restores uniform batches. */
int aem_set_active_counts(aem_handle h, int32_t *n_active_d, const int32_t *pool_n_active_d);

/* Dropout of the fused training step.  The reference never calls eval(): nn.TransformerEncoderLayer(50, 2, 150) trains
 * with p = 0.1 at four sites per layer (attention weights, attention branch, FFN hidden, FFN branch;
 * crowd_nav/common/qnetwork_factory.py:628-629, trainer.py:65-86).  p > 0 makes aem_train_grad / aem_train_step_adam draw
 * inverted-dropout masks from a counter-based generator (one fresh set per call: seed, call counter -- for
 * aem_train_step_adam the device step count --, sample, layer, site, element) and use the same masks in the backward
 * pass; p = 0 (default) is the eval()-mode network.  Calling it resets the call counter. */
int aem_set_train_dropout(aem_handle h, double p, uint64_t seed);

/* One mini-batch of the replay memory (memory.py:4-28 read through DataLoader(memory, batch_size, shuffle=True), trainer.py:70-72):
 * k of the N stored pairs, uniformly at random WITHOUT replacement (Floyd's algorithm on a counter-based generator: the same
 * (seed, counter) gives the same batch).  states_d [N][n][13] f32, values_d [N] f32 -> out_states_d [k][n][13], out_values_d [k],
 * out_index_d [k] i64 (may be NULL). */
int aem_sample_batch(aem_handle h, int N, int k, int n, const float *states_d, const float *values_d, uint64_t seed,
                     uint64_t counter, float *out_states_d, float *out_values_d, int64_t *out_index_d, void *stream);

/* torch.optim.Adam.step() (trainer.py:30 uses betas (0.9, 0.99), eps 1e-3; no weight decay, no amsgrad) on flat vectors of
 * `count` floats: step = 1-based number of this update; the gradient is multiplied by grad_scale first (1 / world_size
 * after a summing all-reduce). */
int aem_adam_step(aem_handle h, size_t count, float *params_d, const float *grad_d, float *exp_avg_d, float *exp_avg_sq_d,
                  double lr, double beta1, double beta2, double eps, int step, double grad_scale, void *stream);

/* One iteration of Trainer.optimize_batch (trainer.py:66-86) with Adam, every per-step quantity resident on the device: draw a
 * mini-batch of k pairs from the replay ring (first N entries), aem_train_grad on it, Adam update.  dev_state_d: u64[2] =
 * {draws so far, optimizer steps so far}, both advanced by the call; adam_scal_d: f32[2] scratch (step size, bias correction);
 * loss_sum_d (may be NULL) accumulates the losses.  All arguments are constant from step to step, so the five launches can be
 * captured once in a CUDA graph and replayed.  batch_states_d [k][n][13] / batch_values_d [k]: the drawn batch (scratch). */
int aem_train_step_adam(aem_handle h, int N, int k, int n, const float *ring_states_d, const float *ring_values_d, uint64_t seed,
                        unsigned long long *dev_state_d, float *adam_scal_d, float *params_d, float *grad_d, float *exp_avg_d,
                        float *exp_avg_sq_d, double lr, double beta1, double beta2, double eps, float *batch_states_d,
                        float *batch_values_d, float *loss_d, float *loss_sum_d, int32_t *steps_d, void *stream);

/* The other two optimizers Trainer.set_learning_rate offers (trainer.py:24-34): optim.SGD(momentum = 0.9) -- momentum_buf_d is
 * initialised by the step == 1 call -- and optim.RMSprop(alpha = 0.9, eps = 1e-8; not centered, no momentum). */
int aem_sgd_step(aem_handle h, size_t count, float *params_d, const float *grad_d, float *momentum_buf_d, double lr,
                 double momentum, int step, double grad_scale, void *stream);
int aem_rmsprop_step(aem_handle h, size_t count, float *params_d, const float *grad_d, float *square_avg_d, double lr,
                     double alpha, double eps, double grad_scale, void *stream);

/* ---- host-buffer call: one full env-step (predict + step) for B environments ----
 * In/out (updated in place): robot_h [B][9], humans_h [B][n][8], global_time_h [B].
 * Out: chosen_h [B] i32, reward_h [B] f64, done_h [B] i32, info_h [B] i32; values_h [B][A] f64 may be NULL.
 * Copies host->device, runs orca + env_lookahead + lookahead + env_apply, copies back, synchronises. */
int aem_decide_step_host(aem_handle h, int B, int n, double gamma_bar, double *robot_h, double *humans_h,
                         double *global_time_h, int32_t *chosen_h, double *reward_h, int32_t *done_h,
                         int32_t *info_h, double *values_h);

/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
long long aem_launch_count(aem_handle h);

/* The scratch of aem_train_grad / aem_train_step_adam belongs to the handle and grows with the largest (batch, n) seen;
 * this counter increments whenever it is reallocated.  A caller that captured aem_train_step_adam into a CUDA graph must
 * re-capture when the value has changed (the graph holds the old pointers): crowd_nav/utils/trainer.py FusedStep. */
long long aem_train_scratch_generation(aem_handle h);

#ifdef __cplusplus
}
#endif
#endif
