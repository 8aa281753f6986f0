// aem_common.cuh -- shared definitions of libaemcarl_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aemcarl_b200.h"

namespace aem {

constexpr int HID = 50;    // GRU hidden / d_model
constexpr int XD = 13;     // rotated joint-state length (cadrl.py:239-274)
constexpr int HXD = 63;    // [h ; x]

// ---- offsets into the raw weight blob (state_dict order, aemcarl_b200/weights.py) ----
constexpr int OFF_Z1W = 0;
constexpr int OFF_Z1B = OFF_Z1W + 100 * 63;
constexpr int OFF_Z2W = OFF_Z1B + 100;
constexpr int OFF_Z2B = OFF_Z2W + 50 * 100;
constexpr int OFF_R1W = OFF_Z2B + 50;
constexpr int OFF_R1B = OFF_R1W + 100 * 63;
constexpr int OFF_R2W = OFF_R1B + 100;
constexpr int OFF_R2B = OFF_R2W + 50 * 100;
constexpr int OFF_QW = OFF_R2B + 50;
constexpr int OFF_QB = OFF_QW + 50 * 63;
constexpr int OFF_PW = OFF_QB + 50;
constexpr int OFF_PB = OFF_PW + 50;
constexpr int OFF_ENC = OFF_PB + 1;
constexpr int ENC_INW = 0;
constexpr int ENC_INB = ENC_INW + 150 * 50;
constexpr int ENC_OW = ENC_INB + 150;
constexpr int ENC_OB = ENC_OW + 50 * 50;
constexpr int ENC_L1W = ENC_OB + 50;
constexpr int ENC_L1B = ENC_L1W + 150 * 50;
constexpr int ENC_L2W = ENC_L1B + 150;
constexpr int ENC_L2B = ENC_L2W + 50 * 150;
constexpr int ENC_N1G = ENC_L2B + 50;
constexpr int ENC_N1B = ENC_N1G + 50;
constexpr int ENC_N2G = ENC_N1B + 50;
constexpr int ENC_N2B = ENC_N2G + 50;
constexpr int ENC_SIZE = ENC_N2B + 50;
constexpr int OFF_A0W = OFF_ENC + 3 * ENC_SIZE;
constexpr int OFF_A0B = OFF_A0W + 100 * 56;
constexpr int OFF_A2W = OFF_A0B + 100;
constexpr int OFF_A2B = OFF_A2W + 50 * 100;
constexpr int OFF_A4W = OFF_A2B + 50;
constexpr int OFF_A4B = OFF_A4W + 50;
constexpr int NUM_WEIGHTS = OFF_A4B + 1;
static_assert(NUM_WEIGHTS == AEM_NUM_WEIGHTS, "weight blob size");
static_assert(ENC_SIZE == 25600, "encoder layer size");

// ---- packed GEMM operand layout ----
// A packed layer is W^T re-tiled for the register-tiled FP32 GEMM of lookahead.cu:
//   packed[k][g][c], k < K, g < 16 column groups, c < TNHP (TNH real columns padded to a multiple of 4)
// plain layer : real column = g * TNH + c
// two-branch  : branch = g >> 3, real column inside the branch = (g & 7) * TNH + c
constexpr int NGROUPS = 16;
constexpr int pad4(int x) { return (x + 3) & ~3; }

// ---- tensor-core path (lookahead_tc.cu): one entry per tcgen05 GEMM ----
// images: img[k/4][n][4] floats (canonical K-major no-swizzle UMMA layout), hi part and lo part, N and K permuted /
// padded as capi.cu pack_tc documents.
struct TcLayer { int off_hi, off_lo, bytes, N, ksteps, bias_off; };
enum { TL_Z1S = 0, TL_QS = 1, TL_Z1 = 2, TL_Z2 = 3, TL_R1 = 4, TL_R2 = 5, TL_Q = 6,
       TL_ENC = 7,      // + 7 * l : INA, INB, OUT, F1A, F1B, F2A, F2B
       TL_A0 = 28, TL_A2 = 29, TL_NUM = 30 };
constexpr int TC_PBLK_FLOATS = 3200;
constexpr int H16_PBLK_FLOATS = 800;       // LayerNorm 672, ponder 57, action_mlp.4 57 as rows of 28 (capi.cu pack_h16)

struct NetParams {
    const float *raw;      // raw blob (biases, LayerNorm, ponder, last action layer)
    const float *packed;   // packed GEMM operands
    int off_z1r1, off_z2r2, off_q;
    int off_in[3], off_out[3], off_f1[3], off_f2[3];
    int off_a0, off_a2;
    int act_fixed, act_steps;
    // tensor-core path
    const float *tc_img;   // weight images
    const float *tc_pblk;  // permuted biases, LayerNorm, ponder and last-layer parameters
    int tc_pblk_floats, tc_pb_ln, tc_pb_pw, tc_pb_a4;
    const TcLayer *tc_layers_d;   // [TL_NUM] in global memory
    // FP16x3 tensor-core path (lookahead_h16.cu); TcLayer offsets are in fp16 elements
    const __half *h16_img;
    const float *h16_pblk;
    int h16_pblk_floats;
    const TcLayer *h16_layers_d;
};

// tile sizes of the packed layers (shared by the host packer and the kernel)
constexpr int TNH_Z1R1 = 13;  // two-branch, 100 + 100 columns, K = 63
constexpr int TNH_Z2R2 = 7;   // two-branch, 50 + 50 columns, K = 100
constexpr int TNH_50 = 4;     // N = 50 layers (q, out_proj, linear2, action_mlp.2)
constexpr int TNH_150 = 10;   // N = 150 layers (in_proj, linear1)
constexpr int TNH_100 = 7;    // N = 100 layer (action_mlp.0)

}  // namespace aem

// ---- handle ----
struct aem_handle_s {
    int device;
    int num_sms;
    float *raw_d;        // [NUM_WEIGHTS]
    float *packed_d;
    size_t packed_floats;
    float *tc_img_d, *tc_pblk_d;
    aem::TcLayer *tc_layers_d;
    __half *h16_img_d;
    float *h16_pblk_d;
    aem::TcLayer *h16_layers_d;
    size_t h16_img_halves;
    size_t tc_img_floats;
    int precision;       // 0 = FP32 FFMA kernel, 1 = tcgen05 3xTF32 kernel, 2 = tcgen05 FP16x3 kernel (2 tiles / SM), 3 = its single-pass mode
    aem::NetParams net;
    bool weights_set, actions_set, params_set;
    double *actions_d;   // [A][2]
    double actions_h[AEM_MAX_ACTIONS * 2];
    int A;
    aem_env_params params;
    double *lut_d;       // [10000] normal pdf table (crowd_sim.py:20-28)
    long long launches;
    // scratch for the host-buffer call
    int cap_B, cap_n;
    double *s_robot, *s_humans, *s_time, *s_newvel, *s_rewards, *s_dmin, *s_hnext, *s_values, *s_oreward;
    int32_t *s_info, *s_best, *s_odone, *s_oinfo;
    void *s_stream;
    // scratch of the training step (train.cu): per-CTA gradient slices, squared errors, halting flags, activation stash
    float *tr_part, *tr_sqerr, *tr_stash, *tr_gru;
    int *tr_flags;
    int tr_grid, tr_B;
    size_t tr_stash_floats, tr_gru_floats;
    int32_t *nact_d;               // per-environment human counts (aem_set_active_counts), or null: every env has n humans
    const int32_t *pool_nact_d;    // the same for the scenario pool of aem_env_reset_done
    float *tail_d;                 // lookahead_h16: staged token-0 rows
    size_t tail_floats;
    double drop_p;                 // dropout of the fused training step (aem_set_train_dropout); 0 = eval() mode
    unsigned long long drop_seed, drop_ctr;
    long long tr_generation;   // bumped whenever the training scratch is reallocated (captured graphs hold its pointers)
};

#define AEM_CUDA_CHECK(expr)                                                      \
    do {                                                                          \
        cudaError_t e__ = (expr);                                                 \
        if (e__ != cudaSuccess) return aem_set_cuda_error(e__, #expr, __FILE__, __LINE__); \
    } while (0)

int aem_set_cuda_error(cudaError_t e, const char *what, const char *file, int line);
int aem_set_error(int code, const char *msg);

// launchers implemented in the .cu files
namespace aem {
struct LookaheadArgs;
cudaError_t launch_lookahead(aem_handle h, const LookaheadArgs &a, cudaStream_t st);
cudaError_t launch_lookahead_tc(aem_handle h, const LookaheadArgs &a, cudaStream_t st);
cudaError_t launch_lookahead_h16(aem_handle h, const LookaheadArgs &a, cudaStream_t st, bool single_pass = false);
cudaError_t launch_transform(const double *robot, const double *humans, int B, int n, float *out, cudaStream_t st);
cudaError_t launch_select(const double *values, const double *robot, int B, int A, int32_t *best, cudaStream_t st);
cudaError_t launch_orca(int B, int n, const double *humans, const double *robot, const aem_env_params &p,
                        double *new_vel, cudaStream_t st, const int32_t *nact = nullptr);
cudaError_t launch_env_lookahead(int B, int n, int A, const double *robot, const double *humans,
                                 const double *new_vel, const double *global_time, const double *actions,
                                 int action_env_stride, const double *lut, const aem_env_params &p, double amax, double *rewards,
                                 int32_t *info, double *dmin, double *humans_next, cudaStream_t st, const int32_t *nact = nullptr,
                                 double *new_vel_out = nullptr);     // new_vel_out: ORCA runs inside the same launch (env_prepare)
cudaError_t launch_env_finish(int B, int n, int A, const double *values, double *robot, double *humans, const double *new_vel,
                              double *global_time, const double *actions, int action_env_stride, const double *rewards,
                              const int32_t *info, const double *dmin, double time_step, int32_t *best, double *out_reward,
                              int32_t *out_done, int32_t *out_info, double *out_dmin, int32_t *nact, const double *pool_robot,
                              const double *pool_humans, int P, int32_t *case_idx, int case_stride,
                              unsigned long long *outcome_counts, const int32_t *pool_nact, cudaStream_t st);
cudaError_t launch_env_apply(int B, int n, int A, double *robot, double *humans, const double *new_vel,
                             double *global_time, const double *actions, int action_env_stride,
                             const int32_t *chosen,
                             const double *rewards, const int32_t *info, const double *dmin, double time_step,
                             double *out_reward, int32_t *out_done, int32_t *out_info, double *out_dmin,
                             cudaStream_t st, const int32_t *nact = nullptr);

cudaError_t launch_env_reset(int B, int n, double *robot, double *humans, double *global_time, const int32_t *done,
                             const int32_t *info, const double *pool_robot, const double *pool_humans, int P,
                             int32_t *case_idx, int case_stride, unsigned long long *outcome_counts, cudaStream_t st,
                             int32_t *nact = nullptr, const int32_t *pool_nact = nullptr);

cudaError_t launch_scenarios(int rule, uint32_t first_seed, int count, int n, const aem_scenario_params &p, double *robot,
                             double *humans, cudaStream_t st, int32_t *n_active = nullptr);
cudaError_t launch_train_grad(aem_handle h, int B, int n, const float *params, const float *states, const float *targets,
                              float *grad, float *loss, float *values, int32_t *steps, cudaStream_t st, float *loss_sum = nullptr,
                              const unsigned long long *drop_ctr_dev = nullptr);
cudaError_t launch_train_step_adam(aem_handle h, int N, int k, int n, const float *ring_states, const float *ring_values, uint64_t seed,
                                   unsigned long long *dev_state, float *adam_scal, float *params, float *grad, float *m, float *v,
                                   double lr, double b1, double b2, double eps, float *batch_states, float *batch_values, float *loss,
                                   float *loss_sum, int32_t *steps, cudaStream_t st);
cudaError_t launch_sgd(size_t count, float *p, const float *g, float *buf, double lr, double momentum, int step, double gscale,
                       cudaStream_t st);
cudaError_t launch_rmsprop(size_t count, float *p, const float *g, float *sq, double lr, double alpha, double eps, double gscale,
                           cudaStream_t st);
cudaError_t launch_sample_batch(int N, int k, int row_floats, const float *states, const float *values, uint64_t seed,
                                uint64_t counter, float *out_states, float *out_values, int64_t *out_index, cudaStream_t st);
cudaError_t launch_adam(size_t count, float *p, const float *g, float *m, float *v, double lr, double b1, double b2, double eps,
                        int step, double gscale, cudaStream_t st);

struct LookaheadArgs {
    NetParams net;
    int mode;            // 0 = lookahead from env state, 1 = rotated states, 2 = GRU-ACT only
    int n, T, S, A;      // humans, tokens per sample, samples per tile, actions
    int total_samples;
    double dt, gamma_bar;
    const double *robot, *humans_next, *actions, *rewards;
    const float *states;   // mode 1: [S][n][13]; mode 2: [S*T][13]
    double *values;
    float *net_values;
    uint8_t *steps;
    float *gru_out;        // mode 2: [S*T][50]
    const int32_t *n_active;   // mode 0, optional: humans of every environment [B] ('mixed' scenes); token rows beyond it are padding
    float *tail;           // lookahead_h16: per-CTA scratch of staged token-0 rows [grid][128][TAIL_ROW]
    long long *dbg;        // optional debug buffer (lookahead_h16.cu: AEM_H16_TIMELINE / AEM_H16_WATCHDOG builds)
    int dbg_tile;          // timeline build: which of CTA 0's tiles is stamped
    int stagger_ns;        // lookahead_h16: start-up delay of the second CTA of every SM (anti-phase the two tiles)
};
}  // namespace aem
