/*
 * aem_oracle.c -- CPU restatement (ORACLE) of the AEMCARL lookahead hot path.
 * TEST INFRASTRUCTURE ONLY (see aem_oracle.h).  Plain C99 + pthreads.
 *
 * Reference citations are file:line in SJWang2015/AEMCARL.
 */
#include "aem_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NA AEMO_NUM_ACTIONS
#define H 50
#define XD 13
#define HX 63
#define MAXT (AEMO_MAX_HUMANS + 1)

int aemo_num_weights(void) { return AEMO_NUM_WEIGHTS; }

/* ------------------------------------------------------------------ */
/* tiny pthread parallel-for (dynamic chunks of 1) -- no OpenMP dep    */
/* ------------------------------------------------------------------ */
#include <pthread.h>
typedef void (*pf_body)(int i, void *ctx);
typedef struct { pf_body body; void *ctx; int n; volatile int next; } pf_job;
static void *pf_worker(void *arg)
{
    pf_job *j = (pf_job *)arg;
    for (;;) {
        int i = __sync_fetch_and_add(&j->next, 1);
        if (i >= j->n) break;
        j->body(i, j->ctx);
    }
    return NULL;
}
static void parallel_for(int n, int nthreads, pf_body body, void *ctx)
{
    if (nthreads > 256) nthreads = 256;
    if (nthreads <= 1 || n <= 1) { for (int i = 0; i < n; ++i) body(i, ctx); return; }
    pf_job job; job.body = body; job.ctx = ctx; job.n = n; job.next = 0;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads - 1; ++t)
        if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) ++started;
    pf_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}


/* ------------------------------------------------------------------ */
/* cadrl.py:129-154  build_action_space (holonomic branch)             */
/* ------------------------------------------------------------------ */
void aemo_build_action_space(double v_pref, int speed_samples, int rotation_samples, double *out)
{
    /* speeds = (exp((i+1)/S) - 1) / (e - 1) * v_pref ; rotations = linspace(0, 2pi, R, endpoint=False)
     * action_space = [(0,0)] + product(rotations, speeds)  (rotation-major) */
    const double e = 2.718281828459045; /* np.e */
    const double two_pi = 2.0 * 3.141592653589793;
    int idx = 0;
    out[0] = 0.0; out[1] = 0.0; idx = 1;
    for (int r = 0; r < rotation_samples; ++r) {
        /* np.linspace(0, 2*pi, R, endpoint=False): step = 2pi/R ; value = r*step */
        double step = two_pi / (double)rotation_samples;
        double rot = (double)r * step;
        for (int s = 0; s < speed_samples; ++s) {
            double speed = (exp((double)(s + 1) / (double)speed_samples) - 1.0) / (e - 1.0) * v_pref;
            out[2 * idx + 0] = speed * cos(rot);
            out[2 * idx + 1] = speed * sin(rot);
            ++idx;
        }
    }
}

/* ------------------------------------------------------------------ */
/* cadrl.py:239-274  rotate (FP32, holonomic: theta' = 0)              */
/* ------------------------------------------------------------------ */
void aemo_rotate(const float *s, float *o)
{
    /* 'px','py','vx','vy','radius','gx','gy','v_pref','theta','px1','py1','vx1','vy1','radius1' */
    float dx = s[5] - s[0];
    float dy = s[6] - s[1];
    float rot = atan2f(dy, dx);
    float c = cosf(rot), sn = sinf(rot);
    o[0] = sqrtf(dx * dx + dy * dy);                   /* dg */
    o[1] = s[7];                                       /* v_pref */
    o[2] = 0.0f;                                       /* theta (holonomic) */
    o[3] = s[4];                                       /* radius */
    o[4] = s[2] * c + s[3] * sn;                       /* vx */
    o[5] = s[3] * c - s[2] * sn;                       /* vy */
    o[6] = (s[9] - s[0]) * c + (s[10] - s[1]) * sn;    /* px1 */
    o[7] = (s[10] - s[1]) * c - (s[9] - s[0]) * sn;    /* py1 */
    o[8] = s[11] * c + s[12] * sn;                     /* vx1 */
    o[9] = s[12] * c - s[11] * sn;                     /* vy1 */
    o[10] = s[13];                                     /* radius1 */
    {
        float ax = s[0] - s[9], ay = s[1] - s[10];
        o[11] = sqrtf(ax * ax + ay * ay);              /* da */
    }
    o[12] = s[4] + s[13];                              /* radius sum */
}

/* ------------------------------------------------------------------ */
/* prepared (transposed) weights so the inner loops vectorise over the */
/* output index while each output keeps a fixed k-ascending FMA order. */
/* ------------------------------------------------------------------ */
typedef struct {
    float z1[HX][100], r1[HX][100];   /* [k][j] */
    float z2[100][H], r2[100][H];
    float q[HX][H];
    float in[3][H][150], out[3][H][H], l1[3][H][150], l2[3][150][H];
    float a0[56][100], a2[100][H];
    const float *W;                   /* original blob for biases etc. */
} prep_t;

static void transpose_into(const float *w, int out, int in, float *dst /* [in][out] */)
{
    for (int j = 0; j < out; ++j)
        for (int k = 0; k < in; ++k) dst[k * out + j] = w[j * in + k];
}

static prep_t *prep_create(const float *W)
{
    prep_t *p = (prep_t *)malloc(sizeof(prep_t));
    p->W = W;
    transpose_into(W + AEMO_OFF_Z1W, 100, HX, &p->z1[0][0]);
    transpose_into(W + AEMO_OFF_R1W, 100, HX, &p->r1[0][0]);
    transpose_into(W + AEMO_OFF_Z2W, H, 100, &p->z2[0][0]);
    transpose_into(W + AEMO_OFF_R2W, H, 100, &p->r2[0][0]);
    transpose_into(W + AEMO_OFF_QW, H, HX, &p->q[0][0]);
    for (int l = 0; l < 3; ++l) {
        const float *E = W + AEMO_OFF_ENC + l * AEMO_ENC_SIZE;
        transpose_into(E + AEMO_ENC_INW, 150, H, &p->in[l][0][0]);
        transpose_into(E + AEMO_ENC_OW, H, H, &p->out[l][0][0]);
        transpose_into(E + AEMO_ENC_L1W, 150, H, &p->l1[l][0][0]);
        transpose_into(E + AEMO_ENC_L2W, H, 150, &p->l2[l][0][0]);
    }
    transpose_into(W + AEMO_OFF_A0W, 100, 56, &p->a0[0][0]);
    transpose_into(W + AEMO_OFF_A2W, H, 100, &p->a2[0][0]);
    return p;
}

/* y[j] = (sum_k x[k] * Wt[k][j]) + b[j], k ascending, fused multiply-add */
static inline void dense(const float *x, int K, const float *Wt, int N, const float *b, float *y)
{
    float acc[200];
    for (int j = 0; j < N; ++j) acc[j] = 0.0f;
    for (int k = 0; k < K; ++k) {
        const float xk = x[k];
        const float *w = Wt + (size_t)k * N;
        for (int j = 0; j < N; ++j) acc[j] = fmaf(xk, w[j], acc[j]);
    }
    for (int j = 0; j < N; ++j) y[j] = acc[j] + b[j];
}

static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
static inline float reluf_(float x) { return x > 0.0f ? x : 0.0f; }

/* components.py:46-53  GRUEXND.forward for one token */
static void gru_cell(const prep_t *p, const float *x, float *h /* in/out [50] */)
{
    const float *W = p->W;
    float hx[HX], t1[100], z[H], r[H], q[H], rhx[HX];
    memcpy(hx, h, H * sizeof(float));
    memcpy(hx + H, x, XD * sizeof(float));
    /* z = sigmoid(convz(hx)); convz = Linear, ReLU, Linear, ReLU (last_relu=True) */
    dense(hx, HX, &p->z1[0][0], 100, W + AEMO_OFF_Z1B, t1);
    for (int j = 0; j < 100; ++j) t1[j] = reluf_(t1[j]);
    dense(t1, 100, &p->z2[0][0], H, W + AEMO_OFF_Z2B, z);
    for (int j = 0; j < H; ++j) z[j] = sigmoidf_(reluf_(z[j]));
    dense(hx, HX, &p->r1[0][0], 100, W + AEMO_OFF_R1B, t1);
    for (int j = 0; j < 100; ++j) t1[j] = reluf_(t1[j]);
    dense(t1, 100, &p->r2[0][0], H, W + AEMO_OFF_R2B, r);
    for (int j = 0; j < H; ++j) r[j] = sigmoidf_(reluf_(r[j]));
    /* q = tanh(convq(cat[r*h, x])) */
    for (int j = 0; j < H; ++j) rhx[j] = r[j] * h[j];
    memcpy(rhx + H, x, XD * sizeof(float));
    dense(rhx, HX, &p->q[0][0], H, W + AEMO_OFF_QB, q);
    for (int j = 0; j < H; ++j) {
        float qq = tanhf(q[j]);
        h[j] = (1.0f - z[j]) * h[j] + z[j] * qq;      /* h = (1 - z) * h + z * q */
    }
}

/* components.py:107-137  ATCBasic.forward for ONE sample of T tokens
 * (predict() feeds one (env, action) sample per forward, so the batch-global
 *  `selector.any()` of the reference is exactly a per-sample test). */
static __thread float g_act_margin;   /* min |accum_p - 0.95| seen by the halting tests of the last sample */

static int gru_act_sample(const prep_t *p, const float *x /*[T][13]*/, int T, int act_fixed, int act_steps,
                          float *out /*[T][50]*/)
{
    g_act_margin = INFINITY;
    const float *W = p->W;
    float h[MAXT][H], acc[MAXT][H], P[MAXT];
    memset(h, 0, sizeof(h));
    if (act_fixed) {
        for (int s = 0; s < act_steps; ++s)
            for (int t = 0; t < T; ++t) gru_cell(p, x + t * XD, h[t]);
        for (int t = 0; t < T; ++t) memcpy(out + t * H, h[t], H * sizeof(float));
        return act_steps;
    }
    memset(acc, 0, sizeof(acc));
    memset(P, 0, sizeof(P));
    int cnt = 0;
    for (int s = 0; s < 3; ++s) {                     /* max_ponder = 3 */
        int any = 0;
        for (int t = 0; t < T; ++t) {
            gru_cell(p, x + t * XD, h[t]);
            float a = 0.0f;
            for (int k = 0; k < H; ++k) a = fmaf(h[t][k], W[AEMO_OFF_PW + k], a);
            float hp = sigmoidf_(a + W[AEMO_OFF_PB]);  /* halten_p */
            P[t] += hp;
            for (int k = 0; k < H; ++k) acc[t][k] += hp * h[t][k];
            if (P[t] < 0.95f) any = 1;                /* accum_p < 1 - epsilon (fp32 compare) */
            if (fabsf(P[t] - 0.95f) < g_act_margin) g_act_margin = fabsf(P[t] - 0.95f);
        }
        ++cnt;
        if (!any) break;
    }
    for (int t = 0; t < T; ++t)
        for (int k = 0; k < H; ++k) out[t * H + k] = acc[t][k] / (float)cnt;
    return cnt;
}

static void layer_norm(float *x, const float *g, const float *b)
{
    float mean = 0.0f, var = 0.0f;
    for (int k = 0; k < H; ++k) mean += x[k];
    mean /= (float)H;
    for (int k = 0; k < H; ++k) { float d = x[k] - mean; var += d * d; }
    var /= (float)H;
    float rstd = 1.0f / sqrtf(var + 1e-5f);
    for (int k = 0; k < H; ++k) x[k] = (x[k] - mean) * rstd * g[k] + b[k];
}

/* torch nn.TransformerEncoderLayer(50, 2, 150), post-norm, relu, eval mode
 * (qnetwork_factory.py:628-629, used :671-677).  last != 0: only token 0 is needed downstream. */
static void encoder_layer(const prep_t *p, int l, float e[][H], int T, int last)
{
    const float *E = p->W + AEMO_OFF_ENC + l * AEMO_ENC_SIZE;
    float qkv[MAXT][150], att[MAXT][H], tmp[150], y[H];
    for (int t = 0; t < T; ++t) dense(e[t], H, &p->in[l][0][0], 150, E + AEMO_ENC_INB, qkv[t]);
    int TQ = last ? 1 : T;
    for (int i = 0; i < TQ; ++i) {
        for (int hd = 0; hd < 2; ++hd) {
            float s[MAXT], m = -INFINITY, sum = 0.0f;
            for (int j = 0; j < T; ++j) {
                float d = 0.0f;
                for (int k = 0; k < 25; ++k) d = fmaf(qkv[i][hd * 25 + k], qkv[j][H + hd * 25 + k], d);
                s[j] = d * 0.2f;                      /* 1/sqrt(head_dim=25) */
                if (s[j] > m) m = s[j];
            }
            for (int j = 0; j < T; ++j) { s[j] = expf(s[j] - m); sum += s[j]; }
            for (int k = 0; k < 25; ++k) {
                float o = 0.0f;
                for (int j = 0; j < T; ++j) o = fmaf(s[j] / sum, qkv[j][2 * H + hd * 25 + k], o);
                att[i][hd * 25 + k] = o;
            }
        }
    }
    for (int i = 0; i < TQ; ++i) {
        dense(att[i], H, &p->out[l][0][0], H, E + AEMO_ENC_OB, y);
        for (int k = 0; k < H; ++k) e[i][k] += y[k];
        layer_norm(e[i], E + AEMO_ENC_N1G, E + AEMO_ENC_N1B);
        dense(e[i], H, &p->l1[l][0][0], 150, E + AEMO_ENC_L1B, tmp);
        for (int k = 0; k < 150; ++k) tmp[k] = reluf_(tmp[k]);
        dense(tmp, 150, &p->l2[l][0][0], H, E + AEMO_ENC_L2B, y);
        for (int k = 0; k < H; ++k) e[i][k] += y[k];
        layer_norm(e[i], E + AEMO_ENC_N2G, E + AEMO_ENC_N2B);
    }
}

/* qnetwork_factory.py:641-687 for one sample; state [n][13] */
static float value_forward_prepared(const prep_t *p, const float *state, int n, int act_fixed, int act_steps,
                                    int *steps)
{
    const float *W = p->W;
    int T = n + 1;
    float x[MAXT][XD], e[MAXT][H];
    /* robot pseudo token (:651-659) from human row 0 */
    memcpy(x[0], state, XD * sizeof(float));
    x[0][6] = 0.0f; x[0][7] = 0.0f; x[0][8] = x[0][4]; x[0][9] = x[0][5];
    x[0][10] = x[0][3]; x[0][11] = 0.0f; x[0][12] = x[0][3];
    memcpy(x[1], state, (size_t)n * XD * sizeof(float));
    int cnt = gru_act_sample(p, &x[0][0], T, act_fixed, act_steps, &e[0][0]);
    if (steps) *steps = cnt;
    for (int l = 0; l < 3; ++l) encoder_layer(p, l, e, T, l == 2);
    float joint[56], a1[100], a2[H];
    memcpy(joint, state, 6 * sizeof(float));          /* self_state = state[:, 0, :6] */
    memcpy(joint + 6, e[0], H * sizeof(float));       /* env_info = encoder output token 0 */
    dense(joint, 56, &p->a0[0][0], 100, W + AEMO_OFF_A0B, a1);
    for (int k = 0; k < 100; ++k) a1[k] = reluf_(a1[k]);
    dense(a1, 100, &p->a2[0][0], H, W + AEMO_OFF_A2B, a2);
    for (int k = 0; k < H; ++k) a2[k] = reluf_(a2[k]);
    float v = 0.0f;
    for (int k = 0; k < H; ++k) v = fmaf(a2[k], W[AEMO_OFF_A4W + k], v);
    return v + W[AEMO_OFF_A4B];
}

void aemo_value_forward(const float *W, const float *state, int n, int act_fixed, int act_steps, float *value,
                        int *steps)
{
    prep_t *p = prep_create(W);
    *value = value_forward_prepared(p, state, n, act_fixed, act_steps, steps);
    free(p);
}

/* distance of the halting decision from its threshold: min over tokens and executed steps of |accum_p - (1 - eps)|.
 * The ACT step count (components.py:132-134) is a hard threshold, so an implementation with different rounding may
 * legitimately take a different number of steps when this margin is at rounding level (documented near-threshold rule). */
float aemo_act_margin(const float *W, const float *state, int n)
{
    prep_t *p = prep_create(W);
    int steps = 0;
    (void)value_forward_prepared(p, state, n, 0, 1, &steps);
    free(p);
    return g_act_margin;
}

typedef struct { const prep_t *p; int n; const float *states; int act_fixed, act_steps; float *values; uint8_t *steps; } vfb_ctx;
static void vfb_body(int s, void *vc)
{
    vfb_ctx *c = (vfb_ctx *)vc;
    int k = 0;
    c->values[s] = value_forward_prepared(c->p, c->states + (size_t)s * c->n * XD, c->n, c->act_fixed, c->act_steps, &k);
    if (c->steps) c->steps[s] = (uint8_t)k;
}

void aemo_value_forward_batch(const float *W, int S, int n, const float *states, int act_fixed, int act_steps,
                              float *values, uint8_t *steps, int nthreads)
{
    prep_t *p = prep_create(W);
    vfb_ctx c = { p, n, states, act_fixed, act_steps, values, steps };
    parallel_for(S, nthreads, vfb_body, &c);
    free(p);
}

void aemo_gru_act(const float *W, int S, int T, const float *x, int act_fixed, int act_steps, float *out,
                  uint8_t *steps)
{
    prep_t *p = prep_create(W);
    for (int s = 0; s < S; ++s) {
        int c = gru_act_sample(p, x + (size_t)s * T * XD, T, act_fixed, act_steps, out + (size_t)s * T * H);
        if (steps) steps[s] = (uint8_t)c;
    }
    free(p);
}

/* multi_human_rl.py:336-372 */
typedef struct {
    const prep_t *p; int n, A; const double *robot, *humans_next, *rewards, *actions; double dt, gamma_bar;
    int act_fixed, act_steps; double *values; float *net_values; int32_t *best; uint8_t *steps;
} la_ctx;
static void la_body(int b, void *vc)
{
    la_ctx *c = (la_ctx *)vc;
    const int n = c->n, A = c->A;
    const double *actions = c->actions;
    const double *rb = c->robot + (size_t)b * 9;
    const double *hb = c->humans_next + (size_t)b * n * 5;
    double max_value = -INFINITY;
    int max_idx = -1;
    for (int a = 0; a < A; ++a) {
        /* cadrl.py:156-181 propagate (holonomic): p' = p + a*dt, v' = a */
        double npx = rb[0] + actions[2 * a] * c->dt;
        double npy = rb[1] + actions[2 * a + 1] * c->dt;
        float st[AEMO_MAX_HUMANS][XD];
        for (int i = 0; i < n; ++i) {
            float in14[14];
            /* torch.Tensor([next_self_state + next_human_state]) : f64 -> f32 */
            in14[0] = (float)npx; in14[1] = (float)npy;
            in14[2] = (float)actions[2 * a]; in14[3] = (float)actions[2 * a + 1];
            in14[4] = (float)rb[4]; in14[5] = (float)rb[5]; in14[6] = (float)rb[6];
            in14[7] = (float)rb[7]; in14[8] = (float)rb[8];
            for (int k = 0; k < 5; ++k) in14[9 + k] = (float)hb[i * 5 + k];
            aemo_rotate(in14, st[i]);
        }
        int k = 0;
        float v = value_forward_prepared(c->p, &st[0][0], n, c->act_fixed, c->act_steps, &k);
        double val = c->rewards[(size_t)b * A + a] + c->gamma_bar * (double)v;
        c->values[(size_t)b * A + a] = val;
        if (c->net_values) c->net_values[(size_t)b * A + a] = v;
        if (c->steps) c->steps[(size_t)b * A + a] = (uint8_t)k;
        if (val > max_value) { max_value = val; max_idx = a; }   /* strict '>' (:370) */
    }
    c->best[b] = max_idx;
}

void aemo_lookahead(const float *W, int B, int n, int A, const double *robot, const double *humans_next,
                    const double *rewards, const double *actions, double dt, double gamma_bar, int act_fixed,
                    int act_steps, double *values, float *net_values, int32_t *best, uint8_t *steps, int nthreads)
{
    prep_t *p = prep_create(W);
    la_ctx c = { p, n, A, robot, humans_next, rewards, actions, dt, gamma_bar, act_fixed, act_steps,
                 values, net_values, best, steps };
    parallel_for(B, nthreads, la_body, &c);
    free(p);
}

/* ================================================================== */
/* ORCA -- published RVO2 v2.0 algorithm (Agent.cpp), FP32.            */
/* Called as in crowd_sim/envs/policy/orca.py:82-132.                  */
/* PARITY UNPINNED against the real library (not available).           */
/* ================================================================== */
#define RVO_EPSILON 0.00001f
typedef struct { float x, y; } v2;
typedef struct { v2 point, direction; } line_t;

static inline v2 V(float x, float y) { v2 r; r.x = x; r.y = y; return r; }
static inline v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 vmul(float s, v2 a) { return V(s * a.x, s * a.y); }
static inline v2 vdiv(v2 a, float s) { float inv = 1.0f / s; return V(a.x * inv, a.y * inv); } /* Vector2::operator/ */
static inline float vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline float vdet(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
static inline float vabssq(v2 a) { return vdot(a, a); }
static inline float vabs(v2 a) { return sqrtf(vdot(a, a)); }
static inline v2 vnormalize(v2 a) { return vdiv(a, vabs(a)); }
static inline float sqrf(float a) { return a * a; }

static int lp1(const line_t *lines, int lineNo, float radius, v2 opt, int dirOpt, v2 *result)
{
    const float dotProduct = vdot(lines[lineNo].point, lines[lineNo].direction);
    const float discriminant = sqrf(dotProduct) + sqrf(radius) - vabssq(lines[lineNo].point);
    if (discriminant < 0.0f) return 0;
    const float sqrtDiscriminant = sqrtf(discriminant);
    float tLeft = -dotProduct - sqrtDiscriminant;
    float tRight = -dotProduct + sqrtDiscriminant;
    for (int i = 0; i < lineNo; ++i) {
        const float denominator = vdet(lines[lineNo].direction, lines[i].direction);
        const float numerator = vdet(lines[i].direction, vsub(lines[lineNo].point, lines[i].point));
        if (fabsf(denominator) <= RVO_EPSILON) {
            if (numerator < 0.0f) return 0;
            continue;
        }
        const float t = numerator / denominator;
        if (denominator >= 0.0f) tRight = fminf(tRight, t);
        else tLeft = fmaxf(tLeft, t);
        if (tLeft > tRight) return 0;
    }
    if (dirOpt) {
        if (vdot(opt, lines[lineNo].direction) > 0.0f)
            *result = vadd(lines[lineNo].point, vmul(tRight, lines[lineNo].direction));
        else
            *result = vadd(lines[lineNo].point, vmul(tLeft, lines[lineNo].direction));
    } else {
        const float t = vdot(lines[lineNo].direction, vsub(opt, lines[lineNo].point));
        if (t < tLeft) *result = vadd(lines[lineNo].point, vmul(tLeft, lines[lineNo].direction));
        else if (t > tRight) *result = vadd(lines[lineNo].point, vmul(tRight, lines[lineNo].direction));
        else *result = vadd(lines[lineNo].point, vmul(t, lines[lineNo].direction));
    }
    return 1;
}

static int lp2(const line_t *lines, int m, float radius, v2 opt, int dirOpt, v2 *result)
{
    if (dirOpt) *result = vmul(radius, opt);
    else if (vabssq(opt) > sqrf(radius)) *result = vmul(radius, vnormalize(opt));
    else *result = opt;
    for (int i = 0; i < m; ++i) {
        if (vdet(lines[i].direction, vsub(lines[i].point, *result)) > 0.0f) {
            const v2 temp = *result;
            if (!lp1(lines, i, radius, opt, dirOpt, result)) { *result = temp; return i; }
        }
    }
    return m;
}

static void lp3(const line_t *lines, int m, int numObstLines, int beginLine, float radius, v2 *result)
{
    float distance = 0.0f;
    line_t proj[AEMO_MAX_HUMANS + 1];
    for (int i = beginLine; i < m; ++i) {
        if (vdet(lines[i].direction, vsub(lines[i].point, *result)) > distance) {
            int np = 0;
            for (int j = 0; j < numObstLines; ++j) proj[np++] = lines[j];
            for (int j = numObstLines; j < i; ++j) {
                line_t line;
                float determinant = vdet(lines[i].direction, lines[j].direction);
                if (fabsf(determinant) <= RVO_EPSILON) {
                    if (vdot(lines[i].direction, lines[j].direction) > 0.0f) continue;
                    line.point = vmul(0.5f, vadd(lines[i].point, lines[j].point));
                } else {
                    line.point = vadd(lines[i].point,
                                      vmul(vdet(lines[j].direction, vsub(lines[i].point, lines[j].point)) / determinant,
                                           lines[i].direction));
                }
                line.direction = vnormalize(vsub(lines[j].direction, lines[i].direction));
                proj[np++] = line;
            }
            const v2 temp = *result;
            if (lp2(proj, np, radius, V(-lines[i].direction.y, lines[i].direction.x), 1, result) < np) *result = temp;
            distance = vdet(lines[i].direction, vsub(lines[i].point, *result));
        }
    }
}

void aemo_orca_agent(const float *self_pos, const float *self_vel, float self_radius, float max_speed,
                     const float *pref_vel, int m, const float *others, float neighbor_dist, int max_neighbors,
                     float time_horizon, float time_step, float *new_vel, int *nb_idx_out, int *num_nb_out,
                     float *lines_out)
{
    const v2 pos = V(self_pos[0], self_pos[1]);
    const v2 vel = V(self_vel[0], self_vel[1]);
    /* Agent::computeNeighbors + insertAgentNeighbor: insertion-sorted by squared distance, capped */
    int nb[AEMO_MAX_HUMANS + 1];
    float nbd[AEMO_MAX_HUMANS + 1];
    int cnt = 0;
    float rangeSq = sqrf(neighbor_dist);
    for (int j = 0; j < m; ++j) {
        const v2 op = V(others[5 * j], others[5 * j + 1]);
        const float distSq = vabssq(vsub(pos, op));
        if (distSq < rangeSq) {
            if (cnt < max_neighbors) { nb[cnt] = j; nbd[cnt] = distSq; ++cnt; }
            int i = cnt - 1;
            while (i != 0 && distSq < nbd[i - 1]) { nb[i] = nb[i - 1]; nbd[i] = nbd[i - 1]; --i; }
            nb[i] = j; nbd[i] = distSq;
            if (cnt == max_neighbors) rangeSq = nbd[cnt - 1];
        }
    }
    /* Agent::computeNewVelocity (no obstacles) */
    line_t lines[AEMO_MAX_HUMANS + 1];
    const float invTimeHorizon = 1.0f / time_horizon;
    for (int i = 0; i < cnt; ++i) {
        const float *o = others + 5 * nb[i];
        const v2 relativePosition = vsub(V(o[0], o[1]), pos);
        const v2 relativeVelocity = vsub(vel, V(o[2], o[3]));
        const float distSq = vabssq(relativePosition);
        const float combinedRadius = self_radius + o[4];
        const float combinedRadiusSq = sqrf(combinedRadius);
        line_t line;
        v2 u;
        if (distSq > combinedRadiusSq) {
            const v2 w = vsub(relativeVelocity, vmul(invTimeHorizon, relativePosition));
            const float wLengthSq = vabssq(w);
            const float dotProduct1 = vdot(w, relativePosition);
            if (dotProduct1 < 0.0f && sqrf(dotProduct1) > combinedRadiusSq * wLengthSq) {
                const float wLength = sqrtf(wLengthSq);
                const v2 unitW = vdiv(w, wLength);
                line.direction = V(unitW.y, -unitW.x);
                u = vmul(combinedRadius * invTimeHorizon - wLength, unitW);
            } else {
                const float leg = sqrtf(distSq - combinedRadiusSq);
                if (vdet(relativePosition, w) > 0.0f) {
                    line.direction = vdiv(V(relativePosition.x * leg - relativePosition.y * combinedRadius,
                                            relativePosition.x * combinedRadius + relativePosition.y * leg), distSq);
                } else {
                    v2 t = vdiv(V(relativePosition.x * leg + relativePosition.y * combinedRadius,
                                  -relativePosition.x * combinedRadius + relativePosition.y * leg), distSq);
                    line.direction = V(-t.x, -t.y);
                }
                const float dotProduct2 = vdot(relativeVelocity, line.direction);
                u = vsub(vmul(dotProduct2, line.direction), relativeVelocity);
            }
        } else {
            const float invTimeStep = 1.0f / time_step;
            const v2 w = vsub(relativeVelocity, vmul(invTimeStep, relativePosition));
            const float wLength = vabs(w);
            const v2 unitW = vdiv(w, wLength);
            line.direction = V(unitW.y, -unitW.x);
            u = vmul(combinedRadius * invTimeStep - wLength, unitW);
        }
        line.point = vadd(vel, vmul(0.5f, u));
        lines[i] = line;
    }
    v2 res;
    int lineFail = lp2(lines, cnt, max_speed, V(pref_vel[0], pref_vel[1]), 0, &res);
    if (lineFail < cnt) lp3(lines, cnt, 0, lineFail, max_speed, &res);
    new_vel[0] = res.x; new_vel[1] = res.y;
    if (num_nb_out) *num_nb_out = cnt;
    if (nb_idx_out) for (int i = 0; i < cnt; ++i) nb_idx_out[i] = nb[i];
    if (lines_out)
        for (int i = 0; i < cnt; ++i) {
            lines_out[4 * i] = lines[i].point.x; lines_out[4 * i + 1] = lines[i].point.y;
            lines_out[4 * i + 2] = lines[i].direction.x; lines_out[4 * i + 3] = lines[i].direction.y;
        }
}

/* orca.py:82-132 for every human of every env (crowd_sim.py:562-568) */
typedef struct { int n; const double *humans, *robot; const aemo_env_params *p; double *new_vel; } orca_ctx;
static void orca_body(int b, void *vc)
{
    orca_ctx *c = (orca_ctx *)vc;
    const int n = c->n;
    const aemo_env_params *p = c->p;
    const double *hb = c->humans + (size_t)b * n * 8;
    const double *rb = c->robot + (size_t)b * 9;
    for (int i = 0; i < n; ++i) {
        const double *me = hb + i * 8;
        float others[(AEMO_MAX_HUMANS + 1) * 5];
        int m = 0;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            const double *o = hb + j * 8;
            others[5 * m] = (float)o[0]; others[5 * m + 1] = (float)o[1];
            others[5 * m + 2] = (float)o[2]; others[5 * m + 3] = (float)o[3];
            others[5 * m + 4] = (float)(o[4] + 0.01 + p->orca_safety_space);
            ++m;
        }
        if (p->robot_visible) {                     /* crowd_sim.py:566-567 */
            others[5 * m] = (float)rb[0]; others[5 * m + 1] = (float)rb[1];
            others[5 * m + 2] = (float)rb[2]; others[5 * m + 3] = (float)rb[3];
            others[5 * m + 4] = (float)(rb[4] + 0.01 + p->orca_safety_space);
            ++m;
        }
        float pos[2] = { (float)me[0], (float)me[1] };
        float vel[2] = { (float)me[2], (float)me[3] };
        /* orca.py:112-115 preferred velocity in f64, then cast */
        double vx = me[5] - me[0], vy = me[6] - me[1];
        double speed = sqrt(vx * vx + vy * vy);
        if (speed > 1.0) { vx /= speed; vy /= speed; }
        float pref[2] = { (float)vx, (float)vy };
        float nv[2];
        aemo_orca_agent(pos, vel, (float)(me[4] + 0.01 + p->orca_safety_space), (float)me[7], pref, m, others,
                        10.0f, 10, 5.0f, (float)p->time_step, nv, NULL, NULL, NULL);
        c->new_vel[((size_t)b * n + i) * 2] = (double)nv[0];
        c->new_vel[((size_t)b * n + i) * 2 + 1] = (double)nv[1];
    }
}

void aemo_orca(int B, int n, const double *humans, const double *robot, const aemo_env_params *p, double *new_vel,
               int nthreads)
{
    orca_ctx c = { n, humans, robot, p, new_vel };
    parallel_for(B, nthreads, orca_body, &c);
}

/* ================================================================== */
/* adaptive occupancy reward  crowd_sim.py:20-219                      */
/* ================================================================== */
#define GRID_RES 0.02
#define GRID_W 600
#define GRID_OFF 6.0
#define PI_D 3.141592653589793

static double g_lut[10000];
static int g_lut_ready = 0;

static pthread_once_t g_lut_once = PTHREAD_ONCE_INIT;
static void lut_fill(void)
{
    /* crowd_sim.py:20-28: norm.pdf((i - 5000) / 1000) */
    for (int i = 0; i < 10000; ++i) {
        double x = (double)(i - 5000) / 1000.0;
        g_lut[i] = exp(-x * x / 2.0) / sqrt(2.0 * PI_D);
    }
    g_lut_ready = 1;
}
static void lut_init(void) { pthread_once(&g_lut_once, lut_fill); }

static inline double lut(double raw)
{
    /* crowd_sim.py:32-38 : int() truncates toward zero */
    double t = raw * 1000.0 + 5000.0;
    /* guard the int conversion against overflow; out of table -> 0 */
    if (!(t > -1.0e9 && t < 1.0e9)) return 0.0;
    int index = (int)t;
    if (index > 9999 || index < 0) return 0.0;
    return g_lut[index];
}

typedef struct {
    int cx, cy, xlo, xhi, ylo, yhi, rw, rh, active;
    double m, heading, sum;
} blob_t;

static inline double blob_weight(const blob_t *b, int x, int y, double pos_var, double dir_var)
{
    /* crowd_sim.py:84-106 */
    double angle;
    if ((y - b->cy) == 0 && (x - b->cx) == 0) angle = b->heading;
    else angle = atan2((double)(y - b->cy), (double)(x - b->cx));
    double delta = angle - b->heading;
    if (delta <= -PI_D) delta += 2.0 * PI_D;
    else if (delta >= PI_D) delta = 2.0 * PI_D - delta;
    double x0 = pos_var * (double)(y - b->cy) / ((double)b->rh + 0.000001);
    double x1 = pos_var * (double)(x - b->cx) / ((double)b->rw + 0.000001);
    double x2 = dir_var * delta / PI_D;
    return lut(x0) * lut(x1) * lut(x2);
}

static inline int in_blob(const blob_t *b, int x, int y)
{
    if (x < b->xlo || x >= b->xhi || y < b->ylo || y >= b->yhi) return 0;
    return !(hypot((double)(x - b->cx), (double)(y - b->cy)) * GRID_RES > b->m);
}

static void blob_setup(blob_t *b, const double *h, double human_timestep, double pos_var, double dir_var, int need_sum)
{
    /* crowd_sim.py:47-82 */
    double p_x = h[0] + GRID_OFF, p_y = h[1] + GRID_OFF;
    b->cx = (int)rint(p_x / GRID_RES);
    b->cy = (int)rint(p_y / GRID_RES);
    double dx = h[2] * human_timestep, dy = h[3] * human_timestep;
    b->m = hypot(dx, dy);
    b->xlo = (int)rint((p_x - b->m) / GRID_RES);
    b->xhi = (int)rint((p_x + b->m) / GRID_RES);
    b->ylo = (int)rint((p_y - b->m) / GRID_RES);
    b->yhi = (int)rint((p_y + b->m) / GRID_RES);
    b->rw = b->xhi - b->xlo;
    b->rh = b->yhi - b->ylo;
    b->heading = atan2(dy, dx);
    b->sum = 0.0;
    b->active = 0;
    if (!need_sum) return;
    double s = 0.0;
    for (int y = b->ylo; y < b->yhi; ++y)
        for (int x = b->xlo; x < b->xhi; ++x) {
            if (hypot((double)(x - b->cx), (double)(y - b->cy)) * GRID_RES > b->m) continue;
            s += blob_weight(b, x, y, pos_var, dir_var);
        }
    b->sum = s;
    b->active = (s != 0.0);                            /* crowd_sim.py:113-116 */
}

/* disc sum for a robot look-ahead position (crowd_sim.py:159-185, 122-132).
 * DEVIATION (documented in DESIGN.md): cells of a human blob whose index falls outside
 * [0, 600) are dropped (the reference wraps negative indices and is out-of-bounds UB above). */
static double disc_sum(const blob_t *blobs, int n, double rx, double ry, double search_radius, double pos_var,
                       double dir_var)
{
    double agent_x = rx + GRID_OFF, agent_y = ry + GRID_OFF;
    int cx = (int)rint(agent_x / GRID_RES), cy = (int)rint(agent_y / GRID_RES);
    double xlr = (agent_x - search_radius) < 0 ? 0.0 : (agent_x - search_radius);
    double ylr = (agent_y - search_radius) < 0 ? 0.0 : (agent_y - search_radius);
    int xlo = (int)rint(xlr / GRID_RES), xhi = (int)rint((agent_x + search_radius) / GRID_RES);
    int ylo = (int)rint(ylr / GRID_RES), yhi = (int)rint((agent_y + search_radius) / GRID_RES);
    double sum = 0.0;
    for (int y = ylo; y < yhi; ++y)
        for (int x = xlo; x < xhi; ++x) {
            if (hypot((double)(x - cx), (double)(y - cy)) * GRID_RES > search_radius) continue;
            if (x < GRID_W && y < GRID_W) {
                double cell = 0.0;                     /* global_gmap starts at 0, += gmap per human */
                for (int i = 0; i < n; ++i) {
                    const blob_t *b = &blobs[i];
                    if (!b->active) continue;
                    if (in_blob(b, x, y)) cell += blob_weight(b, x, y, pos_var, dir_var) / b->sum;
                }
                sum += cell;
            }
        }
    return sum;
}

static int blob_may_touch(const double *h, double human_timestep, double rx, double ry, double reach)
{
    /* conservative bounding test: blob radius m around the human vs. disc(s) within `reach` of (rx, ry);
     * +0.1 m slack covers every index rounding.  Skipping a blob that cannot touch is exact. */
    double m = hypot(h[2] * human_timestep, h[3] * human_timestep);
    double d = hypot(h[0] - rx, h[1] - ry);
    return d <= m + reach + 0.1;
}

double aemo_occupancy(const double *hum, int n, double rx, double ry, double robot_radius, double human_timestep,
                      double pos_var, double dir_var)
{
    lut_init();
    blob_t blobs[AEMO_MAX_HUMANS];
    double sr = robot_radius + (n > 0 ? hum[4] : 0.0);  /* crowd_sim.py:169-173 */
    for (int i = 0; i < n; ++i)
        blob_setup(&blobs[i], hum + 5 * i, human_timestep, pos_var, dir_var,
                   blob_may_touch(hum + 5 * i, human_timestep, rx, ry, sr));
    return disc_sum(blobs, n, rx, ry, sr, pos_var, dir_var);
}

/* crowd_sim/envs/utils/utils.py:4-26 */
static double point_to_segment_dist(double x1, double y1, double x2, double y2, double x3, double y3)
{
    double px = x2 - x1, py = y2 - y1;
    if (px == 0 && py == 0) return sqrt((x3 - x1) * (x3 - x1) + (y3 - y1) * (y3 - y1));
    double u = ((x3 - x1) * px + (y3 - y1) * py) / (px * px + py * py);
    if (u > 1) u = 1; else if (u < 0) u = 0;
    double x = x1 + u * px, y = y1 + u * py;
    return sqrt((x - x3) * (x - x3) + (y - y3) * (y - y3));
}

/* crowd_sim.py:557-651, 676-680 (update=False) for all actions */
typedef struct {
    int n, A; const double *robot, *humans, *new_vel, *global_time, *actions; const aemo_env_params *p;
    double *rewards; int32_t *info; double *dmin_out, *humans_next;
} el_ctx;
static void el_body(int b, void *vc)
{
    el_ctx *c = (el_ctx *)vc;
    const int n = c->n, A = c->A;
    const aemo_env_params *p = c->p;
    const double *actions = c->actions;
    const double *rb = c->robot + (size_t)b * 9;
    const double *hb = c->humans + (size_t)b * n * 8;
    double hum5[AEMO_MAX_HUMANS * 5];
    blob_t blobs[AEMO_MAX_HUMANS];
    for (int i = 0; i < n; ++i) {
        hum5[5 * i] = hb[8 * i]; hum5[5 * i + 1] = hb[8 * i + 1]; hum5[5 * i + 2] = hb[8 * i + 2];
        hum5[5 * i + 3] = hb[8 * i + 3]; hum5[5 * i + 4] = hb[8 * i + 4];
    }
    double sr = rb[4] + (n > 0 ? hum5[4] : 0.0);
    double amax = 0.0;
    for (int a = 0; a < A; ++a) {
        double s = hypot(actions[2 * a], actions[2 * a + 1]) * p->agent_timestep;
        if (s > amax) amax = s;
    }
    for (int i = 0; i < n; ++i)
        blob_setup(&blobs[i], hum5 + 5 * i, p->human_timestep, p->position_variance, p->direction_variance,
                   blob_may_touch(hum5 + 5 * i, p->human_timestep, rb[0], rb[1], sr + amax));
    for (int a = 0; a < A; ++a) {
        double ax = actions[2 * a], ay = actions[2 * a + 1];
        /* collision detection crowd_sim.py:570-591 */
        double dmin = INFINITY;
        int collision = 0;
        for (int i = 0; i < n; ++i) {
            double px = hum5[5 * i] - rb[0], py = hum5[5 * i + 1] - rb[1];
            double vx = hum5[5 * i + 2] - ax, vy = hum5[5 * i + 3] - ay;
            double ex = px + vx * p->time_step, ey = py + vy * p->time_step;
            double cd = point_to_segment_dist(px, py, ex, ey, 0, 0) - hum5[5 * i + 4] - rb[4];
            if (cd < 0) { collision = 1; break; }
            else if (cd < dmin) dmin = cd;
        }
        /* occupancy crowd_sim.py:594-604 */
        double prob = disc_sum(blobs, n, rb[0] + ax * p->agent_timestep, rb[1] + ay * p->agent_timestep, sr,
                               p->position_variance, p->direction_variance);
        /* goal crowd_sim.py:622-623 */
        double ex = rb[0] + ax * p->time_step - rb[5], ey = rb[1] + ay * p->time_step - rb[6];
        int reaching_goal = sqrt(ex * ex + ey * ey) < rb[4];
        double reward; int code;
        if (c->global_time[b] >= p->time_limit - 1) { reward = 0; code = AEMO_INFO_TIMEOUT; }
        else if (collision) { reward = p->collision_penalty; code = AEMO_INFO_COLLISION; }
        else if (reaching_goal) { reward = p->success_reward; code = AEMO_INFO_REACHGOAL; }
        else {
            reward = -0.25 * prob * p->reward_increment;
            code = (dmin < p->discomfort_dist) ? AEMO_INFO_DANGER : AEMO_INFO_NOTHING;
        }
        c->rewards[(size_t)b * A + a] = reward;
        if (c->info) c->info[(size_t)b * A + a] = code;
        if (c->dmin_out) c->dmin_out[(size_t)b * A + a] = dmin;
    }
    if (c->humans_next)
        for (int i = 0; i < n; ++i) {              /* agent.py:63-74 */
            double vx = c->new_vel[((size_t)b * n + i) * 2], vy = c->new_vel[((size_t)b * n + i) * 2 + 1];
            double *o = c->humans_next + ((size_t)b * n + i) * 5;
            o[0] = hum5[5 * i] + vx * p->time_step;
            o[1] = hum5[5 * i + 1] + vy * p->time_step;
            o[2] = vx; o[3] = vy; o[4] = hum5[5 * i + 4];
        }
}

void aemo_env_lookahead(int B, int n, int A, const double *robot, const double *humans, const double *new_vel,
                        const double *global_time, const double *actions, const aemo_env_params *p, double *rewards,
                        int32_t *info, double *dmin_out, double *humans_next, int nthreads)
{
    lut_init();
    el_ctx c = { n, A, robot, humans, new_vel, global_time, actions, p, rewards, info, dmin_out, humans_next };
    parallel_for(B, nthreads, el_body, &c);
}

/* crowd_sim.py:653-669 update=True */
void aemo_env_apply(int B, int n, int A, double *robot, double *humans, const double *new_vel, double *global_time,
                    const double *actions, const int32_t *chosen, const aemo_env_params *p)
{
    (void)A;
    for (int b = 0; b < B; ++b) {
        double *rb = robot + (size_t)b * 9;
        double ax = actions[2 * chosen[b]], ay = actions[2 * chosen[b] + 1];
        rb[0] = rb[0] + ax * p->time_step; rb[1] = rb[1] + ay * p->time_step;   /* agent.py:122-131 */
        rb[2] = ax; rb[3] = ay;
        for (int i = 0; i < n; ++i) {
            double *h = humans + ((size_t)b * n + i) * 8;
            double vx = new_vel[((size_t)b * n + i) * 2], vy = new_vel[((size_t)b * n + i) * 2 + 1];
            h[0] = h[0] + vx * p->time_step; h[1] = h[1] + vy * p->time_step;
            h[2] = vx; h[3] = vy;
        }
        global_time[b] += p->time_step;
    }
}
