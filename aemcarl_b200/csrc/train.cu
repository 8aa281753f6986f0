// train.cu -- the value-network regression step as hand-written kernels:
//   forward of ValueNetworkTransformerAndGRU (qnetwork_factory.py:641-687, components.py:38-137) on a batch of rotated
//   states with the reference's BATCH-GLOBAL adaptive-computation-time halting (components.py:132-134), MSE against the
//   targets (trainer.py:53,76), the complete backward pass, and torch.optim.Adam's update (trainer.py:30).
//
// Shape of the problem: 100 samples x (n + 1) tokens x 113,752 parameters -- every weight is used for only T = n + 1 rows
// per sample, so the step is bound by weight traffic out of L2 and by latency, not by FLOPs.  Design:
//   * train_probe_kernel   one CTA per sample runs the three GRU steps and ORs "some row still below 1 - epsilon after
//                          step k" into two global flags: the batch-global step count K every sample must use.
//   * train_fwdbwd_kernel  one CTA per sample (grid-stride if the batch exceeds the grid): forward with every
//                          intermediate kept in shared memory (80 KB at T = 6; global scratch when it does not fit),
//                          then the hand-derived backward (tests/manual_backward.py is the same derivation in numpy,
//                          checked against autograd).  Weights are staged layer by layer into shared memory with an odd
//                          row pitch, so both products (y = W x and dx = W^T dy) read them without bank conflicts.
//                          Weight gradients go to the CTA's OWN slice part[cta][113,752] -- no atomics.
//   * train_reduce_kernel  grad[e] = sum over CTAs in a fixed order (deterministic), loss = mean squared error.
//   * adam_kernel          one thread per parameter.
// The gradient all-reduce of a data-parallel run sits between train_reduce_kernel and adam_kernel (trainer.py).
#include "aem_common.cuh"

namespace aem {

constexpr int TR_THREADS = 512;
constexpr int TR_WS_FLOATS = 7680;      // largest staged layer: 150 x 51
constexpr int TR_BUF_PITCH = 152;       // temporaries hold up to 150 columns per token row
constexpr int TR_NBUF = 8;
constexpr float ACT_EPS = 0.05f;        // ATCBasic epsilon (qnetwork_factory.py:619 -> components.py:93)

// per-sample activation stash (floats), T tokens
struct TrLayout {
    int x, h, gru, enc, fin, total;
    int gru_step, enc_layer;
};
__host__ __device__ inline TrLayout tr_layout(int T)
{
    TrLayout L;
    int o = 0;
    L.x = o; o += T * 13;
    L.h = o; o += 4 * T * 50;                      // h_0 .. h_3
    L.gru_step = T * 451;                          // az1 100 | az2 50 | z 50 | ar1 100 | ar2 50 | r 50 | q 50 | p 1
    L.gru = o; o += 3 * L.gru_step;
    L.enc_layer = T * 552 + 2 * T * T;             // X 50 | qkv 150 | ctx 50 | xh1 50 | y1 50 | f1 150 | xh2 50 | rs1 | rs2 ; P
    L.enc = o; o += 3 * L.enc_layer;
    L.fin = o; o += T * 50 + 56 + 100 + 50;        // X3 | ain | a1 | a2
    L.total = (o + 3) & ~3;
    return L;
}
// offsets inside one GRU step block (per-row pitch = the field width)
struct GruStash { float *az1, *az2, *z, *ar1, *ar2, *r, *q, *p; };
__device__ inline GruStash gru_stash(float *base, int T)
{
    GruStash g;
    g.az1 = base; g.az2 = g.az1 + T * 100; g.z = g.az2 + T * 50; g.ar1 = g.z + T * 50; g.ar2 = g.ar1 + T * 100;
    g.r = g.ar2 + T * 50; g.q = g.r + T * 50; g.p = g.q + T * 50;
    return g;
}
struct EncStash { float *X, *qkv, *ctx, *xh1, *y1, *f1, *xh2, *rs1, *rs2, *P; };
__device__ inline EncStash enc_stash(float *base, int T)
{
    EncStash e;
    e.X = base; e.qkv = e.X + T * 50; e.ctx = e.qkv + T * 150; e.xh1 = e.ctx + T * 50; e.y1 = e.xh1 + T * 50;
    e.f1 = e.y1 + T * 50; e.xh2 = e.f1 + T * 150; e.rs1 = e.xh2 + T * 50; e.rs2 = e.rs1 + T; e.P = e.rs2 + T;
    return e;
}

__device__ __forceinline__ float sigmoidf_(float a) { return 1.0f / (1.0f + expf(-a)); }

// ---- weight staging: W[O][I] (row-major, global) -> Ws[O][I | 1] ----
template <int O, int I>
__device__ __forceinline__ void stage_w(const float *__restrict__ Wg, float *Ws)
{
    constexpr int P = I | 1;
    static_assert(O * P <= TR_WS_FLOATS, "staging buffer");
    for (int e = threadIdx.x; e < O * I; e += blockDim.x) {
        const int o = e / I, i = e - o * I;
        Ws[o * P + i] = __ldg(Wg + e);
    }
    __syncthreads();
}

// Y[r][o] = act(b[o] + sum_i X[r][i] W[o][i]); ACT: 0 none, 1 ReLU
template <int O, int I, int ACT>
__device__ void lin_fwd(const float *__restrict__ Wg, const float *__restrict__ bg, const float *X, int ldx, float *Y,
                        int ldy, int T, float *Ws)
{
    constexpr int P = I | 1;
    stage_w<O, I>(Wg, Ws);
    for (int it = threadIdx.x; it < T * O; it += blockDim.x) {
        const int r = it / O, o = it - r * O;
        const float *w = Ws + o * P, *x = X + r * ldx;
        float a0 = 0.0f, a1 = 0.0f;
        int i = 0;
#pragma unroll 4
        for (; i + 1 < I; i += 2) {
            a0 = fmaf(x[i], w[i], a0);
            a1 = fmaf(x[i + 1], w[i + 1], a1);
        }
        if (i < I) a0 = fmaf(x[i], w[i], a0);
        float a = a0 + a1 + __ldg(bg + o);
        if (ACT == 1) a = fmaxf(a, 0.0f);
        Y[r * ldy + o] = a;
    }
    __syncthreads();
}

// dX[r][i] (i < ICNT) = (ADD ? dX[r][i] : 0) + sum_o dY[r][o] W[o][i]
template <int O, int I, int ICNT, bool ADD>
__device__ void lin_bwd_x(const float *__restrict__ Wg, const float *dY, int lddy, float *dX, int lddx, int T, float *Ws)
{
    constexpr int P = I | 1;
    stage_w<O, I>(Wg, Ws);
    for (int it = threadIdx.x; it < T * ICNT; it += blockDim.x) {
        const int r = it / ICNT, i = it - r * ICNT;
        const float *dy = dY + r * lddy, *w = Ws + i;
        float a0 = 0.0f, a1 = 0.0f;
        int o = 0;
#pragma unroll 4
        for (; o + 1 < O; o += 2) {
            a0 = fmaf(dy[o], w[o * P], a0);
            a1 = fmaf(dy[o + 1], w[(o + 1) * P], a1);
        }
        if (o < O) a0 = fmaf(dy[o], w[o * P], a0);
        const float a = a0 + a1;
        dX[r * lddx + i] = ADD ? dX[r * lddx + i] + a : a;
    }
    __syncthreads();
}

// gW[o][i] (+)= sum_r dY[r][o] X[r][i];  gb[o] (+)= sum_r dY[r][o]   (the CTA's own gradient slice: plain stores)
template <int O, int I>
__device__ void lin_bwd_w(float *gW, float *gb, const float *dY, int lddy, const float *X, int ldx, int T, bool acc)
{
    for (int e = threadIdx.x; e < O * I; e += blockDim.x) {
        const int o = e / I, i = e - o * I;
        float s = 0.0f;
        for (int r = 0; r < T; ++r) s = fmaf(dY[r * lddy + o], X[r * ldx + i], s);
        gW[e] = acc ? gW[e] + s : s;
    }
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
        float s = 0.0f;
        for (int r = 0; r < T; ++r) s += dY[r * lddy + o];
        gb[o] = acc ? gb[o] + s : s;
    }
    __syncthreads();
}

// ---- LayerNorm over 50 features, one warp per row ----
__device__ void ln_fwd(const float *S, const float *__restrict__ g, const float *__restrict__ b, float *Y, float *xhat,
                       float *rstd, int T)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < T; r += nw) {
        const float v0 = S[r * 50 + lane], v1 = lane + 32 < 50 ? S[r * 50 + lane + 32] : 0.0f;
        float s = v0 + v1;
        for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        const float mu = s * (1.0f / 50.0f);
        const float d0 = v0 - mu, d1 = lane + 32 < 50 ? v1 - mu : 0.0f;
        float q = d0 * d0 + d1 * d1;
        for (int m = 16; m; m >>= 1) q += __shfl_xor_sync(0xffffffffu, q, m);
        const float rs = 1.0f / sqrtf(q * (1.0f / 50.0f) + 1e-5f);
        if (lane == 0) rstd[r] = rs;
        xhat[r * 50 + lane] = d0 * rs;
        Y[r * 50 + lane] = d0 * rs * __ldg(g + lane) + __ldg(b + lane);
        if (lane + 32 < 50) {
            xhat[r * 50 + lane + 32] = d1 * rs;
            Y[r * 50 + lane + 32] = d1 * rs * __ldg(g + lane + 32) + __ldg(b + lane + 32);
        }
    }
    __syncthreads();
}

// dS = rstd (dxhat - mean(dxhat) - xhat mean(dxhat xhat)), dxhat = dY g;  gg += sum_r dY xhat;  gbeta += sum_r dY
__device__ void ln_bwd(const float *dY, const float *xhat, const float *rstd, const float *__restrict__ g, float *dS,
                       float *gg, float *gbeta, int T, bool acc)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = threadIdx.x; j < 50; j += blockDim.x) {
        float sg = 0.0f, sb = 0.0f;
        for (int r = 0; r < T; ++r) {
            sg = fmaf(dY[r * 50 + j], xhat[r * 50 + j], sg);
            sb += dY[r * 50 + j];
        }
        gg[j] = acc ? gg[j] + sg : sg;
        gbeta[j] = acc ? gbeta[j] + sb : sb;
    }
    __syncthreads();               // dS may alias dY
    for (int r = warp; r < T; r += nw) {
        const bool two = lane + 32 < 50;
        const float x0 = xhat[r * 50 + lane], x1 = two ? xhat[r * 50 + lane + 32] : 0.0f;
        const float d0 = dY[r * 50 + lane] * __ldg(g + lane), d1 = two ? dY[r * 50 + lane + 32] * __ldg(g + lane + 32) : 0.0f;
        float m1 = d0 + d1, m2 = d0 * x0 + d1 * x1;
        for (int m = 16; m; m >>= 1) {
            m1 += __shfl_xor_sync(0xffffffffu, m1, m);
            m2 += __shfl_xor_sync(0xffffffffu, m2, m);
        }
        m1 *= (1.0f / 50.0f);
        m2 *= (1.0f / 50.0f);
        const float rs = rstd[r];
        dS[r * 50 + lane] = rs * (d0 - m1 - x0 * m2);
        if (two) dS[r * 50 + lane + 32] = rs * (d1 - m1 - x1 * m2);
    }
    __syncthreads();
}

// ---- the GRU-ACT forward (components.py:46-53,107-137): K steps, everything kept in the stash; A = sum_k p_k h_k ----
// st = stash base; buf = 2 temporaries [T][TR_BUF_PITCH]
// pk (optional): pk[(k - 1) * T + r] = halting probability of row r at step k
__device__ void gru_forward(const float *__restrict__ w, float *st, const TrLayout &L, int T, int K, float *A, float *hxb,
                            float *Ws, float *pk)
{
    const float *x = st + L.x;
    for (int e = threadIdx.x; e < T * 50; e += blockDim.x) {
        st[L.h + e] = 0.0f;
        A[e] = 0.0f;
    }
    __syncthreads();
    for (int k = 1; k <= K; ++k) {
        const float *hp = st + L.h + (k - 1) * T * 50;
        float *hn = st + L.h + k * T * 50;
        GruStash g = gru_stash(st + L.gru + (k - 1) * L.gru_step, T);
        for (int e = threadIdx.x; e < T * 63; e += blockDim.x) {
            const int r = e / 63, j = e - r * 63;
            hxb[r * 64 + j] = j < 50 ? hp[r * 50 + j] : x[r * 13 + (j - 50)];
        }
        __syncthreads();
        lin_fwd<100, 63, 1>(w + OFF_Z1W, w + OFF_Z1B, hxb, 64, g.az1, 100, T, Ws);
        lin_fwd<50, 100, 1>(w + OFF_Z2W, w + OFF_Z2B, g.az1, 100, g.az2, 50, T, Ws);
        lin_fwd<100, 63, 1>(w + OFF_R1W, w + OFF_R1B, hxb, 64, g.ar1, 100, T, Ws);
        lin_fwd<50, 100, 1>(w + OFF_R2W, w + OFF_R2B, g.ar1, 100, g.ar2, 50, T, Ws);
        for (int e = threadIdx.x; e < T * 50; e += blockDim.x) {
            const int r = e / 50, j = e - r * 50;
            const float z = sigmoidf_(g.az2[e]), rr = sigmoidf_(g.ar2[e]);
            g.z[e] = z;
            g.r[e] = rr;
            hxb[r * 64 + j] = rr * hp[e];               // [r * h ; x]
        }
        __syncthreads();
        lin_fwd<50, 63, 0>(w + OFF_QW, w + OFF_QB, hxb, 64, g.q, 50, T, Ws);
        for (int e = threadIdx.x; e < T * 50; e += blockDim.x) {
            const float q = tanhf(g.q[e]);
            g.q[e] = q;
            hn[e] = (1.0f - g.z[e]) * hp[e] + g.z[e] * q;
        }
        __syncthreads();
        {   // halting probability, one warp per row
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
            for (int r = warp; r < T; r += nw) {
                float s = hn[r * 50 + lane] * __ldg(w + OFF_PW + lane);
                if (lane + 32 < 50) s = fmaf(hn[r * 50 + lane + 32], __ldg(w + OFF_PW + lane + 32), s);
                for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
                if (lane == 0) {
                    g.p[r] = sigmoidf_(s + __ldg(w + OFF_PB));
                    if (pk) pk[(k - 1) * T + r] = g.p[r];
                }
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < T * 50; e += blockDim.x) A[e] = fmaf(g.p[e / 50], hn[e], A[e]);
        __syncthreads();
    }
}

// robot pseudo-token + the n human rows (qnetwork_factory.py:651-663)
__device__ void load_tokens(const float *__restrict__ state, int n, float *x)
{
    for (int e = threadIdx.x; e < (n + 1) * 13; e += blockDim.x) {
        const int r = e / 13, j = e - r * 13;
        float v;
        if (r > 0) v = __ldg(state + (r - 1) * 13 + j);
        else if (j < 6) v = __ldg(state + j);
        else if (j == 8) v = __ldg(state + 4);
        else if (j == 9) v = __ldg(state + 5);
        else if (j == 10 || j == 12) v = __ldg(state + 3);
        else v = 0.0f;
        x[e] = v;
    }
    __syncthreads();
}

struct TrainArgs {
    const float *params, *states, *targets;
    float *part;        // [grid][NUM_WEIGHTS]
    float *sqerr;       // [B]
    float *values;      // [B] or null
    int *flags;         // [2]
    float *stash_g;     // [grid][stash] when the stash does not fit in shared memory, else null
    int B, n;
};

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TR_THREADS, 1) train_probe_kernel(TrainArgs a)
{
    extern __shared__ __align__(16) float sm[];
    const int T = a.n + 1;
    TrLayout L = tr_layout(T);
    L.gru_step = 0;                                 // the probe keeps nothing: every step reuses one stash block
    float *Ws = sm, *A = Ws + TR_WS_FLOATS, *hxb = A + T * 50, *pk = hxb + T * 64, *st = pk + ((3 * T + 3) & ~3);
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        load_tokens(a.states + (size_t)b * a.n * 13, a.n, st + L.x);
        gru_forward(a.params, st, L, T, 3, A, hxb, Ws, pk);
        if (threadIdx.x < 32) {
            int un1 = 0, un2 = 0;
            for (int r = threadIdx.x; r < T; r += 32) {
                const float p1 = pk[r];
                const float p2 = p1 + pk[T + r];
                un1 |= p1 < 1.0f - ACT_EPS;
                un2 |= p2 < 1.0f - ACT_EPS;
            }
            un1 = __any_sync(0xffffffffu, un1);
            un2 = __any_sync(0xffffffffu, un2);
            if (threadIdx.x == 0) {
                if (un1) atomicOr(a.flags + 0, 1);
                if (un2) atomicOr(a.flags + 1, 1);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TR_THREADS, 1) train_fwdbwd_kernel(TrainArgs a)
{
    extern __shared__ __align__(16) float sm[];
    const int T = a.n + 1;
    const TrLayout L = tr_layout(T);
    float *Ws = sm;
    float *buf[TR_NBUF];
    for (int i = 0; i < TR_NBUF; ++i) buf[i] = Ws + TR_WS_FLOATS + i * T * TR_BUF_PITCH;
    float *small_ = buf[TR_NBUF - 1] + T * TR_BUF_PITCH;            // 64 floats
    float *st = a.stash_g ? a.stash_g + (size_t)blockIdx.x * L.total : small_ + 64;
    const float *w = a.params;
    float *G = a.part + (size_t)blockIdx.x * NUM_WEIGHTS;
    const int K = a.flags[0] == 0 ? 1 : (a.flags[1] == 0 ? 2 : 3);
    const float invK = 1.0f / (float)K;
    const int tid = threadIdx.x, nt = blockDim.x;

    if (blockIdx.x >= a.B) {                       // more CTAs than samples: a zero slice
        for (int e = tid; e < NUM_WEIGHTS; e += nt) G[e] = 0.0f;
        return;
    }
    int iter = 0;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x, ++iter) {
        const bool acc = iter > 0;
        const float *state = a.states + (size_t)b * a.n * 13;
        // =========================================== forward ===========================================
        load_tokens(state, a.n, st + L.x);
        float *A = buf[0], *hxb = buf[1];
        gru_forward(w, st, L, T, K, A, hxb, Ws, nullptr);
        {
            float *X0 = enc_stash(st + L.enc, T).X;
            for (int e = tid; e < T * 50; e += nt) X0[e] = A[e] * invK;
            __syncthreads();
        }
        for (int l = 0; l < 3; ++l) {
            const float *E = w + OFF_ENC + l * ENC_SIZE;
            EncStash c = enc_stash(st + L.enc + l * L.enc_layer, T);
            float *Xn = l < 2 ? enc_stash(st + L.enc + (l + 1) * L.enc_layer, T).X : st + L.fin;
            lin_fwd<150, 50, 0>(E + ENC_INW, E + ENC_INB, c.X, 50, c.qkv, 150, T, Ws);
            for (int e = tid; e < 2 * T * T; e += nt) {             // scores (q . k) / sqrt(25)
                const int hd = e / (T * T), rc = e - hd * T * T, r = rc / T, cc = rc - r * T;
                const float *q = c.qkv + r * 150 + 25 * hd, *k = c.qkv + cc * 150 + 50 + 25 * hd;
                float s = 0.0f;
#pragma unroll
                for (int d = 0; d < 25; ++d) s = fmaf(q[d], k[d], s);
                c.P[e] = s * 0.2f;
            }
            __syncthreads();
            for (int e = tid; e < 2 * T; e += nt) {                 // softmax over the T keys
                float *p = c.P + e * T;
                float m = p[0];
                for (int cc = 1; cc < T; ++cc) m = fmaxf(m, p[cc]);
                float s = 0.0f;
                for (int cc = 0; cc < T; ++cc) {
                    p[cc] = expf(p[cc] - m);
                    s += p[cc];
                }
                const float inv = 1.0f / s;
                for (int cc = 0; cc < T; ++cc) p[cc] *= inv;
            }
            __syncthreads();
            for (int e = tid; e < T * 50; e += nt) {                // ctx = P v
                const int r = e / 50, j = e - r * 50, hd = j / 25;
                const float *p = c.P + (hd * T + r) * T;
                float s = 0.0f;
                for (int cc = 0; cc < T; ++cc) s = fmaf(p[cc], c.qkv[cc * 150 + 100 + j], s);
                c.ctx[e] = s;
            }
            __syncthreads();
            float *s1 = buf[0], *f2 = buf[1];
            lin_fwd<50, 50, 0>(E + ENC_OW, E + ENC_OB, c.ctx, 50, s1, 50, T, Ws);
            for (int e = tid; e < T * 50; e += nt) s1[e] += c.X[e];
            __syncthreads();
            ln_fwd(s1, E + ENC_N1G, E + ENC_N1B, c.y1, c.xh1, c.rs1, T);
            lin_fwd<150, 50, 1>(E + ENC_L1W, E + ENC_L1B, c.y1, 50, c.f1, 150, T, Ws);
            lin_fwd<50, 150, 0>(E + ENC_L2W, E + ENC_L2B, c.f1, 150, f2, 50, T, Ws);
            for (int e = tid; e < T * 50; e += nt) f2[e] += c.y1[e];
            __syncthreads();
            ln_fwd(f2, E + ENC_N2G, E + ENC_N2B, Xn, c.xh2, c.rs2, T);
        }
        float *X3 = st + L.fin, *ain = X3 + T * 50, *a1 = ain + 56, *a2 = a1 + 100;
        for (int e = tid; e < 56; e += nt) ain[e] = e < 6 ? __ldg(state + e) : X3[e - 6];
        __syncthreads();
        lin_fwd<100, 56, 1>(w + OFF_A0W, w + OFF_A0B, ain, 56, a1, 100, 1, Ws);
        lin_fwd<50, 100, 1>(w + OFF_A2W, w + OFF_A2B, a1, 100, a2, 50, 1, Ws);
        if (tid < 32) {
            float s = a2[tid] * __ldg(w + OFF_A4W + tid);
            if (tid + 32 < 50) s = fmaf(a2[tid + 32], __ldg(w + OFF_A4W + tid + 32), s);
            for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
            if (tid == 0) {
                const float v = s + __ldg(w + OFF_A4B), d = v - __ldg(a.targets + b);
                small_[0] = 2.0f * d / (float)a.B;                  // d mean((v - t)^2) / dv
                a.sqerr[b] = d * d;
                if (a.values) a.values[b] = v;
            }
        }
        __syncthreads();
        // =========================================== backward ==========================================
        const float dv = small_[0];
        float *dX = buf[0], *dh = buf[1], *t2 = buf[2], *t3 = buf[3], *t4 = buf[4], *t5 = buf[5], *t6 = buf[6], *t7 = buf[7];
        {   // action MLP, one row
            float *da2 = t2, *da1 = t3, *dain = t4;
            for (int e = tid; e < 50; e += nt) {
                const float g = dv * a2[e];
                G[OFF_A4W + e] = acc ? G[OFF_A4W + e] + g : g;
                da2[e] = a2[e] > 0.0f ? dv * __ldg(w + OFF_A4W + e) : 0.0f;
            }
            if (tid == 0) G[OFF_A4B] = acc ? G[OFF_A4B] + dv : dv;
            __syncthreads();
            lin_bwd_w<50, 100>(G + OFF_A2W, G + OFF_A2B, da2, 50, a1, 100, 1, acc);
            lin_bwd_x<50, 100, 100, false>(w + OFF_A2W, da2, 50, da1, 100, 1, Ws);
            for (int e = tid; e < 100; e += nt) da1[e] = a1[e] > 0.0f ? da1[e] : 0.0f;
            __syncthreads();
            lin_bwd_w<100, 56>(G + OFF_A0W, G + OFF_A0B, da1, 100, ain, 56, 1, acc);
            lin_bwd_x<100, 56, 56, false>(w + OFF_A0W, da1, 100, dain, 56, 1, Ws);
            for (int e = tid; e < T * 50; e += nt) dX[e] = e < 50 ? dain[6 + e] : 0.0f;    // only token 0 feeds the head
            __syncthreads();
        }
        for (int l = 2; l >= 0; --l) {
            const float *E = w + OFF_ENC + l * ENC_SIZE;
            float *GE = G + OFF_ENC + l * ENC_SIZE;
            EncStash c = enc_stash(st + L.enc + l * L.enc_layer, T);
            float *ds2 = t2, *df1 = t3, *dy1 = t4, *ds1 = t2, *dctx = t4, *dqkv = t3, *dS = t5;
            ln_bwd(dX, c.xh2, c.rs2, E + ENC_N2G, ds2, GE + ENC_N2G, GE + ENC_N2B, T, acc);
            lin_bwd_w<50, 150>(GE + ENC_L2W, GE + ENC_L2B, ds2, 50, c.f1, 150, T, acc);
            lin_bwd_x<50, 150, 150, false>(E + ENC_L2W, ds2, 50, df1, 150, T, Ws);
            for (int e = tid; e < T * 150; e += nt) df1[e] = c.f1[e] > 0.0f ? df1[e] : 0.0f;
            __syncthreads();
            lin_bwd_w<150, 50>(GE + ENC_L1W, GE + ENC_L1B, df1, 150, c.y1, 50, T, acc);
            for (int e = tid; e < T * 50; e += nt) dy1[e] = ds2[e];
            __syncthreads();
            lin_bwd_x<150, 50, 50, true>(E + ENC_L1W, df1, 150, dy1, 50, T, Ws);
            ln_bwd(dy1, c.xh1, c.rs1, E + ENC_N1G, ds1, GE + ENC_N1G, GE + ENC_N1B, T, acc);
            lin_bwd_w<50, 50>(GE + ENC_OW, GE + ENC_OB, ds1, 50, c.ctx, 50, T, acc);
            lin_bwd_x<50, 50, 50, false>(E + ENC_OW, ds1, 50, dctx, 50, T, Ws);
            // attention: dP = dctx v^T ; dv = P^T dctx ; dS = P (dP - sum_c P dP) ; dq = dS k / 5 ; dk = dS^T q / 5
            for (int e = tid; e < 2 * T * T; e += nt) {
                const int hd = e / (T * T), rc = e - hd * T * T, r = rc / T, cc = rc - r * T;
                const float *dc = dctx + r * 50 + 25 * hd, *v = c.qkv + cc * 150 + 100 + 25 * hd;
                float s = 0.0f;
#pragma unroll
                for (int d = 0; d < 25; ++d) s = fmaf(dc[d], v[d], s);
                dS[e] = s;
            }
            __syncthreads();
            for (int e = tid; e < 2 * T; e += nt) {
                float *d = dS + e * T;
                const float *p = c.P + e * T;
                float s = 0.0f;
                for (int cc = 0; cc < T; ++cc) s = fmaf(p[cc], d[cc], s);
                for (int cc = 0; cc < T; ++cc) d[cc] = p[cc] * (d[cc] - s);
            }
            __syncthreads();
            for (int e = tid; e < T * 150; e += nt) {
                const int r = e / 150, j = e - r * 150, part = j / 50, jj = j - part * 50, hd = jj / 25;
                float s = 0.0f;
                if (part == 0) {            // dq[r] = sum_c dS[r][c] k[c] / 5
                    for (int cc = 0; cc < T; ++cc) s = fmaf(dS[(hd * T + r) * T + cc], c.qkv[cc * 150 + 50 + jj], s);
                    s *= 0.2f;
                } else if (part == 1) {     // dk[r] = sum_q dS[q][r] q[q] / 5
                    for (int qq = 0; qq < T; ++qq) s = fmaf(dS[(hd * T + qq) * T + r], c.qkv[qq * 150 + jj], s);
                    s *= 0.2f;
                } else {                    // dv[r] = sum_q P[q][r] dctx[q]
                    for (int qq = 0; qq < T; ++qq) s = fmaf(c.P[(hd * T + qq) * T + r], dctx[qq * 50 + jj], s);
                }
                dqkv[e] = s;
            }
            __syncthreads();
            lin_bwd_w<150, 50>(GE + ENC_INW, GE + ENC_INB, dqkv, 150, c.X, 50, T, acc);
            for (int e = tid; e < T * 50; e += nt) dX[e] = ds1[e];
            __syncthreads();
            lin_bwd_x<150, 50, 50, true>(E + ENC_INW, dqkv, 150, dX, 50, T, Ws);
        }
        // GRU-ACT: out = (sum_k p_k h_k) / K
        for (int e = tid; e < T * 50; e += nt) {
            dX[e] *= invK;                       // dA
            dh[e] = 0.0f;
        }
        __syncthreads();
        const float *x = st + L.x;
        for (int k = K; k >= 1; --k) {
            const bool gacc = acc || k != K;
            const float *hp = st + L.h + (k - 1) * T * 50, *hn = st + L.h + k * T * 50;
            GruStash g = gru_stash(st + L.gru + (k - 1) * L.gru_step, T);
            float *dlog = small_ + 8;            // [T] (T <= 33)
            float *hxb = t2, *rhx = t3, *daq = t4, *dz = t5, *d1 = t6, *drh = t7;
            {   // dlogit[r] = (dA[r] . h_k[r]) p (1 - p)
                const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
                for (int r = warp; r < T; r += nw) {
                    float s = dX[r * 50 + lane] * hn[r * 50 + lane];
                    if (lane + 32 < 50) s = fmaf(dX[r * 50 + lane + 32], hn[r * 50 + lane + 32], s);
                    for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
                    if (lane == 0) dlog[r] = s * g.p[r] * (1.0f - g.p[r]);
                }
            }
            __syncthreads();
            for (int j = tid; j < 51; j += nt) {     // ponder_linear gradient
                float s = 0.0f;
                if (j < 50) for (int r = 0; r < T; ++r) s = fmaf(dlog[r], hn[r * 50 + j], s);
                else for (int r = 0; r < T; ++r) s += dlog[r];
                G[OFF_PW + j] = gacc ? G[OFF_PW + j] + s : s;          // OFF_PB == OFF_PW + 50
            }
            for (int e = tid; e < T * 63; e += nt) {
                const int r = e / 63, j = e - r * 63;
                const float xv = j < 50 ? 0.0f : x[r * 13 + (j - 50)];
                hxb[r * 64 + j] = j < 50 ? hp[r * 50 + j] : xv;
                rhx[r * 64 + j] = j < 50 ? g.r[r * 50 + j] * hp[r * 50 + j] : xv;
            }
            for (int e = tid; e < T * 50; e += nt) {
                const int r = e / 50, j = e - r * 50;
                const float d = dh[e] + g.p[r] * dX[e] + dlog[r] * __ldg(w + OFF_PW + j);      // total dL/dh_k
                const float z = g.z[e], q = g.q[e];
                dz[e] = g.az2[e] > 0.0f ? d * (q - hp[e]) * z * (1.0f - z) : 0.0f;             // through sigmoid and ReLU
                daq[e] = d * z * (1.0f - q * q);
                dh[e] = d * (1.0f - z);                                                          // -> dL/dh_{k-1}, first term
            }
            __syncthreads();
            lin_bwd_w<50, 63>(G + OFF_QW, G + OFF_QB, daq, 50, rhx, 64, T, gacc);
            lin_bwd_x<50, 63, 50, false>(w + OFF_QW, daq, 50, drh, 50, T, Ws);
            for (int e = tid; e < T * 50; e += nt) {
                const float rr = g.r[e];
                dh[e] = fmaf(drh[e], rr, dh[e]);
                drh[e] = g.ar2[e] > 0.0f ? drh[e] * hp[e] * rr * (1.0f - rr) : 0.0f;           // d ar2 (pre-sigmoid, post-ReLU mask)
            }
            __syncthreads();
            // update gate
            lin_bwd_w<50, 100>(G + OFF_Z2W, G + OFF_Z2B, dz, 50, g.az1, 100, T, gacc);
            lin_bwd_x<50, 100, 100, false>(w + OFF_Z2W, dz, 50, d1, 100, T, Ws);
            for (int e = tid; e < T * 100; e += nt) d1[e] = g.az1[e] > 0.0f ? d1[e] : 0.0f;
            __syncthreads();
            lin_bwd_w<100, 63>(G + OFF_Z1W, G + OFF_Z1B, d1, 100, hxb, 64, T, gacc);
            lin_bwd_x<100, 63, 50, true>(w + OFF_Z1W, d1, 100, dh, 50, T, Ws);
            // reset gate
            lin_bwd_w<50, 100>(G + OFF_R2W, G + OFF_R2B, drh, 50, g.ar1, 100, T, gacc);
            lin_bwd_x<50, 100, 100, false>(w + OFF_R2W, drh, 50, d1, 100, T, Ws);
            for (int e = tid; e < T * 100; e += nt) d1[e] = g.ar1[e] > 0.0f ? d1[e] : 0.0f;
            __syncthreads();
            lin_bwd_w<100, 63>(G + OFF_R1W, G + OFF_R1B, d1, 100, hxb, 64, T, gacc);
            lin_bwd_x<100, 63, 50, true>(w + OFF_R1W, d1, 100, dh, 50, T, Ws);
        }
    }
}

// grad[e] = sum_c part[c][e] in CTA order; block 0 also writes the loss and the step count
__global__ void __launch_bounds__(256) train_reduce_kernel(const float *__restrict__ part, int G, float *__restrict__ grad,
                                                            const float *__restrict__ sqerr, int B, const int *flags,
                                                            float *loss, int32_t *steps)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < NUM_WEIGHTS) {
        float s = 0.0f;
        for (int c = 0; c < G; ++c) s += part[(size_t)c * NUM_WEIGHTS + e];
        grad[e] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        float s = 0.0f;
        for (int b = threadIdx.x; b < B; b += 32) s += sqerr[b];
        for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        if (threadIdx.x == 0) {
            if (loss) *loss = s / (float)B;
            if (steps) *steps = flags[0] == 0 ? 1 : (flags[1] == 0 ? 2 : 3);
        }
    }
}

// torch.optim.Adam (no weight decay, no amsgrad), single-tensor formulation of torch/optim/adam.py
__global__ void __launch_bounds__(256) adam_kernel(size_t count, float *__restrict__ p, const float *__restrict__ g,
                                                    float *__restrict__ m, float *__restrict__ v, float step_size, float b1,
                                                    float b2, float eps, float bc2_sqrt, float gscale)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const float gr = g[e] * gscale;
    const float mm = m[e] + (gr - m[e]) * (1.0f - b1);          // exp_avg.lerp_(grad, 1 - beta1)
    const float vv = v[e] * b2 + (1.0f - b2) * gr * gr;         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[e] = mm;
    v[e] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    p[e] = p[e] - step_size * (mm / denom);
}

// ---- launchers ------------------------------------------------------------------------------------------------------
size_t train_smem_bytes(int T, bool stash_in_smem)
{
    const TrLayout L = tr_layout(T);
    size_t f = TR_WS_FLOATS + (size_t)TR_NBUF * T * TR_BUF_PITCH + 64 + (stash_in_smem ? L.total : 0);
    return f * sizeof(float);
}

size_t probe_smem_bytes(int T)
{
    const TrLayout L = tr_layout(T);
    return (TR_WS_FLOATS + (size_t)T * 50 + T * 64 + ((3 * T + 3) & ~3) + L.gru + L.gru_step) * sizeof(float);   // x, h, one step
}

cudaError_t launch_train_grad(aem_handle h, int B, int n, const float *params, const float *states, const float *targets,
                              float *grad, float *loss, float *values, int32_t *steps, cudaStream_t st)
{
    const int T = n + 1;
    const TrLayout L = tr_layout(T);
    const size_t smem_limit = 227 * 1024;
    const bool in_smem = train_smem_bytes(T, true) <= smem_limit;
    const int grid = B < h->num_sms ? B : h->num_sms;
    cudaError_t e;
    if (grid > h->tr_grid || B > h->tr_B || (!in_smem && (size_t)grid * L.total > h->tr_stash_floats)) {   // (re)allocate scratch
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        cudaFree(h->tr_part); cudaFree(h->tr_sqerr); cudaFree(h->tr_flags); cudaFree(h->tr_stash);
        h->tr_part = h->tr_sqerr = h->tr_stash = nullptr; h->tr_flags = nullptr;
        h->tr_grid = grid > h->tr_grid ? grid : h->tr_grid;
        h->tr_B = B > h->tr_B ? B : h->tr_B;
        if ((e = cudaMalloc(&h->tr_part, (size_t)h->tr_grid * NUM_WEIGHTS * sizeof(float))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&h->tr_sqerr, (size_t)h->tr_B * sizeof(float))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&h->tr_flags, 2 * sizeof(int))) != cudaSuccess) return e;
        h->tr_stash_floats = in_smem ? 0 : (size_t)h->tr_grid * tr_layout(AEM_MAX_HUMANS + 1).total;
        if (h->tr_stash_floats && (e = cudaMalloc(&h->tr_stash, h->tr_stash_floats * sizeof(float))) != cudaSuccess) return e;
    }
    TrainArgs a;
    a.params = params; a.states = states; a.targets = targets;
    a.part = h->tr_part; a.sqerr = h->tr_sqerr; a.values = values; a.flags = h->tr_flags;
    a.stash_g = in_smem ? nullptr : h->tr_stash;
    a.B = B; a.n = n;
    if ((e = cudaMemsetAsync(h->tr_flags, 0, 2 * sizeof(int), st)) != cudaSuccess) return e;
    const size_t psm = probe_smem_bytes(T), fsm = train_smem_bytes(T, in_smem);
    if (psm > smem_limit) return cudaErrorInvalidValue;
    if ((e = cudaFuncSetAttribute(train_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(train_fwdbwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm)) != cudaSuccess) return e;
    train_probe_kernel<<<grid, TR_THREADS, psm, st>>>(a);
    train_fwdbwd_kernel<<<grid, TR_THREADS, fsm, st>>>(a);
    train_reduce_kernel<<<(NUM_WEIGHTS + 255) / 256, 256, 0, st>>>(h->tr_part, grid, grad, h->tr_sqerr, B, h->tr_flags, loss, steps);
    return cudaGetLastError();
}

cudaError_t launch_adam(size_t count, float *p, const float *g, float *m, float *v, double lr, double b1, double b2, double eps,
                        int step, double gscale, cudaStream_t st)
{
    const double bc1 = 1.0 - pow(b1, (double)step);             // torch computes these in Python floats
    const double bc2_sqrt = sqrt(1.0 - pow(b2, (double)step));
    adam_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(count, p, g, m, v, (float)(lr / bc1), (float)b1, (float)b2,
                                                                 (float)eps, (float)bc2_sqrt, (float)gscale);
    return cudaGetLastError();
}

}  // namespace aem
