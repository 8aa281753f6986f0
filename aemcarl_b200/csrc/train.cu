// train.cu -- the value-network regression step as hand-written kernels:
//   forward of ValueNetworkTransformerAndGRU (qnetwork_factory.py:641-687, components.py:38-137) on a batch of rotated
//   states with the reference's BATCH-GLOBAL adaptive-computation-time halting (components.py:132-134), MSE against the
//   targets (trainer.py:53,76), the complete backward pass, and torch.optim.Adam's update (trainer.py:30).
//
// Shape of the problem: 100 samples x (n + 1) tokens x 113,752 parameters -- every weight meets only T = n + 1 rows per
// sample, so the step is bound by latency (a chain of ~90 dependent small products per sample), not by FLOPs or HBM.
//   * train_probe_kernel   one CTA per sample runs the three GRU steps and ORs "some row still below 1 - epsilon after
//                          step k" into two global flags: the batch-global step count K every sample must use.  Its GRU
//                          intermediates are handed to the next kernel through an L2-resident scratch (no recompute).
//   * train_fwdbwd_kernel  one CTA (512 threads) per sample (grid-stride if the batch exceeds the grid): forward with every
//                          intermediate kept in shared memory (80 KB at T = 6; a global scratch slice when it does not
//                          fit), then the hand-derived backward (tests/manual_backward.py is the same derivation in numpy,
//                          checked against autograd).  Weights arrive one layer ahead by one cp.async.bulk per layer into a
//                          two-slot ring.  y = W x and dx = W^T dy run on mma.sync.m16n8k8 (3xTF32, FP32-grade); the weight
//                          gradient dW = dy^T x is an FP32 outer product with coalesced stores into the CTA's OWN slice
//                          part[cta][113,752] -- no atomics.
//   * train_reduce_kernel  grad[e] = sum over CTAs in a fixed order (deterministic), loss = mean squared error.
//   * adam_kernel / sgd_kernel / rmsprop_kernel   one thread per parameter (the three optimizers of trainer.py:24-34).
//   * sample_batch_kernel  a uniform mini-batch without replacement from the replay ring + gather, one launch.
// The gradient all-reduce of a data-parallel run sits between train_reduce_kernel and the optimizer kernel (trainer.py).
#include "aem_common.cuh"
#include "tc_common.cuh"

namespace aem {

constexpr int TR_THREADS = 512;
constexpr int TR_WS_FLOATS = 7520;      // largest staged layer: 150 x 50 (+ up to 3 floats of alignment head, + tail)
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
    L.x = o; o += (T * 13 + 3) & ~3;
    L.h = o; o += 4 * T * 50;                      // h_0 .. h_3
    L.gru_step = T * 450 + ((T + 3) & ~3);                       // az1 100 | az2 50 | z 50 | ar1 100 | ar2 50 | r 50 | q 50 | p 1
    L.gru = o; o += 3 * L.gru_step;
    L.enc_layer = T * 550 + 2 * ((T + 3) & ~3) + ((2 * T * T + 3) & ~3);           // X 50 | qkv 150 | ctx 50 | xh1 50 | y1 50 | f1 150 | xh2 50 | rs1 | rs2 ; P
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
    e.f1 = e.y1 + T * 50; e.xh2 = e.f1 + T * 150; e.rs1 = e.xh2 + T * 50; e.rs2 = e.rs1 + ((T + 3) & ~3);
    e.P = e.rs2 + ((T + 3) & ~3);
    return e;
}

__device__ __forceinline__ float sigmoidf_(float a) { return 1.0f / (1.0f + expf(-a)); }

// ---- weight pipeline: W[O][I] (row-major, global) -> shared memory by ONE 1-D bulk copy (TMA), one layer ahead of its use
// Protocol at every layer:  Ws = pipe.take();  pipe.issue(<the NEXT layer's weights>);  <compute with Ws>;  psync();
// issue(): thread 0 arms the buffer's mbarrier with the byte count and starts cp.async.bulk on the 16-byte aligned
// superset of the tensor (blob offsets are only 4-byte aligned); take(): every thread waits on that mbarrier.  The buffer
// being refilled was last read two phases ago, i.e. before the previous psync().
#ifdef AEM_TRAIN_TIMELINE
// debug build only (python -m aemcarl_b200.build with AEM_TRAIN_TIMELINE=1): clock64 of CTA 0 at every barrier phase
__device__ long long g_tr_tl[2048];
__device__ int g_tr_tl_n;
__device__ __forceinline__ void psync()
{
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && g_tr_tl_n < 2048) g_tr_tl[g_tr_tl_n++] = clock64();
}
#else
__device__ __forceinline__ void psync() { __syncthreads(); }
#endif
struct WPipe {
    float *b0, *b1;
    uint64_t *bars;          // [2] in shared memory
    int cur, head0, head1;
    uint32_t par0, par1;
    __device__ __forceinline__ void init(float *buf0, float *buf1, uint64_t *mb)
    {
        b0 = buf0; b1 = buf1; bars = mb; cur = 0; head0 = head1 = 0; par0 = par1 = 0;
        if (threadIdx.x == 0) {
            aemtc::mbar_init(mb, 1);
            aemtc::mbar_init(mb + 1, 1);
            aemtc::fence_mbar_init();
        }
        __syncthreads();
    }
    __device__ __forceinline__ void issue(const float *__restrict__ Wg, int O, int I)
    {
        if (!Wg) return;
        const uintptr_t s = (uintptr_t)Wg, al = s & ~(uintptr_t)15;
        const int head = (int)(s - al) / 4;
        if (cur) head1 = head; else head0 = head;
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)(((s - al) + (size_t)O * I * 4 + 15) & ~(size_t)15);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic reads of this buffer
            aemtc::mbar_expect_tx(bars + cur, bytes);
            aemtc::bulk_g2s(cur ? b1 : b0, (const void *)al, bytes, bars + cur);
        }
    }
    __device__ __forceinline__ const float *take()
    {
        const float *p;
        if (cur) { aemtc::mbar_wait(bars + 1, par1); par1 ^= 1; p = b1 + head1; }
        else { aemtc::mbar_wait(bars, par0); par0 ^= 1; p = b0 + head0; }
        cur ^= 1;
#ifdef AEM_TRAIN_TIMELINE
        if (blockIdx.x == 0 && threadIdx.x == 0 && g_tr_tl_n < 2048) g_tr_tl[g_tr_tl_n++] = -clock64();     // weights arrived
#endif
        return p;
    }
};

// Two of the three products of a layer (y = W x, dx = W^T dy) run on the warp-level tensor-core instruction
// mma.sync.m16n8k8 (TF32 inputs, FP32 accumulate) with the 3xTF32 split (hi = upper 19 bits, lo = x - hi; hi*hi + lo*hi + hi*lo),
// i.e. FP32-grade results.  The 16-wide M dimension always carries WEIGHT-SHAPED indices (outputs or columns), the 8-wide N
// dimension the token rows (T = 6 at the trainer's shape) or the other weight index, so 75 % of every instruction is useful
// work and a 150 x 50 layer is 70 tiles.  (tcgen05 needs M = 128 rows and a commit / wait round trip per dependent step; for a
// chain of ~90 dependent products over 6 rows the synchronous warp-level instruction is the right tool.)  Fragment layout
// (PTX ISA, m16n8k8 .tf32): g = lane / 4, t = lane % 4;  A: a0 (g, t) a1 (g + 8, t) a2 (g, t + 4) a3 (g + 8, t + 4);
// B: b0 (k = t, n = g) b1 (k = t + 4, n = g);  C: c0 (g, 2t) c1 (g, 2t + 1) c2 (g + 8, 2t) c3 (g + 8, 2t + 1).
__device__ __forceinline__ void tf32_split(float x, uint32_t &hi, uint32_t &lo)
{
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
struct FragA { uint32_t hi[4], lo[4]; };
struct FragB { uint32_t hi[2], lo[2]; };
__device__ __forceinline__ FragA make_a(float a0, float a1, float a2, float a3)
{
    FragA f;
    tf32_split(a0, f.hi[0], f.lo[0]); tf32_split(a1, f.hi[1], f.lo[1]);
    tf32_split(a2, f.hi[2], f.lo[2]); tf32_split(a3, f.hi[3], f.lo[3]);
    return f;
}
__device__ __forceinline__ FragB make_b(float b0, float b1)
{
    FragB f;
    tf32_split(b0, f.hi[0], f.lo[0]); tf32_split(b1, f.hi[1], f.lo[1]);
    return f;
}
__device__ __forceinline__ void mma3(float (&c)[4], const FragA &a, const FragB &b)
{
    mma_tf32(c, a.lo, b.hi);
    mma_tf32(c, a.hi, b.lo);
    mma_tf32(c, a.hi, b.hi);
}

// epi(r, o, b[o] + sum_i X[r][i] W[o][i]):  Y^T [O x T] = W [O x I] X^T [I x T];  M = outputs, K = columns, N = rows
template <int O, int I, class Epi>
__device__ __forceinline__ void lin_fwd(const float *Ws, const float *__restrict__ bg, const float *X, int ldx, int T, Epi epi)
{
    static_assert(O * I + 8 <= TR_WS_FLOATS, "staging buffer");
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    constexpr int MT = (O + 15) / 16;
    for (int mt = warp; mt < MT; mt += nw) {
        const int oa = mt * 16 + g, ob = oa + 8;
        const bool va = oa < O, vb = ob < O;
        const float *wa = Ws + (va ? oa : 0) * I, *wb = Ws + (vb ? ob : 0) * I;
        const float bias_a = va ? __ldg(bg + oa) : 0.0f, bias_b = vb ? __ldg(bg + ob) : 0.0f;
        for (int r0 = 0; r0 < T; r0 += 8) {
            const int row = r0 + g;
            const bool rv = row < T;
            const float *x = X + (rv ? row : 0) * ldx;
            float c0[4] = {0.0f, 0.0f, 0.0f, 0.0f}, c1[4] = {0.0f, 0.0f, 0.0f, 0.0f};      // two chains: even / odd K steps
#pragma unroll 2
            for (int i0 = 0; i0 < I; i0 += 16) {
                {
                    const int ia = i0 + t, ib = ia + 4;
                    const bool ka = ia < I, kb = ib < I;
                    const FragA a = make_a(va && ka ? wa[ia] : 0.0f, vb && ka ? wb[ia] : 0.0f, va && kb ? wa[ib] : 0.0f,
                                           vb && kb ? wb[ib] : 0.0f);
                    mma3(c0, a, make_b(rv && ka ? x[ia] : 0.0f, rv && kb ? x[ib] : 0.0f));
                }
                if (i0 + 8 < I) {
                    const int ia = i0 + 8 + t, ib = ia + 4;
                    const bool ka = ia < I, kb = ib < I;
                    const FragA a = make_a(va && ka ? wa[ia] : 0.0f, vb && ka ? wb[ia] : 0.0f, va && kb ? wa[ib] : 0.0f,
                                           vb && kb ? wb[ib] : 0.0f);
                    mma3(c1, a, make_b(rv && ka ? x[ia] : 0.0f, rv && kb ? x[ib] : 0.0f));
                }
            }
            const int ra = r0 + 2 * t, rb = ra + 1;
            if (va && ra < T) epi(ra, oa, c0[0] + c1[0] + bias_a);
            if (va && rb < T) epi(rb, oa, c0[1] + c1[1] + bias_a);
            if (vb && ra < T) epi(ra, ob, c0[2] + c1[2] + bias_b);
            if (vb && rb < T) epi(rb, ob, c0[3] + c1[3] + bias_b);
        }
    }
    psync();
}

// epi(r, i, sum_o dY[r][o] W[o][i]) for i < ICNT:  dX^T [I x T] = W^T [I x O] dY^T [O x T];  M = columns, K = outputs, N = rows
template <int O, int I, int ICNT, class Epi>
__device__ __forceinline__ void lin_bwd_x(const float *Ws, const float *dY, int lddy, int T, Epi epi)
{
    static_assert(O * I + 8 <= TR_WS_FLOATS, "staging buffer");
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    constexpr int MT = (ICNT + 15) / 16;
    for (int mt = warp; mt < MT; mt += nw) {
        const int ia = mt * 16 + g, ib = ia + 8;
        const bool va = ia < ICNT, vb = ib < ICNT;
        const float *wa = Ws + (va ? ia : 0), *wb = Ws + (vb ? ib : 0);
        for (int r0 = 0; r0 < T; r0 += 8) {
            const int row = r0 + g;
            const bool rv = row < T;
            const float *dy = dY + (rv ? row : 0) * lddy;
            float c0[4] = {0.0f, 0.0f, 0.0f, 0.0f}, c1[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 2
            for (int o0 = 0; o0 < O; o0 += 16) {
                {
                    const int oa = o0 + t, ob = oa + 4;
                    const bool ka = oa < O, kb = ob < O;
                    const FragA a = make_a(va && ka ? wa[oa * I] : 0.0f, vb && ka ? wb[oa * I] : 0.0f, va && kb ? wa[ob * I] : 0.0f,
                                           vb && kb ? wb[ob * I] : 0.0f);
                    mma3(c0, a, make_b(rv && ka ? dy[oa] : 0.0f, rv && kb ? dy[ob] : 0.0f));
                }
                if (o0 + 8 < O) {
                    const int oa = o0 + 8 + t, ob = oa + 4;
                    const bool ka = oa < O, kb = ob < O;
                    const FragA a = make_a(va && ka ? wa[oa * I] : 0.0f, vb && ka ? wb[oa * I] : 0.0f, va && kb ? wa[ob * I] : 0.0f,
                                           vb && kb ? wb[ob * I] : 0.0f);
                    mma3(c1, a, make_b(rv && ka ? dy[oa] : 0.0f, rv && kb ? dy[ob] : 0.0f));
                }
            }
            const int ra = r0 + 2 * t, rb = ra + 1;
            if (va && ra < T) epi(ra, ia, c0[0] + c1[0]);
            if (va && rb < T) epi(rb, ia, c0[1] + c1[1]);
            if (vb && ra < T) epi(ra, ib, c0[2] + c1[2]);
            if (vb && rb < T) epi(rb, ib, c0[3] + c1[3]);
        }
    }
    psync();
}

constexpr int TR_RC = 6;        // rows per register chunk of the weight-gradient product (T = 6 at the trainer's shape)

// gW[o][i] (+)= sum_r dY[r][o] X[r][i];  gb[o] (+)= sum_r dY[r][o]   (the CTA's own gradient slice: plain stores).
// One thread per column i and group of rows o: X[.][i] stays in registers while the thread walks its o's.
// No barrier: the caller's next phase only READS dY and X.
template <int O, int I>
__device__ __forceinline__ void lin_bwd_w(float *gW, float *gb, const float *dY, int lddy, const float *X, int ldx, int T, bool acc)
{
    // a thread owns TWO adjacent columns (i0, i0 + 1) and every ng-th output: each broadcast dY load feeds two FMAs (the phase
    // is bound by shared-memory loads), four outputs are in flight (the per-output chain LDS -> FMAs -> STG is serial)
    constexpr int IH = (I + 1) / 2;
    const int ip = threadIdx.x % IH, og = threadIdx.x / IH, ng = blockDim.x / IH;
    const int i0 = 2 * ip, i1 = i0 + 1;
    const bool v1 = i1 < I;
    if (og < ng) {
        for (int r0 = 0; r0 < T; r0 += TR_RC) {
            const int rows = T - r0 < TR_RC ? T - r0 : TR_RC;
            const bool add = acc || r0 > 0;
            float x0[TR_RC], x1[TR_RC];
#pragma unroll
            for (int u = 0; u < TR_RC; ++u) {
                x0[u] = u < rows ? X[(r0 + u) * ldx + i0] : 0.0f;
                x1[u] = u < rows && v1 ? X[(r0 + u) * ldx + i1] : 0.0f;
            }
            const float *dy = dY + r0 * lddy;
            int o = og;
            for (; o + 3 * ng < O; o += 4 * ng) {
                float old0[4], old1[4], s0[4], s1[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    old0[j] = add ? gW[(o + j * ng) * I + i0] : 0.0f;
                    old1[j] = add && v1 ? gW[(o + j * ng) * I + i1] : 0.0f;
                    s0[j] = s1[j] = 0.0f;
                }
#pragma unroll
                for (int u = 0; u < TR_RC; ++u)
                    if (u < rows) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float d = dy[u * lddy + o + j * ng];
                            s0[j] = fmaf(d, x0[u], s0[j]);
                            s1[j] = fmaf(d, x1[u], s1[j]);
                        }
                    }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    gW[(o + j * ng) * I + i0] = old0[j] + s0[j];
                    if (v1) gW[(o + j * ng) * I + i1] = old1[j] + s1[j];
                }
            }
            for (; o < O; o += ng) {
                const float old0 = add ? gW[o * I + i0] : 0.0f, old1 = add && v1 ? gW[o * I + i1] : 0.0f;
                float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
                for (int u = 0; u < TR_RC; ++u)
                    if (u < rows) {
                        const float d = dy[u * lddy + o];
                        s0 = fmaf(d, x0[u], s0);
                        s1 = fmaf(d, x1[u], s1);
                    }
                gW[o * I + i0] = old0 + s0;
                if (v1) gW[o * I + i1] = old1 + s1;
            }
        }
    }
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
        const float old = acc ? gb[o] : 0.0f;
        float s = 0.0f;
        for (int r = 0; r < T; ++r) s += dY[r * lddy + o];
        gb[o] = old + s;
    }
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
    psync();
}

// dS = rstd (dxhat - mean(dxhat) - xhat mean(dxhat xhat)), dxhat = dY g;  gg += sum_r dY xhat;  gbeta += sum_r dY.
// dS must not alias dY (one phase).
__device__ void ln_bwd(const float *dY, const float *xhat, const float *rstd, const float *__restrict__ g, float *dS,
                       float *gg, float *gbeta, int T, bool acc)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = threadIdx.x; j < 50; j += blockDim.x) {
        const float og = acc ? gg[j] : 0.0f, ob = acc ? gbeta[j] : 0.0f;
        float sg = 0.0f, sb = 0.0f;
        for (int r = 0; r < T; ++r) {
            sg = fmaf(dY[r * 50 + j], xhat[r * 50 + j], sg);
            sb += dY[r * 50 + j];
        }
        gg[j] = og + sg;
        gbeta[j] = ob + sb;
    }
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
    psync();
}

// robot pseudo-token + the n human rows (qnetwork_factory.py:651-663); also the x part of hxb = [h ; x] and h_0 = 0, A = 0
__device__ void load_tokens(const float *__restrict__ state, int n, float *x, float *hxb, float *h0, float *A)
{
    const int T = n + 1;
    for (int e = threadIdx.x; e < T * 13; e += blockDim.x) {
        const int r = e / 13, j = e - r * 13;
        float v;
        if (r > 0) v = __ldg(state + (r - 1) * 13 + j);
        else if (j < 6) v = __ldg(state + j);
        else if (j == 8) v = __ldg(state + 4);
        else if (j == 9) v = __ldg(state + 5);
        else if (j == 10 || j == 12) v = __ldg(state + 3);
        else v = 0.0f;
        x[e] = v;
        hxb[r * 64 + 50 + j] = v;
    }
    for (int e = threadIdx.x; e < T * 50; e += blockDim.x) {
        h0[e] = 0.0f;
        A[e] = 0.0f;
        hxb[(e / 50) * 64 + (e % 50)] = 0.0f;
    }
    psync();
}

// ---- the GRU-ACT forward (components.py:46-53,107-137): K steps, everything kept in the stash; A = sum_k p_k h_k ----
// Expects Z1W in flight and hxb = [h_0 ; x]; leaves (Wnext, On, In) in flight.  6 barrier phases per step.
// pk (optional): pk[(k - 1) * T + r] = halting probability of row r at step k
// fixed_steps > 0 (act_fixed): the output is h_K itself, i.e. the "halting probability" is the constant [k == fixed_steps]
// and nothing flows into ponder_linear.
__device__ void gru_forward(const float *__restrict__ w, float *st, const TrLayout &L, int T, int K, float *A, float *hxb,
                            WPipe &pipe, float *pk, const float *Wnext, int On, int In, int fixed_steps)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int k = 1; k <= K; ++k) {
        const float *hp = st + L.h + (k - 1) * T * 50;
        float *hn = st + L.h + k * T * 50;
        GruStash g = gru_stash(st + L.gru + (k - 1) * L.gru_step, T);
        const float *Ws = pipe.take();
        pipe.issue(w + OFF_Z2W, 50, 100);
        lin_fwd<100, 63>(Ws, w + OFF_Z1B, hxb, 64, T, [=](int r, int o, float a) { g.az1[r * 100 + o] = fmaxf(a, 0.0f); });
        Ws = pipe.take();
        pipe.issue(w + OFF_R1W, 100, 63);
        lin_fwd<50, 100>(Ws, w + OFF_Z2B, g.az1, 100, T, [=](int r, int o, float a) {
            a = fmaxf(a, 0.0f);
            g.az2[r * 50 + o] = a;
            g.z[r * 50 + o] = sigmoidf_(a);
        });
        Ws = pipe.take();
        pipe.issue(w + OFF_R2W, 50, 100);
        lin_fwd<100, 63>(Ws, w + OFF_R1B, hxb, 64, T, [=](int r, int o, float a) { g.ar1[r * 100 + o] = fmaxf(a, 0.0f); });
        Ws = pipe.take();
        pipe.issue(w + OFF_QW, 50, 63);
        lin_fwd<50, 100>(Ws, w + OFF_R2B, g.ar1, 100, T, [=](int r, int o, float a) {
            a = fmaxf(a, 0.0f);
            const float rr = sigmoidf_(a);
            g.ar2[r * 50 + o] = a;
            g.r[r * 50 + o] = rr;
            hxb[r * 64 + o] = rr * hp[r * 50 + o];          // [r * h ; x]: nothing reads the h part of hxb in this phase
        });
        Ws = pipe.take();
        if (k < K) pipe.issue(w + OFF_Z1W, 100, 63);
        else pipe.issue(Wnext, On, In);
        lin_fwd<50, 63>(Ws, w + OFF_QB, hxb, 64, T, [=](int r, int o, float a) {
            const float q = tanhf(a), z = g.z[r * 50 + o], h = hp[r * 50 + o];
            g.q[r * 50 + o] = q;
            hn[r * 50 + o] = (1.0f - z) * h + z * q;
        });
        for (int r = warp; r < T; r += nw) {                // halting probability, A += p h_k, hxb = [h_k ; x]
            const bool two = lane + 32 < 50;
            const float h0 = hn[r * 50 + lane], h1 = two ? hn[r * 50 + lane + 32] : 0.0f;
            float s = h0 * __ldg(w + OFF_PW + lane);
            if (two) s = fmaf(h1, __ldg(w + OFF_PW + lane + 32), s);
            for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
            const float p = fixed_steps > 0 ? (k == fixed_steps ? 1.0f : 0.0f) : sigmoidf_(s + __ldg(w + OFF_PB));
            if (lane == 0) {
                g.p[r] = p;
                if (pk) pk[(k - 1) * T + r] = p;
            }
            A[r * 50 + lane] = fmaf(p, h0, A[r * 50 + lane]);
            hxb[r * 64 + lane] = h0;
            if (two) {
                A[r * 50 + lane + 32] = fmaf(p, h1, A[r * 50 + lane + 32]);
                hxb[r * 64 + lane + 32] = h1;
            }
        }
        psync();
    }
}

struct TrainArgs {
    const float *params, *states, *targets;
    float *part;        // [grid][NUM_WEIGHTS]
    float *sqerr;       // [B]
    float *values;      // [B] or null
    int *flags;         // [2]
    float *stash_g;     // [grid][stash] when the stash does not fit in shared memory, else null
    float *gru_g;       // [B][layout.enc]: the probe's GRU-ACT intermediates (x, h_0..3, three step blocks) handed to the
                        // forward / backward kernel, or null (the probe keeps one step block, the main kernel recomputes)
    int B, n;
    int fixed_steps;    // act_fixed (components.py:109-112): run exactly this many GRU steps and hand on h_K; 0 = adaptive
    // dropout of nn.TransformerEncoderLayer (qnetwork_factory.py:628: p = 0.1 at 4 sites per layer, live because the reference
    // never calls eval()): inverted-dropout masks from a counter-based generator, recomputed in the backward pass
    float drop_p;                              // 0 = eval() mode
    unsigned long long drop_seed, drop_ctr;    // ctr = one value per optimisation step
    const unsigned long long *drop_ctr_dev;    // or read it from the device step state (aem_train_step_adam: dev_state[1])
};

// Dropout mask of one element: 0 or 1 / (1 - p).  key = (sample, layer, site, element index); sites: 0 = attention weights
// [2][T][T], 1 = attention branch before the residual [T][50], 2 = FFN hidden after the ReLU [T][150], 3 = FFN branch before
// the residual [T][50].  tests/test_gpu_train.py regenerates the same masks on the host (dropout_masks_host) and feeds them
// to the float64 derivation (tests/manual_backward.py = torch autograd with injected masks).
struct DropGen {
    unsigned long long base;
    unsigned int thresh;        // keep iff (hash >> 40) >= thresh, thresh = p * 2^24
    float scale;
    __device__ __forceinline__ float operator()(int b, int l, int site, int idx) const
    {
        if (thresh == 0u) return 1.0f;
        unsigned long long z = base + ((((unsigned long long)b * 12ull + (unsigned long long)(l * 4 + site)) << 16) | (unsigned long long)idx) *
                                          0xD1B54A32D192ED03ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (unsigned int)(z >> 40) >= thresh ? scale : 0.0f;
    }
};
__device__ __forceinline__ DropGen make_dropgen(const TrainArgs &a)
{
    DropGen g;
    const unsigned long long ctr = a.drop_ctr_dev ? a.drop_ctr_dev[1] : a.drop_ctr;
    g.base = a.drop_seed * 0x9E3779B97F4A7C15ull + ctr * 0xC2B2AE3D27D4EB4Full;
    g.thresh = a.drop_p > 0.0f ? (unsigned int)(a.drop_p * 16777216.0f) : 0u;
    g.scale = a.drop_p > 0.0f ? 1.0f / (1.0f - a.drop_p) : 1.0f;
    return g;
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TR_THREADS, 1) train_probe_kernel(TrainArgs a)
{
    extern __shared__ __align__(16) float sm[];
    const int T = a.n + 1;
    TrLayout L = tr_layout(T);
    if (!a.gru_g) L.gru_step = 0;                   // no hand-off: every step reuses one stash block
    __shared__ __align__(8) uint64_t bars[2];
    WPipe pipe;
    pipe.init(sm, sm + TR_WS_FLOATS, bars);
    float *A = sm + 2 * TR_WS_FLOATS, *hxb = A + T * 50, *pk = hxb + T * 64, *st = pk + ((3 * T + 3) & ~3);
    pipe.issue(a.params + OFF_Z1W, 100, 63);
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        load_tokens(a.states + (size_t)b * a.n * 13, a.n, st + L.x, hxb, st + L.h, A);
        const bool more = b + (int)gridDim.x < a.B;
        gru_forward(a.params, st, L, T, 3, A, hxb, pipe, pk, more ? a.params + OFF_Z1W : nullptr, 100, 63, a.fixed_steps);
        if (threadIdx.x < 32) {
            int un1 = 0, un2 = 0;
            for (int r = threadIdx.x; r < T; r += 32) {
                const float p1 = pk[r];
                const float p2 = p1 + pk[T + r];
                un1 |= p1 < 1.0f - ACT_EPS;
                un2 |= p2 < 1.0f - ACT_EPS;
            }
            if (a.fixed_steps > 0) { un1 = a.fixed_steps >= 2; un2 = a.fixed_steps >= 3; }     // K is given, not voted
            un1 = __any_sync(0xffffffffu, un1);
            un2 = __any_sync(0xffffffffu, un2);
            if (threadIdx.x == 0) {
                if (un1) atomicOr(a.flags + 0, 1);
                if (un2) atomicOr(a.flags + 1, 1);
            }
        }
        if (a.gru_g) {                              // x | h_0..h_3 | step blocks 1..3, as tr_layout lays them out
            float *dst = a.gru_g + (size_t)b * L.enc;
            for (int e = threadIdx.x; e < L.enc; e += blockDim.x) dst[e] = st[e];
        }
        psync();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// STASH_SMEM: the activation stash lives in shared memory (every access is a 32-bit LDS / STS); otherwise in the CTA's
// global scratch slice through generic pointers (large n only).
template <bool STASH_SMEM>
__global__ void __launch_bounds__(TR_THREADS, 1) train_fwdbwd_kernel(TrainArgs a)
{
    extern __shared__ __align__(16) float sm[];
    const int T = a.n + 1;
    const TrLayout L = tr_layout(T);
    __shared__ __align__(8) uint64_t bars[2];
    WPipe pipe;
    pipe.init(sm, sm + TR_WS_FLOATS, bars);
    float *buf[TR_NBUF];
    for (int i = 0; i < TR_NBUF; ++i) buf[i] = sm + 2 * TR_WS_FLOATS + i * T * TR_BUF_PITCH;
    float *small_ = buf[TR_NBUF - 1] + T * TR_BUF_PITCH;            // 64 floats
    float *st = STASH_SMEM ? small_ + 64 : a.stash_g + (size_t)blockIdx.x * L.total;
    const float *w = a.params;
    float *G = a.part + (size_t)blockIdx.x * NUM_WEIGHTS;
    const int K = a.flags[0] == 0 ? 1 : (a.flags[1] == 0 ? 2 : 3);
    const float invK = a.fixed_steps > 0 ? 1.0f : 1.0f / (float)K;     // act_fixed hands on h_K, not the mean
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    const DropGen drop = make_dropgen(a);

    if (blockIdx.x >= a.B) {                       // more CTAs than samples: a zero slice
        for (int e = tid; e < NUM_WEIGHTS; e += nt) G[e] = 0.0f;
        return;
    }
    if (a.gru_g) pipe.issue(w + OFF_ENC + ENC_INW, 150, 50);
    else pipe.issue(w + OFF_Z1W, 100, 63);
    int iter = 0;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x, ++iter) {
        const bool acc = iter > 0, more = b + (int)gridDim.x < a.B;
        const float *state = a.states + (size_t)b * a.n * 13;
        const float *Ws;
        // =========================================== forward ===========================================
        float *A = buf[0], *hxb = buf[1];
        if (a.gru_g) {                              // the probe already ran the GRU steps of this sample: take its stash
            const float *src = a.gru_g + (size_t)b * L.enc;
            for (int e = tid; e < L.enc; e += nt) st[e] = src[e];
            psync();
            for (int e = tid; e < T * 50; e += nt) {                // A = sum_{k <= K} p_k h_k, accumulated like gru_forward
                float s = 0.0f;
                for (int k = 1; k <= K; ++k)
                    s = fmaf(gru_stash(st + L.gru + (k - 1) * L.gru_step, T).p[e / 50], st[L.h + k * T * 50 + e], s);
                A[e] = s;
            }
            psync();
        } else {
            load_tokens(state, a.n, st + L.x, hxb, st + L.h, A);
            gru_forward(w, st, L, T, K, A, hxb, pipe, nullptr, w + OFF_ENC + ENC_INW, 150, 50, a.fixed_steps);
        }
        for (int l = 0; l < 3; ++l) {
            const float *E = w + OFF_ENC + l * ENC_SIZE;
            EncStash c = enc_stash(st + L.enc + l * L.enc_layer, T);
            float *Xn = l < 2 ? enc_stash(st + L.enc + (l + 1) * L.enc_layer, T).X : st + L.fin;
            float *s1 = buf[2], *s2 = buf[3];
            Ws = pipe.take();
            pipe.issue(E + ENC_OW, 50, 50);
            if (l == 0) {                           // the encoder input of layer 0 is the GRU-ACT output A / K
                lin_fwd<150, 50>(Ws, E + ENC_INB, A, 50, T, [=](int r, int o, float v) {
                    c.qkv[r * 150 + o] = fmaf(v - __ldg(E + ENC_INB + o), invK, __ldg(E + ENC_INB + o));
                    if (o < 50) c.X[r * 50 + o] = A[r * 50 + o] * invK;
                });
            } else {
                lin_fwd<150, 50>(Ws, E + ENC_INB, c.X, 50, T, [=](int r, int o, float v) { c.qkv[r * 150 + o] = v; });
            }
            for (int hr = warp; hr < 2 * T; hr += nw) {             // scores, softmax over the T keys, ctx = P v
                const int hd = hr / T, r = hr - hd * T;
                const float *q = c.qkv + r * 150 + 25 * hd;
                float *p = c.P + hr * T;
                float sc[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int cc = lane + 32 * u;
                    float s = -INFINITY;
                    if (cc < T) {
                        const float *k = c.qkv + cc * 150 + 50 + 25 * hd;
                        s = 0.0f;
#pragma unroll
                        for (int d = 0; d < 25; ++d) s = fmaf(q[d], k[d], s);
                        s *= 0.2f;                                  // (q . k) / sqrt(25)
                    }
                    sc[u] = s;
                }
                float m = fmaxf(sc[0], sc[1]);
                for (int x = 16; x; x >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, x));
                const float e0 = lane < T ? expf(sc[0] - m) : 0.0f, e1 = lane + 32 < T ? expf(sc[1] - m) : 0.0f;
                float sum = e0 + e1;
                for (int x = 16; x; x >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, x);
                const float inv = 1.0f / sum;
                if (lane < T) p[lane] = e0 * inv;
                if (lane + 32 < T) p[lane + 32] = e1 * inv;
                __syncwarp();
                if (lane < 25) {                                    // ctx = dropout(P) v  (the stash keeps the undropped P)
                    float s = 0.0f;
                    for (int cc = 0; cc < T; ++cc) s = fmaf(p[cc] * drop(b, l, 0, hr * T + cc), c.qkv[cc * 150 + 100 + 25 * hd + lane], s);
                    c.ctx[r * 50 + 25 * hd + lane] = s;
                }
            }
            psync();
            Ws = pipe.take();
            pipe.issue(E + ENC_L1W, 150, 50);
            lin_fwd<50, 50>(Ws, E + ENC_OB, c.ctx, 50, T, [=](int r, int o, float v) {
                s1[r * 50 + o] = fmaf(v, drop(b, l, 1, r * 50 + o), c.X[r * 50 + o]);
            });
            ln_fwd(s1, E + ENC_N1G, E + ENC_N1B, c.y1, c.xh1, c.rs1, T);
            Ws = pipe.take();
            pipe.issue(E + ENC_L2W, 50, 150);
            lin_fwd<150, 50>(Ws, E + ENC_L1B, c.y1, 50, T, [=](int r, int o, float v) {
                c.f1[r * 150 + o] = fmaxf(v, 0.0f) * drop(b, l, 2, r * 150 + o);
            });
            Ws = pipe.take();
            if (l < 2) pipe.issue(E + ENC_SIZE + ENC_INW, 150, 50);
            else pipe.issue(w + OFF_A0W, 100, 56);
            lin_fwd<50, 150>(Ws, E + ENC_L2B, c.f1, 150, T, [=](int r, int o, float v) {
                s2[r * 50 + o] = fmaf(v, drop(b, l, 3, r * 50 + o), c.y1[r * 50 + o]);
            });
            ln_fwd(s2, E + ENC_N2G, E + ENC_N2B, Xn, c.xh2, c.rs2, T);
        }
        float *X3 = st + L.fin, *ain = X3 + T * 50, *a1 = ain + 56, *a2 = a1 + 100;
        for (int e = tid; e < 56; e += nt) ain[e] = e < 6 ? __ldg(state + e) : X3[e - 6];
        psync();
        Ws = pipe.take();
        pipe.issue(w + OFF_A2W, 50, 100);
        lin_fwd<100, 56>(Ws, w + OFF_A0B, ain, 56, 1, [=](int, int o, float v) { a1[o] = fmaxf(v, 0.0f); });
        Ws = pipe.take();
        pipe.issue(w + OFF_A2W, 50, 100);           // again, for the backward pass
        lin_fwd<50, 100>(Ws, w + OFF_A2B, a1, 100, 1, [=](int, int o, float v) { a2[o] = fmaxf(v, 0.0f); });
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
        psync();
        // =========================================== backward ==========================================
        const float dv = small_[0];
        float *dX = buf[0], *dh = buf[1], *t2 = buf[2], *t3 = buf[3], *t4 = buf[4], *t5 = buf[5], *t6 = buf[6], *t7 = buf[7];
        {   // action MLP, one row
            float *da2 = t2, *da1 = t3;
            for (int e = tid; e < 50; e += nt) {
                const float old = acc ? G[OFF_A4W + e] : 0.0f;
                G[OFF_A4W + e] = old + dv * a2[e];
                da2[e] = a2[e] > 0.0f ? dv * __ldg(w + OFF_A4W + e) : 0.0f;
            }
            if (tid == 0) G[OFF_A4B] = (acc ? G[OFF_A4B] : 0.0f) + dv;
            psync();
            Ws = pipe.take();
            pipe.issue(w + OFF_A0W, 100, 56);
            lin_bwd_w<50, 100>(G + OFF_A2W, G + OFF_A2B, da2, 50, a1, 100, 1, acc);
            lin_bwd_x<50, 100, 100>(Ws, da2, 50, 1, [=](int, int i, float v) { da1[i] = a1[i] > 0.0f ? v : 0.0f; });
            Ws = pipe.take();
            pipe.issue(w + OFF_ENC + 2 * ENC_SIZE + ENC_L2W, 50, 150);
            lin_bwd_w<100, 56>(G + OFF_A0W, G + OFF_A0B, da1, 100, ain, 56, 1, acc);
            for (int e = 50 + tid; e < T * 50; e += nt) dX[e] = 0.0f;                    // only token 0 feeds the head
            lin_bwd_x<100, 56, 56>(Ws, da1, 100, 1, [=](int, int i, float v) { if (i >= 6) dX[i - 6] = v; });
        }
        for (int l = 2; l >= 0; --l) {
            const float *E = w + OFF_ENC + l * ENC_SIZE;
            float *GE = G + OFF_ENC + l * ENC_SIZE;
            EncStash c = enc_stash(st + L.enc + l * L.enc_layer, T);
            float *ds2 = t2, *df1 = t3, *dy1 = t4, *ds1 = t2, *dctx = t4, *dqkv = t3, *dS = t5, *df2 = t6, *dao = t7;
            const float dscale = drop.scale;
            ln_bwd(dX, c.xh2, c.rs2, E + ENC_N2G, ds2, GE + ENC_N2G, GE + ENC_N2B, T, acc);
            for (int e = tid; e < T * 50; e += nt) df2[e] = ds2[e] * drop(b, l, 3, e);      // through the FFN-branch dropout
            psync();
            Ws = pipe.take();
            pipe.issue(E + ENC_L1W, 150, 50);
            lin_bwd_w<50, 150>(GE + ENC_L2W, GE + ENC_L2B, df2, 50, c.f1, 150, T, acc);
            // f1 carries its mask and the ReLU: a kept, positive unit passes the gradient scaled by 1 / (1 - p)
            lin_bwd_x<50, 150, 150>(Ws, df2, 50, T, [=](int r, int i, float v) { df1[r * 150 + i] = c.f1[r * 150 + i] > 0.0f ? v * dscale : 0.0f; });
            Ws = pipe.take();
            pipe.issue(E + ENC_OW, 50, 50);
            lin_bwd_w<150, 50>(GE + ENC_L1W, GE + ENC_L1B, df1, 150, c.y1, 50, T, acc);
            lin_bwd_x<150, 50, 50>(Ws, df1, 150, T, [=](int r, int i, float v) { dy1[r * 50 + i] = ds2[r * 50 + i] + v; });
            ln_bwd(dy1, c.xh1, c.rs1, E + ENC_N1G, ds1, GE + ENC_N1G, GE + ENC_N1B, T, acc);
            for (int e = tid; e < T * 50; e += nt) dao[e] = ds1[e] * drop(b, l, 1, e);      // through the attention-branch dropout
            psync();
            Ws = pipe.take();
            pipe.issue(E + ENC_INW, 150, 50);
            lin_bwd_w<50, 50>(GE + ENC_OW, GE + ENC_OB, dao, 50, c.ctx, 50, T, acc);
            lin_bwd_x<50, 50, 50>(Ws, dao, 50, T, [=](int r, int i, float v) { dctx[r * 50 + i] = v; });
            // attention: dP = dctx v^T ; dS = P (dP - sum_c P dP) ; then dq = dS k / 5 ; dk = dS^T q / 5 ; dv = P^T dctx
            for (int hr = warp; hr < 2 * T; hr += nw) {
                const int hd = hr / T, r = hr - hd * T;
                const float *dc = dctx + r * 50 + 25 * hd, *p = c.P + hr * T;
                float dp[2], s = 0.0f;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int cc = lane + 32 * u;
                    dp[u] = 0.0f;
                    if (cc < T) {
                        const float *v = c.qkv + cc * 150 + 100 + 25 * hd;
                        float x = 0.0f;
#pragma unroll
                        for (int d = 0; d < 25; ++d) x = fmaf(dc[d], v[d], x);
                        x *= drop(b, l, 0, hr * T + cc);            // d/dP of ctx = (P * mask) v
                        dp[u] = x;
                        s = fmaf(p[cc], x, s);
                    }
                }
                for (int x = 16; x; x >>= 1) s += __shfl_xor_sync(0xffffffffu, s, x);
                if (lane < T) dS[hr * T + lane] = p[lane] * (dp[0] - s);
                if (lane + 32 < T) dS[hr * T + lane + 32] = p[lane + 32] * (dp[1] - s);
            }
            psync();
            for (int e = tid; e < T * 150; e += nt) {
                const int r = e / 150, j = e - r * 150, part = j / 50, jj = j - part * 50, hd = jj / 25;
                float s = 0.0f;
                if (part == 0) {            // dq[r] = sum_c dS[r][c] k[c] / 5
                    for (int cc = 0; cc < T; ++cc) s = fmaf(dS[(hd * T + r) * T + cc], c.qkv[cc * 150 + 50 + jj], s);
                    s *= 0.2f;
                } else if (part == 1) {     // dk[r] = sum_q dS[q][r] q[q] / 5
                    for (int qq = 0; qq < T; ++qq) s = fmaf(dS[(hd * T + qq) * T + r], c.qkv[qq * 150 + jj], s);
                    s *= 0.2f;
                } else {                    // dv[r] = sum_q (P * mask)[q][r] dctx[q]
                    for (int qq = 0; qq < T; ++qq)
                        s = fmaf(c.P[(hd * T + qq) * T + r] * drop(b, l, 0, (hd * T + qq) * T + r), dctx[qq * 50 + jj], s);
                }
                dqkv[e] = s;
            }
            psync();
            Ws = pipe.take();
            if (l > 0) pipe.issue(E - ENC_SIZE + ENC_L2W, 50, 150);
            else pipe.issue(w + OFF_QW, 50, 63);
            lin_bwd_w<150, 50>(GE + ENC_INW, GE + ENC_INB, dqkv, 150, c.X, 50, T, acc);
            if (l > 0) {
                lin_bwd_x<150, 50, 50>(Ws, dqkv, 150, T, [=](int r, int i, float v) { dX[r * 50 + i] = ds1[r * 50 + i] + v; });
            } else {                        // into the GRU-ACT: out = (sum_k p_k h_k) / K, so dA = dX / K; dh_K starts at 0
                lin_bwd_x<150, 50, 50>(Ws, dqkv, 150, T, [=](int r, int i, float v) {
                    dX[r * 50 + i] = (ds1[r * 50 + i] + v) * invK;
                    dh[r * 50 + i] = 0.0f;
                });
            }
        }
        const float *x = st + L.x;
        for (int k = K; k >= 1; --k) {
            const bool gacc = acc || k != K;
            const float *hp = st + L.h + (k - 1) * T * 50, *hn = st + L.h + k * T * 50;
            GruStash g = gru_stash(st + L.gru + (k - 1) * L.gru_step, T);
            float *dlog = small_ + 8;            // [T] (T <= 33)
            float *hxb = t2, *rhx = t3, *daq = t4, *dz = t5, *d1 = t6, *dr = t7;
            for (int r = warp; r < T; r += nw) {     // per row: dlogit, total dL/dh_k, the gate derivatives, [h ; x], [r h ; x]
                const float p = g.p[r];
                float s = 0.0f;
                for (int j = lane; j < 50; j += 32) s = fmaf(dX[r * 50 + j], hn[r * 50 + j], s);
                for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
                const float dl = a.fixed_steps > 0 ? 0.0f : s * p * (1.0f - p);     // d/dlogit of p_k (dA . h_k)
                if (lane == 0) dlog[r] = dl;
                for (int j = lane; j < 50; j += 32) {
                    const int e = r * 50 + j;
                    const float d = dh[e] + p * dX[e] + dl * __ldg(w + OFF_PW + j);
                    const float z = g.z[e], q = g.q[e], h = hp[e];
                    dz[e] = g.az2[e] > 0.0f ? d * (q - h) * z * (1.0f - z) : 0.0f;     // through the sigmoid and the ReLU
                    daq[e] = d * z * (1.0f - q * q);
                    dh[e] = d * (1.0f - z);                                            // -> dL/dh_{k-1}, first term
                    hxb[r * 64 + j] = h;
                    rhx[r * 64 + j] = g.r[e] * h;
                }
                if (lane < 13) {
                    hxb[r * 64 + 50 + lane] = x[r * 13 + lane];
                    rhx[r * 64 + 50 + lane] = x[r * 13 + lane];
                }
            }
            psync();
            Ws = pipe.take();
            pipe.issue(w + OFF_Z2W, 50, 100);
            for (int j = tid; j < 51; j += nt) {     // ponder_linear gradient (OFF_PB == OFF_PW + 50)
                const float old = gacc ? G[OFF_PW + j] : 0.0f;
                float s = 0.0f;
                if (j < 50) for (int r = 0; r < T; ++r) s = fmaf(dlog[r], hn[r * 50 + j], s);
                else for (int r = 0; r < T; ++r) s += dlog[r];
                G[OFF_PW + j] = old + s;
            }
            lin_bwd_w<50, 63>(G + OFF_QW, G + OFF_QB, daq, 50, rhx, 64, T, gacc);
            lin_bwd_x<50, 63, 50>(Ws, daq, 50, T, [=](int r, int i, float v) {
                const int e = r * 50 + i;
                const float rr = g.r[e];
                dh[e] = fmaf(v, rr, dh[e]);
                dr[e] = g.ar2[e] > 0.0f ? v * hp[e] * rr * (1.0f - rr) : 0.0f;
            });
            Ws = pipe.take();
            pipe.issue(w + OFF_Z1W, 100, 63);
            lin_bwd_w<50, 100>(G + OFF_Z2W, G + OFF_Z2B, dz, 50, g.az1, 100, T, gacc);
            lin_bwd_x<50, 100, 100>(Ws, dz, 50, T, [=](int r, int i, float v) { d1[r * 100 + i] = g.az1[r * 100 + i] > 0.0f ? v : 0.0f; });
            Ws = pipe.take();
            pipe.issue(w + OFF_R2W, 50, 100);
            lin_bwd_w<100, 63>(G + OFF_Z1W, G + OFF_Z1B, d1, 100, hxb, 64, T, gacc);
            lin_bwd_x<100, 63, 50>(Ws, d1, 100, T, [=](int r, int i, float v) { dh[r * 50 + i] += v; });
            Ws = pipe.take();
            pipe.issue(w + OFF_R1W, 100, 63);
            lin_bwd_w<50, 100>(G + OFF_R2W, G + OFF_R2B, dr, 50, g.ar1, 100, T, gacc);
            lin_bwd_x<50, 100, 100>(Ws, dr, 50, T, [=](int r, int i, float v) { d1[r * 100 + i] = g.ar1[r * 100 + i] > 0.0f ? v : 0.0f; });
            Ws = pipe.take();
            if (k > 1) pipe.issue(w + OFF_QW, 50, 63);
            else if (more && a.gru_g) pipe.issue(w + OFF_ENC + ENC_INW, 150, 50);
            else if (more) pipe.issue(w + OFF_Z1W, 100, 63);
            lin_bwd_w<100, 63>(G + OFF_R1W, G + OFF_R1B, d1, 100, hxb, 64, T, gacc);
            lin_bwd_x<100, 63, 50>(Ws, d1, 100, T, [=](int r, int i, float v) { dh[r * 50 + i] += v; });
        }
    }
}

// grad[e] = sum_c part[c][e] in CTA order; block 0 also writes the loss and the step count
__global__ void __launch_bounds__(256) train_reduce_kernel(const float *__restrict__ part, int G, float *__restrict__ grad,
                                                            const float *__restrict__ sqerr, int B, const int *flags,
                                                            float *loss, int32_t *steps, float *loss_sum)
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
            if (loss_sum) *loss_sum += s / (float)B;
            if (steps) *steps = flags[0] == 0 ? 1 : (flags[1] == 0 ? 2 : 3);
        }
    }
}

// torch.optim.Adam (no weight decay, no amsgrad), single-tensor formulation of torch/optim/adam.py
__global__ void __launch_bounds__(256) adam_kernel(size_t count, float *__restrict__ p, const float *__restrict__ g,
                                                    float *__restrict__ m, float *__restrict__ v, float step_size, float b1,
                                                    float b2, float eps, float bc2_sqrt, float gscale, const float *__restrict__ scal)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    if (scal) {                                      // step size and bias correction from the device-resident step count
        step_size = scal[0];
        bc2_sqrt = scal[1];
    }
    const float gr = g[e] * gscale;
    const float mm = m[e] + (gr - m[e]) * (1.0f - b1);          // exp_avg.lerp_(grad, 1 - beta1)
    const float vv = v[e] * b2 + (1.0f - b2) * gr * gr;         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[e] = mm;
    v[e] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    p[e] = p[e] - step_size * (mm / denom);
}

// torch.optim.SGD(momentum, dampening 0, no Nesterov, no weight decay): buf = g on the first step, else momentum buf + g
__global__ void __launch_bounds__(256) sgd_kernel(size_t count, float *__restrict__ p, const float *__restrict__ g,
                                                   float *__restrict__ buf, float lr, float momentum, int first, float gscale)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const float gr = g[e] * gscale;
    const float b = first ? gr : buf[e] * momentum + gr;
    buf[e] = b;
    p[e] = p[e] - lr * b;
}

// torch.optim.RMSprop(alpha, eps; not centered, no momentum): sq = alpha sq + (1 - alpha) g^2; p -= lr g / (sqrt(sq) + eps)
__global__ void __launch_bounds__(256) rmsprop_kernel(size_t count, float *__restrict__ p, const float *__restrict__ g,
                                                       float *__restrict__ sq, float lr, float alpha, float eps, float gscale)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const float gr = g[e] * gscale;
    const float s = sq[e] * alpha + (1.0f - alpha) * gr * gr;
    sq[e] = s;
    p[e] = p[e] - lr * (gr / (sqrtf(s) + eps));
}

cudaError_t launch_sgd(size_t count, float *p, const float *g, float *buf, double lr, double momentum, int step, double gscale,
                       cudaStream_t st)
{
    sgd_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(count, p, g, buf, (float)lr, (float)momentum, step == 1, (float)gscale);
    return cudaGetLastError();
}

cudaError_t launch_rmsprop(size_t count, float *p, const float *g, float *sq, double lr, double alpha, double eps, double gscale,
                           cudaStream_t st)
{
    rmsprop_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(count, p, g, sq, (float)lr, (float)alpha, (float)eps, (float)gscale);
    return cudaGetLastError();
}

// One uniformly random mini-batch WITHOUT replacement from the replay ring (what next(iter(DataLoader(memory, batch_size,
// shuffle=True))) yields, trainer.py:70-72, memory.py): Floyd's subset algorithm by warp 0 (membership test by ballot), then
// every thread gathers.  Replaces randperm(N) + two index kernels (0.11 ms at N = 100,000) with one ~5 us launch.
__device__ __forceinline__ uint64_t splitmix64(uint64_t &x)
{
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) sample_batch_kernel(int N, int k, int row_floats, const float *__restrict__ states,
                                                            const float *__restrict__ values, uint64_t seed, uint64_t counter,
                                                            float *__restrict__ out_states, float *__restrict__ out_values,
                                                            int64_t *__restrict__ out_index, unsigned long long *dev_state,
                                                            float *adam_scal, double lr, double b1, double b2)
{
    // dev_state (optional, the graph-replayable step): [0] = draw counter, [1] = optimizer step count, both advanced here;
    // adam_scal[0] = lr / (1 - b1^t), adam_scal[1] = sqrt(1 - b2^t) for the adam_kernel that follows in the stream
    extern __shared__ int picked[];                 // [k]
    if (dev_state) counter = dev_state[0];
    __syncthreads();
    if (dev_state && threadIdx.x == 32) {
        dev_state[0] = counter + 1;
        const unsigned long long t = ++dev_state[1];
        if (adam_scal) {
            adam_scal[0] = (float)(lr / (1.0 - pow(b1, (double)t)));
            adam_scal[1] = (float)sqrt(1.0 - pow(b2, (double)t));
        }
    }
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        uint64_t st = seed ^ (counter * 0xd1342543de82ef95ull + 0x2545f4914f6cdd1dull);
        for (int m = 0; m < k; ++m) {
            const int j = N - k + m;                // candidates 0 .. j
            const uint64_t r = splitmix64(st);      // the same value in every lane
            int t = (int)__umul64hi(r, (uint64_t)j + 1);
            bool hit = false;
            for (int q = lane; q < m; q += 32) hit |= picked[q] == t;
            if (__any_sync(0xffffffffu, hit)) t = j;
            if (lane == 0) picked[m] = t;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < k * row_floats; e += blockDim.x) {
        const int i = e / row_floats, c = e - i * row_floats;
        out_states[e] = states[(size_t)picked[i] * row_floats + c];
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        out_values[i] = values[picked[i]];
        if (out_index) out_index[i] = picked[i];
    }
}

cudaError_t launch_sample_batch(int N, int k, int row_floats, const float *states, const float *values, uint64_t seed,
                                uint64_t counter, float *out_states, float *out_values, int64_t *out_index, cudaStream_t st)
{
    sample_batch_kernel<<<1, 256, k * sizeof(int), st>>>(N, k, row_floats, states, values, seed, counter, out_states, out_values,
                                                         out_index, nullptr, nullptr, 0.0, 0.0, 0.0);
    return cudaGetLastError();
}

#ifdef AEM_TRAIN_TIMELINE
extern "C" int aem_debug_train_timeline(long long *out_h, int max_n)
{
    int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_tr_tl_n, sizeof(int));
    if (n > max_n) n = max_n;
    cudaMemcpyFromSymbol(out_h, g_tr_tl, sizeof(long long) * n);
    int zero = 0;
    cudaMemcpyToSymbol(g_tr_tl_n, &zero, sizeof(int));
    return n;
}
#endif

// ---- launchers ------------------------------------------------------------------------------------------------------
size_t train_smem_bytes(int T, bool stash_in_smem)
{
    const TrLayout L = tr_layout(T);
    size_t f = 2 * TR_WS_FLOATS + (size_t)TR_NBUF * T * TR_BUF_PITCH + 64 + (stash_in_smem ? L.total : 0);
    return f * sizeof(float);
}

size_t probe_smem_bytes(int T, bool handoff)
{
    const TrLayout L = tr_layout(T);     // x, h and one step block -- or all three when the stash is handed to the main kernel
    return (2 * TR_WS_FLOATS + (size_t)T * 50 + T * 64 + ((3 * T + 3) & ~3) + (handoff ? L.enc : L.gru + L.gru_step)) * sizeof(float);
}

cudaError_t launch_train_grad(aem_handle h, int B, int n, const float *params, const float *states, const float *targets,
                              float *grad, float *loss, float *values, int32_t *steps, cudaStream_t st, float *loss_sum,
                              const unsigned long long *drop_ctr_dev)
{
    const int T = n + 1;
    const TrLayout L = tr_layout(T);
    const size_t smem_limit = 227 * 1024;
    const bool in_smem = train_smem_bytes(T, true) <= smem_limit;
    const int grid = B < h->num_sms ? B : h->num_sms;
    const bool handoff = probe_smem_bytes(T, true) <= smem_limit;
    cudaError_t e;
    if (grid > h->tr_grid || B > h->tr_B || (handoff && (size_t)B * L.enc > h->tr_gru_floats) || (!in_smem && (size_t)grid * L.total > h->tr_stash_floats)) {   // (re)allocate scratch
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        cudaFree(h->tr_part); cudaFree(h->tr_sqerr); cudaFree(h->tr_flags); cudaFree(h->tr_stash); cudaFree(h->tr_gru);
        h->tr_part = h->tr_sqerr = h->tr_stash = h->tr_gru = nullptr; h->tr_flags = nullptr;
        ++h->tr_generation;
        h->tr_grid = grid > h->tr_grid ? grid : h->tr_grid;
        h->tr_B = B > h->tr_B ? B : h->tr_B;
        if ((e = cudaMalloc(&h->tr_part, (size_t)h->tr_grid * NUM_WEIGHTS * sizeof(float))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&h->tr_sqerr, (size_t)h->tr_B * sizeof(float))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&h->tr_flags, 2 * sizeof(int))) != cudaSuccess) return e;
        h->tr_stash_floats = in_smem ? 0 : (size_t)h->tr_grid * tr_layout(AEM_MAX_HUMANS + 1).total;
        if (h->tr_stash_floats && (e = cudaMalloc(&h->tr_stash, h->tr_stash_floats * sizeof(float))) != cudaSuccess) return e;
        h->tr_gru_floats = handoff ? (size_t)h->tr_B * L.enc : 0;
        if (h->tr_gru_floats && (e = cudaMalloc(&h->tr_gru, h->tr_gru_floats * sizeof(float))) != cudaSuccess) return e;
    }
    TrainArgs a;
    a.params = params; a.states = states; a.targets = targets;
    a.part = h->tr_part; a.sqerr = h->tr_sqerr; a.values = values; a.flags = h->tr_flags;
    a.stash_g = in_smem ? nullptr : h->tr_stash;
    a.gru_g = handoff ? h->tr_gru : nullptr;
    a.B = B; a.n = n;
    a.fixed_steps = h->net.act_fixed ? h->net.act_steps : 0;
    a.drop_p = (float)h->drop_p; a.drop_seed = h->drop_seed;
    a.drop_ctr_dev = drop_ctr_dev;
    a.drop_ctr = drop_ctr_dev ? 0ull : h->drop_ctr++;          // a fresh set of masks per optimisation step
    if ((e = cudaMemsetAsync(h->tr_flags, 0, 2 * sizeof(int), st)) != cudaSuccess) return e;
    const size_t psm = probe_smem_bytes(T, handoff), fsm = train_smem_bytes(T, in_smem);
    if (psm > smem_limit) return cudaErrorInvalidValue;
    if ((e = cudaFuncSetAttribute(train_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm)) != cudaSuccess) return e;
    auto kern = in_smem ? train_fwdbwd_kernel<true> : train_fwdbwd_kernel<false>;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm)) != cudaSuccess) return e;
    train_probe_kernel<<<grid, TR_THREADS, psm, st>>>(a);
    kern<<<grid, TR_THREADS, fsm, st>>>(a);
    train_reduce_kernel<<<(NUM_WEIGHTS + 255) / 256, 256, 0, st>>>(h->tr_part, grid, grad, h->tr_sqerr, B, h->tr_flags, loss, steps, loss_sum);
    return cudaGetLastError();
}

cudaError_t launch_adam(size_t count, float *p, const float *g, float *m, float *v, double lr, double b1, double b2, double eps,
                        int step, double gscale, cudaStream_t st)
{
    const double bc1 = 1.0 - pow(b1, (double)step);             // torch computes these in Python floats
    const double bc2_sqrt = sqrt(1.0 - pow(b2, (double)step));
    adam_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(count, p, g, m, v, (float)(lr / bc1), (float)b1, (float)b2,
                                                                 (float)eps, (float)bc2_sqrt, (float)gscale, nullptr);
    return cudaGetLastError();
}

// One optimisation step whose every per-step quantity lives on the device (draw counter, optimizer step count, bias corrections,
// loss accumulator): sample -> probe -> forward / backward -> reduce -> Adam, five launches and one memset with constant
// arguments, so the sequence can be captured once in a CUDA graph and replayed (Trainer.optimize_batch).
cudaError_t launch_train_step_adam(aem_handle h, int N, int k, int n, const float *ring_states, const float *ring_values, uint64_t seed,
                                   unsigned long long *dev_state, float *adam_scal, float *params, float *grad, float *m, float *v,
                                   double lr, double b1, double b2, double eps, float *batch_states, float *batch_values, float *loss,
                                   float *loss_sum, int32_t *steps, cudaStream_t st)
{
    sample_batch_kernel<<<1, 256, k * sizeof(int), st>>>(N, k, n * 13, ring_states, ring_values, seed, 0ull, batch_states, batch_values,
                                                         nullptr, dev_state, adam_scal, lr, b1, b2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if ((e = launch_train_grad(h, k, n, params, batch_states, batch_values, grad, loss, nullptr, steps, st, loss_sum, dev_state)) != cudaSuccess)
        return e;
    adam_kernel<<<(NUM_WEIGHTS + 255) / 256, 256, 0, st>>>((size_t)NUM_WEIGHTS, params, grad, m, v, 0.0f, (float)b1, (float)b2, (float)eps,
                                                          1.0f, 1.0f, adam_scal);
    return cudaGetLastError();
}

}  // namespace aem
