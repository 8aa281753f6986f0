// lookahead.cu -- fused one-step-lookahead value kernel (FP32 FFMA path), sm_100a.
//
// Replaces the 81-iteration Python loop of MultiHumanRL.predict (crowd_nav/policy/multi_human_rl.py:336-366):
//   propagate (cadrl.py:156-181) -> rotate (cadrl.py:239-274) -> ValueNetworkTransformerAndGRU.forward
//   (qnetwork_factory.py:641-687: robot pseudo-token, ATCBasic/GRUEXND components.py:38-137 with PER-SAMPLE
//   halting, 3 post-norm encoder layers, token 0, action_mlp) -> r + gamma_bar * V.
//
// Work decomposition: a persistent CTA (256 threads, one per SM) walks over tiles of S = floor(128 / T) whole
// (env, action) samples = up to 128 token rows.  All activations of a tile live in shared memory in a
// feature-major layout act[feature][row] (row pitch 128 floats), so every GEMM operand access is a
// conflict-free 128-bit shared load and every layer output a conflict-free 128-bit store.  Each dense layer is
// a register-tiled FP32 GEMM: half-warp = 16 lanes x 8 rows = all 128 rows, 16 half-warps split the (padded)
// output columns; weights stream global/L2 -> shared through a 2-stage cp.async pipeline in k-chunks.
// Accumulation order is k-ascending FMA then + bias, exactly as oracle/aem_oracle.c dense().
#include "aem_common.cuh"

namespace aem {

constexpr int RP = 128;               // rows per tile == row pitch (floats)
constexpr int NTHREADS = 256;
constexpr int BUFA_ROWS = HXD;        // [h(50) | x(13)]  -- later E(50)
constexpr int BUFACC_ROWS = HID;      // ACT accumulator -- later attention output / compact tail buffers
constexpr int BUFH_ROWS = 208;        // z1|r1 hidden (200) -- aliases Z, r*h|x, QKV, FFN hidden, tail buffers
constexpr int WSTAGE = 4096;          // floats per weight stage (16 KB), 2 stages
constexpr int CP = 32;                // compact (per-sample) row pitch of the tail

constexpr int SMEM_FLOATS = (BUFA_ROWS + BUFACC_ROWS + BUFH_ROWS) * RP + 2 * WSTAGE + 2 * RP;
constexpr int SMEM_INTS = RP + 3 * 32 + 4;
constexpr size_t SMEM_BYTES = (size_t)SMEM_FLOATS * 4 + (size_t)SMEM_INTS * 4;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------------------------------------
// Register-tiled GEMM over the 128 rows of a tile.
//   out[row][col] = sum_k A[k][row] * Wp[k][col]   (k ascending, fmaf), handed to epi(col, branch, row0, float4)
// Thread mapping: lane = (ch = lane >> 4, rl = lane & 15); group g = 2 * warp + ch owns TNH columns;
// rows {4rl..4rl+3} and {64+4rl..64+4rl+3}.
// ------------------------------------------------------------------------------------------------
template <int TNH, bool TWO, class Epi>
__device__ __forceinline__ void gemm_big(const float *__restrict__ A0, const float *__restrict__ A1,
                                         const float *__restrict__ Wg, int K, int kbeg, int ncols,
                                         float *__restrict__ wbuf, Epi epi)
{
    constexpr int TNHP = pad4(TNH);
    constexpr int NP = NGROUPS * TNHP;       // floats per k row of the packed operand
    constexpr int KC = WSTAGE / NP;          // k rows per stage
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rl = lane & 15, ch = lane >> 4, g = warp * 2 + ch;
    const float *__restrict__ A = (TWO && g >= 8) ? A1 : A0;

    float acc[8][TNH];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < TNH; ++c) acc[r][c] = 0.0f;

    const int nk = K - kbeg;
    const int nchunks = (nk + KC - 1) / KC;
    auto load_chunk = [&](int c) {
        const int k0 = kbeg + c * KC;
        const int kc = min(KC, K - k0);
        const float4 *src = reinterpret_cast<const float4 *>(Wg + (size_t)k0 * NP);
        float4 *dst = reinterpret_cast<float4 *>(wbuf + (c & 1) * WSTAGE);
        const int nvec = kc * (NP / 4);
        for (int i = tid; i < nvec; i += NTHREADS) cp_async16(dst + i, src + i);
    };
    load_chunk(0);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) load_chunk(c + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const int k0 = kbeg + c * KC;
        const int kc = min(KC, K - k0);
        const float *__restrict__ Wc = wbuf + (c & 1) * WSTAGE + g * TNHP;
        const float *__restrict__ Ak = A + (size_t)k0 * RP + 4 * rl;
#pragma unroll 2
        for (int kk = 0; kk < kc; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(Ak + kk * RP);
            const float4 a1 = *reinterpret_cast<const float4 *>(Ak + kk * RP + 64);
            float w[TNHP];
#pragma unroll
            for (int i = 0; i < TNHP / 4; ++i) {
                const float4 t = *reinterpret_cast<const float4 *>(Wc + kk * NP + 4 * i);
                w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
            }
#pragma unroll
            for (int cidx = 0; cidx < TNH; ++cidx) {
                acc[0][cidx] = fmaf(a0.x, w[cidx], acc[0][cidx]);
                acc[1][cidx] = fmaf(a0.y, w[cidx], acc[1][cidx]);
                acc[2][cidx] = fmaf(a0.z, w[cidx], acc[2][cidx]);
                acc[3][cidx] = fmaf(a0.w, w[cidx], acc[3][cidx]);
                acc[4][cidx] = fmaf(a1.x, w[cidx], acc[4][cidx]);
                acc[5][cidx] = fmaf(a1.y, w[cidx], acc[5][cidx]);
                acc[6][cidx] = fmaf(a1.z, w[cidx], acc[6][cidx]);
                acc[7][cidx] = fmaf(a1.w, w[cidx], acc[7][cidx]);
            }
        }
        __syncthreads();   // stage (c & 1) may be refilled by the next prefetch; A reads of this chunk are done
    }
    const int branch = TWO ? (g >> 3) : 0;
    const int cbase = (TWO ? (g & 7) : g) * TNH;
#pragma unroll
    for (int cidx = 0; cidx < TNH; ++cidx) {
        const int col = cbase + cidx;
        if (col < ncols) {
            epi(col, branch, 4 * rl, make_float4(acc[0][cidx], acc[1][cidx], acc[2][cidx], acc[3][cidx]));
            epi(col, branch, 64 + 4 * rl, make_float4(acc[4][cidx], acc[5][cidx], acc[6][cidx], acc[7][cidx]));
        }
    }
    __syncthreads();
}

// Same packed operand, but only 32 rows (one per lane, compact pitch CP): used for the per-sample tail
// (last encoder layer's token-0 rows and the action MLP).  Warp w owns groups 2w and 2w+1.
template <int TNH, class Epi>
__device__ __forceinline__ void gemm_small(const float *__restrict__ A, const float *__restrict__ Wg, int K,
                                           int ncols, float *__restrict__ wbuf, Epi epi)
{
    constexpr int TNHP = pad4(TNH);
    constexpr int NP = NGROUPS * TNHP;
    constexpr int KC = WSTAGE / NP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float acc[2][TNH];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < TNH; ++c) acc[j][c] = 0.0f;
    const int nchunks = (K + KC - 1) / KC;
    auto load_chunk = [&](int c) {
        const int k0 = c * KC;
        const int kc = min(KC, K - k0);
        const float4 *src = reinterpret_cast<const float4 *>(Wg + (size_t)k0 * NP);
        float4 *dst = reinterpret_cast<float4 *>(wbuf + (c & 1) * WSTAGE);
        const int nvec = kc * (NP / 4);
        for (int i = tid; i < nvec; i += NTHREADS) cp_async16(dst + i, src + i);
    };
    load_chunk(0);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) load_chunk(c + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const int k0 = c * KC;
        const int kc = min(KC, K - k0);
        const float *__restrict__ Wc = wbuf + (c & 1) * WSTAGE + (2 * warp) * TNHP;
        const float *__restrict__ Ak = A + (size_t)k0 * CP + lane;
#pragma unroll 4
        for (int kk = 0; kk < kc; ++kk) {
            const float a = Ak[kk * CP];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float w[TNHP];
#pragma unroll
                for (int i = 0; i < TNHP / 4; ++i) {
                    const float4 t = *reinterpret_cast<const float4 *>(Wc + kk * NP + j * TNHP + 4 * i);
                    w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
                }
#pragma unroll
                for (int cidx = 0; cidx < TNH; ++cidx) acc[j][cidx] = fmaf(a, w[cidx], acc[j][cidx]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int cidx = 0; cidx < TNH; ++cidx) {
            const int col = (2 * warp + j) * TNH + cidx;
            if (col < ncols) epi(col, lane, acc[j][cidx]);
        }
    __syncthreads();
}

// LayerNorm over the 50 features of one row (two-pass, eps = 1e-5), feature-major buffer with given pitch
__device__ __forceinline__ void layer_norm_row(float *buf, int pitch, int row, const float *__restrict__ gam,
                                               const float *__restrict__ bet)
{
    float mean = 0.0f;
    for (int k = 0; k < HID; ++k) mean += buf[k * pitch + row];
    mean = mean / (float)HID;
    float var = 0.0f;
    for (int k = 0; k < HID; ++k) {
        const float d = buf[k * pitch + row] - mean;
        var = fmaf(d, d, var);
    }
    var = var / (float)HID;
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    for (int k = 0; k < HID; ++k)
        buf[k * pitch + row] = (buf[k * pitch + row] - mean) * rstd * __ldg(gam + k) + __ldg(bet + k);
}

// cadrl.py:239-274 rotate (FP32, holonomic): in14 = robot 9 + human 5
__device__ __forceinline__ void rotate13(const float *s, float *o)
{
    const float dx = s[5] - s[0], dy = s[6] - s[1];
    const float rot = atan2f(dy, dx);
    const float c = cosf(rot), sn = sinf(rot);
    o[0] = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    o[1] = s[7];
    o[2] = 0.0f;
    o[3] = s[4];
    o[4] = __fadd_rn(__fmul_rn(s[2], c), __fmul_rn(s[3], sn));
    o[5] = __fsub_rn(__fmul_rn(s[3], c), __fmul_rn(s[2], sn));
    const float rx = s[9] - s[0], ry = s[10] - s[1];
    o[6] = __fadd_rn(__fmul_rn(rx, c), __fmul_rn(ry, sn));
    o[7] = __fsub_rn(__fmul_rn(ry, c), __fmul_rn(rx, sn));
    o[8] = __fadd_rn(__fmul_rn(s[11], c), __fmul_rn(s[12], sn));
    o[9] = __fsub_rn(__fmul_rn(s[12], c), __fmul_rn(s[11], sn));
    o[10] = s[13];
    const float ax = s[0] - s[9], ay = s[1] - s[10];
    o[11] = sqrtf(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)));
    o[12] = s[4] + s[13];
}

__global__ void __launch_bounds__(NTHREADS, 1) lookahead_kernel(const LookaheadArgs args)
{
    extern __shared__ __align__(16) float smem[];
    float *bufA = smem;                                   // [63][128]
    float *bufACC = bufA + BUFA_ROWS * RP;                // [50][128]
    float *bufH = bufACC + BUFACC_ROWS * RP;              // [208][128]
    float *wbuf = bufH + BUFH_ROWS * RP;                  // [2][4096]
    float *sP = wbuf + 2 * WSTAGE;                        // [128] accumulated halting probability
    float *sPv = sP + RP;                                 // [128] halting probability of this step
    int *row_sample = reinterpret_cast<int *>(sPv + RP);  // [128] sample of a row, -1 = padding
    int *s_cnt = row_sample + RP;                         // [32] ACT steps of a sample
    int *s_act = s_cnt + 32;                              // [32] sample still pondering
    int *s_need = s_act + 32;                             // [32]
    int *s_flag = s_need + 32;                            // [0] any sample active
    __shared__ int s_tact[32];                            // tokens of every sample that are real ('mixed' scenes), else T

    const NetParams &net = args.net;
    const float *__restrict__ raw = net.raw;
    const float *__restrict__ pk = net.packed;
    const int tid = threadIdx.x;
    const int T = args.T, S = args.S;
    const int ntiles = (args.total_samples + S - 1) / S;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int s0 = tile * S;
        const int nsamp = min(S, args.total_samples - s0);
        const int nrows = nsamp * T;

        // ---------------- phase 0: inputs -> x (bufA rows 50..62), h = 0, ACC = 0, P = 0 -----------------
        for (int i = tid; i < HID * RP; i += NTHREADS) { bufA[i] = 0.0f; bufACC[i] = 0.0f; }
        if (tid < RP) {
            const int r = tid;
            float o[XD];
#pragma unroll
            for (int k = 0; k < XD; ++k) o[k] = 0.0f;
            int samp = -1;
            if (r < nrows) {
                samp = r / T;
                const int t = r - samp * T;
                const int gs = s0 + samp;
                if (args.mode == 0) {
                    const int b = gs / args.A, a = gs - b * args.A;
                    const int tact = args.n_active ? 1 + min(max(args.n_active[b], 1), args.n) : T;
                    if (t == 0) s_tact[samp] = tact;
                    const double *rb = args.robot + (size_t)b * 9;
                    const double ax = args.actions[2 * a], ay = args.actions[2 * a + 1];
                    const int hi = (t == 0) ? 0 : t - 1;
                    const double *hb = args.humans_next + ((size_t)b * args.n + hi) * 5;
                    float in14[14];
                    in14[0] = (float)__dadd_rn(rb[0], __dmul_rn(ax, args.dt));   // cadrl.py:156-181 propagate (unfused), f64 -> f32
                    in14[1] = (float)__dadd_rn(rb[1], __dmul_rn(ay, args.dt));
                    in14[2] = (float)ax; in14[3] = (float)ay;
                    in14[4] = (float)rb[4]; in14[5] = (float)rb[5]; in14[6] = (float)rb[6];
                    in14[7] = (float)rb[7]; in14[8] = (float)rb[8];
#pragma unroll
                    for (int k = 0; k < 5; ++k) in14[9 + k] = (float)hb[k];
                    rotate13(in14, o);
                    if (t >= tact) {                 // padding row: not a token of this sample
#pragma unroll
                        for (int k = 0; k < XD; ++k) o[k] = 0.0f;
                        samp = -1;
                    }
                } else if (args.mode == 1) {
                    if (t == 0) s_tact[samp] = T;
                    const int hi = (t == 0) ? 0 : t - 1;
                    const float *src = args.states + ((size_t)gs * args.n + hi) * XD;
#pragma unroll
                    for (int k = 0; k < XD; ++k) o[k] = src[k];
                } else {
                    if (t == 0) s_tact[samp] = T;
                    const float *src = args.states + ((size_t)gs * T + t) * XD;
#pragma unroll
                    for (int k = 0; k < XD; ++k) o[k] = src[k];
                }
                if (args.mode != 2 && t == 0) {      // robot pseudo token, qnetwork_factory.py:651-659
                    o[6] = 0.0f; o[7] = 0.0f; o[8] = o[4]; o[9] = o[5]; o[10] = o[3]; o[11] = 0.0f; o[12] = o[3];
                }
            }
#pragma unroll
            for (int k = 0; k < XD; ++k) bufA[(HID + k) * RP + r] = o[k];
            sP[r] = 0.0f;
            row_sample[r] = samp;
        }
        if (tid < 32) { s_cnt[tid] = 0; s_act[tid] = (tid < nsamp) ? 1 : 0; s_need[tid] = 0; }
        __syncthreads();

        // ---------------- phase 1: GRU with adaptive computation time (components.py:107-137) -----------
        const int max_steps = net.act_fixed ? net.act_steps : 3;
        for (int step = 1; step <= max_steps; ++step) {
            const int kbeg = (step == 1) ? HID : 0;    // h == 0 at step 1: the first 50 products are exact zeros
            // z1 | r1 = ReLU(W1 [h;x] + b1)  -> bufH[0..200)
            gemm_big<TNH_Z1R1, true>(bufA, bufA, pk + net.off_z1r1, HXD, kbeg, 100, wbuf,
                [&](int col, int br, int row0, float4 v) {
                    const float b = __ldg(raw + (br ? OFF_R1B : OFF_Z1B) + col);
                    v.x = fmaxf(v.x + b, 0.0f); v.y = fmaxf(v.y + b, 0.0f);
                    v.z = fmaxf(v.z + b, 0.0f); v.w = fmaxf(v.w + b, 0.0f);
                    *reinterpret_cast<float4 *>(bufH + (br * 100 + col) * RP + row0) = v;
                });
            // z = sigmoid(ReLU(W2z z1 + b)), r likewise; writes Z -> bufH[0..50), r*h -> bufH[50..100)
            // (the trailing barrier of the k loop orders these writes after every read of the hidden buffer)
            gemm_big<TNH_Z2R2, true>(bufH, bufH + 100 * RP, pk + net.off_z2r2, 100, 0, HID, wbuf,
                [&](int col, int br, int row0, float4 v) {
                    const float b = __ldg(raw + (br ? OFF_R2B : OFF_Z2B) + col);
                    v.x = sigmoidf_(fmaxf(v.x + b, 0.0f)); v.y = sigmoidf_(fmaxf(v.y + b, 0.0f));
                    v.z = sigmoidf_(fmaxf(v.z + b, 0.0f)); v.w = sigmoidf_(fmaxf(v.w + b, 0.0f));
                    if (br) {
                        const float4 h = *reinterpret_cast<const float4 *>(bufA + col * RP + row0);
                        v.x = __fmul_rn(v.x, h.x); v.y = __fmul_rn(v.y, h.y);
                        v.z = __fmul_rn(v.z, h.z); v.w = __fmul_rn(v.w, h.w);
                        *reinterpret_cast<float4 *>(bufH + (HID + col) * RP + row0) = v;
                    } else {
                        *reinterpret_cast<float4 *>(bufH + col * RP + row0) = v;
                    }
                });
            // x copy behind r*h so that [r*h ; x] is one contiguous operand (bufH rows 50..112)
            for (int i = tid; i < XD * RP; i += NTHREADS) bufH[2 * HID * RP + i] = bufA[HID * RP + i];
            __syncthreads();
            // q = tanh(Wq [r*h ; x] + b);  h = (1 - z) * h + z * q   (in place in bufA rows 0..49)
            gemm_big<TNH_50, false>(bufH + HID * RP, nullptr, pk + net.off_q, HXD, kbeg, HID, wbuf,
                [&](int col, int br, int row0, float4 v) {
                    const float b = __ldg(raw + OFF_QB + col);
                    const float4 z = *reinterpret_cast<const float4 *>(bufH + col * RP + row0);
                    float4 h = *reinterpret_cast<const float4 *>(bufA + col * RP + row0);
                    h.x = __fadd_rn(__fmul_rn(1.0f - z.x, h.x), __fmul_rn(z.x, tanhf(v.x + b)));
                    h.y = __fadd_rn(__fmul_rn(1.0f - z.y, h.y), __fmul_rn(z.y, tanhf(v.y + b)));
                    h.z = __fadd_rn(__fmul_rn(1.0f - z.z, h.z), __fmul_rn(z.z, tanhf(v.z + b)));
                    h.w = __fadd_rn(__fmul_rn(1.0f - z.w, h.w), __fmul_rn(z.w, tanhf(v.w + b)));
                    *reinterpret_cast<float4 *>(bufA + col * RP + row0) = h;
                });
            if (net.act_fixed) {
                if (tid < 32 && tid < nsamp) s_cnt[tid] = step;
                continue;
            }
            // halting probability p = sigmoid(w_p . h + b_p) per row; P += p
            if (tid < RP) {
                const int r = tid;
                const int sp = row_sample[r];
                if (sp >= 0 && s_act[sp]) {
                    float a = 0.0f;
                    for (int k = 0; k < HID; ++k) a = fmaf(bufA[k * RP + r], __ldg(raw + OFF_PW + k), a);
                    const float p = sigmoidf_(a + __ldg(raw + OFF_PB));
                    const float Pn = sP[r] + p;
                    sP[r] = Pn;
                    sPv[r] = p;
                    if (Pn < 0.95f) s_need[sp] = 1;      // accum_p < 1 - epsilon, per (env, action) sample
                }
            }
            if (tid == 0) s_flag[0] = 0;
            __syncthreads();
            // accum_hx += p * h  (only samples that are still pondering)
            for (int i = tid; i < HID * RP; i += NTHREADS) {
                const int r = i & (RP - 1);
                const int sp = row_sample[r];
                if (sp >= 0 && s_act[sp]) bufACC[i] = __fadd_rn(bufACC[i], __fmul_rn(sPv[r], bufA[i]));
            }
            __syncthreads();
            if (tid < 32 && s_act[tid]) {
                s_cnt[tid] = step;
                const int need = s_need[tid];
                s_act[tid] = need;
                s_need[tid] = 0;
                if (need) s_flag[0] = 1;
            }
            __syncthreads();
            if (!s_flag[0]) break;
        }
        __syncthreads();
        // in_mlp output: accum_hx / step_count (adaptive) or the last h (act_fixed) -> E = bufA rows 0..49
        if (!net.act_fixed) {
            for (int i = tid; i < HID * RP; i += NTHREADS) {
                const int sp = row_sample[i & (RP - 1)];
                bufA[i] = (sp >= 0) ? bufACC[i] / (float)s_cnt[sp] : 0.0f;
            }
        }
        __syncthreads();

        if (args.mode == 2) {
            for (int i = tid; i < HID * RP; i += NTHREADS) {
                const int r = i & (RP - 1), k = i >> 7;
                if (r < nrows) args.gru_out[((size_t)s0 * T + r) * HID + k] = bufA[i];
            }
            if (args.steps && tid < nsamp) args.steps[s0 + tid] = (uint8_t)s_cnt[tid];
            __syncthreads();
            continue;
        }

        // ---------------- phase 2: 3 x post-norm TransformerEncoderLayer(50, 2, 150) ---------------------
        for (int l = 0; l < 3; ++l) {
            const float *__restrict__ E = raw + OFF_ENC + l * ENC_SIZE;
            // in_proj: QKV = W_in e + b_in -> bufH rows 0..149
            gemm_big<TNH_150, false>(bufA, nullptr, pk + net.off_in[l], HID, 0, 150, wbuf,
                [&](int col, int br, int row0, float4 v) {
                    const float b = __ldg(E + ENC_INB + col);
                    v.x += b; v.y += b; v.z += b; v.w += b;
                    *reinterpret_cast<float4 *>(bufH + col * RP + row0) = v;
                });
            if (l < 2) {
                // attention, one thread per (row, head): softmax_j(q_i . k_j / 5) v_j -> bufACC rows 0..49
                {
                    const int r = tid & (RP - 1), hd = tid >> 7;
                    const int sp = row_sample[r];
                    if (sp >= 0) {
                        const int rb = sp * T;
                        const int Tk = s_tact[sp];               // keys = the sample's real tokens
                        const float *qp = bufH + (hd * 25) * RP + r;
                        float m = -INFINITY;
                        for (int j = 0; j < Tk; ++j) {
                            const float *kp = bufH + (HID + hd * 25) * RP + rb + j;
                            float d = 0.0f;
#pragma unroll 5
                            for (int k = 0; k < 25; ++k) d = fmaf(qp[k * RP], kp[k * RP], d);
                            m = fmaxf(m, d * 0.2f);
                        }
                        float o[25];
#pragma unroll
                        for (int k = 0; k < 25; ++k) o[k] = 0.0f;
                        float sum = 0.0f;
                        for (int j = 0; j < Tk; ++j) {
                            const float *kp = bufH + (HID + hd * 25) * RP + rb + j;
                            const float *vp = bufH + (2 * HID + hd * 25) * RP + rb + j;
                            float d = 0.0f;
#pragma unroll 5
                            for (int k = 0; k < 25; ++k) d = fmaf(qp[k * RP], kp[k * RP], d);
                            const float p = expf(d * 0.2f - m);
                            sum += p;
#pragma unroll
                            for (int k = 0; k < 25; ++k) o[k] = fmaf(p, vp[k * RP], o[k]);
                        }
                        const float inv = 1.0f / sum;
#pragma unroll
                        for (int k = 0; k < 25; ++k) bufACC[(hd * 25 + k) * RP + r] = o[k] * inv;
                    } else {
#pragma unroll
                        for (int k = 0; k < 25; ++k) bufACC[(hd * 25 + k) * RP + r] = 0.0f;
                    }
                }
                __syncthreads();
                // out_proj + residual
                gemm_big<TNH_50, false>(bufACC, nullptr, pk + net.off_out[l], HID, 0, HID, wbuf,
                    [&](int col, int br, int row0, float4 v) {
                        const float b = __ldg(E + ENC_OB + col);
                        float4 e = *reinterpret_cast<const float4 *>(bufA + col * RP + row0);
                        e.x += v.x + b; e.y += v.y + b; e.z += v.z + b; e.w += v.w + b;
                        *reinterpret_cast<float4 *>(bufA + col * RP + row0) = e;
                    });
                if (tid < RP) layer_norm_row(bufA, RP, tid, E + ENC_N1G, E + ENC_N1B);
                __syncthreads();
                // FFN
                gemm_big<TNH_150, false>(bufA, nullptr, pk + net.off_f1[l], HID, 0, 150, wbuf,
                    [&](int col, int br, int row0, float4 v) {
                        const float b = __ldg(E + ENC_L1B + col);
                        v.x = fmaxf(v.x + b, 0.0f); v.y = fmaxf(v.y + b, 0.0f);
                        v.z = fmaxf(v.z + b, 0.0f); v.w = fmaxf(v.w + b, 0.0f);
                        *reinterpret_cast<float4 *>(bufH + col * RP + row0) = v;
                    });
                gemm_big<TNH_50, false>(bufH, nullptr, pk + net.off_f2[l], 150, 0, HID, wbuf,
                    [&](int col, int br, int row0, float4 v) {
                        const float b = __ldg(E + ENC_L2B + col);
                        float4 e = *reinterpret_cast<const float4 *>(bufA + col * RP + row0);
                        e.x += v.x + b; e.y += v.y + b; e.z += v.z + b; e.w += v.w + b;
                        *reinterpret_cast<float4 *>(bufA + col * RP + row0) = e;
                    });
                if (tid < RP) layer_norm_row(bufA, RP, tid, E + ENC_N2G, E + ENC_N2B);
                __syncthreads();
            } else {
                // ---- last layer: only token 0 of every sample is consumed downstream (qnetwork_factory.py:676-677)
                float *AOc = bufACC;                 // [50][32] attention output of token 0
                float *Ec = bufACC + HID * CP;       // [50][32] residual stream of token 0
                float *Hc = bufH + 150 * RP;         // [150][32] FFN hidden         (QKV stays in rows 0..149)
                if (tid < 2 * nsamp) {
                    const int sp = tid >> 1, hd = tid & 1;
                    const int rb = sp * T;
                    const int Tk = s_tact[sp];
                    const float *qp = bufH + (hd * 25) * RP + rb;
                    float m = -INFINITY;
                    for (int j = 0; j < Tk; ++j) {
                        const float *kp = bufH + (HID + hd * 25) * RP + rb + j;
                        float d = 0.0f;
                        for (int k = 0; k < 25; ++k) d = fmaf(qp[k * RP], kp[k * RP], d);
                        m = fmaxf(m, d * 0.2f);
                    }
                    float o[25];
#pragma unroll
                    for (int k = 0; k < 25; ++k) o[k] = 0.0f;
                    float sum = 0.0f;
                    for (int j = 0; j < Tk; ++j) {
                        const float *kp = bufH + (HID + hd * 25) * RP + rb + j;
                        const float *vp = bufH + (2 * HID + hd * 25) * RP + rb + j;
                        float d = 0.0f;
                        for (int k = 0; k < 25; ++k) d = fmaf(qp[k * RP], kp[k * RP], d);
                        const float p = expf(d * 0.2f - m);
                        sum += p;
#pragma unroll
                        for (int k = 0; k < 25; ++k) o[k] = fmaf(p, vp[k * RP], o[k]);
                    }
                    const float inv = 1.0f / sum;
#pragma unroll
                    for (int k = 0; k < 25; ++k) AOc[(hd * 25 + k) * CP + sp] = o[k] * inv;
                }
                for (int i = tid; i < HID * CP; i += NTHREADS) {
                    const int sp = i & (CP - 1), k = i >> 5;
                    Ec[i] = (sp < nsamp) ? bufA[k * RP + sp * T] : 0.0f;
                    if (sp >= nsamp) AOc[i] = 0.0f;
                }
                __syncthreads();
                gemm_small<TNH_50>(AOc, pk + net.off_out[l], HID, HID, wbuf, [&](int col, int row, float v) {
                    Ec[col * CP + row] += v + __ldg(E + ENC_OB + col);
                });
                if (tid < CP) layer_norm_row(Ec, CP, tid, E + ENC_N1G, E + ENC_N1B);
                __syncthreads();
                gemm_small<TNH_150>(Ec, pk + net.off_f1[l], HID, 150, wbuf, [&](int col, int row, float v) {
                    Hc[col * CP + row] = fmaxf(v + __ldg(E + ENC_L1B + col), 0.0f);
                });
                gemm_small<TNH_50>(Hc, pk + net.off_f2[l], 150, HID, wbuf, [&](int col, int row, float v) {
                    Ec[col * CP + row] += v + __ldg(E + ENC_L2B + col);
                });
                if (tid < CP) layer_norm_row(Ec, CP, tid, E + ENC_N2G, E + ENC_N2B);
                __syncthreads();
            }
        }

        // ---------------- phase 3: action_mlp([self_state(6) ; env(50)]) 56 -> 100 -> 50 -> 1 ----------------
        {
            float *Ec = bufACC + HID * CP;
            float *J = bufH;                    // [56][32]
            float *A1c = bufH + 56 * CP;        // [100][32]
            float *A2c = A1c + 100 * CP;        // [50][32]
            for (int i = tid; i < 56 * CP; i += NTHREADS) {
                const int sp = i & (CP - 1), k = i >> 5;
                float v = 0.0f;
                if (sp < nsamp) v = (k < 6) ? bufA[(HID + k) * RP + sp * T] : Ec[(k - 6) * CP + sp];
                J[i] = v;
            }
            __syncthreads();
            gemm_small<TNH_100>(J, pk + net.off_a0, 56, 100, wbuf, [&](int col, int row, float v) {
                A1c[col * CP + row] = fmaxf(v + __ldg(raw + OFF_A0B + col), 0.0f);
            });
            gemm_small<TNH_50>(A1c, pk + net.off_a2, 100, HID, wbuf, [&](int col, int row, float v) {
                A2c[col * CP + row] = fmaxf(v + __ldg(raw + OFF_A2B + col), 0.0f);
            });
            if (tid < nsamp) {
                float a = 0.0f;
                for (int k = 0; k < HID; ++k) a = fmaf(A2c[k * CP + tid], __ldg(raw + OFF_A4W + k), a);
                const float v = a + __ldg(raw + OFF_A4B);
                const int gs = s0 + tid;
                if (args.net_values) args.net_values[gs] = v;
                if (args.values) args.values[gs] = __dadd_rn(args.rewards[gs], __dmul_rn(args.gamma_bar, (double)v));
                if (args.steps) args.steps[gs] = (uint8_t)s_cnt[tid];
            }
            __syncthreads();
        }
    }
}

// multi_human_rl.py:321-322 (goal short-circuit) + :368-372 (first strict maximum, NaN never wins)
__global__ void select_kernel(const double *__restrict__ values, const double *__restrict__ robot, int B, int A,
                              int32_t *__restrict__ best)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const double *v = values + (size_t)warp * A;
    double bv = -INFINITY;
    int bi = -1;
    for (int a = lane; a < A; a += 32) {
        const double x = v[a];
        if (x > bv) { bv = x; bi = a; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
        const double *rb = robot + (size_t)warp * 9;
        const double dx = rb[0] - rb[5], dy = rb[1] - rb[6];
        if (sqrt(dy * dy + dx * dx) < rb[4]) bi = 0;     // Policy.reach_destination -> ActionXY(0, 0)
        best[warp] = bi;
    }
}

// MultiHumanRL.transform (multi_human_rl.py:423-438): rotate() of the CURRENT joint state -> the replay-memory sample
// [n][13] of every environment.  One thread per (env, human); same rotate13 as the lookahead kernel.
__global__ void transform_kernel(const double *__restrict__ robot, const double *__restrict__ humans, int B, int n,
                                 float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * n) return;
    const int b = i / n;
    const double *rb = robot + (size_t)b * 9;
    const double *hb = humans + (size_t)i * 8;
    float in14[14], o[XD];
#pragma unroll
    for (int k = 0; k < 9; ++k) in14[k] = (float)rb[k];
#pragma unroll
    for (int k = 0; k < 5; ++k) in14[9 + k] = (float)hb[k];
    rotate13(in14, o);
    float *dst = out + (size_t)i * XD;
#pragma unroll
    for (int k = 0; k < XD; ++k) dst[k] = o[k];
}

cudaError_t launch_transform(const double *robot, const double *humans, int B, int n, float *out, cudaStream_t st)
{
    if (B <= 0) return cudaSuccess;
    const int threads = 128;
    transform_kernel<<<(B * n + threads - 1) / threads, threads, 0, st>>>(robot, humans, B, n, out);
    return cudaGetLastError();
}

cudaError_t launch_lookahead(aem_handle h, const LookaheadArgs &a, cudaStream_t st)
{
    static bool attr_set[64] = {false};
    if (!attr_set[h->device & 63]) {
        cudaError_t e = cudaFuncSetAttribute(lookahead_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set[h->device & 63] = true;
    }
    const int ntiles = (a.total_samples + a.S - 1) / a.S;
    if (ntiles <= 0) return cudaSuccess;
    const int grid = ntiles < h->num_sms ? ntiles : h->num_sms;
    lookahead_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_select(const double *values, const double *robot, int B, int A, int32_t *best, cudaStream_t st)
{
    if (B <= 0) return cudaSuccess;
    const int threads = 256;
    const int blocks = (B * 32 + threads - 1) / threads;
    select_kernel<<<blocks, threads, 0, st>>>(values, robot, B, A, best);
    return cudaGetLastError();
}

}  // namespace aem
