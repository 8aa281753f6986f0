// lookahead_tc.cu -- tensor-core variant of the fused lookahead kernel (tcgen05 / TMEM / bulk-TMA, sm_100a).
//
// Same function as lookahead.cu (multi_human_rl.py:336-366 -> qnetwork_factory.py:641-687) but every dense layer is a
// tcgen05.mma.kind::tf32 with M = 128 token rows:
//   * 3xTF32 split (x = hi + lo; D = A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, FP32 accumulate in TMEM) keeps the result
//     within ~1e-6 relative of the FP32 path, so the north-star tolerance (values <= 1e-4 relative, argmax identical
//     except near-ties) still holds; tolerance is stated per kernel in tests/test_gpu_parity.py;
//   * activations never touch shared memory: the row-owner thread reads its accumulator row with tcgen05.ld, applies
//     bias / activation / LayerNorm in registers and writes the next layer's A operand (hi and lo parts) straight
//     back into TENSOR MEMORY with tcgen05.st (A-from-TMEM form of tcgen05.mma);
//   * weights are pre-tiled on the host into the canonical K-major no-swizzle image (img[k/4][n][4], hi and lo) and
//     stream L2 -> shared memory with one cp.async.bulk per image through a 5-deep mbarrier ring;
//   * one extra warp (warp 8) issues the bulk copies and the MMAs; warps 0..7 are the epilogue: two threads per token
//     row (TMEM lane), each owning one half of the columns (= one attention head).
// K / N index orders of every layer are permuted on the host (capi.cu pack_tc) so that each thread's columns are
// contiguous TMEM columns.  Per-(env, action)-sample ACT halting as in lookahead.cu.
#include "aem_common.cuh"
#include "tc_common.cuh"

namespace aem {
namespace tc {

using namespace aemtc;

constexpr int NT = 256;                 // 8 warps: two threads per token row; thread 0 also issues TMA + MMA
constexpr int NBUF = 5;                 // weight ring
constexpr int WBUF_BYTES = 29696;       // >= largest image (Z1: 112 x 64 x 4 = 28672)
constexpr int KV_PITCH = 101;
constexpr int NEV = 4;                  // MMA-commit mbarriers
constexpr int PBLK_MAX = TC_PBLK_FLOATS;

// ---- tensor-memory column map (512 columns x 128 lanes) ----
constexpr uint32_t A1H = 0, A1L = 64, D1 = 128, A2H = 240, A2L = 344, D2 = 448, A3H = 240, A3L = 304;   // GRU
constexpr uint32_t AEH = 0, AEL = 56, DIN = 128, AATH = 320, AATL = 376, DOUT = 432;                     // encoder
constexpr uint32_t DF1A = 112, DF1B = 192, AHAH = 272, AHAL = 352, AHBH = 0, AHBL = 80, DF2 = 432;
constexpr uint32_t DA0 = 128, AJH = 272, AJL = 376, DA2 = 0;                                             // action MLP

constexpr size_t SMEM_BYTES = (size_t)NBUF * WBUF_BYTES + 128 * KV_PITCH * 4 + PBLK_MAX * 4 + 4 * 128 * 4 +
                              (128 + 96 + 4) * 4 + (NBUF + NEV) * 8 + 16;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void ld8f(uint32_t taddr, float *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(taddr)
                 : "memory");
}
// store 8 fp32 values as their (hi, lo) TF32 split into two column ranges
__device__ __forceinline__ void st8_split(uint32_t hi_addr, uint32_t lo_addr, const float *v)
{
    uint32_t h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_tf32(v[i], h[i], l[i]);
    tmem_st8(hi_addr, h);
    tmem_st8(lo_addr, l);
}
__device__ __forceinline__ void st4_split(uint32_t hi_addr, uint32_t lo_addr, const float *v)
{
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(v[i], h[i], l[i]);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hi_addr), "r"(h[0]), "r"(h[1]),
                 "r"(h[2]), "r"(h[3])
                 : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_addr), "r"(l[0]), "r"(l[1]),
                 "r"(l[2]), "r"(l[3])
                 : "memory");
}
// NV values (multiple of 4, compile time) -> hi/lo column ranges
template <int NV>
__device__ __forceinline__ void st_split(uint32_t hi_addr, uint32_t lo_addr, const float *v)
{
#pragma unroll
    for (int i = 0; i + 8 <= NV; i += 8) st8_split(hi_addr + i, lo_addr + i, v + i);
    if (NV % 8) st4_split(hi_addr + (NV & ~7), lo_addr + (NV & ~7), v + (NV & ~7));
}
template <int NV>
__device__ __forceinline__ void ld_cols(uint32_t addr, float *v)
{
#pragma unroll
    for (int i = 0; i < NV; i += 8) ld8f(addr + i, v + i);
    tmem_wait_ld();
}

// streaming epilogue: out[c] = ReLU(D[c] + bias[c]) for NV columns (multiple of 4), written as the hi / lo parts of the
// next A operand.  A rolled loop over 8-column chunks keeps the instruction footprint small.
template <int NV>
__device__ __forceinline__ void relu_stream(uint32_t d_addr, const float *__restrict__ bias, uint32_t hi_addr,
                                            uint32_t lo_addr)
{
#pragma unroll 1
    for (int c = 0; c + 8 <= NV; c += 8) {
        float v[8];
        ld8f(d_addr + c, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j] + bias[c + j], 0.0f);
        st8_split(hi_addr + c, lo_addr + c, v);
    }
    if (NV % 8) {
        constexpr int c = NV & ~7;
        float v[8];
        ld8f(d_addr + c, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j] + bias[c + j], 0.0f);
        st4_split(hi_addr + c, lo_addr + c, v);
    }
}

// fast transcendental forms (MUFU ex2 / rcp, ~2 ulp): the tensor-core path is specified to 1e-4 relative anyway
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

__device__ __forceinline__ void rotate13(const float *s, float *o)   // cadrl.py:239-274, see lookahead.cu
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

// ------------------------------------------------------------------------------------------------
// issuer state (lane 0 of warp 8 only)
// ------------------------------------------------------------------------------------------------
struct Issuer {
    uint32_t tm;
    uint8_t *wbuf;
    uint64_t *wfull;
    const float *img;
    const TcLayer *layers;
    int nloaded, nused;          // chunk counters (2 chunks per layer: hi image, lo image)
    int nfree;                   // chunks whose MMAs have completed (buffers reusable)
    int pf_stage, pf_idx, pf_part;   // prefetch cursor; stage -1 = waiting for the ACT decision, -2 = no more work
    int tiles_left;              // tiles after the one the prefetch cursor is in

    __device__ __forceinline__ int stage_len(int st) const { return st == 0 ? 3 : (st == 1 ? 5 : 23); }
    __device__ __forceinline__ int layer_at(int st, int idx) const
    {
        if (st == 0) return idx == 0 ? TL_Z1S : (idx == 1 ? TL_Z2 : TL_QS);
        if (st == 1) return idx == 0 ? TL_Z1 : (idx == 1 ? TL_Z2 : (idx == 2 ? TL_R1 : (idx == 3 ? TL_R2 : TL_Q)));
        return idx < 21 ? TL_ENC + idx : (idx == 21 ? TL_A0 : TL_A2);
    }
    __device__ __noinline__ void prefetch()
    {
        while (pf_stage >= 0 && nloaded - nfree < NBUF) {
            const TcLayer &L = layers[layer_at(pf_stage, pf_idx)];
            const int buf = nloaded % NBUF;
            mbar_expect_tx(&wfull[buf], (uint32_t)L.bytes);
            bulk_g2s(wbuf + (size_t)buf * WBUF_BYTES, img + (pf_part ? L.off_lo : L.off_hi), (uint32_t)L.bytes,
                     &wfull[buf]);
            ++nloaded;
            if (++pf_part == 2) {
                pf_part = 0;
                if (++pf_idx == stage_len(pf_stage)) {
                    pf_idx = 0;
                    if (pf_stage == 2) {
                        if (tiles_left > 0) { --tiles_left; pf_stage = 0; } else pf_stage = -2;
                    } else {
                        pf_stage = -1;      // next stage depends on the halting decision of this GRU step
                    }
                }
            }
        }
    }
    __device__ __forceinline__ void decided(int next_stage)
    {
        if (pf_stage == -1) { pf_stage = next_stage; pf_idx = 0; pf_part = 0; }
    }
    __device__ __noinline__ const float *take_chunk()
    {
        while (nloaded <= nused) prefetch();
        const int buf = nused % NBUF;
        mbar_wait(&wfull[buf], (uint32_t)((nused / NBUF) & 1));
        ++nused;
        return reinterpret_cast<const float *>(wbuf + (size_t)buf * WBUF_BYTES);
    }
    // D[d] (+)= A * W^T for one layer, 3 passes.  special: the step-1 images consume A columns {24..31, 56..63}.
    __device__ __noinline__ void issue(int layer, uint32_t a_hi, uint32_t a_lo, uint32_t d, bool special,
                                          bool accumulate)
    {
        const TcLayer &L = layers[layer];
        const uint32_t idesc = make_idesc_tf32(128, L.N);
        const uint32_t lbo = (uint32_t)L.N * 16;
        uint32_t acc = accumulate ? 1u : 0u;
        const float *w = take_chunk();
        const uint32_t wa = smem_u32(w);
        for (int ks = 0; ks < L.ksteps; ++ks) {
            const uint32_t acol = special ? (ks == 0 ? 24u : 56u) : (uint32_t)(8 * ks);
            const uint64_t bd = make_smem_desc(wa + ks * 2 * lbo, lbo, 128);
            mma_tf32_ts(tm + d, tm + a_hi + acol, bd, idesc, acc);
            acc = 1u;
            mma_tf32_ts(tm + d, tm + a_lo + acol, bd, idesc, acc);
        }
        const float *wl = take_chunk();
        const uint32_t wla = smem_u32(wl);
        for (int ks = 0; ks < L.ksteps; ++ks) {
            const uint32_t acol = special ? (ks == 0 ? 24u : 56u) : (uint32_t)(8 * ks);
            const uint64_t bd = make_smem_desc(wla + ks * 2 * lbo, lbo, 128);
            mma_tf32_ts(tm + d, tm + a_hi + acol, bd, idesc, 1u);
        }
    }
};

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __noinline__ uint32_t take_chunk_addr(Issuer *iss)
{
    return smem_u32(iss->take_chunk());
}

// Warp-uniform MMA issue of one layer (all 32 lanes of warp 0 execute it, one elected lane issues): N and the number
// of k-steps are compile-time, so every descriptor is base + constant and the loop is fully unrolled -- the single-thread
// issue rate, not the tensor pipe, was the bottleneck of the first version (profiles/r01_lookahead_tc_ncu.md).
template <int N, int KS>
__device__ __forceinline__ void issue_t(Issuer *iss, int lane, uint32_t tm, uint32_t a_hi, uint32_t a_lo, uint32_t d,
                                        bool special, bool accumulate)
{
    constexpr uint32_t LBO = (uint32_t)N * 16;
    constexpr uint64_t DHI = ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t wa = 0;
    if (lane == 0) wa = take_chunk_addr(iss);
    wa = __shfl_sync(0xffffffffu, wa, 0);
    uint32_t base = (wa & 0x3FFFFu) >> 4;
    const bool leader = elect_one();
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const uint32_t acol = special ? (ks == 0 ? 24u : 56u) : (uint32_t)(8 * ks);
        const uint64_t bd = DHI | (uint64_t)(base + ks * ((2 * LBO) >> 4));
        if (leader) {
            mma_tf32_ts(tm + d, tm + a_hi + acol, bd, IDESC, (ks == 0 && !accumulate) ? 0u : 1u);
            mma_tf32_ts(tm + d, tm + a_lo + acol, bd, IDESC, 1u);
        }
    }
    wa = 0;
    if (lane == 0) wa = take_chunk_addr(iss);
    wa = __shfl_sync(0xffffffffu, wa, 0);
    base = (wa & 0x3FFFFu) >> 4;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const uint32_t acol = special ? (ks == 0 ? 24u : 56u) : (uint32_t)(8 * ks);
        const uint64_t bd = DHI | (uint64_t)(base + ks * ((2 * LBO) >> 4));
        if (leader) mma_tf32_ts(tm + d, tm + a_hi + acol, bd, IDESC, 1u);
    }
}

__global__ void __launch_bounds__(NT, 1) lookahead_tc_kernel(const LookaheadArgs args)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *wbuf = smem_raw;
    float *kv = reinterpret_cast<float *>(wbuf + (size_t)NBUF * WBUF_BYTES);   // [128][101]: k(25) v(25) per head
    float *pblk = kv + 128 * KV_PITCH;
    float *sx = pblk + PBLK_MAX;                                                // [4][128] exchange
    int *row_sample = reinterpret_cast<int *>(sx + 4 * 128);
    int *s_cnt = row_sample + 128;
    int *s_act = s_cnt + 32;
    int *s_need = s_act + 32;
    int *s_flag = s_need + 32;
    uint64_t *wfull = reinterpret_cast<uint64_t *>(s_flag + 4);
    uint64_t *evb = wfull + NBUF;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(evb + NEV);

    const NetParams &net = args.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr bool is_issuer_warp = false;   // no dedicated warp: thread 0 issues while the others wait anyway
    const bool is_issuer = (tid == 0);
    const int half = (warp >> 2) & 1, q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const int T = args.T, S = args.S;
    const int ntiles = (args.total_samples + S - 1) / S;

    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) mbar_init(&wfull[i], 1);
        for (int i = 0; i < NEV; ++i) mbar_init(&evb[i], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    for (int i = tid; i < net.tc_pblk_floats; i += NT) pblk[i] = net.tc_pblk[i];
    __shared__ TcLayer s_layers[TL_NUM];
    if (tid < TL_NUM) s_layers[tid] = net.tc_layers_d[tid];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_slot;
    const uint32_t tl = tm + lane_base;     // this thread's lane quarter

    // issuer state lives in shared memory (thread 0 only): local memory would be an L2 round trip per access here,
    // because 216 KB of the 228 KB L1/shared array are carved out as shared memory
    __shared__ Issuer s_iss;
    Issuer &iss = s_iss;
    Issuer *iss_p = &s_iss;
    if (tid == 0) {
        iss.tm = tm; iss.wbuf = wbuf; iss.wfull = wfull; iss.img = net.tc_img; iss.layers = s_layers;
        iss.nloaded = iss.nused = iss.nfree = 0;
        iss.pf_stage = -2; iss.pf_idx = 0; iss.pf_part = 0; iss.tiles_left = 0;
        if ((int)blockIdx.x < ntiles) {
            iss.pf_stage = 0;
            iss.tiles_left = (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x;
        }
    }
    __syncthreads();
    int nev = 0;     // commit events, counted identically by every thread

    // group barrier: every A operand written / every accumulator read before the issuer may launch the next MMAs
#define GROUP_BARRIER()                                   \
    do {                                                  \
        tmem_wait_st();                                   \
        tc_fence_before();                                \
        __syncthreads();                                  \
        if (warp == 0) tc_fence_after();                  \
    } while (0)
#define COMMIT(e) mma_commit(&evb[(e) % NEV])
#define WAIT_EVENT(e)                                                   \
    do {                                                                \
        mbar_wait(&evb[(e) % NEV], (uint32_t)(((e) / NEV) & 1));        \
        tc_fence_after();                                               \
    } while (0)
    // issuer: after its group is complete every chunk it used is reusable
#define ISSUER_DRAIN(e)                                                 \
    do {                                                                \
        iss.prefetch();                                                 \
        mbar_wait(&evb[(e) % NEV], (uint32_t)(((e) / NEV) & 1));        \
        iss.nfree = iss.nused;                                          \
        iss.prefetch();                                                 \
    } while (0)

    const float *PB = pblk;
    const TcLayer *LY = s_layers;

#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int s0 = tile * S;
        const int nsamp = min(S, args.total_samples - s0);
        const int nrows = nsamp * T;

        // per-thread row state
        float xh[7], selfs[3], ACC[25], hn[25], z[25];
        float P = 0.0f;
        int samp = -1, tok = 0;
#pragma unroll
        for (int j = 0; j < 25; ++j) { ACC[j] = 0.0f; hn[j] = 0.0f; }

        // ---------------- phase 0: inputs -> A1 = [h = 0 ; x] ----------------
        if (!is_issuer_warp) {
            float o[XD];
#pragma unroll
            for (int k = 0; k < XD; ++k) o[k] = 0.0f;
            if (row < nrows) {
                samp = row / T;
                tok = row - samp * T;
                const int gs = s0 + samp;
                if (args.mode == 0) {
                    const int b = gs / args.A, a = gs - b * args.A;
                    const double *rb = args.robot + (size_t)b * 9;
                    const double ax = args.actions[2 * a], ay = args.actions[2 * a + 1];
                    const int hi = (tok == 0) ? 0 : tok - 1;
                    const double *hb = args.humans_next + ((size_t)b * args.n + hi) * 5;
                    float in14[14];
                    in14[0] = (float)__dadd_rn(rb[0], __dmul_rn(ax, args.dt));
                    in14[1] = (float)__dadd_rn(rb[1], __dmul_rn(ay, args.dt));
                    in14[2] = (float)ax; in14[3] = (float)ay;
                    in14[4] = (float)rb[4]; in14[5] = (float)rb[5]; in14[6] = (float)rb[6];
                    in14[7] = (float)rb[7]; in14[8] = (float)rb[8];
#pragma unroll
                    for (int k = 0; k < 5; ++k) in14[9 + k] = (float)hb[k];
                    rotate13(in14, o);
                } else if (args.mode == 1) {
                    const int hi = (tok == 0) ? 0 : tok - 1;
                    const float *src = args.states + ((size_t)gs * args.n + hi) * XD;
#pragma unroll
                    for (int k = 0; k < XD; ++k) o[k] = src[k];
                } else {
                    const float *src = args.states + ((size_t)gs * T + tok) * XD;
#pragma unroll
                    for (int k = 0; k < XD; ++k) o[k] = src[k];
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) selfs[i] = half ? o[3 + i] : o[i];   // self_state = first 6 features of human row 0
                if (args.mode != 2 && tok == 0) {
                    o[6] = 0.0f; o[7] = 0.0f; o[8] = o[4]; o[9] = o[5]; o[10] = o[3]; o[11] = 0.0f; o[12] = o[3];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 3; ++i) selfs[i] = 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 7; ++i) xh[i] = half ? (i < 6 ? o[7 + (i < 6 ? i : 0)] : 0.0f) : o[i];
            {
                float v[32];
#pragma unroll
                for (int j = 0; j < 25; ++j) v[j] = 0.0f;
#pragma unroll
                for (int i = 0; i < 7; ++i) v[25 + i] = xh[i];
                st_split<32>(tl + A1H + 32 * half, tl + A1L + 32 * half, v);
            }
            if (half == 0) row_sample[row] = samp;
            if (tid < 32) { s_cnt[tid] = 0; s_act[tid] = (tid < nsamp) ? 1 : 0; s_need[tid] = 0; }
        }

        // ---------------- phase 1: GRU + adaptive computation time ----------------
        const int max_steps = net.act_fixed ? net.act_steps : 3;
#pragma unroll 1
        for (int step = 1; step <= max_steps; ++step) {
            const bool kfull = step > 1;
            // ---- z1 = ReLU(W1z [h;x] + b)
            GROUP_BARRIER();
            const int e_z1 = nev++;
            if (warp == 0) {
                if (kfull) issue_t<112, 8>(iss_p, lane, tm, A1H, A1L, D1, false, false);
                else issue_t<112, 2>(iss_p, lane, tm, A1H, A1L, D1, true, false);
                if (lane == 0) { COMMIT(e_z1); ISSUER_DRAIN(e_z1); }
            }
            __syncwarp();
            {
                WAIT_EVENT(e_z1);
                relu_stream<52>(tl + D1 + 56 * half, PB + LY[TL_Z1].bias_off + 56 * half, tl + A2H + 52 * half,
                                tl + A2L + 52 * half);
            }
            // ---- z = sigmoid(ReLU(W2z z1 + b))   (+ r1 issued back to back: it only depends on A1)
            GROUP_BARRIER();
            const int e_z2 = nev++;
            const int e_r1 = kfull ? nev++ : -1;
            if (warp == 0) {
                issue_t<64, 13>(iss_p, lane, tm, A2H, A2L, D2, false, false);
                if (lane == 0) COMMIT(e_z2);
                __syncwarp();
                if (kfull) {
                    issue_t<112, 8>(iss_p, lane, tm, A1H, A1L, D1, false, false);
                    if (lane == 0) { COMMIT(e_r1); ISSUER_DRAIN(e_r1); }
                } else {
                    if (lane == 0) { ISSUER_DRAIN(e_z2); }
                }
            }
            __syncwarp();
            {
                WAIT_EVENT(e_z2);
                float v[32];
                ld_cols<32>(tl + D2 + 32 * half, v);
                const float *b = PB + LY[TL_Z2].bias_off + 32 * half;
#pragma unroll
                for (int j = 0; j < 25; ++j) z[j] = sigmoidf_(fmaxf(v[j] + b[j], 0.0f));
            }
            if (kfull) {
                // ---- r1 -> A2, r = sigmoid(ReLU(W2r r1 + b)), [r*h ; x] -> A3
                if (!is_issuer_warp) {
                    WAIT_EVENT(e_r1);
                    relu_stream<52>(tl + D1 + 56 * half, PB + LY[TL_R1].bias_off + 56 * half, tl + A2H + 52 * half,
                                    tl + A2L + 52 * half);
                }
                GROUP_BARRIER();
                const int e_r2 = nev++;
                if (warp == 0) {
                    issue_t<64, 13>(iss_p, lane, tm, A2H, A2L, D2, false, false);
                    if (lane == 0) { COMMIT(e_r2); ISSUER_DRAIN(e_r2); }
                }
            __syncwarp();
            {
                    WAIT_EVENT(e_r2);
                    float v[32];
                    ld_cols<32>(tl + D2 + 32 * half, v);
                    const float *b = PB + LY[TL_R2].bias_off + 32 * half;
#pragma unroll
                    for (int j = 0; j < 25; ++j) {
                        const float r = sigmoidf_(fmaxf(v[j] + b[j], 0.0f));
                        v[j] = __fmul_rn(r, hn[j]);          // hn = h of the previous step (registers)
                    }
#pragma unroll
                    for (int i = 0; i < 7; ++i) v[25 + i] = xh[i];
                    st_split<32>(tl + A3H + 32 * half, tl + A3L + 32 * half, v);
                }
            }
            // ---- q = tanh(Wq [r*h ; x] + b);  h = (1 - z) h + z q
            GROUP_BARRIER();
            const int e_q = nev++;
            if (warp == 0) {
                if (kfull) issue_t<64, 8>(iss_p, lane, tm, A3H, A3L, D2, false, false);
                else issue_t<64, 2>(iss_p, lane, tm, A1H, A1L, D2, true, false);
                if (lane == 0) { COMMIT(e_q); ISSUER_DRAIN(e_q); }
            }
            __syncwarp();
            {
                WAIT_EVENT(e_q);
                float v[32];
                ld_cols<32>(tl + D2 + 32 * half, v);
                const float *b = PB + LY[TL_Q].bias_off + 32 * half;
                const float *pw = PB + net.tc_pb_pw + 25 * half;
                float pd = 0.0f;
#pragma unroll
                for (int j = 0; j < 25; ++j) {
                    const float qv = tanhf_(v[j] + b[j]);
                    const float hold = hn[j];
                    hn[j] = __fadd_rn(__fmul_rn(1.0f - z[j], hold), __fmul_rn(z[j], qv));
                    pd = fmaf(hn[j], pw[j], pd);
                    v[j] = hn[j];
                }
#pragma unroll
                for (int i = 0; i < 7; ++i) v[25 + i] = xh[i];
                st_split<32>(tl + A1H + 32 * half, tl + A1L + 32 * half, v);
                if (!net.act_fixed) {
                    sx[half * 128 + row] = pd;
                    __syncthreads();
                    const bool active = samp >= 0 && s_act[samp];
                    if (active) {
                        const float p = sigmoidf_(sx[row] + sx[128 + row] + PB[net.tc_pb_pw + 50]);
                        P += p;
#pragma unroll
                        for (int j = 0; j < 25; ++j) ACC[j] = __fadd_rn(ACC[j], __fmul_rn(p, hn[j]));
                        if (half == 0 && P < 0.95f) s_need[samp] = 1;
                    }
                }
            }
            // ---- per-sample halting decision (components.py:125-134)
            int cont;
            if (net.act_fixed) {
                if (tid < 32 && tid < nsamp) s_cnt[tid] = step;
                cont = step < max_steps;
            } else {
                if (tid == 0) s_flag[0] = 0;
                __syncthreads();
                if (tid < 32 && s_act[tid]) {
                    s_cnt[tid] = step;
                    const int need = s_need[tid];
                    s_act[tid] = need;
                    s_need[tid] = 0;
                    if (need) s_flag[0] = 1;
                }
                __syncthreads();
                cont = s_flag[0] && step < max_steps;
            }
            if (is_issuer) {
                int ns;
                if (cont) ns = 1;
                else if (args.mode == 2) {      // GRU-only mode: the next layers are the next tile's first step
                    if (iss.tiles_left > 0) { --iss.tiles_left; ns = 0; } else ns = -2;
                } else ns = 2;
                iss.decided(ns);
            }
            if (!cont) break;
        }
        __syncthreads();

        // in_mlp output: accum_hx / step_count (adaptive) or last h (act_fixed)
        float xres[25];
        if (!is_issuer_warp) {
            const float cnt = (samp >= 0) ? (float)s_cnt[samp] : 1.0f;
            const float rc = __fdividef(1.0f, cnt);
#pragma unroll
            for (int j = 0; j < 25; ++j) xres[j] = net.act_fixed ? hn[j] : (samp >= 0 ? ACC[j] * rc : 0.0f);
        }

        if (args.mode == 2) {
            if (!is_issuer_warp && row < nrows) {
                float *dst = args.gru_out + ((size_t)s0 * T + row) * HID + 25 * half;
#pragma unroll
                for (int j = 0; j < 25; ++j) dst[j] = xres[j];
                if (args.steps && half == 0 && tok == 0) args.steps[s0 + samp] = (uint8_t)s_cnt[samp];
            }
            __syncthreads();
            continue;
        }

        if (!is_issuer_warp) {
            float v[28];
#pragma unroll
            for (int j = 0; j < 25; ++j) v[j] = xres[j];
            v[25] = v[26] = v[27] = 0.0f;
            st_split<28>(tl + AEH + 28 * half, tl + AEL + 28 * half, v);
        }

        // ---------------- phase 2: 3 x post-norm encoder layer ----------------
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
            const int LB = TL_ENC + 7 * l;
            const float *LN = PB + net.tc_pb_ln + l * 200 + 25 * half;
            // ---- in_proj (one MMA per head) + attention
            GROUP_BARRIER();
            const int e_in = nev++;
            if (warp == 0) {
                issue_t<96, 7>(iss_p, lane, tm, AEH, AEL, DIN, false, false);
                issue_t<96, 7>(iss_p, lane, tm, AEH, AEL, DIN + 96, false, false);
                if (lane == 0) { COMMIT(e_in); ISSUER_DRAIN(e_in); }
            }
            __syncwarp();
            {
                WAIT_EVENT(e_in);
                float qh[32];
                {
                    float k[32], v[32];
                    ld_cols<32>(tl + DIN + 96 * half, qh);
                    ld_cols<32>(tl + DIN + 96 * half + 32, k);
                    ld_cols<32>(tl + DIN + 96 * half + 64, v);
                    const float *b = PB + LY[LB + half].bias_off;
                    float *dst = kv + row * KV_PITCH + 50 * half;
#pragma unroll
                    for (int j = 0; j < 25; ++j) {
                        qh[j] += b[j];
                        dst[j] = k[j] + b[32 + j];
                        dst[25 + j] = v[j] + b[64 + j];
                    }
                }
                __syncthreads();
                float o[28];
#pragma unroll
                for (int j = 0; j < 28; ++j) o[j] = 0.0f;
                if (samp >= 0) {
                    const float *base = kv + (samp * T) * KV_PITCH + 50 * half;
                    float m = -INFINITY;
                    for (int j = 0; j < T; ++j) {
                        const float *kp = base + j * KV_PITCH;
                        float d = 0.0f;
#pragma unroll
                        for (int c = 0; c < 25; ++c) d = fmaf(qh[c], kp[c], d);
                        m = fmaxf(m, d * 0.2f);
                    }
                    float sum = 0.0f;
                    for (int j = 0; j < T; ++j) {
                        const float *kp = base + j * KV_PITCH;
                        float d = 0.0f;
#pragma unroll
                        for (int c = 0; c < 25; ++c) d = fmaf(qh[c], kp[c], d);
                        const float p = __expf(d * 0.2f - m);
                        sum += p;
#pragma unroll
                        for (int c = 0; c < 25; ++c) o[c] = fmaf(p, kp[25 + c], o[c]);
                    }
                    const float inv = __fdividef(1.0f, sum);
#pragma unroll
                    for (int c = 0; c < 25; ++c) o[c] *= inv;
                }
                st_split<28>(tl + AATH + 28 * half, tl + AATL + 28 * half, o);
            }
            // ---- out_proj + residual + LayerNorm1
            GROUP_BARRIER();
            const int e_out = nev++;
            if (warp == 0) {
                issue_t<64, 7>(iss_p, lane, tm, AATH, AATL, DOUT, false, false);
                if (lane == 0) { COMMIT(e_out); ISSUER_DRAIN(e_out); }
            }
            __syncwarp();
            {
                WAIT_EVENT(e_out);
                float v[32];
                ld_cols<32>(tl + DOUT + 32 * half, v);
                const float *b = PB + LY[LB + 2].bias_off + 32 * half;
                // LayerNorm with one exchange: each half sends (sum, sum of squares about its own mean); the halves are
                // combined with the parallel-variance formula, which keeps the two-pass accuracy
                float s = 0.0f;
#pragma unroll
                for (int j = 0; j < 25; ++j) { xres[j] += v[j] + b[j]; s += xres[j]; }
                const float mh = s * 0.04f;
                float ss = 0.0f;
#pragma unroll
                for (int j = 0; j < 25; ++j) { const float dd = xres[j] - mh; ss = fmaf(dd, dd, ss); }
                sx[half * 128 + row] = s;
                sx[256 + half * 128 + row] = ss;
                __syncthreads();
                const float so = sx[(half ^ 1) * 128 + row], sso = sx[256 + (half ^ 1) * 128 + row];
                const float mean = (s + so) * 0.02f;
                const float dm = (s - so) * 0.04f;                      // difference of the two half means
                const float var = (ss + sso + dm * dm * 12.5f) * 0.02f;  // + n_a n_b / (n_a + n_b) * dm^2
#pragma unroll
                for (int j = 0; j < 25; ++j) xres[j] -= mean;
                const float rstd = rsqrtf(var + 1e-5f);
                float a[28];
#pragma unroll
                for (int j = 0; j < 25; ++j) { xres[j] = xres[j] * rstd * LN[j] + LN[50 + j]; a[j] = xres[j]; }
                a[25] = a[26] = a[27] = 0.0f;
                st_split<28>(tl + AEH + 28 * half, tl + AEL + 28 * half, a);
            }
            // ---- FFN in two hidden halves: f1a, f1b issued back to back; f2a overlaps the f1b epilogue
            GROUP_BARRIER();
            const int e_f1a = nev++;
            const int e_f1b = nev++;
            if (warp == 0) {
                issue_t<80, 7>(iss_p, lane, tm, AEH, AEL, DF1A, false, false);
                if (lane == 0) COMMIT(e_f1a);
                __syncwarp();
                issue_t<80, 7>(iss_p, lane, tm, AEH, AEL, DF1B, false, false);
                if (lane == 0) { COMMIT(e_f1b); ISSUER_DRAIN(e_f1b); }
            }
            __syncwarp();
            {
                WAIT_EVENT(e_f1a);
                relu_stream<40>(tl + DF1A + 40 * half, PB + LY[LB + 3].bias_off + 40 * half, tl + AHAH + 40 * half,
                                tl + AHAL + 40 * half);
            }
            GROUP_BARRIER();
            if (warp == 0) {
                issue_t<64, 10>(iss_p, lane, tm, AHAH, AHAL, DF2, false, false);    // completes before f2b (in-order tensor pipe)
            }
            __syncwarp();
            {
                WAIT_EVENT(e_f1b);
                relu_stream<40>(tl + DF1B + 40 * half, PB + LY[LB + 4].bias_off + 40 * half, tl + AHBH + 40 * half,
                                tl + AHBL + 40 * half);
            }
            GROUP_BARRIER();
            const int e_f2 = nev++;
            if (warp == 0) {
                issue_t<64, 10>(iss_p, lane, tm, AHBH, AHBL, DF2, false, true);
                if (lane == 0) { COMMIT(e_f2); ISSUER_DRAIN(e_f2); }
            }
            __syncwarp();
            {
                WAIT_EVENT(e_f2);
                float v[32];
                ld_cols<32>(tl + DF2 + 32 * half, v);
                const float *b = PB + LY[LB + 6].bias_off + 32 * half;
                // LayerNorm with one exchange: each half sends (sum, sum of squares about its own mean); the halves are
                // combined with the parallel-variance formula, which keeps the two-pass accuracy
                float s = 0.0f;
#pragma unroll
                for (int j = 0; j < 25; ++j) { xres[j] += v[j] + b[j]; s += xres[j]; }
                const float mh = s * 0.04f;
                float ss = 0.0f;
#pragma unroll
                for (int j = 0; j < 25; ++j) { const float dd = xres[j] - mh; ss = fmaf(dd, dd, ss); }
                sx[half * 128 + row] = s;
                sx[256 + half * 128 + row] = ss;
                __syncthreads();
                const float so = sx[(half ^ 1) * 128 + row], sso = sx[256 + (half ^ 1) * 128 + row];
                const float mean = (s + so) * 0.02f;
                const float dm = (s - so) * 0.04f;                      // difference of the two half means
                const float var = (ss + sso + dm * dm * 12.5f) * 0.02f;  // + n_a n_b / (n_a + n_b) * dm^2
#pragma unroll
                for (int j = 0; j < 25; ++j) xres[j] -= mean;
                const float rstd = rsqrtf(var + 1e-5f);
                float a[28];
#pragma unroll
                for (int j = 0; j < 25; ++j) { xres[j] = xres[j] * rstd * LN[100 + j] + LN[150 + j]; a[j] = xres[j]; }
                if (l < 2) { a[25] = a[26] = a[27] = 0.0f; }
                else { a[25] = selfs[0]; a[26] = selfs[1]; a[27] = selfs[2]; }   // joint = [self_state ; env]
                st_split<28>(tl + AEH + 28 * half, tl + AEL + 28 * half, a);
            }
        }

        // ---------------- phase 3: action_mlp 56 -> 100 -> 50 -> 1 ----------------
        GROUP_BARRIER();
        const int e_a0 = nev++;
        if (warp == 0) {
            issue_t<112, 7>(iss_p, lane, tm, AEH, AEL, DA0, false, false);
            if (lane == 0) { COMMIT(e_a0); ISSUER_DRAIN(e_a0); }
        }
            __syncwarp();
            {
            WAIT_EVENT(e_a0);
            relu_stream<52>(tl + DA0 + 56 * half, PB + LY[TL_A0].bias_off + 56 * half, tl + AJH + 52 * half,
                            tl + AJL + 52 * half);
        }
        GROUP_BARRIER();
        const int e_a2 = nev++;
        if (warp == 0) {
            issue_t<64, 13>(iss_p, lane, tm, AJH, AJL, DA2, false, false);
            if (lane == 0) { COMMIT(e_a2); ISSUER_DRAIN(e_a2); }
        }
            __syncwarp();
            {
            WAIT_EVENT(e_a2);
            float v[32];
            ld_cols<32>(tl + DA2 + 32 * half, v);
            const float *b = PB + LY[TL_A2].bias_off + 32 * half;
            const float *w4 = PB + net.tc_pb_a4 + 25 * half;
            float pd = 0.0f;
#pragma unroll
            for (int j = 0; j < 25; ++j) pd = fmaf(fmaxf(v[j] + b[j], 0.0f), w4[j], pd);
            sx[half * 128 + row] = pd;
            __syncthreads();
            if (half == 0 && samp >= 0 && tok == 0) {
                const float val = sx[row] + sx[128 + row] + PB[net.tc_pb_a4 + 50];
                const int gs = s0 + samp;
                if (args.net_values) args.net_values[gs] = val;
                if (args.values) args.values[gs] = __dadd_rn(args.rewards[gs], __dmul_rn(args.gamma_bar, (double)val));
                if (args.steps) args.steps[gs] = (uint8_t)s_cnt[samp];
            }
        }
        __syncthreads();
    }

    // teardown
    if (!is_issuer_warp) { tmem_wait_st(); tmem_wait_ld(); }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
#undef GROUP_BARRIER
#undef COMMIT
#undef WAIT_EVENT
#undef ISSUER_DRAIN
}

}  // namespace tc

cudaError_t launch_lookahead_tc(aem_handle h, const LookaheadArgs &a, cudaStream_t st)
{
    static bool attr_set[64] = {false};
    if (!attr_set[h->device & 63]) {
        cudaError_t e = cudaFuncSetAttribute(tc::lookahead_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)tc::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set[h->device & 63] = true;
    }
    const int ntiles = (a.total_samples + a.S - 1) / a.S;
    if (ntiles <= 0) return cudaSuccess;
    const int grid = ntiles < h->num_sms ? ntiles : h->num_sms;
    tc::lookahead_tc_kernel<<<grid, tc::NT, tc::SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

}  // namespace aem
