// lookahead_q16.cu -- tensor-core lookahead kernel, FP16x3 operands, FOUR threads per token row (sm_100a).
//
// Third tensor-core variant (precision mode 3).  Same arithmetic as lookahead_h16.cu (tcgen05.mma.kind::f16, fp16 hi/lo
// split, FP32 accumulation in TMEM, A operands packed in TMEM, in-place conversions, bulk-TMA weight ring) but organised
// to shorten the per-tile dependency chain that bounds the other two variants (profiles/r01_lookahead_h16_ncu.md):
//   * 512 threads: 16 warps, four threads per TMEM lane (token row), each owning one quarter of the columns
//     (13 / 12 / 13 / 12 of 50; a pair of quarters = one attention head) -> half the epilogue per thread;
//   * one CTA per SM with all 512 TMEM columns, which leaves room to fuse: z1|r1 is ONE N = 256 MMA group, z2 and r2 are
//     issued together, linear1 is one N = 192 group -> 23 barrier/issue/commit groups per tile instead of 31.
// Layer layouts (N / K permutation, padding) are defined in capi.cu pack_q16.
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

#include "aem_common.cuh"
#include "tc_common.cuh"

namespace aem {
namespace q16 {

using namespace aemtc;

constexpr int NT = 512;
constexpr int NBUF = 8;
constexpr int WBUF_BYTES = 16384;
constexpr int QKV_PITCH = 132;          // per row: k 4 x 16 | v 4 x 16 (quarters padded to 16) + 4: float4 access, no conflicts
constexpr int NEV = 4;
constexpr int PBLK_MAX = Q16_PBLK_FLOATS;

// ---- tensor-memory column map (512 columns) ----
constexpr uint32_t A1 = 0;              // [h;x] K = 64 : quarter q -> hi cols [8q, 8q+8), lo cols [32+8q, ..)
constexpr uint32_t D1 = 64;             // z1|r1 accumulators N = 256 = [z: 4 x 32 | r: 4 x 32], converted in place (K = 128 each)
constexpr uint32_t D2Z = 320, D2R = 384;    // z2 / r2 accumulators (N = 64 each); q reuses D2Z
constexpr uint32_t A3 = 64;             // [r*h;x] K = 64, aliases the dead hidden operand
constexpr uint32_t AE = 0;              // encoder stream operand K = 64
constexpr uint32_t DIN = 64;            // in_proj accumulators N = 192
constexpr uint32_t AATT = 256, DOUT = 320;
constexpr uint32_t DF1 = 64;            // linear1 accumulators N = 192 = 4 x 48, converted in place (K = 192)
constexpr uint32_t DF2 = 384;
constexpr uint32_t DA0 = 64, DA2 = 320;

// parameter block offsets (floats), fixed by the layer order of capi.cu pack_q16
constexpr int PB_Z1R1 = 0, PB_Z2 = 256, PB_R2 = 320, PB_Q = 384, PB_ENC = 448, PB_ENC_STRIDE = 512;
constexpr int PBE_IN = 0, PBE_OUT = 192, PBE_F1 = 256, PBE_F2 = 448;
constexpr int PB_A0 = 1984, PB_A2 = 2112, PB_LN = 2176, PB_PW = 2776, PB_A4 = 2827;
static_assert(PB_A4 + 51 <= Q16_PBLK_FLOATS, "parameter block");

constexpr size_t SMEM_BYTES = (size_t)NBUF * WBUF_BYTES + 128 * QKV_PITCH * 4 + PBLK_MAX * 4 + 8 * 128 * 4 +
                              (96 + 4) * 4 + (NBUF + NEV) * 8 + 16;

// real column index of slot j of quarter q (50 = 13 + 12 + 13 + 12), -1 = padding
__device__ __forceinline__ constexpr int qsize(int q) { return (q & 1) ? 12 : 13; }
__device__ __forceinline__ constexpr int qstart(int q) { return q == 0 ? 0 : (q == 1 ? 13 : (q == 2 ? 25 : 38)); }

__device__ __forceinline__ void quad_sync(int q4) { asm volatile("bar.sync %0, 128;" ::"r"(q4 + 1) : "memory"); }

__device__ __forceinline__ void ld8f(uint32_t taddr, float *v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void st4u(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3])
                 : "memory");
}
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo)
{
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
__device__ __forceinline__ void st8_pack(uint32_t hi_addr, uint32_t lo_addr, const float *v)
{
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    st4u(hi_addr, h);
    st4u(lo_addr, l);
}
// 16 slots of this thread's quarter -> 8 hi columns + 8 lo columns
__device__ __forceinline__ void st16_pack(uint32_t hi_addr, uint32_t lo_addr, const float *v)
{
    st8_pack(hi_addr, lo_addr, v);
    st8_pack(hi_addr + 4, lo_addr + 4, v + 8);
}

__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

// in-place ReLU conversion of W accumulator columns [base, base + W) -> hi [base, base + W/2), lo [base + W/2, base + W)
template <int W>
__device__ __forceinline__ void relu_inplace(uint32_t base, const float *__restrict__ bias)
{
    constexpr int NB = W / 16;
    float buf[NB * 8];
#pragma unroll
    for (int c = 0; c < NB; ++c) ld8f(base + W / 2 + 8 * c, buf + 8 * c);
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        float v[8];
        ld8f(base + 8 * c, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j] + bias[8 * c + j], 0.0f);
        st8_pack(base + 4 * c, base + W / 2 + 4 * c, v);
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        float *v = buf + 8 * c;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j] + bias[W / 2 + 8 * c + j], 0.0f);
        st8_pack(base + W / 4 + 4 * c, base + W / 2 + W / 4 + 4 * c, v);
    }
}

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
// issuer
// ------------------------------------------------------------------------------------------------
struct LayerInfo { int off_hi, off_lo, bytes, nsplit; };
// Weight ring bookkeeping (shared memory).  Two single-thread roles, neither on the path between a group barrier and the
// group's first MMA:
//   consumer = warp 0: the ring read position travels in registers (IssReg) while a group is issued; after the commit
//              lane 0 publishes the running chunk count in used[group & 1];
//   producer = lane 0 of warp 1: after each group barrier it releases the slots of all PREVIOUS groups (complete: every
//              thread waited on their events before the barrier) and requests the next chunks, in parallel with warp 0
//              issuing and the tensor pipe working; it also switches stages at the ACT decision.
struct Issuer {
    int used[2];                    // consumer -> producer: chunks consumed up to and including group e, at [e & 1]
    int ubuf, upar;                 // consumer: ring slot and mbarrier parity of the next chunk to consume
    int nloaded, nfree;             // producer: chunks requested from L2 / released for refill
    int lbuf;                       // producer: ring slot of the next request
    int pf_cur, pf_end, pf_stage;   // producer: cursor into the chunk table; stage -1 = waiting for the ACT decision, -2 = done
    int tiles_left;
};
struct IssReg { uint32_t buf, par, n; };
constexpr int MAX_CHUNKS = 96;
__device__ __forceinline__ int stage_len(int st) { return st == 0 ? 3 : (st == 1 ? 4 : 14); }
__device__ __forceinline__ int layer_at(int st, int idx)
{
    if (st == 0) return idx == 0 ? QL_Z1 : (idx == 1 ? QL_Z2 : QL_Q);
    if (st == 1) return idx == 0 ? QL_Z1R1 : (idx == 1 ? QL_Z2 : (idx == 2 ? QL_R2 : QL_Q));
    return idx < 12 ? QL_ENC + idx : (idx == 12 ? QL_A0 : QL_A2);
}
// chunk table: the {offset in halves, bytes} of every ring chunk of the three stages, in consumption order
__device__ __noinline__ void build_chunk_table(const LayerInfo *layers, int2 *chunks, int *stage_beg)
{
    int n = 0;
    for (int st = 0; st < 3; ++st) {
        stage_beg[st] = n;
        for (int i = 0; i < stage_len(st); ++i) {
            const LayerInfo L = layers[layer_at(st, i)];
            const int bytes = L.bytes / L.nsplit;
            for (int part = 0; part < 2; ++part)
                for (int sub = 0; sub < L.nsplit; ++sub) chunks[n++] = make_int2((part ? L.off_lo : L.off_hi) + sub * (bytes / 2), bytes);
        }
    }
    stage_beg[3] = n;
}
__device__ __forceinline__ void set_stage(Issuer *s, const int *stage_beg, int st)
{
    s->pf_stage = st;
    if (st >= 0) { s->pf_cur = stage_beg[st]; s->pf_end = stage_beg[st + 1]; }
}
// request as many chunks as the ring has room for (lane 0 of warp 0)
__device__ __noinline__ void prefetch(Issuer *s, const int2 *chunks, const int *stage_beg, uint8_t *wbuf, uint64_t *wfull,
                                         const __half *img)
{
    int stage = s->pf_stage;
    int nloaded = s->nloaded;
    int room = NBUF - (nloaded - s->nfree);
    if (stage < 0 || room <= 0) return;
    int cur = s->pf_cur, end = s->pf_end, lbuf = s->lbuf, tiles_left = s->tiles_left;
    while (room > 0) {
        const int2 c = chunks[cur];
        mbar_expect_tx(&wfull[lbuf], (uint32_t)c.y);
        bulk_g2s(wbuf + (size_t)lbuf * WBUF_BYTES, img + c.x, (uint32_t)c.y, &wfull[lbuf]);
        ++nloaded; --room;
        if (++lbuf == NBUF) lbuf = 0;
        if (++cur == end) {
            if (stage == 2 && tiles_left > 0) { --tiles_left; stage = 0; cur = stage_beg[0]; end = stage_beg[1]; }
            else { stage = (stage == 2) ? -2 : -1; break; }
        }
    }
    s->nloaded = nloaded; s->lbuf = lbuf; s->pf_cur = cur; s->pf_end = end; s->pf_stage = stage; s->tiles_left = tiles_left;
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// KS k-steps; operand columns of k-step ks: a_base + HSTRIDE * (ks / KSH) + 8 * (ks % KSH); lo part LO columns further.
// Executed by all lanes of warp 0 (uniform control flow), the MMAs by the elected lane.
template <int N, int KS, int KSH, int HSTRIDE, int LO, int SPLIT>
__device__ __forceinline__ void issue_t(IssReg &r, uint64_t *wfull, uint32_t ring, uint32_t tm, uint32_t a_base, uint32_t d)
{
    constexpr uint32_t LBO = (uint32_t)N * 16;
    constexpr uint64_t DHI = ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr int KSS = KS / SPLIT;
    static_assert(KS % SPLIT == 0, "split");
    const bool leader = elect_one();
#pragma unroll
    for (int part = 0; part < 2; ++part) {
#pragma unroll
        for (int sub = 0; sub < SPLIT; ++sub) {
            mbar_wait(&wfull[r.buf], r.par);
            const uint32_t base = ((ring + r.buf * (uint32_t)WBUF_BYTES) & 0x3FFFFu) >> 4;
            if (++r.buf == (uint32_t)NBUF) { r.buf = 0; r.par ^= 1u; }
            ++r.n;
#pragma unroll
            for (int kk = 0; kk < KSS; ++kk) {
                const int ks = sub * KSS + kk;
                const uint32_t acol = (uint32_t)(HSTRIDE * (ks / KSH) + 8 * (ks % KSH));
                const uint64_t bd = DHI | (uint64_t)(base + kk * ((2 * LBO) >> 4));
                if (leader) {
                    if (part == 0) {
                        mma_f16_ts(tm + d, tm + a_base + acol, bd, IDESC, ks == 0 ? 0u : 1u);
                        mma_f16_ts(tm + d, tm + a_base + acol + LO, bd, IDESC, 1u);
                    } else {
                        mma_f16_ts(tm + d, tm + a_base + acol, bd, IDESC, 1u);
                    }
                }
            }
        }
    }
}
// operand layouts: contiguous K = 64 (hi 32 cols | lo 32 cols); in place [4 x 32] (K = 128); in place [4 x 48] (K = 192)
#define ISSUE_K64(N, a, d) issue_t<N, 4, 4, 0, 32, ((N) * 64 * 2 > 16384 ? 2 : 1)>(ir, wfull, ring, tm, a, d)
#define ISSUE_K128(N, a, d) issue_t<N, 8, 2, 32, 16, 1>(ir, wfull, ring, tm, a, d)
#define ISSUE_K192(N, a, d) issue_t<N, 12, 3, 48, 24, 2>(ir, wfull, ring, tm, a, d)

__global__ void __launch_bounds__(NT, 1) lookahead_q16_kernel(const LookaheadArgs args)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *wbuf = smem_raw;
    float *qkv = reinterpret_cast<float *>(wbuf + (size_t)NBUF * WBUF_BYTES);
    float *pblk = qkv + 128 * QKV_PITCH;
    float *sx = pblk + PBLK_MAX;                                 // [8][128] exchange
    int *s_cnt = reinterpret_cast<int *>(sx + 8 * 128);
    int *s_act = s_cnt + 32;
    int *s_need = s_act + 32;
    int *s_flag = s_need + 32;
    uint64_t *wfull = reinterpret_cast<uint64_t *>(s_flag + 4);
    uint64_t *evb = wfull + NBUF;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(evb + NEV);
    __shared__ Issuer s_iss;
    __shared__ LayerInfo s_layers[QL_NUM];
    __shared__ int2 s_chunks[MAX_CHUNKS];
    __shared__ int s_stage_beg[4];

    const NetParams &net = args.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qd = warp >> 2, q4 = warp & 3;          // column quarter, lane quarter
    const int hd = qd >> 1, sub = qd & 1;             // attention head and half of it
    const int row = q4 * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const int T = args.T, S = args.S;
    const int ntiles = (args.total_samples + S - 1) / S;
    const int QS = qsize(qd), Q0 = qstart(qd);        // real columns [Q0, Q0 + QS) of 50 belong to this thread

    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) mbar_init(&wfull[i], 1);
        for (int i = 0; i < NEV; ++i) mbar_init(&evb[i], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    for (int i = tid; i < net.q16_pblk_floats; i += NT) pblk[i] = net.q16_pblk[i];
    if (tid < QL_NUM) {
        const TcLayer L = net.q16_layers_d[tid];
        s_layers[tid].off_hi = L.off_hi; s_layers[tid].off_lo = L.off_lo; s_layers[tid].bytes = L.bytes;
        s_layers[tid].nsplit = L.bias_off;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_slot;
    const uint32_t tl = tm + lane_base;
    const uint32_t ring = smem_u32(wbuf);
    if (tid == 32) {
        build_chunk_table(s_layers, s_chunks, s_stage_beg);
        s_iss.nloaded = s_iss.nfree = s_iss.used[0] = s_iss.used[1] = 0;
        s_iss.lbuf = s_iss.ubuf = s_iss.upar = 0;
        s_iss.pf_cur = s_iss.pf_end = 0; s_iss.tiles_left = 0;
        s_iss.pf_stage = -2;
        if ((int)blockIdx.x < ntiles) {
            s_iss.tiles_left = (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x;
            set_stage(&s_iss, s_stage_beg, 0);
            prefetch(&s_iss, s_chunks, s_stage_beg, wbuf, wfull, net.q16_img);
        }
    }
    __syncthreads();
    int nev = 0;

#define GROUP_BARRIER()                     \
    do {                                    \
        TS();                               \
        tmem_wait_st();                     \
        tc_fence_before();                  \
        __syncthreads();                    \
        if (warp == 0) tc_fence_after();    \
        TS();                               \
    } while (0)
#define COMMIT(e) mma_commit(&evb[(e) % NEV])
#define WAIT_EVENT(e)                                                   \
    do {                                                                \
        TS();                                                           \
        mbar_wait(&evb[(e) % NEV], (uint32_t)(((e) / NEV) & 1));        \
        tc_fence_after();                                               \
        TS();                                                           \
    } while (0)
// Issue protocol of warp 0: the ring read position travels in registers while the group's MMAs are issued; after the
// commit, while the tensor pipe works, lane 0 releases the slots of the PREVIOUS group (complete: every thread waited on
// its event before this group's barrier) and requests the next chunks into them.
#define ISSUE_BEGIN() \
    IssReg ir;        \
    ir.buf = (uint32_t)s_iss.ubuf; ir.par = (uint32_t)s_iss.upar; ir.n = 0
#define ISSUE_END(e)                                                                  \
    do {                                                                              \
        if (lane == 0) {                                                              \
            TS();                                                                     \
            COMMIT(e);                                                                \
            s_iss.used[(e) & 1] = s_iss.used[((e) & 1) ^ 1] + (int)ir.n;              \
            s_iss.ubuf = (int)ir.buf; s_iss.upar = (int)ir.par;                       \
            TS();                                                                     \
        }                                                                             \
    } while (0)
#define PRODUCE(e)                                                                    \
    do {                                                                              \
        if (tid == 32) {                                                              \
            s_iss.nfree = s_iss.used[((e) & 1) ^ 1];                                  \
            prefetch(&s_iss, s_chunks, s_stage_beg, wbuf, wfull, net.q16_img);        \
        }                                                                             \
        __syncwarp();                                                                 \
    } while (0)

    const float *PB = pblk;
    // optional timeline (debug): stamps of the first tile of CTA 0; slots [0,256) thread 32, [256,512) thread 0
    int ts_n = 0;
#define TS()                                                                                   \
    do {                                                                                       \
        if (args.dbg && tile == args.stagger_ns * (int)gridDim.x && (tid == 32 || tid == 0) && ts_n < 250) \
            args.dbg[(tid == 0 ? 256 : 0) + ts_n++] = clock64();                               \
    } while (0)

#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int s0 = tile * S;
        const int nsamp = min(S, args.total_samples - s0);
        const int nrows = nsamp * T;

        // per-thread state: 13 (or 12) of the 50 columns of this row
        float xq[4], selfq[2], ACC[13], hn[13], z[13];
        float P = 0.0f;
        int samp = -1, tok = 0;
#pragma unroll
        for (int j = 0; j < 13; ++j) { ACC[j] = 0.0f; hn[j] = 0.0f; z[j] = 0.0f; }

        // ---------------- phase 0: inputs -> A1 = [h = 0 ; x] ----------------
        {
            float o[XD];
#pragma unroll
            for (int k = 0; k < XD; ++k) o[k] = 0.0f;
            float s6[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) s6[k] = 0.0f;
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
                for (int k = 0; k < 6; ++k) s6[k] = o[k];
                if (args.mode != 2 && tok == 0) {
                    o[6] = 0.0f; o[7] = 0.0f; o[8] = o[4]; o[9] = o[5]; o[10] = o[3]; o[11] = 0.0f; o[12] = o[3];
                }
            }
            // x slots of the quarters: q0 -> x0..2, q1 -> x3..6, q2 -> x7..9, q3 -> x10..12
            xq[0] = qd == 0 ? o[0] : (qd == 1 ? o[3] : (qd == 2 ? o[7] : o[10]));
            xq[1] = qd == 0 ? o[1] : (qd == 1 ? o[4] : (qd == 2 ? o[8] : o[11]));
            xq[2] = qd == 0 ? o[2] : (qd == 1 ? o[5] : (qd == 2 ? o[9] : o[12]));
            xq[3] = qd == 1 ? o[6] : 0.0f;
            // self_state slots for the action input: q0 -> s0,s1; q1 -> s2; q2 -> s3,s4; q3 -> s5
            selfq[0] = qd == 0 ? s6[0] : (qd == 1 ? s6[2] : (qd == 2 ? s6[3] : s6[5]));
            selfq[1] = qd == 0 ? s6[1] : (qd == 2 ? s6[4] : 0.0f);
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // slot QS + i (13 or 12 real h slots first)
                if (qd & 1) v[12 + i] = xq[i]; else if (i < 3) v[13 + i] = xq[i];
            }
            st16_pack(tl + A1 + 8 * qd, tl + A1 + 32 + 8 * qd, v);
            if (tid < 32) { s_cnt[tid] = 0; s_act[tid] = (tid < nsamp) ? 1 : 0; s_need[tid] = 0; }
        }

        // ---------------- phase 1: GRU + adaptive computation time ----------------
        const int max_steps = net.act_fixed ? net.act_steps : 3;
#pragma unroll 1
        for (int step = 1; step <= max_steps; ++step) {
            const bool kfull = step > 1;
            // ---- z1 | r1 (one N = 256 group; step 1 needs z1 only)
            GROUP_BARRIER();
            const int e_1 = nev++;
            PRODUCE(e_1);
            if (warp == 0) {
                ISSUE_BEGIN();
                if (kfull) ISSUE_K64(256, A1, D1);
                else ISSUE_K64(128, A1, D1);
                ISSUE_END(e_1);
            }
            __syncwarp();
            WAIT_EVENT(e_1);
            relu_inplace<32>(tl + D1 + 32 * qd, PB + PB_Z1R1 + 32 * qd);
            if (kfull) relu_inplace<32>(tl + D1 + 128 + 32 * qd, PB + PB_Z1R1 + 128 + 32 * qd);
            // ---- z2 and r2 together
            GROUP_BARRIER();
            const int e_2 = nev++;
            PRODUCE(e_2);
            if (warp == 0) {
                ISSUE_BEGIN();
                ISSUE_K128(64, D1, D2Z);
                if (kfull) ISSUE_K128(64, D1 + 128, D2R);
                ISSUE_END(e_2);
            }
            __syncwarp();
            WAIT_EVENT(e_2);
            {
                float vz[16], vr[16];
                ld8f(tl + D2Z + 16 * qd, vz);
                ld8f(tl + D2Z + 16 * qd + 8, vz + 8);
                if (kfull) { ld8f(tl + D2R + 16 * qd, vr); ld8f(tl + D2R + 16 * qd + 8, vr + 8); }
                tmem_wait_ld();
                const float *bz = PB + PB_Z2 + 16 * qd, *br = PB + PB_R2 + 16 * qd;
#pragma unroll
                for (int j = 0; j < 13; ++j) z[j] = sigmoidf_(fmaxf(vz[j] + bz[j], 0.0f));
                if (kfull) {
                    float a3[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) a3[j] = 0.0f;
#pragma unroll
                    for (int j = 0; j < 13; ++j)
                        if (j < QS) a3[j] = __fmul_rn(sigmoidf_(fmaxf(vr[j] + br[j], 0.0f)), hn[j]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (qd & 1) a3[12 + i] = xq[i]; else if (i < 3) a3[13 + i] = xq[i];
                    }
                    st16_pack(tl + A3 + 8 * qd, tl + A3 + 32 + 8 * qd, a3);
                }
            }
            // ---- q, h = (1 - z) h + z q
            GROUP_BARRIER();
            const int e_q = nev++;
            PRODUCE(e_q);
            if (warp == 0) {
                ISSUE_BEGIN();
                ISSUE_K64(64, kfull ? A3 : A1, D2Z);
                ISSUE_END(e_q);
            }
            __syncwarp();
            WAIT_EVENT(e_q);
            {
                float v[16];
                ld8f(tl + D2Z + 16 * qd, v);
                ld8f(tl + D2Z + 16 * qd + 8, v + 8);
                tmem_wait_ld();
                const float *b = PB + PB_Q + 16 * qd;
                const float *pw = PB + PB_PW + Q0;
                float pd = 0.0f;
#pragma unroll
                for (int j = 0; j < 13; ++j) {
                    if (j < QS) {
                        const float qv = tanhf_(v[j] + b[j]);
                        hn[j] = __fadd_rn(__fmul_rn(1.0f - z[j], hn[j]), __fmul_rn(z[j], qv));
                        pd = fmaf(hn[j], pw[j], pd);
                        v[j] = hn[j];
                    }
                }
#pragma unroll
                for (int j = 12; j < 16; ++j)
                    if (j >= QS) v[j] = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (qd & 1) v[12 + i] = xq[i]; else if (i < 3) v[13 + i] = xq[i];
                }
                st16_pack(tl + A1 + 8 * qd, tl + A1 + 32 + 8 * qd, v);
                if (!net.act_fixed) {
                    sx[qd * 128 + row] = pd;
                    quad_sync(q4);
                    const bool active = samp >= 0 && s_act[samp];
                    if (active) {
                        const float p = sigmoidf_(sx[row] + sx[128 + row] + sx[256 + row] + sx[384 + row] + PB[PB_PW + 50]);
                        P += p;
#pragma unroll
                        for (int j = 0; j < 13; ++j) ACC[j] = __fadd_rn(ACC[j], __fmul_rn(p, hn[j]));
                        if (qd == 0 && P < 0.95f) s_need[samp] = 1;
                    }
                }
            }
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
            if (tid == 32 && s_iss.pf_stage == -1) {
                int ns;
                if (cont) ns = 1;
                else if (args.mode == 2) {
                    if (s_iss.tiles_left > 0) { --s_iss.tiles_left; ns = 0; } else ns = -2;
                } else ns = 2;
                set_stage(&s_iss, s_stage_beg, ns);
                prefetch(&s_iss, s_chunks, s_stage_beg, wbuf, wfull, net.q16_img);
            }
            if (!cont) break;
        }
        __syncthreads();

        float xres[13];
        {
            const float cnt = (samp >= 0) ? (float)s_cnt[samp] : 1.0f;
            const float rc = __fdividef(1.0f, cnt);
#pragma unroll
            for (int j = 0; j < 13; ++j) xres[j] = net.act_fixed ? hn[j] : (samp >= 0 ? ACC[j] * rc : 0.0f);
        }
        if (args.mode == 2) {
            if (row < nrows) {
                float *dst = args.gru_out + ((size_t)s0 * T + row) * HID + Q0;
#pragma unroll
                for (int j = 0; j < 13; ++j)
                    if (j < QS) dst[j] = xres[j];
                if (args.steps && qd == 0 && tok == 0) args.steps[s0 + samp] = (uint8_t)s_cnt[samp];
            }
            __syncthreads();
            continue;
        }
        {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (j < 13 && j < QS) ? xres[j < 13 ? j : 0] : 0.0f;
            st16_pack(tl + AE + 8 * qd, tl + AE + 32 + 8 * qd, v);
        }

        // ---------------- phase 2: 3 x post-norm encoder layer ----------------
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
            const int PBL = PB_ENC + PB_ENC_STRIDE * l;
            const float *LN = PB + PB_LN + l * 200 + Q0;
            GROUP_BARRIER();
            const int e_in = nev++;
            PRODUCE(e_in);
            if (warp == 0) {
                ISSUE_BEGIN();
                ISSUE_K64(192, AE, DIN);
                ISSUE_END(e_in);
            }
            __syncwarp();
            WAIT_EVENT(e_in);
            float o[16];
            {
                // D_IN columns: head hd = [q: 32 | k: 32 | v: 32], each 32 = the two quarters of the head padded to 16
                // (13 or 12 real columns, the rest exactly zero).  Every thread of a row can read the whole TMEM row, so
                // q (all 32 slots of its head) stays in registers; k and v (own quarter) go to shared memory.
                const uint32_t base = tl + DIN + 96 * hd;
                const float *b = PB + PBL + PBE_IN + 96 * hd;
                float qh[32];
                {
                    float kq[16], vq[16];
#pragma unroll
                    for (int c = 0; c < 4; ++c) ld8f(base + 8 * c, qh + 8 * c);
                    ld8f(base + 32 + 16 * sub, kq);
                    ld8f(base + 32 + 16 * sub + 8, kq + 8);
                    ld8f(base + 64 + 16 * sub, vq);
                    ld8f(base + 64 + 16 * sub + 8, vq + 8);
                    tmem_wait_ld();
                    float4 *dst = reinterpret_cast<float4 *>(qkv + row * QKV_PITCH + 16 * qd);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float *bk = b + 32 + 16 * sub + 4 * c, *bv = b + 64 + 16 * sub + 4 * c;
                        dst[c] = make_float4(kq[4 * c] + bk[0], kq[4 * c + 1] + bk[1], kq[4 * c + 2] + bk[2], kq[4 * c + 3] + bk[3]);
                        dst[16 + c] = make_float4(vq[4 * c] + bv[0], vq[4 * c + 1] + bv[1], vq[4 * c + 2] + bv[2], vq[4 * c + 3] + bv[3]);
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c) qh[c] = (qh[c] + b[c]) * 0.2f;      // 1 / sqrt(head_dim = 25)
                }
                tc_fence_before();
                __syncthreads();
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = 0.0f;
                if (samp >= 0) {
                    // one pass over the T keys of the sample with a running maximum (streaming softmax)
                    const float *kb = qkv + (samp * T) * QKV_PITCH;
                    float m = -INFINITY, sum = 0.0f;
                    for (int j = 0; j < T; ++j, kb += QKV_PITCH) {
                        const float4 *kp = reinterpret_cast<const float4 *>(kb + 32 * hd);
                        float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float4 k4 = kp[c];
                            d0 = fmaf(qh[4 * c], k4.x, d0); d1 = fmaf(qh[4 * c + 1], k4.y, d1);
                            d2 = fmaf(qh[4 * c + 2], k4.z, d2); d3 = fmaf(qh[4 * c + 3], k4.w, d3);
                        }
                        const float sc = (d0 + d1) + (d2 + d3);
                        const float mn = fmaxf(m, sc);
                        const float rescale = __expf(m - mn), p = __expf(sc - mn);
                        m = mn;
                        sum = fmaf(sum, rescale, p);
                        const float4 *vp = reinterpret_cast<const float4 *>(kb + 64 + 16 * qd);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float4 v4 = vp[c];
                            o[4 * c] = fmaf(p, v4.x, o[4 * c] * rescale); o[4 * c + 1] = fmaf(p, v4.y, o[4 * c + 1] * rescale);
                            o[4 * c + 2] = fmaf(p, v4.z, o[4 * c + 2] * rescale); o[4 * c + 3] = fmaf(p, v4.w, o[4 * c + 3] * rescale);
                        }
                    }
                    const float inv = __fdividef(1.0f, sum);
#pragma unroll
                    for (int c = 0; c < 16; ++c) o[c] *= inv;
                }
            }
            st16_pack(tl + AATT + 8 * qd, tl + AATT + 32 + 8 * qd, o);
            // ---- out_proj + residual + LayerNorm1
            GROUP_BARRIER();
            const int e_out = nev++;
            PRODUCE(e_out);
            if (warp == 0) {
                ISSUE_BEGIN();
                ISSUE_K64(64, AATT, DOUT);
                ISSUE_END(e_out);
            }
            __syncwarp();
            WAIT_EVENT(e_out);
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 1) {
                    // ---- FFN: linear1 (one N = 192 group, converted in place), linear2, residual + LayerNorm2
                    GROUP_BARRIER();
                    const int e_f1 = nev++;
                    PRODUCE(e_f1);
                    if (warp == 0) {
                        ISSUE_BEGIN();
                        ISSUE_K64(192, AE, DF1);
                        ISSUE_END(e_f1);
                    }
                    __syncwarp();
                    WAIT_EVENT(e_f1);
                    relu_inplace<48>(tl + DF1 + 48 * qd, PB + PBL + PBE_F1 + 48 * qd);
                    GROUP_BARRIER();
                    const int e_f2 = nev++;
                    PRODUCE(e_f2);
                    if (warp == 0) {
                        ISSUE_BEGIN();
                        ISSUE_K192(64, DF1, DF2);
                        ISSUE_END(e_f2);
                    }
                    __syncwarp();
                    WAIT_EVENT(e_f2);
                }
                // residual + LayerNorm over the 50 columns held by the four threads of the row
                float v[16];
                const uint32_t dcol = (pass == 0 ? DOUT : DF2) + 16 * qd;
                ld8f(tl + dcol, v);
                ld8f(tl + dcol + 8, v + 8);
                tmem_wait_ld();
                const float *b = PB + PBL + (pass == 0 ? PBE_OUT : PBE_F2) + 16 * qd;
                float s = 0.0f;
#pragma unroll
                for (int j = 0; j < 13; ++j)
                    if (j < QS) { xres[j] += v[j] + b[j]; s += xres[j]; }
                const float mq = s / (float)QS;
                float ss = 0.0f;
#pragma unroll
                for (int j = 0; j < 13; ++j)
                    if (j < QS) { const float dd = xres[j] - mq; ss = fmaf(dd, dd, ss); }
                sx[qd * 128 + row] = s;
                sx[512 + qd * 128 + row] = ss;
                quad_sync(q4);
                float stot = 0.0f;
#pragma unroll
                for (int k = 0; k < 4; ++k) stot += sx[k * 128 + row];
                const float mean = stot * 0.02f;
                float var = 0.0f;            // sum over quarters of ss_k + n_k (mean_k - mean)^2
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float nk = (float)qsize(k);
                    const float dk = sx[k * 128 + row] / nk - mean;
                    var += sx[512 + k * 128 + row] + nk * dk * dk;
                }
                const float rstd = rsqrtf(var * 0.02f + 1e-5f);
                const float *g = LN + (pass == 0 ? 0 : 100);
                float a[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) a[j] = 0.0f;
#pragma unroll
                for (int j = 0; j < 13; ++j)
                    if (j < QS) { xres[j] = (xres[j] - mean) * rstd * g[j] + g[50 + j]; a[j] = xres[j]; }
                if (pass == 1 && l == 2) {          // joint = [self_state ; env] for the action MLP
                    if (qd & 1) a[12] = selfq[0]; else { a[13] = selfq[0]; a[14] = selfq[1]; }
                }
                st16_pack(tl + AE + 8 * qd, tl + AE + 32 + 8 * qd, a);
                quad_sync(q4);      // sx is reused by the next exchange of this row group
            }
        }

        // ---------------- phase 3: action_mlp ----------------
        GROUP_BARRIER();
        const int e_a0 = nev++;
        PRODUCE(e_a0);
        if (warp == 0) {
            ISSUE_BEGIN();
            ISSUE_K64(128, AE, DA0);
            ISSUE_END(e_a0);
        }
        __syncwarp();
        WAIT_EVENT(e_a0);
        relu_inplace<32>(tl + DA0 + 32 * qd, PB + PB_A0 + 32 * qd);
        GROUP_BARRIER();
        const int e_a2 = nev++;
        PRODUCE(e_a2);
        if (warp == 0) {
            ISSUE_BEGIN();
            ISSUE_K128(64, DA0, DA2);
            ISSUE_END(e_a2);
        }
        __syncwarp();
        WAIT_EVENT(e_a2);
        {
            float v[16];
            ld8f(tl + DA2 + 16 * qd, v);
            ld8f(tl + DA2 + 16 * qd + 8, v + 8);
            tmem_wait_ld();
            const float *b = PB + PB_A2 + 16 * qd;
            const float *w4 = PB + PB_A4 + Q0;
            float pd = 0.0f;
#pragma unroll
            for (int j = 0; j < 13; ++j)
                if (j < QS) pd = fmaf(fmaxf(v[j] + b[j], 0.0f), w4[j], pd);
            sx[qd * 128 + row] = pd;
            quad_sync(q4);
            if (qd == 0 && samp >= 0 && tok == 0) {
                const float val = sx[row] + sx[128 + row] + sx[256 + row] + sx[384 + row] + PB[PB_A4 + 50];
                const int gs = s0 + samp;
                if (args.net_values) args.net_values[gs] = val;
                if (args.values) args.values[gs] = __dadd_rn(args.rewards[gs], __dmul_rn(args.gamma_bar, (double)val));
                if (args.steps) args.steps[gs] = (uint8_t)s_cnt[samp];
            }
        }
        __syncthreads();
    }

    tmem_wait_st();
    tmem_wait_ld();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
#undef GROUP_BARRIER
#undef TS
#undef COMMIT
#undef WAIT_EVENT
#undef ISSUE_BEGIN
#undef ISSUE_END
#undef PRODUCE
}

}  // namespace q16

cudaError_t launch_lookahead_q16(aem_handle h, const LookaheadArgs &a, cudaStream_t st)
{
    static bool attr_set[64] = {false};
    if (!attr_set[h->device & 63]) {
        cudaError_t e = cudaFuncSetAttribute(q16::lookahead_q16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)q16::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set[h->device & 63] = true;
    }
    const int ntiles = (a.total_samples + a.S - 1) / a.S;
    if (ntiles <= 0) return cudaSuccess;
    const int grid = ntiles < h->num_sms ? ntiles : h->num_sms;
    if (getenv("AEM_Q16_TIMELINE")) {       // debug: print the clock64 timeline of CTA 0 / tile 0
        LookaheadArgs b = a;
        long long *dbg = nullptr, host[512];
        cudaMalloc(&dbg, sizeof(host));
        cudaMemset(dbg, 0, sizeof(host));
        b.dbg = dbg;
        b.stagger_ns = atoi(getenv("AEM_Q16_TIMELINE"));      // which tile of CTA 0 to stamp
        q16::lookahead_q16_kernel<<<grid, q16::NT, q16::SMEM_BYTES, st>>>(b);
        cudaStreamSynchronize(st);
        cudaMemcpy(host, dbg, sizeof(host), cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        for (int t = 0; t < 2; ++t) {
            fprintf(stderr, "timeline thread %d:", t == 0 ? 32 : 0);
            for (int i = 0; i < 250 && host[256 * t + i]; ++i)
                fprintf(stderr, " %lld", host[256 * t + i] - host[256 * t]);
            fprintf(stderr, "\n");
        }
        return cudaGetLastError();
    }
    q16::lookahead_q16_kernel<<<grid, q16::NT, q16::SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

}  // namespace aem
