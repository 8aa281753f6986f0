// lookahead_h16.cu -- tensor-core lookahead kernel, FP16x3 operands, TWO tiles in flight per SM (sm_100a).
// The shipped default (aem_set_precision mode 2).  History and measurements: profiles/r01_lookahead_h16_final_ncu.md,
// profiles/r02_lookahead_h16_ncu.md.
//
// Same function as lookahead.cu / lookahead_tc.cu, laid out so that one 128-row tile needs only 256 TMEM columns,
// 128 registers / thread and 114 KB of shared memory; two CTAs then share an SM and the hardware overlaps the MMA /
// commit latency of one tile with the epilogue of the other.
//   * operands: x = hi + lo, hi and lo fp16; D = A_hi W_hi + A_lo W_hi + A_hi W_lo with tcgen05.mma.kind::f16 (K = 16
//     per instruction, FP32 accumulation in TMEM): ~21-22 mantissa bits for this network's value ranges
//     (tools/tc_probe16.cu: ~1e-6 relative), same class as 3xTF32;
//   * A operands live in TMEM, two fp16 per 32-bit column (even k in the low half); layers whose output feeds the next
//     GEMM directly (z1, r1, linear1, action_mlp.0) are converted IN PLACE, 16 accumulator columns -> 8 hi + 8 lo operand
//     columns of the same thread, with the ReLU inside the conversions (hi rounds toward zero);
//   * biases ride in the weight images: slot 31 of half 1 of every operand is kept at exactly 1.0 (capi.cu pack_h16);
//   * weight images img[k/8][n][8] (fp16, hi and lo) stream through a 4 x 12 KB bulk-copy ring: warp 0 consumes (ring
//     position in registers, chunk count published before each tcgen05.commit), lane 0 of warp 1 produces (releases the
//     slots after the event, requests the next chunks from a flat table);
//   * rows of a tile: token 0 of every sample first, so the last encoder layer and the action MLP (token 0 only) skip the
//     epilogues of warps without such a row;
//   * one copy of the gate (z / r), FFN-chunk and LayerNorm code (the two CTAs of an SM run half a tile apart: the
//     instruction cache decides), attention with float4 key / value rows and a streaming softmax.
// Round 2: the kernel was bound by the instruction issue of its epilogues (profiles/r01: 16.4 k warp instructions per
// warp and tile, tensor pipe 35 % busy), so every epilogue now works on PAIRS of columns:
//   * packed FP32 arithmetic (FFMA2 / FADD2 / FMUL2: __ffma2_rn ...) for the gate blend, the ACT accumulation, LayerNorm,
//     the attention dot products / weighted sums and the residual of the hi / lo split (5 instructions per pair);
//   * sigma / tanh from ex2.approx + rcp.approx without the range fix-ups of __expf / __fdividef (4 / 3.5 instructions per
//     element instead of 8 / 9), softmax in the log2 domain (log2 e folded into the query scale);
//   * the agent-centric rotation from (dx, dy) / |d| instead of atan2 -> cos / sin;
//   * tcgen05.ld x16, the x slots of the GRU operands split once per tile, parameters as float4 rows;
//   * the halting vote of a GRU step is ONE barrier (__syncthreads_or + per-step flags) that doubles as the group
//     barrier of the next step; every thread counts its own sample's steps.
// Layout of every layer (N / K permutation and padding) is defined in capi.cu pack_h16.
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <unistd.h>

#include "aem_common.cuh"
#include "tc_common.cuh"

namespace aem {
namespace h16 {

using namespace aemtc;

constexpr int NT = 256;
constexpr int NBUF = 4;                 // a two-layer issue group (in_proj heads, linear1 chunks) holds 4 chunks at once
constexpr int WBUF_BYTES = 12288;       // one N = 96, K = 64 image; 16 KB images (N*K = 8192 fp16) stream as two 8 KB k-halves
constexpr int KV_PITCH = 116;           // [k head 0 (28) | k head 1 (28) | v head 0 (28) | v head 1 (28)] + 4: float4 access, conflict-free
static_assert((NBUF & (NBUF - 1)) == 0, "ring slots: power of two");
constexpr int NEV = 4;
constexpr int PBLK_MAX = H16_PBLK_FLOATS;
// Token-0 compaction (SURVEY 8d "exact algorithmic savings"): after the last encoder layer's attention only token 0 of a
// sample is needed, so its rows (attention output, residual stream, self state) are staged in a per-CTA scratch and
// out_proj / FFN / LayerNorm / action MLP run ONCE for the token-0 rows of several tiles (a "tail" of up to 128 rows).
constexpr int TAIL_HALF = 56;            // floats per (row, thread half): o 26 | residual 26 | self 3 | sample index (half 0)
constexpr int TAIL_ROW = 2 * TAIL_HALF;

// ---- tensor-memory column map: 256 columns per CTA ----
constexpr uint32_t A1 = 0;              // [h;x]   K = 64  : hi cols [0,32) lo cols [32,64)
constexpr uint32_t D1 = 64;             // z1 / r1 accumulators N = 128 = [64 | 64]; converted in place to A2 (K = 128, 8 x [hi8 | lo8])
constexpr uint32_t D2 = 192;            // z2 / r2 / q accumulators N = 64
constexpr uint32_t A3 = 64;             // [r*h;x] K = 64 (aliases the dead A2)
constexpr uint32_t AE = 0;              // encoder stream operand K = 64
constexpr uint32_t DIN = 64;            // in_proj accumulators, 2 heads x 96
constexpr uint32_t AATT = 64;           // attention output operand K = 64 (D_IN is consumed by then)
constexpr uint32_t DOUT = 128;          // out_proj accumulators N = 64
constexpr uint32_t DF1A = 64, DF1B = 160;   // linear1 chunks N = 96 = [48 | 48]; converted in place (K = 96, 6 x [hi8 | lo8])
constexpr uint32_t DF2 = 0;             // linear2 accumulators N = 64 (aliases the dead AE)
constexpr uint32_t DA0 = 64;            // action_mlp.0 accumulators N = 128, converted in place
constexpr uint32_t DA2 = 192;           // action_mlp.2 accumulators N = 64

// parameter block (floats, capi.cu pack_h16): every 50-vector is stored as two rows of 28 (25 + 3 zeros), one per thread
// half, so that a thread reads its 25 values as 7 float4.  LayerNorm: [layer][norm1 / norm2][half][gamma 28 | beta 28];
// ponder weights [half][28] + bias; action_mlp.4 weights [half][28] + bias.
// The biases of the MMA layers ride in the weight images: every operand keeps one padding K slot (slot 31 of half 1) at
// exactly 1.0.
constexpr int PB_LN = 0, PB_PW = 672, PB_PWB = 728, PB_A4 = 732, PB_A4B = 788;
static_assert(PB_A4B + 1 <= H16_PBLK_FLOATS, "parameter block");

constexpr size_t SMEM_BYTES = (size_t)NBUF * WBUF_BYTES + 128 * KV_PITCH * 4 + PBLK_MAX * 4 + 4 * 128 * 4 +
                              (96 + 4) * 4 + (NBUF + NEV) * 8 + 16;

constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void pair_sync(int q4) { asm volatile("bar.sync %0, 64;" ::"r"(q4 + 1) : "memory"); }

// ---- packed FP32 pairs (FFMA2 / FADD2 / FMUL2 on sm_100) ----
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, neg2(b)); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoidf_(float x) { return rcpf(1.0f + ex2f(-LOG2E * x)); }
// sigma(relu(v)) of a pair: 1 / (1 + 2^min(-v log2 e, 0))
__device__ __forceinline__ float2 sigmoid_relu2(float2 v)
{
    const float2 t = mul2(v, bc2(-LOG2E));
    const float2 d = add2(f2(ex2f(fminf(t.x, 0.0f)), ex2f(fminf(t.y, 0.0f))), bc2(1.0f));
    return f2(rcpf(d.x), rcpf(d.y));
}
// tanh of a pair: 1 - 2 / (2^(2 v log2 e) + 1)
__device__ __forceinline__ float2 tanh2(float2 v)
{
    const float2 t = mul2(v, bc2(2.0f * LOG2E));
    const float2 d = add2(f2(ex2f(t.x), ex2f(t.y)), bc2(1.0f));
    return fma2(f2(rcpf(d.x), rcpf(d.y)), bc2(-2.0f), bc2(1.0f));
}

// ---- tensor-memory loads / stores of this thread's lane: columns as float pairs ----
__device__ __forceinline__ void ld16p(uint32_t taddr, float2 *v)      // 16 columns -> 8 pairs
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y),
          "=f"(v[4].x), "=f"(v[4].y), "=f"(v[5].x), "=f"(v[5].y), "=f"(v[6].x), "=f"(v[6].y), "=f"(v[7].x), "=f"(v[7].y)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ld8p(uint32_t taddr, float2 *v)       // 8 columns -> 4 pairs
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void ld4p(uint32_t taddr, float2 *v)       // 4 columns -> 2 pairs
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void ld2p(uint32_t taddr, float2 *v)       // 2 columns -> 1 pair
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=f"(v[0].x), "=f"(v[0].y) : "r"(taddr) : "memory");
}
// the 26 leading columns of a 32-column block (25 real + the first padding column) -> 13 pairs
__device__ __forceinline__ void ld26p(uint32_t taddr, float2 *v)
{
    ld16p(taddr, v);
    ld8p(taddr + 16, v + 8);
    ld2p(taddr + 24, v + 12);
}
__device__ __forceinline__ void st8u(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void st4u(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3])
                 : "memory");
}
__device__ __forceinline__ void st1u(uint32_t taddr, uint32_t v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}

// (a.x, a.y) -> packed fp16 pair hi and the packed residual pair lo; a.x sits in the low half (even k)
__device__ __forceinline__ void split2(float2 a, uint32_t &hi, uint32_t &lo)
{
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(a.y), "f"(a.x));
    const float2 r = sub2(a, __half22float2(*reinterpret_cast<const __half2 *>(&hi)));
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r.y), "f"(r.x));
}
// ReLU + split: hi = relu(a) rounded TOWARDS ZERO, so the residual of a positive value is never negative and the ReLU
// can ride in both conversions (a negative value gives hi = 0, residual a < 0 -> lo = 0).
__device__ __forceinline__ void split2_relu(float2 a, uint32_t &hi, uint32_t &lo)
{
    asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(a.y), "f"(a.x));
    const float2 r = sub2(a, __half22float2(*reinterpret_cast<const __half2 *>(&hi)));
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r.y), "f"(r.x));
}
// Operand block of a thread: 16 columns hi at `hi_addr`, 16 columns lo 32 further; column j = K slots (2j, 2j + 1).
// Columns 0..12 are split from the 13 pairs `v`; columns 13..15 (x slots / constants) arrive already split.
// X1 (single-pass mode, aem_set_precision 3): only the hi columns exist.
template <bool X1>
__device__ __forceinline__ void st16_split(uint32_t hi_addr, const float2 *v, uint32_t h13, uint32_t h14, uint32_t h15,
                                           uint32_t l13, uint32_t l14, uint32_t l15)
{
    uint32_t h[16], l[16];
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        if (X1) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(v[i].y), "f"(v[i].x));
        else split2(v[i], h[i], l[i]);
    }
    h[13] = h13; h[14] = h14; h[15] = h15;
    st8u(hi_addr, h);
    st8u(hi_addr + 8, h + 8);
    if (!X1) {
        l[13] = l13; l[14] = l14; l[15] = l15;
        st8u(hi_addr + 32, l);
        st8u(hi_addr + 40, l + 8);
    }
}

// In-place ReLU conversion of this thread's W accumulator columns [base, base + W) into its operand block, 16 columns
// at a time: accumulators [16c, 16c + 16) -> hi columns [16c, 16c + 8), lo columns [16c + 8, 16c + 16).  Every chunk is
// read completely before it is overwritten and touches no other chunk, so the loop stays rolled (instruction-cache
// footprint) and needs no staging registers.  k-step ks of the operand = chunk ks: hi at 16 ks, lo at 16 ks + 8.
template <int W, bool X1>
__device__ __forceinline__ void relu_inplace(uint32_t base)
{
#pragma unroll 1
    for (int c = 0; c < W / 16; ++c) {
        float2 v[8];
        ld16p(base + 16 * c, v);
        tmem_wait_ld();
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (X1) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(v[i].y), "f"(v[i].x));
            else split2_relu(v[i], h[i], l[i]);
        }
        st8u(base + 16 * c, h);
        if (!X1) st8u(base + 16 * c + 8, l);
    }
}

// cadrl.py:239-274 (see lookahead.cu for the literal form).  cos / sin of rot = atan2(dy, dx) are dx / |d| and dy / |d|
// (atan2(0, 0) = 0 -> 1, 0): the same rotation to ~1 ulp without the three transcendental calls.
__device__ __forceinline__ void rotate13(const float *s, float *o)
{
    const float dx = s[5] - s[0], dy = s[6] - s[1];
    const float dg = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    const float inv = dg > 0.0f ? 1.0f / dg : 0.0f;
    const float c = dg > 0.0f ? dx * inv : 1.0f, sn = dy * inv;
    o[0] = dg;
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
// weight ring: consumer = warp 0 (issues the MMAs), producer = lane 0 of warp 1 (requests the bulk copies)
// ------------------------------------------------------------------------------------------------
// Neither role works between a group barrier and the group's first MMA.  The ring read position is a register every
// thread carries (`wn` = chunks consumed so far; each issue_t call advances it by a compile-time count, control flow is
// CTA-uniform): the issuer derives slot and mbarrier parity from it, and the producer warp, after waiting on an event like
// every other thread, releases exactly the slots consumed up to that event's commit and requests the next chunks from a
// flat table of {offset, bytes} built once per CTA -- no shared-memory hand-off and no fence on the issue path.
struct Issuer {
    int nloaded, nfree, lbuf;       // producer: chunks requested / released, ring slot of the next request
    int pf_cur, pf_end, pf_stage;   // producer: cursor into the chunk table; stage -1 = waiting for the ACT decision, -2 = done
    int tiles_left;
    int tiles_done, tail_every;     // the tail stage follows every tail_every-th tile of the CTA (and its last one)
};
constexpr int MAX_CHUNKS = 128;
// stage 0: GRU step 1, 1: GRU steps 2.., 2: encoder layers 0, 1 and the in_proj of layer 2, 3: the tail (out_proj, FFN of layer 2
// and the action MLP on the staged token-0 rows)
__device__ __forceinline__ int stage_len(int st) { return st == 0 ? 3 : (st == 1 ? 5 : (st == 2 ? 16 : 7)); }
__device__ __forceinline__ int layer_at(int st, int idx)
{
    if (st == 0) return idx == 0 ? TL_Z1S : (idx == 1 ? TL_Z2 : TL_QS);
    if (st == 1) return idx == 0 ? TL_Z1 : (idx == 1 ? TL_Z2 : (idx == 2 ? TL_R1 : (idx == 3 ? TL_R2 : TL_Q)));
    if (st == 2) return TL_ENC + idx;
    return idx < 5 ? TL_ENC + 16 + idx : (idx == 5 ? TL_A0 : TL_A2);
}
// chunk table entry: offset into the image in units of 8 halves (16 B) | bytes / 16 << 20
__device__ __noinline__ void build_chunk_table(const TcLayer *layers, uint32_t *chunks, int *stage_beg, int nparts)
{
    int n = 0;
    for (int st = 0; st < 4; ++st) {
        stage_beg[st] = n;
        for (int i = 0; i < stage_len(st); ++i) {
            const TcLayer L = layers[layer_at(st, i)];
            const int nsplit = L.bias_off;                 // h16 tables carry the k-split count in this field
            const int bytes = L.bytes / nsplit;
            for (int part = 0; part < nparts; ++part)       // hi images, then lo images (single-pass mode: hi only)
                for (int sub = 0; sub < nsplit; ++sub)
                    chunks[n++] = (uint32_t)(((part ? L.off_lo : L.off_hi) + sub * (bytes / 2)) >> 3) | ((uint32_t)(bytes >> 4) << 20);
        }
    }
    stage_beg[4] = n;
}
__device__ __forceinline__ void set_stage(Issuer *s, const int *stage_beg, int st)
{
    s->pf_stage = st;
    if (st >= 0) { s->pf_cur = stage_beg[st]; s->pf_end = stage_beg[st + 1]; }
}
// request as many chunks as the ring has room for (producer thread only)
__device__ __noinline__ void prefetch(Issuer *s, const uint32_t *chunks, const int *stage_beg, uint8_t *wbuf, uint64_t *wfull,
                                         const __half *img)
{
    int stage = s->pf_stage;
    int nloaded = s->nloaded;
    int room = NBUF - (nloaded - s->nfree);
    if (stage < 0 || room <= 0) return;
    int cur = s->pf_cur, end = s->pf_end, lbuf = s->lbuf, tiles_left = s->tiles_left;
    while (room > 0) {
        const uint32_t c = chunks[cur];
        const uint32_t bytes = (c >> 20) << 4;
        mbar_expect_tx(&wfull[lbuf], bytes);
        bulk_g2s(wbuf + (size_t)lbuf * WBUF_BYTES, img + (size_t)(c & 0xFFFFFu) * 8, bytes, &wfull[lbuf]);
        ++nloaded; --room;
        if (++lbuf == NBUF) lbuf = 0;
        if (++cur == end) {
            bool next_tile = false;
            if (stage == 2) {               // the tile's own work is streamed: a tail now (same rule as the kernel's), else the next tile
                ++s->tiles_done;
                if (s->tiles_done % s->tail_every == 0 || tiles_left == 0) { stage = 3; cur = stage_beg[3]; end = stage_beg[4]; }
                else next_tile = true;
            } else if (stage == 3) {
                if (tiles_left > 0) next_tile = true;
                else { stage = -2; break; }
            } else { stage = -1; break; }   // a GRU step: the halting vote decides what comes next
            if (next_tile) { --tiles_left; stage = 0; cur = stage_beg[0]; end = stage_beg[1]; }
        }
    }
    s->nloaded = nloaded; s->lbuf = lbuf; s->pf_cur = cur; s->pf_end = end; s->pf_stage = stage; s->tiles_left = tiles_left;
}
// Fast path of the producer, executed by the whole producer warp: the `room` chunks to request do not end a stage, so no
// cursor logic is needed, and lane i requests chunk i (its table word, expect_tx and bulk copy computed in parallel with
// the other lanes').  `nf` = chunks the consumer has finished with.  (Round-2 timelines: the same work as dependent scalar
// code on ONE thread took ~770 cycles, by which its warp arrived late at every group barrier; the general walker above
// costs ~200 dependent instructions per call and remains for the stage ends.)
__device__ __forceinline__ void produce_fast(Issuer *s, int nf, int lane, const uint32_t *chunks, const int *stage_beg,
                                             uint8_t *wbuf, uint64_t *wfull, const __half *img, uint32_t ring_s, uint32_t wfull_s)
{
    const int stage = s->pf_stage, nloaded = s->nloaded, cur = s->pf_cur, end = s->pf_end;
    __syncwarp();                                   // every lane has read the state before lane 0 rewrites it
    const int room = NBUF - (nloaded - nf);
    if (stage < 0 || room <= 0) {
        if (lane == 0) s->nfree = nf;
        return;
    }
    if (cur + room >= end) {
        if (lane == 0) { s->nfree = nf; prefetch(s, chunks, stage_beg, wbuf, wfull, img); }
        return;
    }
    if (lane < room) {
        const uint32_t c = chunks[cur + lane];
        const uint32_t bytes = (c >> 20) << 4;
        const uint32_t slot = (uint32_t)(nloaded + lane) & (uint32_t)(NBUF - 1);
        const uint32_t bar = wfull_s + 8u * slot;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ring_s + slot * (uint32_t)WBUF_BYTES),
                     "l"(img + (size_t)(c & 0xFFFFFu) * 8), "r"(bytes), "r"(bar)
                     : "memory");
    }
    if (lane == 0) {
        s->nfree = nf;
        s->nloaded = nloaded + room;
        s->lbuf = (nloaded + room) & (NBUF - 1);
        s->pf_cur = cur + room;
    }
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

#ifdef AEM_H16_TIMELINE
#define TL_ARG , long long &tl_first
#define TL_PASS , tl_first
#else
#define TL_ARG
#define TL_PASS
#endif
// One layer: KS k-steps (16 K values = 8 operand columns each) per pass.  Operand column of k-step ks:
//   a_base + HSTRIDE * (ks / KSH) + KSTR * (ks % KSH), lo part LO columns further.  SPECIAL: step-1 images use the
//   k-steps {1, 3} of A1 (the x slots).
template <int N, int KS, int KSH, int HSTRIDE, int KSTR, int LO, bool SPECIAL, int SPLIT, bool X1>
__device__ __forceinline__ void issue_t(uint32_t &wn, bool issuer, uint64_t *wfull, uint32_t ring, uint32_t tm, uint32_t a_base,
                                        uint32_t d, bool accumulate TL_ARG)
{
    constexpr uint32_t LBO = (uint32_t)N * 16;
    constexpr uint64_t DHI = ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // F16 x F16 -> F32
    constexpr int KSS = KS / SPLIT;                      // k-steps per streamed chunk
    static_assert(KS % SPLIT == 0, "split");
    // Every thread calls this (uniform control flow) and advances its copy of the ring read position `wn` = chunks consumed
    // so far (slot = wn % NBUF, mbarrier parity = (wn / NBUF) & 1); only the issuer warp waits for the chunks and issues.
    constexpr uint32_t NCH = (uint32_t)SPLIT * (X1 ? 1u : 2u);
    if (issuer) {
    const bool leader = elect_one();
#pragma unroll
    for (int part = 0; part < (X1 ? 1 : 2); ++part) {    // part 0: W_hi (A_hi and A_lo passes), part 1: W_lo (A_hi pass)
#pragma unroll
        for (int sub = 0; sub < SPLIT; ++sub) {
            const uint32_t cn = wn + (uint32_t)(part * SPLIT + sub);
            const uint32_t cbuf = cn & (uint32_t)(NBUF - 1);
            mbar_wait(&wfull[cbuf], (cn / (uint32_t)NBUF) & 1u);         // all lanes, uniform
#ifdef AEM_H16_TIMELINE
            if (tl_first == 0) tl_first = clock64();
#endif
            const uint32_t base = ((ring + cbuf * (uint32_t)WBUF_BYTES) & 0x3FFFFu) >> 4;
#pragma unroll
            for (int kk = 0; kk < KSS; ++kk) {
                const int ks = sub * KSS + kk;
                const uint32_t acol = SPECIAL ? (uint32_t)(8 * (2 * ks + 1))
                                              : (uint32_t)(HSTRIDE * (ks / KSH) + KSTR * (ks % KSH));
                const uint64_t bd = DHI | (uint64_t)(base + kk * ((2 * LBO) >> 4));
                if (leader) {
                    if (part == 0) {
                        mma_f16_ts(tm + d, tm + a_base + acol, bd, IDESC, (ks == 0 && !accumulate) ? 0u : 1u);
                        if (!X1) mma_f16_ts(tm + d, tm + a_base + acol + LO, bd, IDESC, 1u);
                    } else {
                        mma_f16_ts(tm + d, tm + a_base + acol, bd, IDESC, 1u);
                    }
                }
            }
        }
    }
    }
    wn += NCH;
}
// shorthand for the three operand layouts
// (images larger than one ring slot -- N * K = 8192 -- are streamed as two k-halves: SPLIT = 2)
#define ISSUE_K64(N, a, d, acc) issue_t<N, 4, 4, 0, 8, 32, false, ((N) * 64 > 6144 ? 2 : 1), X1>(wn, warp == 0, wfull, ring, tm, a, d, acc TL_PASS)
#define ISSUE_K64S(N, a, d) issue_t<N, 2, 4, 0, 8, 32, true, 1, X1>(wn, warp == 0, wfull, ring, tm, a, d, false TL_PASS)        /* step-1 images   */
#define ISSUE_K128(N, a, d) issue_t<N, 8, 4, 64, 16, 8, false, 2, X1>(wn, warp == 0, wfull, ring, tm, a, d, false TL_PASS)   /* in-place 2 x 4 x [8|8] */
#define ISSUE_K96(N, a, d, acc) issue_t<N, 6, 3, 48, 16, 8, false, 1, X1>(wn, warp == 0, wfull, ring, tm, a, d, acc TL_PASS) /* in-place 2 x 3 x [8|8] */

// X1: the single-pass mode (A_hi W_hi only; opt-in, own tolerance: tests/test_gpu_parity.py)
template <bool X1>
__global__ void __launch_bounds__(NT, 2) lookahead_h16_kernel(const LookaheadArgs args)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *wbuf = smem_raw;
    float *kv = reinterpret_cast<float *>(wbuf + (size_t)NBUF * WBUF_BYTES);
    float *pblk = kv + 128 * KV_PITCH;
    float *sx = pblk + PBLK_MAX;                                 // [4][128] exchange
    int *s_need = reinterpret_cast<int *>(sx + 4 * 128);         // [3 steps][32 samples]: the sample goes on after this step
    uint64_t *wfull = reinterpret_cast<uint64_t *>(s_need + 96 + 4);
    uint64_t *evb = wfull + NBUF;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(evb + NEV);
    __shared__ Issuer s_iss;
    __shared__ uint32_t s_chunks[MAX_CHUNKS];
    __shared__ int s_stage_beg[5];

    const NetParams &net = args.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    for (int i = tid; i < net.h16_pblk_floats; i += NT) pblk[i] = net.h16_pblk[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_slot;
    const uint32_t tl = tm + lane_base;
    const uint32_t ring = smem_u32(wbuf);
    if (tid == 32) {
        build_chunk_table(net.h16_layers_d, s_chunks, s_stage_beg, X1 ? 1 : 2);
        s_iss.nloaded = s_iss.nfree = s_iss.lbuf = 0;
        s_iss.pf_cur = s_iss.pf_end = 0; s_iss.tiles_left = 0;
        s_iss.tiles_done = 0; s_iss.tail_every = max(1, 128 / S);
        s_iss.pf_stage = -2;
        if ((int)blockIdx.x < ntiles) {
            s_iss.tiles_left = (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x;
            set_stage(&s_iss, s_stage_beg, 0);
            prefetch(&s_iss, s_chunks, s_stage_beg, wbuf, wfull, net.h16_img);
        }
    }
    __syncthreads();
    int nev = 0;
    uint32_t wn = 0;            // weight chunks consumed so far: the ring read position, carried by every thread (issue_t)
    // the two CTAs of an SM run identical work: offset the second one so that its epilogues fall into the first one's
    // MMA / commit waits instead of colliding with its epilogues
    if ((int)blockIdx.x >= (int)gridDim.x / 2 && args.stagger_ns > 0) __nanosleep((unsigned)args.stagger_ns);

#define GROUP_BARRIER()                     \
    do {                                    \
        tmem_wait_st();                     \
        tc_fence_before();                  \
        __syncthreads();                    \
        if (warp == 0) tc_fence_after();    \
    } while (0)
#define COMMIT(e) mma_commit(&evb[(e) % NEV])
#define WAIT_EVENT_POLL(e)                                              \
    do {                                                                \
        mbar_wait(&evb[(e) % NEV], (uint32_t)(((e) / NEV) & 1));        \
        tc_fence_after();                                               \
    } while (0)
#define WAIT_EVENT(e) WAIT_EVENT_POLL(e)
#ifdef AEM_H16_TIMELINE
    long long tl_first = 0;     // clock64 when the group's first weight chunk was there
#define ISSUE_BEGIN() tl_first = 0; MARK()
#define ISSUE_END() do { if (args.dbg && tl_slot == 0 && blockIdx.x == 0 && tile == tl_tile && tl_n < 500) { \
        args.dbg[tl_n++] = (tl_first << 12) | 1; args.dbg[tl_n++] = (clock64() << 12) | 2; } } while (0)
#else
#define ISSUE_BEGIN() do { } while (0)
#define ISSUE_END() do { } while (0)
#endif
    // the MMAs are issued by the elected lane of warp 0 (issue_t); the same thread commits them to the group's event
#define ISSUE_COMMIT(e)                                       \
    do {                                                      \
        if (warp == 0) { if (elect_one()) COMMIT(e); __syncwarp(); } \
    } while (0)
// after an event has been waited on: the producer warp releases the ring slots the event's MMAs read (`nf` = chunks
// consumed when the event was committed; every thread carries that count in `wn`) and requests the next chunks
#define PRODUCE_N(nf)                                                                 \
    do {                                                                              \
        if (warp == 1) {                                                              \
            produce_fast(&s_iss, (int)(nf), lane, s_chunks, s_stage_beg, wbuf, wfull, net.h16_img, ring, smem_u32(wfull)); \
            __syncwarp();                                                             \
        }                                                                             \
    } while (0)
#define PRODUCE(e) PRODUCE_N(wn)
    const float *PB = pblk;
#if defined(AEM_H16_TIMELINE)
    // clock64 stamps of three threads of CTA 0 (issuer warp, producer warp, last warp) over one steady-state tile
    int tl_n = 0;
    const int tl_slot = tid == 0 ? 0 : (tid == 32 ? 1 : (tid == 255 ? 2 : -1));
    const int tl_tile = (int)args.dbg_tile * (int)gridDim.x;     // the CTA's dbg_tile-th tile (AEM_H16_TL_TILE, default 2)
#define MARK() do { if (args.dbg && tl_slot >= 0 && blockIdx.x == 0 && tile == tl_tile && tl_n < 500) { \
        args.dbg[tl_slot * 512 + tl_n++] = (clock64() << 12) | (long long)__LINE__; } } while (0)
#elif defined(AEM_H16_MARKS)
#define MARK() do { if (args.dbg && (tid == 0 || tid == 32)) { ((volatile long long *)args.dbg)[2 * blockIdx.x + (tid == 32)] = __LINE__ + 10000ll * tile; __threadfence_system(); } } while (0)
#else
#define MARK() do { } while (0)
#endif

    const int tail_every = max(1, 128 / S);
    float *tail = args.tail + (size_t)blockIdx.x * 128 * TAIL_ROW;
    int staged = 0, ltile = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int s0 = tile * S;
        const int nsamp = min(S, args.total_samples - s0);
        const int nrows = nsamp * T;

        // per-thread state: this thread's 25 hidden columns as 13 pairs (slot 25 is a padding column and stays 0)
        float2 ACC[13], hn[13], z[13];
        float selfs[3], xh0 = 0.0f;
        uint32_t xph[3], xpl[3];                         // x slots 1..6 of this half as split pairs (columns 13..15 of the operand)
        float P = 0.0f;
        int samp = -1, tok = 0, tact = T;                 // tact: real tokens of this row's sample ('mixed' scenes), else T
#pragma unroll
        for (int j = 0; j < 13; ++j) { ACC[j] = bc2(0.0f); hn[j] = bc2(0.0f); }
        MARK();

        // ---------------- phase 0: inputs -> A1 = [h = 0 ; x] ----------------
        {
            float o[XD];
#pragma unroll
            for (int k = 0; k < XD; ++k) o[k] = 0.0f;
#pragma unroll
            for (int i = 0; i < 3; ++i) selfs[i] = 0.0f;
            if (row < nrows) {
                // row order inside a tile: the token-0 (robot) rows of all samples first, then the human tokens sample by
                // sample -- the last encoder layer and the action MLP only need token 0, i.e. the leading warps
                if (row < nsamp) { samp = row; tok = 0; }
                else { const int r2 = row - nsamp; samp = r2 / (T - 1); tok = 1 + r2 - samp * (T - 1); }
                const int gs = s0 + samp;
                if (args.mode == 0) {
                    const int b = gs / args.A, a = gs - b * args.A;
                    if (args.n_active) tact = 1 + min(max(args.n_active[b], 1), args.n);
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
                    if (tok >= tact) {               // padding row: not a token of its sample (no vote, no key, no value)
#pragma unroll
                        for (int k = 0; k < XD; ++k) o[k] = 0.0f;
                    }
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
                for (int i = 0; i < 3; ++i) selfs[i] = half ? o[3 + i] : o[i];
                if (args.mode != 2 && tok == 0) {
                    o[6] = 0.0f; o[7] = 0.0f; o[8] = o[4]; o[9] = o[5]; o[10] = o[3]; o[11] = 0.0f; o[12] = o[3];
                }
            }
            float xh[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) xh[i] = half ? (i < 6 ? o[7 + (i < 6 ? i : 0)] : 1.0f) : o[i];   // slot 31 of half 1 = 1 (bias)
            xh0 = xh[0];
#pragma unroll
            for (int i = 0; i < 3; ++i) split2(f2(xh[1 + 2 * i], xh[2 + 2 * i]), xph[i], xpl[i]);
            // A1 block: columns 0..11 = h = 0, column 12 = (h24 = 0, x0), columns 13..15 = x1..x6
            const uint32_t a1 = tl + A1 + 16 * half;
            float2 v[13];
#pragma unroll
            for (int j = 0; j < 13; ++j) v[j] = bc2(0.0f);
            v[12].y = xh0;
            st16_split<X1>(a1, v, xph[0], xph[1], xph[2], xpl[0], xpl[1], xpl[2]);
            if (tid < 96) s_need[tid] = 0;
        }

        // ---------------- phase 1: GRU + adaptive computation time ----------------
        const int max_steps = net.act_fixed ? net.act_steps : 3;
        bool active = samp >= 0 && tok < tact;
        int cnt = 0;
        MARK(); GROUP_BARRIER();
#pragma unroll 1
        for (int step = 1; step <= max_steps; ++step) {
            const bool kfull = step > 1;
            // ---- the two gates share one copy of the code (instruction-cache footprint): gate 0 = z, gate 1 = r
            //      layer 1 -> in place A2 (K = 128), layer 2 -> z, or r and [r*h ; x] -> A3.  The ring delivers the weight
            //      images in the order Z1, Z2, R1, R2; step 1 has h = 0 and needs z only.
            //      The group barrier of gate 0 is the halting vote of the previous step (or the barrier above).
            // (fully unrolled: with the loop rolled, ptxas 12.9 assigned the tcgen05.ld destinations of the second iteration
            //  to registers that still held z of the first one -- round-2 finding, profiles/r02_lookahead_h16_ncu.md)
#pragma unroll
            for (int gate = 0; gate < 2; ++gate) {
                if (gate == 1 && !kfull) break;
                if (gate) { MARK(); GROUP_BARRIER(); }
                const int e_g1 = nev++;
                ISSUE_BEGIN();
                if (kfull) ISSUE_K64(128, A1, D1, false);
                else ISSUE_K64S(128, A1, D1);
                ISSUE_COMMIT(e_g1);
                ISSUE_END();
                MARK(); WAIT_EVENT(e_g1); MARK();
                PRODUCE(e_g1); MARK();
                relu_inplace<64, X1>(tl + D1 + 64 * half);
                MARK(); GROUP_BARRIER();
                const int e_g2 = nev++;
                ISSUE_BEGIN();
                ISSUE_K128(64, D1, D2);
                ISSUE_COMMIT(e_g2);
                ISSUE_END();
                MARK(); WAIT_EVENT(e_g2); MARK();
                PRODUCE(e_g2); MARK();
                float2 v[13];
                ld26p(tl + D2 + 32 * half, v);
                tmem_wait_ld();
                if (gate == 0) {
#pragma unroll
                    for (int j = 0; j < 13; ++j) z[j] = sigmoid_relu2(v[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 13; ++j) v[j] = mul2(sigmoid_relu2(v[j]), hn[j]);
                    v[12].y = xh0;
                    st16_split<X1>(tl + A3 + 16 * half, v, xph[0], xph[1], xph[2], xpl[0], xpl[1], xpl[2]);
                }
            }
            // ---- q, h = (1 - z) h + z q
            MARK(); GROUP_BARRIER();
            const int e_q = nev++;
            ISSUE_BEGIN();
            if (kfull) ISSUE_K64(64, A3, D2, false);
            else ISSUE_K64S(64, A1, D2);
            ISSUE_COMMIT(e_q);
            ISSUE_END();
            MARK(); WAIT_EVENT(e_q); MARK();
            PRODUCE(e_q); MARK();
            {
                float2 v[13];
                ld26p(tl + D2 + 32 * half, v);
                tmem_wait_ld();
                const float4 *pw = reinterpret_cast<const float4 *>(PB + PB_PW + 28 * half);
                float2 pd = bc2(0.0f);
#pragma unroll
                for (int j = 0; j < 13; ++j) {
                    const float2 qv = tanh2(v[j]);
                    hn[j] = fma2(z[j], sub2(qv, hn[j]), hn[j]);
                    if (j == 12) hn[12].y = 0.0f;                  // padding column
                    const float4 w4 = pw[j >> 1];
                    pd = fma2(hn[j], (j & 1) ? f2(w4.z, w4.w) : f2(w4.x, w4.y), pd);
                    v[j] = hn[j];
                }
                v[12].y = xh0;
                st16_split<X1>(tl + A1 + 16 * half, v, xph[0], xph[1], xph[2], xpl[0], xpl[1], xpl[2]);
                if (!net.act_fixed) {
                    sx[half * 128 + row] = pd.x + pd.y;
                    pair_sync(q4);
                    if (active) {
                        const float p = sigmoidf_(sx[row] + sx[128 + row] + PB[PB_PWB]);
                        P += p;
#pragma unroll
                        for (int j = 0; j < 13; ++j) ACC[j] = fma2(bc2(p), hn[j], ACC[j]);
                        cnt = step;
                        if (half == 0 && P < 0.95f) s_need[(step - 1) * 32 + samp] = 1;
                    }
                } else {
                    cnt = step;
                }
            }
            // ---- halting vote (components.py:132-134, per (env, action) sample): one barrier, which is also the group
            //      barrier of the next step's first gate
            MARK();
            tmem_wait_st();
            tc_fence_before();
            int cont;
            if (net.act_fixed) {
                __syncthreads();
                cont = step < max_steps;
            } else {
                const int any = __syncthreads_or(active && P < 0.95f);
                cont = any && step < max_steps;
                active = active && s_need[(step - 1) * 32 + samp] != 0;      // (padding rows stay inactive)
            }
            if (warp == 0) tc_fence_after();
            if (tid == 32 && s_iss.pf_stage == -1) {
                int ns;
                if (cont) ns = 1;
                else if (args.mode == 2) {
                    if (s_iss.tiles_left > 0) { --s_iss.tiles_left; ns = 0; } else ns = -2;
                } else ns = 2;
                set_stage(&s_iss, s_stage_beg, ns);
                prefetch(&s_iss, s_chunks, s_stage_beg, wbuf, wfull, net.h16_img);
            }
            if (!cont) break;
        }

        float2 xres[13];
        {
            const float rc = rcpf((float)max(cnt, 1));
#pragma unroll
            for (int j = 0; j < 13; ++j) xres[j] = net.act_fixed ? hn[j] : (samp >= 0 ? mul2(ACC[j], bc2(rc)) : bc2(0.0f));
        }
        if (args.mode == 2) {
            if (row < nrows) {
                float *dst = args.gru_out + ((size_t)(s0 + samp) * T + tok) * HID + 25 * half;
#pragma unroll
                for (int j = 0; j < 12; ++j) { dst[2 * j] = xres[j].x; dst[2 * j + 1] = xres[j].y; }
                dst[24] = xres[12].x;
                if (args.steps && half == 0 && tok == 0) args.steps[s0 + samp] = (uint8_t)cnt;
            }
            __syncthreads();
            continue;
        }
        // constant columns 13..15 of the encoder operands: zeros and the bias slot (slot 31 of half 1 = 1.0)
        const uint32_t one_hi = half ? 0x3C000000u : 0u;
        {
            st16_split<X1>(tl + AE + 16 * half, xres, 0u, 0u, one_hi, 0u, 0u, 0u);
        }

        // ---------------- phase 2: 3 x post-norm encoder layer ----------------
        bool tail_ran = false;
        int tail_gs = -1, tail_rows = 0;
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
            MARK(); GROUP_BARRIER();
            const int e_in = nev++;
            ISSUE_BEGIN();
#pragma unroll 1
            for (int hd = 0; hd < 2; ++hd) ISSUE_K64(96, AE, DIN + 96 * hd, false);
            ISSUE_COMMIT(e_in);
            ISSUE_END();
            MARK(); WAIT_EVENT(e_in); MARK();
            PRODUCE(e_in); MARK();
            // exact saving: the last layer needs k / v of every token but q, out_proj, FFN and LayerNorm of
            // token 0 only: warps without a token-0 row skip those epilogues (their MMA rows compute garbage nobody reads)
            float2 o[13];
            {
                // q stays in registers (28 slots, the last 3 zero) scaled by log2(e) / sqrt(head_dim = 25); k and v of this
                // head go to shared memory as 7 float4 each: row = [k head 0 (28) | k head 1 (28) | v head 0 (28) | v head 1
                // (28)].  Accumulator columns 25..31 of every block are padding neurons and exactly zero.
                float2 qh[14];
                {
                    const uint32_t din = tl + DIN + 96 * half;
                    float4 *dst = reinterpret_cast<float4 *>(kv + row * KV_PITCH + 28 * half);
                    float2 t[14];
                    ld16p(din + 32, t); ld8p(din + 48, t + 8); ld4p(din + 56, t + 12);          // k
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 7; ++c) dst[c] = make_float4(t[2 * c].x, t[2 * c].y, t[2 * c + 1].x, t[2 * c + 1].y);
                    ld16p(din + 64, t); ld8p(din + 80, t + 8); ld4p(din + 88, t + 12);          // v
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 7; ++c) dst[14 + c] = make_float4(t[2 * c].x, t[2 * c].y, t[2 * c + 1].x, t[2 * c + 1].y);
                    ld16p(din, qh); ld8p(din + 16, qh + 8); ld4p(din + 24, qh + 12);            // q
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 14; ++c) qh[c] = mul2(qh[c], bc2(0.2f * LOG2E));
                }
                tc_fence_before();
                __syncthreads();       // every thread has consumed D_IN and published k / v
                float2 oo[14];
#pragma unroll
                for (int j = 0; j < 14; ++j) oo[j] = bc2(0.0f);
                if (samp >= 0 && (l < 2 || tok == 0)) {        // last layer: the query of token 0 only
                    // one pass over the T keys of the sample with a running maximum (streaming softmax, log2 domain);
                    // key 0 is the sample's token-0 row, keys 1.. are its block of human rows
                    const float *kb = kv + samp * KV_PITCH + 28 * half;
                    float m, sum = 1.0f;
                    {
                        const float4 *kp = reinterpret_cast<const float4 *>(kb);
                        float2 d0 = bc2(0.0f), d1 = bc2(0.0f);
#pragma unroll
                        for (int c = 0; c < 7; ++c) {
                            const float4 k4 = kp[c];
                            d0 = fma2(qh[2 * c], f2(k4.x, k4.y), d0);
                            d1 = fma2(qh[2 * c + 1], f2(k4.z, k4.w), d1);
                        }
                        d0 = add2(d0, d1);
                        m = d0.x + d0.y;
#pragma unroll
                        for (int c = 0; c < 7; ++c) {
                            const float4 v4 = kp[14 + c];
                            oo[2 * c] = f2(v4.x, v4.y);
                            oo[2 * c + 1] = f2(v4.z, v4.w);
                        }
                    }
                    kb = kv + (nsamp + samp * (T - 1)) * KV_PITCH + 28 * half;
                    for (int j = 1; j < tact; ++j, kb += KV_PITCH) {
                        const float4 *kp = reinterpret_cast<const float4 *>(kb);
                        float2 d0 = bc2(0.0f), d1 = bc2(0.0f);
#pragma unroll
                        for (int c = 0; c < 7; ++c) {
                            const float4 k4 = kp[c];
                            d0 = fma2(qh[2 * c], f2(k4.x, k4.y), d0);
                            d1 = fma2(qh[2 * c + 1], f2(k4.z, k4.w), d1);
                        }
                        d0 = add2(d0, d1);
                        const float sc = d0.x + d0.y;
                        const float mn = fmaxf(m, sc);
                        const float rescale = ex2f(m - mn), p = ex2f(sc - mn);
                        m = mn;
                        sum = fmaf(sum, rescale, p);
                        const float2 r2 = bc2(rescale), p2 = bc2(p);
#pragma unroll
                        for (int c = 0; c < 7; ++c) {
                            const float4 v4 = kp[14 + c];
                            oo[2 * c] = fma2(p2, f2(v4.x, v4.y), mul2(oo[2 * c], r2));
                            oo[2 * c + 1] = fma2(p2, f2(v4.z, v4.w), mul2(oo[2 * c + 1], r2));
                        }
                    }
                    const float2 inv = bc2(rcpf(sum));
#pragma unroll
                    for (int c = 0; c < 13; ++c) oo[c] = mul2(oo[c], inv);
                }
#pragma unroll
                for (int c = 0; c < 13; ++c) o[c] = oo[c];
                o[12].y = 0.0f;
            }
            int nlive = 128;                                // rows that are real in the rest of this layer
            if (l == 2) {
                // ---- stage the token-0 rows; every tail_every-th tile (and the CTA's last one) the tail runs on all staged rows
                if (samp >= 0 && tok == 0) {
                    float *dst = tail + (size_t)(staged + samp) * TAIL_ROW + half * TAIL_HALF;
#pragma unroll
                    for (int j = 0; j < 13; ++j) {
                        *reinterpret_cast<float2 *>(dst + 2 * j) = o[j];
                        *reinterpret_cast<float2 *>(dst + 26 + 2 * j) = xres[j];
                    }
                    dst[52] = selfs[0]; dst[53] = selfs[1]; dst[54] = selfs[2];
                    if (half == 0) {
                        dst[55] = __int_as_float(s0 + samp);
                        if (args.steps) args.steps[s0 + samp] = (uint8_t)cnt;
                    }
                }
                staged += nsamp;
                ++ltile;
                tail_ran = (ltile % tail_every == 0) || (tile + (int)gridDim.x >= ntiles);
                if (!tail_ran) break;
                __syncthreads();                            // the staged rows of this tile are visible; k / v are consumed
                nlive = staged;
                staged = 0;
                {
                    // coalesced copy of the staged rows into the (now free) k / v buffer, then every thread takes its row:
                    // per-thread row loads straight from global memory serialised on their L2 round trips (~18 k cycles per tail)
                    const float4 *g4 = reinterpret_cast<const float4 *>(tail);
                    float4 *s4 = reinterpret_cast<float4 *>(kv);
                    const int n4 = nlive * (TAIL_ROW / 4);
                    float4 t[14];
#pragma unroll
                    for (int i = 0; i < 14; ++i) {
                        const int e = tid + NT * i;
                        t[i] = e < n4 ? __ldcg(g4 + e) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    }
#pragma unroll
                    for (int i = 0; i < 14; ++i) {
                        const int e = tid + NT * i;
                        if (e < 128 * (TAIL_ROW / 4)) s4[(e / (TAIL_ROW / 4)) * (KV_PITCH / 4) + e % (TAIL_ROW / 4)] = t[i];
                    }
                    __syncthreads();
                    const float *src = kv + row * KV_PITCH + half * TAIL_HALF;
#pragma unroll
                    for (int j = 0; j < 13; ++j) {
                        o[j] = *reinterpret_cast<const float2 *>(src + 2 * j);
                        xres[j] = *reinterpret_cast<const float2 *>(src + 26 + 2 * j);
                    }
#pragma unroll
                    for (int i = 0; i < 3; ++i) selfs[i] = src[52 + i];
                    tail_gs = row < nlive ? __float_as_int(kv[row * KV_PITCH + 55]) : -1;
                }
            }
            const bool live = q4 * 32 < nlive;
            if (live) st16_split<X1>(tl + AATT + 16 * half, o, 0u, 0u, one_hi, 0u, 0u, 0u);
            // ---- out_proj + residual + LayerNorm1
            MARK(); GROUP_BARRIER();
            const int e_out = nev++;
            ISSUE_BEGIN();
            ISSUE_K64(64, AATT, DOUT, false);
            ISSUE_COMMIT(e_out);
            ISSUE_END();
            MARK(); WAIT_EVENT(e_out); MARK();
            PRODUCE(e_out); MARK();
            // pass 0: out_proj + residual + LayerNorm1; pass 1: FFN + residual + LayerNorm2 (one copy of the LayerNorm code)
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 1) {
                    // ---- FFN in two hidden chunks, converted in place; linear2 accumulates the two K = 96 halves
                    MARK(); GROUP_BARRIER();
                    const int e_f1 = nev;          // chunk c completes on event e_f1 + c
                    const int e_f2 = nev + 2;
                    nev += 3;
                    ISSUE_BEGIN();
                    ISSUE_K64(96, AE, DF1A, false);
                    ISSUE_COMMIT(e_f1);
                    const uint32_t wn_f1a = wn;     // chunks covered by event e_f1 (the producer may release exactly those)
                    ISSUE_K64(96, AE, DF1B, false);
                    ISSUE_COMMIT(e_f1 + 1);
                    const uint32_t wn_f1b = wn;
                    ISSUE_END();
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        MARK(); WAIT_EVENT(e_f1 + c); MARK();
                        PRODUCE_N(c ? wn_f1b : wn_f1a); MARK();
                        if (live) relu_inplace<48, X1>(tl + (c ? DF1B : DF1A) + 48 * half);
                        MARK(); GROUP_BARRIER();
                        if (warp == 0 && c == 0) WAIT_EVENT_POLL(e_f1 + 1);   // F1B reads AE, which D_F2 aliases: complete before linear2 starts
                        ISSUE_BEGIN();
                        ISSUE_K96(64, c ? DF1B : DF1A, DF2, c != 0);
                        if (c) ISSUE_COMMIT(e_f2);
                        ISSUE_END();
                    }
                    MARK(); WAIT_EVENT(e_f2); MARK();
                    PRODUCE(e_f2); MARK();
                }
                if (!live) continue;
                float2 s2 = bc2(0.0f);
                {
                    float2 v[13];
                    ld26p(tl + (pass ? DF2 : DOUT) + 32 * half, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 13; ++j) { xres[j] = add2(xres[j], v[j]); s2 = add2(s2, xres[j]); }
                }
                const float s = s2.x + s2.y;                    // the padding column is zero
                const float mh = s * 0.04f;
                float2 ss2 = bc2(0.0f);
#pragma unroll
                for (int j = 0; j < 13; ++j) {
                    float2 dd = sub2(xres[j], bc2(mh));
                    if (j == 12) dd.y = 0.0f;
                    ss2 = fma2(dd, dd, ss2);
                }
                const float ss = ss2.x + ss2.y;
                sx[half * 128 + row] = s;
                sx[256 + half * 128 + row] = ss;
                tc_fence_before();
                pair_sync(q4);      // pass 1 also: the partner has finished reading its D_F2 columns (AE overwrites them)
                const float so = sx[(half ^ 1) * 128 + row], sso = sx[256 + (half ^ 1) * 128 + row];
                const float mean = (s + so) * 0.02f;
                const float dm = (s - so) * 0.04f;
                const float rstd = rsqrtf((ss + sso + dm * dm * 12.5f) * 0.02f + 1e-5f);
                const float4 *g4 = reinterpret_cast<const float4 *>(PB + PB_LN + (((l * 2 + pass) * 2 + half) * 56));
                float2 a[13];
#pragma unroll
                for (int j = 0; j < 13; ++j) {
                    const float4 gg = g4[j >> 1], bb = g4[7 + (j >> 1)];
                    const float2 t = mul2(sub2(xres[j], bc2(mean)), bc2(rstd));
                    xres[j] = fma2(t, (j & 1) ? f2(gg.z, gg.w) : f2(gg.x, gg.y), (j & 1) ? f2(bb.z, bb.w) : f2(bb.x, bb.y));
                    a[j] = xres[j];
                }
                uint32_t h13 = 0u, l13 = 0u;
                if (pass == 1 && l == 2) {                                    // joint = [self_state ; env]
                    a[12].y = selfs[0];
                    split2(f2(selfs[1], selfs[2]), h13, l13);
                }
                st16_split<X1>(tl + AE + 16 * half, a, h13, 0u, one_hi, l13, 0u, 0u);
            }
            if (l == 2) tail_rows = nlive;
        }
        if (!tail_ran) {                   // this tile's token-0 rows wait in the scratch for a later tail
            MARK();
            __syncthreads();
            continue;
        }

        // ---------------- phase 3: action_mlp ----------------
        MARK(); GROUP_BARRIER();
        const int e_a0 = nev++;
        ISSUE_BEGIN();
        ISSUE_K64(128, AE, DA0, false);
        ISSUE_COMMIT(e_a0);
        ISSUE_END();
        MARK(); WAIT_EVENT(e_a0); MARK();
        PRODUCE(e_a0); MARK();
        const bool live_a = q4 * 32 < tail_rows;  // the staged token-0 rows
        if (live_a) relu_inplace<64, X1>(tl + DA0 + 64 * half);
        MARK(); GROUP_BARRIER();
        const int e_a2 = nev++;
        ISSUE_BEGIN();
        ISSUE_K128(64, DA0, DA2);
        ISSUE_COMMIT(e_a2);
        ISSUE_END();
        MARK(); WAIT_EVENT(e_a2); MARK();
        PRODUCE(e_a2); MARK();
        if (live_a) {
            const float4 *w4 = reinterpret_cast<const float4 *>(PB + PB_A4 + 28 * half);
            float2 pd = bc2(0.0f);
            float2 v[13];
            ld26p(tl + DA2 + 32 * half, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 13; ++j) {
                const float4 ww = w4[j >> 1];
                pd = fma2(f2(fmaxf(v[j].x, 0.0f), fmaxf(v[j].y, 0.0f)), (j & 1) ? f2(ww.z, ww.w) : f2(ww.x, ww.y), pd);
            }
            sx[half * 128 + row] = pd.x + pd.y;
            pair_sync(q4);
            if (half == 0 && tail_gs >= 0) {
                const float val = sx[row] + sx[128 + row] + PB[PB_A4B];
                const int gs = tail_gs;
                if (args.net_values) args.net_values[gs] = val;
                if (args.values) args.values[gs] = __dadd_rn(args.rewards[gs], __dmul_rn(args.gamma_bar, (double)val));
            }
        }
        MARK();
        __syncthreads();
    }

#ifdef AEM_H16_MARKS
    { const int tile = -1; MARK(); }
#endif
    tmem_wait_st();
    tmem_wait_ld();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
#ifdef AEM_H16_MARKS
    { const int tile = -2; MARK(); }
#endif
#undef GROUP_BARRIER
#undef COMMIT
#undef WAIT_EVENT
#undef WAIT_EVENT_POLL
#undef ISSUE_BEGIN
#undef ISSUE_COMMIT
#undef ISSUE_END
#undef PRODUCE_N
#undef PRODUCE
}

}  // namespace h16

cudaError_t launch_lookahead_h16(aem_handle h, const LookaheadArgs &a, cudaStream_t st, bool single_pass)
{
    static bool attr_set[64] = {false};
    void (*kern)(const LookaheadArgs) = single_pass ? h16::lookahead_h16_kernel<true> : h16::lookahead_h16_kernel<false>;
    if (!attr_set[h->device & 63]) {
        cudaError_t e = cudaFuncSetAttribute(h16::lookahead_h16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)h16::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(h16::lookahead_h16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h16::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set[h->device & 63] = true;
    }
    const int ntiles = (a.total_samples + a.S - 1) / a.S;
    if (ntiles <= 0) return cudaSuccess;
    const int maxgrid = 2 * h->num_sms;        // two CTAs (two tiles in flight) per SM
    const int grid = ntiles < maxgrid ? ntiles : maxgrid;
    const size_t tail_floats = (size_t)maxgrid * 128 * h16::TAIL_ROW;      // staged token-0 rows, L2-resident (17 MB)
    if (h->tail_floats < tail_floats) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return e;
        cudaFree(h->tail_d);
        h->tail_d = nullptr; h->tail_floats = 0;
        if ((e = cudaMalloc(&h->tail_d, tail_floats * sizeof(float))) != cudaSuccess) return e;
        h->tail_floats = tail_floats;
    }
    LookaheadArgs b = a;
    b.tail = h->tail_d;
    static int stagger = -1;
    if (stagger < 0) {
        const char *e = getenv("AEM_H16_STAGGER_NS");
        stagger = e ? atoi(e) : 0;
    }
    b.stagger_ns = stagger;
#if defined(AEM_H16_TIMELINE)
    if (const char *path = getenv("AEM_H16_TIMELINE")) {       // debug build: clock64 stamps of CTA 0, third tile -> file
        long long *host = nullptr;
        cudaHostAlloc((void **)&host, 3 * 512 * sizeof(long long), cudaHostAllocMapped);
        for (int i = 0; i < 3 * 512; ++i) host[i] = 0;
        b.dbg = host;
        b.dbg_tile = getenv("AEM_H16_TL_TILE") ? atoi(getenv("AEM_H16_TL_TILE")) : 2;
        size_t smem = h16::SMEM_BYTES;
        int tl_grid = grid;
        if (getenv("AEM_H16_SOLO")) {           // one CTA per SM: the chain of a tile without a neighbour on the tensor pipe
            smem = 200 * 1024;
            tl_grid = ntiles < h->num_sms ? ntiles : h->num_sms;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        kern<<<tl_grid, h16::NT, smem, st>>>(b);
        cudaError_t e = cudaStreamSynchronize(st);
        if (FILE *f = fopen(path, "w")) {
            for (int s = 0; s < 3; ++s)
                for (int i = 0; i < 512 && host[s * 512 + i]; ++i)
                    fprintf(f, "%d %lld %lld\n", s, host[s * 512 + i] & 4095, host[s * 512 + i] >> 12);
            fclose(f);
        }
        cudaFreeHost(host);
        return e;
    }
#endif
    if (getenv("AEM_H16_WATCHDOG")) {       // debug: report where CTA 0 is stuck instead of hanging
        long long *host = nullptr;
        cudaHostAlloc((void **)&host, 16 * 1024, cudaHostAllocMapped);
        for (int i = 0; i < 2048; ++i) host[i] = 0;
        b.dbg = host;
        kern<<<grid, h16::NT, h16::SMEM_BYTES, st>>>(b);
        for (int i = 0; i < 300; ++i) {
            if (cudaStreamQuery(st) != cudaErrorNotReady) { cudaFreeHost(host); return cudaGetLastError(); }
            struct timespec ts = {0, 10000000};
            nanosleep(&ts, nullptr);
        }
        for (int c = 0; c < grid && c < 1024; ++c) {
            const long long a0 = ((volatile long long *)host)[2 * c], a1 = ((volatile long long *)host)[2 * c + 1];
            if (a0 / 10000 != -2 || a1 / 10000 != -2)
                fprintf(stderr, "h16 watchdog: CTA %d stuck: thread 0 tile %lld line %lld, thread 32 tile %lld line %lld\n", c,
                        a0 / 10000, a0 % 10000, a1 / 10000, a1 % 10000);
        }
        _exit(3);
    }
    kern<<<grid, h16::NT, h16::SMEM_BYTES, st>>>(b);
    return cudaGetLastError();
}

}  // namespace aem
