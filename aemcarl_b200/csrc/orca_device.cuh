// orca_device.cuh -- device code of the per-agent ORCA half-plane linear program (one warp per human), shared by the
// stand-alone ORCA kernel (orca.cu) and the fused environment-preparation kernel (env.cu).
//
//
// Replaces Human.act -> ORCA.predict -> rvo2.PyRVOSimulator.doStep (crowd_sim/envs/policy/orca.py:82-132, called
// for every human at crowd_sim/envs/crowd_sim.py:562-568).  The arithmetic is the published RVO2 v2.0 algorithm
// (Agent::computeNeighbors / insertAgentNeighbor / computeNewVelocity, linearProgram1-3) in FP32.
//
// EVERY FILE THAT INCLUDES THIS HEADER MUST BE COMPILED WITH -fmad=false: every FP32 operation is then an individually rounded IEEE
// add / mul / div / sqrt in the same order as oracle/aem_oracle.c (built with -ffp-contract=off), so the
// result is bit-identical to the CPU oracle.  min / max reductions across lanes are order independent.
//
// Parallel decomposition inside the warp: lane j = candidate neighbour j (<= 32 others).  Neighbour ranking,
// half-plane construction, the inner loop of linearProgram1 (intersection with lines j < i) and the projection
// lines of linearProgram3 run one lane per line; the outer incremental loops are warp-uniform.
#pragma once
#include "aem_common.cuh"

namespace aem {


#define RVO_EPSILON 0.00001f
constexpr unsigned FULL = 0xffffffffu;

struct V2 { float x, y; };
__device__ __forceinline__ V2 mk(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ V2 vadd(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 vsub(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ V2 vmul(float s, V2 a) { return mk(s * a.x, s * a.y); }
__device__ __forceinline__ V2 vdiv(V2 a, float s) { const float inv = 1.0f / s; return mk(a.x * inv, a.y * inv); }
__device__ __forceinline__ float vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float vdet(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float vabssq(V2 a) { return vdot(a, a); }
__device__ __forceinline__ float vabs(V2 a) { return sqrtf(vdot(a, a)); }
__device__ __forceinline__ V2 vnormalize(V2 a) { return vdiv(a, vabs(a)); }
__device__ __forceinline__ float sqrf(float a) { return a * a; }

struct Line { V2 point, direction; };

__device__ __forceinline__ float warp_min(float v)
{
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// linearProgram1: lines in shared memory, lane j handles the constraint of line j (< lineNo). Warp-uniform result.
__device__ __forceinline__ bool lp1(const Line *lines, int lineNo, float radius, V2 opt, bool dirOpt, V2 &result,
                                    int lane)
{
    const Line li = lines[lineNo];
    const float dotProduct = vdot(li.point, li.direction);
    const float discriminant = sqrf(dotProduct) + sqrf(radius) - vabssq(li.point);
    if (discriminant < 0.0f) return false;
    const float sqrtDiscriminant = sqrtf(discriminant);
    float tLeft = -dotProduct - sqrtDiscriminant;
    float tRight = -dotProduct + sqrtDiscriminant;
    bool fail = false;
    float candL = -INFINITY, candR = INFINITY;
    if (lane < lineNo) {
        const Line lj = lines[lane];
        const float denominator = vdet(li.direction, lj.direction);
        const float numerator = vdet(lj.direction, vsub(li.point, lj.point));
        if (fabsf(denominator) <= RVO_EPSILON) {
            if (numerator < 0.0f) fail = true;
        } else {
            const float t = numerator / denominator;
            if (denominator >= 0.0f) candR = t; else candL = t;
        }
    }
    // sequential semantics: fails as soon as a parallel line rejects or tLeft > tRight after some prefix; because
    // tRight only shrinks and tLeft only grows, "any prefix fails" == "parallel reject or final tLeft > tRight".
    // (a parallel reject after an interval failure also returns false, so the order of the two tests is moot)
    const bool anyfail = __any_sync(FULL, fail);
    tRight = fminf(tRight, warp_min(candR));
    tLeft = fmaxf(tLeft, warp_max(candL));
    if (anyfail || tLeft > tRight) return false;
    if (dirOpt) {
        if (vdot(opt, li.direction) > 0.0f) result = vadd(li.point, vmul(tRight, li.direction));
        else result = vadd(li.point, vmul(tLeft, li.direction));
    } else {
        const float t = vdot(li.direction, vsub(opt, li.point));
        if (t < tLeft) result = vadd(li.point, vmul(tLeft, li.direction));
        else if (t > tRight) result = vadd(li.point, vmul(tRight, li.direction));
        else result = vadd(li.point, vmul(t, li.direction));
    }
    return true;
}

__device__ __forceinline__ int lp2(const Line *lines, int m, float radius, V2 opt, bool dirOpt, V2 &result, int lane)
{
    if (dirOpt) result = vmul(radius, opt);
    else if (vabssq(opt) > sqrf(radius)) result = vmul(radius, vnormalize(opt));
    else result = opt;
    for (int i = 0; i < m; ++i) {
        const Line li = lines[i];
        if (vdet(li.direction, vsub(li.point, result)) > 0.0f) {
            const V2 temp = result;
            if (!lp1(lines, i, radius, opt, dirOpt, result, lane)) { result = temp; return i; }
        }
    }
    return m;
}

__device__ __forceinline__ void lp3(const Line *lines, int m, int beginLine, float radius, V2 &result, Line *proj,
                                    int lane)
{
    float distance = 0.0f;
    for (int i = beginLine; i < m; ++i) {
        const Line li = lines[i];
        if (vdet(li.direction, vsub(li.point, result)) > distance) {
            // projection lines, lane j < i; same-direction parallel lines are skipped, order preserved
            bool keep = false;
            Line line;
            line.point = mk(0.f, 0.f); line.direction = mk(0.f, 0.f);
            if (lane < i) {
                const Line lj = lines[lane];
                const float determinant = vdet(li.direction, lj.direction);
                keep = true;
                if (fabsf(determinant) <= RVO_EPSILON) {
                    if (vdot(li.direction, lj.direction) > 0.0f) keep = false;
                    else line.point = vmul(0.5f, vadd(li.point, lj.point));
                } else {
                    line.point = vadd(li.point,
                                      vmul(vdet(lj.direction, vsub(li.point, lj.point)) / determinant, li.direction));
                }
                if (keep) line.direction = vnormalize(vsub(lj.direction, li.direction));
            }
            const unsigned mask = __ballot_sync(FULL, keep);
            __syncwarp();
            if (keep) proj[__popc(mask & ((1u << lane) - 1u))] = line;
            __syncwarp();
            const int np = __popc(mask);
            const V2 temp = result;
            if (lp2(proj, np, radius, mk(-li.direction.y, li.direction.x), true, result, lane) < np) result = temp;
            distance = vdet(li.direction, vsub(li.point, result));
        }
    }
}

// One human's ORCA step by one warp (orca.py:82-132): `hb` = the environment's human rows [n][8], `rb` = its robot row [9],
// `i` = this human, `na` = humans that exist (rows >= na are padding).  `lines` / `proj` = 32 Line slots each of this warp's
// shared memory.  The new velocity is warp-uniform.
__device__ __forceinline__ V2 orca_agent(const double *hb, const double *rb, int i, int na, float time_step, double safety,
                                         int robot_visible, Line *lines, Line *proj, int lane)
{
    const double *me = hb + i * 8;

    const V2 pos = mk((float)me[0], (float)me[1]);
    const V2 vel = mk((float)me[2], (float)me[3]);
    const float radius = (float)(me[4] + 0.01 + safety);
    const float maxSpeed = (float)me[7];
    double pvx = me[5] - me[0], pvy = me[6] - me[1];            // orca.py:112-115 in f64
    const double speed = sqrt(pvx * pvx + pvy * pvy);
    if (speed > 1.0) { pvx /= speed; pvy /= speed; }
    const V2 pref = mk((float)pvx, (float)pvy);

    // candidate j: other humans in index order, then the robot if visible (crowd_sim.py:565-567)
    const int m = na - 1 + (robot_visible ? 1 : 0);
    V2 op = mk(0.f, 0.f), ov = mk(0.f, 0.f);
    float orad = 0.f, distSq = INFINITY;
    bool valid = false;
    if (lane < m) {
        const double *o;
        if (lane < na - 1) o = hb + (lane < i ? lane : lane + 1) * 8;
        else o = rb;
        op = mk((float)o[0], (float)o[1]);
        ov = mk((float)o[2], (float)o[3]);
        orad = (float)(o[4] + 0.01 + safety);
        distSq = vabssq(vsub(pos, op));
        valid = distSq < sqrf(10.0f);                            // neighborDist = 10
    }
    // rank = position in the stable ascending sort by squared distance; keep the 10 nearest (maxNeighbors)
    int rank = 0;
    for (int j = 0; j < 32; ++j) {
        const float dj = __shfl_sync(FULL, distSq, j);
        const bool vj = __shfl_sync(FULL, (int)valid, j) != 0;
        if (vj && (dj < distSq || (dj == distSq && j < lane))) ++rank;
    }
    const bool selected = valid && rank < 10;
    const int cnt = __popc(__ballot_sync(FULL, selected));

    if (selected) {
        const float invTimeHorizon = 1.0f / 5.0f;                // timeHorizon = 5
        const V2 relativePosition = vsub(op, pos);
        const V2 relativeVelocity = vsub(vel, ov);
        const float dsq = vabssq(relativePosition);
        const float combinedRadius = radius + orad;
        const float combinedRadiusSq = sqrf(combinedRadius);
        Line line;
        V2 u;
        if (dsq > combinedRadiusSq) {
            const V2 w = vsub(relativeVelocity, vmul(invTimeHorizon, relativePosition));
            const float wLengthSq = vabssq(w);
            const float dotProduct1 = vdot(w, relativePosition);
            if (dotProduct1 < 0.0f && sqrf(dotProduct1) > combinedRadiusSq * wLengthSq) {
                const float wLength = sqrtf(wLengthSq);
                const V2 unitW = vdiv(w, wLength);
                line.direction = mk(unitW.y, -unitW.x);
                u = vmul(combinedRadius * invTimeHorizon - wLength, unitW);
            } else {
                const float leg = sqrtf(dsq - combinedRadiusSq);
                if (vdet(relativePosition, w) > 0.0f) {
                    line.direction = vdiv(mk(relativePosition.x * leg - relativePosition.y * combinedRadius,
                                             relativePosition.x * combinedRadius + relativePosition.y * leg), dsq);
                } else {
                    const V2 t = vdiv(mk(relativePosition.x * leg + relativePosition.y * combinedRadius,
                                         -relativePosition.x * combinedRadius + relativePosition.y * leg), dsq);
                    line.direction = mk(-t.x, -t.y);
                }
                const float dotProduct2 = vdot(relativeVelocity, line.direction);
                u = vsub(vmul(dotProduct2, line.direction), relativeVelocity);
            }
        } else {
            const float invTimeStep = 1.0f / time_step;
            const V2 w = vsub(relativeVelocity, vmul(invTimeStep, relativePosition));
            const float wLength = vabs(w);
            const V2 unitW = vdiv(w, wLength);
            line.direction = mk(unitW.y, -unitW.x);
            u = vmul(combinedRadius * invTimeStep - wLength, unitW);
        }
        line.point = vadd(vel, vmul(0.5f, u));
        lines[rank] = line;
    }
    __syncwarp();

    V2 res;
    const int lineFail = lp2(lines, cnt, maxSpeed, pref, false, res, lane);
    if (lineFail < cnt) lp3(lines, cnt, lineFail, maxSpeed, res, proj, lane);
    return res;
}

}  // namespace aem
