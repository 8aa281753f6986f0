"""CPU: known-answer and property tests that pin our FP32 restatement of RVO2's ORCA (parity against the real
Python-RVO2 is unpinned: the library is not installable here).  SURVEY.md 8(c) (i)-(vi)."""
import numpy as np
import pytest


def halfplanes_ok(lines, v, tol=1e-4):
    # det(direction, point - v) <= 0  <=>  v on the permitted side
    for pt_x, pt_y, dx, dy in lines:
        if dx * (pt_y - v[1]) - dy * (pt_x - v[0]) > tol:
            return False
    return True


def brute_force(lines, pref, max_speed, grid=401):
    xs = np.linspace(-max_speed, max_speed, grid)
    X, Y = np.meshgrid(xs, xs)
    ok = X * X + Y * Y <= max_speed ** 2 + 1e-12
    for pt_x, pt_y, dx, dy in lines.astype(np.float64):
        ok &= dx * (pt_y - Y) - dy * (pt_x - X) <= 1e-9
    if not ok.any():
        return None
    d = (X - pref[0]) ** 2 + (Y - pref[1]) ** 2
    d[~ok] = np.inf
    i = np.unravel_index(np.argmin(d), d.shape)
    return np.array([X[i], Y[i]]), np.sqrt(d[i])


def test_no_neighbours_gives_clipped_pref_velocity(oracle):
    v = oracle.orca_agent([0, 0], [0, 0], 0.31, 1.0, [0.3, 0.4], np.zeros((0, 5)))
    assert np.allclose(v, [0.3, 0.4])
    v = oracle.orca_agent([0, 0], [0, 0], 0.31, 0.5, [3.0, 4.0], np.zeros((0, 5)))
    assert np.allclose(v, [0.3, 0.4], atol=1e-7)
    far = np.array([[20.0, 0, 0, 0, 0.31]])                     # outside neighborDist = 10
    v = oracle.orca_agent([0, 0], [0, 0], 0.31, 1.0, [1.0, 0.0], far)
    assert np.allclose(v, [1.0, 0.0])


def test_head_on_closed_form(oracle):
    """two agents approaching head-on along x with a lateral offset: both must dodge to opposite sides, symmetric"""
    a = oracle.orca_agent([0, 0.0], [1, 0], 0.31, 1.0, [1, 0], np.array([[3.0, 0.1, -1, 0, 0.31]]))
    b = oracle.orca_agent([3.0, 0.1], [-1, 0], 0.31, 1.0, [-1, 0], np.array([[0.0, 0.0, 1, 0, 0.31]]))
    assert a[1] < 0 and b[1] > 0                                # each yields away from the other
    assert np.allclose(a, -b, atol=1e-6)                        # reciprocity: point symmetry of the pair
    assert np.hypot(*a) <= 1.0 + 1e-6


@pytest.mark.parametrize("seed", range(40))
def test_feasible_and_optimal_vs_brute_force(oracle, seed):
    rng = np.random.RandomState(seed)
    m = rng.randint(1, 8)
    others = np.zeros((m, 5), dtype=np.float32)
    others[:, 0:2] = rng.uniform(-3, 3, (m, 2))
    others[:, 2:4] = rng.uniform(-1, 1, (m, 2))
    others[:, 4] = 0.31
    keep = np.hypot(others[:, 0], others[:, 1]) > 0.7           # not overlapping the agent at the origin
    others = others[keep]
    pref = rng.uniform(-1, 1, 2)
    vel = rng.uniform(-1, 1, 2)
    v, nb, lines = oracle.orca_agent([0, 0], vel, 0.31, 1.0, pref, others, want_lines=True)
    assert np.hypot(*v) <= 1.0 + 1e-5
    bf = brute_force(lines, np.asarray(pref, dtype=np.float64) / max(1.0, np.hypot(*pref)), 1.0)
    if bf is not None and halfplanes_ok(lines, v):
        # LP2 succeeded: closest feasible point to the (clipped) preferred velocity
        p = np.asarray(pref, dtype=np.float64)
        if np.hypot(*p) > 1:
            p = p / np.hypot(*p)
        assert np.hypot(*(v - p)) <= bf[1] + 0.02               # within the brute-force grid resolution
    # neighbours are sorted by distance
    d = np.hypot(others[nb, 0], others[nb, 1])
    assert np.all(np.diff(d) >= 0)


def test_mirror_equivariance(oracle):
    rng = np.random.RandomState(3)
    for _ in range(20):
        m = 4
        others = np.zeros((m, 5), dtype=np.float32)
        others[:, 0:2] = rng.uniform(-3, 3, (m, 2))
        others[:, 2:4] = rng.uniform(-1, 1, (m, 2))
        others[:, 4] = 0.31
        pref, vel = rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 2)
        v = oracle.orca_agent([0, 0], vel, 0.31, 1.0, pref, others)
        om = others.copy()
        om[:, 0] *= -1
        om[:, 2] *= -1
        vm = oracle.orca_agent([0, 0], [-vel[0], vel[1]], 0.31, 1.0, [-pref[0], pref[1]], om)
        assert np.allclose(vm, [-v[0], v[1]], atol=2e-5)


def test_only_ten_nearest_constrain(oracle):
    rng = np.random.RandomState(9)
    ang = rng.uniform(0, 2 * np.pi, 17)
    rad = np.linspace(1.0, 6.0, 17)
    others = np.zeros((17, 5), dtype=np.float32)
    others[:, 0], others[:, 1] = rad * np.cos(ang), rad * np.sin(ang)
    others[:, 2:4] = rng.uniform(-0.5, 0.5, (17, 2))
    others[:, 4] = 0.31
    perm = rng.permutation(17)
    v, nb, lines = oracle.orca_agent([0, 0], [0.2, 0.1], 0.31, 1.0, [1, 0], others[perm], want_lines=True)
    assert len(nb) == 10
    assert sorted(perm[nb].tolist()) == list(range(10))         # the 10 nearest, whatever the input order
    v2 = oracle.orca_agent([0, 0], [0.2, 0.1], 0.31, 1.0, [1, 0], others[:10])
    assert np.array_equal(v, v2)


def test_overlap_uses_time_step_branch(oracle):
    """overlapping discs: the constraint is the cut-off circle of radius R/timeStep (1 / timeStep branch)"""
    others = np.array([[0.4, 0.0, 0.0, 0.0, 0.31]], dtype=np.float32)
    v, nb, lines = oracle.orca_agent([0, 0], [0, 0], 0.31, 1.0, [1, 0], others, time_step=0.2, want_lines=True)
    # w = -rp / dt = (-2, 0); u = (R/dt - |w|) * w_hat = (3.1 - 2) * (-1, 0); point = v + u/2
    assert np.allclose(lines[0][:2], [-0.55, 0.0], atol=1e-5)
    assert np.allclose(lines[0][2:], [0.0, 1.0], atol=1e-6)
    assert v[0] <= -0.55 + 1e-5                                  # pushed away from the overlapping neighbour


# ------------------------------------------------------------------------------------------------------------------
# Independent FP64 solvers (vertex enumeration -- not the incremental algorithm of RVO2) for the two programs of
# Agent::computeNewVelocity:
#   linearProgram2: the point of {speed disc} ∩ {half-planes} closest to the preferred velocity;
#   linearProgram3 (infeasible case): the point of the speed disc that minimises the largest half-plane violation
#                   det(direction_i, point_i - v).
# The oracle's FP32 result must reach the FP64 optimum within 1e-4 (objective) with violations <= 1e-4.
# ------------------------------------------------------------------------------------------------------------------
def _unit(lines):
    """float64 copy of the half-planes with the (FP32-normalised) directions re-normalised in float64"""
    L = np.array(lines, dtype=np.float64).reshape(-1, 4)
    L[:, 2:] /= np.hypot(L[:, 2], L[:, 3])[:, None]
    return L


def _viol(lines, v):
    """signed violation of every half-plane at v (<= 0: satisfied)"""
    L = _unit(lines)
    return L[:, 2] * (L[:, 1] - v[1]) - L[:, 3] * (L[:, 0] - v[0])


def _line_circle(pt, d, R):
    b = pt @ d
    disc = b * b + R * R - pt @ pt
    if disc < 0:
        return []
    s = np.sqrt(disc)
    return [pt + (-b - s) * d, pt + (-b + s) * d]


def exact_lp2(lines, pref, R, tol=1e-9):
    """closest point to `pref` in disc(R) ∩ half-planes by candidate enumeration; None if (numerically) infeasible"""
    L = _unit(lines)
    p = np.asarray(pref, dtype=np.float64)
    cands = [p if np.hypot(*p) <= R else p * (R / np.hypot(*p))]
    for i in range(len(L)):
        pt, d = L[i, :2], L[i, 2:]
        cands.append(pt + ((p - pt) @ d) * d)                       # projection of pref on the boundary line
        cands += _line_circle(pt, d, R)
        for j in range(i):
            pj, dj = L[j, :2], L[j, 2:]
            den = d[0] * dj[1] - d[1] * dj[0]
            if abs(den) > 1e-12:
                t = (dj[0] * (pt[1] - pj[1]) - dj[1] * (pt[0] - pj[0])) / den
                cands.append(pt + t * d)
    best, bd = None, np.inf
    for c in cands:
        if np.hypot(*c) <= R + tol and (len(L) == 0 or _viol(L, c).max() <= tol):
            dist = np.hypot(*(c - p))
            if dist < bd:
                best, bd = c, dist
    return None if best is None else (best, bd)


def exact_minmax(lines, R):
    """min over the disc of max_i violation_i(v) by candidate enumeration (upper-envelope vertices, envelope edges on the
    circle, single-plane minima on the circle)"""
    L = _unit(lines)
    m = len(L)
    g = np.stack([L[:, 3], -L[:, 2]], axis=1)            # violation_i(v) = c_i + g_i . v
    c = L[:, 2] * L[:, 1] - L[:, 3] * L[:, 0]
    cands = [np.zeros(2)]
    for i in range(m):
        cands.append(-R * g[i] / np.hypot(*g[i]))
        for j in range(i):
            a, b = g[i] - g[j], c[j] - c[i]               # a . v = b
            na = np.hypot(*a)
            if na < 1e-12:
                continue
            p0 = a * (b / (na * na))
            d = np.array([-a[1], a[0]]) / na
            cands += _line_circle(p0, d, R)
            for k in range(j):
                A = np.stack([g[i] - g[j], g[i] - g[k]])
                if abs(np.linalg.det(A)) > 1e-12:
                    cands.append(np.linalg.solve(A, np.array([c[j] - c[i], c[k] - c[i]])))
    best, bv = None, np.inf
    for v in cands:
        if np.hypot(*v) <= R + 1e-9:
            f = (c + g @ v).max()
            if f < bv:
                best, bv = v, f
    return best, bv


def crowded_layout(rng, m, spread, overlap):
    others = np.zeros((m, 5), dtype=np.float32)
    ang = rng.uniform(0, 2 * np.pi, m)
    rad = rng.uniform(0.66 if not overlap else 0.3, spread, m)
    others[:, 0], others[:, 1] = rad * np.cos(ang), rad * np.sin(ang)
    others[:, 2:4] = rng.uniform(-1, 1, (m, 2))
    others[:, 4] = 0.31
    return others


@pytest.mark.parametrize("seed", range(60))
def test_lp2_reaches_the_exact_fp64_optimum(oracle, seed):
    """feasible case: result feasible (violation <= 1e-4) and as close to the preferred velocity as the exact optimum
    (<= 1e-4 farther): replaces the 0.02 grid bound"""
    rng = np.random.RandomState(1000 + seed)
    m = rng.randint(1, 11)
    others = crowded_layout(rng, m, 4.0, overlap=False)
    pref, vel = rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 2)
    v, nb, lines = oracle.orca_agent([0, 0], vel, 0.31, 1.0, pref, others, want_lines=True)
    p = np.asarray(pref, dtype=np.float64)
    ex = exact_lp2(lines, p, 1.0)
    _, slack = exact_minmax(lines, 1.0)
    if ex is None or slack > -1e-3:
        pytest.skip("(nearly) infeasible layout: covered by the linearProgram3 test")
    assert np.hypot(*v) <= 1.0 + 1e-5
    assert _viol(lines, v.astype(np.float64)).max() <= 1e-4
    assert np.hypot(*(v - p)) <= ex[1] + 1e-4, (np.hypot(*(v - p)), ex[1])
    assert np.hypot(*(v - ex[0])) <= 2e-3          # the optimum is unique (strictly convex objective)


def test_lp3_minimises_the_largest_violation(oracle):
    """infeasible case (crowded / overlapping layouts): the result must minimise max_i det(direction_i, point_i - v)
    over the speed disc -- checked against the exact FP64 optimum on every infeasible layout of 400 draws"""
    rng = np.random.RandomState(77)
    checked = 0
    for it in range(400):
        m = rng.randint(3, 11)
        others = crowded_layout(rng, m, 1.6, overlap=(it % 2 == 0))
        pref, vel = rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 2)
        v, nb, lines = oracle.orca_agent([0, 0], vel, 0.31, 1.0, pref, others, want_lines=True)
        best, fstar = exact_minmax(lines, 1.0)
        if fstar <= 1e-3:
            continue                                                 # feasible (or marginal): linearProgram2's job
        checked += 1
        assert np.hypot(*v) <= 1.0 + 1e-5
        f = _viol(lines, v.astype(np.float64)).max()
        assert f <= fstar + 1e-4, (it, f, fstar)
    assert checked >= 60, checked


def test_equal_distance_neighbours_keep_index_order(oracle):
    """Ties in squared distance (the documented open point: RVO2's kd-tree order is unknown): the oracle and the CUDA
    kernel keep the LOWER index -- the insertion test is a strict `<` and a full list rejects an equal distance."""
    ang = np.linspace(0, 2 * np.pi, 12, endpoint=False)
    others = np.zeros((12, 5), dtype=np.float32)
    others[:, 0], others[:, 1] = np.float32(2.0) * np.cos(ang), np.float32(2.0) * np.sin(ang)
    d2 = others[:, 0] ** 2 + others[:, 1] ** 2
    others[:, 4] = 0.31
    v, nb, lines = oracle.orca_agent([0, 0], [0.1, 0.0], 0.31, 1.0, [1, 0], others, want_lines=True)
    assert len(nb) == 10
    order = sorted(range(12), key=lambda j: (d2[j], j))[:10]         # stable by (distance, index)
    assert nb.tolist() == order
