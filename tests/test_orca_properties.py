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
