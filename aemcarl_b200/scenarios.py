"""Scenario generation of CrowdSim.reset, bit-compatible with the reference's seeded numpy stream.

Restates crowd_sim/envs/crowd_sim.py:486-552 (reset), :318-388 (generate_random_human_position),
:390-411 (circle crossing), :413-442 (square crossing) and Agent.sample_random_attributes (agent.py:39-45).
`np.random.seed(s)` + `np.random.random()` of the reference is the legacy MT19937 stream, reproduced here with
a private `np.random.RandomState(s)` in the same call order (rejection sampling included), so the start / goal
positions are identical to the reference's for every (phase, case).
"""
import math

import numpy as np

# case_capacity of crowd_sim.py:277 -> counter offsets of :501-505
CASE_CAPACITY = {"train": np.iinfo(np.uint32).max - 2000, "val": 1000, "test": 1000}
COUNTER_OFFSET = {"train": CASE_CAPACITY["val"] + CASE_CAPACITY["test"], "val": 0, "test": CASE_CAPACITY["val"]}


def case_seed(phase, case):
    return COUNTER_OFFSET[phase] + case


def _norm(dx, dy):
    return math.sqrt(dx * dx + dy * dy)


class _Agent(object):
    __slots__ = ("px", "py", "gx", "gy", "radius", "v_pref")

    def __init__(self, px, py, gx, gy, radius, v_pref):
        self.px, self.py, self.gx, self.gy, self.radius, self.v_pref = px, py, gx, gy, radius, v_pref


def _circle_human(rs, agents, circle_radius, discomfort_dist, radius, v_pref, randomize):
    if randomize:
        v_pref = rs.uniform(0.5, 1.5)
        radius = rs.uniform(0.3, 0.5)
    while True:
        angle = rs.random_sample() * np.pi * 2
        px_noise = (rs.random_sample() - 0.5) * v_pref
        py_noise = (rs.random_sample() - 0.5) * v_pref
        px = circle_radius * np.cos(angle) + px_noise
        py = circle_radius * np.sin(angle) + py_noise
        collide = False
        for a in agents:
            min_dist = radius + a.radius + discomfort_dist
            if _norm(px - a.px, py - a.py) < min_dist or _norm(px - a.gx, py - a.gy) < min_dist:
                collide = True
                break
        if not collide:
            break
    return _Agent(px, py, -px, -py, radius, v_pref)


def _square_human(rs, agents, square_width, discomfort_dist, radius, v_pref, randomize):
    if randomize:
        v_pref = rs.uniform(0.5, 1.5)
        radius = rs.uniform(0.3, 0.5)
    sign = -1 if rs.random_sample() > 0.5 else 1
    while True:
        px = rs.random_sample() * square_width * 0.5 * sign
        py = (rs.random_sample() - 0.5) * square_width
        collide = False
        for a in agents:
            if _norm(px - a.px, py - a.py) < radius + a.radius + discomfort_dist:
                collide = True
                break
        if not collide:
            break
    while True:
        gx = rs.random_sample() * square_width * 0.5 * -sign
        gy = (rs.random_sample() - 0.5) * square_width
        collide = False
        for a in agents:
            if _norm(gx - a.gx, gy - a.gy) < radius + a.radius + discomfort_dist:
                collide = True
                break
        if not collide:
            break
    return _Agent(px, py, gx, gy, radius, v_pref)


# 'mixed' rule (crowd_sim.py:337-386): with probability 0.2 a static scene, otherwise a dynamic one; the number of
# humans is drawn from a fixed distribution per kind.
_STATIC_COUNT_P = ((0, 0.05), (1, 0.2), (2, 0.2), (3, 0.3), (4, 0.1), (5, 0.15))
_DYNAMIC_COUNT_P = ((1, 0.3), (2, 0.3), (3, 0.2), (4, 0.1), (5, 0.1))


def _mixed_humans(rs, robot, human_num, circle_radius, square_width, discomfort_dist, radius, v_pref, randomize):
    """Returns (human_num of the episode, list of agents).  A static scene with zero humans still carries one dummy
    human parked at (0, -10) (crowd_sim.py:355-358); static humans have goal = start."""
    static = rs.random_sample() < 0.2
    prob = rs.random_sample()
    for count, p in (_STATIC_COUNT_P if static else _DYNAMIC_COUNT_P):
        if prob - p <= 0:
            human_num = count
            break
        prob -= p
    agents = [robot]
    if static:
        width, height = 4, 8
        if human_num == 0:
            agents.append(_Agent(0, -10, 0, -10, radius, v_pref))
        for _ in range(human_num):
            sign = -1 if rs.random_sample() > 0.5 else 1
            while True:
                px = rs.random_sample() * width * 0.5 * sign
                py = (rs.random_sample() - 0.5) * height
                if all(_norm(px - a.px, py - a.py) >= radius + a.radius + discomfort_dist for a in agents):
                    break
            agents.append(_Agent(px, py, px, py, radius, v_pref))
    else:
        for i in range(human_num):       # the first two cross the circle, the others the square
            make = _circle_human if i < 2 else _square_human
            extent = circle_radius if i < 2 else square_width
            agents.append(make(rs, agents, extent, discomfort_dist, radius, v_pref, randomize))
    return human_num, agents[1:]


def generate_case(phase, case, human_num, rule, circle_radius=4.0, square_width=10.0, discomfort_dist=0.2,
                  robot_radius=0.3, robot_v_pref=1.0, human_radius=0.3, human_v_pref=1.0, randomize_attributes=False):
    """One scenario: (robot[9], humans[human_num][8]) float64, laid out as include/aemcarl_b200.h documents."""
    rs = np.random.RandomState(case_seed(phase, case))
    robot = _Agent(0.0, -circle_radius, 0.0, circle_radius, robot_radius, robot_v_pref)   # crowd_sim.py:512
    agents = [robot]
    if rule == "mixed":        # the number of humans is part of the draw: the result has len(humans) rows
        _, hs_list = _mixed_humans(rs, robot, human_num, circle_radius, square_width, discomfort_dist, human_radius,
                                   human_v_pref, randomize_attributes)
        agents += hs_list
        human_num = len(hs_list)
    for i in range(0 if rule == "mixed" else human_num):
        if rule == "circle_crossing":
            h = _circle_human(rs, agents, circle_radius, discomfort_dist, human_radius, human_v_pref,
                              randomize_attributes)
        elif rule == "square_crossing":
            h = _square_human(rs, agents, square_width, discomfort_dist, human_radius, human_v_pref,
                              randomize_attributes)
        else:
            raise ValueError("Rule doesn't exist")
        agents.append(h)
    rb = np.array([robot.px, robot.py, 0.0, 0.0, robot.radius, robot.gx, robot.gy, robot.v_pref, np.pi / 2])
    hs = np.array([[h.px, h.py, 0.0, 0.0, h.radius, h.gx, h.gy, h.v_pref] for h in agents[1:]], dtype=np.float64)
    return rb, hs.reshape(human_num, 8)


MIXED_MAX_HUMANS = 5       # the largest count either distribution of the 'mixed' rule can draw


def generate_mixed_pool(phase, first_case, count, **kw):
    """'mixed' scenarios stacked for a batch: robot [count][9], humans [count][5][8] (rows beyond a case's human count are
    padding: parked far away, zero velocity) and n_active [count] int32 -- the per-environment human count the kernels take
    through aem_set_active_counts."""
    n = MIXED_MAX_HUMANS
    robots = np.empty((count, 9), dtype=np.float64)
    humans = np.zeros((count, n, 8), dtype=np.float64)
    humans[:, :, 0:2] = humans[:, :, 5:7] = [0.0, -100.0]
    humans[:, :, 4] = kw.get("human_radius", 0.3)
    counts = np.zeros(count, dtype=np.int32)
    for i in range(count):
        robots[i], hs = generate_case(phase, first_case + i, n, "mixed", **kw)
        counts[i] = hs.shape[0]
        humans[i, :hs.shape[0]] = hs
    return robots, humans, counts


def generate_pool(phase, first_case, count, human_num, rule, **kw):
    """Scenarios first_case .. first_case+count-1 stacked: robot [count][9], humans [count][n][8]."""
    if rule == "mixed":
        raise ValueError("'mixed' draws a different number of humans per case: use generate_mixed_pool (padded rows + counts)")
    robots = np.empty((count, 9), dtype=np.float64)
    humans = np.empty((count, human_num, 8), dtype=np.float64)
    for i in range(count):
        robots[i], humans[i] = generate_case(phase, first_case + i, human_num, rule, **kw)
    return robots, humans
