"""State value types (crowd_sim/envs/utils/state.py:1-50).  `self_state + human_state` yields the 14-tuple
(robot 9, human 5) that rotate() consumes (state.py:17-18,36-37)."""


class FullState(object):
    def __init__(self, px, py, vx, vy, radius, gx, gy, v_pref, theta):
        self.px, self.py, self.vx, self.vy, self.radius = px, py, vx, vy, radius
        self.gx, self.gy, self.v_pref, self.theta = gx, gy, v_pref, theta
        self.position = (self.px, self.py)
        self.goal_position = (self.gx, self.gy)
        self.velocity = (self.vx, self.vy)

    def as_tuple(self):
        return (self.px, self.py, self.vx, self.vy, self.radius, self.gx, self.gy, self.v_pref, self.theta)

    def __add__(self, other):
        return other + self.as_tuple()

    def __str__(self):
        return " ".join(str(x) for x in self.as_tuple())


class ObservableState(object):
    def __init__(self, px, py, vx, vy, radius):
        self.px, self.py, self.vx, self.vy, self.radius = px, py, vx, vy, radius
        self.position = (self.px, self.py)
        self.velocity = (self.vx, self.vy)

    def as_tuple(self):
        return (self.px, self.py, self.vx, self.vy, self.radius)

    def __add__(self, other):
        return other + self.as_tuple()

    def __str__(self):
        return " ".join(str(x) for x in self.as_tuple())


class JointState(object):
    def __init__(self, self_state, human_states):
        assert isinstance(self_state, FullState)
        for human_state in human_states:
            assert isinstance(human_state, ObservableState)
        self.self_state = self_state
        self.human_states = human_states
