"""Episode info objects (crowd_sim/envs/utils/info.py); callers test them with isinstance (explorer.py:58-74)."""


class Timeout(object):
    def __str__(self):
        return "Timeout"


class ReachGoal(object):
    def __str__(self):
        return "Reaching goal"


class Danger(object):
    def __init__(self, min_dist):
        self.min_dist = min_dist

    def __str__(self):
        return "Too close"


class Collision(object):
    def __str__(self):
        return "Collision"


class Stationary(object):
    def __str__(self):
        return "No action"


class Nothing(object):
    def __str__(self):
        return ""


def from_code(code, dmin=None):
    """aem info code (include/aemcarl_b200.h AEM_INFO_*) -> info object"""
    if code == 0:
        return Nothing()
    if code == 1:
        return Danger(dmin)
    if code == 2:
        return ReachGoal()
    if code == 3:
        return Collision()
    if code == 4:
        return Timeout()
    raise ValueError("unknown info code %r" % (code,))
