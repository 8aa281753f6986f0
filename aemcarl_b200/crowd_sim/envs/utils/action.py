"""Action value types (crowd_sim/envs/utils/action.py:3-4)."""
from collections import namedtuple

ActionXY = namedtuple("ActionXY", ["vx", "vy"])
ActionRot = namedtuple("ActionRot", ["v", "r"])
