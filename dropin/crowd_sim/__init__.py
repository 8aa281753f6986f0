"""`import crowd_sim` of the reference (crowd_sim/__init__.py) -> aemcarl_b200.crowd_sim.  Put <repo>/dropin on sys.path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _alias import alias_package                          # noqa: E402

alias_package("aemcarl_b200.crowd_sim", "crowd_sim", sys.modules[__name__])
