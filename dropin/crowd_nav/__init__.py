"""`import crowd_nav` of the reference -> aemcarl_b200.crowd_nav.  Put <repo>/dropin on sys.path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _alias import alias_package                          # noqa: E402

alias_package("aemcarl_b200.crowd_nav", "crowd_nav", sys.modules[__name__])
