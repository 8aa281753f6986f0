"""`import rvo2` of the reference (Python-RVO2) -> the adapter over the ORCA kernel.  Put <repo>/dropin on sys.path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aemcarl_b200.rvo2 import PyRVOSimulator              # noqa: E402,F401
