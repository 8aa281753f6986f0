"""aemcarl_b200 -- B200-native (sm_100a) implementation of AEMCARL's lookahead hot path.

Host-side mirror of the reference's crowd_nav Policy / crowd_sim Env API on top of hand-written CUDA kernels
reached through a C-ABI shared library (include/aemcarl_b200.h).  There is NO CPU fallback: every compute entry
point raises if libaemcarl_b200.so is missing or no CUDA device is present.
"""
__version__ = "0.1.0"
