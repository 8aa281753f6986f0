"""crowd_sim/envs/utils/utils.py:4-26 (host copy used by tests; the kernels carry their own FP64 version)."""
import math


def point_to_segment_dist(x1, y1, x2, y2, x3, y3):
    px, py = x2 - x1, y2 - y1
    if px == 0 and py == 0:
        return math.sqrt((x3 - x1) ** 2 + (y3 - y1) ** 2)
    u = ((x3 - x1) * px + (y3 - y1) * py) / (px * px + py * py)
    u = 1 if u > 1 else (0 if u < 0 else u)
    x, y = x1 + u * px, y1 + u * py
    return math.sqrt((x - x3) ** 2 + (y - y3) ** 2)
