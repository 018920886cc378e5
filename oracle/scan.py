"""CPU restatement of the lidar scan -> points step (TEST INFRASTRUCTURE ONLY, like the rest of oracle/).

Follows, beam by beam, ``neupan.scan_to_point`` (neupan/neupan.py:173-222) and ``neupan.scan_to_point_velocity``
(neupan.py:224-281), then the decimation PAN applies when more than ``dune_max_num`` points arrive
(pan.py:171-174 -> util.downsample_decimation, neupan/util/__init__.py:285-305), and the float32 cast of
``np_to_tensor`` (neupan.py:123-126).  Pinned against the reference's own functions (imported with stub modules,
oracle/refload.py) by tests/test_oracle_scan.py and the golden vectors tests/golden/ref_scan.npz.
"""
from math import cos, sin

import numpy as np


def _transform(state):
    """util.get_transform for a (3,1) state (neupan/util/__init__.py: rotation from theta, translation from x, y)."""
    th = float(state[2, 0])
    return state[0:2], np.array([[cos(th), -sin(th)], [sin(th), cos(th)]])


def scan_to_point(state, scan, scan_offset=(0, 0, 0), angle_range=(-np.pi, np.pi), down_sample=1):
    """neupan.py:173-222."""
    cloud = []
    ranges = np.array(scan["ranges"])
    angles = np.linspace(scan["angle_min"], scan["angle_max"], len(ranges))
    for i in range(len(ranges)):
        r, a = ranges[i], angles[i]
        if r < (scan["range_max"] - 0.02) and r > scan["range_min"]:  # :207
            if a > angle_range[0] and a < angle_range[1]:  # :208
                cloud.append(np.array([[r * cos(a)], [r * sin(a)]]))
    if len(cloud) == 0:
        return None
    pts = np.hstack(cloud)
    s_trans, s_R = _transform(np.c_[list(scan_offset)])
    temp = s_R @ pts + s_trans  # :216-217
    trans, R = _transform(state)
    return (R @ temp + trans)[:, ::down_sample]  # :219-220


def scan_to_point_velocity(state, scan, scan_offset=(0, 0, 0), angle_range=(-np.pi, np.pi), down_sample=1):
    """neupan.py:224-281."""
    cloud, vels = [], []
    ranges = np.array(scan["ranges"])
    angles = np.linspace(scan["angle_min"], scan["angle_max"], len(ranges))
    scan_velocity = scan.get("velocity", np.zeros((2, len(ranges))))  # :250
    for i in range(len(ranges)):
        r, a = ranges[i], angles[i]
        if r < (scan["range_max"] - 0.02) and r >= scan["range_min"]:  # :258 (inclusive lower bound)
            if a > angle_range[0] and a < angle_range[1]:
                cloud.append(np.array([[r * cos(a)], [r * sin(a)]]))
                vels.append(scan_velocity[:, i:i + 1])
    if len(cloud) == 0:
        return None, None
    pts = np.hstack(cloud)
    s_trans, s_R = _transform(np.c_[list(scan_offset)])
    temp = s_R.T @ (pts - s_trans)  # :271-273 (inverse sensor offset)
    trans, R = _transform(state)
    return (R @ temp + trans)[:, ::down_sample], np.hstack(vels)[:, ::down_sample]  # :275-279


def decimate(mat, m):
    """util.downsample_decimation (neupan/util/__init__.py:285-305)."""
    n = mat.shape[1]
    if m >= n:
        return mat
    return mat[:, np.linspace(0, n - 1, m).astype(int)]


def scan_batch(states, ranges, scan, scan_offset=(0, 0, 0), angle_range=(-np.pi, np.pi), down_sample=1, max_points=None,
               velocity=None, velocity_mode=None, fn_point=scan_to_point, fn_velocity=scan_to_point_velocity):
    """B scans -> (points (B,2,max_points) f32 zero padded, velocities or None, counts (B,) int32): the layout of
    nb_scan_to_points.  ``fn_point`` / ``fn_velocity`` let the tests run the reference's own functions through the same
    packing."""
    ranges = np.asarray(ranges)
    B, R = ranges.shape
    max_points = R if max_points is None else max_points
    if velocity_mode is None:
        velocity_mode = velocity is not None
    pts = np.zeros((B, 2, max_points), np.float32)
    vel = np.zeros((B, 2, max_points), np.float32) if velocity_mode else None
    cnt = np.zeros(B, np.int32)
    for b in range(B):
        sc = dict(scan, ranges=ranges[b].astype(np.float64))
        st = np.asarray(states[b], dtype=np.float64).reshape(3, 1)
        if velocity_mode:
            if velocity is not None:
                sc["velocity"] = np.asarray(velocity[b], dtype=np.float64)
            p, v = fn_velocity(st, sc, list(scan_offset), list(angle_range), down_sample)
        else:
            p, v = fn_point(st, sc, list(scan_offset), list(angle_range), down_sample), None
        if p is None:
            continue
        p = decimate(p, max_points)
        cnt[b] = p.shape[1]
        pts[b, :, :cnt[b]] = p.astype(np.float32)  # np_to_tensor: float32 (neupan.py:123-126)
        if v is not None:
            vel[b, :, :cnt[b]] = decimate(v, max_points).astype(np.float32)
    return pts, vel, cnt
