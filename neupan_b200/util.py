"""Host-side helpers of the hot path's boundary (mirrors the names of neupan/util/__init__.py).

Only what the PAN path and its API surface need: polygon -> (G, h) half-planes
(util/__init__.py:161-244), column decimation (util:285-305), the ``time_it`` print
decorator (util:29-55), checkpoint path resolution (util:58-94), WrapToPi (util:97-120).
"""
from __future__ import annotations

import os
import sys
import time
from math import pi

import numpy as np

from . import configuration


def time_it(name="Function"):
    """Wall-clock print decorator with the reference's message format (util:29-55).  When the
    instance runs on CUDA the device is synchronised first so the printed time is meaningful."""

    def decorator(func):
        def wrapper(self, *args, **kwargs):
            wrapper.count += 1
            if not configuration.time_print:
                result = func(self, *args, **kwargs)
                wrapper.func_count += 1
                return result
            start = time.time()
            result = func(self, *args, **kwargs)
            dev = getattr(self, "device", None)
            if dev is not None and getattr(dev, "type", "cpu") == "cuda":
                import torch

                torch.cuda.synchronize(dev)
            end = time.time()
            wrapper.func_count += 1
            print(f"{name} execute time {(end - start):.6f} seconds")
            return result

        wrapper.count = 0
        wrapper.func_count = 0
        return wrapper

    return decorator


def file_check(file_name):
    """Resolve a path against cwd, sys.path[0] and the package's parent (util:58-94)."""
    if file_name is None:
        return None
    root_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (file_name, os.path.join(sys.path[0], file_name), os.path.join(os.getcwd(), file_name),
                 os.path.join(root_path, file_name)):
        if os.path.exists(cand):
            return cand
    raise FileNotFoundError("File not found: " + os.path.join(root_path, file_name))


def WrapToPi(rad: float, positive: bool = False) -> float:
    while rad > pi:
        rad -= 2 * pi
    while rad < -pi:
        rad += 2 * pi
    return abs(rad) if positive else rad


def _turn(o, a, b) -> float:
    return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])


def is_convex_and_ordered(points: np.ndarray):
    """(convex?, 'CCW'|'CW'|None) for a 2xN vertex array (util:209-241)."""
    n = points.shape[1]
    if n < 3:
        return False, None
    sign = 0
    for i in range(n):
        c = _turn(points[:, i], points[:, (i + 1) % n], points[:, (i + 2) % n])
        if c == 0:
            continue
        s = 1 if c > 0 else -1
        if sign == 0:
            sign = s
        elif s != sign:
            return False, None
    return True, ("CCW" if sign > 0 else "CW")


def gen_inequal_from_vertex(vertex: np.ndarray):
    """Half-plane form G x <= h of a convex polygon given as 2xN vertices (util:161-206).

    Row i is the (un-normalised) outward normal of edge i -> i+1 in CCW order:
    (dy, -dx), h_i = normal . vertex_i.  A CW polygon is re-ordered the way the reference does
    it (first vertex kept, remaining ones reversed) so row order matches its checkpoints.
    """
    convex, order = is_convex_and_ordered(vertex)
    if not convex:
        print("The polygon constructed by vertex is not convex.")
        return None, None
    v = np.asarray(vertex, dtype=float)
    if order == "CW":
        v = np.hstack([v[:, 0:1], v[:, 1:][:, ::-1]])
    nxt = np.roll(v, -1, axis=1)
    d = nxt - v
    G = np.stack([d[1], -d[0]], axis=1)
    h = np.sum(G * v.T, axis=1, keepdims=True)
    return G, h


def downsample_decimation(mat, m):
    """dim x n -> dim x m by picking columns np.linspace(0, n-1, m).astype(int) (util:285-305).
    Works on numpy arrays and torch tensors (last axis)."""
    n = mat.shape[-1]
    if m >= n:
        return mat
    idx = np.linspace(0, n - 1, m).astype(int)
    return mat[..., idx]


def decimation_indices(n: int, m: int) -> np.ndarray:
    return np.linspace(0, n - 1, m).astype(int)
