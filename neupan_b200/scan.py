"""Batched lidar scan -> obstacle points on the GPU (SURVEY 8f "next" row 2).

``scan_to_points`` is the B-environment form of ``neupan.scan_to_point`` / ``neupan.scan_to_point_velocity``
(neupan/neupan.py:173-281) followed by the decimation to ``dune_max_num`` that PAN applies (pan.py:171-174): it
returns exactly what ``PAN.forward`` takes as ``obs_points``, ``point_velocities`` and ``num_points``, resident on the
device.  The single-environment numpy methods of the reference API live on the ``neupan`` facade (host code).
There is no CPU implementation here: the CUDA library is the product.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def scan_to_points(states, ranges, scan: dict, scan_offset=(0.0, 0.0, 0.0), angle_range=(-np.pi, np.pi), down_sample: int = 1,
                   max_points: int | None = None, velocity=None, velocity_mode: bool | None = None, device=None):
    """states (B,3) [x, y, theta]; ranges (B,R); ``scan`` carries angle_min/angle_max/range_min/range_max like the
    reference's dict; ``velocity`` (B,2,R) per-beam velocities or None.  ``velocity_mode`` selects the semantics of
    scan_to_point_velocity (inverse sensor offset, range >= range_min); default: on iff ``velocity`` is given.
    ``max_points``: capacity of the result and decimation target (the planner's dune_max_num); default R.
    Returns (points (B,2,max_points) f32, velocities (B,2,max_points) f32 or None, counts (B,) int32) on the GPU."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("neupan_b200.scan_to_points needs a CUDA device (no CPU fallback)")
    dev = torch.device(device) if device is not None else (ranges.device if isinstance(ranges, torch.Tensor) and ranges.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    ranges = torch.as_tensor(ranges).to(device=dev, dtype=torch.float32).contiguous()
    if ranges.dim() != 2:
        raise ValueError("ranges must be (B, R)")
    B, R = ranges.shape
    states = torch.as_tensor(states).reshape(B, 3).to(device=dev, dtype=torch.float64).contiguous()
    if velocity is not None:
        velocity = torch.as_tensor(velocity).to(device=dev, dtype=torch.float32).contiguous()
        if velocity.shape != (B, 2, R):
            raise ValueError("velocity must be (B, 2, R)")
    if velocity_mode is None:
        velocity_mode = velocity is not None
    max_points = int(R if max_points is None else max_points)
    cfg = _lib.ScanConfig()
    cfg.angle_min, cfg.angle_max = float(scan["angle_min"]), float(scan["angle_max"])
    cfg.range_min, cfg.range_max = float(scan["range_min"]), float(scan["range_max"])
    cfg.scan_offset = (C.c_double * 3)(*[float(v) for v in scan_offset])
    cfg.angle_range = (C.c_double * 2)(float(angle_range[0]), float(angle_range[1]))
    cfg.down_sample, cfg.velocity_mode = int(down_sample), int(bool(velocity_mode))
    points = torch.empty((B, 2, max_points), dtype=torch.float32, device=dev)
    vel_out = torch.empty((B, 2, max_points), dtype=torch.float32, device=dev) if velocity_mode else None
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    ptr = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.nb_scan_to_points(B, R, ptr(ranges), ptr(velocity), ptr(states), C.byref(cfg), max_points, ptr(points), ptr(vel_out),
                                         ptr(counts), stream))
    return points, vel_out, counts
