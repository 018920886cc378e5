"""Golden vectors for the lidar scan -> points step, produced by the REFERENCE's own functions
(neupan.scan_to_point / neupan.scan_to_point_velocity, /root/reference/neupan/neupan.py:173-281, imported through
oracle/refload.py) and util.downsample_decimation, packed into the nb_scan_to_points layout by oracle.scan.scan_batch.

    python tests/golden/make_golden_scan.py        # needs /root/reference; writes tests/golden/ref_scan.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refload, scan as oscan  # noqa: E402

CASES = {
    # name: (B, R, scan dict, offset, angle_range, down_sample, max_points, velocity_mode)
    "plain": (4, 360, dict(angle_min=-np.pi, angle_max=np.pi, range_min=0.1, range_max=10.0), (0.3, -0.1, 0.2), (-np.pi, np.pi), 1, 360, False),
    "velocity_stride": (3, 257, dict(angle_min=-2.5, angle_max=2.75, range_min=0.5, range_max=8.0), (0.0, 0.2, -0.4), (-2.0, 2.5), 2, 257, True),
    "decimated": (3, 720, dict(angle_min=-np.pi, angle_max=np.pi, range_min=0.1, range_max=12.0), (0.1, 0.0, 0.0), (-np.pi, np.pi), 1, 100, True),
    "decimated_stride3": (2, 1000, dict(angle_min=-3.0, angle_max=3.0, range_min=0.2, range_max=10.0), (0.0, 0.0, 0.0), (-3.0, 3.0), 3, 64, False),
    "nothing_in_range": (2, 90, dict(angle_min=-1.0, angle_max=1.0, range_min=0.1, range_max=5.0), (0.0, 0.0, 0.0), (-np.pi, np.pi), 1, 90, False),
}


def inputs(name, B, R, scan, seed):
    rng = np.random.default_rng(seed)
    hi = scan["range_max"] * (1.15 if name != "nothing_in_range" else 3.0)
    lo = 0.0 if name != "nothing_in_range" else scan["range_max"]
    ranges = rng.uniform(lo, hi, size=(B, R)).astype(np.float32)
    if name == "velocity_stride":  # beams exactly at range_min: kept by scan_to_point_velocity (>=), dropped by scan_to_point (>)
        ranges[:, ::17] = np.float32(scan["range_min"])
    states = np.stack([rng.uniform(-5, 5, B), rng.uniform(-5, 5, B), rng.uniform(-np.pi, np.pi, B)], axis=1)
    velocity = rng.normal(size=(B, 2, R)).astype(np.float32)
    return ranges, states, velocity


def main():
    ref = refload.load_reference()
    planner = ref.neupan  # the class; the two functions do not touch self
    fn_point = lambda st, sc, off, ar, ds: planner.scan_to_point(None, st, sc, off, ar, ds)
    fn_velocity = lambda st, sc, off, ar, ds: planner.scan_to_point_velocity(None, st, sc, off, ar, ds)
    out = {}
    for seed, (name, (B, R, scan, off, ar, ds, mp, vm)) in enumerate(CASES.items()):
        ranges, states, velocity = inputs(name, B, R, scan, 100 + seed)
        pts, vel, cnt = oscan.scan_batch(states, ranges, scan, off, ar, ds, mp, velocity if vm else None, vm, fn_point=fn_point, fn_velocity=fn_velocity)
        out[f"{name}.ranges"], out[f"{name}.states"], out[f"{name}.velocity"] = ranges, states, velocity
        out[f"{name}.points"], out[f"{name}.counts"] = pts, cnt
        if vel is not None:
            out[f"{name}.vel_out"] = vel
        print(name, "counts", cnt)
    np.savez_compressed(os.path.join(HERE, "ref_scan.npz"), **out)


if __name__ == "__main__":
    main()
