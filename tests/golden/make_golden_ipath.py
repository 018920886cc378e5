"""Golden vectors for the per-step InitialPath work, produced by the REFERENCE class itself
(/root/reference/neupan/blocks/initial_path.py, imported through oracle/refload.py; paths enter through set_initial_path,
so the absent gctl curve generator is not needed): closed-loop runs of check_arrive + generate_nom_ref_state.

    python tests/golden/make_golden_ipath.py        # needs /root/reference; writes tests/golden/ref_ipath.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refload  # noqa: E402

T, DT, REF_SPEED = 10, 0.1, 4.0
# name: (kinematics, L, loop, point spacing, index of the gear switch or None, curvature per point, number of points)
SCENARIOS = {
    "diff_exact": ("diff", None, False, 0.4, None, 0.02, 40),
    "diff_short_spacing": ("diff", None, False, 0.3999, 25, 0.03, 45),      # int(0.4 / 0.3999) = 1: index jumps
    "diff_long_spacing": ("diff", None, False, 0.4001, 20, -0.02, 40),      # 0.4 < interval: circle / segment walk
    "diff_dense": ("diff", None, False, 0.13, None, 0.01, 120),              # jumps of 3 points
    "diff_sparse": ("diff", None, False, 0.9, 18, 0.05, 30),                 # walk across long segments
    "acker_two_gears": ("acker", 3.0, False, 0.4, 30, 0.015, 55),
    "acker_walk": ("acker", 3.0, False, 0.55, None, -0.03, 35),
    "omni_exact": ("omni", None, False, 0.4, None, 0.04, 40),
    "diff_loop": ("diff", None, True, 0.4, None, 0.1, 30),
    "acker_loop_gears": ("acker", 3.0, True, 0.45, 12, 0.06, 26),
}
STEPS = 70


def make_path(n, step, gearsplit, curve):
    pts, x, y, th = [], 0.0, 0.0, 0.3
    for i in range(n):
        g = 1.0 if (gearsplit is None or i < gearsplit) else -1.0
        pts.append(np.array([[x], [y], [th], [g]]))
        x += step * np.cos(th) * g
        y += step * np.sin(th) * g
        th += curve
    return pts


def drive(ip, path, seed, record):
    """closed loop: the next state is the first reference state plus noise; velocities are float32-representable"""
    rng = np.random.default_rng(seed)
    ip.set_initial_path([p.copy() for p in path])
    state = np.array([[0.05], [-0.03], [0.25]])
    for k in range(STEPS):
        vel = rng.uniform(-1, 1, (2, T))
        vel[0] += 3.0
        vel = vel.astype(np.float32).astype(np.float64)  # what PAN hands back (float32), held in the float64 array numpy computes with
        arrived = bool(ip.check_arrive(state))
        out = None if arrived else ip.generate_nom_ref_state(state, vel, REF_SPEED)
        record(k, state, vel, arrived, out, ip.point_index, ip.curve_index)
        if arrived:
            break
        state = out[2][:, 1:2].copy() + rng.normal(0, 0.01, (3, 1))


def reference_instance(kin, L, loop):
    refload.load_reference()
    from neupan.blocks.initial_path import InitialPath

    class Robot:
        pass

    rb = Robot()
    rb.kinematics, rb.L = kin, L
    ip = InitialPath.__new__(InitialPath)
    ip.T, ip.dt, ip.ref_speed, ip.robot, ip.loop = T, DT, REF_SPEED, rb, loop
    ip.arrive_threshold, ip.close_threshold, ip.ind_range, ip.arrive_index_threshold, ip.arrive_flag = 0.1, 0.1, 10, 1, False
    ip.initial_path = None
    return ip


def main():
    out = {}
    for seed, (name, (kin, L, loop, step, split, curve, n)) in enumerate(SCENARIOS.items()):
        rows = []
        ip = reference_instance(kin, L, loop)

        def record(k, state, vel, arrived, o, pi, ci):
            z = (np.zeros((3, T + 1)), vel, np.zeros((3, T + 1)), np.zeros(T)) if o is None else o
            rows.append((state.copy(), vel.copy(), float(arrived), z[0].copy(), z[2].copy(), np.asarray(z[3], float).copy(), pi, ci))

        drive(ip, make_path(n, step, split, curve), 500 + seed, record)
        out[f"{name}.states"] = np.stack([r[0][:, 0] for r in rows])
        out[f"{name}.vel"] = np.stack([r[1] for r in rows])
        out[f"{name}.arrived"] = np.array([r[2] for r in rows])
        out[f"{name}.nom_s"] = np.stack([r[3] for r in rows])
        out[f"{name}.ref_s"] = np.stack([r[4] for r in rows])
        out[f"{name}.ref_us"] = np.stack([r[5] for r in rows])
        out[f"{name}.point_index"] = np.array([r[6] for r in rows])
        out[f"{name}.curve_index"] = np.array([r[7] for r in rows])
        out[f"{name}.final_path"] = np.hstack([p for c in ip.curve_list for p in c]).T  # the path as the reference left it (mutated headings)
        print(name, "steps", len(rows), "arrived", bool(rows[-1][2]), "interval", ip.interval, "curves", len(ip.curve_list))
    np.savez_compressed(os.path.join(HERE, "ref_ipath.npz"), **out)


if __name__ == "__main__":
    main()
