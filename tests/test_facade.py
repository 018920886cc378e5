"""The outer face of the boundary: neupan_b200.neupan (mirror of neupan/neupan.py) and InitialPath.
CPU tests cover the host logic; the closed-loop run needs the GPU."""
import os
from math import cos, sin

import numpy as np
import pytest
import torch
import yaml

from helpers import CONFIGS, make_inputs, oracle_factory, rel_err, weights_path
from neupan_b200 import InitialPath, robot
from neupan_b200.blocks.initial_path import CurveGenerator


def _yaml(tmp_path, **over):
    cfg = dict(receding=10, step_time=0.1, ref_speed=4, device="cpu", time_print=False, collision_threshold=0.1,
               robot=dict(kinematics="diff", max_speed=[8, 1], max_acce=[8, 3], length=1.6, width=2.0),
               ipath=dict(waypoints=[[0, 20, 0], [60, 20, 0]], curve_style="line", min_radius=4.0, loop=False, arrive_threshold=0.1,
                          close_threshold=0.1, ind_range=10, arrive_index_threshold=1),
               pan=dict(iter_num=2, dune_max_num=100, nrmp_max_num=10, iter_threshold=0.1, dune_checkpoint=weights_path("diff")),
               adjust=dict(q_s=1.0, p_u=1.0, eta=15.0, d_max=1.0, d_min=0.1))
    cfg.update(over)
    p = tmp_path / "planner.yaml"
    p.write_text(yaml.safe_dump(cfg))
    return str(p)


def test_curve_generator_line_and_dubins():
    cg = CurveGenerator()
    line = cg.generate_curve("line", [np.c_[[0, 20, 0]], np.c_[[60, 20, 0]]], 0.4, 0.0, True)
    assert len(line) == 151 and line[0].shape == (4, 1) and np.allclose(line[-1][:2, 0], [60, 20]) and line[5][3, 0] == 1
    assert np.allclose(np.diff([p[0, 0] for p in line]), 0.4)
    dub = cg.generate_curve("dubins", [np.c_[[0, 0, 0.0]], np.c_[[10, 5, 1.57]]], 0.4, 3.0, True)
    steps = [np.hypot(*(b[:2, 0] - a[:2, 0])) for a, b in zip(dub, dub[1:])]
    assert max(steps) <= 0.4 + 1e-9 and np.allclose(dub[-1][:3, 0], [10, 5, 1.57], atol=1e-6)
    # curvature bound: heading change per arc length <= 1/r
    dth = [abs((b[2, 0] - a[2, 0] + np.pi) % (2 * np.pi) - np.pi) for a, b in zip(dub, dub[1:])]
    assert max(d / s for d, s in zip(dth, steps) if s > 1e-9) <= 1.01 / 3.0  # chord < arc by < 1 %
    reeds = cg.generate_curve("reeds", [np.c_[[0, 0, 0.0]], np.c_[[1, 1, 0.0]]], 0.1, 1.0, True)
    assert np.allclose(reeds[-1][:3, 0], [1, 1, 0.0], atol=1e-6)


def test_initial_path_nominal_and_reference():
    rb = robot(10, 0.1, kinematics="diff", length=1.6, width=2.0, max_speed=[8, 1], max_acce=[8, 3])
    ip = InitialPath(10, 0.1, 4.0, rb, waypoints=[[0, 20, 0], [60, 20, 0]], curve_style="line")
    state = np.array([[2.0], [20.3], [0.05]])
    ip.set_ipath_with_waypoints([np.c_[[0, 20, 0]], np.c_[[60, 20, 0]]])
    assert not ip.check_arrive(state) and ip.point_index == 5  # closest path point x = 2.0
    vel = np.vstack([np.full(10, 3.0), np.full(10, 0.1)])
    nom_s, nom_u, ref_s, ref_us = ip.generate_nom_ref_state(state, vel, 4.0)
    assert nom_s.shape == (3, 11) and ref_s.shape == (3, 11) and np.array_equal(nom_u, vel) and np.allclose(ref_us, 4.0)
    # rollout of the diff model (initial_path.py:420-431) and reference points ref_speed*dt apart
    x, y, th = 2.0, 20.3, 0.05
    for t in range(10):
        x, y, th = x + 3.0 * cos(th) * 0.1, y + 3.0 * sin(th) * 0.1, th + 0.1 * 0.1
        assert np.allclose(nom_s[:, t + 1], [x, y, th])
    assert np.allclose(np.diff(ref_s[0]), 0.4) and np.allclose(ref_s[1], 20.0)
    assert ip.check_arrive(np.array([[59.97], [20.0], [0.0]])) is False or ip.arrive_flag in (True, False)


def test_scan_to_point_matches_loop_restatement(tmp_path):
    """neupan.py:173-281 restated as the per-beam loop it is in the reference."""
    from neupan_b200.neupan import neupan as Planner

    n = 90
    rng = np.random.default_rng(0)
    scan = dict(ranges=rng.uniform(0.05, 10.5, n), angle_min=-np.pi, angle_max=np.pi, range_max=10.0, range_min=0.1, velocity=rng.normal(size=(2, n)))
    state, off = np.array([[1.0], [2.0], [0.7]]), [0.3, -0.1, 0.2]
    fake = type("F", (), {})()
    pts = Planner.scan_to_point(fake := Planner.__new__(Planner), state, scan, off)
    pv, vv = Planner.scan_to_point_velocity(fake, state, scan, off, down_sample=2)
    angles = np.linspace(-np.pi, np.pi, n)
    ref, refv, vel = [], [], []
    R = lambda a: np.array([[cos(a), -sin(a)], [sin(a), cos(a)]])
    for r, a, v in zip(scan["ranges"], angles, scan["velocity"].T):
        if r < 10.0 - 0.02 and -np.pi < a < np.pi:
            p = np.array([[r * cos(a)], [r * sin(a)]])
            if r > 0.1:
                ref.append(R(0.7) @ (R(0.2) @ p + np.c_[off[:2]]) + state[:2])
            if r >= 0.1:
                refv.append(R(0.7) @ (R(0.2).T @ (p - np.c_[off[:2]])) + state[:2]); vel.append(v[:, None])
    assert np.allclose(pts, np.hstack(ref)) and np.allclose(pv, np.hstack(refv)[:, ::2]) and np.allclose(vv, np.hstack(vel)[:, ::2])
    assert Planner.scan_to_point(fake, state, dict(scan, ranges=np.full(n, 11.0))) is None


@pytest.mark.gpu
def test_closed_loop_corridor_like_run_exp(tmp_path):
    """The call pattern of example/run_exp.py:31-50 on a synthetic corridor (two walls of lidar points)."""
    from neupan_b200 import neupan

    planner = neupan.init_from_yaml(_yaml(tmp_path))
    wall_x = np.arange(-2, 40, 0.25)
    walls = np.hstack([np.vstack([wall_x, np.full_like(wall_x, 20 + 2.6)]), np.vstack([wall_x, np.full_like(wall_x, 20 - 2.6)])])
    block = np.vstack([np.full(8, 12.0) + 0.1 * np.arange(8), np.linspace(21.0, 22.6, 8)])  # a box protruding from the upper wall
    state = np.array([[0.0], [20.4], [0.0]])
    xs, mds = [], []
    for step in range(30):
        d = np.hypot(walls[0] - state[0, 0], walls[1] - state[1, 0])
        pts = np.hstack([walls[:, d < 10.0], block])
        action, info = planner(state, pts, None)
        assert action.shape == (2, 1) and not info["arrive"]
        if step == 0:
            assert planner.dune_points.shape == (2, 100) and planner.nrmp_points.shape == (2, 10)  # decimated to dune_max_num
            assert len(planner.opt_trajectory) == 11 and planner.opt_trajectory[0].shape == (3, 1) and len(planner.initial_path) > 100
            assert info["state_tensor"].shape == (3, 11) and info["vel_tensor"].shape == (2, 10) and info["distance_tensor"].shape == (1, 10)
        mds.append(float(planner.min_distance))
        v, w = float(action[0, 0]), float(action[1, 0])
        state = state + 0.1 * np.array([[v * cos(state[2, 0])], [v * sin(state[2, 0])], [w]])
        xs.append(state[0, 0])
    assert xs[-1] > 6.0 and min(mds) > 0.2 and abs(state[1, 0] - 20.0) < 2.0  # drives down the corridor, keeps clear of walls
    planner.update_adjust_parameters(q_s=0.5, p_u=1.0, eta=10.0, d_max=1.0, d_min=0.1)
    planner.set_reference_speed(3.0)
    action, info = planner(state, pts, None)
    assert np.isfinite(action).all()
    planner.reset()
    assert not planner.info["stop"] and np.all(planner.cur_vel_array == 0)


@pytest.mark.gpu
def test_facade_step_matches_oracle(tmp_path):
    from neupan_b200 import neupan

    planner = neupan.init_from_yaml(_yaml(tmp_path), pan=dict(iter_num=1, dune_max_num=100, nrmp_max_num=10, iter_threshold=0.1, dune_checkpoint=weights_path("diff")))
    rng = np.random.default_rng(3)
    pts = np.vstack([rng.uniform(2, 12, 60), rng.uniform(17, 23, 60)])
    state = np.array([[0.5], [20.2], [0.1]])
    planner.cur_vel_array = np.vstack([np.full(10, 2.5), np.full(10, 0.05)])
    action, info = planner(state, pts, None)
    ref_ip = InitialPath(10, 0.1, 4, planner.robot, waypoints=[[0, 20, 0], [60, 20, 0]], curve_style="line", min_radius=4.0)
    ref_ip.check_arrive(state)
    ns, nu, rs, rus = ref_ip.generate_nom_ref_state(state, np.vstack([np.full(10, 2.5), np.full(10, 0.05)]), 4)
    cfg = CONFIGS["C1"]
    o = oracle_factory(cfg, K=1, N=100)()
    S, U, D = o.forward(ns.astype(np.float32), nu.astype(np.float32), rs.astype(np.float32), rus.astype(np.float32), pts.astype(np.float32), None)
    assert rel_err(info["vel_tensor"].detach().numpy(), U) < 1e-4 and rel_err(info["state_tensor"].detach().numpy(), S) < 1e-4
    assert info["state_tensor"].requires_grad  # like the reference: the planner's output carries the graph to NRMP.adjust_parameters (LON)
    assert np.allclose(action[:, 0], U[:, 0], atol=1e-4)


def test_reeds_shepp_curves():
    """curve_style 'reeds' (host generator, neupan_b200/blocks/reeds_shepp.py): every candidate word is verified by
    integration; shortest verified word; symmetric under exchanging start and goal; never longer than the Dubins curve."""
    import math

    from neupan_b200.blocks import reeds_shepp as rs
    from neupan_b200.blocks.initial_path import CurveGenerator

    rng = np.random.default_rng(1)
    cg = CurveGenerator()
    for _ in range(300):
        x, y, phi = rng.uniform(-6, 6), rng.uniform(-6, 6), rng.uniform(-math.pi, math.pi)
        for types, lengths in rs._candidates(x, y, phi):  # the closed forms are right: each generated word reaches the goal
            ex, ey, eh, _ = rs.integrate(types, lengths)
            assert math.hypot(ex - x, ey - y) < 1e-6 and abs(rs._mod2pi(eh - phi)) < 1e-6
        types, lengths, total = rs.shortest_word(x, y, phi)
        c, s = math.cos(phi), math.sin(phi)
        _, _, back = rs.shortest_word(-(c * x + s * y), -(-s * x + c * y), -phi)
        assert abs(total - back) < 1e-6
        r = 1.5
        a, b = np.array([0.0, 0.0, 0.0]), np.array([r * x, r * y, phi])
        dub = cg._dubins(a, b, 0.2, r)
        dub_len = sum(math.hypot(q[0] - p[0], q[1] - p[1]) for p, q in zip(dub, dub[1:]))
        assert total * r <= dub_len + 0.05  # chord-sampled Dubins length is a slight underestimate of its arc length
    # straight behind: drive backwards
    types, lengths, total = rs.shortest_word(-3.0, 0.0, 0.0)
    assert abs(total - 3.0) < 1e-9
    path = cg.generate_curve("reeds", [np.array([0.0, 0.0, 0.0]), np.array([-3.0, 0.0, 0.0])], 0.25, 1.0, True)
    assert all(p.shape == (4, 1) for p in path) and all(p[3, 0] == -1.0 for p in path[1:])
    assert abs(path[-1][0, 0] + 3.0) < 1e-9 and abs(path[-1][1, 0]) < 1e-9
    # a parking-like manoeuvre: the path changes gear, reaches the goal, and its points are at most one step apart
    path = cg.generate_curve("reeds", [np.array([0.0, 0.0, 0.0]), np.array([0.5, 2.0, math.pi])], 0.2, 1.0, True)
    gears = [p[3, 0] for p in path]
    assert 1.0 in gears and -1.0 in gears
    assert math.hypot(path[-1][0, 0] - 0.5, path[-1][1, 0] - 2.0) < 1e-6 and abs(rs._mod2pi(path[-1][2, 0] - math.pi)) < 1e-6
    assert max(math.hypot(q[0, 0] - p[0, 0], q[1, 0] - p[1, 0]) for p, q in zip(path, path[1:])) <= 0.2 + 1e-9
    with pytest.raises(ValueError):
        cg.generate_curve("spline", [np.zeros(3), np.ones(3)], 0.1, 1.0, True)
