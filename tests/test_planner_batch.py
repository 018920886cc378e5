"""PlannerBatch: one control step of B robots on the device = InitialPathBatch.step -> scan_to_points -> PAN.forward
(neupan.forward, neupan/neupan.py:104-167, for a batch).  Checked as the composition of its three separately verified stages
and in a short closed loop."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from helpers import CONFIGS, weights_path

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_ipath", os.path.join(HERE, "golden", "make_golden_ipath.py"))
mgi = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mgi)


def _planner(cfg, B, N):
    from neupan_b200 import PlannerBatch

    return PlannerBatch(B, cfg.T, cfg.dt, 4.0, robot_kwargs=cfg.robot_kwargs, adjust_kwargs=dict(cfg.adjust),
                        pan_kwargs=dict(iter_num=2, dune_max_num=N, nrmp_max_num=cfg.M, dune_checkpoint=weights_path(cfg.model), iter_threshold=0.0, max_points=N))


def _scans(B, R, seed):
    rng = np.random.default_rng(seed)
    scan = dict(angle_min=-np.pi, angle_max=np.pi, range_min=0.1, range_max=10.0)
    ranges = rng.uniform(2.5, 11.0, size=(B, R)).astype(np.float32)  # nothing closer than 2.5 m: no emergency stop
    vel = rng.uniform(-0.5, 0.5, size=(B, 2, R)).astype(np.float32)
    return scan, ranges, vel


def test_one_step_equals_the_three_stages():
    from neupan_b200 import PAN, InitialPathBatch, scan_to_points

    cfg, B, R, N = CONFIGS["C4"], 5, 300, 128
    paths = [mgi.make_path(60, 0.4 + 0.0001 * (b - 2), None, 0.01 * (b - 2)) for b in range(B)]
    states = np.tile(np.array([0.05, -0.03, 0.28]), (B, 1)) + np.arange(B)[:, None] * 0.01
    scan, ranges, vel = _scans(B, R, 3)
    pl = _planner(cfg, B, N)
    pl.set_initial_paths(paths)
    action, info = pl.forward(torch.from_numpy(states), torch.from_numpy(ranges), scan, scan_velocity=torch.from_numpy(vel))
    assert action.shape == (B, 2) and torch.isfinite(action).all() and not info["arrive"].any() and not info["stop"].any()
    # the same step from the parts
    ipb = InitialPathBatch(cfg.T, cfg.dt, "diff", max_envs=B)
    ipb.set_initial_paths(paths)
    nom_s, nom_u, ref_s, ref_us, arrived = ipb.step(torch.from_numpy(states), torch.zeros(B, 2, cfg.T), 4.0)
    pts, pv, cnt = scan_to_points(torch.from_numpy(states), torch.from_numpy(ranges), scan, max_points=N, velocity=torch.from_numpy(vel))
    pan = PAN(cfg.T, cfg.dt, cfg.make_robot(), iter_num=2, dune_max_num=N, nrmp_max_num=cfg.M, dune_checkpoint=weights_path(cfg.model), iter_threshold=0.0,
              adjust_kwargs=dict(cfg.adjust), max_envs=B, max_points=N)
    S, U, D = pan(nom_s, nom_u, ref_s, ref_us, pts, pv, cnt)
    assert torch.equal(info["ref_state_tensor"], ref_s) and torch.equal(info["num_points"], cnt)
    assert torch.equal(info["vel_tensor"], U) and torch.equal(info["state_tensor"], S) and torch.equal(action, U[:, :, 0])
    assert torch.equal(pl.cur_vel, U)  # the velocity memory of the next step (neupan.py:137)
    pl.close(); ipb.close(); pan.close()


def test_short_closed_loop_makes_progress_along_the_paths():
    cfg, B, R, N = CONFIGS["C4"], 4, 200, 64
    paths = [mgi.make_path(80, 0.4, None, 0.008 * (b - 1.5)) for b in range(B)]
    pl = _planner(cfg, B, N)
    pl.set_initial_paths(paths)
    states = np.tile(np.array([0.0, 0.0, 0.3]), (B, 1))
    scan, ranges, vel = _scans(B, R, 4)
    start = states.copy()
    for k in range(8):
        action, info = pl.forward(torch.from_numpy(states), torch.from_numpy(ranges), scan)
        a = action.cpu().numpy().astype(np.float64)
        assert np.isfinite(a).all() and (a[:, 0] >= -1e-3).all() and (a[:, 0] <= 8.0 + 1e-3).all()   # max_speed of the config
        states = states + cfg.dt * np.stack([a[:, 0] * np.cos(states[:, 2]), a[:, 0] * np.sin(states[:, 2]), a[:, 1]], 1)  # diff drive
    st = pl.ipath.read_state()
    assert (st["point_index"].cpu().numpy() >= 1).all()                   # the closest path point moved forward
    assert (np.linalg.norm(states[:, :2] - start[:, :2], axis=1) > 0.5).all()
    assert (pl.pan.status == 0).all()
    pl.close()
