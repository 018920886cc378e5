"""Shared helpers for the parity tests (CPU oracle side)."""
import os

import numpy as np

from neupan_b200.synth import CONFIGS, make_inputs  # noqa: F401
from oracle import dune as od, nrmp as onr, pan as op

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def weights_path(model: str) -> str:
    return os.path.join(GOLDEN, f"weights_{model}.npz")


def robot_spec(cfg):
    rb = cfg.make_robot()
    return rb, onr.RobotSpec(rb.kinematics, rb.G, rb.h, rb.max_speed.reshape(-1), rb.max_acce.reshape(-1), cfg.dt, rb.L)


def oracle_factory(cfg, K=None, iter_threshold=0.0, solver="ipm", N=None, adjust=None, M=None):
    rb, spec = robot_spec(cfg)
    w = od.load_weights(weights_path(cfg.model))
    adj = onr.Adjust(**(adjust or cfg.adjust))
    return lambda: op.OraclePAN(spec, w, T=cfg.T, iter_num=cfg.K if K is None else K, dune_max_num=cfg.N if N is None else N,
                                nrmp_max_num=cfg.M if M is None else M, iter_threshold=iter_threshold, adjust=adj, solver=solver)


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(floor, max|b|): relative to the scale of the reference tensor."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(floor, np.abs(b).max()))
