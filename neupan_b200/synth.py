"""Seeded synthetic workloads for the five BASELINE.json configurations.

One environment = one robot with its own start pose, nominal control sequence, reference line
and obstacle point cloud.  Everything is generated on the CPU with numpy (PCG64, seed
1234 + config id) so the GPU run, the CPU oracle and the golden fixtures see identical bits.

Shapes (batch-leading, float32): nom_s (B,3,T+1), nom_u (B,2,T), ref_s (B,3,T+1), ref_us (B,T),
points (B,2,N), velocities (B,2,N) or None.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .robot import robot


@dataclass
class WorkloadConfig:
    name: str
    cfg_id: int
    robot_kwargs: dict
    adjust: dict
    model: str  # which shipped ObsPointNet checkpoint: diff | acker | polygon
    B: int
    T: int
    N: int
    K: int
    dynamic: bool
    M: int = 10
    dt: float = 0.1
    ref_speed: float = 4.0
    extra: dict = field(default_factory=dict)

    def make_robot(self) -> robot:
        return robot(self.T, self.dt, **self.robot_kwargs)


# adjust / robot values from example/<scenario>/<kin>/planner.yaml of the reference; defaults
# bk=0.1, ro_obs=400 from neupan/blocks/pan.py:70-82.
CONFIGS = {
    "C1": WorkloadConfig("corridor/diff", 1, dict(kinematics="diff", max_speed=[8, 1], max_acce=[8, 3], length=1.6, width=2.0),
                         dict(q_s=1.0, p_u=1.0, eta=15.0, d_max=1.0, d_min=0.1), "diff", B=1, T=10, N=100, K=2, dynamic=False),
    "C2": WorkloadConfig("dyna_obs/acker", 2, dict(kinematics="acker", max_speed=[8, 1], max_acce=[8, 1], length=4.6, width=1.6, wheelbase=3),
                         dict(q_s=0.5, p_u=0.3, eta=15.0, d_max=1.0, d_min=0.1), "acker", B=256, T=10, N=200, K=10, dynamic=True),
    "C3": WorkloadConfig("non_obs/acker", 3, dict(kinematics="acker", max_speed=[8, 1], max_acce=[8, 0.5], length=4.6, width=1.6, wheelbase=3),
                         dict(q_s=1.0, p_u=1.0, eta=15.0, d_max=1.0, d_min=0.1), "acker", B=1024, T=10, N=500, K=10, dynamic=False),
    "C4": WorkloadConfig("dyna_non_obs/diff", 4, dict(kinematics="diff", max_speed=[8, 3], max_acce=[8, 3], length=1.6, width=2.0),
                         dict(q_s=0.5, p_u=1.0, eta=15.0, d_max=1.0, d_min=0.1, bk=1.0), "diff", B=4096, T=10, N=500, K=10, dynamic=True),
    "C5": WorkloadConfig("polygon_robot/omni", 5, dict(kinematics="omni", max_speed=[8, 6.28], max_acce=[2, 2],
                                                       vertices=[[-0.8, -1.0], [-1.8, 1.0], [1.8, 1.0], [0.8, -1.0]]),
                         dict(q_s=1.0, p_u=0.5, eta=15.0, d_max=1.0, d_min=0.1), "polygon", B=2048, T=15, N=1000, K=15, dynamic=False),
}


def rollout(kin: str, s0: np.ndarray, u: np.ndarray, dt: float, L) -> np.ndarray:
    """Nonlinear forward rollout of the nominal controls, the models of
    neupan/blocks/initial_path.py:401-444.  s0 (B,3), u (B,2,T) -> (B,3,T+1) float64."""
    B, _, T = u.shape
    s = np.zeros((B, 3, T + 1))
    s[:, :, 0] = s0
    for t in range(T):
        x, y, th = s[:, 0, t], s[:, 1, t], s[:, 2, t]
        v, w = u[:, 0, t], u[:, 1, t]
        if kin == "diff":
            ds = np.stack([v * np.cos(th), v * np.sin(th), w], 1)
        elif kin == "acker":
            ds = np.stack([v * np.cos(th), v * np.sin(th), v * np.tan(w) / L], 1)
        else:  # omni: u = (speed, heading command)
            ds = np.stack([v * np.cos(w), v * np.sin(w), np.zeros_like(v)], 1)
        s[:, :, t + 1] = s[:, :, t] + ds * dt
    return s


def make_inputs(cfg: WorkloadConfig, B: int | None = None, N: int | None = None, seed: int | None = None, env_offset: int = 0,
                scene: str = "annulus"):
    """Returns a dict of float32 arrays for envs [env_offset, env_offset+B) of the config.

    scene="annulus"   : SURVEY.md 8d -- N points uniform in the annulus 1.5..10 m around the start
                        (dense clutter: every path is blocked, the hinge terms are always active).
    scene="obstacles" : lidar-like -- the N points lie on the boundaries of 4..10 discs (radius
                        0.3..1 m, centres 3.5..10 m from the start, some near the reference line);
                        all points of one disc share one velocity.  Leaves free space, so the PAN
                        iteration settles the way it does in the reference's example scenarios.

    Each env draws from its own child stream (SeedSequence spawn key = env index), so any
    sub-range -- a GPU shard, a CPU-baseline sample -- reproduces exactly the same envs.
    """
    B = cfg.B if B is None else B
    N = cfg.N if N is None else N
    T, dt = cfg.T, cfg.dt
    rb = cfg.make_robot()
    kin = rb.kinematics
    G, h = rb.G, rb.h.reshape(-1)
    gnorm = np.linalg.norm(G, axis=1)
    base = 1234 + cfg.cfg_id if seed is None else seed

    nom_s = np.zeros((B, 3, T + 1)); nom_u = np.zeros((B, 2, T)); ref_s = np.zeros((B, 3, T + 1))
    pts = np.zeros((B, 2, N)); vel = np.zeros((B, 2, N))
    for b in range(B):
        rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(base, spawn_key=(env_offset + b,))))
        x0, y0 = rng.uniform(-5, 5, 2)
        th0 = rng.uniform(-np.pi, np.pi)
        v = rng.uniform(1.0, 4.0) + rng.normal(0, 0.02, T)
        w = rng.uniform(-0.3, 0.3) + rng.normal(0, 0.01, T)
        if kin == "omni":
            w = th0 + w  # heading command around the start heading
        nom_u[b, 0], nom_u[b, 1] = v, w
        k = np.arange(T + 1) * cfg.ref_speed * dt
        ref_s[b, 0], ref_s[b, 1], ref_s[b, 2] = x0 + k * np.cos(th0), y0 + k * np.sin(th0), th0
        nom_s[b, :, 0] = (x0, y0, th0)
        # points: uniform in the annulus 1.5 <= r <= 10 around the start, outside the robot polygon inflated by 0.2 m
        need = np.ones(N, bool)
        loc = np.zeros((2, N))
        c, s_ = np.cos(th0), np.sin(th0)
        if scene == "obstacles":
            n_obs = int(rng.integers(4, 11))
            rc = np.sqrt(rng.uniform(3.5 ** 2, 10.0 ** 2, n_obs)); ac = rng.uniform(-np.pi, np.pi, n_obs)
            ac[: n_obs // 2] = rng.normal(0.0, 0.5, n_obs // 2)  # half of them ahead, around the reference line
            rad = rng.uniform(0.3, 1.0, n_obs)
            which = rng.integers(0, n_obs, N)
            ang = rng.uniform(-np.pi, np.pi, N)
            loc = np.stack([rc[which] * np.cos(ac[which]) + rad[which] * np.cos(ang), rc[which] * np.sin(ac[which]) + rad[which] * np.sin(ang)])
            ov = rng.uniform(-1, 1, (2, n_obs))
            obs_vel = ov[:, which]
            need[:] = False
        while need.any():
            n = int(need.sum())
            r = np.sqrt(rng.uniform(1.5 ** 2, 10.0 ** 2, n))
            a = rng.uniform(-np.pi, np.pi, n)
            cand = np.stack([r * np.cos(a), r * np.sin(a)])  # robot frame
            inside = np.all(G @ cand - h[:, None] <= 0.2 * gnorm[:, None], axis=0)
            idx = np.flatnonzero(need)
            loc[:, idx[~inside]] = cand[:, ~inside]
            need[idx[~inside]] = False
        pts[b, 0] = x0 + c * loc[0] - s_ * loc[1]
        pts[b, 1] = y0 + s_ * loc[0] + c * loc[1]
        vel[b] = obs_vel if scene == "obstacles" else rng.uniform(-1, 1, (2, N))
    nom_s = rollout(kin, nom_s[:, :, 0], nom_u, dt, rb.L)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(nom_s=f32(nom_s), nom_u=f32(nom_u), ref_s=f32(ref_s), ref_us=f32(np.full((B, T), cfg.ref_speed)),
                points=f32(pts), velocities=f32(vel) if cfg.dynamic else None)
