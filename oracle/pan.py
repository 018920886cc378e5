"""CPU restatement of the PAN loop (TEST INFRASTRUCTURE, see oracle/__init__.py).

neupan/blocks/pan.py:109-147 (loop), :215-243 (stop criterion, state persisting across calls).
One ``OraclePAN`` = one environment, exactly like one reference ``PAN`` object; batches are
Python loops over instances (``run_batch``).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import dune as odune
from . import ipm as oipm
from . import nrmp as onrmp


class OraclePAN:
    def __init__(self, robot: onrmp.RobotSpec, weights: dict, T=10, iter_num=2, dune_max_num=100, nrmp_max_num=10,
                 iter_threshold=0.1, adjust: onrmp.Adjust | None = None, solver="ipm"):
        self.robot, self.w = robot, weights
        self.T, self.dt = T, robot.dt
        self.iter_num, self.iter_threshold = iter_num, iter_threshold
        self.dune_max_num, self.nrmp_max_num = dune_max_num, nrmp_max_num
        self.no_obs = nrmp_max_num == 0 or dune_max_num == 0  # pan.py:85
        self.adjust = adjust or onrmp.Adjust()
        self.G = torch.from_numpy(robot.G).float()
        self.h = torch.from_numpy(robot.h.reshape(-1, 1)).float()
        self.solver = solver
        self.current = [None, None, None, None]  # pan.py:100-105
        self.min_distance = float("inf")
        self.dune_points = None
        self.nrmp_points = None
        self.iters_run = 0
        self.t_dune = 0.0
        self.t_nrmp = 0.0
        self.fallbacks = 0
        self.trace = []

    def _solve(self, prob):
        if self.solver == "highs":
            S, U, D, ok = onrmp.solve_highs(prob)
            if not ok:
                raise RuntimeError("HiGHS did not report optimal")
            return S, U, D
        try:
            S, U, D, _ = oipm.solve_ipm(prob)
        except (oipm.IpmFailure, np.linalg.LinAlgError, ValueError):
            # second solver of the same program (HiGHS active set on the lifted form); counted so that
            # callers can report how often the oracle of record had to be replaced
            self.fallbacks += 1
            S, U, D, ok = onrmp.solve_highs(prob)
            if not ok:
                raise RuntimeError("neither the interior point method nor HiGHS solved this NRMP instance")
        return S, U, D

    def forward(self, nom_s, nom_u, ref_s, ref_us, obs_points=None, point_velocities=None, keep_trace=False):
        """All arguments float32 numpy / torch of the reference's unbatched shapes (pan.py:109-126)."""
        t32 = lambda a: None if a is None else torch.as_tensor(np.asarray(a), dtype=torch.float32)
        nom_s, nom_u, ref_s, ref_us = t32(nom_s), t32(nom_u), t32(ref_s), t32(ref_us)
        obs_points, point_velocities = t32(obs_points), t32(point_velocities)
        nom_d = None
        self.iters_run = 0
        self.trace = []
        for _ in range(self.iter_num):
            t0 = time.perf_counter()
            if obs_points is not None and not self.no_obs:
                p0_list, R_list, p_list = odune.point_flow(nom_s, obs_points, point_velocities, self.T, self.dt, self.dune_max_num)
                mu_list, lam_list, sp_list, md, _ = odune.dune_forward(self.w, self.G, self.h, p0_list, R_list, p_list)
                self.min_distance = float(md)
                self.dune_points = p_list[0].numpy()
                self.nrmp_points = sp_list[0][:, :self.nrmp_max_num].numpy()  # nrmp.py:135-138
                fa, fb = odune.nrmp_coefficients(self.h, mu_list, lam_list, sp_list, self.T, self.nrmp_max_num)
                fa, fb = fa.numpy(), fb.numpy()
            else:
                mu_list, lam_list, fa, fb = [], [], None, None
            t1 = time.perf_counter()
            prob = onrmp.build_problem(self.robot, self.adjust, nom_s.numpy(), nom_u.numpy(), ref_s.numpy(), ref_us.numpy(),
                                       fa, fb, 0 if self.no_obs else self.nrmp_max_num)
            S, U, D = self._solve(prob)
            t2 = time.perf_counter()
            self.t_dune += t1 - t0
            self.t_nrmp += t2 - t1
            nom_s = torch.from_numpy(S).float()  # nrmp.py:145-148 cast to float32
            nom_u = torch.from_numpy(U).float()
            nom_d = None if D is None else torch.from_numpy(D).float()
            self.iters_run += 1
            if keep_trace:
                self.trace.append(dict(S=nom_s.numpy().copy(), U=nom_u.numpy().copy(), D=None if nom_d is None else nom_d.numpy().copy(),
                                       fa=fa, fb=fb, min_distance=self.min_distance))
            if self._stop(nom_s, nom_u, mu_list, lam_list):
                break
        return nom_s.numpy(), nom_u.numpy(), (None if nom_d is None else nom_d.numpy())

    def _stop(self, nom_s, nom_u, mu_list, lam_list) -> bool:
        """pan.py:215-243."""
        if self.current[0] is None:
            self.current = [nom_s, nom_u, mu_list, lam_list]
            return False
        if len(mu_list) == 0 or len(self.current[2]) == 0:
            diff = torch.norm(nom_s - self.current[0]) ** 2 + torch.norm(nom_u - self.current[1]) ** 2
        else:
            en = min(mu_list[0].shape[1], self.current[2][0].shape[1], self.nrmp_max_num)
            mu_diff = torch.norm(torch.cat(mu_list)[:, :en] - torch.cat(self.current[2])[:, :en]) / en
            lam_diff = torch.norm(torch.cat(lam_list)[:, :en] - torch.cat(self.current[3])[:, :en]) / en
            diff = mu_diff ** 2 + lam_diff ** 2
        self.current = [nom_s, nom_u, mu_list, lam_list]
        self.last_diff = float(diff)
        return bool(diff < self.iter_threshold)


def run_batch(make_pan, inputs: dict, envs=None):
    """Loops fresh OraclePAN instances over envs of a batch-leading input dict; returns stacked outputs."""
    B = inputs["nom_s"].shape[0]
    envs = range(B) if envs is None else envs
    S, U, D, md, iters = [], [], [], [], []
    for b in envs:
        pan = make_pan()
        vel = None if inputs.get("velocities") is None else inputs["velocities"][b]
        pts = None if inputs.get("points") is None else inputs["points"][b]
        s, u, d = pan.forward(inputs["nom_s"][b], inputs["nom_u"][b], inputs["ref_s"][b], inputs["ref_us"][b], pts, vel)
        S.append(s); U.append(u); D.append(d if d is not None else np.zeros((1, u.shape[1]), np.float32))
        md.append(pan.min_distance); iters.append(pan.iters_run)
    return np.stack(S), np.stack(U), np.stack(D), np.array(md, np.float32), np.array(iters)
