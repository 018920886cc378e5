"""Robot description for the PAN hot path (mirrors neupan/robot/robot.py).

Keeps the reference's constructor keywords and attributes (robot.py:32-71, 318-375): polygon
geometry -> (G, h), speed / acceleration bounds, kinematics tag, wheelbase.  The cvxpy
variable / parameter / cost fragments of the reference (robot.py:73-236) have no counterpart
here: the convex program they define is solved by the CUDA NRMP kernel
(neupan_b200/csrc/nrmp_kernel.cuh), which also evaluates the kinematics linearisation
(robot.py:239-316) on the device.  ``linear_*_model`` are kept as host-side helpers with the
reference's signatures for callers that want A, B, C as tensors.
"""
from __future__ import annotations

from math import cos, inf, sin, tan
from typing import Optional, Union

import numpy as np
import torch

from .util import gen_inequal_from_vertex

KINEMATICS_ID = {"diff": 0, "acker": 1, "omni": 2}


class robot:
    def __init__(
        self,
        receding: int = 10,
        step_time: float = 0.1,
        kinematics: Optional[str] = None,
        vertices: Optional[Union[list, np.ndarray]] = None,
        max_speed: list = [inf, inf],
        max_acce: list = [inf, inf],
        wheelbase: Optional[float] = None,
        length: Optional[float] = None,
        width: Optional[float] = None,
        **kwargs,
    ):
        if kinematics is None:
            raise ValueError("kinematics is required")  # robot.py:46-47
        if kinematics not in KINEMATICS_ID:
            raise ValueError("kinematics currently only supports acker, diff or omni")

        self.shape = None
        self.vertices = self.cal_vertices(vertices, length, width, wheelbase)
        self.G, self.h = gen_inequal_from_vertex(self.vertices)

        self.T = receding
        self.dt = step_time
        self.L = wheelbase
        self.kinematics = kinematics
        self.max_speed = np.c_[max_speed].astype(float) if isinstance(max_speed, (list, tuple)) else np.asarray(max_speed, float).reshape(2, 1)
        self.max_acce = np.c_[max_acce].astype(float) if isinstance(max_acce, (list, tuple)) else np.asarray(max_acce, float).reshape(2, 1)

        if kinematics == "acker" and self.max_speed[1] >= 1.57:  # robot.py:63-66
            print(f"Warning: max steering angle of acker robot is {self.max_speed[1]} rad, which is larger than 1.57 rad, so it is limited to 1.57 rad")
            self.max_speed[1] = 1.57

        self.speed_bound = self.max_speed
        self.acce_bound = self.max_acce * self.dt
        self.name = kwargs.get("name", self.kinematics + "_robot" + "_default")

    # ---- geometry (robot.py:318-375) -------------------------------------------------
    def cal_vertices_from_length_width(self, length, width, wheelbase=None):
        wheelbase = 0 if wheelbase is None else wheelbase
        x0 = -(length - wheelbase) / 2
        y0 = -width / 2
        return np.array([[x0, x0 + length, x0 + length, x0], [y0, y0, y0 + width, y0 + width]], dtype=float)

    def cal_vertices(self, vertices=None, length=None, width=None, wheelbase=None):
        if vertices is not None:
            if isinstance(vertices, list):
                vertices_np = np.array(vertices, dtype=float).T
            elif isinstance(vertices, np.ndarray):
                vertices_np = vertices
            else:
                raise ValueError("vertices must be a list or numpy array")
        else:
            self.shape = "rectangle"
            vertices_np = self.cal_vertices_from_length_width(length, width, wheelbase)
            self.length, self.width, self.wheelbase = length, width, wheelbase
        assert vertices_np.shape[1] >= 3, "vertices must be a numpy array of shape (2, N), N >= 3"
        return vertices_np

    # ---- kinematics linearisation, host-side helpers (robot.py:239-316) ----------------
    def generate_state_parameter_value(self, nom_s, nom_u, qs_ref_s, pu_ref_us):
        out = [nom_s, qs_ref_s, pu_ref_us]
        As, Bs, Cs = [], [], []
        for t in range(self.T):
            st, ut = nom_s[:, t:t + 1], nom_u[:, t:t + 1]
            if self.kinematics == "acker":
                A, B, C = self.linear_ackermann_model(st, ut, self.dt, self.L)
            elif self.kinematics == "diff":
                A, B, C = self.linear_diff_model(st, ut, self.dt)
            else:
                A, B, C = self.linear_omni_model(ut, self.dt)
            As.append(A); Bs.append(B); Cs.append(C)
        return out + As + Bs + Cs

    @staticmethod
    def _abc(A, B, C, like):
        dev = like.device if isinstance(like, torch.Tensor) else "cpu"
        mk = lambda x: torch.tensor(x, dtype=torch.float32, device=dev)
        return mk(A), mk(B), mk(C)

    def linear_ackermann_model(self, nom_st, nom_ut, dt, L):
        phi = float(nom_st[2, 0]); v = float(nom_ut[0, 0]); psi = float(nom_ut[1, 0])
        k = v * dt / (L * cos(psi) ** 2)
        A = [[1, 0, -v * dt * sin(phi)], [0, 1, v * dt * cos(phi)], [0, 0, 1]]
        B = [[cos(phi) * dt, 0], [sin(phi) * dt, 0], [tan(psi) * dt / L, k]]
        C = [[phi * v * sin(phi) * dt], [-phi * v * cos(phi) * dt], [-psi * k]]
        return self._abc(A, B, C, nom_st)

    def linear_diff_model(self, nom_state, nom_u, dt):
        phi = float(nom_state[2, 0]); v = float(nom_u[0, 0])
        A = [[1, 0, -v * dt * sin(phi)], [0, 1, v * dt * cos(phi)], [0, 0, 1]]
        B = [[cos(phi) * dt, 0], [sin(phi) * dt, 0], [0, dt]]
        C = [[phi * v * sin(phi) * dt], [-phi * v * cos(phi) * dt], [0]]
        return self._abc(A, B, C, nom_state)

    def linear_omni_model(self, nom_u, dt):
        phi = float(nom_u[1, 0]); v = float(nom_u[0, 0])
        A = [[1, 0, 0], [0, 1, 0], [0, 0, 1]]
        B = [[cos(phi) * dt, -v * sin(phi) * dt], [sin(phi) * dt, v * cos(phi) * dt], [0, 0]]
        C = [[phi * v * sin(phi) * dt], [-phi * v * cos(phi) * dt], [0]]
        return self._abc(A, B, C, nom_u)
