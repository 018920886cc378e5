"""NRMP layer facade (mirrors neupan/blocks/nrmp.py:33-392).

Keeps the reference's adjustable parameters (q_s, p_u, eta, d_max, d_min as leaf tensors with
``requires_grad=True``, ``adjust_parameters`` list, ``update_adjust_parameters_value``) and ``points``.  The
convex program the reference builds with cvxpy (nrmp.py:263-383) is solved by the CUDA NRMP kernel
(neupan_b200/csrc/nrmp_kernel.cuh); there is no cvxpy / cvxpylayers dependency.  When autograd is recording,
``PAN.forward`` runs the native solve in differentiable mode and ``loss.backward()`` reaches these leaves through
``nb_pan_backward`` (adjoint solves on the device) -- the role of CvxpyLayer's backward in the reference (LON).
"""
from __future__ import annotations

from typing import Union

import numpy as np
import torch


def _scalar(v, grad=True):
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().item() if v.numel() == 1 else v.detach().cpu().numpy()
    return torch.tensor(v, dtype=torch.float32, requires_grad=grad)


class NRMP(torch.nn.Module):
    def __init__(self, receding: int, step_time: float, robot, nrmp_max_num: int = 10, eta: float = 10.0, d_max: float = 1.0,
                 d_min: float = 0.1, q_s: Union[float, list, np.ndarray] = 1.0, p_u: float = 1.0, ro_obs: float = 400, bk: float = 0.1,
                 **kwargs) -> None:
        super().__init__()
        self.T, self.dt, self.robot = receding, step_time, robot
        self.G = torch.from_numpy(np.asarray(robot.G)).float()
        self.h = torch.from_numpy(np.asarray(robot.h)).float()
        self.max_num = nrmp_max_num
        self.no_obs = not nrmp_max_num > 0
        self.eta, self.d_max, self.d_min = _scalar(eta), _scalar(d_max), _scalar(d_min)
        if isinstance(q_s, (list, tuple, np.ndarray)):
            q = np.array(q_s, dtype=np.float32).flatten()
            if q.shape[0] != 3:
                raise ValueError(f"q_s must be a scalar or a 3-element list/array, got {q.shape[0]} elements")  # nrmp.py:86-87
            self.q_s = torch.from_numpy(q).reshape(3, 1).requires_grad_(True)
        else:
            self.q_s = _scalar(q_s)
        self.p_u = _scalar(p_u)
        self.ro_obs, self.bk = float(ro_obs), float(bk)
        self.solver = kwargs.get("solver", "ECOS")  # accepted for yaml compatibility; unused
        self.obstacle_points = None
        self.version = 0  # bumped on every update so PAN re-sends the values to the device
        self._refresh()

    def _refresh(self):
        self.adjust_parameters = [self.q_s, self.p_u] if self.no_obs else [self.q_s, self.p_u, self.eta, self.d_max, self.d_min]
        self.version += 1

    def q_vector(self) -> np.ndarray:
        q = self.q_s.detach().cpu().numpy().astype(np.float32).reshape(-1)
        return np.repeat(q, 3) if q.size == 1 else q

    def update_adjust_parameters_value(self, **kwargs):
        """nrmp.py:171-217 (with its unbound-``value`` slip for scalar q_s fixed: a scalar stays a
        scalar, a list given to a scalar-initialised layer uses its first element)."""
        if "q_s" in kwargs:
            q_new = kwargs["q_s"]
            if self.q_s.dim() == 0:
                if isinstance(q_new, (list, tuple, np.ndarray)):
                    print(f"q_s should be a scalar when initialized as scalar, got list/array with {len(q_new)} elements. Using the first element: {q_new[0]}")
                    q_new = q_new[0]
                self.q_s = _scalar(q_new)
            else:
                if not isinstance(q_new, (list, tuple, np.ndarray, torch.Tensor)):
                    raise ValueError(f"q_s must be a 3d list, np.ndarray, or torch.Tensor, got {type(q_new)}")
                q = np.array(q_new.detach().cpu() if isinstance(q_new, torch.Tensor) else q_new, dtype=np.float32).flatten()
                if q.shape[0] != 3:
                    raise ValueError(f"q_s must be a scalar or a 3-element list/array, got {q.shape[0]} elements")
                self.q_s = torch.from_numpy(q).reshape(3, 1).requires_grad_(True)
        for name in ("p_u", "eta", "d_max", "d_min"):
            if name in kwargs:
                setattr(self, name, _scalar(kwargs[name]))
        self._refresh()

    def generate_adjust_parameter_value(self):
        return self.adjust_parameters

    @property
    def points(self):
        return self.obstacle_points
