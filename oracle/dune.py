"""CPU restatement of the DUNE half of the PAN hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows, op for op, the reference's torch code so that on CPU it is bit-identical to it:

* point flow + robot-frame transform .... neupan/blocks/pan.py:150-212, neupan/util/__init__.py:285-305
* ObsPointNet ........................... neupan/blocks/obs_point_net.py:25-49
* DUNE.forward / cal_objective_distance . neupan/blocks/dune.py:58-127

Single environment (the reference has no batch axis); ``oracle/pan.py`` loops over envs.
Pinned by tests/test_oracle_dune.py against the imported reference and tests/golden/*.npz.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# keys of the reference checkpoint (nn.Sequential indices, obs_point_net.py:31-46)
_LINEAR = (0, 3, 5, 8, 10, 13)
_LNORM = (1, 6, 11)


def load_weights(path: str) -> dict:
    """Reads either a reference ``.pth`` state_dict (dune.py:141-144) or a golden ``.npz``
    with the same 18 keys and returns {key: float32 torch tensor}."""
    if path.endswith(".npz"):
        z = np.load(path)
        return {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files if k.startswith("MLP.")}
    sd = torch.load(path, map_location="cpu")
    return {k: v.float() for k, v in sd.items()}


def obs_point_net(w: dict, x: torch.Tensor) -> torch.Tensor:
    """x: (rows, 2) float32 -> mu: (rows, E).  obs_point_net.py:31-49 (LayerNorm eps = 1e-5)."""
    h = F.linear(x, w["MLP.0.weight"], w["MLP.0.bias"])
    h = torch.tanh(F.layer_norm(h, (32,), w["MLP.1.weight"], w["MLP.1.bias"], 1e-5))
    h = torch.relu(F.linear(h, w["MLP.3.weight"], w["MLP.3.bias"]))
    h = F.linear(h, w["MLP.5.weight"], w["MLP.5.bias"])
    h = torch.tanh(F.layer_norm(h, (32,), w["MLP.6.weight"], w["MLP.6.bias"], 1e-5))
    h = torch.relu(F.linear(h, w["MLP.8.weight"], w["MLP.8.bias"]))
    h = F.linear(h, w["MLP.10.weight"], w["MLP.10.bias"])
    h = torch.tanh(F.layer_norm(h, (32,), w["MLP.11.weight"], w["MLP.11.bias"], 1e-5))
    return torch.relu(F.linear(h, w["MLP.13.weight"], w["MLP.13.bias"]))


def decimate(mat: torch.Tensor, m: int) -> torch.Tensor:
    """util/__init__.py:285-305: column pick np.linspace(0, n-1, m).astype(int)."""
    n = mat.shape[1]
    if m >= n:
        return mat
    idx = np.linspace(0, n - 1, m).astype(int)
    return mat[:, idx]


def point_flow(nom_s: torch.Tensor, obs_points: torch.Tensor, point_velocities, T: int, dt: float, dune_max_num: int):
    """pan.py:150-212.  Returns (p0_list, R_list, p_list), each of length T+1."""
    if point_velocities is None:
        point_velocities = torch.zeros_like(obs_points)
    if obs_points.shape[1] > dune_max_num:
        obs_points = decimate(obs_points, dune_max_num)
        point_velocities = decimate(point_velocities, dune_max_num)
    p_list, p0_list, R_list = [], [], []
    for i in range(T + 1):
        p_t = obs_points + i * (point_velocities * dt)  # pan.py:182
        state = nom_s[:, i].reshape((3, 1))
        trans = state[0:2]
        theta = state[2, 0]
        # pan.py:208: R built from 0-d tensors through torch.tensor -> float32
        R = torch.tensor([[torch.cos(theta), -torch.sin(theta)], [torch.sin(theta), torch.cos(theta)]])
        p0 = R.T @ (p_t - trans)  # pan.py:210
        p_list.append(p_t)
        p0_list.append(p0)
        R_list.append(R)
    return p0_list, R_list, p_list


def objective_distance(G: torch.Tensor, h: torch.Tensor, mu: torch.Tensor, p0: torch.Tensor) -> torch.Tensor:
    """dune.py:109-127: dist_n = mu_n^T (G p0_n - h)."""
    temp = (G @ p0 - h).T.unsqueeze(2)
    muT = mu.T.unsqueeze(1)
    distance = torch.squeeze(torch.bmm(muT, temp))
    if distance.ndim == 0:
        distance = distance.unsqueeze(0)
    return distance


def dune_forward(w: dict, G: torch.Tensor, h: torch.Tensor, p0_list, R_list, p_list, stable: bool = True):
    """dune.py:58-106.  Returns (mu_list, lam_list, sort_point_list, min_distance, dist_list).

    ``stable=True`` sorts with a stable argsort (ties -> lower index first); the reference
    calls torch.argsort without ``stable`` (dune.py:100), whose order among exact ties is
    unspecified -- any tie order is a valid reference output, we fix the deterministic one.
    ``dist_list`` (unsorted distances) is extra, for tolerance-aware tests.
    """
    T1 = len(p0_list)
    total_points = torch.hstack(p0_list)
    with torch.no_grad():
        total_mu = obs_point_net(w, total_points.T).T
    mu_list, lam_list, sort_point_list, dist_list = [], [], [], []
    min_distance = None
    for index in range(T1):
        n = p0_list[index].shape[1]
        mu = total_mu[:, index * n:(index + 1) * n]
        R = R_list[index]
        p0 = p0_list[index]
        lam = -R @ G.T @ mu  # dune.py:89
        distance = objective_distance(G, h, mu, p0)
        if index == 0:
            min_distance = torch.min(distance)
        sort_indices = torch.argsort(distance, stable=stable)
        mu_list.append(mu[:, sort_indices])
        lam_list.append(lam[:, sort_indices])
        sort_point_list.append(p_list[index][:, sort_indices])
        dist_list.append(distance)
    return mu_list, lam_list, sort_point_list, min_distance, dist_list


def nrmp_coefficients(h: torch.Tensor, mu_list, lam_list, point_list, T: int, max_num: int):
    """nrmp.py:220-261: fa_t = lam^T[:M], fb_t = (lam^T p + mu^T h)[:M] from list entry t+1,
    rows pn..M padded with row 0.  Returns (fa (T,M,2), fb (T,M)) float32."""
    fa_out = torch.zeros((T, max_num, 2))
    fb_out = torch.zeros((T, max_num, 1))
    if not mu_list:
        return fa_out, fb_out[:, :, 0]
    for t in range(T):
        mu, lam, point = mu_list[t + 1], lam_list[t + 1], point_list[t + 1]
        fa = lam.T
        temp = torch.bmm(lam.T.unsqueeze(1), point.T.unsqueeze(2)).squeeze(1)
        fb = temp + mu.T @ h
        pn = min(mu.shape[1], max_num)
        fa_out[t, :pn, :] = fa[:pn, :]
        fb_out[t, :pn, :] = fb[:pn, :]
        fa_out[t, pn:, :] = fa[0, :]
        fb_out[t, pn:, :] = fb[0, :]
    return fa_out, fb_out[:, :, 0]
