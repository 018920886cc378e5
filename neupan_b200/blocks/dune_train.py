"""DUNE training (SURVEY 8f "next" row 4): same data set, losses, schedule, checkpoints and log as the reference's
``DUNETrain`` (neupan/blocks/dune_train.py:60-384), without cvxpy.

The reference labels every sampled point by solving the cone program (10)  max mu'(Gp - h)  s.t. ||G'mu|| <= 1, mu >= 0  with
cvxpy/ECOS, one solve per point (``prob_solve``, :134-140) -- 100,000 solves for the default data set.  That program is the
dual of the distance from p to the robot polygon, so the labels are available in closed form; ``closed_form_labels`` evaluates
them for the whole data set at once (verified against oracle/dune_label.py's certificate-carrying restatement by
tests/test_dune_train.py).  The training loop itself is the reference's: Adam(lr, weight_decay 1e-4), per-batch loss
MSE(mu) + MSE(distance) + MSE(fa) + MSE(fb)  with a random rotation R per batch (:281-362), 80/20 split, lr decay,
``model_<epoch>.pth`` checkpoints (same ``state_dict`` keys as the shipped models) and ``results.txt``.

Two backends:
* ``backend="native"`` (default when a CUDA device is present): labels by ``nb_dune_labels`` and every epoch by
  ``nb_dune_train_epoch`` -- hand-written CUDA (csrc/dune_train_kernel.cuh): forward, four-term loss, backward and Adam of one
  epoch in a single persistent CTA with the parameters resident in shared memory;
* ``backend="torch"``: the same loop in eager torch on the model's device (what round 1 shipped) -- kept because the reference
  trains on CPU-only machines too, and as the checker of the native kernels (tests/test_gpu_train.py).
"""
from __future__ import annotations

import math
import os
import pickle

import numpy as np
import torch


def polygon_vertices(G: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """vertex e = intersection of the rows e and e+1 of {x : G x <= h} (rows in cyclic order, gen_inequal_from_vertex)"""
    E = G.shape[0]
    A = torch.stack([torch.stack([G[e], G[(e + 1) % E]]) for e in range(E)])          # (E,2,2)
    b = torch.stack([torch.stack([h[e], h[(e + 1) % E]]) for e in range(E)])          # (E,2)
    return torch.linalg.solve(A, b.unsqueeze(-1)).squeeze(-1)                          # (E,2)


def closed_form_labels(G: torch.Tensor, h: torch.Tensor, points: torch.Tensor):
    """(mu* (n,E), distance (n,)) of program (10) for n points (n,2); float64 inside."""
    G, h, P = G.double(), h.double().reshape(-1), points.double()
    E, n = G.shape[0], P.shape[0]
    V = polygon_vertices(G, h)
    a, b = V.roll(1, 0), V                                  # row e runs from vertex e-1 to vertex e
    d = b - a                                               # (E,2)
    t = (((P[:, None, :] - a[None]) * d[None]).sum(-1) / (d * d).sum(-1)[None]).clamp(0.0, 1.0)   # (n,E)
    x = a[None] + t[..., None] * d[None]                    # closest point of every edge
    dist_e = (P[:, None, :] - x).norm(dim=-1)
    dist, e = dist_e.min(dim=1)                             # closest edge
    te = t.gather(1, e[:, None]).squeeze(1)
    r = P @ G.T - h[None]                                   # (n,E)
    inside = (r <= 0).all(dim=1)
    mu = torch.zeros(n, E, dtype=torch.float64, device=P.device)
    rows = torch.arange(n, device=P.device)
    interior = (te > 0) & (te < 1)
    mu[rows, e] = torch.where(interior, 1.0 / G[e].norm(dim=1), torch.zeros_like(dist))
    vert = ~interior & ~inside
    if vert.any():
        ev = e[vert]
        f = torch.where(te[vert] == 0, (ev - 1) % E, (ev + 1) % E)
        xv = x[vert, ev]
        nvec = (P[vert] - xv) / dist[vert, None].clamp_min(1e-300)
        M = torch.stack([G[ev], G[f]], dim=-1)              # columns G_e', G_f'
        sol = torch.linalg.solve(M, nvec.unsqueeze(-1)).squeeze(-1).clamp_min(0.0)
        mu[rows[vert], ev] = sol[:, 0]
        mu[rows[vert], f] = sol[:, 1]
    mu[inside] = 0.0
    value = (mu * r).sum(1)
    value[inside] = 0.0
    return mu, value


class DUNETrain:
    def __init__(self, model, robot_G, robot_h, checkpoint_path, backend=None, device=None) -> None:
        self.model = model
        self.device = next(model.parameters()).device
        if backend is None:
            backend = "native" if torch.cuda.is_available() else "torch"
        if backend not in ("native", "torch"):
            raise ValueError("backend must be 'native' or 'torch'")
        self.backend = backend
        self._native = None
        self.cuda_device = torch.device("cuda", torch.cuda.current_device()) if (backend == "native" and device is None) else (torch.device(device) if device is not None else None)
        self.G = torch.as_tensor(robot_G, dtype=torch.float32, device=self.device)
        self.h = torch.as_tensor(robot_h, dtype=torch.float32, device=self.device).reshape(-1, 1)
        self.checkpoint_path = checkpoint_path
        self.loss_fn = torch.nn.MSELoss()
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=1e-4, weight_decay=1e-4)  # dune_train.py:72
        self.loss_of_epoch, self.loss_list = 0, []

    # ---- native backend ------------------------------------------------------------------------------------------
    def _handle(self):
        """nb_dune_train handle holding the parameters + Adam moments on the device (created from the model's current weights)."""
        if self._native is None:
            import ctypes as C

            from .. import _lib

            lib = _lib.load()
            w = _lib.pack_weights(self.model.state_dict())
            G = np.ascontiguousarray(self.G.detach().cpu().numpy(), dtype=np.float32)
            h = np.ascontiguousarray(self.h.detach().cpu().numpy(), dtype=np.float32).reshape(-1)
            hd = C.c_void_p()
            _lib.check(lib.nb_dune_train_create(int(G.shape[0]), G.ctypes.data, h.ctypes.data, w.ctypes.data, int(w.size), int(self.cuda_device.index or 0), C.byref(hd)))
            self._native, self._nw = hd, int(w.size)
        return self._native

    def sync_model(self):
        """Copies the device parameters back into ``self.model`` (checkpoints, validation in torch, PAN)."""
        if self._native is None:
            return
        from .. import _lib

        flat = np.empty(self._nw, np.float32)
        _lib.check(_lib.load().nb_dune_train_get_weights(self._native, flat.ctypes.data))
        self.model.load_state_dict(_lib.unpack_weights(flat, self.model.state_dict()))

    def close(self):
        if self._native is not None:
            from .. import _lib

            _lib.load().nb_dune_train_destroy(self._native)
            self._native = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data (dune_train.py:100-140) ---------------------------------------------------------------------------
    def generate_data_set(self, data_size=10000, data_range=(-50, -50, 50, 50)):
        rand_p = np.random.uniform(low=data_range[:2], high=data_range[2:], size=(data_size, 2))
        if self.backend == "native":
            import ctypes as C

            from .. import _lib

            dev = self.cuda_device
            pts64 = torch.from_numpy(rand_p).to(dev)
            E = int(self.G.shape[0])
            pts, mu, dist = torch.empty((data_size, 2), device=dev), torch.empty((data_size, E), device=dev), torch.empty(data_size, device=dev)
            G = np.ascontiguousarray(self.G.detach().cpu().numpy(), dtype=np.float32)
            h = np.ascontiguousarray(self.h.detach().cpu().numpy(), dtype=np.float32).reshape(-1)
            with torch.cuda.device(dev):
                stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                _lib.check(_lib.load().nb_dune_labels(E, G.ctypes.data, h.ctypes.data, data_size, C.c_void_p(pts64.data_ptr()), C.c_void_p(pts.data_ptr()),
                                                      C.c_void_p(mu.data_ptr()), C.c_void_p(dist.data_ptr()), stream))
            return pts, mu, dist
        pts = torch.from_numpy(rand_p).to(self.device)
        mu, dist = closed_form_labels(self.G, self.h, pts)
        return pts.float(), mu.float(), dist.float()

    # ---- losses (dune_train.py:281-375) -------------------------------------------------------------------------
    def _losses(self, pts, label_mu, label_dist, theta):
        out_mu = self.model(pts)                                        # (b,E)
        temp = pts @ self.G.T - self.h.reshape(1, -1)                   # G p - h
        dist = (out_mu * temp).sum(1)
        R = torch.tensor([[math.cos(theta), -math.sin(theta)], [math.sin(theta), math.cos(theta)]], dtype=torch.float32, device=self.device)
        lam = lambda m: -(m @ self.G) @ R.T                             # (-R G' mu)' rows
        fa, fa_l = lam(out_mu), lam(label_mu)
        hb = self.h.reshape(1, -1)
        fb, fb_l = (fa * pts).sum(1) + (out_mu * hb).sum(1), (fa_l * pts).sum(1) + (label_mu * hb).sum(1)
        return self.loss_fn(out_mu, label_mu), self.loss_fn(dist, label_dist), self.loss_fn(fa, fa_l), self.loss_fn(fb, fb_l)

    def train_one_epoch(self, data, batch_size, validate=False, thetas=None):
        """One pass over `data`; `thetas` (one rotation angle per batch) defaults to fresh np.random.uniform(0, 2 pi) draws."""
        pts, mus, dists = data
        nbatch = (pts.shape[0] + batch_size - 1) // batch_size
        if thetas is None:
            thetas = np.random.uniform(0, 2 * np.pi, nbatch)
        if self.backend == "native":
            import ctypes as C

            from .. import _lib

            dev = self.cuda_device
            prep = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
            pts, mus, dists = prep(pts), prep(mus), prep(dists)
            th = np.ascontiguousarray(thetas, dtype=np.float32)
            losses = (C.c_double * 4)()
            with torch.cuda.device(dev):
                stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                _lib.check(_lib.load().nb_dune_train_epoch(self._handle(), C.c_void_p(pts.data_ptr()), C.c_void_p(mus.data_ptr()), C.c_void_p(dists.data_ptr()),
                                                           int(pts.shape[0]), int(batch_size), th.ctypes.data, float(self.optimizer.param_groups[0]["lr"]),
                                                           int(bool(validate)), C.byref(losses), stream))
            return tuple(float(v) for v in losses)
        sums, nb = [0.0, 0.0, 0.0, 0.0], 0
        for i in range(0, pts.shape[0], batch_size):
            self.optimizer.zero_grad()
            with torch.set_grad_enabled(not validate):
                parts = self._losses(pts[i:i + batch_size], mus[i:i + batch_size], dists[i:i + batch_size], float(thetas[i // batch_size]))
                if not validate:
                    sum(parts).backward()
                    self.optimizer.step()
            sums = [s + p.item() for s, p in zip(sums, parts)]
            nb += 1
        return tuple(s / max(nb, 1) for s in sums)

    # ---- schedule, checkpoints, log (dune_train.py:142-279) -------------------------------------------------------
    def start(self, data_size: int = 100000, data_range=(-25, -25, 25, 25), batch_size: int = 256, epoch: int = 5000, valid_freq: int = 100,
              save_freq: int = 500, lr: float = 5e-5, lr_decay: float = 0.5, decay_freq: int = 1500, save_loss: bool = False, **kwargs):
        os.makedirs(self.checkpoint_path, exist_ok=True)
        head = (f"data_size: {data_size}, data_range: {list(data_range)}, batch_size: {batch_size}, epoch: {epoch}, valid_freq: {valid_freq}, "
                f"save_freq: {save_freq}, lr: {lr}, lr_decay: {lr_decay}, decay_freq: {decay_freq}, robot_G: {self.G}, robot_h: {self.h}")
        with open(os.path.join(self.checkpoint_path, "train_dict.pkl"), "wb") as f:
            pickle.dump(dict(data_size=data_size, data_range=list(data_range), batch_size=batch_size, epoch=epoch, valid_freq=valid_freq, save_freq=save_freq,
                             lr=lr, lr_decay=lr_decay, decay_freq=decay_freq, robot_G=self.G.cpu(), robot_h=self.h.cpu()), f)
        with open(os.path.join(self.checkpoint_path, "results.txt"), "a") as f:
            print(head + "\n", file=f)
        self.optimizer.param_groups[0]["lr"] = float(lr)
        pts, mus, dists = self.generate_data_set(data_size, data_range)
        perm = torch.randperm(data_size).to(pts.device)               # random_split 80 / 20
        n_train = int(data_size * 0.8)
        tr, va = perm[:n_train], perm[n_train:n_train + int(data_size * 0.2)]
        train, valid = (pts[tr], mus[tr], dists[tr]), (pts[va], mus[va], dists[va])
        full_model_name = None
        fmt = lambda v: "{:.2e}".format(v)
        for i in range(epoch + 1):
            self.model.train(True)
            losses = self.train_one_epoch(train, batch_size, False)
            if i % valid_freq == 0:
                self.model.eval()
                vl = self.train_one_epoch(valid, batch_size, True)
                with open(os.path.join(self.checkpoint_path, "results.txt"), "a") as f:
                    print("Epoch {}/{} learning rate {} \n---------------------------------\nLosses:\n"
                          "  Mu Loss:          {} | Validate Mu Loss:          {}\n  Distance Loss:    {} | Validate Distance Loss:    {}\n"
                          "  Fa Loss:          {} | Validate Fa Loss:          {}\n  Fb Loss:          {} | Validate Fb Loss:          {}\n".format(
                              i, epoch, self.optimizer.param_groups[0]["lr"], fmt(losses[0]).ljust(10), fmt(vl[0]).rjust(10), fmt(losses[1]).ljust(10),
                              fmt(vl[1]).rjust(10), fmt(losses[2]).ljust(10), fmt(vl[2]).rjust(10), fmt(losses[3]).ljust(10), fmt(vl[3]).rjust(10)), file=f)
            if i % save_freq == 0:
                self.sync_model()
                full_model_name = os.path.join(self.checkpoint_path, f"model_{i}.pth")
                torch.save({k: v.detach().cpu() for k, v in self.model.state_dict().items()}, full_model_name)
            if (i + 1) % decay_freq == 0:
                self.optimizer.param_groups[0]["lr"] *= lr_decay
                with open(os.path.join(self.checkpoint_path, "results.txt"), "a") as f:
                    print("current learning rate:", self.optimizer.param_groups[0]["lr"], file=f)
            self.loss_of_epoch = sum(losses)
            self.loss_list.append(self.loss_of_epoch)
            if save_loss:
                with open(os.path.join(self.checkpoint_path, "loss.pkl"), "wb") as f:
                    pickle.dump(self.loss_list, f)
        self.sync_model()
        self.model.eval()
        return full_model_name
