"""Groundwork for SURVEY 8f row 4 (DUNE training on the GPU): closed-form labels of program (10) (oracle/dune_label.py) carry
a solver-free optimality certificate on the polygons of the shipped robots, and the shipped networks were trained on them."""
import numpy as np
import pytest
import torch

from helpers import CONFIGS, robot_spec, weights_path
from oracle import dune as od, dune_label as dl


@pytest.mark.parametrize("cname", ["C1", "C2", "C5"])
def test_labels_are_dual_feasible_with_zero_gap(cname):
    rb, _ = robot_spec(CONFIGS[cname])
    G, h = rb.G, rb.h
    rng = np.random.default_rng(3)
    pts = rng.uniform(-25, 25, size=(400, 2))  # the training range of example/dune_train/*.yaml
    mus, vals = dl.labels(G, h, pts)
    for p, mu, v in zip(pts, mus, vals):
        infeas, gap = dl.certificate(G, h, p, mu, v)
        assert infeas < 1e-12 and gap < 1e-10 * max(1.0, v)
    inside = (G @ pts.T - h.reshape(-1, 1) <= 0).all(0)
    assert (vals[inside] == 0).all() and (vals[~inside] > 0).all()
    assert ((mus > 0).sum(1) <= 2).all()  # one edge or one vertex supports the maximiser


@pytest.mark.parametrize("cname", ["C1", "C2"])
def test_shipped_networks_approximate_these_labels(cname):
    """ObsPointNet was trained by the reference on exactly these labels (dune_train.py:100-140): far from the robot its mu and
    the distance mu'(Gp - h) are close to them -- a sanity link between the label oracle and the checkpoints."""
    cfg = CONFIGS[cname]
    rb, _ = robot_spec(cfg)
    w = od.load_weights(weights_path(cfg.model))
    rng = np.random.default_rng(5)
    pts = rng.uniform(-20, 20, size=(600, 2))
    pts = pts[np.array([dl.primal_distance(rb.G, rb.h, p) for p in pts]) > 1.0]
    mu_net = od.obs_point_net(w, torch.from_numpy(pts).float()).numpy()                 # (n, E)
    d_net = (mu_net * (pts @ rb.G.T - rb.h.reshape(1, -1))).sum(1)
    mus, vals = dl.labels(rb.G, rb.h, pts)
    assert np.median(np.abs(d_net - vals) / vals) < 0.05
    assert np.median(np.abs(mu_net - mus).max(1)) < 0.05
