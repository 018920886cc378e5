"""DUNE training (host/torch): the batched closed-form labels equal the certificate-carrying restatement of
oracle/dune_label.py; a short training run lowers the reference's four-term loss, writes the reference's artefacts, and its
checkpoint loads like a shipped model."""
import os

import numpy as np
import pytest
import torch

from helpers import CONFIGS, robot_spec
from oracle import dune_label as dl


@pytest.mark.parametrize("cname", ["C1", "C2", "C5"])
def test_batched_closed_form_labels_equal_the_oracle(cname):
    from neupan_b200.blocks.dune_train import closed_form_labels

    rb, _ = robot_spec(CONFIGS[cname])
    rng = np.random.default_rng(9)
    pts = rng.uniform(-25, 25, size=(3000, 2))
    pts[:50] *= 0.02  # some points inside / very close to the robot
    mu, dist = closed_form_labels(torch.from_numpy(rb.G), torch.from_numpy(rb.h), torch.from_numpy(pts))
    mu_o, dist_o = dl.labels(rb.G, rb.h, pts)
    assert np.abs(dist.numpy() - dist_o).max() < 1e-10 and np.abs(mu.numpy() - mu_o).max() < 1e-9
    for p, m, v in zip(pts[:400], mu.numpy()[:400], dist.numpy()[:400]):
        infeas, gap = dl.certificate(rb.G, rb.h, p, m, v)
        assert infeas < 1e-9 and gap < 1e-9 * max(1.0, v)


def test_short_training_run(tmp_path):
    from neupan_b200.blocks.dune import DUNE
    from neupan_b200.blocks.dune_train import DUNETrain

    torch.manual_seed(0)
    np.random.seed(0)
    rb, _ = robot_spec(CONFIGS["C1"])
    dune = DUNE(10, None, rb, 100, dict(direct_train=True))  # fresh weights (dune.py:146-152 with direct_train)
    name = dune.train_dune(dict(model_name="unit", checkpoint_dir=str(tmp_path), data_size=4000, data_range=[-10, -10, 10, 10], batch_size=128, epoch=120,
                                valid_freq=40, save_freq=60, lr=1e-3, lr_decay=0.5, decay_freq=110, save_loss=True))
    tr = dune.train_model
    assert name == os.path.join(str(tmp_path), "unit", "model_120.pth") and os.path.exists(name)
    assert tr.loss_list[-1] < 0.5 * tr.loss_list[0]                      # the four-term loss goes down
    assert len(tr.loss_list) == 121
    log = open(os.path.join(str(tmp_path), "unit", "results.txt")).read()
    assert "Epoch 120/120" in log and "current learning rate" in log and "Validate Fb Loss" in log
    for f in ("train_dict.pkl", "loss.pkl", "model_0.pth", "model_60.pth"):
        assert os.path.exists(os.path.join(str(tmp_path), "unit", f))
    sd = torch.load(name, map_location="cpu")
    assert set(sd) == set(dune.model.state_dict()) and sd["MLP.13.weight"].shape == (4, 32)
    again = DUNE(10, name, rb, 100, None)                                  # loads like a shipped checkpoint
    x = torch.randn(5, 2)
    assert torch.allclose(again.model(x), dune.model(x))
    second = dune.train_dune(dict(model_name="unit", checkpoint_dir=str(tmp_path), data_size=500, epoch=0, save_freq=1, valid_freq=1))
    assert os.path.dirname(second).endswith("unit_1")                      # an existing directory is not overwritten
