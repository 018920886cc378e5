"""DUNE training on the device: the label kernel against the certificate-carrying oracle, and the native epoch
kernel (forward + four-term loss + backward + Adam in hand-written CUDA) against the same loop in eager torch on the CPU -- the
restatement of the reference's DUNETrain.train_one_epoch (dune_train.py:281-362) -- on identical data, batches and rotations."""
import numpy as np
import pytest
import torch

from gpu_helpers import record
from helpers import CONFIGS
from neupan_b200.blocks.dune_train import DUNETrain
from neupan_b200.blocks.obs_point_net import ObsPointNet
from oracle import dune_label as ol

pytestmark = pytest.mark.gpu


def _robot(cname):
    rb = CONFIGS[cname].make_robot()
    return np.asarray(rb.G, np.float32), np.asarray(rb.h, np.float32).reshape(-1)


@pytest.mark.parametrize("cname", ["C1", "C2", "C5"])
def test_label_kernel_matches_oracle(cname):
    G, h = _robot(cname)
    torch.manual_seed(0); np.random.seed(1)
    tr = DUNETrain(ObsPointNet(2, G.shape[0]), G, h, "/tmp/unused", backend="native")
    pts, mu, dist = tr.generate_data_set(4000, [-8, -8, 8, 8])  # includes points inside the robot
    P = pts.cpu().numpy().astype(np.float64)
    mu_o, d_o = ol.labels(G.astype(np.float64), h.astype(np.float64), P)
    # the kernel labels the float64 sample, the oracle here sees its float32 rounding: compare at float32 resolution
    err_mu, err_d = np.abs(mu.cpu().numpy() - mu_o).max(), np.abs(dist.cpu().numpy() - d_o).max()
    record("dune_label_kernel", config=cname, max_err_mu=float(err_mu), max_err_dist=float(err_d), inside=int((d_o == 0).sum()))
    assert err_mu < 5e-5 and err_d < 5e-5
    assert (d_o == 0).sum() > 10  # the inside branch was exercised


@pytest.mark.parametrize("cname,n", [("C1", 1024), ("C5", 1000)])
def test_native_epochs_match_the_torch_loop(cname, n):
    G, h = _robot(cname)
    E = G.shape[0]
    torch.manual_seed(3); np.random.seed(4)
    m_nat, m_ref = ObsPointNet(2, E), ObsPointNet(2, E)
    m_ref.load_state_dict(m_nat.state_dict())
    nat, ref = DUNETrain(m_nat, G, h, "/tmp/unused", backend="native"), DUNETrain(m_ref, G, h, "/tmp/unused", backend="torch")
    for t in (nat, ref):
        t.optimizer.param_groups[0]["lr"] = 1e-3
    data = nat.generate_data_set(n, [-25, -25, 25, 25])
    cpu = tuple(t.cpu() for t in data)
    rng = np.random.default_rng(5)
    worst = 0.0
    for ep in range(4):
        th = rng.uniform(0, 2 * np.pi, (n + 255) // 256)
        ln = nat.train_one_epoch(data, 256, False, thetas=th)
        lr_ = ref.train_one_epoch(cpu, 256, False, thetas=th)
        rel = max(abs(a - b) / max(abs(b), 1e-12) for a, b in zip(ln, lr_))
        worst = max(worst, rel)
        assert rel < 2e-3, (ep, ln, lr_)
    vn, vr = nat.train_one_epoch(data, 256, True, thetas=th), ref.train_one_epoch(cpu, 256, True, thetas=th)
    assert max(abs(a - b) / max(abs(b), 1e-12) for a, b in zip(vn, vr)) < 2e-3
    nat.sync_model()
    dw = max(float((a - b).abs().max()) for a, b in zip(m_nat.state_dict().values(), m_ref.state_dict().values()))
    record("dune_train_native_vs_torch", config=cname, n=n, worst_rel_loss_diff=float(worst), max_weight_diff=dw, loss_first=float(sum(lr_)), )
    assert dw < 5e-4  # 16 Adam steps with lr 1e-3 move weights by ~1e-2: the two runs stay together to float32 noise


def test_start_runs_the_reference_schedule_and_saves_a_loadable_checkpoint(tmp_path):
    G, h = _robot("C1")
    torch.manual_seed(0); np.random.seed(0)
    model = ObsPointNet(2, G.shape[0])
    tr = DUNETrain(model, G, h, str(tmp_path / "ckpt"), backend="native")
    name = tr.start(data_size=4096, data_range=[-25, -25, 25, 25], batch_size=256, epoch=30, valid_freq=10, save_freq=30, lr=2e-3, decay_freq=20)
    assert name.endswith("model_30.pth")
    sd = torch.load(name, map_location="cpu")
    assert set(sd.keys()) == set(model.state_dict().keys())
    assert tr.loss_list[-1] < 0.9 * tr.loss_list[0]
    record("dune_train_start", loss_first=float(tr.loss_list[0]), loss_last=float(tr.loss_list[-1]))
