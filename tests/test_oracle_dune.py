"""Pins oracle/dune.py (the CPU restatement of the DUNE half) against the reference's own code:
golden vectors produced by that code (tests/golden/ref_dune_*.npz, made by make_golden.py) and,
when /root/reference is present, the live import."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, weights_path
from oracle import dune as od
from oracle.refload import reference_available


@pytest.mark.parametrize("model", ["diff", "acker", "polygon"])
@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_dune_matches_reference_golden(model, case):
    z = np.load(f"{GOLDEN}/ref_dune_{model}.npz")
    w = od.load_weights(weights_path(model))
    G = torch.from_numpy(z["G"]).float(); h = torch.from_numpy(z["h"]).float()
    nom_s = torch.from_numpy(z[f"{case}_nom_s"]); pts = torch.from_numpy(z[f"{case}_points"])
    vel = z[f"{case}_vel"]; vel = None if vel.size == 0 else torch.from_numpy(vel)
    T = nom_s.shape[1] - 1
    p0, R, pl = od.point_flow(nom_s, pts, vel, T, 0.1, int(z[f"{case}_dune_max_num"]))
    mu, lam, sp, md, _ = od.dune_forward(w, G, h, p0, R, pl, stable=False)
    assert np.array_equal(torch.stack(p0).numpy(), z[f"{case}_p0"])
    assert np.array_equal(torch.stack(R).numpy(), z[f"{case}_R"])
    assert np.array_equal(torch.stack(mu).numpy(), z[f"{case}_mu"])  # bit-for-bit
    assert np.array_equal(torch.stack(lam).numpy(), z[f"{case}_lam"])
    assert np.array_equal(torch.stack(sp).numpy(), z[f"{case}_sorted_points"])
    assert np.float32(md) == z[f"{case}_min_distance"]
    assert np.array_equal(pl[0].numpy(), z[f"{case}_dune_points"])
    # the deterministic (stable) order differs from the reference order only inside exact ties
    mu_s, lam_s, sp_s, _, dist = od.dune_forward(w, G, h, p0, R, pl, stable=True)
    for t in range(T + 1):
        d_ref = od.objective_distance(G, h, mu[t], None if False else (R[t].T @ (sp[t] - nom_s[0:2, t:t + 1])))
        d_stb = od.objective_distance(G, h, mu_s[t], (R[t].T @ (sp_s[t] - nom_s[0:2, t:t + 1])))
        assert torch.equal(d_ref, d_stb)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present on this machine")
def test_dune_matches_live_reference():
    import types

    from oracle.refload import REFERENCE_ROOT, load_reference

    load_reference()
    from neupan.blocks import DUNE, PAN
    from neupan.robot import robot as RefRobot

    rr = RefRobot(10, 0.1, kinematics="acker", length=4.6, width=1.6, wheelbase=3)
    ck = f"{REFERENCE_ROOT}/example/model/acker_robot_default/model_5000.pth"
    dune = DUNE(10, ck, rr, 80, {})
    fake = types.SimpleNamespace(T=10, dt=0.1, dune_max_num=80, printed=True, print_once=lambda *_: None)
    fake.point_state_transform = types.MethodType(PAN.point_state_transform, fake)
    g = torch.Generator().manual_seed(3)
    nom_s = torch.randn(3, 11, generator=g); pts = 6 * torch.randn(2, 200, generator=g); vel = torch.randn(2, 200, generator=g)
    pf, Rl, pl = PAN.generate_point_flow(fake, nom_s, pts, vel)
    mu_r, lam_r, sp_r = dune(pf, Rl, pl)
    w = od.load_weights(ck)
    p0, R, pl2 = od.point_flow(nom_s, pts, vel, 10, 0.1, 80)
    mu, lam, sp, md, _ = od.dune_forward(w, torch.from_numpy(rr.G).float(), torch.from_numpy(rr.h).float(), p0, R, pl2, stable=False)
    assert all(torch.equal(a, b) for a, b in zip(mu, mu_r))
    assert all(torch.equal(a, b) for a, b in zip(lam, lam_r))
    assert all(torch.equal(a, b) for a, b in zip(sp, sp_r))
    assert float(md) == float(dune.min_distance)
    # weights fixture == checkpoint
    wz = od.load_weights(weights_path("acker"))
    assert all(torch.equal(w[k], wz[k]) for k in w)
