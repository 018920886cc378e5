"""Size-independent properties of the CUDA path at the FULL north-star size (C4: B = 4096 envs, T = 10, N = 500 points),
where the CPU oracle cannot follow every environment: selection invariants of the DUNE half (sortedness, membership,
consistency of lam / distance with mu, invariance under permutations and rigid motions of the scene) and feasibility /
optimality of the NRMP half (bounds, linearised dynamics, KKT certificate and oracle agreement on random subsets).
One PAN iteration (K = 1), so that the checks are about one well-defined map."""
import numpy as np
import pytest
import torch

from gpu_helpers import make_pan, to_cuda
from helpers import CONFIGS, make_inputs, rel_err, robot_spec
from oracle import ipm as oi, nrmp as onr

pytestmark = pytest.mark.gpu
CFG = CONFIGS["C4"]
B = CFG.B  # 4096


@pytest.fixture(scope="module")
def full():
    inp = make_inputs(CFG, B=B)
    t = to_cuda(inp)
    pan = make_pan(CFG, K=1, max_envs=B)
    S, U, D = pan(t["nom_s"], t["nom_u"], t["ref_s"], t["ref_us"], t["points"], t["velocities"])
    sel = pan.read_selection()
    return dict(inp=inp, t=t, pan=pan, S=S, U=U, D=D, sel=sel, status=pan.status.clone(), min_distance=pan.min_distance.clone())


def _rot(th):
    c, s = torch.cos(th), torch.sin(th)
    return torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)  # (..., 2, 2)


def test_selection_is_sorted_complete_and_consistent(full):
    sel, t = full["sel"], full["t"]
    rb, _ = robot_spec(CFG)
    G = torch.from_numpy(rb.G).float().cuda(); h = torch.from_numpy(rb.h).float().cuda().reshape(-1)
    d = sel["distance"]
    assert torch.isfinite(d).all() and (sel["count"] == CFG.M).all()
    assert (d[..., 1:] >= d[..., :-1]).all()                        # ascending (dune.py:100-104)
    assert torch.equal(full["min_distance"], d[:, 0, 0])             # dune.py:97-98
    assert (sel["mu"] >= 0).all()                                    # ReLU head
    R = _rot(t["nom_s"][:, 2, :])                                    # (B, T+1, 2, 2)
    lam = -torch.einsum("btij,ej,btme->btmi", R, G, sel["mu"])       # lam = -R G^T mu   (dune.py:89)
    assert (lam - sel["lam"]).abs().max() < 2e-5 * max(1.0, float(sel["lam"].abs().max()))
    p0 = torch.einsum("btji,btmj->btmi", R, sel["points"] - t["nom_s"][:, :2, :].transpose(1, 2)[:, :, None, :])  # R^T (p - s)
    dist = (sel["mu"] * (torch.einsum("ej,btmj->btme", G, p0) - h)).sum(-1)   # mu^T (G p0 - h)   (dune.py:119-122)
    assert (dist - d).abs().max() < 5e-5


def test_selected_points_are_flowed_input_points(full):
    """every selected point is bit-for-bit one of p + t (v dt) (pan.py:178-186)"""
    sel, t = full["sel"], full["t"]
    steps = torch.arange(CFG.T + 1, device="cuda", dtype=torch.float32)
    for b0 in range(0, B, 512):
        sl = slice(b0, b0 + 512)
        pts, vel = t["points"][sl], t["velocities"][sl]                       # (b, 2, N)
        flow = pts[:, None] + steps[None, :, None, None] * (vel * np.float32(CFG.dt))[:, None]   # (b, T+1, 2, N)
        diff = (sel["points"][sl].transpose(2, 3)[..., None] - flow[:, :, :, None, :]).abs().amax(2)  # (b, T+1, M, N)
        assert (diff.amin(-1) == 0).all()


def test_selection_is_invariant_under_point_permutation(full):
    t, sel = full["t"], full["sel"]
    g = torch.Generator(device="cuda").manual_seed(5)
    perm = torch.rand(B, CFG.N, device="cuda", generator=g).argsort(-1)
    idx = perm[:, None, :].expand(-1, 2, -1)
    pan = make_pan(CFG, K=1, max_envs=B)
    pan(t["nom_s"], t["nom_u"], t["ref_s"], t["ref_us"], t["points"].gather(2, idx).contiguous(), t["velocities"].gather(2, idx).contiguous())
    s2 = pan.read_selection()
    assert torch.equal(s2["distance"], sel["distance"])   # each point's distance does not depend on its position in the batch
    strict = (sel["distance"][..., 1:] > sel["distance"][..., :-1]).all(-1)   # no exact ties: the order is determined
    assert torch.equal(s2["points"][strict], sel["points"][strict]) and torch.equal(s2["mu"][strict], sel["mu"][strict])
    pan.close()


def test_selection_is_equivariant_under_rigid_motion(full):
    t, sel = full["t"], full["sel"]
    phi = torch.tensor(0.7, device="cuda"); sh = torch.tensor([3.0, -2.0], device="cuda")
    Rm = _rot(phi)
    mv = lambda p: torch.einsum("ij,bjn->bin", Rm, p) + sh[None, :, None]
    nom = t["nom_s"].clone()
    nom[:, :2] = mv(t["nom_s"][:, :2]); nom[:, 2] = t["nom_s"][:, 2] + phi
    ref = t["ref_s"].clone()
    ref[:, :2] = mv(t["ref_s"][:, :2]); ref[:, 2] = t["ref_s"][:, 2] + phi
    pan = make_pan(CFG, K=1, max_envs=B)
    pan(nom, t["nom_u"], ref, t["ref_us"], mv(t["points"]).contiguous(), torch.einsum("ij,bjn->bin", Rm, t["velocities"]).contiguous())
    s2 = pan.read_selection()
    assert (s2["distance"] - sel["distance"]).abs().max() < 1e-4     # the robot-frame coordinates are unchanged up to rounding
    assert (s2["distance"][..., 0] - sel["distance"][..., 0]).abs().mean() < 2e-6
    pan.close()


def test_nrmp_outputs_are_feasible_everywhere(full):
    inp, S, U, D = full["inp"], full["S"], full["U"], full["D"]
    rb, spec = robot_spec(CFG)
    assert (full["status"] == 0).all()
    assert torch.isfinite(S).all() and torch.isfinite(U).all() and torch.isfinite(D).all()
    assert torch.equal(S[:, :, 0], full["t"]["nom_s"][:, :, 0])                       # initial state constraint (robot.py:234)
    ms, ma = torch.from_numpy(spec.max_speed).float().cuda(), torch.from_numpy(spec.max_acce).float().cuda() * np.float32(CFG.dt)
    assert (U.abs() <= ms[None, :, None] + 1e-5).all()                                # robot.py:216-218
    assert ((U[:, :, 1:] - U[:, :, :-1]).abs() <= ma[None, :, None] + 1e-5).all()     # robot.py:219-222
    adj = CFG.adjust
    Dv = D.reshape(B, -1)
    assert (Dv >= adj["d_min"] - 1e-6).all() and (Dv <= adj["d_max"] + 1e-6).all()   # nrmp.py:347-352


def test_nrmp_dynamics_and_optimality_on_random_subsets(full):
    inp, sel = full["inp"], {k: v.cpu().numpy() for k, v in full["sel"].items()}
    S, U, D = full["S"].cpu().numpy(), full["U"].cpu().numpy(), full["D"].cpu().numpy().reshape(B, -1)
    rb, spec = robot_spec(CFG)
    h = rb.h.reshape(-1).astype(np.float32)
    rng = np.random.default_rng(11)
    for b in rng.choice(B, 192, replace=False):   # linearised dynamics (robot.py:224-236, 272-316)
        A, Bm, Cm = onr.linearise(spec.kinematics, inp["nom_s"][b], inp["nom_u"][b], CFG.dt, spec.L)
        for t in range(CFG.T):
            nxt = A[t].astype(np.float64) @ S[b, :, t] + Bm[t].astype(np.float64) @ U[b, :, t] + Cm[t].astype(np.float64).reshape(3)
            assert np.abs(nxt - S[b, :, t + 1]).max() < 2e-5 * max(1.0, np.abs(S[b]).max())
    for b in rng.choice(B, 24, replace=False):    # KKT certificate + the float64 oracle on the same program
        fa = sel["lam"][b, 1:]                                                       # (T, M, 2)   fa = lam^T   (nrmp.py:243)
        fb = (sel["lam"][b, 1:] * sel["points"][b, 1:]).sum(-1) + sel["mu"][b, 1:] @ h   # lam^T p + mu^T h   (nrmp.py:244-247)
        prob = onr.build_problem(spec, onr.Adjust(**CFG.adjust), inp["nom_s"][b], inp["nom_u"][b], inp["ref_s"][b], inp["ref_us"][b],
                                 fa.astype(np.float32), fb.astype(np.float32), CFG.M)
        res, viol = onr.kkt_certificate(prob, S[b].astype(np.float64), U[b].astype(np.float64), D[b].astype(np.float64), act_tol=1e-5)
        assert viol < 1e-5
        So, Uo, Do, _ = oi.solve_ipm(prob)
        assert rel_err(S[b], So) < 1e-4 and rel_err(U[b], Uo) < 1e-4 and rel_err(D[b], Do[0]) < 1e-4
