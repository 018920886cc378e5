"""Groundwork for SURVEY 8f row 3 (differentiable NRMP): the float64 sensitivity of the NRMP solution to the adjust
parameters (oracle/nrmp_grad.py, implicit differentiation of the optimality conditions) against central finite differences of
the interior point oracle."""
import numpy as np
import pytest
import torch

from helpers import CONFIGS, make_inputs, robot_spec, weights_path
from oracle import dune as od, nrmp as onr, nrmp_grad as og


def _problem(cname, b, N=80):
    cfg = CONFIGS[cname]
    inp = make_inputs(cfg, B=b + 1, N=N, scene="obstacles")
    rb, spec = robot_spec(cfg)
    w = od.load_weights(weights_path(cfg.model))
    G, h = torch.from_numpy(rb.G).float(), torch.from_numpy(rb.h).float()
    vel = None if inp["velocities"] is None else torch.from_numpy(inp["velocities"][b])
    p0, R, pl = od.point_flow(torch.from_numpy(inp["nom_s"][b]), torch.from_numpy(inp["points"][b]), vel, cfg.T, cfg.dt, 10 ** 9)
    mu, lam, sp, md, dist = od.dune_forward(w, G, h, p0, R, pl)
    fa, fb = od.nrmp_coefficients(h, mu, lam, sp, cfg.T, cfg.M)
    prob = onr.build_problem(spec, onr.Adjust(**cfg.adjust), inp["nom_s"][b], inp["nom_u"][b], inp["ref_s"][b], inp["ref_us"][b], fa.numpy(), fb.numpy(), cfg.M)
    ref_s, ref_us = inp["ref_s"][b].astype(np.float64), inp["ref_us"][b].astype(np.float64)
    return og.with_theta(prob, ref_s, ref_us, og.theta_of(prob)), ref_s, ref_us  # float64-consistent gamma_a / gamma_b


@pytest.mark.parametrize("cname,b", [("C1", 0), ("C4", 0), ("C4", 3), ("C2", 1)])
def test_implicit_sensitivity_matches_finite_differences(cname, b):
    prob, ref_s, ref_us = _problem(cname, b)
    sens = og.solution_sensitivity(prob, ref_s, ref_us)
    checked = 0
    for i, name in enumerate(og.THETA):
        (fS, fU, fD), same = og.finite_difference(prob, ref_s, ref_us, i)
        if not same:
            continue  # the active set changes within the finite-difference stencil: the solution map has a kink there
        scale = max(1e-3, np.abs(fS).max(), np.abs(fU).max(), np.abs(fD).max())
        assert np.abs(sens["dS"][i] - fS).max() < 2e-4 * scale, (name, np.abs(sens["dS"][i] - fS).max(), scale)
        assert np.abs(sens["dU"][i] - fU).max() < 2e-4 * scale, name
        assert np.abs(sens["dD"][i] - fD).max() < 2e-4 * scale, name
        checked += 1
    assert checked >= 4


def test_sensitivities_are_not_trivially_zero():
    prob, ref_s, ref_us = _problem("C4", 0)
    sens = og.solution_sensitivity(prob, ref_s, ref_us)
    assert np.abs(sens["dU"][3]).max() > 1e-3      # p_u moves the speed profile
    assert sens["hinge"].any() or np.abs(sens["dD"][4]).max() == 0.0   # eta acts on D only through active hinges / bounds


def test_adjoint_backward_equals_contracted_sensitivities():
    prob, ref_s, ref_us = _problem("C4", 3)
    sens = og.solution_sensitivity(prob, ref_s, ref_us)
    rng = np.random.default_rng(0)
    gS, gU, gD = rng.normal(size=(3, prob.T + 1)), rng.normal(size=(2, prob.T)), rng.normal(size=prob.T)
    gS[:, 0] = 0.0  # the initial state is a constant
    want = np.array([(sens["dS"][i] * gS).sum() + (sens["dU"][i] * gU).sum() + (sens["dD"][i] * gD).sum() for i in range(7)])
    got = og.backward(prob, ref_s, ref_us, gS, gU, gD)
    assert np.abs(got - want).max() < 1e-8 * max(1.0, np.abs(want).max())


def _chain_problems(cname, env, K=2, N=80, scene="obstacles"):
    """The K NRMP problems one OraclePAN.forward solves for an environment (float64-consistent gamma_a / gamma_b)."""
    from helpers import oracle_factory

    cfg = CONFIGS[cname]
    inp = make_inputs(cfg, B=1, N=N, scene=scene, env_offset=env)
    pan = oracle_factory(cfg, K=K, N=N)()
    probs, orig = [], pan._solve
    pan._solve = lambda prob: (probs.append(prob), orig(prob))[1]
    vel = None if inp["velocities"] is None else inp["velocities"][0]
    pan.forward(inp["nom_s"][0], inp["nom_u"][0], inp["ref_s"][0], inp["ref_us"][0], inp["points"][0], vel)
    ref_s, ref_us = inp["ref_s"][0].astype(np.float64), inp["ref_us"][0].astype(np.float64)
    return [og.with_theta(p, ref_s, ref_us, og.theta_of(p)) for p in probs], ref_s, ref_us, inp


def test_state_gradient_of_one_solve_matches_finite_differences():
    """dL/dpara_s (the path by which the reference's autograd reaches the previous PAN iteration)."""
    import dataclasses

    from oracle import ipm

    prob, ref_s, ref_us = _problem("C4", 0)
    rng = np.random.default_rng(1)
    gS, gU, gD = rng.normal(size=(3, prob.T + 1)), rng.normal(size=(2, prob.T)), rng.normal(size=prob.T)
    gS[:, 0] = 0.0
    _, g_para = og.backward_full(prob, ref_s, ref_us, gS, gU, gD)

    def loss(q):
        S, U, D, _ = ipm.solve_ipm(q)
        return float((gS * S).sum() + (gU * U).sum() + (gD * D.reshape(-1)).sum())

    checked = 0
    for (r, t) in [(0, 1), (1, 3), (2, 5), (0, prob.T), (1, prob.T)]:
        fd = []
        for h in (1e-4, 1e-5):
            lp, lm = [], []
            for sgn, acc in ((+1, lp), (-1, lm)):
                ns = prob.nom_s.copy(); ns[r, t] += sgn * h
                acc.append(loss(dataclasses.replace(prob, nom_s=ns)))
            fd.append((lp[0] - lm[0]) / (2 * h))
        if abs(fd[0] - fd[1]) > 1e-4 * max(1.0, abs(fd[0])):
            continue  # kink inside the stencil
        assert abs(g_para[r, t] - fd[1]) < 2e-4 * max(1.0, abs(fd[1])), (r, t, g_para[r, t], fd)
        checked += 1
    assert checked >= 3 and np.abs(g_para).max() > 1e-6


@pytest.mark.parametrize("cname,env", [("C4", 0), ("C1", 0), ("C5", 1)])
def test_chain_gradient_matches_finite_differences_of_the_frozen_chain(cname, env):
    """backward_chain = gradient of the K-solve chain in which only theta and nom_s (= previous S) move."""
    probs, ref_s, ref_us, _ = _chain_problems(cname, env, K=2)
    T = probs[0].T
    rng = np.random.default_rng(2)
    wS, wU, wD = rng.normal(size=(3, T + 1)), rng.normal(size=(2, T)), rng.normal(size=T)
    wS[:, 0] = 0.0
    got = og.backward_chain(probs, ref_s, ref_us, wS, wU, wD)
    th0 = og.theta_of(probs[0])
    checked = 0
    for i in range(7):
        fd = []
        for rel in (1e-4, 1e-5):
            h = rel * max(1.0, abs(th0[i]))
            tp, tm = th0.copy(), th0.copy()
            tp[i] += h; tm[i] -= h
            fd.append((og.chain_loss(probs, ref_s, ref_us, tp, wS, wU, wD) - og.chain_loss(probs, ref_s, ref_us, tm, wS, wU, wD)) / (2 * h))
        if abs(fd[0] - fd[1]) > 1e-3 * max(1.0, abs(fd[0])):
            continue
        assert abs(got[i] - fd[1]) < 1e-3 * max(1.0, abs(fd[1])), (og.THETA[i], got[i], fd)
        checked += 1
    assert checked >= 4
