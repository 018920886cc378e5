"""Generates the committed fixtures under tests/golden/ .  Run HERE (needs /root/reference):

    python tests/golden/make_golden.py

* weights_<robot>.npz ........ the 18 tensors of example/model/<robot>/model_5000.pth (the input
                               fixture parity is measured with)
* ref_dune_<robot>.npz ....... outputs of the reference's OWN code (PAN.generate_point_flow,
                               DUNE.forward, imported unmodified through oracle/refload.py) on
                               seeded inputs, incl. a decimated and a single-point case
* ref_misc.npz ............... reference gen_inequal_from_vertex, robot.linear_*_model,
                               PAN.stop_criteria, downsample_decimation outputs
* oracle_pan_<cfg>.npz ....... end-to-end PAN outputs of the CPU oracle (oracle/pan.py, IPM
                               solver) on the first envs of each BASELINE config -- oracle-made,
                               NOT reference-made (the reference's NRMP needs cvxpylayers+ECOS)
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle.refload import load_reference, REFERENCE_ROOT  # noqa: E402

load_reference()
from neupan.blocks import DUNE, PAN  # noqa: E402  (the reference's classes)
from neupan.robot import robot as RefRobot  # noqa: E402
from neupan.util import gen_inequal_from_vertex as ref_gen, downsample_decimation as ref_decim  # noqa: E402

from neupan_b200.synth import CONFIGS, make_inputs  # noqa: E402
from oracle import dune as od, nrmp as onr, pan as op  # noqa: E402

MODELS = {"diff": "diff_robot_default", "acker": "acker_robot_default", "polygon": "polygon_robot"}
ROBOT_CFG = {"diff": "C4", "acker": "C2", "polygon": "C5"}


def ckpt(model):
    return os.path.join(REFERENCE_ROOT, "example", "model", MODELS[model], "model_5000.pth")


def main():
    torch.manual_seed(0)
    # ---- weights ------------------------------------------------------------------
    for m in MODELS:
        sd = torch.load(ckpt(m), map_location="cpu")
        np.savez(os.path.join(HERE, f"weights_{m}.npz"), **{k: v.numpy() for k, v in sd.items()})

    # ---- reference DUNE half ------------------------------------------------------
    for m, cname in ROBOT_CFG.items():
        cfg = CONFIGS[cname]
        ref_robot = RefRobot(cfg.T, cfg.dt, **cfg.robot_kwargs)
        out = {}
        for case, (N, dmax, dyn) in {"a": (64, 100, True), "b": (300, 100, False), "c": (1, 100, True), "d": (7, 100, True)}.items():
            dune = DUNE(cfg.T, ckpt(m), ref_robot, dmax, {})
            fake_pan = types.SimpleNamespace(T=cfg.T, dt=cfg.dt, dune_max_num=dmax, printed=True, print_once=lambda *_: None)
            fake_pan.point_state_transform = types.MethodType(PAN.point_state_transform, fake_pan)
            inp = make_inputs(cfg, B=1, N=N, seed=900 + ord(case))
            nom_s = torch.from_numpy(inp["nom_s"][0]); pts = torch.from_numpy(inp["points"][0])
            vel = torch.from_numpy(make_inputs(CONFIGS["C4"], B=1, N=N, seed=77)["velocities"][0]) if dyn else None
            pf, Rl, pl = PAN.generate_point_flow(fake_pan, nom_s, pts, vel)
            mu_l, lam_l, sp_l = dune(pf, Rl, pl)
            out[f"{case}_nom_s"] = nom_s.numpy(); out[f"{case}_points"] = pts.numpy()
            out[f"{case}_vel"] = np.zeros((0,), np.float32) if vel is None else vel.numpy()
            out[f"{case}_dune_max_num"] = np.int64(dmax)
            out[f"{case}_p0"] = torch.stack(pf).numpy(); out[f"{case}_R"] = torch.stack(Rl).numpy()
            out[f"{case}_mu"] = torch.stack(mu_l).numpy(); out[f"{case}_lam"] = torch.stack(lam_l).numpy()
            out[f"{case}_sorted_points"] = torch.stack(sp_l).numpy()
            out[f"{case}_min_distance"] = np.float32(dune.min_distance)
            out[f"{case}_dune_points"] = dune.points.numpy()
        out["G"] = ref_robot.G; out["h"] = ref_robot.h
        np.savez(os.path.join(HERE, f"ref_dune_{m}.npz"), **out)

    # ---- misc reference functions -------------------------------------------------
    misc = {}
    polys = {"rect_diff": np.array([[-0.8, 0.8, 0.8, -0.8], [-1.0, -1.0, 1.0, 1.0]]),
             "polygon_cw": np.array([[-0.8, -1.0], [-1.8, 1.0], [1.8, 1.0], [0.8, -1.0]]).T,
             "pentagon": np.array([[0, 0], [1, 0], [2, 1], [1, 3], [-1, 2]], float).T}
    for k, v in polys.items():
        G, h = ref_gen(v)
        misc[f"poly_{k}_v"], misc[f"poly_{k}_G"], misc[f"poly_{k}_h"] = v, G, h
    rng = np.random.default_rng(5)
    for kin, kw in {"diff": dict(length=1.6, width=2.0), "acker": dict(length=4.6, width=1.6, wheelbase=3),
                    "omni": dict(vertices=[[-0.8, -1.0], [-1.8, 1.0], [1.8, 1.0], [0.8, -1.0]])}.items():
        rr = RefRobot(10, 0.1, kinematics=kin, **kw)
        ns = torch.from_numpy(rng.uniform(-3, 3, (3, 11)).astype(np.float32))
        nu = torch.from_numpy(rng.uniform(-1, 4, (2, 10)).astype(np.float32))
        vals = rr.generate_state_parameter_value(ns, nu, ns, nu[0])
        misc[f"lin_{kin}_nom_s"], misc[f"lin_{kin}_nom_u"] = ns.numpy(), nu.numpy()
        misc[f"lin_{kin}_A"] = torch.stack(vals[3:13]).numpy()
        misc[f"lin_{kin}_B"] = torch.stack(vals[13:23]).numpy()
        misc[f"lin_{kin}_C"] = torch.stack(vals[23:33]).numpy()
    misc["decim_idx_500_100"] = ref_decim(np.arange(500)[None, :], 100)[0]
    misc["decim_idx_1080_100"] = ref_decim(np.arange(1080)[None, :], 100)[0]
    # stop criterion (pan.py:215-243) on random sorted mu/lam lists
    fake = types.SimpleNamespace(current_nom_values=[None, None, None, None], nrmp_max_num=10, iter_threshold=0.1)
    seq = []
    for it in range(3):
        s = torch.from_numpy(rng.normal(size=(3, 11)).astype(np.float32)); u = torch.from_numpy(rng.normal(size=(2, 10)).astype(np.float32))
        mu = [torch.from_numpy((0.3 * rng.random((4, 25))).astype(np.float32)) for _ in range(11)]
        lam = [torch.from_numpy((0.3 * rng.normal(size=(2, 25))).astype(np.float32)) for _ in range(11)]
        prev = fake.current_nom_values
        flag = PAN.stop_criteria(fake, s, u, mu, lam)
        if prev[0] is not None:
            en = 10
            d = (torch.norm(torch.cat(mu)[:, :en] - torch.cat(prev[2])[:, :en]) / en) ** 2 + (torch.norm(torch.cat(lam)[:, :en] - torch.cat(prev[3])[:, :en]) / en) ** 2
        else:
            d = torch.tensor(np.nan)
        misc[f"stop_{it}_mu"] = torch.stack(mu).numpy(); misc[f"stop_{it}_lam"] = torch.stack(lam).numpy()
        misc[f"stop_{it}_flag"] = np.bool_(flag); misc[f"stop_{it}_diff"] = np.float32(d)
    np.savez(os.path.join(HERE, "ref_misc.npz"), **misc)

    # ---- oracle PAN traces (oracle-made) ----------------------------------------------
    # Per-iteration trace (S_k, U_k, D_k, min_distance_k) of K PAN iterations.  Consumers check
    # ONE iteration at a time (feed trace[k-1], compare trace[k]): the K-fold composition is
    # chaotic in cluttered scenes (a 1-ulp change of A_t grows to 1e-1 within 10 iterations, see
    # DESIGN.md "parity methodology"), the single iteration is not.
    for cname, nenv in {"C1": 2, "C2": 4, "C3": 3, "C4": 4, "C5": 2}.items():
        cfg = CONFIGS[cname]
        rb = cfg.make_robot()
        spec = onr.RobotSpec(rb.kinematics, rb.G, rb.h, rb.max_speed.reshape(-1), rb.max_acce.reshape(-1), cfg.dt, rb.L)
        w = od.load_weights(os.path.join(HERE, f"weights_{cfg.model}.npz"))
        for scene in ("annulus", "obstacles"):
            inp = make_inputs(cfg, B=nenv, scene=scene)
            K = min(cfg.K, 6)
            tr = {k: [] for k in ("S", "U", "D", "min_distance")}
            for b in range(nenv):
                pan = op.OraclePAN(spec, w, T=cfg.T, iter_num=K, dune_max_num=cfg.N, nrmp_max_num=cfg.M, iter_threshold=0.0,
                                   adjust=onr.Adjust(**cfg.adjust), solver="ipm")
                vel = None if inp["velocities"] is None else inp["velocities"][b]
                pan.forward(inp["nom_s"][b], inp["nom_u"][b], inp["ref_s"][b], inp["ref_us"][b], inp["points"][b], vel, keep_trace=True)
                for k in tr:
                    tr[k].append(np.stack([np.asarray(it[k], np.float32) for it in pan.trace]))
            np.savez(os.path.join(HERE, f"oracle_pan_{cname}_{scene}.npz"), **{k: np.stack(v) for k, v in tr.items()},
                     n_env=np.int64(nenv), K=np.int64(K))
            print(cname, scene, "done", tr["S"][0].shape)


if __name__ == "__main__":
    main()
