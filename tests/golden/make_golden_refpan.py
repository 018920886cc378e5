"""Golden outputs of the REFERENCE'S OWN ``PAN.forward`` (neupan/blocks/pan.py, dune.py, nrmp.py, robot.py executing unmodified;
only cvxpy / cvxpylayers are replaced by the numeric shim oracle/cvx_shim.py, whose solver is HiGHS + an exact active-set polish).
One PAN iteration (iter_num = 1) per environment, so that the comparison with the CUDA path is not blurred by the chaotic K-fold
composition.  Run here (needs /root/reference):  python tests/golden/make_golden_refpan.py"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import CONFIGS, GOLDEN, make_inputs, weights_path  # noqa: E402
from oracle import dune as od, refload  # noqa: E402

refload.load_reference()
from neupan.blocks.pan import PAN as RefPAN  # noqa: E402
from neupan.robot import robot as RefRobot  # noqa: E402

out = {}
for cname, nenv, N in (("C1", 4, 100), ("C2", 4, 120), ("C4", 4, 150), ("C5", 3, 150)):
    cfg = CONFIGS[cname]
    inp = make_inputs(cfg, B=nenv, N=N, scene="obstacles")
    w = od.load_weights(weights_path(cfg.model))
    with tempfile.TemporaryDirectory() as tmp:
        ck = os.path.join(tmp, "model.pth")
        torch.save(dict(w), ck)
        S, U, D, md = [], [], [], []
        for b in range(nenv):
            rb = RefRobot(cfg.T, cfg.dt, **cfg.robot_kwargs)
            rp = RefPAN(cfg.T, cfg.dt, rb, iter_num=1, dune_max_num=N, nrmp_max_num=cfg.M, dune_checkpoint=ck, iter_threshold=0.0, adjust_kwargs=dict(cfg.adjust))
            t = lambda a: None if a is None else torch.from_numpy(a[b])
            s, u, d = rp(t(inp["nom_s"]), t(inp["nom_u"]), t(inp["ref_s"]), t(inp["ref_us"]), t(inp["points"]), t(inp["velocities"]))
            S.append(s.detach().numpy()); U.append(u.detach().numpy()); D.append(d.detach().numpy()); md.append(float(rp.min_distance))
    out[f"{cname}_S"], out[f"{cname}_U"], out[f"{cname}_D"], out[f"{cname}_md"] = np.stack(S), np.stack(U), np.stack(D), np.array(md, np.float32)
    out[f"{cname}_N"] = np.int64(N)
    print(cname, "done")
# the reference's default iter_num = 2 (every example yaml), BASELINE config 1 (corridor / diff): whole forward, 6 environments
cfg = CONFIGS["C1"]
inp = make_inputs(cfg, B=6, N=100, scene="obstacles")
w = od.load_weights(weights_path(cfg.model))
with tempfile.TemporaryDirectory() as tmp:
    ck = os.path.join(tmp, "model.pth")
    torch.save(dict(w), ck)
    S, U, D, md = [], [], [], []
    for b in range(6):
        rb = RefRobot(cfg.T, cfg.dt, **cfg.robot_kwargs)
        rp = RefPAN(cfg.T, cfg.dt, rb, iter_num=2, dune_max_num=100, nrmp_max_num=cfg.M, dune_checkpoint=ck, iter_threshold=0.0, adjust_kwargs=dict(cfg.adjust))
        t = lambda a: None if a is None else torch.from_numpy(a[b])
        s, u, d = rp(t(inp["nom_s"]), t(inp["nom_u"]), t(inp["ref_s"]), t(inp["ref_us"]), t(inp["points"]), t(inp["velocities"]))
        S.append(s.detach().numpy()); U.append(u.detach().numpy()); D.append(d.detach().numpy()); md.append(float(rp.min_distance))
out["C1k2_S"], out["C1k2_U"], out["C1k2_D"], out["C1k2_md"] = np.stack(S), np.stack(U), np.stack(D), np.array(md, np.float32)
np.savez_compressed(os.path.join(GOLDEN, "ref_pan.npz"), **out)
