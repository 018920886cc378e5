"""GPU diagnostics of the differentiable path (developer tool): per-environment gradient of a random linear loss vs the oracle chain,
for K = 1 and K = 2.  python tools/grad_diag.py C5"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import test_gpu_grad as tg  # noqa: E402
from gpu_helpers import make_pan, to_cuda  # noqa: E402
from helpers import CONFIGS, make_inputs  # noqa: E402
from oracle import nrmp_grad as og  # noqa: E402

np.set_printoptions(precision=5, linewidth=200, suppress=False)
cname = sys.argv[1] if len(sys.argv) > 1 else "C5"
cfg = CONFIGS[cname]
B, N = 6, 80
inp = make_inputs(cfg, B=B, N=N, scene="obstacles")
rng = np.random.default_rng(7)
wS, wU, wD = rng.normal(size=(B, 3, cfg.T + 1)), rng.normal(size=(B, 2, cfg.T)), rng.normal(size=(B, 1, cfg.T))
wS[:, :, 0] = 0.0
c = lambda a: torch.from_numpy(a).float().cuda()
for K in (1, 2):
    for tol_env in ("", "1e-14"):
        if tol_env:
            os.environ["NB_NRMP_GAP_TOL"] = tol_env
        pan = make_pan(cfg, K=K, N=N, max_envs=B)
        t = to_cuda(inp)
        S, U, D = pan(t["nom_s"], t["nom_u"], t["ref_s"], t["ref_us"], t["points"], t["velocities"])
        ((S * c(wS)).sum() + (U * c(wU)).sum() + (D * c(wD)).sum()).backward()
        got = pan.last_grad_theta.cpu().numpy().astype(np.float64)
        print("K", K, "gap_tol", tol_env or "default", "status", pan.status.cpu().tolist(), "ipm", pan.ipm_iterations.cpu().tolist())
        for b in range(B):
            probs, ref_s, ref_us = tg._oracle_chain(cfg, inp, b, K, N)
            want = og.backward_chain(probs, ref_s, ref_us, wS[b], wU[b], wD[b, 0])
            Uo = np.stack([p for p in [probs[-1]]])  # noqa: F841
            print("  env", b, "rel err %.2e" % (np.abs(got[b] - want).max() / max(1.0, np.abs(want).max())), "got", got[b], "want", want)
        pan.close()
