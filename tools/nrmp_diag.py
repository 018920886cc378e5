"""GPU diagnostics of the NRMP solve inside PAN (developer tool): interior point iterations and status per PAN iteration,
cold vs warm start.  python tools/nrmp_diag.py C4 [B]  -> prints one JSON line per (warm, K)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from gpu_helpers import make_pan, run_pan  # noqa: E402
from helpers import CONFIGS, make_inputs  # noqa: E402

cname = sys.argv[1] if len(sys.argv) > 1 else "C4"
cfg = CONFIGS[cname]
B = int(sys.argv[2]) if len(sys.argv) > 2 else min(cfg.B, 2048)
scene = sys.argv[3] if len(sys.argv) > 3 else "annulus"
inp = make_inputs(cfg, B=B, scene=scene)
for warm in (0, 1):
    for K in (1, 2, 3, 5, cfg.K):
        pan = make_pan(cfg, K=K, max_envs=B, nrmp_warm=warm)
        run_pan(pan, inp)
        it = pan.ipm_iterations.cpu().numpy()
        st = pan.status.cpu().numpy()
        bad = np.flatnonzero(st != 0)
        print(json.dumps(dict(config=cname, scene=scene, warm=warm, K=K, ipm_mean=float(it.mean()), ipm_max=int(it.max()), ipm_p99=float(np.percentile(it, 99)),
                              hist={int(k): int(v) for k, v in zip(*np.unique(it, return_counts=True))}, bad_envs=[int(b) for b in bad[:20]],
                              bad_status=[int(s) for s in st[bad][:20]], bad_iters=[int(i) for i in it[bad][:20]])))
        pan.close()
