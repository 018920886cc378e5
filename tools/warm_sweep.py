"""NRMP warm-start experiments (developer tool): time a full C4 control step and report the interior point iteration histogram of the
last PAN iteration for several settings of the developer switches, in one process.
    python tools/warm_sweep.py [B]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gpu_helpers import make_pan, to_cuda  # noqa: E402
from helpers import CONFIGS, make_inputs  # noqa: E402

cfg = CONFIGS["C4"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else cfg.B
sets = [to_cuda(make_inputs(cfg, B=B, env_offset=s * B)) for s in range(2)]
SETTINGS = [
    dict(warm=0),
    dict(warm=1),
    dict(warm=1, NB_NRMP_RESTART_IT="8", NB_NRMP_RESTART_GAP="1e-4"),
    dict(warm=1, NB_NRMP_RESTART_IT="6", NB_NRMP_RESTART_GAP="1e-3"),
    dict(warm=1, NB_NRMP_RESTART_IT="6", NB_NRMP_RESTART_GAP="1e-4"),
    dict(warm=1, NB_NRMP_RESTART_IT="5", NB_NRMP_RESTART_GAP="1e-3"),
    dict(warm=1, NB_NRMP_RESTART_IT="4", NB_NRMP_RESTART_GAP="1e-2"),
    dict(warm=1, NB_NRMP_RESTART_IT="4", NB_NRMP_RESTART_GAP="1e-3"),
]
KEYS = ("NB_NRMP_RESTART_IT", "NB_NRMP_RESTART_GAP")
ref = None
for st in SETTINGS:
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in st.items():
        if k != "warm":
            os.environ[k] = v
    pan = make_pan(cfg, K=cfg.K, max_envs=B, nrmp_warm=st["warm"])
    def step(t):
        with torch.no_grad():
            return pan(t["nom_s"], t["nom_u"], t["ref_s"], t["ref_us"], t["points"], t["velocities"])
    for i in range(3):
        out = step(sets[i & 1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(6):
        out = step(sets[i & 1])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 6
    it = pan.ipm_iterations.cpu().numpy()
    stt = pan.status.cpu().numpy()
    S = out[0].cpu().numpy()
    if ref is None:
        ref = S
    dev = float(np.abs(S - ref).max())
    print(json.dumps(dict(setting=st, ms_per_step=round(ms, 3), ipm_mean=round(float(it.mean()), 3), ipm_max=int(it.max()), bad=int((stt != 0).sum()),
                          max_dev_vs_cold=dev, tail={int(k): int(v) for k, v in zip(*np.unique(it, return_counts=True)) if k >= 15})), flush=True)
    pan.close()
