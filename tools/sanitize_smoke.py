"""Small PAN run for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import numpy as np, torch
from gpu_helpers import make_pan, run_pan
from helpers import CONFIGS, make_inputs
# N = 300: two slots, then a pass with one; N = 600: keys in shared memory (n > 512)
for cname, B, N in (("C4", 6, 70), ("C5", 3, 40), ("C2", 5, 33), ("C4", 3, 300), ("C3", 2, 600)):
    cfg = CONFIGS[cname]
    inp = make_inputs(cfg, B=B, N=N, scene="obstacles")
    for dk in (2, 1, 0):
        pan = make_pan(cfg, K=2, N=N, max_envs=B, dune_kernel=dk)
        S, U, D, md = run_pan(pan, inp)
        assert np.isfinite(S).all() and (pan.status.cpu().numpy() == 0).all()
        pan.close()
torch.cuda.synchronize()
print("sanitize smoke ok")
