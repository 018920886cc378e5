"""Small runs covering every kernel path, for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import numpy as np, torch
from gpu_helpers import make_pan, run_pan, to_cuda
from helpers import CONFIGS, make_inputs
# N = 300: two slots, then a pass with one; N = 600: keys in shared memory (n > 512)
for cname, B, N in (("C4", 6, 70), ("C5", 3, 40), ("C2", 5, 33), ("C4", 3, 300), ("C3", 2, 600)):
    cfg = CONFIGS[cname]
    inp = make_inputs(cfg, B=B, N=N, scene="obstacles")
    for dk in (3, 2, 1, 0):
        pan = make_pan(cfg, K=2, N=N, max_envs=B, dune_kernel=dk, nrmp_warm=1 if dk == 2 else 0)
        S, U, D, md = run_pan(pan, inp)
        assert np.isfinite(S).all() and (pan.status.cpu().numpy() == 0).all()
        pan.close()
# persistent NRMP warps (more environments than resident warps is not needed: the counter path is taken when grid > resident) + adjoint
for cname, B in (("C4", 5), ("C5", 3), ("C1", 2)):
    cfg = CONFIGS[cname]
    inp = make_inputs(cfg, B=B, N=50, scene="obstacles")
    pan = make_pan(cfg, K=2, N=50, max_envs=B)
    t = to_cuda(inp)
    S, U, D = pan(t["nom_s"], t["nom_u"], t["ref_s"], t["ref_us"], t["points"], t["velocities"])
    (S.sum() + U.sum() + D.sum()).backward()
    assert all(torch.isfinite(p.grad).all() for p in pan.nrmp_layer.adjust_parameters)
    pan.close()
# training kernels (ragged last batch)
from neupan_b200.blocks.dune_train import DUNETrain
from neupan_b200.blocks.obs_point_net import ObsPointNet
for cname in ("C1", "C5"):
    rb = CONFIGS[cname].make_robot()
    G, h = np.asarray(rb.G, np.float32), np.asarray(rb.h, np.float32).reshape(-1)
    tr = DUNETrain(ObsPointNet(2, G.shape[0]), G, h, "/tmp/unused", backend="native")
    data = tr.generate_data_set(600, [-25, -25, 25, 25])
    for _ in range(2):
        losses = tr.train_one_epoch(data, 256, False)
    tr.train_one_epoch(data, 256, True)
    assert np.isfinite(losses).all()
    tr.sync_model(); tr.close()
torch.cuda.synchronize()
print("sanitize smoke ok")
