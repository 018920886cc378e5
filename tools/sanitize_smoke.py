"""Small runs covering every kernel path, for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import numpy as np, torch
from gpu_helpers import make_pan, run_pan, to_cuda
from helpers import CONFIGS, make_inputs
SECTIONS = os.environ.get("NB_SMOKE_SECTIONS", "variants,screen,adjoint,train").split(",")
# N = 300: two slots, then a pass with one; N = 600: keys in shared memory (n > 512)
for cname, B, N in (("C4", 6, 70), ("C5", 3, 40), ("C2", 5, 33), ("C4", 3, 300), ("C3", 2, 600)) if "variants" in SECTIONS else ():
    cfg = CONFIGS[cname]
    inp = make_inputs(cfg, B=B, N=N, scene="obstacles")
    for dk in (3, 2, 1, 0):
        pan = make_pan(cfg, K=2, N=N, max_envs=B, dune_kernel=dk, nrmp_warm=1 if dk == 2 else 0)
        S, U, D, md = run_pan(pan, inp)
        assert np.isfinite(S).all() and (pan.status.cpu().numpy() == 0).all()
        pan.close()
# round 2: the screening pipeline (both screening kernels, step-0 reuse, paired refine, exact remainder), the stop criterion as its own
# kernel (B >= 256), sub-batches on internal streams (B >= 128) and the chunked upload (nb_pan_forward_h2d / _host)
if "screen" in SECTIONS:
    for cname, B, N, K in (("C4", 7, 70, 3), ("C4", 3, 300, 2), ("C3", 2, 600, 2), ("C2", 260, 40, 3), ("C5", 130, 36, 2)):
        cfg = CONFIGS[cname]
        inp = make_inputs(cfg, B=B, N=N, scene="obstacles")
        for mma, skip in ((1, 1), (0, 1), (1, 0)):
            pan = make_pan(cfg, K=K, N=N, max_envs=B, dune_kernel=4, dune_screen_mma=mma, dune_skip_t0=skip, iter_threshold=0.05 if B > 100 else 0.0)
            S, U, D, md = run_pan(pan, inp)
            assert np.isfinite(S).all() and (pan.status.cpu().numpy() == 0).all()
            if B > 100 and mma and skip:
                pin = {k: (None if v is None else torch.from_numpy(np.ascontiguousarray(v)).pin_memory()) for k, v in inp.items()}
                with torch.no_grad():
                    for dev_out in (True, False):
                        out = pan(pin["nom_s"], pin["nom_u"], pin["ref_s"], pin["ref_us"], pin["points"], pin["velocities"], device_out=dev_out)
                        torch.cuda.synchronize()
                        assert torch.isfinite(out[0]).all()
            pan.close()
# persistent NRMP warps (more environments than resident warps is not needed: the counter path is taken when grid > resident) + adjoint
for cname, B in (("C4", 5), ("C5", 3), ("C1", 2)) if "adjoint" in SECTIONS else ():
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
for cname in ("C1", "C5") if "train" in SECTIONS else ():
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
