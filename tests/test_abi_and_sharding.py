"""No-GPU checks of the boundary: the shared library loads and exports every symbol the header
declares, the ctypes mirror of the config struct matches, the package refuses to compute without
a CUDA device, and the multi-process sharding / gather logic works (gloo, world size 2)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from helpers import CONFIGS, weights_path
from neupan_b200 import _lib, build as nb_build
from neupan_b200.parallel import gather_results, pack_results, shard_range, unpack_results

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    nb_build.build()  # no-op when neupan_b200/lib/libneupan_b200.so is up to date
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "neupan_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(nb_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 15
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.nb_version() == 1
    assert [lib.nb_weight_count(e) for e in (3, 4, 8)] == [4611, 4644, 4776]


def test_config_struct_layout_matches_header():
    # C layout: 8 int32, float, (pad to 8), 2+2*2+2 doubles, 3+4 floats, tail pad to 8
    assert C.sizeof(_lib.PanConfig) == 8 * 4 + 4 + 4 + 8 * 8 + 7 * 4 + 4
    assert _lib.PanConfig.step_time.offset == 40 and _lib.PanConfig.q_s.offset == 104


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_cpu_fallback(lib):
    cfg = _lib.PanConfig()
    cfg.receding, cfg.max_envs, cfg.step_time, cfg.nrmp_max_num = 10, 1, 0.1, 0
    h = C.c_void_p()
    assert lib.nb_pan_create(C.byref(cfg), None, 0, None, None, C.byref(h)) == _lib.NB_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.nb_last_error()
    from neupan_b200 import PAN
    cfgw = CONFIGS["C1"]
    pan = PAN(cfgw.T, cfgw.dt, cfgw.make_robot(), dune_checkpoint=weights_path("diff"))
    with pytest.raises(_lib.NeupanB200Error):
        pan(torch.zeros(3, 11), torch.zeros(2, 10), torch.zeros(3, 11), torch.zeros(10), torch.zeros(2, 5))


def test_bad_arguments_are_rejected(lib):
    h = C.c_void_p()
    cfg = _lib.PanConfig()
    cfg.receding, cfg.max_envs, cfg.step_time = 0, 1, 0.1
    assert lib.nb_pan_create(C.byref(cfg), None, 0, None, None, C.byref(h)) == _lib.NB_ERR_INVALID
    cfg.receding, cfg.kinematics = 10, 7
    assert lib.nb_pan_create(C.byref(cfg), None, 0, None, None, C.byref(h)) == _lib.NB_ERR_INVALID
    cfg.kinematics, cfg.nrmp_max_num, cfg.edge_dim = 0, 10, 4
    assert lib.nb_pan_create(C.byref(cfg), None, 0, None, None, C.byref(h)) == _lib.NB_ERR_INVALID  # weights missing
    assert lib.nb_pan_forward(None, 1, 1, *([None] * 13), None) == _lib.NB_ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(_lib.NB_ERR_INVALID)


def test_pack_weights_order():
    from oracle import dune as od
    w = od.load_weights(weights_path("polygon"))
    flat = _lib.pack_weights(w)
    assert flat.shape == (4644,) and flat.dtype == np.float32
    assert np.array_equal(flat[:64], w["MLP.0.weight"].numpy().ravel()) and np.array_equal(flat[-4:], w["MLP.13.bias"].numpy())


def test_shard_range_partitions():
    for total, world in ((4096, 8), (10, 4), (3, 8), (1, 1)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _worker(rank, world, total, T, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    S, U = torch.randn(total, 3, T + 1, generator=g), torch.randn(total, 2, T, generator=g)
    D, md = torch.randn(total, 1, T, generator=g), torch.randn(total, generator=g)
    lo, hi = shard_range(total, rank, world)
    got = gather_results(pack_results(S[lo:hi], U[lo:hi], D[lo:hi], md[lo:hi]), total)
    s, u, d, m = unpack_results(got, T)
    q.put((rank, bool(torch.equal(s, S) and torch.equal(u, U) and torch.equal(d, D) and torch.equal(m, md))))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_world_size_2_gloo(total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + total) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, total, 10, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert res == [(0, True), (1, True)]


def test_scan_and_ipath_struct_layouts_match_header():
    # nb_scan_config: 4 + 3 + 2 doubles, 2 int32;  nb_ipath_config: 8 int32, 4 doubles
    assert C.sizeof(_lib.ScanConfig) == 9 * 8 + 2 * 4 and _lib.ScanConfig.down_sample.offset == 72
    assert C.sizeof(_lib.IpathConfig) == 8 * 4 + 4 * 8 and _lib.IpathConfig.step_time.offset == 32


def test_scan_and_ipath_reject_bad_arguments_before_touching_the_device(lib):
    cfg = _lib.ScanConfig(angle_min=-1.0, angle_max=1.0, range_min=0.1, range_max=5.0, down_sample=1)
    one = C.c_void_p(8)  # never dereferenced: validation comes first
    assert lib.nb_scan_to_points(1, 16, None, None, one, C.byref(cfg), 16, one, None, one, None) == _lib.NB_ERR_INVALID
    assert lib.nb_scan_to_points(1, 0, one, None, one, C.byref(cfg), 16, one, None, one, None) == _lib.NB_ERR_INVALID
    cfg.down_sample = 0
    assert lib.nb_scan_to_points(1, 16, one, None, one, C.byref(cfg), 16, one, None, one, None) == _lib.NB_ERR_INVALID
    assert b"down_sample" in lib.nb_last_error()
    cfg.down_sample = 1
    assert lib.nb_scan_to_points(1, 60000, one, None, one, C.byref(cfg), 16, one, None, one, None) == _lib.NB_ERR_CAPACITY
    h = C.c_void_p()
    ic = _lib.IpathConfig(receding=0, kinematics=0, max_envs=1, step_time=0.1)
    assert lib.nb_ipath_create(C.byref(ic), C.byref(h)) == _lib.NB_ERR_INVALID
    ic.receding, ic.kinematics = 10, 1  # acker without a wheelbase
    assert lib.nb_ipath_create(C.byref(ic), C.byref(h)) == _lib.NB_ERR_INVALID
    assert lib.nb_ipath_step(None, 1, one, one, 4.0, one, one, one, one, one, None) == _lib.NB_ERR_INVALID


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_scan_and_ipath_have_no_cpu_fallback(lib):
    from neupan_b200 import InitialPathBatch, scan_to_points

    h = C.c_void_p()
    ic = _lib.IpathConfig(receding=10, kinematics=0, max_envs=1, step_time=0.1, ind_range=10, arrive_index_threshold=1)
    assert lib.nb_ipath_create(C.byref(ic), C.byref(h)) == _lib.NB_ERR_NO_DEVICE
    with pytest.raises(RuntimeError):
        scan_to_points(torch.zeros(1, 3), torch.ones(1, 8), dict(angle_min=-1, angle_max=1, range_min=0.1, range_max=5))
    with pytest.raises(RuntimeError):
        InitialPathBatch(10, 0.1, "diff")
