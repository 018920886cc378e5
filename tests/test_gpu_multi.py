"""Multi-GPU parity on real hardware: a batch sharded over 2 GPUs with the NCCL gather
returns, bit for bit, what one GPU returns for the whole batch.  Needs >= 2 CUDA devices (skipped otherwise; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, cname, out_path):
    import torch.distributed as dist

    from gpu_helpers import make_pan
    from helpers import CONFIGS, make_inputs
    from neupan_b200.parallel import ShardedPAN, shard_range

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = CONFIGS[cname]
    lo, hi = shard_range(total, rank, world)
    inp = make_inputs(cfg, B=hi - lo, env_offset=lo)
    pan = make_pan(cfg, K=3, max_envs=hi - lo)
    sp = ShardedPAN(pan, total)
    dev = {k: (None if v is None else torch.from_numpy(v).cuda()) for k, v in inp.items()}
    packed = sp.step(dev["nom_s"], dev["nom_u"], dev["ref_s"], dev["ref_us"], dev["points"], dev["velocities"])
    host = {k: (None if v is None else torch.from_numpy(v).pin_memory()) for k, v in inp.items()}
    packed_h = sp.step(host["nom_s"], host["nom_u"], host["ref_s"], host["ref_us"], host["points"], host["velocities"])
    if rank == 0:
        np.savez(out_path, packed=packed.cpu().numpy(), packed_host=packed_h.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("cname,total", [("C4", 37), ("C5", 10)])
def test_two_gpu_shard_equals_single_gpu_bitwise(cname, total, tmp_path):
    import torch.multiprocessing as mp

    from gpu_helpers import make_pan, run_pan
    from helpers import CONFIGS, make_inputs
    from neupan_b200.parallel import pack_results, unpack_results

    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), total, cname, out), nprocs=2, join=True)
    z = np.load(out)
    cfg = CONFIGS[cname]
    inp = make_inputs(cfg, B=total)
    pan = make_pan(cfg, K=3, max_envs=total)
    S, U, D, md = run_pan(pan, inp)
    single = pack_results(torch.from_numpy(S), torch.from_numpy(U), torch.from_numpy(D), torch.from_numpy(md)).numpy()
    assert z["packed"].shape == single.shape
    assert np.array_equal(z["packed"], single)  # odd total: the shards differ in size (19 + 18) and are padded for the collective
    assert np.array_equal(z["packed_host"], single)
    S2, U2, D2, md2 = unpack_results(torch.from_numpy(z["packed"]), cfg.T)
    assert np.array_equal(S2.numpy(), S) and np.array_equal(md2.numpy(), md)
