#!/usr/bin/env python
"""bench.py -- PAN control steps/sec on the BASELINE.json north-star workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C4]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" = one batched PAN.forward (iter_num fixed at the config's K by iter_threshold = 0) over
B environments per GPU (weak scaling: every rank owns B envs of the same seeded family, env ids
offset by rank*B; at the end of each step the per-env trajectories are gathered with one NCCL
all_gather, inside the timed region).  Prints ONE JSON line (rank 0).

  value ...... env-steps/s, inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e ........ same metric through the public API with pinned HOST tensors (H2D + D2H inside)
  roofline ... DUNE kernel (the dominant launch): algorithmic GEMM FLOPs / CUDA-event duration
  cpu_baseline the CPU oracle (port of the reference path) on the box's host cores, bounded sample

--impl reference times that CPU path alone, on all host cores (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "PAN control steps/sec (batched envs, T x N x K)"
UNIT = "env-steps/s"
F_PT = 8576  # GEMM FLOPs per point-step at E=4: 2*(2*32 + 4*32*32 + 32*4)   


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        z = json.load(open(path))
        return dict(hbm_gbs=z["hbm_gbs"], tflops=z["bf16_tflops_sustained"], which="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, which="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


# --------------------------------------------------------------------------------------------
# CPU path (the oracle port of the reference: oracle/pan.py) -- cpu_baseline and --impl reference
# --------------------------------------------------------------------------------------------
def _cpu_worker(args):
    cname, env_ids, K = args
    import torch

    torch.set_num_threads(1)
    from helpers import CONFIGS, make_inputs, oracle_factory

    cfg = CONFIGS[cname]
    mk = oracle_factory(cfg, K=K)
    t0 = time.perf_counter()
    for e in env_ids:
        inp = make_inputs(cfg, B=1, env_offset=e)
        pan = mk()
        vel = None if inp["velocities"] is None else inp["velocities"][0]
        pan.forward(inp["nom_s"][0], inp["nom_u"][0], inp["ref_s"][0], inp["ref_us"][0], inp["points"][0], vel)
    return time.perf_counter() - t0


class CpuPool:
    """Persistent worker processes (one per core, one thread each); start-up and imports are paid
    once in a warm-up task so that the timed region only contains oracle work."""

    def __init__(self, cname: str, K: int, procs: int):
        import multiprocessing as mp

        for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
            os.environ[var] = "1"  # inherited by the spawned workers
        self.cname, self.K, self.procs = cname, K, procs
        self.pool = mp.get_context("spawn").Pool(procs)
        self.pool.map(_cpu_worker, [(cname, [10 ** 6 + i], 1) for i in range(procs)])  # warm-up: imports, page-in

    def rate(self, n_envs: int, first_env: int = 0):
        """env-steps/s over n_envs environments spread over the workers."""
        chunks = [list(range(first_env + i, first_env + n_envs, self.procs)) for i in range(self.procs)]
        chunks = [c for c in chunks if c]
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, [(self.cname, c, self.K) for c in chunks], chunksize=1)
        wall = time.perf_counter() - t0
        return n_envs / wall, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    from helpers import CONFIGS

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.workload]
    cores = os.cpu_count() or 1
    per_step = 2 * cores
    pool = CpuPool(args.workload, cfg.K, cores)
    for w in range(min(args.warmup, 1)):
        pool.rate(cores, first_env=5 * 10 ** 5)
    times = []
    for i in range(max(1, min(args.steps, 4))):
        rate, wall = pool.rate(per_step, first_env=i * per_step)
        times.append(wall)
    pool.close()
    ms = 1e3 * float(np.mean(times))
    value = per_step / (ms / 1e3)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=len(times), warmup=min(args.warmup, 1), ms_per_step=ms, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32 (MLP) / f64 (QP)", data="synthetic", impl="reference",
                config=dict(workload=f"{args.workload} {cfg.name}: T={cfg.T} N={cfg.N} K={cfg.K} M={cfg.M}", envs_per_step=per_step),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port",
                                  sample=f"{per_step} envs/step x {len(times)} steps of {args.workload} on {cores} worker processes; oracle/pan.py (reference DUNE code restated + float64 IPM for the ECOS solve; cvxpylayers/ECOS not installable)"),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from gpu_helpers import make_pan
    from helpers import CONFIGS, make_inputs
    from neupan_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS[args.workload]
    B = args.envs or cfg.B
    T, N, K = cfg.T, cfg.N, cfg.K

    # rotating input sets (distinct envs) so consecutive steps never see the same data; plus an L2 flush
    n_sets = 2
    sets = []
    for s in range(n_sets):
        inp = make_inputs(cfg, B=B, env_offset=(rank * n_sets + s) * B)
        sets.append({k: (None if v is None else torch.from_numpy(v)) for k, v in inp.items()})
    dsets = [{k: (None if v is None else v.to(dev)) for k, v in s.items()} for s in sets]
    hsets = [{k: (None if v is None else v.pin_memory()) for k, v in s.items()} for s in sets]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    pan = make_pan(cfg, K=K, iter_threshold=args.iter_threshold, max_envs=B, overlap=args.overlap, dune_kernel=args.dune_kernel)
    lib = _lib.load()
    per_env = 3 * (T + 1) + 2 * T + T + 1
    packed = torch.empty(B, per_env, device=dev)
    gathered = torch.empty(world * B, per_env, device=dev) if world > 1 else None

    def step(i, host=False):
        d = (hsets if host else dsets)[i % n_sets]
        S, U, D = pan(d["nom_s"], d["nom_u"], d["ref_s"], d["ref_us"], d["points"], d["velocities"])
        if host:
            return S
        if world > 1:  # the one exchange of the path: gather per-env trajectories
            torch.cat([S.reshape(B, -1), U.reshape(B, -1), D.reshape(B, -1), pan.min_distance.reshape(B, 1)], 1, out=packed)
            dist.all_gather_into_tensor(gathered, packed)
        return S

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        flush.zero_()
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.nb_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.zero_()  # L2 flush, outside the per-step events
        ev[i][0].record()
        step(i)
        ev[i][1].record()
    barrier()
    launches = (lib.nb_launch_count() - l0) // max(1, args.steps)
    step_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
    ms = float(step_ms.item())
    clocks = sampler.stop() if rank == 0 else None
    status_bad = int((pan.status != 0).sum().item())

    # ---- e2e: host tensors through the public API (H2D + compute + D2H per step) ----------------
    for i in range(2):
        step(i, host=True)
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(2, min(args.steps, 5))
    for i in range(n_e2e):
        step(i, host=True)
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / n_e2e], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    h2d = 4 * B * (2 * 3 * (T + 1) + 2 * T + T + (2 * N) * (2 if cfg.dynamic else 1))
    d2h = 4 * B * (3 * (T + 1) + 2 * T + T + 1) + 8 * B

    # ---- roofline of the dominant kernel (DUNE), CUDA events around back-to-back launches ----------
    roof = None
    if rank == 0:
        import ctypes as C

        d = dsets[0]
        p = lambda x: None if x is None else C.c_void_p(x.data_ptr())
        md = torch.empty(B, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        reps = 5
        for _ in range(2):
            _lib.check(lib.nb_dune_forward(pan._handle, B, N, p(d["nom_s"]), p(d["points"]), p(d["velocities"]), None, p(md), st))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for r in range(reps):
            flush.zero_()
            a.record()
            _lib.check(lib.nb_dune_forward(pan._handle, B, N, p(d["nom_s"]), p(d["points"]), p(d["velocities"]), None, p(md), st))
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        dune_ms = tot / reps
        flops = float(F_PT) * B * (T + 1) * N
        pk = peaks()
        ach = flops / (dune_ms * 1e-3) / 1e12
        alg_bytes = 4.0 * B * ((2 * N) * (2 if cfg.dynamic else 1) + 3 * (T + 1)) + 4.0 * B * (T + 1) * cfg.M * 9
        # executed tensor work (3 passes of the fp16 hi/lo split): per 128-point tile 4 layers x 6 UMMA (128x32x16) + head 6 UMMA (128x16x16)
        if args.dune_kernel == 2:
            tiles = B * (T + 1) * ((N + 127) // 128)
            exec_flops = tiles * 5 * 7 * 2.0 * 128 * 32 * 16  # 5 dense layers x (bias product + 6 UMMA 128x32x16)
            kname = "dune_tcp_kernel (tcgen05.mma kind::f16, A from TMEM, fp16 hi/lo split, 3 passes + bias product; two tiles in flight per CTA; FFMA2/FHFMA epilogues; SASS UTCHMMA / LDTM / STTM)"
        else:
            tiles = B * (T + 1) * ((N + 31) // 32 * 2 + 1)
            exec_flops = tiles * (4 * 24 + 6) * 2.0 * 16 * 8 * 16
            kname = "dune_mma_kernel (mma.sync m16n8k16 f16, hi/lo split, 3 passes; SASS HMMA.16816.F32)" if args.dune_kernel == 1 else "dune_kernel (all-FP32 FFMA reference variant)"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dune_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload)
        roof = dict(bound="tensor", achieved=ach, peak=pk["tflops"], unit="TFLOP/s", frac=ach / pk["tflops"], traffic=traffic,
                    kernel=kname, kernel_ms=dune_ms, share_of_step=K * dune_ms / ms, peak_source=pk["which"],
                    algorithmic_flops_per_launch=flops, tensor_executed_tflops=exec_flops / (dune_ms * 1e-3) / 1e12,
                    algorithmic_bytes_per_launch=alg_bytes, hbm_gbs_if_bytes_only=alg_bytes / (dune_ms * 1e-3) / 1e9, hbm_peak_gbs=pk["hbm_gbs"],
                    note="compute-bound path (SURVEY 8d): HBM < 1% utilised; the GEMMs are 32-wide slices between per-point LayerNorm/tanh, "
                         "so the kernel is bound by instruction issue and the MUFU pipe (2 MUFU per tanh), not by the tensor pipe -- DESIGN.md 3.1")

    # ---- cpu baseline (rank 0, N = 1 only), bounded sample ------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        pool = CpuPool(args.workload, K, cores)
        n = 2 * cores
        rate, wall = pool.rate(n)
        pool.close()
        cpu = dict(value=rate, unit=UNIT, cores=cores, kind="port",
                   sample=f"{n} envs of {args.workload} (K={K}) over {cores} worker processes (1 thread each), {wall:.1f} s wall; oracle/pan.py")

    if rank == 0:
        value = world * B / (ms * 1e-3)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32-accurate fp16 hi/lo split on tcgen05 tensor cores (ObsPointNet) / f64 (NRMP interior point)",
                    data="synthetic",
                    config=dict(workload=f"{args.workload} {cfg.name}: B={B}/GPU T={T} N={N} K={K} M={cfg.M} dyn={cfg.dynamic}" + (f" iter_threshold={args.iter_threshold} (early stop active)" if args.iter_threshold > 0 else ""), global_batch=world * B,
                                parallelism=f"env-sharded x{world}, one all_gather of {per_env} floats/env per step",
                                l2="2 rotating input sets + 256 MiB flush between timed steps", scene="annulus (SURVEY 8d)"),
                    e2e=dict(value=world * B / (float(e2e_ms.item()) * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             ms_per_step=float(e2e_ms.item()), api="neupan_b200.PAN.forward on pinned host tensors -> nb_pan_forward_host"),
                    gpu_launches=int(launches), clocks=clocks, roofline=roof, cpu_baseline=cpu, envs_with_solver_status=status_bad,
                    control_steps_per_s=world / (ms * 1e-3))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# --workload scan: the lidar scan -> points producer (SURVEY 8f "next" row 2), measured to the same bar
def run_scan(args):
    """B scans of R beams with per-beam velocities -> points for PAN (decimated to max_points = 500, C4's N).
    metric: scans/s; roofline: HBM (a byte-moving kernel).  Algorithmic bytes per scan: 4R (ranges) + 24 (state) read,
    per output column 8 (point) + 8 (velocity) written and 8 (velocity gather) read, 4 (count) written."""
    import time

    import torch

    from neupan_b200 import _lib, scan_to_points
    from oracle import scan as oscan

    B, R, MP = args.envs or 16384, 1080, 500
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    scan = dict(angle_min=-np.pi, angle_max=np.pi, range_min=0.1, range_max=10.0)
    off, n_sets = (0.25, 0.0, 0.1), 3
    rng = np.random.default_rng(4321)
    host = [dict(ranges=torch.from_numpy(rng.uniform(0.0, 11.5, size=(B, R)).astype(np.float32)).pin_memory(),
                 velocity=torch.from_numpy(rng.uniform(-1, 1, size=(B, 2, R)).astype(np.float32)).pin_memory(),
                 states=torch.from_numpy(np.stack([rng.uniform(-5, 5, B), rng.uniform(-5, 5, B), rng.uniform(-np.pi, np.pi, B)], 1)).pin_memory())
            for _ in range(n_sets)]
    devs = [{k: v.to(dev) for k, v in h.items()} for h in host]  # 3 x 212 MB of inputs: larger than the 126 MB L2
    lib = _lib.load()

    def step(i, from_host=False):
        d = (host if from_host else devs)[i % n_sets]
        if from_host:
            d = {k: v.to(dev, non_blocking=True) for k, v in d.items()}
        pts, vel, cnt = scan_to_points(d["states"], d["ranges"], scan, off, max_points=MP, velocity=d["velocity"])
        return (pts, vel, cnt.cpu()) if from_host else (pts, vel, cnt)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    l0 = lib.nb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = step(i)
    e1.record()
    torch.cuda.synchronize()
    launches = lib.nb_launch_count() - l0
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop()
    n_out = float(out[2].double().mean())
    # end to end: pinned host scans in, counts read back (points stay on the device for PAN)
    for i in range(2):
        step(i, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i, True)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    alg = B * (4.0 * R + 24 + n_out * 24 + 4)
    pk = peaks()
    roof = dict(bound="hbm", achieved=alg / (ms * 1e-3) / 1e9, peak=pk["hbm_gbs"], unit="GB/s", frac=alg / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], traffic=None,
                kernel="scan_to_points_kernel (one CTA per scan: ballot/popc ordered compaction into a shared-memory list, FP64 transform, coalesced stores)",
                kernel_ms=ms, share_of_step=1.0, peak_source=pk["which"], algorithmic_bytes_per_launch=alg,
                note="the velocity array (8R B per scan) is only gathered at the kept beams; if it had to be streamed completely the bytes would be 3x")
    line = dict(metric="lidar scans/sec -> obstacle points (batched envs)", value=B / (ms * 1e-3), unit="scans/s", n_gpus=1, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64 arithmetic, f32 in/out", data="synthetic",
                config=dict(workload=f"scan: B={B} scans x R={R} beams -> max_points={MP} (decimated), velocities on", global_batch=B,
                            l2=f"{n_sets} rotating input sets of {B * R * 12 / 1e6:.0f} MB each (> L2)"),
                gpu_launches=int(launches), clocks=clocks, roofline=roof,
                e2e=dict(value=B / (e2e_ms * 1e-3), unit="scans/s", h2d_bytes_per_step=int(B * (R * 12 + 24)), d2h_bytes_per_step=int(B * 4), ms_per_step=e2e_ms,
                         api="neupan_b200.scan_to_points on pinned host tensors (points stay on the device for PAN.forward)"))
    if not args.no_cpu:
        n = min(B, 4096)
        t0 = time.perf_counter()
        oscan.scan_batch(host[0]["states"].numpy()[:n], host[0]["ranges"].numpy()[:n], scan, off, max_points=MP, velocity=host[0]["velocity"].numpy()[:n])
        wall = time.perf_counter() - t0
        line["cpu_baseline"] = dict(value=n / wall, unit="scans/s", cores=1, kind="port", sample=f"{n} scans, {wall:.1f} s wall; oracle/scan.py (the reference's per-beam Python loop restated)")
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# --workload ipath: check_arrive + generate_nom_ref_state for B robots (SURVEY 8f "next" row 1)
def run_ipath(args):
    """metric: env path-steps/s.  Algorithmic bytes per env and step: 24 (state) + 8T (velocities) read, the path points the
    step touches (closest-point window + T reference points, 32 B each) read, 4(3(T+1)) x 2 + 4(2T) + 4T + 12 written."""
    import time

    import torch

    from neupan_b200 import InitialPathBatch, _lib
    from oracle import ipath as oip

    B, T, NPTS = args.envs or 16384, 10, 200
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    rng = np.random.default_rng(99)

    def path(b):
        pts, x, y, th = np.empty((NPTS, 4)), 0.0, 0.0, 0.3
        step, curve = 0.4 + 0.0002 * ((b % 5) - 2), 0.02 * ((b % 7) - 3)
        for i in range(NPTS):
            pts[i] = (x, y, th, 1.0 if i < 150 else -1.0)
            g = pts[i, 3]
            x += step * np.cos(th) * g; y += step * np.sin(th) * g; th += curve
        return pts

    protos = [path(b) for b in range(35)]
    paths = [protos[b % 35] for b in range(B)]
    ipb = InitialPathBatch(T, 0.1, "diff", max_envs=B, device=dev)
    ipb.set_initial_paths(paths)
    lib = _lib.load()
    n_sets = 4
    sets = []
    for s in range(n_sets):
        k = rng.integers(0, 120, B)
        st = np.stack([np.array([protos[b % 35][k[b], 0], protos[b % 35][k[b], 1], protos[b % 35][k[b], 2]]) for b in range(B)]) + rng.normal(0, 0.02, (B, 3))
        vel = rng.uniform(-1, 1, (B, 2, T)).astype(np.float32); vel[:, 0] += 3.0
        sets.append((torch.from_numpy(st).pin_memory(), torch.from_numpy(vel).pin_memory()))
    dsets = [(a.to(dev), b.to(dev)) for a, b in sets]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(i, host=False):
        a, b = (sets if host else dsets)[i % n_sets]
        if host:
            a, b = a.to(dev, non_blocking=True), b.to(dev, non_blocking=True)
        out = ipb.step(a, b, 4.0)
        return out[4].cpu() if host else out

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    l0, tot = lib.nb_launch_count(), 0.0
    for i in range(args.steps):
        flush.zero_()  # the path window of an env is re-read every step: flush L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(i); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    launches = lib.nb_launch_count() - l0
    ms = tot / args.steps
    clocks = sampler.stop()
    step(0, True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i, True)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    alg = B * (24 + 8.0 * T + 32.0 * (10 + T) + 4.0 * (6 * (T + 1) + 2 * T + T) + 12)
    pk = peaks()
    roof = dict(bound="hbm", achieved=alg / (ms * 1e-3) / 1e9, peak=pk["hbm_gbs"], unit="GB/s", frac=alg / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], traffic=None,
                kernel="ipath_step_kernel (one thread per environment: sequential closest-point window + T-step rollout / reference walk in FP64)",
                kernel_ms=ms, share_of_step=1.0, peak_source=pk["which"], algorithmic_bytes_per_launch=alg,
                note="latency-bound sequential per-environment logic with FP64 trigonometry; ~1 KB of traffic per environment")
    line = dict(metric="initial-path steps/sec (check_arrive + generate_nom_ref_state, batched envs)", value=B / (ms * 1e-3), unit="env-steps/s", n_gpus=1,
                steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64 arithmetic, f32 out",
                data="synthetic", config=dict(workload=f"ipath: B={B} envs, T={T}, paths of {NPTS} points with a gear switch", global_batch=B, l2="L2 flushed between timed iterations"),
                gpu_launches=int(launches), clocks=clocks, roofline=roof,
                e2e=dict(value=B / (e2e_ms * 1e-3), unit="env-steps/s", h2d_bytes_per_step=int(B * (24 + 8 * T)), d2h_bytes_per_step=int(B * 4), ms_per_step=e2e_ms,
                         api="neupan_b200.InitialPathBatch.step on pinned host tensors (trajectories stay on the device for PAN.forward)"))
    if not args.no_cpu:
        n = 512
        os_ = [oip.OracleInitialPath(T, 0.1, "diff") for _ in range(n)]
        for b, o in enumerate(os_):
            o.set_initial_path([r.reshape(4, 1).copy() for r in protos[b % 35]])
        st, vel = sets[0][0].numpy(), sets[0][1].numpy().astype(np.float64)
        t0 = time.perf_counter()
        for b, o in enumerate(os_):
            if not o.check_arrive(st[b].reshape(3, 1)):
                o.generate_nom_ref_state(st[b].reshape(3, 1), vel[b], 4.0)
        wall = time.perf_counter() - t0
        line["cpu_baseline"] = dict(value=n / wall, unit="env-steps/s", cores=1, kind="port", sample=f"{n} envs, {wall:.2f} s wall; oracle/ipath.py (the reference's numpy code restated)")
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# --workload control: the whole control step of B robots on the device (PlannerBatch = ipath -> scan -> PAN)
def run_control(args):
    """C4's robot / adjust values and B = 4096, K = 10, T = 10; obstacles arrive as lidar scans (R = 1080 beams with per-beam
    velocities, decimated to 500 points).  value: device-timed with scans and states resident; e2e: pinned host states + scans in,
    actions out."""
    import time

    import torch

    from helpers import CONFIGS, weights_path
    from neupan_b200 import PlannerBatch, _lib

    cfg = CONFIGS["C4"]
    B, R, N = args.envs or cfg.B, 1080, cfg.N
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    rng = np.random.default_rng(2024)
    pl = PlannerBatch(B, cfg.T, cfg.dt, 4.0, robot_kwargs=cfg.robot_kwargs, adjust_kwargs=dict(cfg.adjust), device=dev,
                      pan_kwargs=dict(iter_num=cfg.K, dune_max_num=N, nrmp_max_num=cfg.M, dune_checkpoint=weights_path(cfg.model), iter_threshold=0.0, max_points=N))

    def path(j):
        pts, x, y, th = np.empty((150, 4)), 0.0, 0.0, 0.3
        for i in range(150):
            pts[i] = (x, y, th, 1.0)
            x += 0.4 * np.cos(th); y += 0.4 * np.sin(th); th += 0.004 * (j - 8)
        return pts

    protos = [path(j) for j in range(17)]
    pl.set_initial_paths([protos[b % 17] for b in range(B)])
    scan = dict(angle_min=-np.pi, angle_max=np.pi, range_min=0.1, range_max=10.0)
    n_sets, sets = 2, []
    for s in range(n_sets):
        k = rng.integers(0, 60, B)
        st = np.stack([protos[b % 17][k[b], :3] for b in range(B)]) + rng.normal(0, 0.02, (B, 3))
        sets.append(dict(states=torch.from_numpy(st).pin_memory(), ranges=torch.from_numpy(rng.uniform(2.0, 11.5, (B, R)).astype(np.float32)).pin_memory(),
                         vel=torch.from_numpy(rng.uniform(-1, 1, (B, 2, R)).astype(np.float32)).pin_memory()))
    dsets = [{k: v.to(dev) for k, v in d.items()} for d in sets]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.load()

    def step(i, host=False):
        d = (sets if host else dsets)[i % n_sets]
        if host:
            d = {k: v.to(dev, non_blocking=True) for k, v in d.items()}
        action, _ = pl.forward(d["states"], d["ranges"], scan, scan_velocity=d["vel"])
        return action.cpu() if host else action

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    l0, tot = lib.nb_launch_count(), 0.0
    for i in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(i); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    launches = (lib.nb_launch_count() - l0) // args.steps
    ms = tot / args.steps
    clocks = sampler.stop()
    step(0, True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i, True)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    print(json.dumps(dict(metric="robot control steps/sec (initial path + scan -> points + PAN, batched envs)", value=B / (ms * 1e-3), unit="env-steps/s", n_gpus=1,
                          steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                          dtype="f64 (path, scan) / fp16 hi-lo tcgen05 (ObsPointNet) / f64 (NRMP)", data="synthetic",
                          config=dict(workload=f"control: C4 robot, B={B}, T={cfg.T}, K={cfg.K}, scans of R={R} beams -> N={N} points, velocities on", global_batch=B,
                                      l2="L2 flushed between timed iterations"),
                          gpu_launches=int(launches), clocks=clocks,
                          e2e=dict(value=B / (e2e_ms * 1e-3), unit="env-steps/s", h2d_bytes_per_step=int(B * (R * 12 + 24)), d2h_bytes_per_step=int(B * 8), ms_per_step=e2e_ms,
                                   api="neupan_b200.PlannerBatch.forward on pinned host states + scans -> actions on the host"))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--envs", type=int, default=0, help="override B per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--dune-kernel", type=int, default=2, help="NB_OPT_DUNE_KERNEL: 0 fp32 ffma, 1 mma.sync, 2 tcgen05")
    ap.add_argument("--iter-threshold", type=float, default=0.0, help="PAN stop criterion (pan.py:243); 0 forces exactly K iterations (the headline), the reference default is 0.1")
    ap.add_argument("--overlap", type=int, default=1, help="env sub-batches pipelined on internal streams (NB_OPT_OVERLAP)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "control" and args.impl == "ours":
        run_control(args)
    elif args.workload == "ipath" and args.impl == "ours":
        run_ipath(args)
    elif args.workload == "scan" and args.impl == "ours":
        run_scan(args)
    elif args.workload == "scan":  # CPU arm of the scan stage: the per-beam loop of the reference, restated (oracle/scan.py)
        from oracle import scan as oscan
        rng = np.random.default_rng(4321)
        n, R, MP = 64, 1080, 500
        scan = dict(angle_min=-np.pi, angle_max=np.pi, range_min=0.1, range_max=10.0)
        ranges, vel = rng.uniform(0.0, 11.5, size=(n, R)).astype(np.float32), rng.uniform(-1, 1, size=(n, 2, R)).astype(np.float32)
        states = np.stack([rng.uniform(-5, 5, n), rng.uniform(-5, 5, n), rng.uniform(-np.pi, np.pi, n)], 1)
        times = []
        for _ in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            oscan.scan_batch(states, ranges, scan, (0.25, 0.0, 0.1), max_points=MP, velocity=vel)
            times.append(time.perf_counter() - t0)
        dt = float(np.mean(times[args.warmup:]))
        v = n / dt
        print(json.dumps(dict(metric="lidar scans/sec -> obstacle points (batched envs)", value=v, unit="scans/s", n_gpus=1, steps=args.steps, warmup=args.warmup,
                              ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                              config=dict(workload=f"scan: R={R} beams -> max_points={MP}", scans_per_step=n),
                              cpu_baseline=dict(value=v, unit="scans/s", cores=1, kind="port", sample=f"{n} scans/step; oracle/scan.py (per-beam Python loop of neupan.py:173-281)"),
                              e2e=dict(value=v, unit="scans/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
