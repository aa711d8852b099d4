#!/usr/bin/env python
"""bench.py -- batched QP solves/sec on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE config 2 -- 65,536 independent random
dense QPs with nz=32, nl=8, nv=64 per GPU, default FBstab options, cold start.
A "step" is one complete batched solve of that shard.  With N GPUs every rank
solves its own 65,536-instance shard of one N*65,536-instance batch (instances
are independent: no data-path collective) and the packed results are gathered
to rank 0 over NCCL inside the timed region -> "scaling": "weak".

  value   solves/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e     the same through the public C-ABI with HOST (pinned) buffers: H2D of
          the problem data and D2H of the results inside the timed region
  roofline  FP64 (the binding roof for this path) against a
          DFMA peak measured in this run, plus the HBM fraction
  cpu_baseline  the CPU oracle (restated reference, no Eigen in this image) on
          the host cores, on a bounded prefix of the same instances

`--impl reference` times that CPU implementation alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NZ, NL, NV = 32, 8, 64
BATCH = 65536
CONFIG_ID = 2
METRIC = "batched QP solves/sec (FBstabDense nz=32 nl=8 nv=64, 65,536 instances per GPU)"
UNIT = "solves/s"
WORKLOAD = ("batched FBstabDense 65,536 random dense QPs nz=32 nl=8 nv=64 per GPU "
            "(BASELINE config 2), default options, cold start")


# ---- algorithmic work per solve ---------------------
def work_flops(newton, prox, backtracks, check_feasibility=True):
    nz, nl, nv = NZ, NL, NV
    f_res = 2 * nz * nz + 4 * nl * nz + 2 * nv * nz
    f_init = nv * nz * (nz + 1) + nv * nz + (nz + nl) ** 3 / 3.0
    f_solve = 4 * nv * nz + 2 * (nz + nl) ** 2
    w = newton * (f_init + f_solve + 2 * f_res) + backtracks * f_res
    w = w + (2 * prox + 2) * f_res
    if check_feasibility:
        w = w + prox * f_res
    return w


def bytes_per_solve():
    data = NZ * NZ + NL * NZ + NV * NZ + NZ + NL + NV
    return 8 * (data + (NZ + NL + NV) + (NZ + NL + 2 * NV)) + 48


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                o = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                     "--format=csv,noheader,nounits"], capture_output=True, text=True,
                    timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names)
                   if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows),
                "samples": len(self.rows), "reasons": reasons}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(count, threads, first=0):
    """Times the CPU oracle on `count` instances of the workload."""
    import fbstab_b200 as fb
    from oracle import binding as ob
    ob.build()
    d = fb.problems.random_dense_qp(NZ, NL, NV, count=count, config=CONFIG_ID,
                                    first=first, nthreads=threads)
    args = [d[k] for k in fb.problems.DENSE_FIELDS]
    t0 = time.perf_counter()
    out, z, l, v, y = ob.dense_solve_batch(NZ, NL, NV, *args, nthreads=threads)
    dt = time.perf_counter() - t0
    return count / dt, dt, out, z


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    # ~1.5 ms per solve per core: size each step for a few seconds of work
    count = max(256, min(BATCH, 1024 * threads // 4))
    for _ in range(args.warmup):
        cpu_reference_run(min(count, 256), threads)
    t_total = 0.0
    for s in range(args.steps):
        _, dt, _, _ = cpu_reference_run(count, threads, first=s * count)
        t_total += dt
    value = args.steps * count / t_total
    sample = (f"{count} instances per step (prefix of the {BATCH}-instance shard), "
              f"{threads} host threads, one solver per thread")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "nz": NZ, "nl": NL, "nv": NV,
                   "instances_per_step": count},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample,
                         "note": "restated reference (oracle/), Eigen is not in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=BATCH, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import fbstab_b200 as fb

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    W = max(args.warmup, 3)
    K = args.steps
    threads = max(1, host_threads() // max(world, 1))

    # ---- synthetic inputs: this rank's shard of the global batch, pinned host
    pin = lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
    d_host = fb.problems.random_dense_qp(NZ, NL, NV, count=B, config=CONFIG_ID,
                                         first=rank * B, nthreads=threads, alloc=pin)
    d_dev = {k: torch.from_numpy(a).to(dev) for k, a in d_host.items()}
    solver = fb.FBstabDense(NZ, NL, NV, max_batch=B, device=local_rank)
    f64 = lambda n: torch.zeros(n, dtype=torch.float64, device=dev)
    z, l, v, y = f64(B * NZ), f64(B * NL), f64(B * NV), f64(B * NV)
    out = torch.zeros(B * fb.OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    rec = (NZ + NL + 2 * NV) * 8 + fb.OUT_DTYPE.itemsize
    packed = torch.empty(B * rec, dtype=torch.uint8, device=dev)
    gathered = ([torch.empty_like(packed) for _ in range(world)]
                if (world > 1 and rank == 0) else None)
    stream = torch.cuda.current_stream()

    def step_device():
        z.zero_(), l.zero_(), v.zero_()  # cold start
        solver.solve_batch(d_dev, z, l, v, y=y, out=out, stream=stream.cuda_stream)
        if world > 1:  # result gather to rank 0 (the only collective on the path)
            off = 0
            for t in (z, l, v, y):
                nb = t.numel() * 8
                packed[off:off + nb].copy_(t.view(torch.uint8))
                off += nb
            packed[off:].copy_(out)
            dist.gather(packed, gathered, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ker0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ker1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    e0.record(stream)
    for s in range(K):
        z.zero_(), l.zero_(), v.zero_()
        ker0[s].record(stream)
        solver.solve_batch(d_dev, z, l, v, y=y, out=out, stream=stream.cuda_stream)
        ker1[s].record(stream)
        if world > 1:
            off = 0
            for t in (z, l, v, y):
                nb = t.numel() * 8
                packed[off:off + nb].copy_(t.view(torch.uint8))
                off += nb
            packed[off:].copy_(out)
            dist.gather(packed, gathered, dst=0)
    e1.record(stream)
    barrier()
    clocks = sampler.summary()
    ms_total = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(ker0, ker1)]))
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    value = world * B * K / (ms_total * 1e-3)
    launches = K * solver.last_launches

    o = np.frombuffer(out.cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE)
    flags = np.bincount(o["eflag"], minlength=6).tolist()

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region
    zh, lh, vh, yh = pin(B * NZ), pin(B * NL), pin(B * NV), pin(B * NV)
    oh = np.frombuffer(torch.empty(B * fb.OUT_DTYPE.itemsize, dtype=torch.uint8,
                                   pin_memory=True).numpy(), dtype=fb.OUT_DTYPE)

    def step_host():
        zh[:] = 0
        lh[:] = 0
        vh[:] = 0
        solver.solve_batch(d_host, zh, lh, vh, y=yh, out=oh, stream=stream.cuda_stream)

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / float(te.item())
    h2d = sum(a.nbytes for a in d_host.values()) + zh.nbytes + lh.nbytes + vh.nbytes
    d2h = zh.nbytes + lh.nbytes + vh.nbytes + yh.nbytes + oh.nbytes
    assert (oh["eflag"] == o["eflag"]).all(), "host and device paths disagree"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the persistent solve kernel)
    dfma, dmma = fb.capi.fp64_peak(local_rank)
    W_total = float(work_flops(o["newton_iters"].astype(np.float64),
                               o["prox_iters"].astype(np.float64),
                               o["ls_backtracks"].astype(np.float64)).sum())
    B_total = float(bytes_per_solve()) * B
    ach_tf = W_total / (kernel_ms * 1e-3) / 1e12
    ach_gbs = B_total / (kernel_ms * 1e-3) / 1e9
    hbm_peak, hbm_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        pass
    roofline = {
        "bound": "fp64", "achieved": ach_tf, "peak": dfma, "unit": "TFLOP/s",
        "frac": ach_tf / dfma, "traffic": None,
        "peak_source": "DFMA loop measured in this run (fbstab_fp64_peak); DMMA "
                       f"mma.sync peak {dmma:.1f} TFLOP/s",
        "kernel": solver.path, "kernel_ms": kernel_ms,
        "algorithmic_flops_per_launch": W_total,
        "algorithmic_bytes_per_launch": B_total,
        "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach_gbs / hbm_peak, "peak_source": hbm_src},
    }

    # ---- CPU baseline: the oracle on a bounded prefix of the same instances
    cpu = None
    if not args.no_cpu:
        cores = host_threads()
        count = max(256, min(B, 768 * cores // 2))
        cv, cdt, co, cz = cpu_reference_run(count, cores, first=0)
        match = bool((co["newton_iters"] == o["newton_iters"][:count]).all() and
                     (co["prox_iters"] == o["prox_iters"][:count]).all() and
                     (co["eflag"] == o["eflag"][:count]).all())
        zg = z[:count * NZ].cpu().numpy()
        err = float(np.abs(zg - cz).max() / max(1.0, np.abs(cz).max()))
        cpu = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {count} of the {B} instances of rank 0's shard, "
                         f"{cdt:.1f} s on {cores} host threads (one solver per thread)",
               "note": "restated reference (oracle/): Eigen is not in this image",
               "trajectory_matches_gpu": match, "max_rel_solution_diff": err}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_total / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "nz": NZ, "nl": NL, "nv": NV,
                   "instances_per_gpu": B, "global_batch": world * B,
                   "l2": "inputs (1.9 GB per GPU) exceed the 126 MB L2; no flush needed",
                   "exit_flags": flags, "path": solver.path,
                   "newton_iters_mean": float(o["newton_iters"].mean())},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
