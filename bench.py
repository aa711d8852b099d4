#!/usr/bin/env python
"""bench.py -- batched QP solves/sec on B200 (BASELINE.json metric).

Default workload (config.workload): BASELINE config 2 -- 65,536 independent
random dense QPs with nz=32, nl=8, nv=64 per GPU, default FBstab options, cold
start.  `--config 3a|3b|4a|4b|5` selects the other BASELINE configs (servo
motor / double integrator N=50, spacecraft / copolymerisation N=100, dense
nz=512) with the same JSON line; they are reported in DESIGN.md, the driver's
bench line is config 2.

A "step" is one complete batched solve of this rank's shard.  With N GPUs every
rank solves its own shard (instances are independent: no data-path collective)
and the packed results are gathered to rank 0 over NCCL inside the timed region.
Configs 2-4 keep the per-GPU batch fixed ("scaling": "weak"); config 5 shards
its 1,024 instances across the GPUs as BASELINE.json words it ("strong").

  value   solves/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e     the same through the public C-ABI with HOST (pinned) buffers: H2D of
          the problem data and D2H of the results inside the timed region
  roofline  FP64 (the binding roof for this path) against the
          DFMA / DMMA peak measured in this run, plus the HBM fraction
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

UNIT = "solves/s"

# name -> (kind, spec, instances, generator config id, rho, scaling, cpu sample per core)
CONFIGS = {
    "2": ("dense", (32, 8, 64), 65536, 2, None, "weak", 2048),
    "3a": ("mpc", ("servo_motor", 50), 16384, 3, 0.02, "weak", 96),
    "3b": ("mpc", ("double_integrator", 50), 16384, 3, -0.1, "weak", 512),
    "4a": ("mpc", ("spacecraft", 100), 4096, 4, 0.05, "weak", 2),
    "4b": ("mpc", ("copolymerization", 100), 4096, 4, 0.05, "weak", 6),
    "5": ("dense", (512, 128, 1024), 1024, 5, None, "strong", 1),
}
OCP_DIMS = {"servo_motor": (4, 1, 4), "double_integrator": (2, 1, 6),
            "spacecraft": (6, 3, 12), "copolymerization": (18, 5, 10)}


class Workload:
    """Sizes, generator and the algorithmic work model of one BASELINE config."""

    def __init__(self, name):
        self.name = name
        (self.kind, spec, self.batch, self.cfg, self.rho, self.scaling,
         self.cpu_per_core) = CONFIGS[name]
        if self.kind == "dense":
            self.nz, self.nl, self.nv = spec
            self.label = (f"batched FBstabDense {self.batch:,} random dense QPs nz={self.nz} "
                          f"nl={self.nl} nv={self.nv} (BASELINE config {name}), default "
                          "options, cold start")
            self.desc = {"nz": self.nz, "nl": self.nl, "nv": self.nv}
        else:
            self.ocp, self.N = spec
            self.nx, self.nu, self.nc = OCP_DIMS[self.ocp]
            K = self.N + 1
            self.nz, self.nl, self.nv = K * (self.nx + self.nu), K * self.nx, K * self.nc
            self.label = (f"batched FBstabMpc {self.ocp} OCP horizon N={self.N}, "
                          f"{self.batch:,} instances (BASELINE config {name}; x0 = nominal + "
                          + (f"{self.rho}*U(-1,1)" if self.rho > 0 else f"{-self.rho}*U(0,1)") +
                          "), default options, cold start")
            self.desc = {"ocp": self.ocp, "N": self.N, "nx": self.nx, "nu": self.nu,
                         "nc": self.nc, "rho": self.rho}
        self.metric = f"batched QP solves/sec ({self.label.split(' (BASELINE')[0]})"

    # ---- problem data ------------------------------------------------------
    def generate(self, fb, count, first, threads, alloc=None):
        if self.kind == "dense":
            return fb.problems.random_dense_qp(self.nz, self.nl, self.nv, count=count,
                                               config=self.cfg, first=first,
                                               nthreads=threads, alloc=alloc)
        return fb.problems.ocp_batch(self.ocp, self.N, count=count, config=self.cfg,
                                     rho=self.rho, first=first, alloc=alloc)[1]

    def solver(self, fb, max_batch, device):
        if self.kind == "dense":
            return fb.FBstabDense(self.nz, self.nl, self.nv, max_batch=max_batch, device=device)
        return fb.FBstabMpc(self.N, self.nx, self.nu, self.nc, max_batch=max_batch,
                            device=device)

    def cpu_solve(self, ob, fb, d, threads):
        if self.kind == "dense":
            return ob.dense_solve_batch(self.nz, self.nl, self.nv,
                                        *[d[k] for k in fb.problems.DENSE_FIELDS],
                                        nthreads=threads)
        return ob.mpc_solve_batch(self.N, self.nx, self.nu, self.nc,
                                  [d[k] for k in fb.problems.MPC_FIELDS], nthreads=threads)

    # ---- algorithmic work per solve ----------------
    def flops(self, newton, prox, backtracks, check_feasibility=True):
        if self.kind == "dense":
            nz, nl, nv = self.nz, self.nl, self.nv
            f_res = 2 * nz * nz + 4 * nl * nz + 2 * nv * nz
            f_init = nv * nz * (nz + 1) + nv * nz + (nz + nl) ** 3 / 3.0
            f_solve = 4 * nv * nz + 2 * (nz + nl) ** 2
        else:
            N, nx, nu, nc = self.N, self.nx, self.nu, self.nc
            ns = nx + nu
            f_res = (N + 1) * (2 * ns * ns + 2 * nc * ns) + 4 * N * nx * ns
            f_init = (N + 1) * ((10 / 3.0) * nx ** 3 + 4 * nx * nx * nu + 2 * nx * nu * nu +
                                nu ** 3 / 3.0 + nc * ns * ns + nc * ns)
            f_solve = (N + 1) * (11 * nx * nx + 10 * nx * nu + 3 * nu * nu + 4 * nc * ns)
        w = newton * (f_init + f_solve + 2 * f_res) + backtracks * f_res
        w = w + (2 * prox + 2) * f_res
        if check_feasibility:
            w = w + prox * f_res
        return w

    def data_doubles(self):
        if self.kind == "dense":
            nz, nl, nv = self.nz, self.nl, self.nv
            return nz * nz + nl * nz + nv * nz + nz + nl + nv
        N, nx, nu, nc = self.N, self.nx, self.nu, self.nc
        K = N + 1
        return (K * (nx * nx + nu * nu + nu * nx + nx + nu + nc * nx + nc * nu + nc) +
                N * (nx * nx + nx * nu + nx) + nx)

    def bytes_per_solve(self):
        nz, nl, nv = self.nz, self.nl, self.nv
        return 8 * (self.data_doubles() + (nz + nl + nv) + (nz + nl + 2 * nv)) + 48


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
            self.stop_flag.wait(0.1)

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


def cpu_reference_run(wl, count, threads, first=0):
    """Times the CPU oracle on `count` instances of the workload."""
    import fbstab_b200 as fb
    from oracle import binding as ob
    ob.build()
    d = wl.generate(fb, count, first, threads)
    t0 = time.perf_counter()
    out, z, l, v, y = wl.cpu_solve(ob, fb, d, threads)
    dt = time.perf_counter() - t0
    return count / dt, dt, out, z


def run_reference(args, wl):
    """--impl reference: the reference's CPU algorithm on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    count = max(threads, min(wl.batch, wl.cpu_per_core * threads))
    for _ in range(args.warmup):
        cpu_reference_run(wl, max(threads, count // 16), threads)
    t_total = 0.0
    for s in range(args.steps):
        _, dt, _, _ = cpu_reference_run(wl, count, threads, first=s * count)
        t_total += dt
    value = args.steps * count / t_total
    sample = (f"{count} instances per step (prefix of the {wl.batch}-instance batch), "
              f"{threads} host threads, one solver per thread")
    line = {
        "impl": "reference", "metric": wl.metric, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": wl.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict({"workload": wl.label, "instances_per_step": count}, **wl.desc),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample,
                         "note": "restated reference (oracle/), Eigen is not in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


_JSON_FD = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, whatever libraries print (NCCL's
    version banner reached stdout at 8 ranks despite NCCL_DEBUG_FILE): keep the
    real stdout for the result line and point fd 1 at stderr for everyone else."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    wl = Workload(args.config)
    if args.impl == "reference":
        run_reference(args, wl)
        return

    import torch
    import torch.distributed as dist
    import fbstab_b200 as fb
    from fbstab_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    total = args.batch or wl.batch
    if wl.scaling == "weak":
        B, first, global_batch = total, rank * total, world * total
    else:
        lo, hi = sharding.shard_range(total, world, rank)
        B, first, global_batch = hi - lo, lo, total
    cap = max(sharding.shard_sizes(global_batch, world))
    W = max(args.warmup, 3)
    K = args.steps
    threads = max(1, host_threads() // max(world, 1))
    nz, nl, nv = wl.nz, wl.nl, wl.nv

    # ---- synthetic inputs: this rank's shard of the global batch, pinned host
    pin = lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
    d_host = wl.generate(fb, B, first, threads, alloc=pin)
    d_dev = {k: torch.from_numpy(a).to(dev) for k, a in d_host.items()}
    solver = wl.solver(fb, B, local_rank)
    f64 = lambda n: torch.zeros(n, dtype=torch.float64, device=dev)
    z, l, v, y = f64(B * nz), f64(B * nl), f64(B * nv), f64(B * nv)
    out = torch.zeros(B * fb.OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step_device(ev0=None, ev1=None):
        z.zero_(), l.zero_(), v.zero_()  # cold start
        if ev0 is not None:
            ev0.record(stream)
        solver.solve_batch(d_dev, z, l, v, y=y, out=out, stream=stream.cuda_stream)
        if ev1 is not None:
            ev1.record(stream)
        if world > 1:  # result gather to rank 0 (the only collective on the path)
            packed = sharding.pack(torch, z, l, v, y, out, B, cap, (nz, nl, nv))
            bufs = [torch.empty_like(packed) for _ in range(world)] if rank == 0 else None
            dist.gather(packed, bufs, dst=0)

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
        step_device(ker0[s], ker1[s])
    e1.record(stream)
    barrier()
    clocks = sampler.summary()
    ms_total = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(ker0, ker1)]))
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    value = global_batch * K / (ms_total * 1e-3)
    launches = K * solver.last_launches

    o = np.frombuffer(out.cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE)
    flags = np.bincount(o["eflag"], minlength=6).tolist()

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region
    # (z, l, v are in/out: every timed step gets its own zeroed cold-start buffers,
    # prepared before the clock starts -- clearing 54 MB of host memory is not part
    # of a solve)
    Ke = max(1, min(K, 3))
    yh = pin(B * nv)
    oh = np.frombuffer(torch.empty(B * fb.OUT_DTYPE.itemsize, dtype=torch.uint8,
                                   pin_memory=True).numpy(), dtype=fb.OUT_DTYPE)
    warm = []
    for _ in range(Ke + 1):
        bufs = (pin(B * nz), pin(B * nl), pin(B * nv))
        for b_ in bufs:
            b_[:] = 0
        warm.append(bufs)

    def step_host(i):
        zh, lh, vh = warm[i]
        solver.solve_batch(d_host, zh, lh, vh, y=yh, out=oh, stream=stream.cuda_stream)

    step_host(Ke)
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        step_host(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    zh, lh, vh = warm[Ke - 1]
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = global_batch * Ke / float(te.item())
    h2d = sum(a.nbytes for a in d_host.values()) + zh.nbytes + lh.nbytes + vh.nbytes
    d2h = zh.nbytes + lh.nbytes + vh.nbytes + yh.nbytes + oh.nbytes
    assert (oh["eflag"] == o["eflag"]).all(), "host and device paths disagree"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the persistent solve kernel)
    dfma, dmma = fb.capi.fp64_peak(local_rank)
    W_total = float(wl.flops(o["newton_iters"].astype(np.float64),
                             o["prox_iters"].astype(np.float64),
                             o["ls_backtracks"].astype(np.float64)).sum())
    B_total = float(wl.bytes_per_solve()) * B
    ach_tf = W_total / (kernel_ms * 1e-3) / 1e12
    ach_gbs = B_total / (kernel_ms * 1e-3) / 1e9
    hbm_peak, hbm_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        pass
    tensor_bound = args.config == "5"  # A'GammaA + Cholesky updates run as DMMA
    peak = dmma if tensor_bound else dfma
    # DRAM traffic (dram__bytes_read + dram__bytes_write per instance) from the committed
    # ncu --set full captures: profiles/r1_dense_small_ncu_full.txt (29.7 KB), the ncu
    # section of profiles/r1_dense_large_phases.txt (775 GB / 296 instances) and
    # profiles/r1_mpc_lane_ncu.txt (34.7 GB / 4,736 instances, common-data path)
    per_instance = {"2": 29659.0, "5": 2.618e9, "3a": 7.33e6}
    traffic = per_instance[args.config] * B if args.config in per_instance else None
    roofline = {
        "bound": "tensor" if tensor_bound else "fp64", "achieved": ach_tf, "peak": peak,
        "unit": "TFLOP/s", "frac": ach_tf / peak, "traffic": traffic,
        "peak_source": f"measured in this run (fbstab_fp64_peak): DFMA loop {dfma:.1f}, "
                       f"FP64 mma.sync (DMMA) loop {dmma:.1f} TFLOP/s",
        "kernel": solver.path, "kernel_ms": kernel_ms,
        "algorithmic_flops_per_launch": W_total,
        "algorithmic_bytes_per_launch": B_total,
        "work_model": "SURVEY.md 8(a) W_flops with this run's per-instance newton / prox / "
                      "backtrack counters",
        "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach_gbs / hbm_peak, "peak_source": hbm_src},
    }

    # ---- CPU baseline: the oracle on a bounded prefix of the same instances
    cpu = None
    if not args.no_cpu:
        cores = host_threads()
        count = max(1, min(B, wl.cpu_per_core * cores))
        cv, cdt, co, cz = cpu_reference_run(wl, count, cores, first=first)
        match = ((co["newton_iters"] == o["newton_iters"][:count]) &
                 (co["prox_iters"] == o["prox_iters"][:count]) &
                 (co["eflag"] == o["eflag"][:count]))
        zg = z[:count * nz].cpu().numpy()
        okf = (co["eflag"] == 0) & (o["eflag"][:count] == 0) & match
        Z, CZ = zg.reshape(count, nz), cz.reshape(count, nz)
        err = float(max([np.abs(Z[i] - CZ[i]).max() / max(1.0, np.abs(CZ[i]).max())
                         for i in np.nonzero(okf)[0]] or [0.0]))
        cpu = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {count} of the {B} instances of rank 0's shard, "
                         f"{cdt:.1f} s on {cores} host threads (one solver per thread)",
               "note": "restated reference (oracle/): Eigen is not in this image",
               "same_flags": bool((co["eflag"] == o["eflag"][:count]).all()),
               "same_trajectory_frac": float(match.mean()),
               "max_rel_solution_diff_same_trajectory": err}

    line = {
        "metric": wl.metric, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_total / K,
        "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict({"workload": wl.label, "instances_per_gpu": B,
                        "global_batch": global_batch,
                        "l2": f"inputs ({sum(a.nbytes for a in d_host.values()) / 1e6:.0f} MB "
                              "per GPU) exceed the 126 MB L2; no flush needed",
                        "exit_flags": flags, "path": solver.path,
                        "newton_iters_mean": float(o["newton_iters"].mean())}, **wl.desc),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
