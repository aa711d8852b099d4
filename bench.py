#!/usr/bin/env python
"""bench.py -- batched QP solves/sec on B200 (BASELINE.json metric).

Headline workload (config.workload): BASELINE config 2 -- 65,536 independent
random dense QPs with nz=32, nl=8, nv=64 per GPU, default FBstab options, cold
start.  The same run then measures every other BASELINE config with fewer steps
and reports them under `per_config` (same keys: value, e2e, roofline, parity):

  3a / 3b  servo motor / double integrator OCP, N=50, 16,384 instances
  4a / 4b  spacecraft / copolymerisation OCP, N=100, 4,096 instances
  4a40     spacecraft OCP at the reference's own test horizon N=40 (converges;
           at N=100 every instance runs into the Newton cap, like the reference)
  5        dense nz=512 nl=128 nv=1024, 1,024 instances (sharded: strong scaling)
  1        ONE dense QP nz=50 nl=10 nv=100: CPU latency of the restated reference
           (median of 101 solves) next to the engine's single-instance latency
  3a-sparse  the instances of 3a restated as general sparse QPs (FBstabSparse, SURVEY
           8(f4): the reference plans this solver; common pattern, values per instance)

`--config X` makes X the headline, `--per-config none` skips the rest.

A "step" is one complete batched solve of this rank's shard.  With N GPUs every
rank solves its own shard (instances are independent: no data-path collective)
and the result rows are gathered to rank 0 by the library's NCCL gather
(fbstab_multi_gpu_gather) inside the timed region -- on a second, high-priority
stream, so that the gather of step s overlaps the solve of step s+1 (the result
buffers are double-buffered; the region ends after the last gather).  `--scaling weak` keeps the
per-GPU batch fixed, `strong` shards the config's batch; default: weak for
configs 2-4, strong for config 5, as BASELINE.json words them.

  value   solves/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e     the same through the public C-ABI with HOST (pinned) buffers: H2D of
          the problem data and D2H of the results inside the timed region
  roofline  FP64 (the binding roof for this path) against the
          DFMA / DMMA peak measured in this run, plus the HBM fraction
  cpu_baseline  the CPU oracle (restated reference, no Eigen in this image) on
          the host cores, on a bounded prefix of the same instances, with the
          parity of the GPU results against it

`--impl reference` times that CPU implementation alone, on the same instances.
"""
import argparse
import hashlib
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

# name -> (kind, spec, instances, generator config id, rho, scaling, oracle sample)
CONFIGS = {
    "2": ("dense", (32, 8, 64), 65536, 2, None, "weak", 32768),
    "3a": ("mpc", ("servo_motor", 50), 16384, 3, 0.02, "weak", 2048),
    "3b": ("mpc", ("double_integrator", 50), 16384, 3, -0.1, "weak", 4096),
    "4a": ("mpc", ("spacecraft", 100), 4096, 4, 0.05, "weak", 4096),
    "4a40": ("mpc", ("spacecraft", 40), 4096, 4, 0.01, "weak", 1024),
    "4b": ("mpc", ("copolymerization", 100), 4096, 4, 0.05, "weak", 1024),
    # config 3a's instances restated as GENERAL sparse QPs (FBstabSparse)
    "3a-sparse": ("sparse", ("servo_motor", 50), 16384, 3, 0.02, "weak", 512),
    "5": ("dense", (512, 128, 1024), 1024, 5, None, "strong", 128),
}
PER_CONFIG_ORDER = ["3a", "3b", "4a", "4a40", "4b", "5", "3a-sparse"]
OCP_DIMS = {"servo_motor": (4, 1, 4), "double_integrator": (2, 1, 6),
            "spacecraft": (6, 3, 12), "copolymerization": (18, 5, 10)}


class Workload:
    """Sizes, generator and the algorithmic work model of one BASELINE config."""

    def __init__(self, name):
        self.name = name
        (self.kind, spec, self.batch, self.cfg, self.rho, self.scaling,
         self.oracle_sample) = CONFIGS[name]
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
            if self.kind == "sparse":
                self.label = self.label.replace("batched FBstabMpc", "batched FBstabSparse:")
                self.label = self.label.replace(
                    "(BASELINE config 3a-sparse", "restated as general sparse QPs on one "
                    "pattern (the instances of BASELINE config 3a")
                self._pattern = self._perm = self._fac = None
            self.desc = {"ocp": self.ocp, "N": self.N, "nx": self.nx, "nu": self.nu,
                         "nc": self.nc, "rho": self.rho}
        self.metric = f"batched QP solves/sec ({self.label.split(' (BASELINE')[0]})"

    # ---- problem data ------------------------------------------------------
    def generate(self, problems, count, first, threads, alloc=None):
        if self.kind == "dense":
            return problems.random_dense_qp(self.nz, self.nl, self.nv, count=count,
                                            config=self.cfg, first=first,
                                            nthreads=threads, alloc=alloc)
        if self.kind == "sparse":
            dims, d = problems.ocp_batch(self.ocp, self.N, count=count, config=self.cfg,
                                         rho=self.rho, first=first)
            _, pat, vals = problems.ocp_as_sparse_qp(dims, d, count, alloc=alloc)
            if self._pattern is None:
                self._pattern = pat
            assert all(np.array_equal(a, b) for a, b in zip(pat, self._pattern))
            return vals
        return problems.ocp_batch(self.ocp, self.N, count=count, config=self.cfg,
                                  rho=self.rho, first=first, alloc=alloc)[1]

    def pattern(self, problems):
        if self._pattern is None:
            self.generate(problems, 1, 0, 1)
        return self._pattern

    def solver(self, fb, max_batch, device):
        if self.kind == "dense":
            return fb.FBstabDense(self.nz, self.nl, self.nv, max_batch=max_batch, device=device)
        if self.kind == "sparse":
            s = fb.FBstabSparse(self.nz, self.nl, self.nv, self.pattern(fb.problems),
                                max_batch=max_batch, device=device)
            self._perm = s.analysis()[3]  # the CPU leg eliminates in the same order
            self._fac = s.factor_pattern()
            return s
        return fb.FBstabMpc(self.N, self.nx, self.nu, self.nc, max_batch=max_batch,
                            device=device)

    def cpu_solve(self, ob, problems, d, threads):
        if self.kind == "dense":
            return ob.dense_solve_batch(self.nz, self.nl, self.nv,
                                        *[d[k] for k in problems.DENSE_FIELDS],
                                        nthreads=threads)
        if self.kind == "sparse":
            return ob.sparse_solve_batch(self.nz, self.nl, self.nv, self.pattern(problems),
                                         [d[k] for k in problems.SPARSE_FIELDS],
                                         perm=self._perm, nthreads=threads)
        return ob.mpc_solve_batch(self.N, self.nx, self.nu, self.nc,
                                  [d[k] for k in problems.MPC_FIELDS], nthreads=threads)

    # ---- algorithmic work per solve ----------------
    def flops(self, newton, prox, backtracks, check_feasibility=True):
        if self.kind == "dense":
            nz, nl, nv = self.nz, self.nl, self.nv
            f_res = 2 * nz * nz + 4 * nl * nz + 2 * nv * nz
            f_init = nv * nz * (nz + 1) + nv * nz + (nz + nl) ** 3 / 3.0
            f_solve = 4 * nv * nz + 2 * (nz + nl) ** 2
        elif self.kind == "sparse":
            # the formulas of SURVEY 8(a) rows a3, a9, a10 with non-zero counts for the
            # products and the operation count of the up-looking LDL' for the factorisation
            import fbstab_b200.problems as problems
            Hp, Hi, Gp, Gi, Ap, Ai = self.pattern(problems)
            nnzH = 2 * len(Hi) - int((np.repeat(np.arange(self.nz), np.diff(Hp)) == Hi).sum())
            nnzG, nnzA = len(Gi), len(Ai)
            n = self.nz + self.nl + self.nv
            f_res = 2 * nnzH + 4 * nnzG + 2 * nnzA
            if self._fac is not None:
                cnt = np.diff(self._fac[0]).astype(np.float64)
                f_fac, nnzL = float((cnt * cnt + 3 * cnt).sum()), float(len(self._fac[1]))
            else:
                f_fac, nnzL = 0.0, 0.0
            f_init = nnzA + 4 * self.nv + f_fac
            f_solve = 4 * nnzL + n + 4 * nnzA
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
        if self.kind == "sparse":
            import fbstab_b200.problems as problems
            Hp, Hi, Gp, Gi, Ap, Ai = self.pattern(problems)
            return len(Hi) + len(Gi) + len(Ai) + self.nz + self.nl + self.nv
        N, nx, nu, nc = self.N, self.nx, self.nu, self.nc
        K = N + 1
        return (K * (nx * nx + nu * nu + nu * nx + nx + nu + nc * nx + nc * nu + nc) +
                N * (nx * nx + nx * nu + nx) + nx)

    def bytes_per_solve(self):
        nz, nl, nv = self.nz, self.nl, self.nv
        return 8 * (self.data_doubles() + (nz + nl + nv) + (nz + nl + 2 * nv)) + 48

    def static_config(self, world, scaling, per_config):
        """The part of `config` that both arms print: a description, no results."""
        B = self.batch if scaling == "weak" else None
        c = {"workload": self.label, "baseline_config": self.name, "scaling": scaling,
             "instances_per_gpu": B if B is not None else f"{self.batch} / n_gpus",
             "global_batch": self.batch * world if scaling == "weak" else self.batch,
             "per_config": per_config}
        c.update(self.desc)
        return c


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


def gpu_local_cpus(torch, dev_index):
    """The CPUs NVML names as local to this GPU, within the set this process may run on
    (None when unknown or when that is every allowed CPU).  Pinned host buffers are
    allocated from a thread confined to them, so that first touch puts the pages on the
    GPU's NUMA node: on a two-socket box the H2D copies of the e2e leg otherwise cross the
    socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(dev_index)
        bus = "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        local = cpus & allowed
        return local if local and local != allowed else None
    except Exception:
        return None


class on_cpus:
    """Runs the calling thread on `cpus` for the duration of the block (no-op for None)."""

    def __init__(self, cpus):
        self.cpus, self.prev = cpus, None

    def __enter__(self):
        if self.cpus:
            try:
                self.prev = os.sched_getaffinity(0)
                os.sched_setaffinity(0, self.cpus)
            except Exception:
                self.prev = None
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            os.sched_setaffinity(0, self.prev)
        return False


def csrc_hash():
    """SHA-256 over the kernel sources: profiles/traffic.json is tied to a build."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "fbstab_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(name, instances):
    """DRAM bytes per launch from the committed ncu capture of THIS build
    (profiles/traffic.json, regenerated by tools/capture_traffic.sh); None when the
    kernel sources changed since the capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)
        e = t["configs"][name]
        if t.get("csrc_sha256") != csrc_hash():
            return None
        return float(e["dram_bytes_per_instance"]) * instances
    except Exception:
        return None


# ---- CPU legs (the only users of oracle/) ----------------------------------------
def cpu_reference_run(wl, count, threads, first=0):
    """Times the CPU oracle on instances first .. first+count-1 of the workload.  The
    problem data comes from the generator library alone (no CUDA engine mapped)."""
    import fbstab_b200.problems as problems
    from oracle import binding as ob
    ob.build()
    d = wl.generate(problems, count, first, threads)
    t0 = time.perf_counter()
    out, z, l, v, y = wl.cpu_solve(ob, problems, d, threads)
    dt = time.perf_counter() - t0
    return count / dt, dt, out, z


def reference_sources_check(wl, count, threads, first, co):
    """The checker checked: on a prefix of the oracle's sample, the reference's OWN sources
    (oracle/_ref/libfbstab_ref.so: the reference's algorithm code compiled against a stand-in
    for Eigen, `make -C oracle _ref`; a prebuilt file on the GPU box) must end every instance
    with the oracle's exit flag and iteration counts.  None when the library is not there;
    never raises (the bench line does not depend on it)."""
    try:
        import fbstab_b200.problems as problems
        from oracle import binding as ob
        if wl.kind == "sparse" or ob.ref_lib() is None:
            return None
        n = min(count, 4 if wl.name == "5" else 256 if wl.kind == "dense" else 128)
        d = wl.generate(problems, n, first, threads)
        if wl.kind == "dense":
            ro = ob.ref_dense_solve_batch(wl.nz, wl.nl, wl.nv,
                                          *[d[k] for k in problems.DENSE_FIELDS],
                                          nthreads=threads)[0]
        else:
            ro = ob.ref_mpc_solve_batch(wl.N, wl.nx, wl.nu, wl.nc,
                                        [d[k] for k in problems.MPC_FIELDS],
                                        nthreads=threads)[0]
        o = co[:n]
        same = ((ro["eflag"] == o["eflag"]) & (ro["newton_iters"] == o["newton_iters"]) &
                (ro["prox_iters"] == o["prox_iters"]))
        return {"instances": int(n), "same_flags": bool((ro["eflag"] == o["eflag"]).all()),
                "same_trajectory_frac": float(same.mean()),
                "what": "oracle (the CPU baseline and parity checker) against the reference's "
                        "own algorithm sources compiled on a stand-in for Eigen (oracle/_ref)"}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}


def cpu_latency_config1(repeats=101):
    """BASELINE config 1: ONE dense QP nz=50 nl=10 nv=100 on the CPU (the reference's own
    runnable case): median wall time of `repeats` cold-start solves of the restated
    reference, next to its SolverOut.solve_time (fbstab_dense.h:139-141)."""
    import fbstab_b200.problems as problems
    from oracle import binding as ob
    ob.build()
    nz, nl, nv = 50, 10, 100
    d = problems.random_dense_qp(nz, nl, nv, count=1, config=1)
    p = ob.Problem.dense(*[d[k] for k in problems.DENSE_FIELDS])
    wall, inner, out = [], [], None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out, x, _ = p.solve()
        wall.append(time.perf_counter() - t0)
        inner.append(out["solve_time"])
    return {"workload": "FBstabDense single random dense QP nz=50 nl=10 nv=100 "
                        "(BASELINE config 1), default options, cold start",
            "cpu_latency_ms_median": 1e3 * float(np.median(wall)),
            "cpu_solve_time_ms_median": 1e3 * float(np.median(inner)),
            "repeats": repeats, "kind": "port", "exit_flag": int(out["eflag"]),
            "newton_iters": int(out["newton_iters"]), "prox_iters": int(out["prox_iters"]),
            "data": d}


def run_reference(args, wl, scaling, per_config):
    """--impl reference: the reference's CPU algorithm on the host cores, on the FIRST
    instances of the batch the GPU arm solves -- the same range every step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    # bounded sample: about one second per step for the headline shape
    per_step = {"2": 2048, "3a": 96, "3b": 512, "4a": 2, "4a40": 8, "4b": 6, "5": 1}[wl.name]
    count = max(threads, min(wl.batch, per_step * threads))
    for _ in range(args.warmup):
        cpu_reference_run(wl, max(threads, count // 16), threads)
    t_total = 0.0
    for s in range(args.steps):
        _, dt, _, _ = cpu_reference_run(wl, count, threads, first=0)
        t_total += dt
    value = args.steps * count / t_total
    sample = (f"instances 0..{count - 1} of the {wl.batch}-instance batch every step, "
              f"{threads} host threads, one solver per thread")
    # The reference's own sources on a stand-in for Eigen (oracle/_ref), same instances, one
    # timed pass: reported beside the port.  The line's value is the FASTER of the two CPU
    # implementations of the reference algorithm (so far always the port: the stand-in's
    # eager temporaries cost 1.1-3.6x), which is the one that is kinder to the CPU.
    kind, note, ref_side = "port", "restated reference (oracle/), Eigen is not in this image", None
    try:
        import fbstab_b200.problems as problems
        from oracle import binding as ob
        if wl.kind != "sparse" and ob.ref_lib() is not None:
            d = wl.generate(problems, count, 0, threads)
            t0 = time.perf_counter()
            if wl.kind == "dense":
                ro = ob.ref_dense_solve_batch(wl.nz, wl.nl, wl.nv,
                                              *[d[k] for k in problems.DENSE_FIELDS],
                                              nthreads=threads)[0]
            else:
                ro = ob.ref_mpc_solve_batch(wl.N, wl.nx, wl.nu, wl.nc,
                                            [d[k] for k in problems.MPC_FIELDS],
                                            nthreads=threads)[0]
            v_ref = count / (time.perf_counter() - t0)
            ref_side = {"value": v_ref, "unit": UNIT, "kind": "reference-sources",
                        "what": "the reference's own algorithm sources compiled against "
                                "oracle/eigen_shim (a stand-in for Eigen), oracle/_ref",
                        "exit_flags": np.bincount(ro["eflag"], minlength=6).tolist()}
            if v_ref > value:
                value, kind = v_ref, "reference"
                note = ("the reference's own sources on a stand-in for Eigen (oracle/_ref) "
                        "were faster than the restated port on this sample")
                t_total = args.steps * count / value
    except Exception as e:  # noqa: BLE001
        ref_side = {"error": str(e)[:200]}
    line = {
        "impl": "reference", "metric": wl.metric, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": wl.static_config(args.gpus, scaling, per_config),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": sample, "note": note, "reference_sources": ref_side},
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


def _log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ---- parity of the GPU results against the oracle ---------------------------------
def parity_block(wl, o, zg, co, cz, count):
    """o / zg: GPU out records and solutions of the first `count` instances;
    co / cz: the oracle's.  Nothing is hidden: off-trajectory instances are counted,
    their Newton-count differences are histogrammed and their solution error is
    reported next to the same-trajectory maximum."""
    nz = wl.nz
    same_flags = co["eflag"] == o["eflag"]
    match = ((co["newton_iters"] == o["newton_iters"]) & (co["prox_iters"] == o["prox_iters"]) &
             (co["ls_backtracks"] == o["ls_backtracks"]) & same_flags)
    Z, CZ = zg.reshape(count, nz), cz.reshape(count, nz)
    err = np.abs(Z - CZ).max(1) / np.maximum(1.0, np.abs(CZ).max(1))
    succ = (co["eflag"] == 0) & (o["eflag"] == 0)
    dn = (o["newton_iters"].astype(np.int64) - co["newton_iters"])[~match]
    hist = {str(int(k)): int(v) for k, v in zip(*np.unique(dn, return_counts=True))}
    mx = lambda m: float(err[m].max()) if m.any() else 0.0
    return {
        "sample_instances": int(count),
        "same_flags": bool(same_flags.all()),
        "flag_mismatches": int((~same_flags).sum()),
        "same_trajectory_frac": float(match.mean()),
        "off_trajectory_instances": int((~match).sum()),
        "off_trajectory_newton_diff_histogram": hist,
        "max_rel_solution_diff_all_success": mx(succ),
        "max_rel_solution_diff_same_trajectory": mx(succ & match),
        "max_rel_solution_diff_off_trajectory": mx(succ & ~match),
        "success_instances": int(succ.sum()),
    }


def trajectory_floor(wl):
    """The committed oracle-vs-oracle (FMA on / off) same-trajectory fraction of the
    family (tests/golden/trajectory_floor.json): the floor for any implementation."""
    key = {"2": "dense_32_8_64", "3a": "servo_motor_N50", "3b": "double_integrator_N50",
           "4a40": "spacecraft_N40", "4b": "copolymerization_N100",
           "5": "dense_512_128_1024", "3a-sparse": "servo_motor_N50_sparse"}.get(wl.name)
    try:
        with open(os.path.join(ROOT, "tests", "golden", "trajectory_floor.json")) as fh:
            return json.load(fh)["families"][key]["same_trajectory_frac"]
    except Exception:
        return None


# ---- one config on the GPUs ----------------------------------------------------------
def measure(env, wl, K, W, scaling, with_cpu, with_e2e=True):
    torch, dist, fb, mg = env["torch"], env["dist"], env["fb"], env["mg"]
    rank, world, local_rank, dev = env["rank"], env["world"], env["local_rank"], env["dev"]
    from fbstab_b200 import sharding
    total = wl.batch
    if scaling == "weak":
        B, first, global_batch = total, rank * total, world * total
    else:
        lo, hi = sharding.shard_range(total, world, rank)
        B, first, global_batch = hi - lo, lo, total
    threads = max(1, host_threads() // max(world, 1))
    nz, nl, nv = wl.nz, wl.nl, wl.nv
    t_gen = time.perf_counter()
    local_cpus = env.get("local_cpus")

    def pin(n):  # pinned, first touched from a CPU next to this rank's GPU
        with on_cpus(local_cpus):
            return torch.empty(max(n, 1), dtype=torch.float64, pin_memory=True).numpy()[:n]

    d_host = wl.generate(fb.problems, B, first, threads, alloc=pin)
    d_dev = {k: torch.from_numpy(a).to(dev) for k, a in d_host.items()}
    t_gen = time.perf_counter() - t_gen
    solver = wl.solver(fb, max(B, 1), local_rank)
    f64 = lambda n: torch.zeros(n, dtype=torch.float64, device=dev)
    u8 = lambda n: torch.zeros(n, dtype=torch.uint8, device=dev)
    # rank 0 solves straight into its rows of the global arrays; the gather fills the rest.
    # With several ranks the result buffers are DOUBLE-BUFFERED: the gather of step s runs
    # on a second, high-priority stream while step s+1 solves into the other set (the
    # collective overlaps the next batch's compute; every gather completes inside the
    # timed region, which ends only after the main stream has waited for the gather
    # stream).
    nset = 2 if world > 1 else 1
    sets = []
    for _ in range(nset):
        if world > 1 and rank == 0:
            full = (f64(global_batch * nz), f64(global_batch * nl), f64(global_batch * nv),
                    f64(global_batch * nv), u8(global_batch * fb.OUT_DTYPE.itemsize))
            sl = lambda t, w, full=full: t[first * w:(first + B) * w]
            bufs = (sl(full[0], nz), sl(full[1], nl), sl(full[2], nv), sl(full[3], nv),
                    sl(full[4], fb.OUT_DTYPE.itemsize))
        else:
            full = None
            bufs = (f64(B * nz), f64(B * nl), f64(B * nv), f64(B * nv),
                    u8(B * fb.OUT_DTYPE.itemsize))
        sets.append((bufs, full))
    stream = torch.cuda.current_stream()
    gstream = torch.cuda.Stream(device=dev, priority=-1) if world > 1 else None
    solved = [torch.cuda.Event() for _ in range(nset)]
    gathered = [torch.cuda.Event() for _ in range(nset)]
    state = {"step": 0, "pending": [False] * nset}

    def step_device(ev0=None, ev1=None):
        k = state["step"] % nset
        state["step"] += 1
        (z, l, v, y, out), full = sets[k]
        if state["pending"][k]:  # the gather that last read this set has to be done
            stream.wait_event(gathered[k])
        z.zero_(), l.zero_(), v.zero_()  # cold start
        if ev0 is not None:
            ev0.record(stream)
        if B:
            solver.solve_batch(d_dev, z, l, v, y=y, out=out, stream=stream.cuda_stream)
        if ev1 is not None:
            ev1.record(stream)
        if world > 1:  # result gather to rank 0 (the only collective on the path)
            solved[k].record(stream)
            gstream.wait_event(solved[k])
            mg.gather(global_batch, (nz, nl, nv), (z, l, v, y, out), full, root=0,
                      stream=gstream.cuda_stream)
            gathered[k].record(gstream)
            state["pending"][k] = True

    def last_set():
        return sets[(state["step"] - 1) % nset]

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
    if gstream is not None:
        stream.wait_stream(gstream)  # the timed region ends after the last gather
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

    (z, l, v, y, out), full = last_set()
    o = np.frombuffer(out.cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE)
    res = {"value": value, "ms_per_step": ms_total / K, "kernel_ms": kernel_ms, "clocks": clocks,
           "launches": int(launches), "B": B, "global_batch": global_batch, "path": solver.path,
           "flags": np.bincount(o["eflag"], minlength=6).tolist() if B else [0] * 6,
           "newton_mean": float(o["newton_iters"].mean()) if B else 0.0,
           "input_mb": sum(a.nbytes for a in d_host.values()) / 1e6, "gen_s": t_gen}

    # ---- gathered bytes == a single-GPU solve of the same instances (rank 0 re-solves a
    # prefix of EVERY rank's range on its own GPU and compares every byte)
    if world > 1:
        ok = True
        if rank == 0:
            fo = np.frombuffer(full[4].cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE)
            res["flags"] = np.bincount(fo["eflag"], minlength=6).tolist()
            for r in range(world):
                if scaling == "weak":
                    rf, rc = r * total, total
                else:
                    lo, hi = sharding.shard_range(total, world, r)
                    rf, rc = lo, hi - lo
                n = min(rc, 256 if wl.nz < 2000 else 32)
                # the lane kernel serves large batches only (a small prefix would take the
                # CTA kernel, whose factors differ by rounding): re-solve the whole range
                if "mpc-lane" in solver.path:
                    n = rc
                if n == 0:
                    continue
                dd = wl.generate(fb.problems, n, rf, threads)
                ddv = {k: torch.from_numpy(a).to(dev) for k, a in dd.items()}
                z1, l1, v1 = f64(n * nz), f64(n * nl), f64(n * nv)
                o1, y1 = solver.solve_batch(ddv, z1, l1, v1) if n <= max(B, 1) else (None, None)
                torch.cuda.synchronize()
                if o1 is None:
                    continue
                for got, ref, w in ((full[0], z1, nz), (full[1], l1, nl), (full[2], v1, nv),
                                    (full[3], y1, nv)):
                    ok = ok and bool(torch.equal(got[rf * w:(rf + n) * w], ref))
                a = fo[rf:rf + n].copy()
                b = np.frombuffer(o1.cpu().numpy().tobytes(), dtype=fb.OUT_DTYPE).copy()
                a["solve_time"] = b["solve_time"] = 0
                ok = ok and a.tobytes() == b.tobytes()
        res["gather_identical"] = ok if rank == 0 else None

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region
    # (z, l, v are in/out: every timed step gets its own zeroed cold-start buffers,
    # prepared before the clock starts -- clearing host memory is not part of a solve)
    if with_e2e and B:
        Ke = max(1, min(K, 3))
        yh = pin(B * nv)
        with on_cpus(local_cpus):
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
        h2d = sum(a.nbytes for a in d_host.values()) + zh.nbytes + lh.nbytes + vh.nbytes
        d2h = zh.nbytes + lh.nbytes + vh.nbytes + yh.nbytes + oh.nbytes
        assert (oh["eflag"] == o["eflag"]).all(), "host and device paths disagree"
        res["e2e"] = {"value": global_batch * Ke / float(te.item()), "unit": UNIT,
                      "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}
        # the host-side ceiling of e2e: all ranks copy their pinned inputs to their GPUs at
        # the same time, nothing else running (aggregate GB/s over the box's PCIe / host
        # memory system); e2e cannot exceed global_batch / (h2d bytes / this rate)
        big = max(d_host, key=lambda k_: d_host[k_].nbytes)
        src_t = torch.from_numpy(d_host[big])
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            d_dev[big].copy_(src_t, non_blocking=True)
        barrier()
        tb = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        agg = world * 2 * d_host[big].nbytes / float(tb.item()) / 1e9
        res["e2e"]["host_h2d_aggregate_gbs"] = agg
        res["e2e"]["pinned_buffers"] = (
            f"first touched on the {len(local_cpus)} CPUs NVML names as local to the GPU"
            if local_cpus else "default placement (GPU-local CPUs unknown or all of them)")
        res["e2e"]["host_ceiling"] = global_batch / (h2d / (agg / world * 1e9))
        if wl.kind == "mpc":
            # the same solve with ONE copy of the stage data (fbstab_mpc_batch_solve_shared):
            # what a caller who knows its plants are identical ships over PCIe
            one = {k: (a if k == "x0" else a[:solver.field_sizes[k]]) for k, a in d_host.items()}

            def step_shared(i):
                zh_, lh_, vh_ = warm[i]
                zh_[:] = 0
                lh_[:] = 0
                vh_[:] = 0
                solver.solve_batch_shared(one, zh_, lh_, vh_, y=yh, out=oh,
                                          stream=stream.cuda_stream)

            step_shared(Ke)
            barrier()
            for i in range(Ke):
                warm[i][0][:] = 0
                warm[i][1][:] = 0
                warm[i][2][:] = 0
            t0 = time.perf_counter()
            for i in range(Ke):
                zh_, lh_, vh_ = warm[i]
                solver.solve_batch_shared(one, zh_, lh_, vh_, y=yh, out=oh,
                                          stream=stream.cuda_stream)
            barrier()
            ts = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            res["e2e_shared_stage_data"] = {
                "value": global_batch * Ke / float(ts.item()), "unit": UNIT,
                "h2d_bytes_per_step": int(sum(a.nbytes for a in one.values()) + zh.nbytes +
                                          lh.nbytes + vh.nbytes),
                "d2h_bytes_per_step": int(d2h),
                "same_flags_as_wire_format": bool((oh["eflag"] == o["eflag"]).all())}
    else:
        res["e2e"] = None

    if rank != 0:
        return res

    # ---- CPU baseline + parity on a bounded prefix of rank 0's instances
    W_counters = (o["newton_iters"].astype(np.float64), o["prox_iters"].astype(np.float64),
                  o["ls_backtracks"].astype(np.float64))
    work_model = ("SURVEY.md 8(a) W_flops with this run's per-instance newton / prox / "
                  "backtrack counters")
    cpu = None
    if with_cpu and B:
        cores = host_threads()
        count = max(1, min(B, wl.oracle_sample))
        cv, cdt, co, cz = cpu_reference_run(wl, count, cores, first=first)
        zg = z[:count * nz].cpu().numpy()
        par = parity_block(wl, o[:count], zg, co, cz, count)
        par["trajectory_floor_fma"] = trajectory_floor(wl)
        cpu = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {count} of the {B} instances of rank 0's shard, "
                         f"{cdt:.1f} s on {cores} host threads (one solver per thread)",
               "note": "restated reference (oracle/): Eigen is not in this image"}
        cpu.update(par)
        cpu["oracle_vs_reference_sources"] = reference_sources_check(wl, count, cores, first, co)
        # the oracle's counters where it was run: wasted GPU work earns no credit
        for k, f in enumerate(("newton_iters", "prox_iters", "ls_backtracks")):
            W_counters[k][:count] = co[f]
        work_model = (f"SURVEY.md 8(a) W_flops: the oracle's newton / prox / backtrack counters "
                      f"on the {count} sampled instances, this run's on the rest")
    res["cpu_baseline"] = cpu

    # ---- roofline of the dominant kernel (the persistent solve kernel)
    dfma, dmma = env["peaks"]
    W_total = float(wl.flops(*W_counters).sum())
    B_total = float(wl.bytes_per_solve()) * B
    ach_tf = W_total / (kernel_ms * 1e-3) / 1e12
    ach_gbs = B_total / (kernel_ms * 1e-3) / 1e9
    tensor_bound = wl.name == "5"  # A'GammaA + Cholesky updates run as DMMA
    peak = dmma if tensor_bound else dfma
    traffic = measured_traffic(wl.name, B)
    res["roofline"] = {
        "bound": "tensor" if tensor_bound else "fp64", "achieved": ach_tf, "peak": peak,
        "unit": "TFLOP/s", "frac": ach_tf / peak, "traffic": traffic,
        "traffic_source": ("profiles/traffic.json (ncu dram bytes of this build, csrc "
                           f"{csrc_hash()})" if traffic is not None else
                           "null: no ncu capture of this build (profiles/traffic.json is stale)"),
        "peak_source": f"measured in this run (fbstab_fp64_peak): DFMA loop {dfma:.1f}, "
                       f"FP64 mma.sync (DMMA) loop {dmma:.1f} TFLOP/s",
        "kernel": solver.path, "kernel_ms": kernel_ms,
        "algorithmic_flops_per_launch": W_total,
        "algorithmic_bytes_per_launch": B_total,
        "work_model": work_model,
        "hbm": {"achieved": ach_gbs, "peak": env["hbm_peak"], "unit": "GB/s",
                "frac": ach_gbs / env["hbm_peak"], "peak_source": env["hbm_src"]},
    }
    return res


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--per-config", default="all",
                    help="'all', 'none' or a comma-separated list of configs to add")
    ap.add_argument("--no-cpu", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    wl = Workload(args.config)
    scaling = args.scaling or wl.scaling
    if args.per_config == "all":
        extra = [c for c in PER_CONFIG_ORDER if c != args.config] + ["1"]
    elif args.per_config == "none":
        extra = []
    else:
        extra = [c for c in args.per_config.split(",") if c and c != args.config]
    if args.impl == "reference":
        run_reference(args, wl, scaling, extra)
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
    mg = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        mg = sharding.MultiGpu(rank, world, local_rank, dist=dist, torch=torch)
    hbm_peak, hbm_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak, hbm_src = float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        pass
    env = {"torch": torch, "dist": dist, "fb": fb, "mg": mg, "rank": rank, "world": world,
           "local_rank": local_rank, "dev": dev, "peaks": fb.capi.fp64_peak(local_rank),
           "local_cpus": gpu_local_cpus(torch, local_rank),
           "hbm_peak": hbm_peak, "hbm_src": hbm_src}
    W = max(args.warmup, 3)
    K = args.steps
    with_cpu = not args.no_cpu and world == 1

    t0 = time.perf_counter()
    head = measure(env, wl, K, W, scaling, with_cpu)
    _log(f"config {wl.name}: {head['value']:.4g} solves/s in {time.perf_counter() - t0:.1f} s")
    per = {}
    for name in extra:
        t0 = time.perf_counter()
        if name == "1":
            if rank == 0 and with_cpu:
                c1 = cpu_latency_config1()
                d = c1.pop("data")
                s1 = fb.FBstabDense(50, 10, 100, max_batch=1, device=local_rank)
                lat = []
                for _ in range(21):
                    z1, l1, v1 = np.zeros(50), np.zeros(10), np.zeros(100)
                    t1 = time.perf_counter()
                    o1, _ = s1.solve_batch(d, z1, l1, v1)
                    lat.append(time.perf_counter() - t1)
                c1["b200_single_instance_latency_ms_median"] = 1e3 * float(np.median(lat))
                c1["b200_same_trajectory"] = bool(
                    int(o1["newton_iters"][0]) == c1["newton_iters"] and
                    int(o1["eflag"][0]) == c1["exit_flag"])
                c1["note"] = ("a single QP cannot fill a GPU: this row is the reference's own "
                              "CPU-runnable case, reported as latency, not a throughput claim")
                per["1"] = c1
            continue
        w2 = Workload(name)
        r = measure(env, w2, max(1, min(K, 2)), 3, args.scaling or w2.scaling, with_cpu)
        if rank == 0:
            entry = {"workload": w2.label, "scaling": args.scaling or w2.scaling,
                     "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
                     "steps": max(1, min(K, 2)), "warmup": 3, "instances_per_gpu": r["B"],
                     "global_batch": r["global_batch"], "exit_flags": r["flags"],
                     "newton_iters_mean": r["newton_mean"], "path": r["path"],
                     "e2e": r["e2e"], "roofline": r.get("roofline"),
                     "cpu_baseline": r.get("cpu_baseline"), "gpu_launches": r["launches"],
                     "clocks": r["clocks"]}
            if "e2e_shared_stage_data" in r:
                entry["e2e_shared_stage_data"] = r["e2e_shared_stage_data"]
            if "gather_identical" in r:
                entry["gather_identical"] = r["gather_identical"]
            per[name] = entry
        _log(f"config {name}: {r['value']:.4g} solves/s in {time.perf_counter() - t0:.1f} s "
             f"(data generation {r['gen_s']:.1f} s)")

    if rank == 0:
        cfg = wl.static_config(world, scaling, extra)
        line = {
            "metric": wl.metric, "value": head["value"], "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "results": {"instances_per_gpu": head["B"], "global_batch": head["global_batch"],
                        "l2": f"inputs ({head['input_mb']:.0f} MB per GPU) exceed the 126 MB L2; "
                              "no flush needed",
                        "exit_flags": head["flags"], "path": head["path"],
                        "newton_iters_mean": head["newton_mean"]},
            "clocks": head["clocks"], "e2e": head["e2e"], "gpu_launches": head["launches"],
            "roofline": head.get("roofline"), "cpu_baseline": head.get("cpu_baseline"),
            "per_config": per,
        }
        if "gather_identical" in head:
            line["results"]["gather_identical"] = head["gather_identical"]
        if world > 1:
            line["results"]["gather"] = (
                "NCCL gather of every step's results to rank 0 inside the timed region, on a "
                "high-priority stream: it overlaps the next step's solve (double-buffered "
                "result arrays); the region ends after the last gather")
        _emit(line)
    if mg is not None:
        mg.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
