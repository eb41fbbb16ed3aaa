#!/usr/bin/env python3
"""bench.py -- headline benchmark of the ERI hot path.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

A STEP is one Schwarz-screened direct J/K build (every surviving shell quartet evaluated
once, digested into J and K, one all-reduce for N>1) of the workload below on synthetic
input.  metric = ERI shell-quartets/s (BASELINE.json).  `value` is measured with D already
resident in HBM; `e2e` goes through the reference-facing call with HOST buffers
(JK_direct(J, K, basis, D): H2D of D, compute, D2H of J and K inside the timed region).

The reference arm (--impl reference) times the reference's own CPU implementation of the
path (libpyquante2's coulomb_repulsion compiled unmodified into oracle/_ref, driven by the
restated basis.rs loop nest, OpenMP over all host cores) on a bounded seeded sample of the
same workload's surviving shell quartets.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_waters, basis, tau, BASELINE.json config it realises)
    "h2o96_631g": (96, "6-31G", 1e-10, "configs[4]: synthetic (H2O)_96 6-31G, N=1248, Schwarz-screened direct J/K"),
    "h2o96_sto3g": (96, "STO-3G", 1e-10, "configs[4]: synthetic (H2O)_96 STO-3G, N=672, Schwarz-screened direct J/K"),
    "h2o32_631g": (32, "6-31G", 1e-10, "configs[3]: synthetic (H2O)_32 6-31G, N=416, direct J/K"),
    "h2o32_631gs": (32, "6-31G*", 1e-10, "configs[3]: synthetic (H2O)_32 6-31G*, N=608 (s/p/d), direct J/K"),
    "h2o10_sto3g": (10, "STO-3G", 0.0, "configs[2]: synthetic (H2O)_10 STO-3G, N=70, ERI + J/K"),
    # dense-tensor mode only: the largest clusters whose N^4 tensor is a sensible share of HBM
    "h2o12_631gs": (12, "6-31G*", 0.0, "dense tensor: synthetic (H2O)_12 6-31G*, N=228 (s/p/d), 21.6 GB"),
    "h2o20_631g": (20, "6-31G", 0.0, "dense tensor: synthetic (H2O)_20 6-31G, N=260, 36.6 GB"),
}
DEFAULT_WORKLOAD = "h2o96_631g"
METRIC = "ERI shell-quartets/sec (Schwarz-screened direct J/K build)"
UNIT = "shell-quartets/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--boys", default="reference", choices=["reference", "exact"],
                    help="reference = libpyquante2 Fgamma (1e-12 parity); exact = converged Boys")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline budget")
    ap.add_argument("--mode", default="jk", choices=["jk", "tensor"],
                    help="jk = Schwarz-screened direct J/K build (the headline); tensor = build_I + "
                         "JK_inmem on the dense N^4 tensor (HBM-bound; N=1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled oracle parity block")
    ap.add_argument("--parity-elements", type=int, default=8)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------
# workload + sampling of surviving shell quartets (shared by both arms)
# ---------------------------------------------------------------------------------------
def load_geometry():
    """rchem_b200/geometry.py loaded BY PATH: the reference arm must not import the package
    (importing it maps librchem_b200.so into the process)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "_rchem_geometry", os.path.join(ROOT, "rchem_b200", "geometry.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_workload(name):
    geo = load_geometry()
    n_waters, basis, tau, desc = WORKLOADS[name]
    z, x = geo.water_cluster(n_waters)
    return z, x, basis, tau, desc


def sample_shell_quartets(shell_l, shell_first, sa, sb, Q, tau, count, seed=20261017):
    """Uniform sample of the surviving canonical shell quartets (pair p >= pair q in kernel
    order, Q_p*Q_q >= tau) as FUNCTION quartets: returns (function_quartets[int32, n x 4],
    n_shell_quartets).  Every Cartesian component of a sampled shell quartet is included.
    Shells reported with l = -1 are fused sp shells (s, px, py, pz); such a quartet counts as the
    segmented (s/p/d) shell quartets it covers -- the unit of the metric."""
    rng = np.random.default_rng(seed)
    npair = len(Q)
    ncart = lambda l: 4 if l == -1 else (l + 1) * (l + 2) // 2
    nvar = lambda l: 2 if l == -1 else 1
    out, got = [], 0
    while got < count:
        p = rng.integers(0, npair, size=4 * count)
        q = rng.integers(0, npair, size=4 * count)
        keep = (q <= p) & (Q[p] * Q[q] >= tau)
        for pp, qq in zip(p[keep], q[keep]):
            fa = [shell_first[sa[pp]] + i for i in range(ncart(shell_l[sa[pp]]))]
            fb = [shell_first[sb[pp]] + i for i in range(ncart(shell_l[sb[pp]]))]
            fc = [shell_first[sa[qq]] + i for i in range(ncart(shell_l[sa[qq]]))]
            fd = [shell_first[sb[qq]] + i for i in range(ncart(shell_l[sb[qq]]))]
            out.extend((a, b, c, d) for a in fa for b in fb for c in fc for d in fd)
            got += (nvar(shell_l[sa[pp]]) * nvar(shell_l[sb[pp]]) * nvar(shell_l[sa[qq]])
                    * nvar(shell_l[sb[qq]]))
            if got >= count:
                break
    return np.array(out, dtype=np.int32), got


def cpu_time_sample(orc, obasis, fq, n_shell, budget_s):
    """Times the oracle (reference kernel when oracle/_ref is present) on the sampled
    function quartets, repeating the whole sample until ~budget_s of wall time is used;
    returns (shell_quartets/s, shell quartets evaluated, seconds)."""
    t0 = time.perf_counter()
    orc.eval_quartets(obasis, fq, want_values=False)
    dt = max(time.perf_counter() - t0, 1e-6)
    repeats = int(max(1, min(1000, budget_s / dt)))
    t0 = time.perf_counter()
    for _ in range(repeats):
        orc.eval_quartets(obasis, fq, want_values=False)
    dt = time.perf_counter() - t0
    return n_shell * repeats / dt, n_shell * repeats, dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks"}

    def __init__(self, index):
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc

    z, x, basis_name, tau, desc = make_workload(args.workload)
    obasis = orc.make_basis(z, x, basis_name)
    kind = "port"
    if orc.ref_lib() is not None:
        orc.use_reference_kernel(True)
        kind = "reference"
    orc.set_num_threads(host_cores())
    # pair list + Schwarz bounds on the CPU (exact Boys, like the product), by the oracle
    shell_l, shell_first = [], []
    i = 0
    while i < obasis.n:
        l = int(obasis.powers[i].sum())
        shell_l.append(l)
        shell_first.append(i)
        i += (l + 1) * (l + 2) // 2
    shell_l, shell_first = np.array(shell_l), np.array(shell_first)
    ns = len(shell_l)
    ii, jj = np.tril_indices(ns)
    swap = shell_l[ii] < shell_l[jj]
    sa, sb = np.where(swap, jj, ii), np.where(swap, ii, jj)
    ncart = lambda l: (l + 1) * (l + 2) // 2
    diag = []
    owner = []
    for p in range(len(sa)):
        fa = [shell_first[sa[p]] + k for k in range(ncart(shell_l[sa[p]]))]
        fb = [shell_first[sb[p]] + k for k in range(ncart(shell_l[sb[p]]))]
        for a in fa:
            for b in fb:
                diag.append((a, b, a, b))
                owner.append(p)
    vals = orc.eval_quartets(obasis, np.array(diag, dtype=np.int32), orc.BOYS_EXACT)
    Q = np.zeros(len(sa))
    np.maximum.at(Q, np.array(owner), np.abs(vals))
    Q = np.sqrt(Q)
    # torchrun exports OMP_NUM_THREADS=1 to its workers: ask for every core explicitly and report
    # the thread count OpenMP actually runs with
    cores = orc.set_num_threads(host_cores())
    per_step = max(2.0, min(args.cpu_seconds, 60.0 / max(1, args.steps + args.warmup)))
    fq, n_shell = sample_shell_quartets(shell_l, shell_first, sa, sb, Q, tau, 20000)
    rates, times = [], []
    for s in range(args.warmup + args.steps):
        rate, used, dt = cpu_time_sample(orc, obasis, fq, n_shell, per_step)
        if s >= args.warmup:
            rates.append(rate)
            times.append(dt)
    value = float(np.mean(rates))
    sample = (f"{n_shell} seeded random surviving shell quartets of the workload (all Cartesian "
              f"components, all primitives) evaluated {used // n_shell}x per step; libpyquante2 "
              f"coulomb_repulsion via basis.rs loop order, OpenMP x{cores} threads "
              f"(set explicitly; {host_cores()} cores visible)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "basis": basis_name, "schwarz_tau": tau,
                   "nbf": int(obasis.n), "boys": "reference (libpyquante2 Fgamma)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def profile_counters(workload, boys, world):
    """ncu-derived per-step counters of a workload (DRAM bytes, FP64 warp instructions), from
    the committed summary profiles/counters.json (written by tools/counters_from_ncu.py from an
    ncu launch list of this very command); None when that configuration was not profiled."""
    path = os.path.join(ROOT, "profiles", "counters.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(f"{workload}/{boys}/{world}")
    except (OSError, ValueError):
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist

    import rchem_b200 as rc
    from rchem_b200 import geometry as geo
    from rchem_b200 import parallel

    if rc.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (rchem_b200 has no CPU path)")
    rank, world, local = parallel.init_distributed("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    z, x, basis_name, tau, desc = make_workload(args.workload)
    basis = rc.Basis.new(z, x, basis_name)
    basis.set_device(local)
    basis.set_schwarz_tau(tau)
    boys_mode = rc.BOYS_REFERENCE if args.boys == "reference" else rc.BOYS_EXACT
    basis.set_boys(boys_mode)
    n = basis.nbf
    D_np = geo.synthetic_density(n)
    D_host = torch.from_numpy(D_np).pin_memory()
    D_dev = D_host.to(dev)
    JK_dev = torch.zeros((2, n, n), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    basis.set_stream(stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        flush.fill_(1.0)
        parallel.jk_direct_distributed(basis, D_dev, JK_dev, rank, world)

    J_host = torch.empty((n, n), dtype=torch.float64).pin_memory()
    K_host = torch.empty((n, n), dtype=torch.float64).pin_memory()
    JK_host = torch.empty((2, n, n), dtype=torch.float64).pin_memory()
    D_in = torch.empty((n, n), dtype=torch.float64, device=dev)

    # N > 1, end to end: the reference-facing call itself drives all N GPUs.  Rank 0 owns one
    # more handle with RCHEM_OPT_NGPUS = N (single process, devices 0..N-1: one H2D of D, peer
    # copies over NVLink, peer-to-peer reduction of [J|K] on device 0, one D2H); the other ranks
    # only flush their GPU's L2 and wait on a CPU (gloo) barrier, so nothing of theirs runs on the
    # GPUs while rank 0's call is timed.
    multi, cpu_group = None, None
    if world > 1:
        cpu_group = dist.new_group(backend="gloo")
        if rank == 0 and rc.device_count() >= world:
            multi = rc.Basis.new(z, x, basis_name)
            multi.set_device(0)
            multi.set_schwarz_tau(tau)
            multi.set_boys(boys_mode)
            multi.set_gpus(world)
        flag = torch.tensor([1 if (rank != 0 or multi is not None) else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=cpu_group)
        use_multi = bool(flag.item())
    e2e_times = []

    def step_e2e():
        flush.fill_(1.0)
        torch.cuda.synchronize()
        if world == 1:
            # the reference-facing call: JK_direct(&mut J, &mut K, &basis, &D) with host buffers
            t0 = time.perf_counter()
            rc.JK_direct(J_host.numpy(), K_host.numpy(), basis, D_host.numpy())
            e2e_times.append(time.perf_counter() - t0)
        elif use_multi:
            dist.barrier(group=cpu_group)
            if rank == 0:
                t0 = time.perf_counter()
                rc.JK_direct(J_host.numpy(), K_host.numpy(), multi, D_host.numpy())
                e2e_times.append(time.perf_counter() - t0)
            dist.barrier(group=cpu_group)
        else:
            # (fewer devices visible to rank 0 than ranks: one process per GPU + NCCL)
            dist.barrier()
            t0 = time.perf_counter()
            if rank == 0:
                D_in.copy_(D_host, non_blocking=True)
            dist.broadcast(D_in, src=0)
            parallel.jk_direct_distributed(basis, D_in, JK_dev, rank, world)
            if rank == 0:
                JK_host.copy_(JK_dev, non_blocking=True)
            torch.cuda.synchronize()
            e2e_times.append(time.perf_counter() - t0)

    # ---- first call: builds pair data, Schwarz bounds, Boys tables and the task tables ----------
    barrier()
    t0 = time.perf_counter()
    step_resident()
    barrier()
    first_call_ms = 1e3 * (time.perf_counter() - t0)
    setup_ms = basis.stats()["setup_ms"]
    for _ in range(max(args.warmup - 1, 0)):
        step_resident()
    barrier()

    # ---- timed: device-resident -------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    kernel_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
        if rank == 0 and world == 1:
            kernel_ms.append(basis.stats()["kernel_ms"])  # syncs on the library's end event
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    if world == 1 and not kernel_ms:
        kernel_ms = [basis.stats()["kernel_ms"]]
    # subtract nothing: the L2 flush (~0.1 ms) is part of the step
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    # checksum of the result every rank holds after the all-reduce (SCALE: N = 2/4/8 must
    # reproduce N = 1 to rounding)
    J_res, K_res = JK_dev[0].clone(), JK_dev[1].clone()
    checksum = {"sum_J": float(J_res.sum().item()), "sum_K": float(K_res.sum().item()),
                "norm_J": float(torch.linalg.norm(J_res).item()),
                "norm_K": float(torch.linalg.norm(K_res).item()),
                "trace_JD": float((J_res * D_dev).sum().item()),
                "trace_KD": float((K_res * D_dev).sum().item())}

    # ---- timed: end to end through the host-buffer API ---------------------------------------
    # (each step's timed region is exactly the user-facing call; L2 flushes in between)
    step_e2e()
    del e2e_times[:]
    for _ in range(args.steps):
        step_e2e()
    barrier()
    t = torch.tensor([sum(e2e_times) if e2e_times else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    if world == 1 or use_multi:
        J_e2e, K_e2e = J_host.numpy().copy(), K_host.numpy().copy()
    else:
        J_e2e, K_e2e = JK_host[0].numpy().copy(), JK_host[1].numpy().copy()
    # ---- the other Boys flavour, same workload, device-resident (reported alongside) ----------
    other = rc.BOYS_EXACT if args.boys == "reference" else rc.BOYS_REFERENCE
    basis.set_boys(other)
    step_resident()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    other_ms = float(t.item()) / args.steps
    other_kernel_ms = basis.stats()["kernel_ms"]
    basis.set_boys(boys_mode)
    clocks = sampler.stop()

    # ---- whole-job counts (exact per rank, summed) -------------------------------------------------
    st = basis.stats()
    counts = torch.tensor([st["shell_quartets"], st["prim_quartets"], st["integrals"],
                           st["model_flops"], st["launches"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    sq, pq, ints, flops, launches = [float(v) for v in counts.tolist()]
    basis.use_own_stream()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peak_best, peak_avg = rc.fp64_peak(local, 10)
    ms_per_step = ms_total / args.steps
    value = sq / (ms_per_step * 1e-3)
    e2e_value = sq / (e2e_s / args.steps)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "basis": basis_name, "schwarz_tau": tau, "nbf": n,
                   "boys": "reference (libpyquante2 Fgamma)" if args.boys == "reference"
                   else "exact (tabulated Taylor; <=2e-8 from libpyquante2)",
                   "shell_quartets_per_step": sq, "shell_quartets_unscreened": st["shell_quartets_all"],
                   "primitive_quartets_per_step": pq, "integrals_per_step": ints,
                   "l2_flush": "256 MiB fill between steps",
                   "parallelism": f"quartet blocks round-robin over {world} GPU(s) + 1 all-reduce of [J|K]"},
        "integrals_per_s": ints / (ms_per_step * 1e-3),
        "primitive_quartets_per_s": pq / (ms_per_step * 1e-3),
        "jk_build_ms": ms_per_step,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * n * n,
                "d2h_bytes_per_step": 16 * n * n,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "rchem_jk_direct (host buffers)" if world == 1 else
                       (f"rchem_jk_direct (host buffers) with RCHEM_OPT_NGPUS={world}: ONE call of one "
                        "process drives all GPUs (H2D of D once, NVLink peer copies, peer-to-peer "
                        "reduction of [J|K] on device 0, one D2H)" if use_multi else
                        "rank 0: pinned D -> H2D, NCCL broadcast of D; every rank: "
                        "rchem_jk_direct_device; all-reduce of [J|K]; rank 0: D2H")},
        "gpu_launches": int(launches * args.steps),
        "clocks": clocks,
        "setup": {"first_call_ms": first_call_ms, "library_setup_ms": setup_ms,
                  "note": "one-off per basis, outside the timed steps: shell-pair batches, Schwarz "
                          "bounds (GPU), Boys tables, implicit quartet-list tables; first_call_ms = "
                          "first J/K build including all of it"},
        "checksum": checksum,
    }
    # ---- roofline: FP64 pipe --------------------------------------------------------------------
    if world == 1:
        k_ms = float(np.mean(kernel_ms))
        achieved = flops / (k_ms * 1e-3) / 1e12
    else:
        k_ms = ms_per_step
        achieved = flops / (k_ms * 1e-3) / 1e12 / world
    other_name = "exact_boys" if args.boys == "reference" else "reference_boys"
    line[other_name] = {
        "note": "same workload with the other Boys flavour (SURVEY H1: exact = converged tabulated "
                "Boys, <=2e-8 from libpyquante2; reference = libpyquante2 Fgamma, 1e-12 parity)",
        "value": sq / (other_ms * 1e-3), "unit": UNIT, "ms_per_step": other_ms,
        "roofline_frac": (flops / ((other_kernel_ms if world == 1 else other_ms) * 1e-3) / 1e12
                          / (world if world > 1 else 1)) / peak_avg,
    }
    prof = profile_counters(args.workload, args.boys, world)
    line["roofline"] = {
        "bound": "fp64", "achieved": achieved, "peak": peak_avg, "unit": "TFLOP/s",
        "frac": achieved / peak_avg,
        "traffic": prof["dram_bytes_per_step"] if prof else None,
        "traffic_source": prof["source"] if prof else None,
        "peak_source": "builder-measured in this run: DFMA microbenchmark rchem_fp64_peak (avg of 10; "
                       f"best {peak_best:.2f}); MEASURED_PEAKS.json has no FP64 entry",
        "kernel": "eri_jk_block_kernel<la,lb,lc,ld,boys> + eri_jk_light_multi_kernel<..> (all class "
                  "instantiations of one step; the (ps|ss) block kernel is the largest share)",
        "kernel_ms_per_step": k_ms,
        "algorithmic_flops_per_step": flops,
        "flop_model": "SURVEY 8(d): sum over surviving quartets of K2_bra*K2_ket*P(class)+H(class)",
        "per_gpu": world > 1,
    }
    if prof and prof.get("fp64_warp_inst_per_step"):
        # instruction-based view: FP64 warp instructions actually executed (ncu) against the
        # FP64 pipe's issue rate (2 warp instructions per SM and clock = 64 lanes) over the LIVE
        # step time -- independent of the flop model
        sm_hz = 1e6 * (clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965.0)
        line["roofline"]["fp64_pipe_active"] = (
            prof["fp64_warp_inst_per_step"] / world / (148 * 2.0 * sm_hz * k_ms * 1e-3))
        line["roofline"]["fp64_pipe_active_source"] = (
            "FP64 warp instructions per step from " + prof["source"] + " over the live kernel time")
    # ---- parity: sampled J/K elements of THIS run's result against oracle rows -----------------
    if not args.no_parity:
        from oracle import oracle as orc
        from oracle import parity

        if orc.ref_lib() is not None:
            orc.use_reference_kernel(True)
        orc.set_num_threads(host_cores())
        obasis = orc.make_basis(z, x, basis_name)
        t0 = time.perf_counter()
        ej, ek, elements = parity.sampled_jk_errors(orc, obasis, basis, D_np, J_e2e, K_e2e, tau,
                                                    count=args.parity_elements)
        line["parity"] = {
            "max_abs_err": max(ej, ek), "max_abs_err_J": ej, "max_abs_err_K": ek,
            "tolerance": 1e-12, "elements": len(elements),
            "checked": "J[mu,nu], K[mu,nu] of the e2e (host-buffer) result vs rows rebuilt from "
                       "oracle integrals (libpyquante2 coulomb_repulsion when oracle/_ref is present) "
                       "under the same Schwarz list; outside the timed region",
            "boys": args.boys, "seconds": time.perf_counter() - t0}
    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc

        kind = "port"
        if orc.ref_lib() is not None:
            orc.use_reference_kernel(True)
            kind = "reference"
        threads = orc.set_num_threads(host_cores())
        obasis = orc.make_basis(z, x, basis_name)
        sa, sb, _, Q = basis.schwarz()
        l, first = basis.shells()
        fq, n_shell = sample_shell_quartets(l, first, sa, sb, Q, tau, 20000)
        rate, used, dt = cpu_time_sample(orc, obasis, fq, n_shell, args.cpu_seconds)
        line["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{n_shell} seeded random surviving shell quartets of the workload, all "
                      f"components and primitives, evaluated {used // n_shell}x ({dt:.1f} s); "
                      f"libpyquante2 coulomb_repulsion in basis.rs loop order, OpenMP x{threads}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_tensor(args):
    """Dense-tensor mode: a step = build_I (every canonical quartet evaluated, all 8 permutation
    images written: N^4 doubles) followed by JK_inmem (the tensor read once).  Both are bound by
    HBM: algorithmic bytes = 8 N^4 written, then 8 N^4 read."""
    import torch

    import rchem_b200 as rc
    from rchem_b200 import geometry as geo

    if rc.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (rchem_b200 has no CPU path)")
    if args.gpus != 1:
        raise SystemExit("--mode tensor is a one-GPU measurement")
    name = args.workload if args.workload != DEFAULT_WORKLOAD else "h2o10_sto3g"
    z, x, basis_name, tau, desc = make_workload(name)
    dev = torch.device("cuda", 0)
    basis = rc.Basis.new(z, x, basis_name)
    basis.set_boys(rc.BOYS_REFERENCE if args.boys == "reference" else rc.BOYS_EXACT)
    n = basis.nbf
    stream = torch.cuda.current_stream()
    basis.set_stream(stream.cuda_stream)
    I = torch.empty((n,) * 4, dtype=torch.float64, device=dev)
    D = torch.from_numpy(geo.synthetic_density(n)).to(dev)
    JK = torch.empty((2, n, n), dtype=torch.float64, device=dev)
    nbytes = 8.0 * n ** 4
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    hbm = 6552.3
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)"
    except (OSError, ValueError, KeyError):
        peak_src = "fallback 6552.3 GB/s (MEASURED_PEAKS.json absent)"

    def timed(fn, reps):
        ms = []
        for _ in range(reps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.mean(ms))

    build = lambda: basis.build_I_device(I.data_ptr())
    inmem = lambda: rc.jk_inmem_device(n, I.data_ptr(), D.data_ptr(), JK.data_ptr(), stream.cuda_stream)
    for _ in range(max(args.warmup, 1)):
        build()
        inmem()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    build_ms = timed(build, args.steps)
    st = basis.stats()
    inmem_ms = timed(inmem, args.steps)
    clocks = sampler.stop()
    # parity of the whole step at a size the oracle can do; otherwise sampled elements
    J_dev, K_dev = JK[0].cpu().numpy(), JK[1].cpu().numpy()
    parity = None
    if not args.no_parity:
        from oracle import oracle as orc
        from oracle import parity as par

        if orc.ref_lib() is not None:
            orc.use_reference_kernel(True)
        orc.set_num_threads(host_cores())
        ob = orc.make_basis(z, x, basis_name)
        seg = rc.Basis.new(z, x, basis_name)
        ej, ek, els = par.sampled_jk_errors(orc, ob, seg, D.cpu().numpy(), J_dev, K_dev, 0.0,
                                            count=args.parity_elements)
        rng = np.random.default_rng(5)
        q = rng.integers(0, n, size=(4000, 4)).astype(np.int32)
        vals = orc.eval_quartets(ob, q, orc.BOYS_REFERENCE if args.boys == "reference" else orc.BOYS_EXACT)
        got = I[q[:, 0], q[:, 1], q[:, 2], q[:, 3]].cpu().numpy()
        parity = {"max_abs_err_I": float(np.abs(got - vals).max()), "sampled_integrals": len(q),
                  "max_abs_err_J": ej, "max_abs_err_K": ek, "elements": len(els), "tolerance": 1e-12}
    basis.use_own_stream()
    step_ms = build_ms + inmem_ms
    line = {
        "metric": METRIC.replace("Schwarz-screened direct J/K build", "dense tensor build_I + JK_inmem"),
        "value": st["shell_quartets"] / (step_ms * 1e-3), "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "mode": "tensor", "basis": basis_name, "nbf": n,
                   "tensor_bytes": nbytes, "shell_quartets_per_step": st["shell_quartets"],
                   "l2_flush": "256 MiB fill before every timed launch group"},
        "build_I_ms": build_ms, "jk_inmem_ms": inmem_ms,
        "roofline": {"bound": "hbm", "achieved": nbytes / (build_ms * 1e-3) / 1e9, "peak": hbm,
                     "unit": "GB/s", "frac": nbytes / (build_ms * 1e-3) / 1e9 / hbm,
                     "traffic": None, "peak_source": peak_src,
                     "kernel": "build_I: eri_kernel<..,tensor> launches + tensor_fill_kernel",
                     "algorithmic_bytes": nbytes,
                     "note": "8 N^4 bytes written once; the quartet kernels are FP64-bound below N~100"},
        "roofline_jk_inmem": {"bound": "hbm", "achieved": nbytes / (inmem_ms * 1e-3) / 1e9, "peak": hbm,
                              "unit": "GB/s", "frac": nbytes / (inmem_ms * 1e-3) / 1e9 / hbm,
                              "kernel": "jk_inmem_kernel", "algorithmic_bytes": nbytes},
        "gpu_launches": int(st["launches"] * args.steps + args.steps),
        "clocks": clocks, "parity": parity,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "tensor":
        return run_tensor(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
