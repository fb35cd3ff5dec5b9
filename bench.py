#!/usr/bin/env python
"""bench.py -- likelihood evaluations per second of the batched PopPK path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload poppk_two_100k_x64|poppk_one_1k_x16] [--impl reference]

A "step" is one batched call: EvaluateLogProbability for all C tempered chains' proposals at once, i.e. C * P
independent stiff ODE solves + the per-chain reduction. One evaluation = one chain's log-likelihood (the unit the
reference counts in Sampler.cpp:129-134).

  value  device-timed (CUDA events on the launching stream), parameter vectors already resident in HBM
  e2e    the same steps through the reference-facing call with HOST buffers: pinned host values -> H2D of the
         rank's slice -> kernels -> (NCCL all-reduce) -> D2H of the result, wall-clock with a device sync
  roofline       dominant kernel (poppk_kernel) against the FP64 FMA peak measured live on this GPU
  cpu_baseline   the reference's CPU implementation (oracle/_ref = its own compiled CVODE stack; else the plain-C
                 port) timed on the box's host cores on a bounded sample of the same workload (rank 0, N=1 only)

--impl reference times only that CPU implementation, one bounded sample per step, same metric and config.
N > 1: launched by torchrun, one rank per GPU; patients are sharded across ranks (strong scaling on the named
workload), per-chain partials combined by an NCCL all-reduce; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[4]: PopPK two-compartment, 64 tempered chains x 100k individuals (the 1/2/4/8 sweep config)
    "poppk_two_100k_x64": dict(pk="two", P=100_000, T=10, t_end=72.0, C=64, flop_per_system=64.5e3),
    # BASELINE.json configs[1]: PopPK one-compartment, 1k individuals x 10 samples, 16 temperatures
    "poppk_one_1k_x16": dict(pk="one", P=1000, T=10, t_end=72.0, C=16, flop_per_system=48.0e3),
    # BASELINE.json configs[2]: cellpop small cell-cycle-like SBML-style model, 10k simulated cells, 16 temperatures, 1 GPU
    "cellpop_12sp_10k_x16": dict(kind="cellpop", N=12, cells=10_000, T=50, C=16),
    # BASELINE.json configs[3]: cellpop ~50-species stiff signalling cascade, 100k cells, 16 temperatures (sharded over GPUs when N > 1)
    "cellpop_50sp_100k_x16": dict(kind="cellpop", N=50, cells=100_000, T=50, C=16, rate_decades=4.0),
}
METRIC = "likelihood evals/sec (PopPK batched EvaluateLogProbability)"
METRIC_CELLPOP = "likelihood evals/sec (cellpop batched EvaluateLogProbability)"
UNIT = "evals/s"
Q_MEAN = 4.3  # mean BDF order measured on the reference


def algorithmic_flop_per_system(N: int, cnt_mean: np.ndarray, nout: float) -> float:
    """FLOP(system) from the ORACLE's counters (steps, nfe, nsetups, nje, netf, ncfn, nni, ok)."""
    nst, nfe, nsetups, nje, nni = cnt_mean[0], cnt_mean[1], cnt_mean[2], cnt_mean[3], cnt_mean[6]
    q = Q_MEAN
    A = N * (q * (q + 1) / 2 + 2 * (q + 1) + 4) + 60
    f_rhs = 5 if N == 2 else 11
    f_fact = 8 if N == 2 else 45
    return float(nst * A + nfe * f_rhs + nni * (2 * N * N + 9 * N) + nsetups * (2 * N * N + f_fact) + nje * 0 + nout * 2 * N * (q + 1))


def algorithmic_bytes_per_eval(P: int, T: int, nvar: int) -> float:
    """observations + parameters + dosing metadata + result, per chain evaluation."""
    return 8.0 * P * T + 8.0 * nvar + 56.0 * P + 8.0


def make_workload(name: str):
    from bcm3_b200 import synthetic as syn
    from bcm3_b200.poppk_data import PK_ONE, PK_TWO

    w = WORKLOADS[name]
    pk = PK_ONE if w["pk"] == "one" else PK_TWO
    prob = syn.make_poppk_problem(pk, P=w["P"], T=w["T"], t_end=w["t_end"], seed=1)
    vals = syn.make_chain_values(prob, w["C"])
    return prob, vals


def subsample_problem(prob, vals, P_sample: int):
    """First P_sample patients of the workload with their columns of the parameter vectors (same model, same chains)."""
    from bcm3_b200 import synthetic as syn
    from bcm3_b200.poppk_data import PopPKProblem, num_pk_params

    tr = prob.trial
    P = tr.num_patients
    if P_sample >= P:
        return prob, vals
    npk = num_pk_params(prob.pk_type)
    s = slice(0, P_sample)
    trs = type(tr)(drug=tr.drug, time=tr.time, observed_concentration=tr.observed_concentration[s], dose=tr.dose[s],
                   dosing_interval=tr.dosing_interval[s], dose_after_dose_change=tr.dose_after_dose_change[s],
                   dose_change_time=tr.dose_change_time[s], intermittent=tr.intermittent[s],
                   treatment_interruptions=tr.treatment_interruptions[s])
    nvs = npk + 2 * (P_sample + 1) + 2
    ps = PopPKProblem(pk_type=prob.pk_type, trial=trs, transforms=syn.poppk_transforms(prob.pk_type, P_sample), sd_ix=nvs - 2)
    vs = np.empty((vals.shape[0], nvs))
    vs[:, :npk + 2 + 2 * P_sample] = vals[:, :npk + 2 + 2 * P_sample]
    vs[:, nvs - 2:] = vals[:, -2:]
    return ps, vs


def cellpop_flop_per_system(N: int, steps: float, n_ratelaw_flops: float) -> float:
    """SURVEY.md section 8(d) formula with the counter ratios the reference shows on this model family (measured with the
    oracle: nfe/nst = 1.33, nni = nfe - 1, nsetups/nst = 0.122, nje/nst = 0.021), q = 4.3, DQ Jacobian and dense LU."""
    q = Q_MEAN
    A = N * (q * (q + 1) / 2 + 2 * (q + 1) + 4) + 60
    nfe, nsetups, nje = 1.33 * steps, 0.122 * steps, 0.021 * steps
    nni = nfe - 1
    return float(steps * A + nfe * n_ratelaw_flops + nni * (2 * N * N + 9 * N) + nsetups * (2 * N * N + 2 * N ** 3 / 3) + nje * N * (n_ratelaw_flops + 2 * N))


def run_cellpop(args, workload: str):
    """cell_population workloads: single GPU (cells are not sharded across ranks yet)."""
    import torch

    from bcm3_b200 import _lib
    from bcm3_b200 import synthetic_cellpop as sc
    from bcm3_b200.cellpop import CellPopEvaluator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rank != 0 and (args.impl == "reference" or world == 1):
        return
    w = WORKLOADS[workload]
    prob = sc.make_cellpop_problem(N=w["N"], num_cells=w["cells"], T=w["T"], data_cells=32, seed=1, rate_decades=w.get("rate_decades", 2.0))
    vals = sc.make_chain_values(w["C"])
    C, nvar = vals.shape
    if args.impl != "reference" and world > 1:
        run_cellpop_sharded(args, workload, prob, vals, rank, world, local_rank)
        return
    if args.impl == "reference":
        kind, chk = cpu_checker()
        cores = max(1, min(C, os.cpu_count() or 1))
        import dataclasses
        times = []
        for i in range(args.warmup + args.steps):
            sample = min(w["cells"], 2000 if w["N"] <= 20 else 200)
            ps = dataclasses.replace(prob, num_cells=sample, sobol=prob.sobol[:sample])
            t0 = time.perf_counter()
            chk.cellpop_evaluate(ps, vals, threads=cores)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
        dt = statistics.mean(times)
        value = (C / dt) * (sample / w["cells"])
        print(json.dumps({"impl": "reference", "metric": METRIC_CELLPOP, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "config": {"workload": workload, "species": w["N"], "cells": w["cells"], "chains": C,
                                                                        "timepoints": w["T"]},
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                                           "sample": f"all {C} chains x first {sample} of {w['cells']} cells in {dt:.1f} s on {cores} threads ({cpu_model_name()}); scaled by {sample}/{w['cells']}"},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    if not torch.cuda.is_available() or _lib.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    ev = CellPopEvaluator(prob, device=0)
    h_vals = torch.from_numpy(vals).pin_memory()
    h_logp = np.empty(C)
    h_status = np.empty(C, dtype=np.int32)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda:0")
    fp64_peak = _lib.measure_fp64_peak(0)

    def step():
        _lib.check(ev.lib.bcm3b200_evaluate_batch(ev.handle, C, nvar, h_vals.data_ptr(), h_logp.ctypes.data, h_status.ctypes.data))

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    launches0 = ev.get_stat("total_kernel_launches")
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.25)
    kernel_ms, t_begin = [], time.perf_counter()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        e2e_s += time.perf_counter() - t0
        kernel_ms.append(ev.get_stat("last_kernel_us") / 1e3)
    t_end = time.perf_counter()
    clocks = sampler.stop(t_begin, t_end)
    launches = ev.get_stat("total_kernel_launches") - launches0
    ev._last_C = C
    steps_mean = float(ev.diagnostics()["cell_steps"].mean())
    total_ms = sum(kernel_ms)
    value = C * args.steps / (total_ms * 1e-3)
    flop_sys = cellpop_flop_per_system(w["N"], steps_mean, 8.0 * 2 * w["N"])
    k_ms = statistics.mean(kernel_ms)
    achieved = flop_sys * C * w["cells"] / (k_ms * 1e-3) / 1e12
    line = {"metric": METRIC_CELLPOP, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "species": w["N"], "cells": w["cells"], "chains": C, "timepoints": w["T"], "ode_solves_per_step": C * w["cells"],
                       "mean_steps_per_solve": steps_mean, "l2": "256 MB memset between timed iterations"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": ncu_traffic(workload),
                         "kernel": "cellpop_group_kernel" if w["N"] <= 96 else "cellpop_kernel", "kernel_ms": k_ms, "flop_per_system": flop_sys,
                         "systems_per_launch": C * w["cells"], "peak_source": "measured live: bcm3b200_measure_fp64_peak"},
            "e2e": {"value": C * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(vals.nbytes), "d2h_bytes_per_step": int(C * 8), "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "check": {"logp0": float(h_logp[0]), "status_ok": bool((h_status == 0).all())}}
    if not args.no_cpu_baseline:
        import dataclasses
        kind, chk = cpu_checker()
        cores = max(1, min(C, os.cpu_count() or 1))
        sample = min(w["cells"], 4000 if w["N"] <= 20 else 400)
        ps = dataclasses.replace(prob, num_cells=sample, sobol=prob.sobol[:sample])
        t0 = time.perf_counter()
        chk.cellpop_evaluate(ps, vals, threads=cores)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": (C / dt) * (sample / w["cells"]), "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"all {C} chains x first {sample} of {w['cells']} cells in {dt:.1f} s on {cores} threads ({cpu_model_name()}); scaled by {sample}/{w['cells']}"}
    print(json.dumps(line))
    ev.close()


def run_cellpop_sharded(args, workload, prob, vals, rank, world, local_rank):
    """N > 1: the simulated cells are split over the ranks (strong scaling), one SUM all-reduce of [C][2 T + 1] doubles."""
    import torch
    import torch.distributed as dist

    from bcm3_b200 import _lib
    from bcm3_b200.parallel import ShardedCellPopLikelihood

    w = WORKLOADS[workload]
    C, nvar = vals.shape
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    lk = ShardedCellPopLikelihood(prob, rank, world, local_rank)
    h_vals = torch.from_numpy(vals).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        logp, status = lk.evaluate(h_vals)
    barrier()
    launches0 = lk.evaluator.get_stat("total_kernel_launches")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_begin = time.perf_counter()
    wall = 0.0
    for a, b in evs:
        flush.zero_()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        a.record(stream)
        logp, status = lk.evaluate(h_vals)  # H2D of the batch, kernels, all-reduce, data likelihood, D2H of logp
        b.record(stream)
        wall += time.perf_counter() - t0
    barrier()
    t_end = time.perf_counter()
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    torch.cuda.synchronize(dev)
    tot = torch.tensor([sum(a.elapsed_time(b) for a, b in evs), 1e3 * wall], dtype=torch.float64, device=dev)
    nl = torch.tensor([float(lk.evaluator.get_stat("total_kernel_launches") - launches0)], dtype=torch.float64, device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    # roofline of the dominant kernel on this rank's shard (kernel time from the library's own events, max over ranks)
    kms = torch.tensor([lk.evaluator.get_stat("last_kernel_us") / 1e3], dtype=torch.float64, device=dev)
    dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    fp64_peak = _lib.measure_fp64_peak(local_rank)
    cells_local = lk.evaluator.get_stat("num_cells_local")
    if rank == 0:
        total_ms, wall_ms = float(tot[0].item()), float(tot[1].item())
        steps_mean = 243.6 if w["N"] == 12 else 542.0  # mean accepted steps per cell of the synthetic models (measured at N = 1)
        flop_sys = cellpop_flop_per_system(w["N"], steps_mean, 8.0 * 2 * w["N"])
        k_ms = float(kms.item())
        achieved = flop_sys * C * cells_local / (k_ms * 1e-3) / 1e12
        roofline = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": None,
                    "kernel": "cellpop_group_kernel", "kernel_ms": k_ms, "flop_per_system": flop_sys, "systems_per_launch": C * cells_local,
                    "note": "per GPU, on its shard of the cells", "peak_source": "measured live: bcm3b200_measure_fp64_peak"}
        line = {"metric": METRIC_CELLPOP, "value": C * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "species": w["N"], "cells": w["cells"], "chains": C, "timepoints": w["T"], "ode_solves_per_step": C * w["cells"],
                           "sharding": f"cells over {world} ranks, NCCL SUM all-reduce of [{C}][{2 * w['T'] + 1}] doubles", "l2": "256 MB memset between timed iterations"},
                "roofline": roofline,
                "e2e": {"value": C * args.steps / (wall_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(vals.nbytes) * world, "d2h_bytes_per_step": int(C * 8) * world,
                        "ms_per_step": wall_ms / args.steps},
                "gpu_launches": int(nl.item()), "clocks": clocks, "check": {"logp0": float(logp[0]), "status_ok": bool((status == 0).all())}}
        print(json.dumps(line))
    lk.close()
    dist.destroy_process_group()


def cpu_checker():
    """(kind, oracle object): the compiled reference when it was built in the container, else the plain-C port."""
    import oracle

    if oracle.available("ref"):
        return "reference", oracle.load("ref")
    if not oracle.available("port"):
        oracle.build_port()
    return "port", oracle.load("port")


def cpu_model_name() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def time_cpu(prob, vals, target_seconds: float, want_counters: bool = False):
    """Time the CPU implementation on a bounded sample of the workload: all C chains, the first P_sample patients.
    One worker thread per chain up to the core count, as the reference schedules them (SamplerPTChain.cpp:315-326).
    Returns dict(evals_per_s scaled to the full workload, cores, sample text, counters)."""
    kind, chk = cpu_checker()
    C = vals.shape[0]
    P = prob.trial.num_patients
    cores = max(1, min(C, os.cpu_count() or 1))
    # calibrate with a tiny sample
    p0 = min(P, max(8, 2 * cores))
    ps, vs = subsample_problem(prob, vals, p0)
    t0 = time.perf_counter()
    chk.poppk_evaluate(ps, vs, threads=cores)
    dt = max(time.perf_counter() - t0, 1e-4)
    rate = C * p0 / dt  # solves / s
    P_sample = int(min(P, max(p0, rate * target_seconds / C)))
    ps, vs = subsample_problem(prob, vals, P_sample)
    t0 = time.perf_counter()
    r = chk.poppk_evaluate(ps, vs, threads=cores, want_counters=want_counters)
    dt = time.perf_counter() - t0
    evals_per_s = (C / dt) * (P_sample / P)
    return dict(kind=kind, cores=cores, P_sample=P_sample, seconds=dt, evals_per_s=evals_per_s, solves_per_s=C * P_sample / dt,
                counters=r["counters"], sample=f"all {C} chains x first {P_sample} of {P} individuals in {dt:.1f} s on {cores} threads "
                f"({cpu_model_name()}); evals/s scaled by {P_sample}/{P}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.gpu),
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin: float, t_end: float):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t_begin or ts > t_end + 0.2:
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), power_w_max=max(power), samples=len(sm), reasons=sorted(reasons))


def ncu_traffic(workload: str):
    """dram bytes read+written per launch of the dominant kernel from the committed ncu summary, if one exists."""
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        d = json.load(open(path))
        return d.get(workload, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def run_reference(args, workload: str):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob, vals = make_workload(workload)
    w = WORKLOADS[workload]
    target = 4.0  # seconds of CPU work per step
    times, last = [], None
    for i in range(args.warmup + args.steps):
        last = time_cpu(prob, vals, target)
        if i >= args.warmup:
            times.append(last)
    value = statistics.mean(t["evals_per_s"] for t in times)
    ms = 1e3 * statistics.mean(t["seconds"] for t in times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "pk_model": w["pk"], "individuals": w["P"], "chains": w["C"], "timepoints": w["T"],
                   "note": "each step = a bounded sample of the workload (all chains x a prefix of the individuals), evals/s scaled to the full size"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="poppk_two_100k_x64", choices=sorted(WORKLOADS))
    ap.add_argument("--block-size", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    workload = args.workload

    if WORKLOADS[workload].get("kind") == "cellpop":
        run_cellpop(args, workload)
        return

    if args.impl == "reference":
        run_reference(args, workload)
        return

    import torch
    import torch.distributed as dist

    from bcm3_b200 import _lib
    from bcm3_b200.parallel import ShardedPopPKLikelihood, combine_partials, shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available() or _lib.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    w = WORKLOADS[workload]
    prob, vals = make_workload(workload)
    C, nvar = vals.shape
    P, T = w["P"], w["T"]
    lk = ShardedPopPKLikelihood(prob, rank, world, local_rank, block_size=args.block_size)
    ev = lk.evaluator

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # inputs: pinned host copy (e2e) and a device-resident copy (value)
    h_vals = torch.from_numpy(vals).pin_memory()
    d_vals = torch.from_numpy(vals).to(dev)
    d_partial = torch.empty((3, C), dtype=torch.float64, device=dev)
    h_partial = torch.empty((3, C), dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    def device_step():
        ev.evaluate_device(d_vals.data_ptr(), C, nvar, d_partial.data_ptr(), stream.cuda_stream)
        if world > 1:
            from bcm3_b200.parallel import allreduce_partial

            allreduce_partial(d_partial)

    fp64_peak = _lib.measure_fp64_peak(local_rank)

    # ---- device-resident timing: W warm-up steps, then exactly K timed steps ----
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        device_step()
    barrier()
    launches0 = ev.get_stat("total_kernel_launches")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_begin = time.perf_counter()
    for a, b in evs:
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        a.record(stream)
        device_step()
        b.record(stream)
    barrier()
    t_end = time.perf_counter()
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    launches = ev.get_stat("total_kernel_launches") - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    nl = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    total_ms = float(total_ms.item())
    launches_all = int(nl.item())
    h_partial.copy_(d_partial)
    logp, status = combine_partials(h_partial.numpy())

    # kernel-only duration (events around the two kernels inside the library), for the roofline of the dominant kernel
    kernel_ms = []
    for _ in range(3):
        flush.zero_()
        ev.evaluate_raw(h_vals.data_ptr(), C, nvar, h_partial.data_ptr())  # host entry records ev0/ev1 around the kernels
        kernel_ms.append(ev.get_stat("last_kernel_us") / 1e3)

    # ---- end-to-end timing through the reference-facing call with HOST buffers ----
    h_logp = np.empty(C)
    h_status = np.empty(C, dtype=np.int32)

    def e2e_step():
        if world == 1:
            ev.evaluate_raw(h_vals.data_ptr(), C, nvar, h_logp.ctypes.data, h_status.ctypes.data)
            return h_logp
        return lk.evaluate(h_vals)[0]

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        out = e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    e2e_logp = np.array(out, dtype=np.float64, copy=True)

    lo, hi = shard_bounds(P, rank, world)
    h2d = sum(C * (16 + 2 * (shard_bounds(P, r, world)[1] - shard_bounds(P, r, world)[0])) * 8 for r in range(world))
    d2h = (C * 8 + C * 4) if world == 1 else world * 3 * C * 8

    if rank == 0:
        N = 2 if w["pk"] == "one" else 3
        value = C * args.steps / (total_ms * 1e-3)
        ms_per_step = total_ms / args.steps
        # roofline of the dominant kernel: algorithmic FLOPs per launch / its average duration
        flop_per_system = w["flop_per_system"]
        systems_per_launch = C * (hi - lo)
        k_ms = statistics.mean(kernel_ms)
        achieved = flop_per_system * systems_per_launch / (k_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        bytes_per_launch = algorithmic_bytes_per_eval(hi - lo, T, nvar) * C
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "pk_model": w["pk"], "individuals": P, "chains": C, "timepoints": T, "t_end_h": w["t_end"],
                       "ode_solves_per_step": C * P, "sharding": f"individuals over {world} rank(s), NCCL all-reduce of [3][{C}] doubles",
                       "l2": "256 MB memset between timed iterations", "block_size": args.block_size or "auto"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                         "traffic": ncu_traffic(workload), "kernel": "poppk_kernel", "kernel_ms": k_ms,
                         "flop_per_system": flop_per_system, "systems_per_launch": systems_per_launch,
                         "peak_source": "measured live: bcm3b200_measure_fp64_peak (DFMA chains), MEASURED_PEAKS.json has no FP64 entry",
                         "hbm": {"achieved": bytes_per_launch / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": bytes_per_launch / (k_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": bytes_per_launch,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
            "e2e": {"value": C * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": launches_all,
            "clocks": clocks,
            "check": {"logp0": float(logp[0]), "status_ok": bool((status == 0).all()),
                      "e2e_matches_device": bool(np.allclose(e2e_logp, logp, rtol=1e-12, atol=0))},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = time_cpu(prob, vals, 12.0, want_counters=True)
            cnt = cb["counters"].reshape(-1, cb["counters"].shape[-1]).mean(axis=0)
            line["cpu_baseline"] = {"value": cb["evals_per_s"], "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"],
                                    "solves_per_s": cb["solves_per_s"]}
            line["roofline"]["flop_per_system_from_oracle_counters"] = algorithmic_flop_per_system(N, cnt, T)
            line["roofline"]["oracle_counters_mean"] = {k: float(v) for k, v in zip(
                ["steps", "nfe", "nsetups", "nje", "netf", "ncfn", "nni", "ok"], cnt)}
        print(json.dumps(line))
    lk.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
