#!/usr/bin/env python
"""bench.py -- likelihood evaluations per second of the batched PopPK / cellpop paths (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload ...] [--impl reference] [--no-secondary]

The ONE JSON line is the headline workload (default: BASELINE config 5, PopPK two-compartment, 100 000 individuals x 64
chains); its "secondary" list carries the cell_population half of the metric, each entry a complete line of its own (value,
e2e, roofline with oracle-counted FLOPs, cpu_baseline, clocks): BASELINE config 3 (12 species, 10 000 cells, 16 chains; N = 1
only -- it is a one-GPU configuration) and config 4 (50 species, 100 000 cells, 16 chains, cells sharded over the --gpus N ranks).

A "step" is one batched call: EvaluateLogProbability for all C tempered chains' proposals at once, i.e. C * P
independent stiff ODE solves + the per-chain reduction. One evaluation = one chain's log-likelihood (the unit the
reference counts in Sampler.cpp:129-134).

  value  device-timed (CUDA events on the launching stream), parameter vectors already resident in HBM
  e2e    the same steps through the reference-facing call with HOST buffers: pinned host values -> H2D of the
         rank's slice -> kernels -> (the library's NCCL all-gather + rank-order combination) -> D2H of the result, wall-clock with a device sync
  roofline       dominant kernel (poppk_kernel) against the FP64 FMA peak measured live on this GPU
  cpu_baseline   the reference's CPU implementation (oracle/_ref = its own compiled CVODE stack; else the plain-C
                 port) timed on the box's host cores on a bounded sample of the same workload (rank 0, N=1 only)

--impl reference times only that CPU implementation, one bounded sample per step, same metric and config.
N > 1: launched by torchrun, one rank per GPU; patients are sharded across ranks (strong scaling on the named
workload), per-chain partials combined inside the library (NCCL all-gather + rank-order combination); time = max over ranks.
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


# floating-point operations of one call of the prelude helpers (cellpop_prelude.cuh = SolverCodeGenerator.cpp:122-295), main branch
HELPER_FLOPS = {"hill_function_fixedn2": 4, "hill_function_fixedn4": 6, "hill_function_fixedn10": 10, "hill_function_fixedn16": 10,
                "hill_function_fixedn100": 18, "hill_function": 4 + 2 * 20, "michaelis_menten_function": 4, "safepow": 20, "synthcap": 5, "tQSSA": 9}


def count_rhs_flops(code: str) -> int:
    """F_rhs of SURVEY.md section 8(d) "counted from the generated text": the arithmetic operators of generated_derivative
    (binary + - * /; a sign in front of a term of an `out[i] = +a-b` line counts as the addition it stands for) plus the
    operation count of every helper call."""
    import re

    body = code[code.index("{") + 1:]
    end = body.find("EXPORT_PREFIX void generated_jacobian")
    if end >= 0:
        body = body[:end]
    body = re.sub(r"\[[^\]]*\]", "[]", body)          # indices are not arithmetic
    body = re.sub(r"\d+\.\d+(e[+-]?\d+)?", "K", body)   # literals (no sign inside)
    flops = body.count("*") + body.count("/") + body.count("+") + body.count("-")
    flops -= len(re.findall(r"=\s*[+-]", body))         # the leading sign of an assembled sum is not an operation
    for name, cost in HELPER_FLOPS.items():
        flops += cost * len(re.findall(r"\b" + name + r"\(", body))
    return int(flops)


def cellpop_flop_per_system(N: int, cnt_mean: np.ndarray, f_rhs: float, nout: float) -> float:
    """FLOP(system) from the ORACLE's counters of a sample of the same workload
    (steps, nfe, nsetups, nje, netf, ncfn, nni, ok), with the difference-quotient Jacobian N (F_rhs + 2 N) and the dense LU 2 N^3 / 3."""
    nst, nfe, nsetups, nje, nni = cnt_mean[0], cnt_mean[1], cnt_mean[2], cnt_mean[3], cnt_mean[6]
    q = Q_MEAN
    A = N * (q * (q + 1) / 2 + 2 * (q + 1) + 4) + 60
    return float(nst * A + nfe * f_rhs + nni * (2 * N * N + 9 * N) + nsetups * (2 * N * N + 2 * N ** 3 / 3) + nje * N * (f_rhs + 2 * N) + nout * 2 * N * (q + 1))


COUNTER_NAMES = ["steps", "nfe", "nsetups", "nje", "netf", "ncfn", "nni", "ok"]


def cellpop_cpu_sample(prob, vals, sample: int, cores: int):
    """The reference's CPU implementation on the first `sample` cells of the workload, all chains: evals/s scaled to the full
    population (cells are i.i.d. quasi-random draws) and the solver's counters, from which the algorithmic FLOPs come."""
    import dataclasses

    kind, chk = cpu_checker()
    ps = dataclasses.replace(prob, num_cells=sample, sobol=prob.sobol[:sample])
    t0 = time.perf_counter()
    r = chk.cellpop_counters(ps, vals, threads=cores)
    dt = time.perf_counter() - t0
    C = vals.shape[0]
    cnt = r["counters"].reshape(-1, r["counters"].shape[-1]).mean(axis=0)
    return dict(kind=kind, seconds=dt, evals_per_s=(C / dt) * (sample / prob.num_cells), counters_mean=cnt,
                sample=f"all {C} chains x first {sample} of {prob.num_cells} cells in {dt:.1f} s on {cores} threads ({cpu_model_name()}); scaled by {sample}/{prob.num_cells}")


def cellpop_line(args, workload: str, steps: int, warmup: int, ctx: dict):
    """One complete JSON line (a dict, rank 0; None elsewhere) for a cell_population workload: single GPU, or the cells split over
    the ranks of the torchrun launch (strong scaling, one all-gather + rank-order sum of [C][2 T + 1] doubles inside the library)."""
    import torch

    from bcm3_b200 import _lib
    from bcm3_b200 import synthetic_cellpop as sc

    rank, world, local_rank = ctx["rank"], ctx["world"], ctx["local_rank"]
    w = WORKLOADS[workload]
    prob = sc.make_cellpop_problem(N=w["N"], num_cells=w["cells"], T=w["T"], data_cells=32, seed=1, rate_decades=w.get("rate_decades", 2.0))
    vals = sc.make_chain_values(w["C"])
    C, nvar = vals.shape
    cores = max(1, min(C, os.cpu_count() or 1))
    f_rhs = count_rhs_flops(prob.derivative_code)
    config = {"workload": workload, "species": w["N"], "cells": w["cells"], "chains": C, "timepoints": w["T"], "ode_solves_per_step": C * w["cells"],
              "l2": "256 MB memset between timed iterations"}
    base = {"metric": METRIC_CELLPOP, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic"}

    if args.impl == "reference":
        if rank != 0:
            return None
        sample = min(w["cells"], 2000 if w["N"] <= 20 else 200)
        runs = [cellpop_cpu_sample(prob, vals, sample, cores) for _ in range(warmup + steps)][warmup:]
        value = statistics.mean(r["evals_per_s"] for r in runs)
        return dict(base, impl="reference", value=value, ms_per_step=1e3 * statistics.mean(r["seconds"] for r in runs), config=config,
                    cpu_baseline={"value": value, "unit": UNIT, "cores": cores, "kind": runs[-1]["kind"], "sample": runs[-1]["sample"]},
                    e2e={"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0)

    dev = torch.device("cuda", local_rank)
    h_vals = torch.from_numpy(vals).pin_memory()
    flush = ctx["flush"]
    fp64_peak = ctx["fp64_peak"]
    if world == 1:
        from bcm3_b200.cellpop import CellPopEvaluator

        ev = CellPopEvaluator(prob, device=local_rank)
        h_logp = np.empty(C)
        h_status = np.empty(C, dtype=np.int32)

        def step():
            _lib.check(ev.lib.bcm3b200_evaluate_batch(ev.handle, C, nvar, h_vals.data_ptr(), h_logp.ctypes.data, h_status.ctypes.data))
            return h_logp, h_status
        closer = ev
    else:
        from bcm3_b200.parallel import ShardedCellPopLikelihood

        lk = ShardedCellPopLikelihood(prob, rank, world, local_rank)
        ev = lk.evaluator

        def step():
            return lk.evaluate(h_vals)  # ONE C-ABI call: H2D of the batch, kernels, library exchange, data likelihood, D2H of logp
        closer = lk
        config["sharding"] = f"cells over {world} ranks; library exchange: NCCL all-gather of [{C}][{2 * w['T'] + 1}] doubles per rank + rank-order sum"

    for _ in range(warmup):
        flush.zero_()
        step()
    ctx["barrier"]()
    launches0 = ev.get_stat("total_kernel_launches")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    kernel_ms, wall = [], 0.0
    ctx["barrier"]()
    t_begin = time.perf_counter()
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        logp, status = step()
        wall += time.perf_counter() - t0
        kernel_ms.append(ev.get_stat("last_kernel_us") / 1e3)
    ctx["barrier"]()
    t_end = time.perf_counter()
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    launches = ev.get_stat("total_kernel_launches") - launches0
    cells_local = ev.get_stat("num_cells_local")
    # device time of a step = the library's own events around its kernels (host-buffer entry: the H2D of the 768-byte batch
    # precedes the first event), max over ranks; e2e = wall clock around the call, max over ranks
    tot = ctx["max_over_ranks"]([sum(kernel_ms), 1e3 * wall, statistics.mean(kernel_ms)])
    nl = ctx["sum_over_ranks"]([float(launches)])[0]
    closer.close()
    if rank != 0:
        return None
    total_ms, wall_ms, k_ms = tot
    line = dict(base, value=C * steps / (total_ms * 1e-3), ms_per_step=total_ms / steps, config=config,
                e2e={"value": C * steps / (wall_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(vals.nbytes) * world, "d2h_bytes_per_step": int(C * 8) * world,
                     "ms_per_step": wall_ms / steps},
                gpu_launches=int(nl), clocks=clocks, check={"logp0": float(logp[0]), "status_ok": bool((np.asarray(status) == 0).all())})
    # CPU arm of the same run (rank 0's host cores) on a bounded sample; its counters give the ALGORITHMIC FLOPs of the roofline
    sample = min(w["cells"], 4000 if w["N"] <= 20 else 400)
    cb = cellpop_cpu_sample(prob, vals, sample, cores)
    cnt = cb["counters_mean"]
    flop_sys = cellpop_flop_per_system(w["N"], cnt, f_rhs, w["T"])
    achieved = flop_sys * C * cells_local / (k_ms * 1e-3) / 1e12
    line["cpu_baseline"] = {"value": cb["evals_per_s"], "unit": UNIT, "cores": cores, "kind": cb["kind"], "sample": cb["sample"]}
    line["config"]["mean_steps_per_solve"] = float(cnt[0])
    line["roofline"] = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": ncu_traffic(workload),
                        "kernel": "cellpop_group_kernel", "kernel_ms": k_ms, "flop_per_system": flop_sys, "systems_per_launch": C * cells_local,
                        "rhs_flops_from_generated_text": f_rhs, "oracle_counters_mean": {k: float(v) for k, v in zip(COUNTER_NAMES, cnt)},
                        "flops_from": f"SURVEY 8(d) formula on the {cb['kind']} CPU run's counters of this run's sample ({sample} cells x {C} chains)",
                        "note": "per GPU, on its shard of the cells" if world > 1 else "whole workload on one GPU",
                        "peak_source": "measured live: bcm3b200_measure_fp64_peak (DFMA chains), MEASURED_PEAKS.json has no FP64 entry"}
    return line


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


def reference_line(args, workload: str):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores (rank 0 only)."""
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
    # the prefix sample is scaled linearly to the full population: checked here once with a sample of twice the size
    double = time_cpu(prob, vals, 2 * target)
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "pk_model": w["pk"], "individuals": w["P"], "chains": w["C"], "timepoints": w["T"],
                   "note": "each step = a bounded sample of the workload (all chains x a prefix of the individuals), evals/s scaled to the full size"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"],
                         "build": "oracle/_ref: the reference's CVODE 5.3.0 + src/odecommon, -O3 -march=x86-64-v3 (the reference's own flag is -march=native)",
                         "scaling_check": {"sample_individuals": [last["P_sample"], double["P_sample"]], "evals_per_s": [last["evals_per_s"], double["evals_per_s"]]}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def poppk_line(args, workload: str, steps: int, warmup: int, ctx: dict):
    """One complete JSON line (a dict, rank 0; None elsewhere) for a PopPK workload."""
    import torch

    from bcm3_b200.parallel import ShardedPopPKLikelihood, combine_partials, shard_bounds

    rank, world, local_rank = ctx["rank"], ctx["world"], ctx["local_rank"]
    dev = torch.device("cuda", local_rank)
    barrier = ctx["barrier"]
    line = None
    w = WORKLOADS[workload]
    prob, vals = make_workload(workload)
    C, nvar = vals.shape
    P, T = w["P"], w["T"]
    lk = ShardedPopPKLikelihood(prob, rank, world, local_rank, block_size=args.block_size)
    ev = lk.evaluator

    # inputs: pinned host copy (e2e) and a device-resident copy (value)
    h_vals = torch.from_numpy(vals).pin_memory()
    d_vals = torch.from_numpy(vals).to(dev)
    d_partial = torch.empty((3, C), dtype=torch.float64, device=dev)
    h_partial = torch.empty((3, C), dtype=torch.float64).pin_memory()
    flush = ctx["flush"]
    stream = torch.cuda.current_stream(dev)

    def device_step():
        ev.evaluate_device(d_vals.data_ptr(), C, nvar, d_partial.data_ptr(), stream.cuda_stream)
        if world > 1:  # the library's exchange: one NCCL all-gather + rank-order combination, on the same stream
            ev.exchange(d_partial.data_ptr(), C, stream.cuda_stream)

    fp64_peak = ctx["fp64_peak"]

    # ---- device-resident timing: W warm-up steps, then exactly K timed steps ----
    for _ in range(warmup):
        flush.zero_()
        device_step()
    barrier()
    launches0 = ev.get_stat("total_kernel_launches")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
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
    total_ms = ctx["max_over_ranks"]([sum(step_ms)])[0]
    launches_all = int(ctx["sum_over_ranks"]([float(launches)])[0])
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
    for _ in range(steps):
        flush.zero_()
        out = e2e_step()
    barrier()
    e2e_s = ctx["max_over_ranks"]([time.perf_counter() - t0])[0]
    e2e_logp = np.array(out, dtype=np.float64, copy=True)

    lo, hi = shard_bounds(P, rank, world)
    h2d = sum(C * (16 + 2 * (shard_bounds(P, r, world)[1] - shard_bounds(P, r, world)[0])) * 8 for r in range(world))
    d2h = (C * 8 + C * 4) if world == 1 else world * 3 * C * 8

    if rank == 0:
        N = 2 if w["pk"] == "one" else 3
        value = C * steps / (total_ms * 1e-3)
        ms_per_step = total_ms / steps
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
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "pk_model": w["pk"], "individuals": P, "chains": C, "timepoints": T, "t_end_h": w["t_end"],
                       "ode_solves_per_step": C * P, "sharding": f"individuals over {world} rank(s); library exchange: NCCL all-gather of [3][{C}] doubles per rank + rank-order combination",
                       "l2": "256 MB memset between timed iterations", "block_size": args.block_size or "auto"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                         "traffic": ncu_traffic(workload), "kernel": "poppk_kernel", "kernel_ms": k_ms,
                         "flop_per_system": flop_per_system, "systems_per_launch": systems_per_launch,
                         "peak_source": "measured live: bcm3b200_measure_fp64_peak (DFMA chains), MEASURED_PEAKS.json has no FP64 entry",
                         "hbm": {"achieved": bytes_per_launch / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": bytes_per_launch / (k_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": bytes_per_launch,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
            "e2e": {"value": C * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / steps},
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
    lk.close()
    return line


def make_context(args):
    """Process-wide state shared by the lines of one run: rank layout, torch.distributed (plumbing: barriers and max-over-ranks
    of the timings), the L2 flush buffer, the FP64 peak of this GPU."""
    import torch
    import torch.distributed as dist

    from bcm3_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = dict(rank=rank, world=world, local_rank=local_rank)
    if args.impl == "reference":
        return ctx
    if world == 1 and args.gpus > 1:
        raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available() or _lib.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_over_ranks(values, op):
        t = torch.tensor(list(values), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=op)
        return [float(x) for x in t.tolist()]

    ctx.update(barrier=barrier, max_over_ranks=lambda v: reduce_over_ranks(v, dist.ReduceOp.MAX),
               sum_over_ranks=lambda v: reduce_over_ranks(v, dist.ReduceOp.SUM),
               flush=torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev),  # > 126 MB L2
               fp64_peak=_lib.measure_fp64_peak(local_rank))
    return ctx


# the cell_population half of BASELINE's metric, reported next to the PopPK headline: (workload, runs at N > 1, steps cap)
SECONDARY = [("cellpop_12sp_10k_x16", False, 8), ("cellpop_50sp_100k_x16", True, 2)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="poppk_two_100k_x64", choices=sorted(WORKLOADS))
    ap.add_argument("--block-size", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="only the headline workload")
    args = ap.parse_args()
    workload = args.workload
    ctx = make_context(args)
    rank, world = ctx["rank"], ctx["world"]
    warmup = max(args.warmup, 3)

    def one_line(name, steps):
        if WORKLOADS[name].get("kind") == "cellpop":
            return cellpop_line(args, name, steps, warmup if args.impl != "reference" else args.warmup, ctx)
        if args.impl == "reference":
            return reference_line(args, name) if rank == 0 else None
        return poppk_line(args, name, steps, warmup, ctx)

    line = one_line(workload, args.steps)
    if workload == "poppk_two_100k_x64" and not args.no_secondary:
        secondary = []
        for name, multi_gpu, cap in SECONDARY:
            if world > 1 and not multi_gpu:
                continue
            sec = one_line(name, min(args.steps, cap))
            if sec is not None:
                secondary.append(sec)
        if line is not None:
            line["secondary"] = secondary
    if rank == 0 and line is not None:
        print(json.dumps(line))
    if world > 1 and args.impl != "reference":
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
