"""What the per-cell time_course data kind costs on top of the integration: wall time of evaluate() for the same cells read as
a population average and as per-cell trajectories with the matching (cell-likelihood kernel + n^2 doubles per chain to the host +
the matching, one chain per host thread). usage: python tools/time_course_timing.py [cells ...]"""
import dataclasses
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcm3_b200 import _lib, synthetic_cellpop as sc  # noqa: E402
from bcm3_b200.cellpop import CellPopEvaluator  # noqa: E402

C = 16
for n in [int(a) for a in sys.argv[1:]] or [128, 512, 2048]:
    tc = sc.make_time_course_problem(N=8, num_cells=n, T=20, seed=71)
    avg = dataclasses.replace(tc, data_kind="time_course_population_average", observed=np.nanmean(tc.observed, axis=0)[None, :])
    vals = sc.make_chain_values(C, seed=71)
    out = {}
    for name, prob in (("population_average", avg), ("time_course", tc)):
        ev = CellPopEvaluator(prob)
        ev.evaluate(vals)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            logp, _ = ev.evaluate(vals)
            best = min(best, time.perf_counter() - t0)
        out[name] = (best, ev.get_stat("last_kernel_us") / 1e3, logp[0])
        ev.close()
    # the matching alone, one chain on one core
    rng = np.random.default_rng(n)
    cost = rng.normal(0.0, 30.0, (n, n))
    t0 = time.perf_counter()
    _lib.match_cells(cost)
    t_match = time.perf_counter() - t0
    print(f"cells {n:5d} x {C} chains: evaluate {out['population_average'][0] * 1e3:8.2f} ms as population average (integration kernels "
          f"{out['population_average'][1]:.2f} ms), {out['time_course'][0] * 1e3:8.2f} ms as time_course; one random {n} x {n} matching on one core "
          f"{t_match * 1e3:.2f} ms; logp0 {out['time_course'][2]:.6f}", flush=True)
