"""Development report (run under gpurun): cell_population GPU path against the compiled reference, with the reference's own
noise floor next to every number. One line per (case, kernel): per-chain relative error of logp, the floor, the fraction of cells
whose step counts equal the reference's and the fraction on which the reference's own two builds agree."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcm3_b200 import synthetic_cellpop as sc  # noqa: E402
from bcm3_b200.cellpop import CellPopEvaluator  # noqa: E402
from tests.util import CELLPOP_GOLDEN_NAMES, load_cellpop_golden, reference_noise_floor_cellpop, rel_err  # noqa: E402


COMPILE_ONLY = "--compile-only" in sys.argv  # build every kernel library the report needs into the in-tree cache (no GPU)


def gpu(prob, vals, **kw):
    if COMPILE_ONLY:
        CellPopEvaluator(prob, compile_only=True, **kw).close()
        return None, None
    ev = CellPopEvaluator(prob, **kw)
    logp, status = ev.evaluate(vals)
    d = ev.diagnostics()
    ev.close()
    return logp, d


def line(name, logp, d, want_logp, want_steps, want_avg, floor, floor_steps):
    if logp is None:
        return
    err = rel_err(logp, want_logp)
    print(f"{name:46s} rel err {np.array2string(err, precision=2)}  abs {np.abs(logp - want_logp).max():.2e}  floor {np.array2string(np.asarray(floor), precision=2)}  "
          f"steps== {(d['cell_steps'] == want_steps).mean():.3f} (ref builds {floor_steps:.3f})  avg diff {np.abs(d['population_average'] - want_avg).max():.2e}", flush=True)


for name in CELLPOP_GOLDEN_NAMES:
    prob, gold = load_cellpop_golden(name)
    for kw in (dict(), dict(rhs_lanes=False)):
        logp, d = gpu(prob, gold["values"], **kw)
        line(name + (" scalar-rhs" if kw else ""), logp, d, gold["logp"], gold["cell_steps"], gold["population_average"], gold["noise_floor"], float(gold["noise_floor_step_match"]))

for N, decades in ((3, 2.0), (5, 2.0), (7, 2.0), (12, 2.0), (16, 3.0), (24, 3.0), (33, 3.0), (50, 4.0)):
    prob = sc.make_cellpop_problem(N=N, num_cells=96, T=12, data_cells=4, seed=40 + N, rate_decades=decades)
    vals = sc.make_chain_values(2, seed=N)
    ra, floor, fsteps = reference_noise_floor_cellpop(prob, vals, threads=4) if not COMPILE_ONLY else (dict(logp=None, cell_steps=None, population_average=None), None, 0)
    kernels = [dict(), dict(rhs_lanes=False)] + ([dict(kernel="warp"), dict(kernel="thread")] if N <= 12 else [])
    for kw in kernels:
        logp, d = gpu(prob, vals, **kw)
        line(f"fresh N={N} {kw or 'default'}", logp, d, ra["logp"], ra["cell_steps"], ra["population_average"], floor, fsteps)
