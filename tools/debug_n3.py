"""Development aid (run under gpurun): per-cell differences between the GPU path and the compiled reference on the fresh
3-species problem of tools/gpu_parity_report.py."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prob = sc.make_cellpop_problem(N=N, num_cells=96, T=12, data_cells=4, seed=40 + N, rate_decades=2.0)
vals = sc.make_chain_values(2, seed=N)
ref = oracle.load("ref")
a = ref.cellpop_evaluate(prob, vals, threads=4, want_cell_values=True, want_steps=True, want_average=True)
for kernel in ("auto", "thread"):
    ev = CellPopEvaluator(prob, kernel=kernel)
    logp, status = ev.evaluate(vals)
    d = ev.diagnostics()
    ev.close()
    print(kernel, "logp", logp, "ref", a["logp"])
    cv, rv = d["cell_values"], a["cell_values"]  # [C][T][cells]
    diff = np.abs(cv - rv)
    for c in range(cv.shape[0]):
        per_cell = np.nanmax(diff[c], axis=0)
        order = np.argsort(-per_cell)[:8]
        print(" chain", c, "max per-cell diff:", [(int(i), float(f"{per_cell[i]:.2e}"), int(d["cell_steps"][c, i]), int(a["cell_steps"][c, i])) for i in order])
        print("   median per-cell diff %.2e, cells with different steps %d" % (np.median(per_cell), (d["cell_steps"][c] != a["cell_steps"][c]).sum()))
        print("   avg gpu", d["population_average"][c][:6], "\n   avg ref", a["population_average"][c][:6])
        worst = order[0]
        print("   worst cell trajectory gpu", cv[c, :, worst], "\n   ref", rv[c, :, worst])
