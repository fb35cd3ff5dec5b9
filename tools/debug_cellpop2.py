import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator
np.set_printoptions(linewidth=220, precision=6)
prob = sc.make_cellpop_problem(N=12, num_cells=64, T=20, data_cells=4)
vals = sc.make_chain_values(2)
# (1) determinism
ev = CellPopEvaluator(prob)
ev.evaluate(vals); d1 = ev.diagnostics(); ev.evaluate(vals); d2 = ev.diagnostics()
print("run-to-run identical:", np.array_equal(d1["cell_values"], d2["cell_values"], equal_nan=True), np.array_equal(d1["cell_steps"], d2["cell_steps"]))
ev.close()
# (2) all cells share one non-trivial sobol point
for row in (7, 8, 10, 3):
    p2 = sc.make_cellpop_problem(N=12, num_cells=64, T=20, data_cells=4)
    p2.sobol = np.tile(prob.sobol[row], (64, 1))
    ev = CellPopEvaluator(p2); lg, _ = ev.evaluate(vals); d = ev.diagnostics(); ev.close()
    r = oracle.load("port").cellpop_evaluate(p2, vals, want_cell_values=True, want_steps=True)
    print("row", row, "gpu steps", d["cell_steps"][0, :4], "cpu", r["cell_steps"][0, :4], "max diff", np.nanmax(np.abs(d["cell_values"] - r["cell_values"])),
          "all gpu cells equal:", (d["cell_steps"][0] == d["cell_steps"][0, 0]).all())
