import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator
np.set_printoptions(linewidth=220, precision=6)
prob = sc.make_cellpop_problem(N=12, num_cells=64, T=20, data_cells=4)
vals = sc.make_chain_values(2)
ev = CellPopEvaluator(prob); ev.evaluate(vals); het = ev.diagnostics(); ev.close()
cpu = oracle.load("port").cellpop_evaluate(prob, vals, want_cell_values=True, want_steps=True)
worst = []
for row in range(64):
    p2 = sc.make_cellpop_problem(N=12, num_cells=4, T=20, data_cells=4)
    p2.sobol = np.tile(prob.sobol[row], (4, 1))
    ev = CellPopEvaluator(p2); ev.evaluate(vals); d = ev.diagnostics(); ev.close()
    same = np.array_equal(d["cell_values"][:, :, 0], het["cell_values"][:, :, row], equal_nan=True)
    diff = np.nanmax(np.abs(d["cell_values"][:, :, 0] - cpu["cell_values"][:, :, row]))
    worst.append((diff, row, same, d["cell_steps"][:, 0].tolist(), cpu["cell_steps"][:, row].tolist()))
print("homogeneous GPU == heterogeneous GPU for all rows:", all(w[2] for w in worst))
for w in sorted(worst, reverse=True)[:6]:
    print("row %d: gpu-vs-cpu max diff %.3e  gpu steps %s cpu steps %s" % (w[1], w[0], w[3], w[4]))
