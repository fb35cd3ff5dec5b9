import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator
np.set_printoptions(linewidth=220, precision=6)
novar = len(sys.argv) > 1 and sys.argv[1] == "novar"
prob = sc.make_cellpop_problem(N=12, num_cells=64, T=20, data_cells=4)
if novar:
    prob.variability = []
    prob.sobol = np.zeros((64, 0))
mode = sys.argv[1] if len(sys.argv) > 1 else ""
if mode.startswith("keep"):
    keep = [int(ch) for ch in mode[4:]]
    prob.variability = [prob.variability[k] for k in keep]
    prob.sobol = np.ascontiguousarray(prob.sobol[:, keep])
vals = sc.make_chain_values(2)
ev = CellPopEvaluator(prob)
logp, status = ev.evaluate(vals)
d = ev.diagnostics()
r = oracle.load("port").cellpop_evaluate(prob, vals, threads=2, want_cell_values=True, want_steps=True, want_average=True)
print("gpu logp", logp, "cpu", r["logp"])
print("steps gpu", d["cell_steps"][0, :16]); print("steps cpu", r["cell_steps"][0, :16])
diff = np.abs(d["cell_values"] - r["cell_values"])
print("max diff per timepoint (chain 0)", np.nanmax(diff[0], axis=1))
c, t, i = np.unravel_index(np.nanargmax(diff), diff.shape)
print("worst: chain", c, "t", t, "cell", i, d["cell_values"][c, t, i], r["cell_values"][c, t, i])
print("cell traj gpu", d["cell_values"][c, :, i]); print("cell traj cpu", r["cell_values"][c, :, i])
