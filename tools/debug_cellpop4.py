import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator
np.set_printoptions(linewidth=220, precision=12)
base = sc.make_cellpop_problem(N=12, num_cells=64, T=20, data_cells=4)
code = sc.SIGNATURE + "\n{\n\tOdeReal ratelaws[3];\n\tratelaws[0] = parameters[0];\n\tratelaws[1] = parameters[2];\n\tratelaws[2] = (parameters[1]*constant_species[0]);\n\tout[0] = +ratelaws[0];\n\tout[1] = +ratelaws[1];\n\tout[2] = +ratelaws[2];\n}\n"
vals = sc.make_chain_values(2)
for obs in ([0], [1], [2]):
    import dataclasses
    p = dataclasses.replace(base, derivative_code=code, num_species=3, initial_conditions=np.zeros(3), obs_species=obs, timepoints=np.array([0.0, 1.0, 2.0]), observed=np.zeros((1, 3)))
    ev = CellPopEvaluator(p); ev.evaluate(vals); d = ev.diagnostics(); ev.close()
    r = oracle.load("port").cellpop_evaluate(p, vals, want_cell_values=True, want_steps=True)
    diff = np.abs(d["cell_values"] - r["cell_values"]) / np.maximum(np.abs(r["cell_values"]), 1e-300)
    print("obs", obs, "max rel diff", np.nanmax(diff), "row16 gpu", d["cell_values"][0, :, 16], "cpu", r["cell_values"][0, :, 16])
