import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator
prob = sc.make_cellpop_problem(N=12, num_cells=64, T=20, data_cells=4)
vals = sc.make_chain_values(2)
for rep, name in ((0, "steps"), (1, "nfe"), (2, "nsetups"), (3, "nje")):
    os.environ["BCM3B200_CELLPOP_REPORT"] = str(rep)
    ev = CellPopEvaluator(prob); ev.evaluate(vals); d = ev.diagnostics(); ev.close()
    r = oracle.load("port").cellpop_evaluate(prob, vals, want_steps=True)
    print(f"{name:8s} gpu mean {d['cell_steps'].mean():8.2f} cpu mean {r['cell_steps'].mean():8.2f}   row16: gpu {d['cell_steps'][:,16]} cpu {r['cell_steps'][:,16]}  row0: gpu {d['cell_steps'][:,0]} cpu {r['cell_steps'][:,0]}")
