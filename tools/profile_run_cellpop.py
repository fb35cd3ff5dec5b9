"""Short single-GPU run for ncu: a few batched evaluations of a mid-size cell_population workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator
kernel = sys.argv[1] if len(sys.argv) > 1 else "auto"
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
prob = sc.make_cellpop_problem(N=12, num_cells=cells, T=50, data_cells=8)
vals = sc.make_chain_values(8)
ev = CellPopEvaluator(prob, kernel=kernel)
for i in range(2):
    logp, status = ev.evaluate(vals)
    print(i, logp[:2], ev.get_stat("last_kernel_us"), "us")
ev.close()
