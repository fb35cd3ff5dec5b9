"""Short single-GPU run for ncu: a few batched evaluations of a 50-species stiff cell_population workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
prob = sc.make_cellpop_problem(N=N, num_cells=cells, T=50, data_cells=8, rate_decades=4.0)
vals = sc.make_chain_values(2)
ev = CellPopEvaluator(prob)
for i in range(2):
    logp, status = ev.evaluate(vals)
    print("N", N, i, logp[:2], ev.get_stat("last_kernel_us"), "us")
ev.close()
