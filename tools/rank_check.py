"""Development check (run under gpurun): effect of ranking the patients by expected work (poppk_rank_kernel) at several
batch sizes and on a heterogeneous trial (mixed dosing intervals)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_TWO
from bcm3_b200.poppk import PopPKEvaluator
cases = [(100000, 16, False), (100000, 16, True), (5000, 16, False), (2500, 16, False), (1000, 16, False)] if len(sys.argv) < 2 else [(int(sys.argv[1]), int(sys.argv[2]), False)]
for P, C, het in cases:
    prob = syn.make_poppk_problem(PK_TWO, P=P, T=10, t_end=72.0, seed=1, heterogeneous=het)
    vals = syn.make_chain_values(prob, C)
    for flag in (True, False):
        ev = PopPKEvaluator(prob, sort_patients=flag)
        ts = []
        for i in range(4):
            logp, st = ev.evaluate(vals)
            ts.append(ev.get_stat("last_kernel_us") / 1e3)
        print(f"P={P} C={C} heterogeneous={het} ranked={flag}: kernel ms {min(ts):.3f} logp0 {logp[0]!r}")
        ev.close()
