"""Short single-GPU run for ncu: a few batched evaluations of a mid-size PopPK workload."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_ONE, PK_TWO
from bcm3_b200.poppk import PopPKEvaluator
pk = PK_TWO if (len(sys.argv) < 2 or sys.argv[1] == "two") else PK_ONE
P = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
C = int(sys.argv[3]) if len(sys.argv) > 3 else 16
prob = syn.make_poppk_problem(pk, P=P, T=10, t_end=72.0, seed=1)
vals = syn.make_chain_values(prob, C)
ev = PopPKEvaluator(prob)
for i in range(3):
    logp, status = ev.evaluate(vals)
    print(i, logp[:2], ev.get_stat("last_kernel_us"), "us")
ev.close()
