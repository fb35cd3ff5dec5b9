"""Development check (run under gpurun): effect of ranking the patients on a heterogeneous trial (mixed dosing intervals)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_TWO
from bcm3_b200.poppk import PopPKEvaluator
for het in (False, True):
    prob = syn.make_poppk_problem(PK_TWO, P=100000, T=10, t_end=72.0, seed=1, heterogeneous=het)
    vals = syn.make_chain_values(prob, 16)
    for flag in (True, False):
        ev = PopPKEvaluator(prob, sort_patients=flag)
        for i in range(3):
            logp, st = ev.evaluate(vals)
        print("heterogeneous", het, "ranked", flag, "kernel ms", ev.get_stat("last_kernel_us") / 1e3, "logp0", logp[0])
        ev.close()
