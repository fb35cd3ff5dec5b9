"""Times the PopPK kernel with the two grid orders (option chain_fastest_grid) on one rank's share of config 5 at 8 GPUs
(12 500 individuals x 64 chains) and on the whole config; checks that the results are the same bits."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_TWO
from bcm3_b200.poppk import PopPKEvaluator
for P in (12500, 25000, 100000):
    prob = syn.make_poppk_problem(PK_TWO, P=P, T=10, t_end=72.0, seed=1)
    vals = syn.make_chain_values(prob, 64)
    res = {}
    for flag in (0, 1):
        ev = PopPKEvaluator(prob)
        ev.set_option("chain_fastest_grid", flag)
        best = 1e18
        for _ in range(5):
            logp, _ = ev.evaluate(vals)
            best = min(best, ev.get_stat("last_kernel_us"))
        ev.close()
        res[flag] = (best / 1e3, logp)
    print(f"P={P}: grid (blocks, chains) {res[0][0]:.2f} ms, grid (chains, blocks) {res[1][0]:.2f} ms, same bits {np.array_equal(res[0][1], res[1][1])}", flush=True)
