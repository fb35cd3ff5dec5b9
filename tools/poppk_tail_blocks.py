"""The tail of the block schedule at the per-rank size of an 8-GPU run of config 5 (12 500 patients x 64 chains = 800 k threads =
14.08 waves of 148 x 384): kernel time against the block size, on one GPU. usage: python tools/poppk_tail_blocks.py [P ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_TWO
from bcm3_b200.poppk import PopPKEvaluator

for P in [int(a) for a in sys.argv[1:]] or [12500, 25000, 100000]:
    prob = syn.make_poppk_problem(PK_TWO, P=P, T=10, t_end=72.0, seed=1)
    vals = syn.make_chain_values(prob, 64)
    ref = None
    for b in (128, 96, 64, 32):
        ev = PopPKEvaluator(prob, block_size=b)
        times = []
        for it in range(5):
            logp, status = ev.evaluate(vals)
            times.append(ev.get_stat("last_kernel_us") / 1e3)
        ev.close()
        if ref is None:
            ref = logp.copy()
        print(f"P={P:6d} C=64 block={b:3d}: kernel best {min(times[1:]):8.3f} ms  median {np.median(times[1:]):8.3f} ms  same bits as block 128: {np.array_equal(ref, logp)}", flush=True)
