"""Times + checks one build variant of the library (BCM3B200_LIB) on the two bench workloads."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_ONE, PK_TWO
from bcm3_b200.poppk import PopPKEvaluator
tag = os.environ.get("BCM3B200_LIB", "default")
blocks = [int(b) for b in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0"])]
chk = oracle.load("port")
for pk, P, C in ((PK_TWO, 20000, 16), (PK_ONE, 1000, 16)):
    prob = syn.make_poppk_problem(pk, P=P, T=10, t_end=72.0, seed=1)
    vals = syn.make_chain_values(prob, C)
    want = None
    for b in blocks:
        ev = PopPKEvaluator(prob, block_size=b)
        best = 1e9
        for it in range(4):
            logp, status = ev.evaluate(vals)
            best = min(best, ev.get_stat("last_kernel_us") / 1e3)
        ev.close()
        if want is None:
            sub = slice(0, 2)
            want = chk.poppk_evaluate(prob, vals[sub], threads=2)["logp"] if P <= 20000 else None
        rel = np.abs(logp[:2] - want) / np.abs(want)
        print(f"{os.path.basename(tag):28s} pk={pk} P={P} C={C} block={b:3d}: kernel {best:8.3f} ms  max rel err vs port {rel.max():.2e}", flush=True)
