"""Times cfg 3 (12 species, 10 000 cells, 16 chains) under the BCM3B200_CELLPOP_* build overrides given on the command line,
one process per variant (the overrides are read when the model's kernel library is built).
usage: python tools/cellpop_env_variants.py "LOCKSTEP=0" "GROUP_WARPS=8" "GROUP=8,LOCKSTEP=1" """
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = r"""
import sys
sys.path.insert(0, %r)
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator
prob = sc.make_cellpop_problem(N=12, num_cells=10000, T=50, data_cells=16, seed=5)
vals = sc.make_chain_values(16, seed=5)
ev = CellPopEvaluator(prob)
best = 1e18
for i in range(4):
    logp, status = ev.evaluate(vals)
    best = min(best, ev.get_stat("last_kernel_us"))
print("kernel %%.2f ms  logp0 %%.10g" %% (best / 1e3, logp[0]))
ev.close()
""" % ROOT

for spec in ["default"] + sys.argv[1:]:
    env = dict(os.environ)
    env["BCM3B200_CACHE"] = os.path.join(ROOT, "gpurun_out", "variant_cache")
    if spec != "default":
        for kv in spec.split(","):
            k, v = kv.split("=")
            env["BCM3B200_CELLPOP_" + k] = v
    r = subprocess.run([sys.executable, "-c", RUN], env=env, capture_output=True, text=True)
    print(f"{spec:32s} {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
