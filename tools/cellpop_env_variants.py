"""Times a cell_population workload under the BCM3B200_CELLPOP_* build overrides given on the command line, one process per
variant (the overrides are read when the model's kernel library is built). Kernel times are best-of-4 without an L2 flush:
for comparing variants, not bench values.
usage: python tools/cellpop_env_variants.py [--n 12|50 --cells 10000 --chains 16 --decades 2.0] "GROUP_LOCKSTEP=0" "LU_SKIP_ZEROS=0,SOLVE_SLOTTED=0" ... """
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=12)
ap.add_argument("--cells", type=int, default=10000)
ap.add_argument("--chains", type=int, default=16)
ap.add_argument("--decades", type=float, default=2.0)
ap.add_argument("--compile-only", action="store_true", help="build the variants' kernel libraries into the in-tree cache (no GPU needed)")
ap.add_argument("variants", nargs="*")
args = ap.parse_args()
RUN = r"""
import sys
sys.path.insert(0, %r)
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator
prob = sc.make_cellpop_problem(N=%d, num_cells=%d, T=50, data_cells=16, seed=1, rate_decades=%r)
vals = sc.make_chain_values(%d)
if %r:
    CellPopEvaluator(prob, compile_only=True).close()
    print("compiled")
    sys.exit(0)
ev = CellPopEvaluator(prob)
best = 1e18
for i in range(4):
    logp, status = ev.evaluate(vals)
    best = min(best, ev.get_stat("last_kernel_us"))
print("kernel %%.2f ms  logp0 %%.12g  logp1 %%.12g" %% (best / 1e3, logp[0], logp[1]))
ev.close()
""" % (ROOT, args.n, args.cells if not args.compile_only else 8, args.decades, args.chains, args.compile_only)

for spec in ["default"] + args.variants:
    env = dict(os.environ)
    if spec != "default":
        for kv in spec.split(","):
            k, v = kv.split("=")
            env["BCM3B200_CELLPOP_" + k] = v
    r = subprocess.run([sys.executable, "-c", RUN], env=env, capture_output=True, text=True)
    print(f"N={args.n} cells={args.cells} {spec:44s} {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
