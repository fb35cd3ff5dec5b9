"""Development check (run under gpurun): group-kernel variants, timing at config-3 size and odd sizes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator

def timing(N, cells, T, C, reps=3, check=False, decades=2.0):
    prob = sc.make_cellpop_problem(N=N, num_cells=cells, T=T, data_cells=8, rate_decades=decades)
    vals = sc.make_chain_values(C)
    try:
        ev = CellPopEvaluator(prob, kernel="group")
        ts = []
        for it in range(reps):
            logp, status = ev.evaluate(vals)
            ts.append(ev.get_stat('last_kernel_us') / 1e3)
        d = ev.diagnostics()
        msg = f"N={N} cells={cells} C={C} env={ {k: v for k, v in os.environ.items() if k.startswith('BCM3B200_CELLPOP')} }: kernel ms {min(ts):.2f} steps mean {d['cell_steps'].mean():.1f} ok {d['cell_status'].mean():.3f}"
        if check:
            r = oracle.load("ref" if oracle.available("ref") else "port").cellpop_evaluate(prob, vals, threads=16)
            msg += f" max rel logp {np.max(np.abs(logp - r['logp']) / np.abs(r['logp'])):.2e}"
        print(msg, flush=True)
        ev.close()
    except Exception as e:
        print(f"N={N} FAILED: {e}", flush=True)

if __name__ == "__main__":
    what = sys.argv[1]
    if what == "smoke":
        prob = sc.make_cellpop_problem(N=12, num_cells=64, T=12, data_cells=4, seed=2)
        vals = sc.make_chain_values(2)
        r = oracle.load("ref").cellpop_evaluate(prob, vals, threads=2, want_cell_values=True, want_steps=True)
        for k in ("group", "warp", "thread"):
            ev = CellPopEvaluator(prob, kernel=k)
            logp, st = ev.evaluate(vals)
            d = ev.diagnostics()
            ev.close()
            diff = np.nanmax(np.abs(d["cell_values"] - r["cell_values"]), axis=1)  # [C][cells]
            worst = np.argsort(diff.ravel())[-5:]
            print(k, "logp", logp, "ref", r["logp"], "steps equal", (d["cell_steps"] == r["cell_steps"]).mean())
            for w in worst:
                c, cell = divmod(int(w), prob.num_cells)
                print(f"   chain {c} cell {cell}: max |dv| {diff[c, cell]:.2e} steps gpu {d['cell_steps'][c, cell]} ref {r['cell_steps'][c, cell]}")
    elif what == "sizes":
        for N in (3, 6, 7, 16, 24, 33, 50):
            timing(N, 300, 20, 2, reps=1, check=True, decades=3.0)
    else:
        timing(12, 10000, 50, 16)
