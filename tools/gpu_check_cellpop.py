"""Development check (run under gpurun): cell_population GPU path vs the CPU checkers."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator

def run(N, cells, T, C, decades=2.0, kinds=("ref", "port"), kernel="auto"):
    prob = sc.make_cellpop_problem(N=N, num_cells=cells, T=T, data_cells=8, rate_decades=decades)
    vals = sc.make_chain_values(C)
    t0 = time.time(); ev = CellPopEvaluator(prob, kernel=kernel); print(f"kernel={kernel} N={N} cells={cells} T={T} C={C}: setup {time.time()-t0:.1f}s")
    for it in range(2):
        t0 = time.time(); logp, status = ev.evaluate(vals); dt = time.time() - t0
    d = ev.diagnostics()
    print(f"  gpu logp[:4]={logp[:4]} e2e {dt*1e3:.1f} ms kernel {ev.get_stat('last_kernel_us')/1e3:.2f} ms steps mean {d['cell_steps'].mean():.1f} status ok {d['cell_status'].mean():.3f}")
    for kind in kinds:
        if not oracle.available(kind): continue
        t0 = time.time()
        r = oracle.load(kind).cellpop_evaluate(prob, vals, threads=16, want_cell_values=True, want_steps=True, want_average=True)
        dt = time.time() - t0
        rel = np.abs(logp - r["logp"]) / np.abs(r["logp"])
        m = ~np.isnan(r["cell_values"])
        nanmatch = (np.isnan(d["cell_values"]) == np.isnan(r["cell_values"])).all()
        cv = np.abs(d["cell_values"][m] - r["cell_values"][m])
        print(f"  vs {kind} ({dt:.2f}s cpu): max rel logp {rel.max():.2e} abs {np.abs(logp-r['logp']).max():.2e}; steps identical {(d['cell_steps']==r['cell_steps']).mean():.3f}; "
              f"cell values max abs diff {cv.max():.2e}; avg max abs diff {np.abs(d['population_average']-r['population_average']).max():.2e}; nan match {nanmatch}")
    ev.close()

if __name__ == "__main__":
    ks = sys.argv[1].split(",") if len(sys.argv) > 1 else ["group"]
    for k in ks:
        run(12, 200, 20, 3, kernel=k)
        run(5, 300, 20, 3, kernel=k)
        run(12, 2000, 50, 8, kernel=k, kinds=("port",))
        run(24, 500, 20, 2, decades=3.0, kernel=k)
        run(12, 10000, 50, 16, kinds=("ref",), kernel=k)
