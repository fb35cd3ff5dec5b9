"""Development check (run under gpurun): GPU path vs the CPU checkers on seeded problems, with timings."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_ONE, PK_TWO
from bcm3_b200.poppk import PopPKEvaluator

def compare(pk, het, P=1000, C=4, kinds=("ref", "port")):
    prob = syn.make_poppk_problem(pk, P=P, T=10, t_end=72.0, heterogeneous=het, missing_fraction=0.1 if het else 0.0)
    vals = syn.make_chain_values(prob, C)
    ev = PopPKEvaluator(prob, diagnostics=True)
    t0 = time.time(); logp, status = ev.evaluate(vals); dt = time.time() - t0
    d = ev.diagnostics()
    print(f"pk={pk} het={het} P={P} C={C}: gpu logp[:4]={logp[:4]} status={status[:4]} ({dt*1e3:.1f} ms, kernel {ev.get_stat('last_kernel_us')} us)")
    for kind in kinds:
        if not oracle.available(kind):
            print("  oracle", kind, "not available"); continue
        r = oracle.load(kind).poppk_evaluate(prob, vals, threads=8, want_conc=True, want_counters=True, want_patient_ll=True)
        rel = np.abs(logp - r["logp"]) / np.abs(r["logp"])
        cg, co = d["counters"].astype(np.int64), r["counters"]
        same = (cg == co).all(axis=2).mean()
        steps_same = (cg[..., 0] == co[..., 0]).mean()
        m = ~np.isnan(r["conc"])
        nanmatch = (np.isnan(d["conc"]) == np.isnan(r["conc"])).all()
        crel = np.abs(d["conc"][m] - r["conc"][m]) / np.maximum(np.abs(r["conc"][m]), 1e-300)
        print(f"  vs {kind}: max rel logp err {rel.max():.3e}; counters identical {same:.4f}; steps identical {steps_same:.4f}; "
              f"conc max rel {crel.max():.2e}, frac>1e-7 {(crel>1e-7).mean():.4f}; nan pattern match {nanmatch}")
        print("     mean counters gpu", cg.reshape(-1, 8).mean(0).round(3))
        print("     mean counters cpu", co.reshape(-1, 8).mean(0).round(3))
    ev.close()

if __name__ == "__main__":
    for pk in (PK_ONE, PK_TWO):
        for het in (False, True):
            compare(pk, het)
    # timing at larger sizes (no diagnostics)
    for pk, P, C in ((PK_ONE, 1000, 16), (PK_ONE, 100000, 16), (PK_TWO, 100000, 64)):
        prob = syn.make_poppk_problem(pk, P=P, T=10, t_end=72.0)
        vals = syn.make_chain_values(prob, C)
        ev = PopPKEvaluator(prob)
        for it in range(3):
            t0 = time.time(); logp, status = ev.evaluate(vals); dt = time.time() - t0
            print(f"timing pk={pk} P={P} C={C}: e2e {dt*1e3:.1f} ms, kernel {ev.get_stat('last_kernel_us')/1e3:.2f} ms -> {C/dt:.1f} evals/s; logp[0]={logp[0]:.6f}")
        ev.close()
