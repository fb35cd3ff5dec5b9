"""Development check (run under gpurun): PopPK model variants against their golden fixtures."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcm3_b200.poppk import PopPKEvaluator
from tests.util import load_golden, rel_err
for name in ("poppk_one_biphasic", "poppk_two_biphasic", "poppk_one_transit", "poppk_two_transit"):
    prob, gold = load_golden(name)
    ev = PopPKEvaluator(prob, diagnostics=True)
    logp, st = ev.evaluate(gold["values"])
    d = ev.diagnostics()
    ev.close()
    same = (d["counters"].astype(np.int64) == gold["counters"]).all(axis=-1)
    pl = rel_err(d["patient_ll"], gold["patient_ll"])
    print(name, "logp", logp, "gold", gold["logp"], "rel", rel_err(logp, gold["logp"]).max(), "counters same", same.mean(), "worst patient rel", pl.max())
    w = np.unravel_index(np.argmax(pl), pl.shape)
    print("   worst", w, "gpu", d["patient_ll"][w], "gold", gold["patient_ll"][w], "counters gpu", d["counters"][w], "gold", gold["counters"][w],
          "interval", prob.trial.dosing_interval[w[1]], "dose", prob.trial.dose[w[1]], "interm", prob.trial.intermittent[w[1]])

import oracle
from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_ONE_BIPHASIC, PK_ONE_TRANSIT
np.set_printoptions(precision=6, linewidth=200)
for pk in (PK_ONE_BIPHASIC, PK_ONE_TRANSIT):
    prob = syn.make_poppk_problem(pk, P=2, T=12, t_end=60.0, seed=3)
    vals = syn.make_chain_values(prob, 1)
    ev = PopPKEvaluator(prob, diagnostics=True)
    logp, st = ev.evaluate(vals)
    d = ev.diagnostics()
    ev.close()
    r = oracle.load("port").poppk_evaluate(prob, vals, want_conc=True, want_counters=True)
    print("pk", pk, "values head", vals[0, :10])
    print(" gpu conc", d["conc"][0, 0])
    print(" ref conc", r["conc"][0, 0])
    print(" counters gpu", d["counters"][0, 0], "ref", r["counters"][0, 0])
