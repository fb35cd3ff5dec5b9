import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_ONE, PK_TWO, PopPKProblem
from bcm3_b200.poppk import PopPKEvaluator
pk, P, C = PK_ONE, 1000, 16
prob = syn.make_poppk_problem(pk, P=P, T=10, t_end=72.0, seed=1)
vals = syn.make_chain_values(prob, C)
ev = PopPKEvaluator(prob, diagnostics=True)
logp, _ = ev.evaluate(vals); d = ev.diagnostics(); ev.close()
rng = np.random.default_rng(0)
perm = rng.permutation(P)
tr = prob.trial; npk = 4
tr2 = type(tr)(drug=tr.drug, time=tr.time, observed_concentration=tr.observed_concentration[perm], dose=tr.dose[perm],
               dosing_interval=tr.dosing_interval[perm], dose_after_dose_change=tr.dose_after_dose_change[perm],
               dose_change_time=tr.dose_change_time[perm], intermittent=tr.intermittent[perm],
               treatment_interruptions=tr.treatment_interruptions[perm])
prob2 = PopPKProblem(pk_type=pk, trial=tr2, transforms=prob.transforms, sd_ix=prob.sd_ix)
vals2 = vals.copy()
pp = vals[:, npk + 2:npk + 2 + 2 * P].reshape(C, P, 2)
vals2[:, npk + 2:npk + 2 + 2 * P] = pp[:, perm, :].reshape(C, 2 * P)
ev2 = PopPKEvaluator(prob2, diagnostics=True)
logp2, _ = ev2.evaluate(vals2); d2 = ev2.diagnostics(); ev2.close()
ll1 = d["patient_ll"][:, perm]; ll2 = d2["patient_ll"]
diff = np.abs(ll1 - ll2)
print("num differing patient_ll:", (diff > 0).sum(), "of", diff.size, "max", diff.max())
c1 = d["counters"][:, perm]; c2 = d2["counters"]
neq = (c1 != c2).any(axis=2)
print("counters differ:", neq.sum())
idx = np.argwhere(diff > 0)[:10]
for c, j in idx:
    print(c, j, "orig pos", perm[j], "lane", perm[j] % 32, "->", j % 32, ll1[c, j], ll2[c, j], c1[c, j], c2[c, j])
    print("   conc", d["conc"][c, perm[j]][:4], d2["conc"][c, j][:4])
