"""Generates tests/golden/pharmaco_*.npz with the reference's own compartment model (oracle/_ref: src/pharmaco/
PharmacokineticModel.cpp + Eigen's matrix exponential, compiled in place) behind the restated population glue.
Run where /root/reference is mounted:  python tests/golden/make_golden_pharmaco.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from bcm3_b200 import pharmaco as ph  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "pharmaco_plain": dict(P=48, T=10, seed=3),
    "pharmaco_peripheral": dict(P=48, T=10, peripheral=True, seed=4),
    "pharmaco_transit3_bioavailability": dict(P=40, T=12, num_transit=3, bioavailability=True, seed=5),
    "pharmaco_peripheral_transit4": dict(P=32, T=10, peripheral=True, num_transit=4, seed=6),
    # two transit compartments: the reference links the chain only for more than two (PharmacokineticModel.cpp:215)
    "pharmaco_transit2_quirk": dict(P=24, T=8, num_transit=2, seed=7),
}
# likelihood.xml type="pharmaco_single" (src/pharmaco/PharmacoLikelihoodSingle.cpp): one patient, the variables are its rates
SINGLE_CASES = {
    "pharmaco_single_plain": dict(seed=3),
    "pharmaco_single_peripheral_no_excretion": dict(peripheral=True, excretion=False, seed=4),
    "pharmaco_single_transit3": dict(num_transit=3, seed=5),
    "pharmaco_single_biphasic": dict(biphasic_absorption=True, seed=6),
    "pharmaco_single_metabolite": dict(metabolite=True, seed=7),
    "pharmaco_single_everything": dict(peripheral=True, num_transit=3, biphasic_absorption=True, metabolite=True, seed=8),
}


def main():
    ref = oracle.load("ref")
    for name, kw in list(CASES.items()) + list(SINGLE_CASES.items()):
        if name in SINGLE_CASES:
            prob = ph.make_pharmaco_single_problem(**kw)
            vals = ph.make_pharmaco_single_values(prob, 6, seed=100 + kw["seed"])
        else:
            prob = ph.make_pharmaco_problem(**kw)
            vals = ph.make_pharmaco_values(prob, 3, seed=100 + kw["seed"])
        r = ref.pharmaco_evaluate(prob, vals, threads=1, want_conc=True, want_patient_ll=True)
        tr = prob.trial
        out = dict(drug=np.array(tr.drug), time=tr.time, observed_concentration=tr.observed_concentration, dose=tr.dose,
                   dosing_interval=tr.dosing_interval, dose_after_dose_change=tr.dose_after_dose_change, dose_change_time=tr.dose_change_time,
                   intermittent=tr.intermittent, treatment_interruptions=tr.treatment_interruptions, variable_names=np.array(prob.variable_names),
                   transforms=prob.transforms, peripheral_compartment=np.array(prob.peripheral_compartment),
                   num_transit_compartments=np.array(prob.num_transit_compartments), bioavailability=np.array(prob.bioavailability),
                   single=np.array(prob.single), biphasic_absorption=np.array(prob.biphasic_absorption), metabolite=np.array(prob.metabolite),
                   values=vals, logp=r["logp"], conc=r["conc"], patient_ll=r["patient_ll"])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "logp", r["logp"], os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
