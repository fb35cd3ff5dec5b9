"""CPU suite for the pharmaco_population path: the plain-C++ restatement (oracle/pharmaco_port.cpp, own matrix exponential)
against the golden vectors of the reference's compiled compartment model, and the compiled reference against its own goldens."""
import numpy as np
import pytest

from tests.util import PHARMACO_GOLDEN_NAMES, load_pharmaco_golden, rel_err


@pytest.mark.parametrize("name", PHARMACO_GOLDEN_NAMES)
def test_port_matches_reference_golden(port, name):
    prob, gold = load_pharmaco_golden(name)
    r = port.pharmaco_evaluate(prob, gold["values"], threads=2, want_conc=True, want_patient_ll=True)
    assert rel_err(r["logp"], gold["logp"]).max() < 1e-12
    assert (np.isnan(r["conc"]) == np.isnan(gold["conc"])).all()
    m = ~np.isnan(gold["conc"])
    assert np.abs(r["conc"][m] - gold["conc"][m]).max() <= 1e-11 * np.abs(gold["conc"][m]).max()
    assert np.abs(r["patient_ll"] - gold["patient_ll"]).max() < 1e-9


@pytest.mark.parametrize("name", PHARMACO_GOLDEN_NAMES)
def test_reference_reproduces_its_golden(ref, name):
    prob, gold = load_pharmaco_golden(name)
    r = ref.pharmaco_evaluate(prob, gold["values"], threads=1, want_conc=True)
    assert np.array_equal(r["logp"], gold["logp"]) and np.array_equal(r["conc"], gold["conc"], equal_nan=True)


def test_transit_chain_quirk_is_kept(port):
    """PharmacokineticModel.cpp:215: the transit compartments are linked only when there are MORE than two of them -- with two,
    nothing reaches the central compartment and every simulated concentration is zero. Restated as it is."""
    prob, gold = load_pharmaco_golden("pharmaco_transit2_quirk")
    m = ~np.isnan(gold["conc"])
    assert (gold["conc"][m] == 0.0).all()
