"""GPU suite for the cell_population path: the CUDA integrators (compiled per model from the generated RHS text; the lane-group
mapping is the default, the one-cell-per-warp and one-cell-per-thread mappings are kept selectable) through the C ABI, against
the golden vectors of the compiled reference and against the CPU checker on fresh inputs."""
import numpy as np
import pytest

from bcm3_b200 import synthetic_cellpop as sc
from tests.util import CELLPOP_GOLDEN_NAMES, cellpop_logp_close, cellpop_rtol, load_cellpop_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Evaluator(built):
    from bcm3_b200 import _lib
    from bcm3_b200.cellpop import CellPopEvaluator

    assert _lib.device_count() > 0
    return CellPopEvaluator


@pytest.mark.parametrize("name", CELLPOP_GOLDEN_NAMES)
def test_matches_reference_golden(Evaluator, name):
    _check_golden(Evaluator, name, "auto")


@pytest.mark.parametrize("kernel", ["warp", "thread"])
def test_other_mappings_match_reference_golden(Evaluator, kernel):
    _check_golden(Evaluator, "cellpop_n12_normal", kernel)


def _check_golden(Evaluator, name, kernel):
    prob, gold = load_cellpop_golden(name)
    ev = Evaluator(prob, kernel=kernel)
    logp, status = ev.evaluate(gold["values"])
    d = ev.diagnostics()
    ev.close()
    assert (status == 0).all() and (d["cell_status"] == 1).all()
    assert cellpop_logp_close(logp, gold["logp"], prob.num_timepoints, prob.num_replicates, rtol=cellpop_rtol(name))
    assert (np.isnan(d["cell_values"]) == np.isnan(gold["cell_values"])).all()
    m = ~np.isnan(gold["cell_values"])
    assert np.abs(d["cell_values"][m] - gold["cell_values"][m]).max() < 5e-5
    assert np.abs(d["population_average"] - gold["population_average"]).max() < 5e-6
    assert abs(d["cell_steps"].mean() / gold["cell_steps"].mean() - 1.0) < 0.02
    assert (d["cell_steps"] == gold["cell_steps"]).mean() >= (0.02 if "stiff" in name else 0.7)


def test_config3_shape_against_cpu_checker(Evaluator, port):
    """BASELINE config 3 shape at a size the checker finishes in seconds: 12 species, 2 000 cells, 50 timepoints, 8 chains."""
    prob = sc.make_cellpop_problem(N=12, num_cells=2000, T=50, data_cells=16, seed=5)
    vals = sc.make_chain_values(8, seed=5)
    ev = Evaluator(prob)
    logp, status = ev.evaluate(vals)
    again, _ = ev.evaluate(vals)
    d = ev.diagnostics()
    ev.close()
    want = port.cellpop_evaluate(prob, vals, threads=8, want_average=True, want_steps=True)
    assert np.array_equal(logp, again)  # deterministic
    assert cellpop_logp_close(logp, want["logp"], 50, 1, rtol=1e-6)
    assert np.abs(d["population_average"] - want["population_average"]).max() < 1e-6
    assert abs(d["cell_steps"].mean() / want["cell_steps"].mean() - 1.0) < 0.01


def test_failed_cell_and_chain_independence(Evaluator):
    import dataclasses

    prob, gold = load_cellpop_golden("cellpop_n12_normal")
    ev = Evaluator(dataclasses.replace(prob, solver_max_steps=20))
    logp, status = ev.evaluate(gold["values"])
    ev.close()
    assert np.all(logp == -np.inf) and (status == 0).all()
    ev = Evaluator(prob)
    full, _ = ev.evaluate(gold["values"])
    single = np.array([ev.evaluate(gold["values"][c:c + 1])[0][0] for c in range(len(full))])
    ev.close()
    assert np.array_equal(full, single)
