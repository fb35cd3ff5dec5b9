"""CPU suite for the cell_population rows: the plain-C restatement against golden vectors made by
the reference's own compiled ODESolverCVODE, the compiled reference against its own fixtures, the synthetic model generator,
and the device translation unit of a generated model (cross-compiled, not run)."""
import dataclasses
import os

import numpy as np
import pytest

from bcm3_b200 import synthetic_cellpop as sc
from tests.util import CELLPOP_GOLDEN_NAMES, assert_logp_parity, cellpop_step_match_floor, load_cellpop_golden


def test_cellpop_golden_fixtures_exist():
    assert len(CELLPOP_GOLDEN_NAMES) >= 4


@pytest.mark.parametrize("name", CELLPOP_GOLDEN_NAMES)
def test_port_matches_reference_golden(port, name):
    prob, gold = load_cellpop_golden(name)
    r = port.cellpop_evaluate(prob, gold["values"], threads=2, want_cell_values=True, want_steps=True, want_average=True)
    # pure relative error per chain, bounded by the north-star 1e-6 or the reference's own measured reproducibility
    assert_logp_parity(r["logp"], gold["logp"], gold["noise_floor"], name)
    assert (np.isnan(r["cell_values"]) == np.isnan(gold["cell_values"])).all()
    m = np.isfinite(gold["cell_values"])  # a log ratio is -inf where both species still are 0
    assert (np.isfinite(r["cell_values"]) == m).all()
    # single trajectories: tolerance level (rtol = atol = 4.8e-7 per step, a few hundred steps); a log ratio divides by small numbers
    assert np.abs(r["cell_values"][m] - gold["cell_values"][m]).max() < (5e-4 if prob.log_ratio_denominator is not None else 5e-5)
    avg_ok = np.isfinite(gold["population_average"])
    assert np.abs(r["population_average"][avg_ok] - gold["population_average"][avg_ok]).max() < (5e-4 if prob.log_ratio_denominator is not None else 5e-6)
    # step counts: identical for most cells; the stiff 24-species case flips more decisions
    same = (r["cell_steps"] == gold["cell_steps"]).mean()
    assert same >= cellpop_step_match_floor(gold)
    assert abs(r["cell_steps"].mean() / gold["cell_steps"].mean() - 1.0) < 0.02


@pytest.mark.parametrize("name", CELLPOP_GOLDEN_NAMES)
def test_compiled_reference_reproduces_golden(ref, name):
    prob, gold = load_cellpop_golden(name)
    r = ref.cellpop_evaluate(prob, gold["values"], threads=1, want_cell_values=True, want_steps=True)
    assert np.array_equal(r["logp"], gold["logp"])
    assert np.array_equal(r["cell_steps"], gold["cell_steps"])
    assert np.array_equal(r["cell_values"], gold["cell_values"], equal_nan=True)


def test_late_entry_cells_are_absent_before_entry(port):
    prob, gold = load_cellpop_golden("cellpop_n5_late_entry")
    before = prob.timepoints < prob.entry_time
    assert before.any()
    assert np.isnan(gold["cell_values"][:, before, :]).all() and not np.isnan(gold["cell_values"][:, ~before, :]).any()
    assert np.all(gold["population_average"][:, before] == 0.0)  # nothing is notified for those timepoints


def test_failed_cell_gives_minus_infinity(port):
    prob, gold = load_cellpop_golden("cellpop_n12_normal")
    p2 = dataclasses.replace(prob, solver_max_steps=20)  # ODESolverCVODE.cpp:440-446 => Simulate fails => -inf (Experiment.cpp:356-358)
    r = port.cellpop_evaluate(p2, gold["values"])
    assert np.all(r["logp"] == -np.inf)


def test_generated_text_matches_generator_conventions():
    code = sc.cascade_code(6, seed=3)
    assert code.startswith(sc.SIGNATURE)
    assert "OdeReal ratelaws[12];" in code and "out[5] = +ratelaws[10]-ratelaws[11];" in code
    assert "EXPORT_PREFIX void generated_jacobian(OdeMatrixReal& out" in code
    # std::to_string: every literal has exactly six decimals
    import re

    for lit in re.findall(r"\d+\.\d+", code):
        assert len(lit.split(".")[1]) == 6


def test_device_module_compiles_for_sm100a(built, tmp_path, monkeypatch):
    """The per-model translation unit (prelude + generated text + cellpop_warp.cuh) cross-compiles with nvcc for sm_100a."""
    from bcm3_b200.cellpop import CellPopEvaluator

    monkeypatch.setenv("BCM3B200_CACHE", str(tmp_path))
    prob, _ = load_cellpop_golden("cellpop_n5_late_entry")
    ev = CellPopEvaluator(prob, compile_only=True)
    ev.close()
    built_dirs = [d for d in os.listdir(tmp_path) if d.startswith("cellpop_")]
    assert len(built_dirs) == 1
    assert os.path.exists(tmp_path / built_dirs[0] / "libcellpop_model.so")
    src = open(tmp_path / built_dirs[0] / "model.cu").read()
    assert "#define CP_N 5" in src and "template <class OUT, class SP, class CS, class PP, class NS>" in src and "generated_jacobian" not in src
    # a text without the generator's signature is rejected
    bad = dataclasses.replace(prob, derivative_code="void f() {}")
    from bcm3_b200 import _lib

    with pytest.raises(_lib.Bcm3B200Error, match="generated_derivative signature"):
        CellPopEvaluator(bad, compile_only=True)


def test_lane_parallel_rhs_text_is_bit_identical_on_the_host(built, tmp_path, monkeypatch):
    """The library regroups the generated statements by expression shape so that the lanes of a cell's group evaluate
    different reactions at the same time (cellpop_host.cuh::cellpop_lane_rhs). Here the regrouped text of a stiff 33-species
    cascade (all rate-law shapes, stoichiometric coefficients added) is compiled for the HOST next to the original text, the
    "lanes" run one after the other, and both must give the same BITS on random states."""
    import ctypes
    import subprocess

    from bcm3_b200.cellpop import CellPopEvaluator

    monkeypatch.setenv("BCM3B200_CACHE", str(tmp_path))
    prob = sc.make_cellpop_problem(N=33, num_cells=8, T=6, data_cells=2, seed=73, rate_decades=3.0)
    # give two species a sum with coefficients and a leading minus, and one an empty sum, as the generator can emit them
    code = prob.derivative_code.replace("out[5] = +ratelaws[10]-ratelaws[11];", "out[5] = -2.000000*ratelaws[11]+ratelaws[10]+0.500000*ratelaws[3];")
    code = code.replace("out[7] = +ratelaws[14]-ratelaws[15];", "out[7] = 0.0;")
    assert code != prob.derivative_code
    prob = dataclasses.replace(prob, derivative_code=code)
    CellPopEvaluator(prob, compile_only=True).close()
    (d,) = [d for d in os.listdir(tmp_path) if d.startswith("cellpop_")]
    src = open(tmp_path / d / "model.cu").read()
    assert "#define CP_RHS_LANES 1" in src and "#define CP_NUM_RATELAWS 66" in src
    lanes = src[src.index("// ---- lane-parallel form"):src.index('#include "cellpop_group.cuh"')]
    here = os.path.dirname(os.path.abspath(__file__))
    harness = ('#include <cmath>\n#include <limits>\n#include <vector>\n#include "cellpop_prelude.h"\n#define EXPORT_PREFIX extern "C"\n'
               "struct OdeMatrixReal { double dummy; double& operator()(int, int) { return dummy; } };\n" + code +
               "\n#define __device__\n#define __forceinline__ inline\ntemplate <class T> static inline T __ldg(const T* p) { return *p; }\n" + lanes +
               """
template <int G> static void run(double* out, const double* y, const double* cs, const double* p, const double* ns, int N)
{
	std::vector<double> rl(CP_NUM_RATELAWS, std::numeric_limits<double>::quiet_NaN());
	for (int lg = 0; lg < G; lg++) generated_ratelaws_lanes<G>(lg, rl.data(), y, cs, p, ns);
	for (int i = 0; i < N; i++) out[i] = generated_assemble(i, rl.data());
}
extern "C" void lanes_eval(int G, double* out, const double* y, const double* cs, const double* p, const double* ns, int N)
{
	if (G == 2) run<2>(out, y, cs, p, ns, N); else if (G == 16) run<16>(out, y, cs, p, ns, N); else run<32>(out, y, cs, p, ns, N);
}
""")
    (tmp_path / "harness.cpp").write_text(harness)
    so = str(tmp_path / "harness.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++14", "-fPIC", "-shared", "-w", "-I", os.path.join(os.path.dirname(here), "oracle"),
                    "-o", so, str(tmp_path / "harness.cpp")], check=True)
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(3)
    N = 33
    for trial in range(50):
        y = rng.uniform(-0.05, 1.2, N) if trial % 3 else rng.uniform(-1.0, 3.0, N)  # includes the helpers' negative / saturated branches
        cs = np.array([rng.uniform(0.0, 2.0)])
        p = 10.0 ** rng.uniform(-1, 1, 6)
        ns = np.zeros(1)
        want = np.full(N, np.nan)
        lib.generated_derivative(want.ctypes.data_as(dp), y.ctypes.data_as(dp), cs.ctypes.data_as(dp), p.ctypes.data_as(dp), ns.ctypes.data_as(dp))
        for G in (2, 16, 32):
            got = np.full(N, np.nan)
            lib.lanes_eval(G, got.ctypes.data_as(dp), y.ctypes.data_as(dp), cs.ctypes.data_as(dp), p.ctypes.data_as(dp), ns.ctypes.data_as(dp), N)
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (trial, G)
    assert want[7] == 0.0


def test_unrecognised_generated_text_keeps_the_scalar_form(built, tmp_path, monkeypatch):
    """A statement the regrouping does not understand (here: a rate law that reads another rate law) makes it give up; the model
    is then compiled from the text as it stands."""
    from bcm3_b200.cellpop import CellPopEvaluator

    monkeypatch.setenv("BCM3B200_CACHE", str(tmp_path))
    prob, _ = load_cellpop_golden("cellpop_n5_late_entry")
    code = prob.derivative_code.replace("\tratelaws[3] = ", "\tratelaws[3] = 0.0*ratelaws[2]+")
    assert code != prob.derivative_code
    CellPopEvaluator(dataclasses.replace(prob, derivative_code=code), compile_only=True).close()
    (d,) = [d for d in os.listdir(tmp_path) if d.startswith("cellpop_")]
    src = open(tmp_path / d / "model.cu").read()
    assert "CP_RHS_LANES 1" not in src and "generated_ratelaws_lanes" not in src


# ---- the per-cell time_course likelihood's matching (bcm3b200_match_cells: host code, no device needed) ----

def _matching_cases():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matching_cases.npz"))
    costs, matches, at_c, at_m = z["costs"], z["matches"], 0, 0
    for n in z["sizes"]:
        n = int(n)
        yield costs[at_c:at_c + n * n].reshape(n, n), matches[at_m:at_m + n]
        at_c += n * n
        at_m += n


def test_matching_reproduces_the_reference_matchings():
    """The product's restatement of the reference's Hungarian implementation (bcm3_b200/csrc/matching_host.cuh) against matchings
    the reference's own compiled function returned (tests/golden/make_golden_matching.py) -- 90 matrices, a fifth of which the
    reference does NOT match optimally: the matching is part of the likelihood's value, so it is the reference's that counts."""
    from bcm3_b200 import _lib

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matching_cases.npz"))
    assert int(z["suboptimal"]) > 0  # the fixture does exercise the quirk
    count = 0
    for cost, want in _matching_cases():
        got = _lib.match_cells(cost)
        assert got is not None and np.array_equal(got, want)
        assert sorted(got.tolist()) == list(range(len(got)))  # a perfect matching
        count += 1
    assert count == 90


def test_matching_against_the_compiled_reference_on_fresh_matrices(ref):
    from bcm3_b200 import _lib

    rng = np.random.default_rng(7)
    for trial in range(150):
        n = int(rng.integers(1, 60))
        cost = rng.normal(0.0, rng.uniform(0.2, 50.0), (n, n)) + rng.uniform(-500.0, 500.0)
        if trial % 3 == 0:
            cost = np.round(cost)
        assert np.array_equal(_lib.match_cells(cost), ref.hungarian_match(cost))


def test_time_course_likelihood_is_the_matched_sum_not_the_optimum(ref):
    """DataLikelihoodTimeCourse::Evaluate sums the cell likelihoods of the matching its Hungarian call returns (.cpp:323-336); on
    this fixture that is below the optimal assignment's sum -- the checker and the golden carry the reference's value."""
    from scipy.optimize import linear_sum_assignment

    prob, gold = load_cellpop_golden("cellpop_time_course_n8_normal")
    assert prob.data_kind == "time_course" and prob.observed.shape == (prob.num_cells, prob.num_timepoints)
    sim = gold["cell_values"]  # [C][T][cells]
    n, sd = prob.num_cells, prob.stdev
    below = 0
    for c in range(sim.shape[0]):
        d = prob.observed[:, :, None] - sim[c][None, :, :]                      # [observed][T][simulated]
        lik = (-np.log(sd) - 0.9189385332046727 - d * d / (2 * sd * sd)).sum(axis=1)
        match = ref.hungarian_match(-lik)
        assert abs(lik[np.arange(n), match].sum() * prob.weight - gold["logp"][c]) <= 1e-9 * abs(gold["logp"][c])
        r, col = linear_sum_assignment(-lik)
        assert lik[r, col].sum() >= lik[np.arange(n), match].sum() - 1e-9
        below += lik[r, col].sum() > lik[np.arange(n), match].sum() + 1e-6
    assert below > 0


def test_matching_edge_cases():
    from bcm3_b200 import _lib

    assert np.array_equal(_lib.match_cells(np.array([[3.5]])), [0])
    assert np.array_equal(_lib.match_cells(np.array([[-1e9]])), [0])
    two = _lib.match_cells(np.array([[0.0, 5.0], [5.0, 0.0]]))
    assert np.array_equal(two, [0, 1])
    # all entries equal: every matching is optimal, the restatement returns the reference's (the identity: greedy over the tight edges)
    assert np.array_equal(_lib.match_cells(np.zeros((7, 7))), np.arange(7))
    # costs that differ by less than 1 are all "tight" for the reference's integer test: the greedy first pass decides
    # -- the smallest example of the quirk: the reference's compiled function answers [0, 1] (cost 1.1) where [1, 0] costs 0.1
    c = np.array([[0.2, 0.0], [0.1, 0.9]])
    assert np.array_equal(_lib.match_cells(c), [0, 1])


def test_variable_indices_of_the_data_likelihood_are_validated():
    """stdev_ix & co. are read on the device as transformed[chain][ix]: an index past the variables is refused at finalize."""
    from bcm3_b200._lib import Bcm3B200Error
    from bcm3_b200.cellpop import CellPopEvaluator

    prob = sc.make_cellpop_problem(N=6, num_cells=4, T=6, data_cells=2, seed=3)
    for field in ("stdev_ix", "offset_ix", "scale_ix"):
        with pytest.raises(Bcm3B200Error, match="out of range"):
            CellPopEvaluator(dataclasses.replace(prob, **{field: prob.num_variables}), compile_only=True).close()
    tc = sc.make_time_course_problem(N=6, num_cells=4, T=6, seed=3, extra_marker_species=((2,),))
    tc.extra_markers[0].scale_ix = 50
    with pytest.raises(Bcm3B200Error, match="out of range"):
        CellPopEvaluator(tc, compile_only=True).close()
