"""GPU suite for the cell_population path: the CUDA integrators (compiled per model from the generated RHS text; the lane-group
mapping is the default, the one-cell-per-warp and one-cell-per-thread mappings are kept selectable) through the C ABI, against
the golden vectors of the compiled reference and against the CPU checker on fresh inputs."""
import dataclasses
import os

import numpy as np
import pytest

from bcm3_b200 import synthetic_cellpop as sc
import oracle
from tests.util import (CELLPOP_GOLDEN_NAMES, _strict_oracle, assert_logp_parity, cellpop_step_match_floor, load_cellpop_golden,
                        reference_noise_floor_cellpop, rel_err)

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
    # the bar: PURE relative error <= 1e-6 per chain, or twice the reference's own measured irreproducibility on this fixture
    # (tests/golden/measure_noise_floor.py) where that is larger
    assert_logp_parity(logp, gold["logp"], gold["noise_floor"], name)
    if prob.log_ratio_denominator is not None:
        # the device reports numerator and denominator rows; the checker's per-cell value is their log ratio (and its "population
        # average" the average of that, which no likelihood uses)
        with np.errstate(all="ignore"):
            den = d["marker_values"][-1]
            d["cell_values"] = 0.4342944819032518 * np.where(den < 1e-16, np.log(d["cell_values"] / 1e-16), np.log(d["cell_values"] / den))
        fin = np.isfinite(gold["cell_values"])
        assert np.abs(d["cell_values"][fin] - gold["cell_values"][fin]).max() < 5e-4
    else:
        assert (np.isnan(d["cell_values"]) == np.isnan(gold["cell_values"])).all()
        m = ~np.isnan(gold["cell_values"])
        # single trajectories at tolerance level; a daughter starts from its parent's state at the step that crossed the division
        # threshold (Cell.cpp:507-512, no interpolation without stored integration points): a step that differs in its last bits
        # moves that start
        assert np.abs(d["cell_values"][m] - gold["cell_values"][m]).max() < (5e-4 if prob.divide_cells else 5e-5)
        assert np.abs(d["population_average"] - gold["population_average"]).max() < (5e-5 if prob.divide_cells else 5e-6)
    assert abs(d["cell_steps"].mean() / gold["cell_steps"].mean() - 1.0) < 0.02
    assert (d["cell_steps"] == gold["cell_steps"]).mean() >= cellpop_step_match_floor(gold)


def _fresh_reference(checker, prob, vals, threads=8):
    """(reference result, per-chain noise floor, step-match fraction of the reference's own builds) for fresh inputs; with only
    the plain-C port present (no compiled reference) the floor is unknown and the bar is the plain 1e-6."""
    m = reference_noise_floor_cellpop(prob, vals, threads=threads)
    if m is not None:
        return m
    return checker.cellpop_evaluate(prob, vals, threads=threads, want_average=True, want_steps=True, want_cell_values=True), None, 0.5


def test_config3_shape_against_cpu_checker(Evaluator, checker):
    """BASELINE config 3 shape at a size the checker finishes in seconds: 12 species, 2 000 cells, 50 timepoints, 8 chains."""
    prob = sc.make_cellpop_problem(N=12, num_cells=2000, T=50, data_cells=16, seed=5)
    vals = sc.make_chain_values(8, seed=5)
    ev = Evaluator(prob)
    logp, status = ev.evaluate(vals)
    again, _ = ev.evaluate(vals)
    d = ev.diagnostics()
    ev.close()
    want, floor, _ = _fresh_reference(checker, prob, vals)
    assert np.array_equal(logp, again)  # deterministic
    assert_logp_parity(logp, want["logp"], floor, "config-3 shape")
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


def test_sharded_cells_reproduce_the_unsharded_result(Evaluator):
    """Two handles owning half of the cells each (what two ranks hold), partials summed as the all-reduce would, the
    finish step on each: same logp as the single handle, to round-off (sum / count versus sum of x / count)."""
    import torch

    prob, gold = load_cellpop_golden("cellpop_n5_late_entry")
    vals = np.ascontiguousarray(gold["values"])
    C = vals.shape[0]
    ev = Evaluator(prob)
    want, _ = ev.evaluate(vals)
    ev.close()
    shards = [Evaluator(prob, shard_rank=r, shard_count=2) for r in range(2)]
    width = shards[0].get_stat("partial_doubles_per_chain")
    assert width == 2 * prob.num_timepoints + 1
    assert shards[0].get_stat("num_cells_local") + shards[1].get_stat("num_cells_local") == prob.num_cells
    stream = torch.cuda.current_stream().cuda_stream
    parts = []
    for s in shards:
        d = torch.empty((C, width), dtype=torch.float64, device="cuda:0")
        s.enqueue(vals.ctypes.data, C, vals.shape[1], d.data_ptr(), stream)
        parts.append(d)
    torch.cuda.synchronize()
    total = parts[0] + parts[1]
    for s in shards:
        logp, status = s.finish(total.data_ptr(), C, stream)
        assert (status == 0).all()
        assert np.abs(logp - want).max() <= 1e-9 * np.maximum(np.abs(want), 1.0).max()
    with pytest.raises(Exception):
        shards[0].evaluate(vals)  # a sharded handle only yields partials
    for s in shards:
        s.close()


@pytest.mark.parametrize("N,decades", [(3, 2.0), (7, 2.0), (16, 3.0), (33, 3.0), (50, 4.0)])
def test_lane_group_shapes_against_cpu_checker(Evaluator, checker, N, decades):
    """The lane-group mapping at sizes that exercise every shape: 2, 4 (padded), 8, 16 and 32 lanes per cell, with and
    without block lock-step, inlined and called rate-law helpers -- against the compiled reference, with the tolerance the
    reference's own reproducibility on the very same inputs allows (measured here with its two builds)."""
    prob = sc.make_cellpop_problem(N=N, num_cells=96, T=12, data_cells=4, seed=40 + N, rate_decades=decades)
    vals = sc.make_chain_values(2, seed=N)
    ev = Evaluator(prob)
    logp, status = ev.evaluate(vals)
    d = ev.diagnostics()
    ev.close()
    want, floor, _ = _fresh_reference(checker, prob, vals, threads=4)
    assert (status == 0).all() and (d["cell_status"] == 1).all()
    assert_logp_parity(logp, want["logp"], floor, f"N={N}")
    assert np.abs(d["population_average"] - want["population_average"]).max() < 2e-5
    assert abs(d["cell_steps"].mean() / want["cell_steps"].mean() - 1.0) < 0.02


@pytest.mark.parametrize("case", ["cellpop_sbml_cell_cycle", "cellpop_n24_stiff", "fresh_n33"])
def test_lane_parallel_rhs_is_bit_identical(Evaluator, case):
    """The regrouped right-hand side (cellpop_host.cuh::cellpop_lane_rhs: lanes evaluate different reactions of one shape at the
    same time, species sums assembled term by term in the order of the text) performs the same IEEE operations on the same
    operands as the generated text evaluated as it stands: every trajectory value and every step count must be the same bits.
    Includes the fixture whose text came out of the reference's own generator (21 shapes for 24 reactions)."""
    if case == "fresh_n33":
        prob = sc.make_cellpop_problem(N=33, num_cells=64, T=12, data_cells=4, seed=73, rate_decades=3.0)
        vals = sc.make_chain_values(2, seed=33)
    else:
        prob, gold = load_cellpop_golden(case)
        vals = gold["values"]
    out = []
    for lanes in (False, "always"):
        ev = Evaluator(prob, rhs_lanes=lanes)
        logp, _ = ev.evaluate(vals)
        d = ev.diagnostics()
        ev.close()
        out.append((logp, d))
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1]["cell_values"], out[1][1]["cell_values"], equal_nan=True)
    assert np.array_equal(out[0][1]["cell_steps"], out[1][1]["cell_steps"])


def test_config3_full_size_properties(Evaluator, checker):
    """BASELINE config 3 at full size (12 species, 10 000 cells, 50 timepoints, 16 chains), through properties that do not
    need the checker to run the whole thing: (1) the first 192 cells' trajectories equal the checker's, (2) reversing the
    order of the cells (rows of the quasi-random table) leaves every log-likelihood unchanged to round-off, (3) a chain's
    result does not depend on the batch it is evaluated in, (4) two shards combine to the unsharded result."""
    import dataclasses

    import torch

    prob = sc.make_cellpop_problem(N=12, num_cells=10_000, T=50, data_cells=32, seed=1)
    vals = sc.make_chain_values(16)
    ev = Evaluator(prob)
    logp, status = ev.evaluate(vals)
    d = ev.diagnostics()
    single, _ = ev.evaluate(vals[5:6])
    ev.close()
    assert (status == 0).all() and (d["cell_status"] == 1).all() and np.isfinite(logp).all()
    assert single[0] == logp[5]
    # (1)
    sub = dataclasses.replace(prob, num_cells=192, sobol=prob.sobol[:192])
    want = checker.cellpop_evaluate(sub, vals[:4], threads=4, want_cell_values=True, want_steps=True)
    got = d["cell_values"][:4, :, :192]
    assert (np.isnan(got) == np.isnan(want["cell_values"])).all()
    m = ~np.isnan(got)
    assert np.abs(got[m] - want["cell_values"][m]).max() < 5e-5
    assert (d["cell_steps"][:4, :192] == want["cell_steps"]).mean() > 0.7
    # (2)
    ev = Evaluator(dataclasses.replace(prob, sobol=prob.sobol[::-1].copy()))
    rev, _ = ev.evaluate(vals)
    ev.close()
    assert rel_err(rev, logp).max() < 1e-9
    # (4)
    shards = [Evaluator(prob, shard_rank=r, shard_count=2) for r in range(2)]
    width = 2 * prob.num_timepoints + 1
    stream = torch.cuda.current_stream().cuda_stream
    parts = []
    for s in shards:
        part = torch.empty((16, width), dtype=torch.float64, device="cuda:0")
        s.enqueue(vals.ctypes.data, 16, vals.shape[1], part.data_ptr(), stream)
        parts.append(part)
    torch.cuda.synchronize()
    total = parts[0] + parts[1]
    combined, _ = shards[0].finish(total.data_ptr(), 16, stream)
    for s in shards:
        s.close()
    assert rel_err(combined, logp).max() < 1e-9


def test_config4_full_size_against_the_reference(Evaluator, checker):
    """BASELINE.json config 4 as named: ~50-species stiff network, 100 000 cells, 16 chains. (1) ONE WHOLE CHAIN at full size
    -- every one of the 100 000 cells integrated by the compiled reference on the host cores -- against the GPU's per-chain
    log-likelihood and population average; (2) the first 200 cells' trajectories of two chains; (3) the whole 16-chain batch
    runs, and a chain's result does not depend on the batch it is evaluated in."""
    import dataclasses

    prob = sc.make_cellpop_problem(N=50, num_cells=100_000, T=50, data_cells=32, seed=1, rate_decades=4.0)
    vals = sc.make_chain_values(16)
    threads = max(1, os.cpu_count() or 1)
    ev = Evaluator(prob)
    one, status = ev.evaluate(vals[3:4])
    d = ev.diagnostics()
    assert status[0] == 0 and (d["cell_status"] == 1).all()
    # (1): the floor is the reference's own movement under its compiler flags and 1-ulp inputs, on this very input
    m = reference_noise_floor_cellpop(prob, vals[3:4], threads=threads)
    if m is None:
        want, floor = checker.cellpop_evaluate(prob, vals[3:4], threads=threads, want_average=True, want_steps=True), None
    else:
        want, floor, _ = m
    assert_logp_parity(one, want["logp"], floor, "config 4, one whole chain")
    assert np.abs(d["population_average"] - want["population_average"]).max() < 1e-6
    assert abs(d["cell_steps"].mean() / want["cell_steps"].mean() - 1.0) < 0.005
    # (2)
    sub = dataclasses.replace(prob, num_cells=200, sobol=prob.sobol[:200])
    ws = checker.cellpop_evaluate(sub, vals[[3]], threads=threads, want_cell_values=True)
    got = d["cell_values"][:, :, :200]
    assert (np.isnan(got) == np.isnan(ws["cell_values"])).all()
    mk = ~np.isnan(got)
    # single trajectories of a stiff 50-species network move by a few 1e-5 between the reference's own builds (step-size decisions
    # flip on round-off; measured 5.4e-5 between GPU and reference on one of these 200 cells): the bound is three times what the
    # reference's -ffp-contract=off build differs from it on the same cells, and 5e-5 at least
    bound = 5e-5
    strict = os.path.join(os.path.dirname(oracle.REF_LIB), "libbcm3ref_strict.so")
    if oracle.available("ref") and os.path.exists(strict):
        wb = _strict_oracle(strict).cellpop_evaluate(sub, vals[[3]], threads=threads, want_cell_values=True)
        bound = max(bound, 3.0 * np.abs(wb["cell_values"][mk] - ws["cell_values"][mk]).max())
    assert np.abs(got[mk] - ws["cell_values"][mk]).max() < bound
    assert np.median(np.abs(got[mk] - ws["cell_values"][mk])) < 1e-7
    # (3)
    logp, status = ev.evaluate(vals)
    ev.close()
    assert (status == 0).all() and np.isfinite(logp).all()
    assert logp[3] == one[0]


def test_stdev_relative_to_scale(Evaluator, checker):
    """<data stdev_relative_to_scale="true"> (DataLikelihoodBase.cpp:151-153): the standard deviation is multiplied by the
    data scale -- equal to the run that is given the product directly, and equal to the CPU checker."""
    import dataclasses

    base = sc.make_cellpop_problem(N=8, num_cells=96, T=10, data_cells=4, seed=9)
    vals = sc.make_chain_values(3, seed=9)
    prob = dataclasses.replace(base, scale=1.7, offset=0.02, stdev_ix=None, stdev=0.2, stdev_relative_to_scale=True)
    ev = Evaluator(prob)
    got, _ = ev.evaluate(vals)
    ev.close()
    ev = Evaluator(dataclasses.replace(prob, stdev=0.2 * 1.7, stdev_relative_to_scale=False))
    same, _ = ev.evaluate(vals)
    ev.close()
    ev = Evaluator(dataclasses.replace(prob, stdev_relative_to_scale=False))
    other, _ = ev.evaluate(vals)
    ev.close()
    assert np.allclose(got, same, rtol=1e-13, atol=0) and not np.allclose(got, other, rtol=1e-3)
    want, floor, _ = _fresh_reference(checker, prob, vals, threads=2)
    assert_logp_parity(got, want["logp"], floor, "stdev_relative_to_scale")


def test_entry_time_variability_takes_a_dimension_and_nothing_else(Evaluator, checker):
    """<variable entry_time=...> of a cell_variability block: the reference reads it and gives it a quasi-random dimension
    but never applies it (VariabilityDescription::ApplyVariabilityEntryTime has no caller) -- the result equals the run
    without that variable on the remaining columns of the table."""
    import dataclasses
    from bcm3_b200.cellpop_data import Variability

    base = sc.make_cellpop_problem(N=8, num_cells=96, T=10, data_cells=4, seed=9)
    vals = sc.make_chain_values(3, seed=9)
    rng = np.random.default_rng(1)
    extra = rng.uniform(0.05, 0.95, size=(base.num_cells, 1))
    sobol = np.concatenate([base.sobol[:, :1], extra, base.sobol[:, 1:]], axis=1)
    variability = [base.variability[0], Variability(apply="additive", entry_time=True, scale_fixed=0.3)] + list(base.variability[1:])
    prob = dataclasses.replace(base, sobol=sobol, variability=variability)
    for kernel in ("auto", "warp", "thread"):
        ev = Evaluator(prob, kernel=kernel)
        got, _ = ev.evaluate(vals)
        ev.close()
        ev = Evaluator(base, kernel=kernel)
        want, _ = ev.evaluate(vals)
        ev.close()
        assert np.array_equal(got, want), kernel
    assert np.array_equal(checker.cellpop_evaluate(prob, vals)["logp"], checker.cellpop_evaluate(base, vals)["logp"])


def test_dividing_population_against_the_reference(Evaluator, checker):
    """<experiment divide_cells="true">: fresh inputs against the compiled reference -- which cells exist at which timepoint, the
    number of cells ever created per chain, the population average; and the reference's failure when the population outgrows
    max_cells (CellPopulation::AddNewCell returns no cell => the chain's log-likelihood is -inf): same chains fail."""
    import dataclasses

    prob = sc.make_dividing_problem(M=6, num_cells=12, max_cells=400, t_end=5.0, T=14, seed=11)
    vals = sc.make_chain_values(4, seed=21)
    want = checker.cellpop_evaluate(prob, vals, threads=4, want_cell_values=True, want_steps=True, want_average=True)
    ev = Evaluator(prob)
    got, status = ev.evaluate(vals)
    d = ev.diagnostics()
    ev.close()
    assert d["cell_values"].shape == want["cell_values"].shape == (4, 14, 400)
    assert (np.isnan(d["cell_values"]) == np.isnan(want["cell_values"])).all()
    assert ((d["cell_steps"] > 0).sum(axis=1) == (want["cell_steps"] > 0).sum(axis=1)).all()
    assert ((want["cell_steps"] > 0).sum(axis=1) > 3 * 12).all()  # at least two generations of daughters everywhere
    assert np.abs(d["population_average"] - want["population_average"]).max() < 1e-6
    assert_logp_parity(got, want["logp"], None, "dividing population")
    # the same population with room for 40 cells only
    small = dataclasses.replace(prob, max_cells=40)
    want = checker.cellpop_evaluate(small, vals, threads=4)
    ev = Evaluator(small)
    got, status = ev.evaluate(vals)
    ev.close()
    assert np.isneginf(want["logp"]).all() and np.isneginf(got).all() and (status == 0).all()


def test_solver_max_timestep(Evaluator, checker):
    """<experiment solver_max_timestep=> -> CVodeSetMaxStep (Experiment.cpp:413, Cell.cpp:73, cvode.c:1121-1122, 3142-3143): the
    step-size ceiling changes the step sequence (more steps) and, at round-off level, the result -- as in the reference."""
    import dataclasses

    base = sc.make_cellpop_problem(N=8, num_cells=96, T=10, data_cells=4, seed=9)
    vals = sc.make_chain_values(3, seed=9)
    prob = dataclasses.replace(base, solver_max_timestep=0.05)
    want, floor, _ = _fresh_reference(checker, prob, vals, threads=4)
    free = checker.cellpop_evaluate(base, vals, threads=4, want_steps=True)
    assert want["cell_steps"].mean() > 1.15 * free["cell_steps"].mean()  # the ceiling binds
    ev = Evaluator(prob)
    got, status = ev.evaluate(vals)
    d = ev.diagnostics()
    ev.close()
    assert (status == 0).all()
    assert_logp_parity(got, want["logp"], floor, "solver_max_timestep")
    assert abs(d["cell_steps"].mean() / want["cell_steps"].mean() - 1.0) < 0.01


# ---- <data type="time_course">: per-cell trajectories, every observed cell matched to one simulated cell ----

def test_time_course_fresh_problem_against_the_reference(Evaluator, checker):
    """160 cells with an observed trajectory each (missing values, Student-t error model): the [observed x simulated] block of
    cell log-likelihoods comes from the device, the matching from the host restatement of the reference's Hungarian call; the
    checker runs the reference's own compiled solver and (oracle/_ref) its own compiled matching."""
    prob = dataclasses.replace(sc.make_time_course_problem(N=8, num_cells=160, T=14, seed=51, missing_fraction=0.1), error_model="student_t4",
                               scale=1.05, offset=-0.01, weight=0.7)
    vals = sc.make_chain_values(5, seed=51)
    ev = Evaluator(prob)
    logp, status = ev.evaluate(vals)
    again, _ = ev.evaluate(vals)
    ev.close()
    want, floor, _ = _fresh_reference(checker, prob, vals)
    assert (status == 0).all() and np.isfinite(logp).all()
    assert np.array_equal(logp, again)
    assert_logp_parity(logp, want["logp"], floor, "time_course, 160 cells")


def test_time_course_failed_cell_gives_minus_infinity(Evaluator):
    prob = dataclasses.replace(sc.make_time_course_problem(N=6, num_cells=16, T=8, seed=52), solver_max_steps=15)
    ev = Evaluator(prob)
    logp, _ = ev.evaluate(sc.make_chain_values(2, seed=52))
    ev.close()
    assert np.all(logp == -np.inf)  # Simulate() fails => the experiment's likelihood is -inf (Experiment.cpp:356-358)


def test_time_course_refuses_what_is_not_built(Evaluator):
    from bcm3_b200._lib import Bcm3B200Error

    prob = sc.make_time_course_problem(N=6, num_cells=16, T=8, seed=53)
    with pytest.raises(Bcm3B200Error):  # fewer observed than simulated cells: the reference refuses it too (DataLikelihoodTimeCourse.cpp:178-187)
        Evaluator(dataclasses.replace(prob, observed=prob.observed[:10])).close()
    with pytest.raises(Bcm3B200Error):  # every observed cell is compared with every simulated cell: not split over ranks
        Evaluator(prob, shard_rank=0, shard_count=2).close()


# ---- <data type="time_points">: at every timepoint its own set of observed cells ----

def test_time_points_fresh_problem_against_the_reference(Evaluator, checker):
    """120 simulated cells, 48-120 observed cells per timepoint (rectangular matchings), values relative to timepoint 1."""
    prob = dataclasses.replace(sc.make_time_points_problem(N=8, num_cells=120, T=10, seed=55, relative_to=1), weight=1.3, scale=0.9, stdev=0.02)
    vals = sc.make_chain_values(4, seed=55)
    ev = Evaluator(prob)
    logp, status = ev.evaluate(vals)
    again, _ = ev.evaluate(vals)
    ev.close()
    want, floor, _ = _fresh_reference(checker, prob, vals)
    assert (status == 0).all() and np.isfinite(logp).all()
    assert np.array_equal(logp, again)
    assert_logp_parity(logp, want["logp"], floor, "time_points, 120 cells")


def test_time_points_without_enough_simulated_cells_is_minus_infinity(Evaluator, checker):
    """Cells that enter at t = 1: at the timepoints before that no simulated cell has a value while the data has cells:
    DataLikelihoodTimePoints.cpp:241-245 gives -inf."""
    prob = dataclasses.replace(sc.make_time_points_problem(N=6, num_cells=16, T=8, seed=56), entry_time=1.5)
    assert (prob.timepoints < 1.5).any()
    vals = sc.make_chain_values(2, seed=56)
    ev = Evaluator(prob)
    logp, _ = ev.evaluate(vals)
    ev.close()
    assert np.all(logp == -np.inf)
    assert np.all(checker.cellpop_evaluate(prob, vals)["logp"] == -np.inf)
    # with the early observations removed the same problem is finite and agrees
    obs = prob.observed.copy()
    obs[:, prob.timepoints < 1.5] = np.nan
    p2 = dataclasses.replace(prob, observed=obs)
    ev = Evaluator(p2)
    logp, _ = ev.evaluate(vals)
    ev.close()
    want, floor, _ = _fresh_reference(checker, p2, vals)
    assert np.isfinite(logp).all()
    assert_logp_parity(logp, want["logp"], floor, "time_points, late entry")


@pytest.mark.parametrize("name", ["cellpop_time_course_n6_t4_missing", "cellpop_time_points_n8_normal"])
def test_per_cell_blocks_in_chunks_of_chains_give_the_same_bits(Evaluator, name, monkeypatch):
    """The [observed x simulated] blocks go to the host a bounded number of chains at a time (1 GiB); one chain at a time here."""
    prob, gold = load_cellpop_golden(name)
    ev = Evaluator(prob)
    whole, _ = ev.evaluate(gold["values"])
    monkeypatch.setenv("BCM3B200_CELL_LIKELIHOOD_DOUBLES", str(prob.num_cells * prob.num_cells))
    chunked, _ = ev.evaluate(gold["values"])
    ev.close()
    assert np.array_equal(whole, chunked)
    assert_logp_parity(chunked, gold["logp"], gold["noise_floor"], name)


def test_time_course_with_a_single_cell(Evaluator, checker):
    prob = sc.make_time_course_problem(N=6, num_cells=1, T=8, seed=58)
    vals = sc.make_chain_values(3, seed=58)
    ev = Evaluator(prob)
    logp, status = ev.evaluate(vals)
    ev.close()
    want, floor, _ = _fresh_reference(checker, prob, vals)
    assert (status == 0).all() and np.isfinite(logp).all()
    assert_logp_parity(logp, want["logp"], floor, "time_course, one cell")


def test_mitotic_only_average_on_non_dividing_cells(Evaluator, checker):
    """include_only_cells_that_went_through_mitosis without division: the event code of the kernel only records which cells'
    nuclear envelope species fell below 0.5 after some accepted step; the average runs over those cells."""
    base = sc.make_dividing_problem(M=5, num_cells=40, max_cells=40, t_end=1.5, T=12, seed=7)  # by t = 1.5 only some cells have entered mitosis
    prob = dataclasses.replace(base, divide_cells=False, max_cells=0, cytokinesis_species=None, division_reset_species=(),
                               sobol=base.sobol[:40], include_only_cells_that_went_through_mitosis=True, nuclear_envelope_species=6)
    vals = sc.make_chain_values(4, seed=9)
    ev = Evaluator(prob)
    logp, status = ev.evaluate(vals)
    d = ev.diagnostics()
    ev.close()
    want, floor, _ = _fresh_reference(checker, prob, vals)
    everyone = checker.cellpop_evaluate(dataclasses.replace(prob, include_only_cells_that_went_through_mitosis=False), vals)["logp"]
    assert (status == 0).all() and np.isfinite(logp).all()
    assert not np.allclose(want["logp"], everyone)  # some cells have not entered mitosis by the end: the two averages differ
    assert_logp_parity(logp, want["logp"], floor, "mitotic cells only")
    assert np.abs(d["population_average"] - want["population_average"]).max() < 5e-5
