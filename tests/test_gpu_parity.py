"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the golden vectors of the compiled reference,
against the CPU checker on fresh seeded inputs, and through size-independent properties at BASELINE.json's full sizes.

The bar (BASELINE.json north_star): relative error <= 1e-6 on every per-chain log-likelihood at matched rtol/atol."""
import numpy as np
import pytest

from bcm3_b200 import synthetic as syn
from bcm3_b200.poppk_data import PK_ONE, PK_TWO, PopPKProblem
from tests.util import GOLDEN_NAMES, assert_matches_golden, counter_match_floor, load_golden, rel_err

pytestmark = pytest.mark.gpu

LOGP_RTOL = 1e-6


@pytest.fixture(scope="module")
def Evaluator(built):
    from bcm3_b200 import _lib
    from bcm3_b200.poppk import PopPKEvaluator

    assert _lib.device_count() > 0, "no CUDA device: the GPU suite must run on the B200 box"
    return PopPKEvaluator


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_matches_reference_golden(Evaluator, name):
    prob, gold = load_golden(name)
    ev = Evaluator(prob, diagnostics=True)
    logp, status = ev.evaluate(gold["values"])
    d = ev.diagnostics()
    ev.close()
    assert (status == 0).all()
    assert_matches_golden(logp, d["conc"], d["counters"], gold, logp_tol=LOGP_RTOL, min_counter_match=counter_match_floor(name))
    assert (np.isneginf(d["patient_ll"]) == np.isneginf(gold["patient_ll"])).all()
    m = ~np.isnan(gold["conc"])
    rel = rel_err(d["conc"][m], gold["conc"][m])
    assert np.median(rel) < 1e-9
    assert (rel > 1e-6).mean() < 1.5 * (1.0 - counter_match_floor(name))


@pytest.mark.parametrize("pk,het,seed", [(PK_ONE, False, 101), (PK_ONE, True, 102), (PK_TWO, False, 103), (PK_TWO, True, 104)])
def test_matches_cpu_checker_on_seeded_inputs(Evaluator, checker, pk, het, seed):
    prob = syn.make_poppk_problem(pk, P=777, T=9, t_end=96.0, heterogeneous=het, missing_fraction=0.1 if het else 0.0, seed=seed)
    vals = syn.make_chain_values(prob, 5, seed=seed)
    ev = Evaluator(prob, diagnostics=True)
    logp, status = ev.evaluate(vals)
    d = ev.diagnostics()
    ev.close()
    want = checker.poppk_evaluate(prob, vals, threads=4, want_counters=True, want_patient_ll=True)
    assert rel_err(logp, want["logp"]).max() <= LOGP_RTOL
    same = (d["counters"].astype(np.int64) == want["counters"]).all(axis=2).mean()
    assert same >= 0.97
    # the sum the kernels reduce equals the sum of the per-patient terms they report
    assert rel_err(logp, d["patient_ll"].sum(axis=1)).max() < 1e-12


@pytest.mark.parametrize("block", [32, 64, 128, 256, 384])
def test_block_size_does_not_change_results(Evaluator, block):
    prob, gold = load_golden("poppk_two_hetero")
    ev = Evaluator(prob, block_size=block)
    logp, _ = ev.evaluate(gold["values"])
    ev.close()
    assert rel_err(logp, gold["logp"]).max() <= LOGP_RTOL


def test_edge_cases(Evaluator, checker):
    # empty trial, single patient, ragged sizes that do not fill a warp, timepoint at t = 0, all-missing observations
    for P in (0, 1, 31, 33):
        prob = syn.make_poppk_problem(PK_ONE, P=P, T=5, t_end=48.0, seed=7)
        if P > 1:
            prob.trial.time[0] = 0.0
            prob.trial.observed_concentration[1, :] = np.nan
            prob = PopPKProblem(pk_type=prob.pk_type, trial=prob.trial, transforms=prob.transforms, sd_ix=prob.sd_ix)
        vals = syn.make_chain_values(prob, 3, seed=P)
        ev = Evaluator(prob)
        logp, status = ev.evaluate(vals)
        ev.close()
        want = checker.poppk_evaluate(prob, vals)["logp"]
        assert rel_err(logp, want).max() <= LOGP_RTOL, P
        assert (status == 0).all()


def test_nan_and_minus_infinity_semantics(Evaluator, checker):
    """A failed solve gives -inf (cpp:400-408); a NaN log-likelihood is flagged (Sampler.cpp:172-178); a NaN that the
    reference's serial loop never reaches (it breaks at the first -inf, cpp:438) must not surface."""
    from tests.util import make_nan_inf_case

    prob, vals = make_nan_inf_case()
    want = checker.poppk_evaluate(prob, vals)["logp"]
    assert want[0] == -np.inf and want[1] == -np.inf and np.isnan(want[2])
    for block in (32, 64):
        ev = Evaluator(prob, block_size=block)
        logp, status = ev.evaluate(vals)
        ev.close()
        assert logp[0] == -np.inf and logp[1] == -np.inf and np.isnan(logp[2])
        assert status.tolist() == [0, 0, 1]


def test_chain_independence_and_determinism(Evaluator):
    prob = syn.make_poppk_problem(PK_TWO, P=500, T=10, t_end=72.0, seed=21)
    vals = syn.make_chain_values(prob, 6, seed=21)
    ev = Evaluator(prob, block_size=64)
    a, _ = ev.evaluate(vals)
    b, _ = ev.evaluate(vals)
    single = np.array([ev.evaluate(vals[c:c + 1])[0][0] for c in range(6)])
    rev, _ = ev.evaluate(vals[::-1].copy())
    ev.close()
    assert np.array_equal(a, b)                # run-to-run bit-identical (fixed reduction order)
    assert np.array_equal(a, single)           # a chain's result does not depend on its batch
    assert np.array_equal(a, rev[::-1])


def test_shards_add_up(Evaluator):
    """Contiguous patient shards (what each rank of a multi-GPU run owns) recombine to the unsharded result."""
    from bcm3_b200.parallel import combine_partials

    prob = syn.make_poppk_problem(PK_ONE, P=1000, T=10, t_end=72.0, heterogeneous=True, seed=31)
    vals = syn.make_chain_values(prob, 4, seed=31)
    full = Evaluator(prob)
    want, _ = full.evaluate(vals)
    full.close()
    import torch

    total = None
    for r in range(3):
        ev = Evaluator(prob, shard_rank=r, shard_count=3)
        d_vals = torch.from_numpy(vals).cuda()
        d_partial = torch.empty((3, 4), dtype=torch.float64, device="cuda")
        ev.evaluate_device(d_vals.data_ptr(), 4, prob.num_variables, d_partial.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        p = d_partial.cpu().numpy()
        # the host-values entry must give the same partial as the device-values entry
        pinned = torch.from_numpy(vals).pin_memory()
        d_partial2 = torch.empty_like(d_partial)
        ev.enqueue(pinned.data_ptr(), 4, prob.num_variables, d_partial2.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(p, d_partial2.cpu().numpy())
        ev.close()
        total = p if total is None else np.stack([total[0] + p[0], np.minimum(total[1], p[1]), np.minimum(total[2], p[2])])
    got, status = combine_partials(total)
    assert rel_err(got, want).max() < 1e-13


@pytest.mark.parametrize("pk,P,C", [(PK_ONE, 1000, 16), (PK_TWO, 100000, 64)])
def test_full_size_properties(Evaluator, checker, pk, P, C):
    """BASELINE.json configs 2 and 5 at full size: (i) a random subsample of patients, evaluated alone by the CPU
    checker, reproduces the per-patient terms; (ii) permuting the patients permutes nothing but the summation order;
    (iii) the simulated concentrations stay within CVODE's own accuracy of the exact solution of the linear model."""
    prob = syn.make_poppk_problem(pk, P=P, T=10, t_end=72.0, seed=1)
    vals = syn.make_chain_values(prob, C)
    diag = P * C <= 2_000_000
    ev = Evaluator(prob, diagnostics=diag)
    logp, status = ev.evaluate(vals)
    assert (status == 0).all() and np.isfinite(logp).all()

    # (ii) permutation of patients (observations, dosing data and per-patient variables move together)
    rng = np.random.default_rng(0)
    perm = rng.permutation(P)
    tr = prob.trial
    npk = 4 if pk == PK_ONE else 6
    tr2 = type(tr)(drug=tr.drug, time=tr.time, observed_concentration=tr.observed_concentration[perm], dose=tr.dose[perm],
                   dosing_interval=tr.dosing_interval[perm], dose_after_dose_change=tr.dose_after_dose_change[perm],
                   dose_change_time=tr.dose_change_time[perm], intermittent=tr.intermittent[perm],
                   treatment_interruptions=tr.treatment_interruptions[perm])
    prob2 = PopPKProblem(pk_type=pk, trial=tr2, transforms=prob.transforms, sd_ix=prob.sd_ix)
    vals2 = vals.copy()
    pp = vals[:, npk + 2:npk + 2 + 2 * P].reshape(C, P, 2)
    vals2[:, npk + 2:npk + 2 + 2 * P] = pp[:, perm, :].reshape(C, 2 * P)
    ev2 = Evaluator(prob2, diagnostics=diag)  # same kernel instantiation as `ev`
    logp2, _ = ev2.evaluate(vals2)
    ev2.close()
    assert rel_err(logp2, logp).max() < 1e-11

    # (i) subsample against the CPU checker
    sub = np.sort(rng.choice(P, size=min(P, 400), replace=False))
    trs = type(tr)(drug=tr.drug, time=tr.time, observed_concentration=tr.observed_concentration[sub], dose=tr.dose[sub],
                   dosing_interval=tr.dosing_interval[sub], dose_after_dose_change=tr.dose_after_dose_change[sub],
                   dose_change_time=tr.dose_change_time[sub], intermittent=tr.intermittent[sub],
                   treatment_interruptions=tr.treatment_interruptions[sub])
    ns = len(sub)
    nvs = npk + 2 * (ns + 1) + 2
    probs = PopPKProblem(pk_type=pk, trial=trs, transforms=syn.poppk_transforms(pk, ns), sd_ix=nvs - 2)
    # the tolerances depend on the minimum dose of the trial, identical here (constant dose)
    assert probs.atol == prob.atol
    cs = [0, C - 1]
    vs = np.empty((len(cs), nvs))
    vs[:, :npk + 2] = vals[cs, :npk + 2]
    vs[:, npk + 2:npk + 2 + 2 * ns] = pp[cs][:, sub, :].reshape(len(cs), 2 * ns)
    vs[:, nvs - 2:] = vals[cs, -2:]
    want = checker.poppk_evaluate(probs, vs, threads=4, want_patient_ll=True)
    evs = Evaluator(probs)
    got, _ = evs.evaluate(vs)
    evs.close()
    assert rel_err(got, want["logp"]).max() <= LOGP_RTOL
    if diag:
        d = ev.diagnostics()
        assert np.abs(d["patient_ll"][cs][:, sub] - want["patient_ll"]).max() < 1e-2
        assert rel_err(d["patient_ll"][cs][:, sub].sum(axis=1), want["logp"]).max() <= LOGP_RTOL
    ev.close()

    # (iv) WHOLE chains at full size against the reference: the first and the last chain over all P patients (at config 5
    # the compiled reference needs ~10 s per chain on one host thread)
    whole = checker.poppk_evaluate(prob, vals[cs], threads=2)["logp"]
    assert rel_err(logp[cs], whole).max() <= LOGP_RTOL


@pytest.mark.parametrize("name", ["poppk_two_hetero", "poppk_one_hetero", "poppk_one_maxsteps"])
def test_ranking_patients_by_absorption_rate_does_not_change_results(Evaluator, name):
    """Large batches integrate each chain's patients in order of ka (rank kernel + radix sort) so that the lanes of a warp
    take similar numbers of steps; the per-patient results are the same, only the summation order changes."""
    prob, gold = load_golden(name)
    out = {}
    for flag in (True, False):
        ev = Evaluator(prob, sort_patients=flag, diagnostics=True)
        logp, _ = ev.evaluate(gold["values"])
        out[flag] = (logp, ev.diagnostics()["patient_ll"])
        ev.close()
    assert rel_err(out[True][0], gold["logp"]).max() <= LOGP_RTOL
    assert np.array_equal(out[True][1], out[False][1], equal_nan=True)   # per-patient terms bit-identical
    assert rel_err(out[True][0], out[False][0]).max() < 1e-12


def test_ranked_large_batch_matches_the_cpu_checker(Evaluator, checker):
    """Above the library's own threshold (P * C >= 60 000 systems) the ranking is on by default."""
    prob = syn.make_poppk_problem(PK_TWO, P=7200, T=10, t_end=72.0, seed=33, heterogeneous=True, missing_fraction=0.05)
    vals = syn.make_chain_values(prob, 16, seed=33)
    ev = Evaluator(prob)
    logp, status = ev.evaluate(vals)
    launches = ev.get_stat("last_kernel_launches")
    # a chain evaluated alone falls below the ranking threshold and runs with another block size: its log-likelihood must
    # still have the same BITS (patient terms are summed in patient order whatever the launch shape), or Metropolis-Hastings
    # decisions could differ between a batched and a chain-by-chain run
    single = np.array([ev.evaluate(vals[c:c + 1])[0][0] for c in (0, 7, 15)])
    ev.close()
    assert np.array_equal(single, logp[[0, 7, 15]])
    want = checker.poppk_evaluate(prob, vals, threads=8)["logp"]
    assert launches == 3  # rank kernel + integrator + chain reduce (the sort passes are the library's)
    assert (status == 0).all() and rel_err(logp, want).max() <= LOGP_RTOL


def test_device_entry_with_odd_row_stride(Evaluator, checker):
    """bcm3b200_evaluate_batch_device on a caller's [C][nvar] block: for the biphasic models nvar = 2 P + 11 is odd, so every
    second chain's per-patient block is only 8-byte aligned (and so is a block that starts at an odd double): the kernel must
    not assume 16-byte alignment there."""
    import torch

    prob, gold = load_golden("poppk_two_biphasic")
    vals = np.ascontiguousarray(gold["values"])
    C, nvar = vals.shape
    assert nvar % 2 == 1
    ev = Evaluator(prob)
    want, _ = ev.evaluate(vals)
    stream = torch.cuda.current_stream().cuda_stream
    for shift in (0, 1):  # block at a 16-byte boundary / 8 bytes off it
        flat = torch.zeros(C * nvar + 2, dtype=torch.float64, device="cuda")
        d_vals = flat[shift:shift + C * nvar]
        d_vals.copy_(torch.from_numpy(vals).reshape(-1))
        assert (d_vals.data_ptr() % 16 == 0) == (shift == 0)
        d_partial = torch.empty((3, C), dtype=torch.float64, device="cuda")
        ev.evaluate_device(d_vals.data_ptr(), C, nvar, d_partial.data_ptr(), stream)
        torch.cuda.synchronize()
        got, status = ev.combine_partials(d_partial.cpu().numpy())
        assert np.array_equal(got, want) and (status == 0).all()
    ev.close()
    assert rel_err(want, gold["logp"]).max() <= LOGP_RTOL


def test_two_batches_in_flight_on_one_handle(Evaluator):
    """bcm3b200_enqueue_batch does not synchronise: a second batch enqueued while the first is still queued behind its
    kernels must not disturb it (the library stages nothing of its own between the caller's buffer and the device)."""
    import torch

    prob = syn.make_poppk_problem(PK_TWO, P=3000, T=10, t_end=72.0, seed=41)
    a = syn.make_chain_values(prob, 8, seed=41)
    b = syn.make_chain_values(prob, 8, seed=42)
    ev = Evaluator(prob)
    want_a, _ = ev.evaluate(a)
    want_b, _ = ev.evaluate(b)
    ha, hb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
    pa = torch.empty((3, 8), dtype=torch.float64, device="cuda")
    pb = torch.empty((3, 8), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    ev.enqueue(ha.data_ptr(), 8, a.shape[1], pa.data_ptr(), stream)
    ev.enqueue(hb.data_ptr(), 8, b.shape[1], pb.data_ptr(), stream)
    torch.cuda.synchronize()
    got_a, _ = ev.combine_partials(pa.cpu().numpy())
    got_b, _ = ev.combine_partials(pb.cpu().numpy())
    ev.close()
    assert np.array_equal(got_a, want_a) and np.array_equal(got_b, want_b) and not np.array_equal(want_a, want_b)


def test_both_biphasic_type_strings_select_the_two_compartment_model(Evaluator):
    """LikelihoodPopPKTrajectory.cpp:73-76: type="one_biphasic_uptake" selects PKMT_TwoCompartmentBiphasicUptake, exactly
    like "two_biphasic_uptake" -- the same likelihood.xml must give the same log-likelihood here."""
    prob, gold = load_golden("poppk_two_biphasic")
    out = []
    for type_string in ("one_biphasic_uptake", "two_biphasic_uptake"):
        ev = Evaluator(prob, type_string=type_string)
        out.append(ev.evaluate(gold["values"])[0])
        ev.close()
    assert np.array_equal(out[0], out[1])
    assert rel_err(out[0], gold["logp"]).max() <= LOGP_RTOL


@pytest.mark.parametrize("fixed", [dict(fixed_vod=45.0), dict(fixed_periphery_fwd=0.3, fixed_periphery_bwd=0.08), dict(fixed_vod=60.0, fixed_periphery_bwd=0.1)])
def test_fixed_pk_model_attributes(Evaluator, checker, fixed):
    """<pk_model volume_of_distribution= k_periphery_fwd= k_periphery_bwd=> (cpp:64-67, 122-130, 285-294): each fixed attribute
    shortens the prior by one variable while the vector is still read at the all-sampled positions; k_periphery_bwd alone
    counts as fixed but is not used (cpp:288 tests the forward rate only). Checked against the compiled reference glue."""
    base = syn.make_poppk_problem(PK_TWO, P=300, T=8, t_end=72.0, heterogeneous=True, seed=51)
    vals = syn.make_chain_values(base, 4, seed=51)
    n = len(fixed)
    prob = PopPKProblem(pk_type=PK_TWO, trial=base.trial, transforms=base.transforms[:-n], sd_ix=base.num_variables - n - 2, **fixed)
    vals = np.ascontiguousarray(vals[:, :-n])
    ev = Evaluator(prob)
    logp, status = ev.evaluate(vals)
    ev.close()
    want = checker.poppk_evaluate(prob, vals, threads=4)["logp"]
    assert np.isfinite(want).all() and (status == 0).all()
    assert rel_err(logp, want).max() <= LOGP_RTOL
    plain = checker.poppk_evaluate(base, syn.make_chain_values(base, 4, seed=51), threads=4)["logp"]  # nothing fixed
    assert not np.allclose(plain, want, rtol=1e-3)


# ---- pharmacokinetic_trajectory: the likelihood of one patient, a batch = one ODE system per chain ----
from tests.util import SINGLE_GOLDEN_NAMES  # noqa: E402


@pytest.mark.parametrize("name", SINGLE_GOLDEN_NAMES)
def test_single_patient_matches_reference_golden(Evaluator, name):
    """likelihood.xml type="pharmacokinetic_trajectory" (LikelihoodPharmacokineticTrajectory.cpp:259-352) on the same kernel:
    the chain's variables are the patient's rates, every timepoint is simulated; goldens from the compiled reference."""
    prob, gold = load_golden(name)
    ev = Evaluator(prob, diagnostics=True)
    logp, status = ev.evaluate(gold["values"])
    d = ev.diagnostics()
    ev.close()
    assert (status == 0).all()
    assert (np.isneginf(logp) == np.isneginf(gold["logp"])).all()
    assert rel_err(logp, gold["logp"]).max() < LOGP_RTOL
    assert (d["counters"] == gold["counters"]).all(axis=-1).mean() >= 0.8  # 5 systems: at most one may differ in a counter
    m = ~np.isnan(gold["conc"]) & np.isfinite(gold["logp"])[:, None, None]
    assert np.median(rel_err(d["conc"][m], gold["conc"][m])) < 1e-9


@pytest.mark.parametrize("pk", [PK_ONE, PK_TWO])
def test_single_patient_fresh_inputs_and_nan(Evaluator, checker, pk):
    """64 chains on a fresh patient against the checker. A NaN rate makes the solver fail: -inf, as in the reference; a NaN
    standard deviation gives a NaN log-likelihood (status 1) that the reference's sampler turns into an error (Sampler.cpp:172-178)."""
    prob = syn.make_single_patient_problem(pk, seed=77 + pk, T=16, t_end=144.0)
    vals = syn.make_single_patient_values(prob, 64, seed=5)
    vals[3, 2] = np.nan
    vals[5, 8] = np.nan
    ev = Evaluator(prob)
    logp, status = ev.evaluate(vals)
    ev.close()
    want = checker.poppk_evaluate(prob, vals, threads=4)["logp"]
    assert np.isneginf(logp[3]) and np.isneginf(want[3]) and status[3] == 0
    assert np.isnan(logp[5]) and np.isnan(want[5]) and status[5] == 1
    ok = ~np.isin(np.arange(64), (3, 5))
    assert (status[ok] == 0).all() and rel_err(logp[ok], want[ok]).max() < LOGP_RTOL


def test_single_patient_refuses_a_population(Evaluator):
    import dataclasses

    from bcm3_b200._lib import Bcm3B200Error

    pop = syn.make_poppk_problem(PK_ONE, P=2, T=6)
    prob = syn.make_single_patient_problem(PK_ONE)
    prob.trial = dataclasses.replace(pop.trial)
    with pytest.raises(Bcm3B200Error, match="ONE patient"):
        Evaluator(prob)
