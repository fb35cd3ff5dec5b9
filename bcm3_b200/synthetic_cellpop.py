"""Synthetic cell_population workloads of the shapes BASELINE.json names (configs 3 and 4).

The reference ships no SBML model and no data, so a model is synthesised here directly as the TEXT the reference's code
generator would emit for it (src/sbml/SBMLModel.cpp:291-365: `ratelaws[r] = ...; out[i] = +ratelaws[..]-ratelaws[..];`,
numeric constants printed with std::to_string's 6 decimals, helper calls hill_function_fixedn*, michaelis_menten_function,
synthcap, tQSSA from the prelude of SolverCodeGenerator.cpp:122-295). The network is a signalling cascade with feedback:
species i is produced in a saturating way from species i-1 and degraded linearly, rate constants spread over several
decades so that the system is stiff. Nothing here touches the oracle or the GPU.
"""
from __future__ import annotations

import math
import re

import numpy as np

from .cellpop_data import CellPopProblem, Variability
from .poppk_data import TRANSFORM_LOG10, TRANSFORM_NONE

SIGNATURE = ("EXPORT_PREFIX void generated_derivative(OdeReal* out, const OdeReal* species, const OdeReal* constant_species, "
             "const OdeReal* parameters, const OdeReal* non_sampled_parameters)")

# sampled variables of the synthetic models (prior.xml order)
VAR_K_IN, VAR_K_CASCADE, VAR_K_DEG, VAR_K_FEEDBACK, VAR_VARIABILITY_SCALE, VAR_STDEV = range(6)
NUM_VARIABLES = 6


def _lit(x: float) -> str:
    """std::to_string(double): fixed, 6 decimals (SBMLRatelaws.cpp:757-763) -- the quirk that 1e-8 prints as 0.000000 is kept."""
    return f"{x:.6f}"


def cascade_code(N: int, seed: int = 0, rate_decades: float = 2.0):
    """Returns (text, info) for an N-species cascade. Reaction 2i produces species i, reaction 2i+1 degrades it."""
    rng = np.random.default_rng(seed)
    lines = [SIGNATURE, "{", f"\tOdeReal ratelaws[{2 * N}];"]
    for i in range(N):
        speed = 10.0 ** rng.uniform(-rate_decades / 2, rate_decades / 2)  # time scale of this level
        kscale = 4.0 * speed                                               # production gain over degradation ~ 4
        K = rng.uniform(0.15, 0.35)
        if i == 0:
            prod = f"((parameters[{VAR_K_IN}]*constant_species[0])*synthcap(species[0]))"
        else:
            kind = i % 4
            if kind == 1:
                prod = f"(((parameters[{VAR_K_CASCADE}]*{_lit(kscale)})*species[{i - 1}])*(1.000000-species[{i}]))"
            elif kind == 2:
                prod = f"michaelis_menten_function((parameters[{VAR_K_CASCADE}]*{_lit(kscale)}),{_lit(K)},species[{i - 1}],(1.000000-species[{i}]))"
            elif kind == 3:
                prod = f"(((parameters[{VAR_K_CASCADE}]*{_lit(kscale)})*hill_function_fixedn2(species[{i - 1}],{_lit(K)}))*(1.000000-species[{i}]))"
            else:
                prod = f"tQSSA((parameters[{VAR_K_CASCADE}]*{_lit(kscale)}),{_lit(K)},species[{i - 1}],(1.000000-species[{i}]))"
        lines.append(f"\tratelaws[{2 * i}] = {prod};")
        dscale = speed
        if i == 0 and N > 2:
            # negative feedback from the end of the cascade onto the first species
            deg = f"(((parameters[{VAR_K_DEG}]*{_lit(dscale)})*species[{i}])*(1.000000+(parameters[{VAR_K_FEEDBACK}]*species[{N - 1}])))"
        else:
            deg = f"((parameters[{VAR_K_DEG}]*{_lit(dscale)})*species[{i}])"
        lines.append(f"\tratelaws[{2 * i + 1}] = {deg};")
    for i in range(N):
        lines.append(f"\tout[{i}] = +ratelaws[{2 * i}]-ratelaws[{2 * i + 1}];")
    lines.append("}")
    lines.append("")
    # the generator also emits generated_jacobian(OdeMatrixReal& out, ...): present in real input, never used (Cell.cpp:57-76)
    lines.append("EXPORT_PREFIX void generated_jacobian(OdeMatrixReal& out, const OdeReal* species, const OdeReal* constant_species, "
                 "const OdeReal* parameters, const OdeReal* non_sampled_parameters)")
    lines.append("{")
    lines.append(f"\tout(0, 0) = -((parameters[{VAR_K_DEG}]*1.000000));")
    lines.append("}")
    return "\n".join(lines) + "\n"


DIVISION_SPECIES = ("cytokinesis", "nuclear_envelope", "G1S_break", "G2_break", "spindle_components", "assembled_spindle", "chromatid_separation")
DIVISION_RESET_VALUES = (0.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0)  # Cell::SetInitialConditionsFromOtherCell, Cell.cpp:127-133


def division_code(M: int, seed: int = 0, rate_decades: float = 1.0, with_apoptosis: bool = False):
    """An M-species cascade (cascade_code) followed by the seven species the reference's dividing cells are built around
    (Cell.cpp:119-148, 463-538) -- cytokinesis, nuclear_envelope, G1S_break, G2_break, spindle_components, assembled_spindle,
    chromatid_separation, in that order at indices M .. M + 6 -- and optionally `apoptosis` at M + 7. The cascade's last species
    releases the G1/S break, that releases the G2 break, the nuclear envelope breaks down, the spindle assembles, chromatids
    separate and cytokinesis accumulates until it passes 1: the cell divides, the daughters start with the seven species reset."""
    base = cascade_code(M, seed=seed, rate_decades=rate_decades)
    body = base[base.index("{") + 1: base.index("}")]
    laws = [ln for ln in body.split("\n") if ln.strip().startswith("ratelaws[")]
    outs = [ln for ln in body.split("\n") if ln.strip().startswith("out[")]
    K, L = VAR_K_CASCADE, M - 1
    cyt, ne, g1s, g2, sc_, asp, chs = (M + i for i in range(7))
    r0 = 2 * M
    extra = [
        f"(((parameters[{K}]*1.500000)*species[{g1s}])*species[{L}])",                                   # r0+0: G1S_break released by the cascade
        f"(((parameters[{K}]*2.000000)*species[{g2}])*(1.000000-species[{g1s}]))",                       # r0+1: G2_break released after G1/S
        f"(((parameters[{K}]*3.000000)*species[{ne}])*(1.000000-species[{g2}]))",                        # r0+2: nuclear envelope breakdown
        f"((parameters[{K}]*1.000000)*(1.000000-species[{g2}]))",                                         # r0+3: spindle components made
        f"(((parameters[{K}]*4.000000)*species[{sc_}])*(1.000000-species[{ne}]))",                       # r0+4: spindle assembly
        f"(((parameters[{K}]*2.000000)*hill_function_fixedn4(species[{asp}],0.300000))*(1.000000-species[{chs}]))",  # r0+5: chromatid separation
        f"((parameters[{K}]*1.200000)*hill_function_fixedn2(species[{chs}],0.500000))",                  # r0+6: cytokinesis accumulates
    ]
    if with_apoptosis:
        extra.append(f"((parameters[{VAR_K_DEG}]*0.110000)*(1.000000+species[0]))")                   # r0+7: slow death signal
    nr = r0 + len(extra)
    lines = [SIGNATURE, "{", f"\tOdeReal ratelaws[{nr}];"] + laws + [f"\tratelaws[{r0 + i}] = {e};" for i, e in enumerate(extra)] + outs
    lines += [f"\tout[{cyt}] = +ratelaws[{r0 + 6}];", f"\tout[{ne}] = -ratelaws[{r0 + 2}];", f"\tout[{g1s}] = -ratelaws[{r0 + 0}];",
              f"\tout[{g2}] = -ratelaws[{r0 + 1}];", f"\tout[{sc_}] = +ratelaws[{r0 + 3}]-ratelaws[{r0 + 4}];", f"\tout[{asp}] = +ratelaws[{r0 + 4}];",
              f"\tout[{chs}] = +ratelaws[{r0 + 5}];"]
    if with_apoptosis:
        lines.append(f"\tout[{M + 7}] = +ratelaws[{r0 + 7}];")
    lines.append("}")
    return "\n".join(lines) + "\n"


def make_dividing_problem(M: int = 5, num_cells: int = 24, max_cells: int = 160, T: int = 16, t_end: float = 12.0, seed: int = 3,
                          with_apoptosis: bool = False, data_cells: int = 4) -> CellPopProblem:
    """<experiment divide_cells="true" num_cells= max_cells=> around division_code: the read-out is the population average of the
    cascade's last species; observations are synthetic (a smooth curve plus noise: their values only shift the likelihood)."""
    N = M + 7 + (1 if with_apoptosis else 0)
    code = division_code(M, seed=seed, with_apoptosis=with_apoptosis)
    transforms = np.full(NUM_VARIABLES, TRANSFORM_LOG10, dtype=np.int32)
    transforms[VAR_VARIABILITY_SCALE] = TRANSFORM_NONE
    timepoints = t_end * np.arange(T) / (T - 1.0)
    ic = np.zeros(N)
    ic[0] = 0.05
    for k, v in enumerate(DIVISION_RESET_VALUES):
        ic[M + k] = v
    variability = [Variability(apply="multiplicative_log", model_parameter=VAR_K_CASCADE, scale_ix=VAR_VARIABILITY_SCALE),
                   Variability(apply="multiplicative_log", model_parameter=VAR_K_DEG, scale_ix=VAR_VARIABILITY_SCALE, negate=True),
                   Variability(apply="additive", initial_condition_species=1, scale_fixed=math.log(0.01), only_initial_cells=True)]
    sobol = sobol_points(100 * num_cells, len(variability))  # VariabilityPseudoRandomIterator.cpp:17
    rng = np.random.default_rng(seed + 500)
    observed = (0.8 * (1.0 - np.exp(-timepoints / 2.0)))[None, :] + 0.02 * rng.standard_normal((1, T))
    return CellPopProblem(derivative_code=code, num_species=N, initial_conditions=ic, transforms=transforms, num_cells=num_cells, timepoints=timepoints,
                          observed=observed, obs_species=[M - 1], constant_species=np.array([1.0]), sobol=sobol, variability=variability,
                          stdev_ix=VAR_STDEV, divide_cells=True, max_cells=max_cells, cytokinesis_species=M,
                          apoptosis_species=(M + 7) if with_apoptosis else None, division_reset_species=tuple(M + k for k in range(7)))


# ---- a Python evaluation of the generated text (only to synthesise observations) ---------------------------------
def _py_helpers():
    def hill(x, k, n):
        if x <= 0:
            return 0.0
        return x ** n / (x ** n + k ** n)

    def mm(kcat, KM, e, s):
        if e <= 0:
            return 0.0
        if s + KM < 0.1 * KM:
            bound = -KM + 0.1 * KM
            offset = e * kcat * bound / (0.01 * KM) - e * kcat * bound / (KM + bound)
            return e * kcat * s / (0.01 * KM) - offset
        return kcat * e * s / (KM + s)

    def tq(k, km, e, s):
        ekms = e + km + s
        return 0.5 * k * (ekms - math.sqrt(max(ekms * ekms - 4 * e * s, 0.0)))

    return dict(
        hill_function=hill, hill_function_fixedn2=lambda x, k: hill(x, k, 2), hill_function_fixedn4=lambda x, k: hill(x, k, 4),
        hill_function_fixedn10=lambda x, k: hill(x, k, 10), hill_function_fixedn16=lambda x, k: hill(x, k, 16),
        hill_function_fixedn100=lambda x, k: hill(x, k, 100), safepow=lambda x, n: 0.0 if x < 0 else x ** n, exp=math.exp, log=math.log,
        michaelis_menten_function=mm, tQSSA=tq, synthcap=lambda x: 1.0 if x <= 0 else 1.0 - x ** 10)


def python_rhs(code: str, N: int):
    body = code[code.index("{") + 1: code.index("}")]
    stmts = [s.strip() for s in body.split(";") if s.strip() and not s.strip().startswith("OdeReal ")]
    src = "\n".join(stmts)
    compiled = compile(src, "<generated_derivative>", "exec")
    helpers = _py_helpers()
    nr = int(re.search(r"ratelaws\[(\d+)\];", code).group(1))

    def f(t, y, constant_species, parameters, non_sampled_parameters=()):
        env = dict(helpers)
        env.update(species=y, constant_species=constant_species, parameters=parameters, non_sampled_parameters=non_sampled_parameters,
                   ratelaws=[0.0] * nr, out=[0.0] * N)
        exec(compiled, env)
        return env["out"]

    return f


def sobol_points(n: int, d: int) -> np.ndarray:
    """Unscrambled Sobol points, skipping the all-zero first point (its normal quantile is -inf)."""
    from scipy.stats import qmc

    if d == 0:
        return np.zeros((n, 0))
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pts = qmc.Sobol(d=d, scramble=False).random(n + 1)[1:]
    return np.ascontiguousarray(pts)


def default_values() -> np.ndarray:
    """Reference ("true") values of the sampled variables, in the sampler's (untransformed) parametrisation."""
    v = np.zeros(NUM_VARIABLES)
    v[VAR_K_IN] = math.log10(0.8)
    v[VAR_K_CASCADE] = math.log10(1.5)
    v[VAR_K_DEG] = math.log10(0.6)
    v[VAR_K_FEEDBACK] = math.log10(0.5)
    v[VAR_VARIABILITY_SCALE] = math.log(0.25)  # log of the s.d. of the per-cell log-normal factors
    v[VAR_STDEV] = math.log10(0.02)
    return v


def make_cellpop_problem(N: int = 12, num_cells: int = 10_000, T: int = 50, t_end: float = 8.0, seed: int = 1,
                         rate_decades: float = 2.0, data_cells: int = 64, replicates: int = 1, missing_fraction: float = 0.0,
                         two_species_readout: bool = False) -> CellPopProblem:
    """cfg 3 (N~12, 10k cells) / cfg 4 (N~50, 100k cells, rate_decades=6) style experiment."""
    from scipy.integrate import solve_ivp
    from scipy.special import ndtri

    code = cascade_code(N, seed=seed, rate_decades=rate_decades)
    transforms = np.full(NUM_VARIABLES, TRANSFORM_LOG10, dtype=np.int32)
    transforms[VAR_VARIABILITY_SCALE] = TRANSFORM_NONE
    timepoints = t_end * (np.arange(T) + 1.0) / T
    timepoints[0] = 0.0 if T > 4 else timepoints[0]  # one timepoint at the entry time: takes the initial condition
    ic = np.zeros(N)
    ic[0] = 0.05
    const = np.array([1.0])
    variability = [Variability(apply="multiplicative_log", model_parameter=VAR_K_IN, scale_ix=VAR_VARIABILITY_SCALE),
                   Variability(apply="multiplicative_log", model_parameter=VAR_K_DEG, scale_ix=VAR_VARIABILITY_SCALE, negate=True),
                   Variability(apply="additive", initial_condition_species=1, scale_fixed=math.log(0.01))]
    sobol = sobol_points(num_cells, len(variability))
    obs_species = [N - 1, N - 2] if two_species_readout else [N - 1]

    # synthetic observations: population average of `data_cells` cells at the reference parameters + noise
    v = default_values()
    tv = np.where(transforms == TRANSFORM_LOG10, 10.0 ** v, v)
    f = python_rhs(code, N)
    acc = np.zeros(T)
    rng = np.random.default_rng(seed + 1000)
    u = sobol_points(data_cells, len(variability))
    for ci in range(data_cells):
        p = tv.copy()
        y0 = ic.copy()
        z = ndtri(u[ci]) * math.exp(tv[VAR_VARIABILITY_SCALE])
        p[VAR_K_IN] *= math.exp(z[0])
        p[VAR_K_DEG] *= math.exp(-z[1])
        y0[1] += ndtri(u[ci][2]) * 0.01
        sol = solve_ivp(lambda t, y: f(t, y, const, p), (0.0, float(timepoints[-1])), y0, method="LSODA", t_eval=timepoints, rtol=1e-7, atol=1e-9)
        acc += sol.y[obs_species].sum(axis=0)
    avg = acc / data_cells
    observed = avg[None, :] + 0.02 * rng.standard_normal((replicates, T))
    if missing_fraction > 0:
        observed[rng.uniform(size=observed.shape) < missing_fraction] = np.nan
    return CellPopProblem(derivative_code=code, num_species=N, initial_conditions=ic, transforms=transforms, num_cells=num_cells,
                          timepoints=timepoints, observed=observed, obs_species=obs_species, constant_species=const, sobol=sobol,
                          variability=variability, stdev_ix=VAR_STDEV)


def make_time_course_problem(N: int = 8, num_cells: int = 24, T: int = 12, t_end: float = 8.0, seed: int = 41, noise: float = 0.03,
                             missing_fraction: float = 0.0, two_species_readout: bool = False, rate_decades: float = 2.0,
                             extra_marker_species: tuple = (), log_ratio_denominator: int | None = None) -> CellPopProblem:
    """<data type="time_course">: one observed trajectory per cell (live-cell imaging), as many observed as simulated cells. The
    observations are the trajectories of `num_cells` cells at the reference parameters (their own quasi-random draws, in a shuffled
    order) plus noise: the likelihood has to find out which simulated cell goes with which observed one."""
    from scipy.integrate import solve_ivp
    from scipy.special import ndtri

    base = make_cellpop_problem(N=N, num_cells=num_cells, T=T, t_end=t_end, seed=seed, data_cells=2, two_species_readout=two_species_readout,
                                rate_decades=rate_decades)
    v = default_values()
    tv = np.where(base.transforms == TRANSFORM_LOG10, 10.0 ** v, v)
    f = python_rhs(base.derivative_code, N)
    rng = np.random.default_rng(seed + 2000)
    u = sobol_points(2 * num_cells, len(base.variability))[num_cells:]  # other draws than the simulated cells will take
    observed = np.empty((num_cells, T))
    marker_observed = [np.empty((num_cells, T)) for _ in extra_marker_species]
    for ci in range(num_cells):
        p = tv.copy()
        y0 = base.initial_conditions.copy()
        z = ndtri(u[ci]) * math.exp(tv[VAR_VARIABILITY_SCALE])
        p[VAR_K_IN] *= math.exp(z[0])
        p[VAR_K_DEG] *= math.exp(-z[1])
        y0[1] += ndtri(u[ci][2]) * 0.01
        sol = solve_ivp(lambda t, y: f(t, y, base.constant_species, p), (0.0, float(base.timepoints[-1])), y0, method="LSODA", t_eval=base.timepoints,
                        rtol=1e-7, atol=1e-9)
        observed[ci] = sol.y[base.obs_species].sum(axis=0)
        if log_ratio_denominator is not None:  # use_log_ratio: a ratiometric reporter, species_name="a/b"
            with np.errstate(divide="ignore", invalid="ignore"):  # both species are 0 at t = 0: the caller leaves that timepoint out
                observed[ci] = np.log10(observed[ci] / np.maximum(sol.y[log_ratio_denominator], 1e-16))
        for mo, sp in zip(marker_observed, extra_marker_species):
            mo[ci] = sol.y[list(sp)].sum(axis=0)
    order = rng.permutation(num_cells)
    observed = observed[order] + noise * rng.standard_normal(observed.shape)
    if missing_fraction > 0:
        observed[rng.uniform(size=observed.shape) < missing_fraction] = np.nan
    import dataclasses

    from .cellpop_data import Marker

    # further markers (species_name="a;b;c"): the same observed cells in the same order, every marker with its own scale and noise
    markers = []
    for l, (mo, sp) in enumerate(zip(marker_observed, extra_marker_species), start=1):
        scale = 1.0 + 0.4 * l
        mobs = scale * mo[order] + 0.02 * l + noise * scale * rng.standard_normal(mo.shape)
        if missing_fraction > 0:
            mobs[rng.uniform(size=mobs.shape) < missing_fraction] = np.nan
        markers.append(Marker(obs_species=list(sp), observed=mobs, stdev=noise * 1.5 * scale, scale=scale, offset=0.02 * l))
    return dataclasses.replace(base, observed=observed, data_kind="time_course", stdev_ix=None, stdev=noise * 1.5, extra_markers=markers,
                               log_ratio_denominator=log_ratio_denominator)


def make_time_points_problem(N: int = 8, num_cells: int = 24, T: int = 8, t_end: float = 8.0, seed: int = 45, noise: float = 0.03,
                             min_fraction: float = 0.4, relative_to: int | None = None, extra_marker_species: tuple = ()) -> CellPopProblem:
    """<data type="time_points">: snapshots -- at every timepoint a different number of observed cells (fixed-cell imaging, flow
    cytometry), each value matched to one simulated cell at that time. `observed` is [num_cells slots][T]; the slots a timepoint
    does not fill are NaN, in no particular order."""
    import dataclasses

    tc = make_time_course_problem(N=N, num_cells=num_cells, T=T, t_end=t_end, seed=seed, noise=noise, extra_marker_species=extra_marker_species)
    rng = np.random.default_rng(seed + 3000)
    observed = tc.observed.copy()
    for ti in range(T):
        keep = int(rng.integers(max(1, int(min_fraction * num_cells)), num_cells + 1))
        drop = rng.permutation(num_cells)[keep:]
        observed[drop, ti] = np.nan
    if relative_to is not None:
        ref = np.nanmean(tc.observed[:, relative_to])
        observed = observed / ref
    markers = []
    for mk in tc.extra_markers:  # a cell that is absent at a timepoint is absent in every marker; a few values missing on top of that
        mobs = np.where(np.isnan(observed), np.nan, mk.observed)
        mobs[rng.uniform(size=mobs.shape) < 0.05] = np.nan
        markers.append(dataclasses.replace(mk, observed=mobs))
    return dataclasses.replace(tc, observed=observed, data_kind="time_points", value_relative_to_timepoint_ix=relative_to, extra_markers=markers)


def make_chain_values(C: int, seed: int = 20261018) -> np.ndarray:
    base = default_values()
    out = np.empty((C, NUM_VARIABLES))
    for c in range(C):
        rng = np.random.default_rng(seed + c)
        out[c] = base + rng.normal(0.0, 0.05, NUM_VARIABLES)
    return out
