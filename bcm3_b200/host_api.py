"""ctypes binding of libbcm3host.so: the C++ mirror of the reference's plugin / sampler interface (bcm3_b200/host)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libbcm3host.so")

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        # libbcm3host.so links against libbcm3b200.so next to it ($ORIGIN rpath)
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _err():
    return C.create_string_buffer(1024)


def run_pt(prior_xml: str, likelihood_xml: str, config_text: str, batched: bool = True, seed: int = 1, max_rows: int = 200000):
    """Parallel-tempered run with a host likelihood. Returns (rows[n, 3 + nvar], stats dict)."""
    lib = load()
    nvar = varset_info(prior_xml)[0]
    out = np.zeros((max_rows, nvar + 3))
    nrows = C.c_size_t()
    stats = (C.c_size_t * 4)()
    err = _err()
    rc = lib.bcm3host_run_pt(prior_xml.encode(), likelihood_xml.encode(), config_text.encode(), int(batched), C.c_ulonglong(seed),
                             out.ctypes.data_as(C.c_void_p), C.c_size_t(max_rows), C.byref(nrows), stats, err, C.c_size_t(1024))
    if rc != 0:
        raise RuntimeError(f"bcm3host_run_pt failed ({rc}): {err.value.decode()}")
    return out[: min(nrows.value, max_rows)], dict(evaluations=stats[0], batched_calls=stats[1], chains=stats[2], blocks=stats[3])


def run_pt_with_handlers(prior_xml: str, likelihood_xml: str, config_text: str, tsv_file: str | None = None, batched: bool = True, seed: int = 1,
                         max_rows: int = 200000):
    """run_pt with the reference's sample sinks attached: a SampleHandlerTSV file (posterior-chain samples as tab-separated text) and
    a SampleHandlerStoreMaxAPosteriori. Returns (rows, stats, map) with map = dict(lposterior, llikelihood, values)."""
    lib = load()
    nvar = varset_info(prior_xml)[0]
    out = np.zeros((max_rows, nvar + 3))
    best = np.zeros(2 + nvar)
    nrows = C.c_size_t()
    stats = (C.c_size_t * 4)()
    err = _err()
    rc = lib.bcm3host_run_pt_with_handlers(prior_xml.encode(), likelihood_xml.encode(), config_text.encode(), int(batched), C.c_ulonglong(seed),
                                           (tsv_file or "").encode(), best.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                           C.c_size_t(max_rows), C.byref(nrows), stats, err, C.c_size_t(1024))
    if rc != 0:
        raise RuntimeError(f"bcm3host_run_pt_with_handlers failed ({rc}): {err.value.decode()}")
    return (out[: min(nrows.value, max_rows)], dict(evaluations=stats[0], batched_calls=stats[1], chains=stats[2], blocks=stats[3]),
            dict(lposterior=best[0], llikelihood=best[1], values=best[2:].copy()))


def run_pt_poppk(prior_xml: str, likelihood_xml: str, config_text: str, trial, batched: bool = True, seed: int = 1, device: int = 0,
                 device_count: int = 1, max_rows: int = 100000):
    lib = load()
    nvar = varset_info(prior_xml)[0]
    out = np.zeros((max_rows, nvar + 3))
    nrows = C.c_size_t()
    stats = (C.c_size_t * 4)()
    err = _err()
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (
        trial.time, trial.observed_concentration, trial.dose, trial.dosing_interval, trial.dose_after_dose_change,
        trial.dose_change_time, trial.intermittent, trial.treatment_interruptions)]
    ptrs = [a.ctypes.data_as(C.c_void_p) for a in arrs]
    rc = lib.bcm3host_run_pt_poppk(prior_xml.encode(), likelihood_xml.encode(), config_text.encode(), int(batched), C.c_ulonglong(seed),
                                   C.c_size_t(trial.num_patients), C.c_size_t(trial.num_timepoints), *ptrs, int(device), int(device_count),
                                   out.ctypes.data_as(C.c_void_p), C.c_size_t(max_rows), C.byref(nrows), stats, err, C.c_size_t(1024))
    if rc != 0:
        raise RuntimeError(f"bcm3host_run_pt_poppk failed ({rc}): {err.value.decode()}")
    return out[: min(nrows.value, max_rows)], dict(evaluations=stats[0], batched_calls=stats[1], chains=stats[2], blocks=stats[3])


def tree_cluster(distance: np.ndarray, cut_height: float) -> np.ndarray:
    """Complete-linkage clustering cut at a height, as the Turek blocking strategy uses it: block index of every item."""
    lib = load()
    d = np.ascontiguousarray(distance, dtype=np.float64)
    out = np.empty(d.shape[0], dtype=np.int32)
    lib.bcm3host_tree_cluster(d.ctypes.data_as(C.c_void_p), C.c_size_t(d.shape[0]), C.c_double(cut_height), out.ctypes.data_as(C.c_void_p))
    return out


def evaluate(prior_xml: str, likelihood_xml: str, values: np.ndarray, batched: bool):
    lib = load()
    values = np.ascontiguousarray(values, dtype=np.float64)
    logp = np.empty(values.shape[0])
    err = _err()
    rc = lib.bcm3host_evaluate(prior_xml.encode(), likelihood_xml.encode(), values.ctypes.data_as(C.c_void_p), C.c_size_t(values.shape[0]),
                               int(batched), logp.ctypes.data_as(C.c_void_p), err, C.c_size_t(1024))
    if rc != 0:
        raise RuntimeError(f"bcm3host_evaluate failed ({rc}): {err.value.decode()}")
    return logp


def pharmaco_evaluate(prior_xml: str, likelihood_xml: str, trial, values, batched: bool = True, device: int = 0):
    """likelihood.xml type="pharmaco_population" through LikelihoodFactory and the plugin class; the trial arrays are what the NetCDF
    reader would supply (PharmacoPatient.cpp:24-46)."""
    lib = load()
    vals = np.ascontiguousarray(values, dtype=np.float64)
    P, T = trial.num_patients, trial.num_timepoints
    arr = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    keep = [arr(trial.time), arr(trial.observed_concentration), arr(trial.dose), arr(trial.dosing_interval), arr(trial.dose_after_dose_change),
            arr(trial.dose_change_time), arr(trial.intermittent), arr(trial.treatment_interruptions)]
    logp = np.empty(vals.shape[0])
    err = C.create_string_buffer(1024)
    rc = lib.bcm3host_pharmaco_evaluate(prior_xml.encode(), likelihood_xml.encode(), C.c_size_t(P), C.c_size_t(T), *[k.ctypes.data_as(C.c_void_p) for k in keep],
                                        int(device), vals.ctypes.data_as(C.c_void_p), C.c_size_t(vals.shape[0]), int(batched),
                                        logp.ctypes.data_as(C.c_void_p), err, C.c_size_t(1024))
    if rc != 0:
        raise RuntimeError(f"bcm3host_pharmaco_evaluate failed ({rc}): {err.value.decode()}")
    return logp


def varset_info(prior_xml: str, lookup: str | None = None):
    lib = load()
    n = C.c_size_t()
    idx = C.c_size_t()
    tr = (C.c_int * 65536)()
    rc = lib.bcm3host_varset_info(prior_xml.encode(), lookup.encode() if lookup else None, C.byref(n), tr, C.c_size_t(65536), C.byref(idx))
    if rc != 0:
        raise RuntimeError("bcm3host_varset_info failed")
    return n.value, list(tr[: min(n.value, 65536)]), (idx.value if lookup else None)


def cellpop_evaluate(prior_xml: str, likelihood_xml: str, problem, species_names, values=None, batched: bool = True, device: int = 0,
                     compile_only: bool = False):
    """CellPopulationLikelihoodB200 through LikelihoodFactory: likelihood.xml + the generated model / data set / quasi-random
    table of `problem` (bcm3_b200.cellpop_data.CellPopProblem). Returns (logp or None, descriptor handed to the C ABI)."""
    lib = load()
    p = problem
    names = (C.c_char_p * len(species_names))(*[s.encode() for s in species_names])
    ic = np.ascontiguousarray(p.initial_conditions, dtype=np.float64)
    cs = np.ascontiguousarray(p.constant_species, dtype=np.float64)
    tp = np.ascontiguousarray(p.timepoints, dtype=np.float64)
    obs = np.ascontiguousarray(p.observed, dtype=np.float64)
    sob = np.ascontiguousarray(p.sobol, dtype=np.float64)
    vals = np.zeros((1, p.num_variables)) if values is None else np.ascontiguousarray(values, dtype=np.float64)
    logp = np.empty(vals.shape[0])
    desc = C.create_string_buffer(4096)
    err = _err()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.bcm3host_cellpop_evaluate(prior_xml.encode(), likelihood_xml.encode(), p.derivative_code.encode(), C.c_size_t(p.num_species), names,
                                       vp(ic), C.c_size_t(cs.size), vp(cs), C.c_size_t(tp.size), C.c_size_t(obs.shape[0]), vp(tp), vp(obs),
                                       C.c_size_t(sob.size), vp(sob), int(device), int(compile_only), vp(vals), C.c_size_t(vals.shape[0]),
                                       int(batched), vp(logp), desc, C.c_size_t(4096), err, C.c_size_t(1024))
    if rc != 0:
        raise RuntimeError(f"bcm3host_cellpop_evaluate failed ({rc}): {err.value.decode()}")
    return (None if compile_only else logp), desc.value.decode()


def run_pt_cellpop(prior_xml: str, likelihood_xml: str, config_text: str, problem, species_names, batched: bool = True, seed: int = 1,
                   device: int = 0, max_rows: int = 100000):
    """Parallel-tempered run of the C++ sampler on CellPopulationLikelihoodB200 (one batched call per mutate round)."""
    lib = load()
    p = problem
    nvar = varset_info(prior_xml)[0]
    out = np.zeros((max_rows, nvar + 3))
    nrows = C.c_size_t()
    stats = (C.c_size_t * 4)()
    err = _err()
    names = (C.c_char_p * len(species_names))(*[s.encode() for s in species_names])
    ic = np.ascontiguousarray(p.initial_conditions, dtype=np.float64)
    cs = np.ascontiguousarray(p.constant_species, dtype=np.float64)
    tp = np.ascontiguousarray(p.timepoints, dtype=np.float64)
    obs = np.ascontiguousarray(p.observed, dtype=np.float64)
    sob = np.ascontiguousarray(p.sobol, dtype=np.float64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.bcm3host_run_pt_cellpop(prior_xml.encode(), likelihood_xml.encode(), config_text.encode(), int(batched), C.c_ulonglong(seed),
                                     p.derivative_code.encode(), C.c_size_t(p.num_species), names, vp(ic), C.c_size_t(cs.size), vp(cs),
                                     C.c_size_t(tp.size), C.c_size_t(obs.shape[0]), vp(tp), vp(obs), C.c_size_t(sob.size), vp(sob), int(device),
                                     vp(out), C.c_size_t(max_rows), C.byref(nrows), stats, err, C.c_size_t(1024))
    if rc != 0:
        raise RuntimeError(f"bcm3host_run_pt_cellpop failed ({rc}): {err.value.decode()}")
    return out[: min(nrows.value, max_rows)], dict(evaluations=stats[0], batched_calls=stats[1], chains=stats[2], blocks=stats[3])


def gmm_fit(samples, num_components: int, seed: int = 1, ess_factor: float = 1.0):
    """GaussianMixture::Fit (the fit behind proposal_type=gaussian_mixture). Returns dict(weights, means, covariances, aic, logl) or None."""
    lib = load()
    x = np.ascontiguousarray(samples, dtype=np.float64)
    n, D = x.shape
    K = num_components
    w, mu, cov, st = np.empty(K), np.empty((K, D)), np.empty((K, D, D)), np.empty(2)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.bcm3host_gmm_fit(vp(x), C.c_size_t(n), C.c_size_t(D), C.c_size_t(K), C.c_ulonglong(seed), C.c_double(ess_factor), vp(w), vp(mu), vp(cov), vp(st))
    if rc != 0:
        return None
    return dict(weights=w, means=mu, covariances=cov, aic=st[0], logl=st[1])


def gmm_evaluate(weights, means, covariances, x):
    """log pdf [m] and responsibilities [m][K] of an explicit mixture at x[m][D]."""
    lib = load()
    w = np.ascontiguousarray(weights, dtype=np.float64)
    mu = np.ascontiguousarray(means, dtype=np.float64)
    cov = np.ascontiguousarray(covariances, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    K, D = mu.shape
    lp, resp = np.empty(x.shape[0]), np.empty((x.shape[0], K))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.bcm3host_gmm_evaluate(C.c_size_t(K), C.c_size_t(D), vp(w), vp(mu), vp(cov), vp(x), C.c_size_t(x.shape[0]), vp(lp), vp(resp))
    if rc != 0:
        raise RuntimeError("mixture covariance is not positive definite")
    return lp, resp


def symmetric_eigen(a):
    lib = load()
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    vals, vecs = np.empty(n), np.empty((n, n))
    lib.bcm3host_symmetric_eigen(a.ctypes.data_as(C.c_void_p), C.c_size_t(n), vals.ctypes.data_as(C.c_void_p), vecs.ctypes.data_as(C.c_void_p))
    return vals, vecs.T  # column-major -> columns are eigenvectors


class CellPopSession:
    """CellPopulationLikelihoodB200 with any number of <experiment> / <data> elements: likelihood.xml goes through
    LikelihoodFactory, then the generated model, the data sets and the quasi-random tables are supplied the way the
    reference's SBML / NetCDF readers would (`set_model`, `set_data`, `set_sobol`), then `post_initialize`."""

    def __init__(self, prior_xml: str, likelihood_xml: str):
        self.lib = lib = load()
        lib.bcm3host_cellpop_open.restype = C.c_void_p
        lib.bcm3host_cellpop_layout.restype = C.c_size_t
        err = _err()
        self.nvar = varset_info(prior_xml)[0]
        self.handle = lib.bcm3host_cellpop_open(prior_xml.encode(), likelihood_xml.encode(), err, C.c_size_t(1024))
        if not self.handle:
            raise RuntimeError(f"bcm3host_cellpop_open failed: {err.value.decode()}")
        nd = (C.c_size_t * 64)()
        n = lib.bcm3host_cellpop_layout(C.c_void_p(self.handle), nd, C.c_size_t(64))
        self.num_data_sets = [int(nd[i]) for i in range(n)]

    def set_model(self, problem, species_names, experiment: int = -1) -> None:
        p = problem
        names = (C.c_char_p * len(species_names))(*[s.encode() for s in species_names])
        ic = np.ascontiguousarray(p.initial_conditions, dtype=np.float64)
        cs = np.ascontiguousarray(p.constant_species, dtype=np.float64)
        rc = self.lib.bcm3host_cellpop_set_model(C.c_void_p(self.handle), C.c_long(experiment), p.derivative_code.encode(), C.c_size_t(p.num_species),
                                                 names, ic.ctypes.data_as(C.c_void_p), C.c_size_t(cs.size), cs.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("bcm3host_cellpop_set_model: no such experiment")

    def set_model_with_non_sampled(self, problem, species_names, experiment: int = -1) -> None:
        """set_model that also hands over the model's non-sampled parameter values (SBMLModel::AddNonSampledParameters side)."""
        p = problem
        names = (C.c_char_p * p.num_species)(*[n.encode() for n in species_names])
        ic = np.ascontiguousarray(p.initial_conditions, dtype=np.float64)
        cs = np.ascontiguousarray(p.constant_species, dtype=np.float64)
        ns = np.ascontiguousarray(p.non_sampled_parameters, dtype=np.float64)
        rc = self.lib.bcm3host_cellpop_set_model_ns(C.c_void_p(self.handle), C.c_long(experiment), p.derivative_code.encode(), C.c_size_t(p.num_species), names,
                                                    ic.ctypes.data_as(C.c_void_p), C.c_size_t(cs.size), cs.ctypes.data_as(C.c_void_p), C.c_size_t(ns.size),
                                                    ns.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("bcm3host_cellpop_set_model_ns: no such experiment")

    def add_non_sampled_parameters(self, names) -> None:
        """bcm3::Likelihood::AddNonSampledParameters (Likelihood.h:18)."""
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        if self.lib.bcm3host_cellpop_add_non_sampled(C.c_void_p(self.handle), C.c_size_t(len(names)), arr) != 0:
            raise RuntimeError("AddNonSampledParameters failed")

    def set_non_sampled_parameters(self, values) -> None:
        """bcm3::Likelihood::SetNonSampledParameters (Likelihood.h:19)."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        self.lib.bcm3host_cellpop_set_non_sampled(C.c_void_p(self.handle), C.c_size_t(v.size), v.ctypes.data_as(C.c_void_p))

    def output_evaluation_statistics(self, path: str) -> None:
        """bcm3::Likelihood::OutputEvaluationStatistics (Likelihood.h:22)."""
        self.lib.bcm3host_cellpop_output_statistics(C.c_void_p(self.handle), path.encode())

    def set_data(self, experiment: int, data_set: int, timepoints, observed) -> None:
        tp = np.ascontiguousarray(timepoints, dtype=np.float64)
        obs = np.ascontiguousarray(observed, dtype=np.float64)
        obs = obs.reshape(-1, tp.size) if tp.size else obs.reshape(1, 0)
        rc = self.lib.bcm3host_cellpop_set_data(C.c_void_p(self.handle), C.c_size_t(experiment), C.c_size_t(data_set), C.c_size_t(tp.size),
                                                C.c_size_t(obs.shape[0]), tp.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("bcm3host_cellpop_set_data: no such experiment / data set")

    def set_sobol(self, experiment: int, table) -> None:
        sob = np.ascontiguousarray(table, dtype=np.float64)
        rc = self.lib.bcm3host_cellpop_set_sobol(C.c_void_p(self.handle), C.c_size_t(experiment), C.c_size_t(sob.size), sob.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("bcm3host_cellpop_set_sobol: no such experiment")

    def share_integration(self, share: bool) -> None:
        """False: one handle -- one integration of the experiment's cells -- per <data> element (the default shares one per experiment)."""
        self.lib.bcm3host_cellpop_share_integration.restype = None
        self.lib.bcm3host_cellpop_share_integration(C.c_void_p(self.handle), int(share))

    def post_initialize(self, device: int = 0, compile_only: bool = False) -> None:
        err = _err()
        rc = self.lib.bcm3host_cellpop_post_initialize(C.c_void_p(self.handle), int(device), int(compile_only), err, C.c_size_t(1024))
        if rc != 0:
            raise RuntimeError(f"PostInitialize failed: {err.value.decode()}")

    def descriptor(self, experiment: int, data_set: int) -> str:
        buf = C.create_string_buffer(4096)
        if self.lib.bcm3host_cellpop_descriptor(C.c_void_p(self.handle), C.c_size_t(experiment), C.c_size_t(data_set), buf, C.c_size_t(4096)) != 0:
            raise RuntimeError("bcm3host_cellpop_descriptor: no such experiment / data set")
        return buf.value.decode()

    def fixed_parameters(self, experiment: int) -> list:
        """The <set_parameter> elements of an experiment: [(name, value)], to be applied to the cell model before code generation."""
        self.lib.bcm3host_cellpop_fixed_parameter.restype = C.c_size_t
        out, k = [], 0
        while True:
            buf, v = C.create_string_buffer(256), C.c_double()
            n = self.lib.bcm3host_cellpop_fixed_parameter(C.c_void_p(self.handle), C.c_size_t(experiment), C.c_size_t(k), buf, C.c_size_t(256), C.byref(v))
            if k >= n:
                return out
            out.append((buf.value.decode(), v.value))
            k += 1

    def evaluate(self, values, batched: bool = True) -> np.ndarray:
        vals = np.ascontiguousarray(values, dtype=np.float64).reshape(-1, self.nvar)
        logp = np.empty(vals.shape[0])
        err = _err()
        rc = self.lib.bcm3host_cellpop_session_evaluate(C.c_void_p(self.handle), vals.ctypes.data_as(C.c_void_p), C.c_size_t(vals.shape[0]),
                                                        int(batched), logp.ctypes.data_as(C.c_void_p), err, C.c_size_t(1024))
        if rc != 0:
            raise RuntimeError(f"evaluate failed ({rc}): {err.value.decode()}")
        return logp

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.bcm3host_cellpop_close(C.c_void_p(self.handle))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
