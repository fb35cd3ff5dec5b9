"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU checkers (oracle/oracle_api.h).

``load("ref")``  -> oracle/_ref/libbcm3ref.so   the reference's own compiled CVODE/odecommon stack
``load("port")`` -> oracle/_build/liboracle.so  the plain-C restatement (oracle/cvode_bdf.c, oracle/poppk_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; nothing under bcm3_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libbcm3ref.so")
PORT_LIB = os.path.join(HERE, "_build", "liboracle.so")
REFERENCE_ROOT = "/root/reference"

NUM_COUNTERS = 8
CNT_STEPS, CNT_NFE, CNT_NSETUPS, CNT_NJE, CNT_NETF, CNT_NCFN, CNT_NNI, CNT_OK = range(8)


class _PopPKProblem(C.Structure):
    _fields_ = [
        ("pk_type", C.c_int32),
        ("num_patients", C.c_int32),
        ("num_timepoints", C.c_int32),
        ("num_variables", C.c_int32),
        ("sd_ix", C.c_int32),
        ("max_steps", C.c_int32),
        ("fixed_vod", C.c_double),
        ("fixed_periphery_fwd", C.c_double),
        ("fixed_periphery_bwd", C.c_double),
        ("mol_weight", C.c_double),
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("time", C.c_void_p),
        ("observed_concentration", C.c_void_p),
        ("dose", C.c_void_p),
        ("dosing_interval", C.c_void_p),
        ("dose_after_dose_change", C.c_void_p),
        ("dose_change_time", C.c_void_p),
        ("intermittent", C.c_void_p),
        ("skipped_days", C.c_void_p),
        ("simulate_until", C.c_void_p),
        ("transforms", C.c_void_p),
    ]


def build_port() -> str:
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    return PORT_LIB


def build_ref() -> str | None:
    """Compile the reference's own sources in place; only possible where /root/reference is mounted."""
    if not os.path.isdir(REFERENCE_ROOT):
        return REF_LIB if os.path.exists(REF_LIB) else None
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "ref")], check=True)
    return REF_LIB


def available(kind: str) -> bool:
    return os.path.exists(REF_LIB if kind == "ref" else PORT_LIB)


class Oracle:
    def __init__(self, kind: str):
        path = REF_LIB if kind == "ref" else PORT_LIB
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built (run `make -C oracle` / `make -C oracle/ref`)")
        self.kind = kind
        self.lib = C.CDLL(path)
        self.lib.oracle_poppk_evaluate.restype = C.c_int
        self.lib.oracle_poppk_evaluate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_void_p, C.c_int]
        self.lib.oracle_kind.restype = C.c_char_p
        assert self.lib.oracle_kind().decode() == kind

    def poppk_evaluate(self, problem, values, threads: int = 1, want_conc=False, want_patient_ll=False, want_counters=False):
        """problem: bcm3_b200.poppk_data.PopPKProblem; values [C, nvar] -> dict(logp[C], conc, patient_ll, counters)."""
        tr = problem.trial
        P, T = tr.num_patients, tr.num_timepoints
        values = np.ascontiguousarray(values, dtype=np.float64)
        if values.ndim == 1:
            values = values[None, :]
        nC, nvar = values.shape
        assert nvar == problem.num_variables
        keep = dict(
            time=np.ascontiguousarray(tr.time, dtype=np.float64),
            obs=np.ascontiguousarray(tr.observed_concentration, dtype=np.float64),
            dose=np.ascontiguousarray(tr.dose, dtype=np.float64),
            di=np.ascontiguousarray(tr.dosing_interval, dtype=np.float64),
            dac=np.ascontiguousarray(tr.dose_after_dose_change, dtype=np.float64),
            dct=np.ascontiguousarray(tr.dose_change_time, dtype=np.float64),
            inter=np.ascontiguousarray(tr.intermittent, dtype=np.int32),
            skipped=np.ascontiguousarray(problem.skipped_days, dtype=np.uint32),
            su=np.ascontiguousarray(problem.simulate_until, dtype=np.int32),
            transforms=np.ascontiguousarray(problem.transforms, dtype=np.int32),
        )
        s = _PopPKProblem(
            pk_type=problem.pk_type, num_patients=P, num_timepoints=T, num_variables=nvar, sd_ix=problem.sd_ix,
            max_steps=problem.max_steps, fixed_vod=problem.fixed_vod, fixed_periphery_fwd=problem.fixed_periphery_fwd,
            fixed_periphery_bwd=problem.fixed_periphery_bwd, mol_weight=problem.mol_weight, rtol=problem.rtol,
            atol=problem.atol,
            time=keep["time"].ctypes.data, observed_concentration=keep["obs"].ctypes.data, dose=keep["dose"].ctypes.data,
            dosing_interval=keep["di"].ctypes.data, dose_after_dose_change=keep["dac"].ctypes.data,
            dose_change_time=keep["dct"].ctypes.data, intermittent=keep["inter"].ctypes.data,
            skipped_days=keep["skipped"].ctypes.data, simulate_until=keep["su"].ctypes.data,
            transforms=keep["transforms"].ctypes.data,
        )
        logp = np.empty(nC, dtype=np.float64)
        conc = np.empty((nC, P, T), dtype=np.float64) if want_conc else None
        pll = np.empty((nC, P), dtype=np.float64) if want_patient_ll else None
        cnt = np.zeros((nC, P, NUM_COUNTERS), dtype=np.int64) if want_counters else None
        rc = self.lib.oracle_poppk_evaluate(
            C.byref(s), nC, values.ctypes.data, logp.ctypes.data,
            conc.ctypes.data if conc is not None else None,
            pll.ctypes.data if pll is not None else None,
            cnt.ctypes.data if cnt is not None else None, int(threads))
        if rc != 0:
            raise RuntimeError(f"oracle_poppk_evaluate failed: {rc}")
        return dict(logp=logp, conc=conc, patient_ll=pll, counters=cnt)


_cache: dict[str, Oracle] = {}


def load(kind: str = "ref") -> Oracle:
    if kind not in _cache:
        _cache[kind] = Oracle(kind)
    return _cache[kind]
