"""CPU suite: the drop-in boundary. The C-ABI library must load, export every symbol include/bcm3b200.h declares,
validate its inputs, and refuse to compute without a CUDA device (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bcm3_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "bcm3b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bcm3b200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    from bcm3_b200 import _lib

    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/bcm3b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms


def test_argument_validation_without_device(built):
    from bcm3_b200 import _lib

    lib = _lib.load()
    h = C.c_void_p()
    assert lib.bcm3b200_create(b"no_such_model", b"", 0, 1, C.byref(h)) == -4
    assert b"unknown model kind" in lib.bcm3b200_last_error()
    assert lib.bcm3b200_create(b"pop_pk_trajectory", b"type=seven", 10, 1, C.byref(h)) == -4
    assert lib.bcm3b200_create(b"pop_pk_trajectory", b"type=one;drug=lapatinib", 23, 1, C.byref(h)) == -1
    assert lib.bcm3b200_create(None, None, 0, 1, C.byref(h)) == -1


def test_combine_partials_reproduces_serial_loop(built):
    """partial rows (finite sum, first -inf index, first NaN index) -> the value the reference's serial
    `logp += patient_logllh; if (logp == -inf) break;` loop returns (LikelihoodPopPKTrajectory.cpp:427-440)."""
    from bcm3_b200 import _lib

    lib = _lib.load()
    inf = np.inf
    partial = np.array([[-10.0, -20.0, -30.0, -40.0],  # sums
                        [inf, 7.0, 7.0, inf],          # first -inf
                        [inf, inf, 3.0, 9.0]])         # first NaN
    logp = np.empty(4)
    status = np.empty(4, dtype=np.int32)
    assert lib.bcm3b200_combine_partials(4, partial.ctypes.data, logp.ctypes.data, status.ctypes.data) == 0
    assert logp[0] == -10.0 and logp[1] == -inf and np.isnan(logp[2]) and np.isnan(logp[3])
    assert status.tolist() == [0, 0, 1, 1]
    # NaN after the first -inf is never reached by the serial loop
    partial[:, 2] = [-30.0, 3.0, 7.0]
    lib.bcm3b200_combine_partials(4, partial.ctypes.data, logp.ctypes.data, status.ctypes.data)
    assert logp[2] == -inf and status[2] == 0


def test_no_cpu_fallback(built):
    """On a box without a GPU the product path must fail loudly instead of computing on the host."""
    from bcm3_b200 import _lib
    from bcm3_b200.poppk import PopPKEvaluator

    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.Bcm3B200Error) as e:
        PopPKEvaluator(syn.make_poppk_problem(P=8))
    assert e.value.code == -3


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under bcm3_b200/ may reference it."""
    for base, _, files in os.walk(os.path.join(ROOT, "bcm3_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "libbcm3ref" not in text and "oracle/" not in text, f
