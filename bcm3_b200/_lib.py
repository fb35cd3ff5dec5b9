"""ctypes binding of the C ABI in include/bcm3b200.h (libbcm3b200.so, built in-tree by __graft_entry__.build()).

There is deliberately no fallback: if the shared library is missing the import of this module fails,
and every compute entry point raises when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BCM3B200_LIB", os.path.join(PKG_DIR, "libbcm3b200.so"))  # override: kernel build experiments

# every symbol declared in include/bcm3b200.h
EXPORTS = [
    "bcm3b200_create",
    "bcm3b200_match_cells",
    "bcm3b200_set_data",
    "bcm3b200_set_text",
    "bcm3b200_get_cell_diagnostics",
    "bcm3b200_finalize",
    "bcm3b200_evaluate_batch",
    "bcm3b200_evaluate_batch_device",
    "bcm3b200_enqueue_batch",
    "bcm3b200_combine_partials",
    "bcm3b200_comm_unique_id",
    "bcm3b200_comm_init",
    "bcm3b200_exchange_partials",
    "bcm3b200_cellpop_finish",
    "bcm3b200_get_diagnostics",
    "bcm3b200_set_option",
    "bcm3b200_get_stat",
    "bcm3b200_destroy",
    "bcm3b200_last_error",
    "bcm3b200_device_count",
    "bcm3b200_measure_fp64_peak",
    "bcm3b200_host_alloc",
    "bcm3b200_host_free",
]


class Bcm3B200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"bcm3b200 error {code}: {message}")
        self.code = code


_lib = None


def _prefer_bundled_nccl() -> None:
    """A Python process may import torch AFTER the library has loaded NCCL; torch must then find the NCCL it was built
    against (the pip package nvidia-nccl, same soname as the system library). Point the library at that copy unless the
    caller chose one (BCM3B200_NCCL_LIB). A C++ host has no such concern and gets the system libnccl.so.2."""
    if os.environ.get("BCM3B200_NCCL_LIB"):
        return
    import importlib.util

    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for d in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(d, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["BCM3B200_NCCL_LIB"] = cand
            return


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    _prefer_bundled_nccl()
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, sz, dp, ip = C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.bcm3b200_create.argtypes = [C.c_char_p, C.c_char_p, sz, C.c_int, C.POINTER(vp)]
    lib.bcm3b200_set_data.argtypes = [vp, C.c_char_p, vp, C.POINTER(sz), C.c_int]
    lib.bcm3b200_set_text.argtypes = [vp, C.c_char_p, C.c_char_p, sz]
    lib.bcm3b200_get_cell_diagnostics.argtypes = [vp, vp, vp, vp, vp]
    lib.bcm3b200_finalize.argtypes = [vp]
    lib.bcm3b200_evaluate_batch.argtypes = [vp, sz, sz, vp, vp, vp]
    lib.bcm3b200_evaluate_batch_device.argtypes = [vp, sz, sz, vp, vp, vp]
    lib.bcm3b200_enqueue_batch.argtypes = [vp, sz, sz, vp, vp, vp]
    lib.bcm3b200_combine_partials.argtypes = [sz, vp, vp, vp]
    lib.bcm3b200_comm_unique_id.argtypes = [vp, sz]
    lib.bcm3b200_comm_init.argtypes = [vp, vp, sz]
    lib.bcm3b200_exchange_partials.argtypes = [vp, sz, vp, vp]
    lib.bcm3b200_cellpop_finish.argtypes = [vp, sz, vp, vp, vp, vp]
    lib.bcm3b200_get_diagnostics.argtypes = [vp, vp, vp, vp]
    lib.bcm3b200_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    lib.bcm3b200_get_stat.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64)]
    lib.bcm3b200_destroy.argtypes = [vp]
    lib.bcm3b200_destroy.restype = None
    lib.bcm3b200_last_error.restype = C.c_char_p
    lib.bcm3b200_device_count.restype = C.c_int
    lib.bcm3b200_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.bcm3b200_measure_fp64_peak.restype = C.c_int
    lib.bcm3b200_host_alloc.argtypes = [sz]
    lib.bcm3b200_host_alloc.restype = vp
    lib.bcm3b200_host_free.argtypes = [vp]
    lib.bcm3b200_host_free.restype = None
    lib.bcm3b200_match_cells.argtypes = [C.c_int, vp, vp]
    lib.bcm3b200_match_cells.restype = C.c_int
    for name in ("create", "set_data", "set_text", "get_cell_diagnostics", "finalize", "evaluate_batch", "evaluate_batch_device", "enqueue_batch", "combine_partials", "cellpop_finish", "comm_unique_id", "comm_init", "exchange_partials",
                 "get_diagnostics", "set_option", "get_stat"):
        getattr(lib, "bcm3b200_" + name).restype = C.c_int
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise Bcm3B200Error(rc, load().bcm3b200_last_error().decode(errors="replace"))


def match_cells(cost):
    """The simulated cell (column) the per-cell time_course likelihood assigns to every observed cell (row) of a complete [n][n]
    cost matrix: bcm3b200_match_cells, host code only. None when no perfect matching was found."""
    import numpy as np

    cost = np.ascontiguousarray(cost, dtype=np.float64)
    n = cost.shape[0]
    assert cost.shape == (n, n)
    out = np.full(n, -1, dtype=np.int32)
    rc = load().bcm3b200_match_cells(n, cost.ctypes.data, out.ctypes.data)
    return out if rc == 0 else None


def measure_fp64_peak(device: int = 0) -> float:
    v = C.c_double()
    check(load().bcm3b200_measure_fp64_peak(device, C.byref(v)))
    return float(v.value)


def device_count() -> int:
    return int(load().bcm3b200_device_count())


COMM_ID_BYTES = 128  # BCM3B200_COMM_ID_BYTES


def comm_unique_id() -> bytes:
    """Rank 0's call: the id every rank passes to comm_init (an ncclUniqueId as plain bytes)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    check(load().bcm3b200_comm_unique_id(buf, COMM_ID_BYTES))
    return buf.raw
