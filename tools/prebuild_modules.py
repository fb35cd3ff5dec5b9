"""Compile (nvcc, sm_100a) the per-model kernel libraries that the test-suite, smoke() and bench.py ask for, into the in-tree
cache (bcm3_b200/codegen_cache), so that a GPU box that receives the tree does not spend its time in nvcc. Safe to run on
a machine without a GPU; every model is compiled in its own process, several at a time."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r"""
import sys, dataclasses
sys.path.insert(0, {root!r})
import numpy as np
from bcm3_b200 import synthetic_cellpop as sc
from bcm3_b200.cellpop import CellPopEvaluator
from tests.util import load_cellpop_golden
kind, arg, kernel = {kind!r}, {arg!r}, {kernel!r}
if kind == "plugin":  # the host plugin's two-experiment set-up, one integration per experiment (arg) or per data set
    from tests.util import cellpop_two_experiment_setup, open_cellpop_session
    s = open_cellpop_session(*cellpop_two_experiment_setup())
    s.share_integration(arg)
    s.post_initialize(compile_only=True)
    s.close()
    sys.exit(0)
if kind == "golden":
    prob, _ = load_cellpop_golden(arg)
else:
    prob = sc.make_cellpop_problem(**arg)
CellPopEvaluator(prob, compile_only=True, kernel=kernel).close()
"""


def jobs():
    sys.path.insert(0, ROOT)
    from tests.util import CELLPOP_GOLDEN_NAMES

    out = [("golden", n, "auto") for n in CELLPOP_GOLDEN_NAMES]
    out += [("golden", "cellpop_n12_normal", "warp"), ("golden", "cellpop_n12_normal", "thread")]
    out += [("synthetic", dict(N=12, num_cells=8, T=50, data_cells=2, seed=5), "auto")]          # config-3 shape (bench, tests)
    out += [("synthetic", dict(N=12, num_cells=8, T=12, data_cells=2, seed=2), k) for k in ("auto", "warp")]  # smoke
    out += [("synthetic", dict(N=8, num_cells=8, T=10, data_cells=2, seed=9), k) for k in ("auto", "warp", "thread")]  # host plugin tests, entry-time test
    out += [("synthetic", dict(N=n, num_cells=8, T=12, data_cells=2, seed=40 + n, rate_decades=d), "auto")
            for n, d in ((3, 2.0), (7, 2.0), (16, 3.0), (33, 3.0), (50, 4.0))]
    out += [("plugin", True, "auto"), ("plugin", False, "auto")]
    return out


def main(parallel: int = 6) -> int:
    pending = jobs()
    running, failed = [], 0
    while pending or running:
        while pending and len(running) < parallel:
            kind, arg, kernel = pending.pop(0)
            code = SNIPPET.format(root=ROOT, kind=kind, arg=arg, kernel=kernel)
            running.append(subprocess.Popen([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE))
        p = running.pop(0)
        _, err = p.communicate()
        if p.returncode != 0:
            failed += 1
            sys.stderr.write(err.decode(errors="replace")[-800:] + "\n")
    return failed


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
