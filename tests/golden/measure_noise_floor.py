"""Measures how far the REFERENCE moves under its own compiler flags, per golden fixture, and stores the result in the
fixture (.npz key `noise_floor`, one relative difference per chain) -- the floor below which a parity assertion against the
compiled reference has no meaning.

Three builds of the same reference sources evaluate every fixture's inputs:
  A  oracle/_ref/libbcm3ref.so          -O3 -march=x86-64-v3 (FMA contraction on: what the goldens were made with), generated RHS strict
  B  oracle/_ref/libbcm3ref_strict.so   the same + -ffp-contract=off (oracle/ref/Makefile target `strict`), generated RHS strict
  C  build A with the generated right-hand side compiled as the reference's CMake would (-O3 -march, contraction on)
noise_floor[c] = max(|logp_A - logp_B|, |logp_A - logp_C|) / |logp_A|  (C only exists for the cell_population fixtures).

The GPU tests assert  |gpu - golden| / |golden| <= max(1e-6, noise_floor)  (tests/util.py::parity_tolerance): the north-star
bar wherever the reference itself is reproducible at that level, the reference's own reproducibility elsewhere.
Run where /root/reference is mounted:  make -C oracle/ref strict && python tests/golden/measure_noise_floor.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from tests.util import CELLPOP_GOLDEN_NAMES, GOLDEN_NAMES, GOLDEN_DIR, load_cellpop_golden, load_golden  # noqa: E402

STRICT = os.path.join(ROOT, "oracle", "_ref", "libbcm3ref_strict.so")


def resave(name, **extra):
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    z = dict(np.load(path))
    z.update(extra)
    np.savez_compressed(path, **z)


def main():
    a = oracle.load("ref")
    b = oracle.Oracle("ref", STRICT)
    for name in GOLDEN_NAMES:
        prob, gold = load_golden(name)
        la = a.poppk_evaluate(prob, gold["values"], threads=2, want_counters=True)
        lb = b.poppk_evaluate(prob, gold["values"], threads=2, want_counters=True)
        assert np.array_equal(la["logp"], gold["logp"])
        fin = np.isfinite(la["logp"])
        floor = np.zeros(len(la["logp"]))
        floor[fin] = np.abs(la["logp"][fin] - lb["logp"][fin]) / np.abs(la["logp"][fin])
        same = (la["counters"] == lb["counters"]).all(axis=-1).mean()
        resave(name, noise_floor=floor, noise_floor_counter_match=np.float64(same))
        print(f"{name:40s} floor {floor.max():.2e}  systems with identical counters between the two builds {same:.4f}")
    for name in CELLPOP_GOLDEN_NAMES:
        prob, gold = load_cellpop_golden(name)
        oracle.rhs_build = "strict"
        ra = a.cellpop_evaluate(prob, gold["values"], threads=1, want_steps=True, want_cell_values=True)
        rb = b.cellpop_evaluate(prob, gold["values"], threads=1, want_steps=True, want_cell_values=True)
        oracle.rhs_build = "contracted"
        rc = a.cellpop_evaluate(prob, gold["values"], threads=1, want_steps=True, want_cell_values=True)
        oracle.rhs_build = "strict"
        assert np.array_equal(ra["logp"], gold["logp"])
        la = ra["logp"]
        f_solver = np.abs(la - rb["logp"]) / np.abs(la)
        f_rhs = np.abs(la - rc["logp"]) / np.abs(la)
        floor = np.maximum(f_solver, f_rhs)
        m = ~np.isnan(ra["cell_values"])
        traj = max(np.abs(ra["cell_values"][m] - rb["cell_values"][m]).max(), np.abs(ra["cell_values"][m] - rc["cell_values"][m]).max())
        steps = min((ra["cell_steps"] == rb["cell_steps"]).mean(), (ra["cell_steps"] == rc["cell_steps"]).mean())
        resave(name, noise_floor=floor, noise_floor_solver=f_solver, noise_floor_rhs=f_rhs, noise_floor_trajectory=np.float64(traj),
               noise_floor_step_match=np.float64(steps))
        print(f"{name:40s} floor {floor.max():.2e} (solver flags {f_solver.max():.2e}, RHS flags {f_rhs.max():.2e})  "
              f"max trajectory difference {traj:.2e}  cells with identical step counts {steps:.3f}")


if __name__ == "__main__":
    main()
