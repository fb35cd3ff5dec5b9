"""Measures how far the REFERENCE moves under its own compiler flags, per golden fixture, and stores the result in the
fixture (.npz key `noise_floor`, one relative difference per chain) -- the floor below which a parity assertion against the
compiled reference has no meaning.

Three builds of the same reference sources evaluate every fixture's inputs:
  A  oracle/_ref/libbcm3ref.so          -O3 -march=x86-64-v3 (FMA contraction on: what the goldens were made with), generated RHS strict
  B  oracle/_ref/libbcm3ref_strict.so   the same + -ffp-contract=off (oracle/ref/Makefile target `strict`), generated RHS strict
  C  build A with the generated right-hand side compiled as the reference's CMake would (-O3 -march, contraction on)
  D  build A evaluated at inputs moved by ONE unit in the last place (every sampled value -> nextafter, both directions):
     the conditioning of the reference's own result. For small models the builds A/B/C differ in very few operations (a
     3-state model runs the closed-form inverse, nothing to contract), so A == B == C on most cells says nothing about how
     sensitive the step-size/order decisions are -- the fresh 3-species case of tools/gpu_parity_report.py agrees bit for
     bit between the builds and still moves by 4.7e-5 (relative, per-chain logp) under a 1-ulp change of its inputs.
noise_floor[c] = max over B, C, D of |logp_A - logp_X| / |logp_A|  (C and D only for the cell_population fixtures).

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
    only = sys.argv[1:]  # fixture names: measure just these
    for name in GOLDEN_NAMES:
        if only and name not in only:
            continue
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
        if only and name not in only:
            continue
        prob, gold = load_cellpop_golden(name)
        oracle.rhs_build = "strict"
        ra = a.cellpop_evaluate(prob, gold["values"], threads=1, want_steps=True, want_cell_values=True)
        rb = b.cellpop_evaluate(prob, gold["values"], threads=1, want_steps=True, want_cell_values=True)
        oracle.rhs_build = "contracted"
        rc = a.cellpop_evaluate(prob, gold["values"], threads=1, want_steps=True, want_cell_values=True)
        oracle.rhs_build = "strict"
        assert np.array_equal(ra["logp"], gold["logp"])
        la = ra["logp"]

        def rel(other):  # relative difference per chain; a chain that is -inf in both runs (a failed evaluation) counts as identical
            with np.errstate(invalid="ignore"):
                d = np.abs(la - other) / np.abs(la)
            return np.where(np.isinf(la) & (la == other), 0.0, d)

        f_solver = rel(rb["logp"])
        f_rhs = rel(rc["logp"])
        f_ulp = np.zeros_like(la)
        steps_ulp = 1.0
        for direction in (np.inf, -np.inf):
            rd = a.cellpop_evaluate(prob, np.nextafter(gold["values"], direction), threads=1, want_steps=True)
            f_ulp = np.maximum(f_ulp, rel(rd["logp"]))
            steps_ulp = min(steps_ulp, (ra["cell_steps"] == rd["cell_steps"]).mean())
        floor = np.maximum(np.maximum(f_solver, f_rhs), f_ulp)
        m = ~np.isnan(ra["cell_values"])
        traj = max(np.abs(ra["cell_values"][m] - rb["cell_values"][m]).max(), np.abs(ra["cell_values"][m] - rc["cell_values"][m]).max())
        steps = min((ra["cell_steps"] == rb["cell_steps"]).mean(), (ra["cell_steps"] == rc["cell_steps"]).mean(), steps_ulp)
        resave(name, noise_floor=floor, noise_floor_solver=f_solver, noise_floor_rhs=f_rhs, noise_floor_ulp=f_ulp, noise_floor_trajectory=np.float64(traj),
               noise_floor_step_match=np.float64(steps))
        print(f"{name:40s} floor {floor.max():.2e} (solver flags {f_solver.max():.2e}, RHS flags {f_rhs.max():.2e}, 1-ulp inputs {f_ulp.max():.2e})  "
              f"max trajectory difference {traj:.2e}  cells with identical step counts {steps:.3f}")


if __name__ == "__main__":
    main()
