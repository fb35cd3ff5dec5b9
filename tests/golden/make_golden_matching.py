"""Generates tests/golden/matching_cases.npz: complete cost matrices and the matching the REFERENCE's own compiled
hungarianMinimumWeightPerfectMatching (dependencies/hungarian2/hungarian.cpp, linked into oracle/_ref) returns for them, with
the edge list built as DataLikelihoodTimeCourse::Evaluate builds it. The matrices cover what the per-cell likelihood produces
(negated sums of log-densities: tens to thousands, negative and positive), ties and the sizes 1..40. Run where /root/reference
is mounted:  python tests/golden/make_golden_matching.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402


def cases(seed=20261019, count=90):
    rng = np.random.default_rng(seed)
    out = []
    for trial in range(count):
        n = int(rng.integers(1, 41))
        kind = trial % 5
        if kind == 0:
            cost = rng.normal(0.0, 5.0, (n, n))
        elif kind == 1:
            cost = rng.uniform(-300.0, 50.0, (n, n))
        elif kind == 2:
            cost = np.round(rng.normal(0.0, 3.0, (n, n)))  # many ties
        elif kind == 3:
            cost = rng.normal(-200.0, 0.5, (n, n)) * rng.uniform(0.1, 10.0)
        else:  # a planted assignment under noise: what matching observed to simulated trajectories looks like
            perm = rng.permutation(n)
            cost = rng.uniform(20.0, 400.0, (n, n))
            cost[np.arange(n), perm] = rng.uniform(-60.0, -20.0, n)
        out.append(cost)
    return out


def main():
    ref = oracle.load("ref")
    mats = cases()
    sizes = np.array([m.shape[0] for m in mats], dtype=np.int32)
    flat = np.concatenate([m.ravel() for m in mats])
    matches = np.concatenate([ref.hungarian_match(m) for m in mats]).astype(np.int32)
    from scipy.optimize import linear_sum_assignment

    suboptimal = 0
    for m, n in zip(mats, sizes):
        got = ref.hungarian_match(m)
        r, c = linear_sum_assignment(m)
        suboptimal += m[np.arange(n), got].sum() > m[r, c].sum() + 1e-9
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "matching_cases.npz"), sizes=sizes, costs=flat, matches=matches,
                        suboptimal=np.int32(suboptimal))
    print(len(mats), "matrices;", int(suboptimal), "of the reference's matchings are not the minimum-cost matching")


if __name__ == "__main__":
    main()
