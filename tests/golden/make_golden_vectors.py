"""Regenerates tests/golden/injected_*.npz: seeded read-placement lists and the count
tables the CPU oracle gives for them.  The CUDA path must
reproduce the tables bit for bit from the same placements; the oracle must keep
reproducing them (regression pin).   python tests/golden/make_golden_vectors.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from conftest import make_params, small_spec  # noqa: E402
from process_b200 import _abi as A  # noqa: E402
from process_b200.synth import synth_forest  # noqa: E402

CASES = {
    "errorless_single": dict(seed=0, params=dict(coverage=4.0, purity=0.7)),
    "random_quality_paired": dict(seed=1, params=dict(coverage=4.0, purity=0.5, insert_size_mean=180,
                                                      insert_size_stddev=9, sequencer=A.PCS_SEQ_BASIC_RANDOM,
                                                      error_rate=0.03, preneoplastic_in_normal=1)),
}

if __name__ == "__main__":
    for name, case in CASES.items():
        f = synth_forest(small_spec(case["seed"]))
        P = make_params(**case["params"])
        r = oracle.simulate(f, P, trace_cap=400_000, trace_masks=True)
        nz = np.flatnonzero(r["masks"].any(axis=1))
        np.savez_compressed(os.path.join(HERE, f"injected_{name}.npz"), trace=r["trace"],
                            mask_rows=nz.astype(np.uint32), mask_vals=r["masks"][nz],
                            occ=r["occ"], cov=r["cov"], forest_seed=case["seed"], read_size=P.read_size)
        print(name, len(r["trace"]), "reads", int(r["occ"].sum()), "occurrences")
