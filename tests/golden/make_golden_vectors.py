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

def distribution_fixture():
    """oracle free-run depth / occurrence tables of one fixed forest: the GPU sampler's
    tables must be statistically indistinguishable from these (KS on depth and VAF)."""
    f = synth_forest(small_spec(7, chr_names=["1"], chr_len=[2_000_000], chr_n_alleles=[2], sample_cells=[30, 50],
                                germline_density=1.5e-3, cna_len=(50_000, 300_000)))
    P = make_params(coverage=100.0, purity=0.8, sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=0.01, seed=21)
    r = oracle.simulate(f, P, n_threads=8)
    np.savez_compressed(os.path.join(HERE, "distribution_oracle.npz"), occ=r["occ"].astype(np.uint16),
                        cov=r["cov"].astype(np.uint16), n_reads=r["n_reads"], forest_seed=7)
    print("distribution fixture", r["n_reads"], "reads", r["cov"].mean())


if __name__ == "__main__":
    distribution_fixture()
    for name, case in CASES.items():
        f = synth_forest(small_spec(case["seed"]))
        P = make_params(**case["params"])
        r = oracle.simulate(f, P, trace_cap=400_000, trace_masks=True)
        nz = np.flatnonzero(r["masks"].any(axis=1))
        np.savez_compressed(os.path.join(HERE, f"injected_{name}.npz"), trace=r["trace"],
                            mask_rows=nz.astype(np.uint32), mask_vals=r["masks"][nz],
                            occ=r["occ"], cov=r["cov"], forest_seed=case["seed"], read_size=P.read_size)
        print(name, len(r["trace"]), "reads", int(r["occ"].sum()), "occurrences")
