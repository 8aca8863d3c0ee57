import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def small_spec(seed=0, **kw):
    from process_b200.synth import SynthSpec
    base = dict(chr_names=["1", "2", "X"], chr_len=[300_000, 200_000, 150_000], chr_n_alleles=[2, 2, 1],
                sample_cells=[6, 7, 8], germline_density=2e-3, n_preneo_snv=20, n_preneo_indel=10,
                node_snv_mean=5, n_clones=3, clone_cna=4, wgd_clones=2, cna_len=(5000, 60000), seed=seed)
    base.update(kw)
    return SynthSpec(**base)


@pytest.fixture(scope="session")
def small_forest():
    from process_b200.synth import synth_forest
    return synth_forest(small_spec(0))


def make_params(**kw):
    from process_b200 import _abi as A
    d = dict(seed=7, coverage=30.0, purity=1.0, read_size=150, insert_size_mean=0, insert_size_stddev=10,
             sequencer=A.PCS_SEQ_ERRORLESS, error_rate=0.0, with_normal_sample=1, preneoplastic_in_normal=0,
             normal_only=0, shard_rank=0, shard_count=1)
    d.update(kw)
    mask = d.pop("chr_mask", None)
    p = A.SeqParams(**d)
    if mask is not None:
        arr = np.ascontiguousarray(mask, dtype=np.uint8)
        p._mask_keep = arr
        p.chr_mask = A.ptr(arr, __import__("ctypes").c_uint8)
    return p
