"""The GPU sampler against CLOSED-FORM expectations (tests/closed_form.py) instead of against the oracle's
Monte Carlo: E[depth] and E[occurrences] of every (sample, row) from the explicit genomes.

Written in round 1 after the GPU budget was spent, so it has never run on a B200.  Until one run has confirmed it,
it is marked xfail(strict=False): it runs with the rest of `-m gpu`, a pass shows as XPASS and a failure cannot
turn the suite red.  PCS_EXTRA_GPU_TESTS=1 makes it an ordinary test.  Its CPU twin, on the oracle, is
test_oracle_golden.py::test_oracle_matches_closed_form_expectations."""
import os

import numpy as np
import pytest

from process_b200 import _lib as L
from process_b200.synth import synth_forest

from conftest import make_params

pytestmark = [pytest.mark.gpu]
if os.environ.get("PCS_EXTRA_GPU_TESTS") != "1":
    pytestmark.append(pytest.mark.xfail(reason="first run on a B200: not yet confirmed", strict=False))


@pytest.mark.parametrize("purity,insert", [(0.7, None), (1.0, None), (0.7, (180, 9))])
def test_sampler_matches_closed_form_expectations(purity, insert):
    import closed_form as CF
    f = synth_forest(CF.snv_only_spec())
    coverage, R = 3000.0, 100
    e_cov, e_occ = CF.expected_tables(f, coverage, purity, R, insert=insert)
    kw = dict(insert_size_mean=insert[0], insert_size_stddev=insert[1]) if insert else {}
    ctx = L.Context(0)
    dev = L.Forest(ctx, f)
    occ, cov, st = dev.simulate(make_params(coverage=coverage, purity=purity, read_size=R, seed=11, **kw))
    dev.close()
    ctx.close()
    assert st.n_reads > 1_000_000
    for obs, exp in ((cov, e_cov), (occ, e_occ)):
        z, impossible = CF.z_scores(obs, exp)
        assert impossible == 0
        assert len(z) > 1500 and abs(z.mean()) < 0.1 and 0.9 < z.std() < 1.1 and np.abs(z).max() < 5.5
        assert abs(obs.sum() / exp.sum() - 1) < 2e-3
