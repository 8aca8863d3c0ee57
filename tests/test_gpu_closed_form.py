"""The GPU sampler against CLOSED-FORM expectations (tests/closed_form.py) instead of against the oracle's
Monte Carlo: E[depth] and E[occurrences] of every (sample, row), and E[reads] of every haplotype, from the
explicit genomes.  These are the tests that see the sampling LAW (the bit-exact recounts only see the counting):
round 1's biased haplotype draw failed two of them on the B200.  Their CPU twin, on the oracle, is
test_oracle_golden.py::test_oracle_matches_closed_form_expectations."""
import numpy as np
import pytest
from scipy import stats

from process_b200 import _abi as A
from process_b200 import _lib as L
from process_b200.synth import synth_forest

from conftest import make_params

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def ctx():
    c = L.Context(0)
    yield c
    c.close()


CASES = [(p, ins, pre) for p in (0.5, 0.7, 0.8, 0.9) for ins in (None, (180, 9)) for pre in (0, 1)] + [(1.0, None, 0)]


@pytest.mark.parametrize("purity,insert,preneo", CASES)
def test_sampler_matches_closed_form_expectations(ctx, purity, insert, preneo):
    import closed_form as CF
    f = synth_forest(CF.snv_only_spec())
    coverage, R = 3000.0, 100
    e_cov, e_occ = CF.expected_tables(f, coverage, purity, R, insert=insert, preneoplastic_in_normal=bool(preneo))
    kw = dict(insert_size_mean=insert[0], insert_size_stddev=insert[1]) if insert else {}
    dev = L.Forest(ctx, f)
    occ, cov, st = dev.simulate(make_params(coverage=coverage, purity=purity, read_size=R, seed=11,
                                            preneoplastic_in_normal=preneo, **kw))
    dev.close()
    assert st.n_reads > 1_000_000
    for obs, exp in ((cov, e_cov), (occ, e_occ)):
        z, impossible = CF.z_scores(obs, exp)
        assert impossible == 0
        assert len(z) > 1500 and abs(z.mean()) < 0.1 and 0.9 < z.std() < 1.1 and np.abs(z).max() < 5.5
        assert abs(obs.sum() / exp.sum() - 1) < 2e-3
    # heterozygous germline SNVs of a sample with contaminant cells: the two germline alleles of the normal cell
    # are drawn 1 : 1 (round 1: 55.6 : 44.4 at purity 0.8), so the pooled VAF sits on its expectation
    germ = np.zeros(f.n_mut, bool)
    germ[f.germ_mut[(f.germ_allele_mask == 1) | (f.germ_allele_mask == 2)]] = True
    for s in range(occ.shape[0]):
        want = e_occ[s, germ].sum() / e_cov[s, germ].sum()
        got = occ[s, germ].sum() / cov[s, germ].sum()
        assert abs(got / want - 1) < 4e-3, (s, got, want)


def test_cna_dense_forest_at_purity_one(ctx):
    """every CNA piece has its own sampling entries: the entry that does not own the whole draw range is the case
    the 32-bit scale got wrong even without contaminant cells"""
    import closed_form as CF
    f = synth_forest(CF.cna_dense_spec())
    coverage, R = 3000.0, 100
    e_cov, e_occ = CF.expected_tables(f, coverage, 1.0, R)
    dev = L.Forest(ctx, f)
    occ, cov, st = dev.simulate(make_params(coverage=coverage, purity=1.0, read_size=R, seed=5))
    dev.close()
    for obs, exp in ((cov, e_cov), (occ, e_occ)):
        z, impossible = CF.z_scores(obs, exp)
        assert impossible == 0
        assert len(z) > 1500 and abs(z.mean()) < 0.1 and 0.9 < z.std() < 1.1 and np.abs(z).max() < 5.5


@pytest.mark.parametrize("purity,preneo,spec", [(0.7, 0, "plain"), (0.8, 1, "plain"), (0.9, 0, "plain"), (1.0, 0, "cna"),
                                                (0.5, 0, "cna")])
def test_every_haplotype_gets_its_share_of_the_reads(ctx, purity, preneo, spec):
    """per-haplotype chi-square from pcs_plan_trace: reads placed on every (sample, chromosome, cell, allele)
    against N / W * w * (valid starts) -- uniform inside a class"""
    import closed_form as CF
    f = synth_forest(CF.snv_only_spec() if spec == "plain" else CF.cna_dense_spec())
    coverage, R = 1200.0, 100
    exp = CF.expected_haplotype_reads(f, coverage, purity, R, preneoplastic_in_normal=bool(preneo))
    dev = L.Forest(ctx, f)
    plan = L.Plan(dev, make_params(coverage=coverage, purity=purity, read_size=R, seed=3, preneoplastic_in_normal=preneo))
    occ, cov, st = plan.run()
    rec, _ = plan.trace(cap=int(st.n_reads) + 8)
    plan.close()
    dev.close()
    assert len(rec) == st.n_reads
    key = (rec["sample"].astype(np.uint64) << 48) | (rec["chr"].astype(np.uint64) << 40) | \
          (rec["flags"].astype(np.uint64) << 36) | (rec["cell"].astype(np.uint64) << 12) | rec["allele"].astype(np.uint64)
    ks, counts = np.unique(key, return_counts=True)
    got = dict(zip(ks.tolist(), counts.tolist()))
    obs, want = [], []
    for (s, c, kind, cell, a), e in exp.items():
        k = (s << 48) | (c << 40) | (kind << 36) | (cell << 12) | a
        o = got.pop(k, 0)
        if e == 0:
            assert o == 0
        elif e >= 50:
            obs.append(o)
            want.append(e)
    assert not got, "reads on haplotypes the explicit genomes do not have"
    obs, want = np.asarray(obs, float), np.asarray(want, float)
    assert len(obs) > 60
    chi2 = ((obs - want) ** 2 / want).sum()
    assert stats.chi2.sf(chi2, len(obs)) > 0.01, (chi2, len(obs))
    # and no haplotype is off by more than noise (the old draw lost up to half of the last haplotype's share)
    assert np.abs((obs - want) / np.sqrt(want)).max() < 5.0
    # the normal cell's germline alleles of chromosome "1" in sample 0: 1 : 1
    if purity < 1 and not preneo:
        n0 = exp[(0, 0, A.PCS_PLACE_NORMAL_PLAIN, 0, 0)]
        k0 = (0 << 48) | (0 << 40) | (A.PCS_PLACE_NORMAL_PLAIN << 36) | 0
        c0, c1 = dict(zip(ks.tolist(), counts.tolist()))[k0], dict(zip(ks.tolist(), counts.tolist()))[k0 | 1]
        assert n0 > 10_000 and abs(c0 / c1 - 1) < 6 / np.sqrt(n0)


def _general_cases():
    from test_oracle_golden import GENERAL_CASES
    return GENERAL_CASES


@pytest.mark.parametrize("seqm,rate,insert,extra", _general_cases())
def test_sampler_matches_the_general_closed_form(ctx, seqm, rate, insert, extra):
    """the GPU twin of test_oracle_golden.py::test_oracle_matches_the_general_closed_form: indels, both error
    models, paired reads, FACS groups, normal_only -- all six sequencer x pairing variants against expectations
    written down without any sampler"""
    import closed_form as CF
    from test_oracle_golden import general_case
    f, pkw, ckw, leaf_group, n_groups = general_case(seqm, rate, insert, extra)
    coverage, R, purity = 2000.0, 100, 0.7
    e_cov, e_occ = CF.expected_tables_general(f, coverage, purity, R, **ckw)
    dev = L.Forest(ctx, f)
    if leaf_group is not None:
        dev.set_groups(leaf_group, n_groups)
    occ, cov, st = dev.simulate(make_params(coverage=coverage, purity=purity, read_size=R, seed=17, **pkw))
    dev.close()
    assert cov.shape == e_cov.shape and st.n_reads > 500_000
    for obs, exp in ((cov, e_cov), (occ, e_occ)):
        z, impossible = CF.z_scores(obs, exp)
        assert impossible == 0
        assert len(z) > 400 and abs(z.mean()) < 0.12 and 0.9 < z.std() < 1.1 and np.abs(z).max() < 5.5
        assert abs(obs.sum() / exp.sum() - 1) < 3e-3
