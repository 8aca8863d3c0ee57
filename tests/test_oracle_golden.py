"""The CPU oracle against the hand-computed micro forest and the committed vectors.

The reference ships no tests or golden vectors for this path and its arithmetic
(RACES@1142937) is not in the image: parity is UNPINNED by the reference, so the
hand-computed known answers below are what pins the oracle."""
import json
import os

import numpy as np
import pytest

import oracle
from process_b200 import _abi as A
from process_b200.synth import synth_forest

from conftest import make_params, small_spec
from golden import micro_forest as MF

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_micro_forest_json_matches_its_script():
    with open(os.path.join(GOLD, "micro_forest.json")) as fh:
        j = json.load(fh)
    assert j["forest"] == MF.FOREST and j["expected_cov"] == MF.EXPECTED_COV and j["expected_occ"] == MF.EXPECTED_OCC
    assert [list(r[:6]) + [r[6]] for r in MF.READS] == [list(r[:6]) + [r[6]] for r in j["reads"]]


def test_micro_forest_explicit_genomes():
    f = MF.forest()
    which = {"tumour": A.PCS_PLACE_TUMOUR, "plain": A.PCS_PLACE_NORMAL_PLAIN, "preneo": A.PCS_PLACE_NORMAL_PRENEO}
    for key, exp in MF.EXPECTED_GENOMES.items():
        kind, cell = key.split(":")
        frags, sids = oracle.cell_genome(f, which[kind], int(cell), 0)
        assert sorted(frags) == sorted(tuple(x) for x in exp["frags"]), key
        assert sorted(sids) == sorted(tuple(x) for x in exp["sids"]), key


def test_micro_forest_hand_computed_counts():
    f = MF.forest()
    rec, masks = MF.placements()
    occ, cov = oracle.count_injected(f, 3, MF.READ_SIZE, rec, masks)
    assert occ.tolist() == MF.EXPECTED_OCC
    assert cov.tolist() == MF.EXPECTED_COV
    # each read alone, against the per-read contributions listed in the fixture's docstring
    occ1, cov1 = oracle.count_injected(f, 3, MF.READ_SIZE, rec[2:3], masks[2:3])  # read C
    assert occ1[0].tolist() == [0, 0, 1, 0, 0, 0, 0, 0] and cov1[0].tolist() == [0, 0, 1, 0, 0, 0, 0, 0]
    occ1, cov1 = oracle.count_injected(f, 3, MF.READ_SIZE, rec[6:7], masks[6:7])  # read E
    assert occ1[1].tolist() == [0, 0, 0, 0, 0, 1, 0, 0] and cov1[1].tolist() == [0, 0, 0, 0, 0, 1, 0, 0]


def test_micro_forest_hand_computed_sam_content():
    f = MF.forest()
    rec, masks = MF.placements()
    alt_off = np.concatenate([[0], np.cumsum([len(a) for a in MF.ALT])]).astype(np.uint32)
    seq, qual, cigar, nc, ln = oracle.materialize(f, np.asarray([0, 1000], np.uint64), MF.REF.encode(), alt_off,
                                                  "".join(MF.ALT).encode(), MF.READ_SIZE, A.PCS_SEQ_BASIC_CONSTANT, 1e-3,
                                                  rec, masks)
    for i, (want_seq, want_cigar) in MF.EXPECTED_SAM.items():
        got_cigar = "".join(f"{int(op) >> 4}{'MID'[int(op) & 3]}" for op in cigar[i, :nc[i]])
        assert seq[i, :ln[i]].tobytes().decode() == want_seq, i
        assert got_cigar == want_cigar, i
    # constant-quality model: Phred 30 everywhere except the erroneous base
    assert qual[5].tobytes().decode() == "????#?????" and qual[0].tobytes().decode() == "?" * 10
    # reads clamped by the end of their fragment (H, J) are shorter than read_size
    assert ln[9] == 5 and ln[11] == 3


@pytest.mark.parametrize("name", ["errorless_single", "random_quality_paired"])
def test_oracle_reproduces_committed_vectors(name):
    z = np.load(os.path.join(GOLD, f"injected_{name}.npz"))
    f = synth_forest(small_spec(int(z["forest_seed"])))
    masks = np.zeros((len(z["trace"]), A.PCS_ERRMASK_WORDS), np.uint32)
    masks[z["mask_rows"]] = z["mask_vals"]
    occ, cov = oracle.count_injected(f, z["occ"].shape[0], int(z["read_size"]), z["trace"], masks)
    assert np.array_equal(occ, z["occ"]) and np.array_equal(cov, z["cov"])


def test_oracle_recount_of_its_own_reads_is_identical():
    f = synth_forest(small_spec(2))
    for kw in (dict(), dict(sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=0.05, insert_size_mean=150),
               dict(normal_only=1, with_normal_sample=0, preneoplastic_in_normal=1)):
        P = make_params(coverage=6.0, purity=0.6, **kw)
        r = oracle.simulate(f, P, trace_cap=300_000, trace_masks=True)
        occ, cov = oracle.count_injected(f, r["occ"].shape[0], P.read_size, r["trace"], r["masks"])
        assert np.array_equal(occ, r["occ"]) and np.array_equal(cov, r["cov"])


def test_oracle_coverage_purity_and_errors_behave():
    f = synth_forest(small_spec(3, chr_names=["1"], chr_len=[400_000], chr_n_alleles=[2], wgd_clones=0, n_clones=0,
                                clone_cna=0))
    P = make_params(coverage=60.0, purity=1.0)
    r = oracle.simulate(f, P)
    # mean depth over loci equals the requested coverage (A9), per sample
    assert np.allclose(r["cov"].mean(axis=1), 60.0, rtol=0.03)
    germ_het = np.zeros(f.n_mut, bool)
    germ_het[f.germ_mut[f.germ_allele_mask != 3]] = True
    germ_hom = np.zeros(f.n_mut, bool)
    germ_hom[f.germ_mut[f.germ_allele_mask == 3]] = True
    vaf = r["occ"].sum(axis=0) / np.maximum(r["cov"].sum(axis=0), 1)
    assert abs(vaf[germ_het].mean() - 0.5) < 0.02 and vaf[germ_hom].min() == 1.0
    # sequencing errors only ever remove occurrences
    Pe = make_params(coverage=60.0, purity=1.0, sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=0.2)
    re_ = oracle.simulate(f, Pe)
    vaf_e = re_["occ"].sum(axis=0) / np.maximum(re_["cov"].sum(axis=0), 1)
    snv = (f.mut_ref_len == 1) & (f.mut_alt_len == 1)
    assert abs(vaf_e[germ_hom & snv].mean() - 0.8) < 0.02
    # random quality model: same mean error rate
    Pq = make_params(coverage=60.0, purity=1.0, sequencer=A.PCS_SEQ_BASIC_RANDOM, error_rate=0.2)
    rq = oracle.simulate(f, Pq)
    vaf_q = rq["occ"].sum(axis=0) / np.maximum(rq["cov"].sum(axis=0), 1)
    assert abs(vaf_q[germ_hom & snv].mean() - 0.8) < 0.03
    # the normal sample never shows somatic SIDs; purity 0.5 halves tumour-only VAFs
    assert r["occ"][-1][~(germ_het | germ_hom)].sum() == 0


def test_oracle_rejects_malformed_input():
    f = MF.forest()
    bad = np.zeros(1, A.PLACEMENT_DTYPE)
    bad["start"] = 100
    bad["allele"] = 7
    with pytest.raises(oracle.OracleError):
        oracle.count_injected(f, 1, 10, bad)
    bad["allele"] = 0
    bad["cell"] = 1
    bad["start"] = 100  # cell 1 lost [1,150] of allele 0
    with pytest.raises(oracle.OracleError):
        oracle.count_injected(f, 1, 10, bad)
    with pytest.raises(oracle.OracleError):
        oracle.simulate(f, make_params(insert_size_mean=10, insert_size_stddev=10))  # sd^2 > mean


@pytest.mark.parametrize("purity,error_rate,insert,preneo", [(0.7, 0.0, None, 0), (1.0, 0.0, None, 0), (0.7, 0.1, None, 0),
                                                             (0.7, 0.0, (180, 9), 0), (0.5, 0.0, None, 1)])
def test_oracle_matches_closed_form_expectations(purity, error_rate, insert, preneo):
    """E[depth] and E[occurrences] of every (sample, row), written down in closed form from the explicit genomes
    (tests/closed_form.py: amplified, deleted and WGD-doubled alleles, purity, the normal sample, reads that fall
    off a fragment end, constant-quality sequencing errors, paired reads with the Binomial insert law), against a
    3000x oracle run: Poisson z-scores centred, unit variance, no outlier, and
    nothing counted where the expectation is exactly zero"""
    import closed_form as CF
    f = synth_forest(CF.snv_only_spec())
    assert (f.ev_kind == A.PCS_EV_WGD).sum() >= 1 and (f.ev_kind == A.PCS_EV_CNA_AMP).sum() >= 1 and \
        (f.ev_kind == A.PCS_EV_CNA_DEL).sum() >= 1
    coverage, R = 3000.0, 100
    e_cov, e_occ = CF.expected_tables(f, coverage, purity, R, insert=insert, preneoplastic_in_normal=bool(preneo))
    kw = dict(sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=error_rate) if error_rate else {}
    kw["preneoplastic_in_normal"] = preneo
    if insert:
        kw.update(insert_size_mean=insert[0], insert_size_stddev=insert[1])
    r = oracle.simulate(f, make_params(coverage=coverage, purity=purity, read_size=R, seed=11, **kw), n_threads=4)
    # constant-quality errors: an SNV occurrence survives with probability 1 - error_rate; depth is unaffected
    for obs, exp in ((r["cov"], e_cov), (r["occ"], e_occ * (1 - error_rate))):
        z, impossible = CF.z_scores(obs, exp)
        assert impossible == 0
        assert len(z) > 1500 and abs(z.mean()) < 0.1 and 0.9 < z.std() < 1.1 and np.abs(z).max() < 5.0
        assert abs(obs.sum() / exp.sum() - 1) < 2e-3
