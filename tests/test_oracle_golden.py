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


def test_placement_rule_deviation_is_the_dropped_templates_and_nothing_else():
    """A11 is RACES-internal (unpinned).  The product draws a start uniformly over the fragment and drops a template
    that runs past its end; SURVEY.md Appendix A first had "uniform over the valid starts".  Both are selectable
    in the oracle: the difference is a coverage deficit of (template length - 1) / fragment length per fragment,
    concentrated within one read length of fragment ends (CNA breakpoints), and nothing else."""
    f = synth_forest(small_spec(2, chr_names=["1"], chr_len=[400_000], chr_n_alleles=[2], sample_cells=[10, 12],
                                cna_len=(4000, 30000), clone_cna=6))
    P = make_params(coverage=400.0, purity=1.0, read_size=150, seed=3)
    try:
        oracle.set_placement_rule(0)
        a = oracle.simulate(f, P, n_threads=4)
        oracle.set_placement_rule(1)
        b = oracle.simulate(f, P, n_threads=4)
    finally:
        oracle.set_placement_rule(0)
    want_reads = round(400.0 * 400_000 / 150) * 3
    assert b["n_reads"] == want_reads and a["n_reads"] < want_reads
    deficit = 1 - a["n_reads"] / b["n_reads"]
    # every molecule loses (R - 1) of its starts: a few 1e-4 here, ~5e-6 on a WGS-sized genome
    assert 1e-5 < deficit < 5e-3
    assert abs(a["cov"].mean() / b["cov"].mean() - 1) < 3 * deficit + 2e-3


GENERAL_CASES = [
    # sequencer, error rate, insert, extra
    (A.PCS_SEQ_ERRORLESS, 0.0, None, {}),
    (A.PCS_SEQ_BASIC_CONSTANT, 0.05, None, {}),
    (A.PCS_SEQ_BASIC_RANDOM, 0.05, None, {}),
    (A.PCS_SEQ_ERRORLESS, 0.0, (180, 9), {}),
    (A.PCS_SEQ_BASIC_CONSTANT, 0.05, (180, 9), {}),
    (A.PCS_SEQ_BASIC_RANDOM, 0.05, (180, 9), {}),
    (A.PCS_SEQ_BASIC_CONSTANT, 0.05, None, {"groups": True, "preneoplastic_in_normal": 1}),
    (A.PCS_SEQ_BASIC_RANDOM, 0.05, None, {"normal_only": 1, "preneoplastic_in_normal": 1}),
]


def general_case(seqm, rate, insert, extra):
    """(forest, params kwargs, closed-form kwargs, leaf_group, n_groups) of one GENERAL_CASES entry"""
    import closed_form as CF
    f = synth_forest(CF.indel_spec())
    extra = dict(extra)
    leaf_group = n_groups = None
    if extra.pop("groups", False):  # FACS-like repartition: every sample split in two by cell parity
        leaf_group = (f.leaf_sample * 2 + np.arange(f.n_leaves) % 2).astype(np.uint32)
        n_groups = 2 * f.n_samples
    pkw = dict(sequencer=seqm, error_rate=rate, **extra)
    ckw = dict(error_rate=rate, random_quality=seqm == A.PCS_SEQ_BASIC_RANDOM, leaf_group=leaf_group, n_groups=n_groups,
               preneoplastic_in_normal=bool(extra.get("preneoplastic_in_normal", 0)), normal_only=bool(extra.get("normal_only", 0)))
    if insert:
        pkw.update(insert_size_mean=insert[0], insert_size_stddev=insert[1])
        ckw["insert"] = insert
    if extra.get("normal_only"):
        pkw["with_normal_sample"] = 0
    return f, pkw, ckw, leaf_group, n_groups


@pytest.mark.parametrize("seqm,rate,insert,extra", GENERAL_CASES)
def test_oracle_matches_the_general_closed_form(seqm, rate, insert, extra):
    """SNVs AND indels (A12/A13 frame shifts, SIDs hidden inside a carried deletion), both error models (A14: the
    occurrence survives iff none of the SID's bases the read holds is an error; random-quality ramp), paired reads,
    FACS groups, normal_only: the expectations are written down by exhaustive enumeration of the starts on a token
    model of every haplotype (tests/closed_form.py::expected_tables_general) that shares nothing with the oracle's
    read walk.  All six sequencer x pairing variants."""
    import closed_form as CF
    f, pkw, ckw, leaf_group, n_groups = general_case(seqm, rate, insert, extra)
    assert ((f.mut_ref_len != 1) | (f.mut_alt_len != 1)).sum() > 100
    coverage, R, purity = 2000.0, 100, 0.7
    e_cov, e_occ = CF.expected_tables_general(f, coverage, purity, R, **ckw)
    r = oracle.simulate(f, make_params(coverage=coverage, purity=purity, read_size=R, seed=11, **pkw),
                        leaf_group=leaf_group, n_groups=n_groups, n_threads=4)
    assert r["cov"].shape == e_cov.shape
    for obs, exp in ((r["cov"], e_cov), (r["occ"], e_occ)):
        z, impossible = CF.z_scores(obs, exp)
        assert impossible == 0
        assert len(z) > 400 and abs(z.mean()) < 0.12 and 0.9 < z.std() < 1.1 and np.abs(z).max() < 5.0
        assert abs(obs.sum() / exp.sum() - 1) < 3e-3


def test_general_closed_form_reduces_to_the_snv_one():
    import closed_form as CF
    f = synth_forest(CF.snv_only_spec())
    for kw in (dict(), dict(insert=(180, 9)), dict(preneoplastic_in_normal=True)):
        a = CF.expected_tables(f, 3000.0, 0.7, 100, **kw)
        b = CF.expected_tables_general(f, 3000.0, 0.7, 100, **kw)
        assert np.allclose(a[0], b[0], rtol=0, atol=1e-6) and np.allclose(a[1], b[1], rtol=0, atol=1e-6)
