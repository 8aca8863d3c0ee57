"""BASELINE.json configurations at full size on the GPU, checked through size-independent
properties (the oracle cannot expand these in seconds): shard additivity, mean coverage,
germline VAF laws, purity scaling, GPU tables == oracle recount of one traced shard."""
import numpy as np
import pytest

import oracle
from process_b200 import _abi as A
from process_b200 import _lib as L
from process_b200.synth import config_spec, synth_forest

from conftest import make_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = L.Context(0)
    yield c
    c.close()


def germline_sets(f):
    het = np.zeros(f.n_mut, bool)
    hom = np.zeros(f.n_mut, bool)
    two = f.chr_n_alleles[f.mut_chr[f.germ_mut]] == 2
    het[f.germ_mut[(f.germ_allele_mask != 3) & two]] = True
    hom[f.germ_mut[(f.germ_allele_mask == 3) | ~two]] = True
    return het, hom


def test_config1_demo_chr22_50x_full_size(ctx):
    """configs[0]: chr22-sized genome, 4 samples (~1.3k cells), errorless, 50x, + normal sample: 85.5 M reads."""
    f = synth_forest(config_spec("C1"))
    dev = L.Forest(ctx, f)
    P = make_params(coverage=50.0, purity=1.0, seed=0)
    occ, cov, st = dev.simulate(P)
    n_out = 5
    assert occ.shape == (n_out, f.n_mut)
    assert abs(st.n_reads - n_out * round(50.0 * int(f.chr_len[0]) / 150)) <= n_out * 200
    assert np.allclose(cov.mean(axis=1), 50.0, rtol=5e-3)
    het, hom = germline_sets(f)
    snv = (f.mut_ref_len == 1) & (f.mut_alt_len == 1)
    normal = n_out - 1
    # errorless: a homozygous germline SNV is on every read that spans it; het ones on half of them
    assert np.array_equal(occ[normal, hom & snv], cov[normal, hom & snv])
    vaf_het = occ[normal, het].sum() / cov[normal, het].sum()
    assert abs(vaf_het - 0.5) < 2e-3
    somatic = (f.mut_nature_mask & ((1 << A.PCS_NATURE_DRIVER) | (1 << A.PCS_NATURE_PASSENGER) |
                                    (1 << A.PCS_NATURE_PRENEOPLASTIC))) != 0
    assert occ[normal, somatic].sum() == 0
    # pre-neoplastic SIDs sit on the trunk: clonal in every tumour sample, never above depth
    assert (occ <= cov).all()
    pre = ((f.mut_nature_mask >> A.PCS_NATURE_PRENEOPLASTIC) & 1) == 1
    assert occ[:4, pre].sum() > 0
    # 8 shards add up to the whole, bit for bit
    occ_s, cov_s = np.zeros_like(occ), np.zeros_like(cov)
    for r in range(8):
        o, c, _ = dev.simulate(make_params(coverage=50.0, purity=1.0, seed=0, shard_rank=r, shard_count=8))
        occ_s += o
        cov_s += c
    assert np.array_equal(occ_s, occ) and np.array_equal(cov_s, cov)
    # one shard, traced and recounted by the oracle on explicit genomes
    Pr = make_params(coverage=50.0, purity=1.0, seed=0, shard_rank=3, shard_count=8)
    plan = L.Plan(dev, Pr)
    o, c, s = plan.run()
    rec, _ = plan.trace(cap=int(s.n_reads) + 8)
    o2, c2 = oracle.count_injected(f, n_out, 150, rec)
    assert np.array_equal(o, o2) and np.array_equal(c, c2)
    plan.close()
    dev.close()


def test_config2_basic_illumina_200x_purity(ctx):
    """configs[1]: same forest, BasicIlluminaSequencer(1e-3), 200x, purity 0.8, + simulate_normal_seq."""
    f = synth_forest(config_spec("C1"))
    dev = L.Forest(ctx, f)
    het, hom = germline_sets(f)
    snv = (f.mut_ref_len == 1) & (f.mut_alt_len == 1)
    res = {}
    for kind in (A.PCS_SEQ_BASIC_CONSTANT, A.PCS_SEQ_BASIC_RANDOM):
        occ, cov, st = dev.simulate(make_params(coverage=200.0, purity=0.8, sequencer=kind, error_rate=1e-3, seed=1))
        assert np.allclose(cov.mean(axis=1), 200.0, rtol=5e-3)
        # an error on the SNV base hides the occurrence: VAF of homozygous germline SNVs = 1 - error_rate
        miss = 1.0 - occ[-1, hom & snv].sum() / cov[-1, hom & snv].sum()
        assert abs(miss - 1e-3) < 1.5e-4, miss
        res[kind] = occ
    # purity: pre-neoplastic (clonal, tumour only) VAF drops by the tumour DNA share; the normal sample has none
    occ1, cov1, _ = dev.simulate(make_params(coverage=200.0, purity=1.0, seed=2))
    occ8, cov8, _ = dev.simulate(make_params(coverage=200.0, purity=0.8, seed=2))
    pre = (((f.mut_nature_mask >> A.PCS_NATURE_PRENEOPLASTIC) & 1) == 1) & snv
    # Closed form (A9/A10): the templates of a chromosome are split over the sample's DNA in proportion to length x
    # cell weight (tumour cell p / n_T, the normal cell 1 - p), so the reads that carry a tumour-only SID scale by
    # p L_T / (p L_T + (1 - p) L_N) whatever the locus -- L_T: mean DNA of the sample's tumour cells on the
    # chromosome (CNAs and WGD included, from the flattened view), L_N = 2 x chr_len.
    flat = L.Flat(f)
    frag_len = {}
    def dna(cell):
        total = 0
        for _, _, fs in flat.cell_haps(0, cell, 0):
            if fs not in frag_len:
                frag_len[fs] = sum(e - b + 1 for b, e in flat.fragset(fs))
            total += frag_len[fs]
        return total
    L_N = 2.0 * int(f.chr_len[0])
    for s in range(4):
        cells = np.flatnonzero(f.leaf_sample == s)
        L_T = float(np.mean([dna(int(c)) for c in cells]))
        want = 0.8 * L_T / (0.8 * L_T + 0.2 * L_N)
        got = occ8[s, pre].sum() / occ1[s, pre].sum()
        assert abs(got / want - 1) < 0.015, (s, got, want, L_T / L_N)
        v1 = occ1[s, pre].sum() / cov1[s, pre].sum()
        v8 = occ8[s, pre].sum() / cov8[s, pre].sum()
        assert 0.75 * v1 < v8 < 0.92 * v1, (s, v1, v8)
    del flat
    assert occ8[-1, pre].sum() == 0
    # simulate_normal_seq: one sample, germline only
    occn, covn, stn = dev.simulate(make_params(coverage=200.0, normal_only=1, with_normal_sample=0,
                                               sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=1e-3, seed=3))
    assert occn.shape[0] == 1 and abs(covn.mean() / 200.0 - 1) < 5e-3
    assert occn[0, ~(het | hom)].sum() == 0
    dev.close()


def test_config3_wgs_80x_full_size(ctx):
    """configs[2] (the bench workload): GRCh38-length genome, 3 x 1000 cells + normal, 80x: 6.59 G reads."""
    f = synth_forest(config_spec("C3"))
    dev = L.Forest(ctx, f)
    P = make_params(coverage=80.0, purity=1.0, seed=0)
    plan = L.Plan(dev, P)
    occ, cov, st = plan.run()
    want = sum(round(80.0 * int(n) / 150) for n in f.chr_len) * 4
    assert abs(st.n_reads - want) < 1e-5 * want
    assert st.sum_occurrences == int(occ.sum(dtype=np.uint64))
    # per chromosome and sample the mean depth is the requested coverage (A9), X/Y (one allele) included
    for c in range(f.n_chr):
        rows = f.mut_chr == c
        assert np.allclose(cov[:, rows].mean(axis=1), 80.0, rtol=2e-2), c
    assert abs(cov.mean() / 80.0 - 1) < 5e-3
    het, hom = germline_sets(f)
    snv = (f.mut_ref_len == 1) & (f.mut_alt_len == 1)
    assert np.array_equal(occ[-1, hom & snv], cov[-1, hom & snv])
    assert abs(occ[-1, het].sum() / cov[-1, het].sum() - 0.5) < 5e-4
    assert (occ <= cov).all()
    # same plan twice: identical tables; two halves add up
    occ2, cov2, _ = plan.run()
    assert np.array_equal(occ, occ2) and np.array_equal(cov, cov2)
    acc_o, acc_c = np.zeros_like(occ), np.zeros_like(cov)
    for r in range(2):
        o, c, _ = dev.simulate(make_params(coverage=80.0, purity=1.0, seed=0, shard_rank=r, shard_count=2))
        acc_o += o
        acc_c += c
    assert np.array_equal(acc_o, occ) and np.array_equal(acc_c, cov)
    plan.close()
    dev.close()


def test_config4_eight_samples_200x_sharded_eight_ways(ctx):
    """configs[3]: 8 samples x 5000 cells, 200x WGS (37 G reads), as its 8 shards on one GPU: the shards must
    add up to the unsharded tables, and one traced shard of one chromosome must survive the oracle's recount."""
    f = synth_forest(config_spec("C4"))
    dev = L.Forest(ctx, f)
    kw = dict(coverage=200.0, purity=0.9, seed=4)
    occ, cov, st = dev.simulate(make_params(**kw))
    assert occ.shape[0] == 9
    want = sum(round(200.0 * int(n) / 150) for n in f.chr_len) * 9
    assert abs(st.n_reads - want) < 1e-5 * want
    assert abs(cov.mean() / 200.0 - 1) < 5e-3
    acc_o, acc_c, reads = np.zeros_like(occ), np.zeros_like(cov), 0
    for r in range(8):
        o, c, s = dev.simulate(make_params(shard_rank=r, shard_count=8, **kw))
        acc_o += o
        acc_c += c
        reads += s.n_reads
        assert abs(s.n_reads / (st.n_reads / 8) - 1) < 0.05  # LPT balance (on cost, not reads): near-linear scaling is possible
    assert reads == st.n_reads
    assert np.array_equal(acc_o, occ) and np.array_equal(acc_c, cov)
    het, hom = germline_sets(f)
    assert abs(occ[-1, het].sum() / cov[-1, het].sum() - 0.5) < 5e-4
    # chromosome 21 only, shard 5 of 64, traced and recounted on explicit genomes of all 40 000 cells
    mask = np.zeros(f.n_chr, np.uint8)
    mask[f.chr_names.index("21")] = 1
    Pr = make_params(chr_mask=mask, shard_rank=5, shard_count=64, **kw)
    plan = L.Plan(dev, Pr)
    o, c, s = plan.run()
    rec, _ = plan.trace(cap=int(s.n_reads) + 8)
    assert len(rec) == s.n_reads > 1_000_000
    o2, c2 = oracle.count_injected(f, 9, 150, rec)
    assert np.array_equal(o, o2) and np.array_equal(c, c2)
    plan.close()
    dev.close()


def test_config5_stress_1e5_cells_wgd_300x(ctx):
    """configs[4]: 100 000 cells, WGD in every clone, 200 CNAs, ~1e6 SNVs per genome, tumour + normal at 300x.
    Explicit genomes are out of reach (1e11 SIDs); the haplotype-interval view holds it in 0.4 GB."""
    f = synth_forest(config_spec("C5"))
    dev = L.Forest(ctx, f)
    info = dev.info()
    assert info["n_haplotypes"] > 10_000_000 and info["device_bytes"] < 1 << 30
    P = make_params(coverage=300.0, purity=0.9, seed=5, sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=1e-3)
    occ, cov, st = dev.simulate(P)
    want = sum(round(300.0 * int(n) / 150) for n in f.chr_len) * 2
    assert abs(st.n_reads - want) < 1e-5 * want
    assert np.allclose(cov.mean(axis=1), 300.0, rtol=5e-3)
    assert (occ <= cov).all()
    het, hom = germline_sets(f)
    snv = (f.mut_ref_len == 1) & (f.mut_alt_len == 1)
    # normal sample: germline laws with the 1e-3 error rate
    assert abs(occ[1, het & snv].sum() / cov[1, het & snv].sum() - 0.5 * (1 - 1e-3)) < 5e-4
    assert abs(occ[1, hom & snv].sum() / cov[1, hom & snv].sum() - (1 - 1e-3)) < 2e-4
    somatic = (f.mut_nature_mask & ((1 << A.PCS_NATURE_DRIVER) | (1 << A.PCS_NATURE_PASSENGER) |
                                    (1 << A.PCS_NATURE_PRENEOPLASTIC))) != 0
    assert occ[1, somatic].sum() == 0
    # tumour sample: trunk passengers are clonal, heterozygous before any copy number change:
    # VAF well inside (0.2, 0.6) at purity 0.9 whatever WGD/CNAs did afterwards
    trunk = np.zeros(f.n_mut, bool)
    root_events = slice(int(f.node_event_off[0]), int(f.node_event_off[1]))
    trunk[f.ev_mut[root_events][f.ev_kind[root_events] == A.PCS_EV_SID]] = True
    vaf = occ[0, trunk & snv].sum() / cov[0, trunk & snv].sum()
    assert 0.2 < vaf < 0.6, vaf
    # shards add up
    acc_o, acc_c = np.zeros_like(occ), np.zeros_like(cov)
    for r in range(4):
        Pr = make_params(coverage=300.0, purity=0.9, seed=5, sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=1e-3,
                         shard_rank=r, shard_count=4)
        o, c, _ = dev.simulate(Pr)
        acc_o += o
        acc_c += c
    assert np.array_equal(acc_o, occ) and np.array_equal(acc_c, cov)
    dev.close()
