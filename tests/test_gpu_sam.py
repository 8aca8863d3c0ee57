"""SAM output (write_SAM = TRUE): reads materialised on the GPU against the oracle's materialisation of the
same placements, the files' structure, CREATE / UPDATE directory semantics, tables == what the files hold."""
import os

import numpy as np
import pytest

import oracle
from process_b200 import _abi as A
from process_b200 import _lib as L
from process_b200 import api
from process_b200.synth import read_fasta, synth_forest, write_reference_fasta

from conftest import make_params, small_spec

pytestmark = pytest.mark.gpu


def sequences(f, fasta):
    ref = read_fasta(fasta)
    off = np.zeros(f.n_chr + 1, np.uint64)
    parts = []
    for c, name in enumerate(f.chr_names):
        off[c] = sum(len(p) for p in parts)
        parts.append(ref[name])
    off[f.n_chr] = sum(len(p) for p in parts)
    return off, "".join(parts).encode()


@pytest.mark.parametrize("seqm,rate,insert", [(A.PCS_SEQ_ERRORLESS, 0.0, 0), (A.PCS_SEQ_BASIC_CONSTANT, 0.02, 0),
                                               (A.PCS_SEQ_BASIC_RANDOM, 0.02, 170), (A.PCS_SEQ_BASIC_CONSTANT, 0.05, 120)])
def test_materialised_reads_match_oracle(tmp_path, seqm, rate, insert):
    f = synth_forest(small_spec(2, indel_frac=0.3, germline_indel_frac=0.2))
    fasta = write_reference_fasta(f, str(tmp_path / "ref.fa"), seed=1)
    ctx = L.Context(0)
    dev = L.Forest(ctx, f)
    assert dev.load_fasta(fasta) == f.n_chr
    alt_off, alt_bytes = f.alt_table()
    dev.set_alt(alt_off, alt_bytes)
    P = make_params(coverage=6.0, purity=0.7, sequencer=seqm, error_rate=rate, insert_size_mean=insert, read_size=100,
                    preneoplastic_in_normal=1)
    plan = L.Plan(dev, P)
    occ, cov, st = plan.run()
    rec, masks, seq, qual, cigar, nc, ln = plan.materialize(cap=int(st.n_reads) + 16)
    assert len(rec) == st.n_reads > 10_000
    # the materialised reads are the counted reads: the oracle's recount of them gives the tables
    occ2, cov2 = oracle.count_injected(f, plan.info.n_out_samples, P.read_size, rec, masks)
    assert np.array_equal(occ, occ2) and np.array_equal(cov, cov2)
    # and their content is what the explicit genomes say
    ref_off, ref_bytes = sequences(f, fasta)
    oseq, oqual, ocig, onc, oln = oracle.materialize(f, ref_off, ref_bytes, alt_off, alt_bytes, P.read_size, seqm, rate,
                                                     rec, masks)
    assert np.array_equal(ln, oln) and (ln == P.read_size).all()
    assert np.array_equal(nc, onc) and np.array_equal(cigar, ocig)
    assert np.array_equal(seq, oseq)
    if seqm != A.PCS_SEQ_BASIC_RANDOM:
        assert np.array_equal(qual, oqual)
    else:
        q = qual.astype(np.int32) - 33
        assert q.min() >= 2 and q.max() <= 41 and abs((10.0 ** (-q / 10.0)).mean() / rate - 1) < 0.25
    assert (nc > 1).sum() > 100  # indels were exercised
    err_frac = np.unpackbits(masks.view(np.uint8), bitorder="little").sum() / (len(rec) * P.read_size)
    assert (rate == 0 and err_frac == 0) or abs(err_frac / rate - 1) < 0.1
    plan.close()
    dev.close()
    ctx.close()


def parse_sam(path):
    header, reads = [], []
    with open(path) as fh:
        for ln in fh:
            (header if ln.startswith("@") else reads).append(ln.rstrip("\n").split("\t"))
    return header, reads


def test_write_sam_files_and_directory_modes(tmp_path):
    f = synth_forest(small_spec(3))
    write_reference_fasta(f, str(tmp_path / "ref.fa"), seed=2)
    out = str(tmp_path / "ProCESS_SAM")
    r = api.simulate_seq(f, coverage=2.5, write_SAM=True, output_dir=out, seed=3, purity=0.8,
                         sequencer=api.BasicIlluminaSequencer(1e-2, False))
    files = sorted(os.listdir(out))
    assert files == sorted(f"chr_{n}.sam" for n in f.chr_names)
    names = f.sample_names + ["normal_sample"]
    total = 0
    per_sample = {n: 0 for n in names}
    ref = read_fasta(f.reference_path)
    for c, name in enumerate(f.chr_names):
        header, reads = parse_sam(os.path.join(out, f"chr_{name}.sam"))
        assert header[0][0] == "@HD" and header[1] == ["@SQ", f"SN:{name}", f"LN:{int(f.chr_len[c])}"]
        assert [h[1] for h in header if h[0] == "@RG"] == [f"ID:{n}" for n in names]
        total += len(reads)
        qnames = set()
        for rd in reads[:5000]:
            assert len(rd) == 12 and rd[2] == name and rd[1] == "0" and rd[6:9] == ["*", "0", "0"]
            assert len(rd[9]) == len(rd[10]) == 150
            per = rd[11]
            assert per.startswith("RG:Z:") and per[5:] in per_sample
            qnames.add(rd[0])
            if rd[5] == "150M":  # no indel: at most a few mismatches against the reference
                pos = int(rd[3])
                mism = sum(a != b for a, b in zip(rd[9], ref[name][pos - 1:pos + 149]))
                assert mism <= 12
        assert len(qnames) == min(len(reads), 5000) and all(q.startswith("r") for q in qnames)
        for rd in reads:
            per_sample[rd[11][5:]] += 1
    assert total == r["_stats"]["n_reads"]
    want = sum(round(2.5 * int(n) / 150) for n in f.chr_len)
    assert all(0.97 * want <= v <= want for v in per_sample.values())  # templates falling off a fragment end are dropped
    # CREATE refuses an existing directory, UPDATE numbers the new files (vignettes/sequencing.Rmd:283-309)
    with pytest.raises(ValueError, match="already exists"):
        api.simulate_seq(f, coverage=1.0, write_SAM=True, output_dir=out, seed=4)
    api.simulate_seq(f, coverage=1.0, write_SAM=True, update_SAM=True, output_dir=out, seed=4, chromosomes=["1"])
    assert "chr_1_1.sam" in os.listdir(out)
    # paired reads: flags, mates, template length
    out2 = str(tmp_path / "paired")
    api.simulate_seq(f, coverage=30.0, write_SAM=True, output_dir=out2, seed=5, insert_size_mean=200, insert_size_stddev=10,
                     chromosomes=["2"], with_normal_sample=False, filename_prefix="x_", template_name_prefix="tpl")
    _, reads = parse_sam(os.path.join(out2, "x_2.sam"))
    by_name = {}
    for rd in reads:
        by_name.setdefault(rd[0], []).append(rd)
    assert all(len(v) == 2 and k.startswith("tpl") for k, v in by_name.items())
    for a, b in list(by_name.values())[:2000]:
        first, second = (a, b) if a[1] == "99" else (b, a)
        assert first[1] == "99" and second[1] == "147" and first[6] == "="
        assert first[7] == second[3] and second[7] == first[3]
        assert int(first[8]) == -int(second[8]) == int(second[3]) - int(first[3]) + 150
    # insert sizes follow get_bin_dist(200, 10): Binomial(t = 200/p, p = 1 - 100/200)  (src/seq_simulation.cpp:431-451)
    from scipy import stats
    ins = np.asarray([int(a[8] if a[1] == "99" else b[8]) - 300 for a, b in by_name.values()])
    t, pr = int(200 / 0.5), 0.5
    assert len(ins) > 15_000 and abs(ins.mean() - t * pr) < 0.35 and abs(ins.std() - 10.0) < 0.3
    obs = np.bincount(ins, minlength=t + 1)[150:251]
    exp = stats.binom(t, pr).pmf(np.arange(150, 251)) * len(ins)
    keep = exp > 5
    chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    assert stats.chi2(keep.sum() - 1).sf(chi2) > 1e-3
    # simulate_normal_seq writes SAM by default (src/sequencing.cpp:275-276)
    n = api.simulate_normal_seq(f, coverage=2.0, seed=6, output_dir=str(tmp_path / "normal"))
    assert sorted(os.listdir(str(tmp_path / "normal"))) == sorted(f"chr_{x}.sam" for x in f.chr_names)
    hdr, _ = parse_sam(str(tmp_path / "normal" / "chr_1.sam"))
    assert [h[1] for h in hdr if h[0] == "@RG"] == ["ID:normal_sample"]
    assert n["parameters"]["write_SAM"] is True
    api.release_device_cache()
