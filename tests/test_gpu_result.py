"""Result assembly on the device (/root/reference/src/seq_simulation.cpp:92-181) and the
sample-by-sample pipelined call behind pcs_simulate()/pcs_simulate_result(): both must give, bit for bit, what the
one-piece plan + full tables + host compaction give."""
import numpy as np
import pytest

import oracle
from process_b200 import _abi as A
from process_b200 import _lib as L
from process_b200.synth import synth_forest

from conftest import make_params, small_spec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = L.Context(0)
    yield c
    c.close()


def host_frame(dev, occ, cov, include, P):
    """the reference construction on the host: active rows (pcs_active_rows), gathered columns, VAF"""
    rows = dev.active_rows(occ, include, P)
    o, c = occ[:, rows].astype(np.int32), cov[:, rows].astype(np.int32)
    vaf = np.divide(o, c, out=np.zeros(o.shape, np.float64), where=c != 0)
    return rows, o, c, vaf


CASES = [dict(), dict(sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=0.02, insert_size_mean=220),
         dict(sequencer=A.PCS_SEQ_BASIC_RANDOM, error_rate=0.02, preneoplastic_in_normal=1),
         dict(normal_only=1, with_normal_sample=0), dict(with_normal_sample=0, purity=1.0), dict(chr_mask=[0, 1, 1])]


@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("include", [False, True])
def test_pipelined_call_and_device_result_equal_the_one_piece_plan(ctx, kw, include):
    f = synth_forest(small_spec(4))
    dev = L.Forest(ctx, f)
    P = make_params(**{**dict(coverage=9.0, purity=0.7, seed=21), **kw})  # low coverage: many rows never sequenced
    plan = L.Plan(dev, P)
    occ, cov, st = plan.run()                      # one-piece plan, full tables
    occ2, cov2, st2 = dev.simulate(P)              # planned and launched sample by sample
    assert np.array_equal(occ, occ2) and np.array_equal(cov, cov2)
    assert st2.n_reads == st.n_reads and st2.sum_occurrences == st.sum_occurrences and st2.sum_depth == st.sum_depth
    want = host_frame(dev, occ, cov, include, P)
    for res in (plan.result(include), dev.simulate_result(P, include)[0]):
        got = res.fetch()
        assert res.n_samples == occ.shape[0] and res.n_rows == len(want[0])
        for g, w in zip(got, want):
            assert g.dtype == w.dtype and np.array_equal(g, w)
        assert res.d2h_bytes == len(want[0]) * (4 + occ.shape[0] * 16)
        res.close()
    if not include:
        assert 0 < len(want[0]) < f.n_mut  # the compaction had something to drop
    else:
        assert len(want[0]) >= len(host_frame(dev, occ, cov, False, P)[0])
        masked = np.zeros(f.n_mut, bool) if "chr_mask" not in kw else (f.mut_chr == 0)
        assert not masked[want[0]].any()  # rows of chromosomes that were not sequenced never appear
    # VAF is optional; a result without it refuses to fetch it
    res, _ = dev.simulate_result(P, include, with_vaf=False)
    rows, o, c, v = res.fetch()
    assert v is None and np.array_equal(rows, want[0]) and np.array_equal(o, want[1])
    res.close()
    plan.close()
    dev.close()


def test_last_sample_launched_in_chromosome_groups_gives_the_same_tables():
    """pcs_simulate() launches the last sample's tiles in groups of whole chromosomes so that a group's rows cross
    the link while the next group is sampled (big jobs only; PCS_SPLIT_LAST=force: always).  Same tables, bit for bit.
    The switch is read once per process: a child process runs the forced case."""
    import subprocess, sys, os, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from process_b200 import _abi as A, _lib as L
        from process_b200.synth import synth_forest
        from conftest import make_params, small_spec
        ctx = L.Context(0)
        for seed, kw in ((4, {}), (2, dict(sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=0.02)), (1, dict(chr_mask=[0, 1, 1]))):
            f = synth_forest(small_spec(seed))
            dev = L.Forest(ctx, f)
            P = make_params(**{**dict(coverage=9.0, purity=0.7, seed=21), **kw})
            plan = L.Plan(dev, P)
            occ, cov, st = plan.run()
            occ2, cov2, st2 = dev.simulate(P)
            assert np.array_equal(occ, occ2) and np.array_equal(cov, cov2), "tables differ"
            assert st2.n_reads == st.n_reads and st2.sum_occurrences == st.sum_occurrences and st2.sum_depth == st.sum_depth
            assert st2.kernel_launches > st.kernel_launches, (st2.kernel_launches, st.kernel_launches)
            plan.close(); dev.close()
        print("split ok")
    """) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PCS_SPLIT_LAST="force")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "split ok" in out.stdout, out.stdout + out.stderr


def test_empty_result_and_shards(ctx):
    f = synth_forest(small_spec(1))
    dev = L.Forest(ctx, f)
    res, st = dev.simulate_result(make_params(coverage=0.0))
    rows, o, c, v = res.fetch()
    assert res.n_rows == 0 and len(rows) == 0 and o.shape == (4, 0) and st.n_reads == 0
    res.close()
    # the pipelined call shards inside every sample: the shards still add up to the unsharded tables
    P = make_params(coverage=20.0, purity=0.8, sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=0.01)
    occ, cov, st = dev.simulate(P)
    acc_o, acc_c, reads = np.zeros_like(occ), np.zeros_like(cov), 0
    for r in range(3):
        o, c, s = dev.simulate(make_params(coverage=20.0, purity=0.8, sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=0.01,
                                           shard_rank=r, shard_count=3))
        acc_o += o
        acc_c += c
        reads += s.n_reads
    assert reads == st.n_reads and np.array_equal(acc_o, occ) and np.array_equal(acc_c, cov)
    dev.close()


def test_api_frame_equals_the_host_construction(ctx):
    """api.simulate_seq builds its frame from the device result and the native column builders: same frame as the
    full tables + pcs_active_rows + per-row strings"""
    import pandas as pd
    from process_b200 import api
    f = synth_forest(small_spec(2))
    f.reference_path = __file__  # any existing file: no SAM is written
    for include in (False, True):
        r = api.simulate_seq(f, coverage=6.0, purity=0.9, seed=5, chromosomes=["1", "X"],
                             include_non_sequenced_mutations=include)
        df = r["mutations"]
        dev = L.Forest(ctx, f)
        P = make_params(coverage=6.0, purity=0.9, seed=5, chr_mask=[1, 0, 1])
        occ, cov, _ = dev.simulate(P)
        names = list(f.sample_names) + ["normal_sample"]
        want = api._frame_from_tables(f, dev, occ, cov, names, include, P)
        dev.close()
        assert list(df.columns) == list(want.columns) and (df.dtypes == want.dtypes).all() and df.equals(want)
        ref, alt = f.row_strings(dev_rows := np.flatnonzero(np.isin(f.mut_pos, df["chr_pos"].to_numpy())))
        assert set(df["ref"]) <= set(ref) and (df[[c for c in df.columns if c.endswith(".VAF")]].to_numpy() <= 1).all()
    api.release_device_cache()


@pytest.mark.parametrize("insert", [0, 250])
@pytest.mark.parametrize("bin_bp", [256, 4096])
def test_coverage_track_describes_the_reads_of_the_plan(ctx, insert, bin_bp):
    """pcs_plan_coverage_track: the binned depth track is drawn with the Philox counters of the
    counting kernels, so it must be, bin for bin, the track of the very reads pcs_plan_trace lists"""
    f = synth_forest(small_spec(6))
    dev = L.Forest(ctx, f)
    P = make_params(coverage=40.0, purity=0.8, read_size=150, insert_size_mean=insert, seed=9)
    plan = L.Plan(dev, P)
    occ, cov, st = plan.run()
    off, track = plan.coverage_track(bin_bp)
    rec, _ = plan.trace(cap=int(st.n_reads) + 8)
    want = np.zeros(track.shape, np.int64)
    R = P.read_size
    for k in range(R):  # one base at a time: trivially right
        pos = rec["start"].astype(np.int64) + k
        np.add.at(want, (rec["sample"].astype(np.int64), off[rec["chr"]].astype(np.int64) + pos // bin_bp), 1)
    assert track.sum() == st.n_reads * R and np.array_equal(track.astype(np.int64), want)
    # mean depth of the track = the requested coverage, for every sample
    G = float(f.chr_len.sum())
    assert np.allclose(track.sum(axis=1) / G, 40.0, rtol=2e-2)
    plan.close()
    dev.close()
