"""Multi-GPU reduction fused into the sampler's flush: shards accumulate straight into ONE pair of
tables, which another process maps through CUDA IPC (over NVLink when it sits on another GPU).
Runs on a single GPU too: the second shard is a child process mapping the parent's tables."""
import os
import subprocess
import sys

import numpy as np
import pytest

from process_b200 import _abi as A
from process_b200 import _lib as L
from process_b200.synth import synth_forest

from conftest import make_params, small_spec

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from process_b200 import _lib as L
from process_b200.synth import synth_forest
from conftest import make_params, small_spec
handle = bytes.fromhex(sys.argv[1]); n_depth = int(sys.argv[2]); device = int(sys.argv[3])
f = synth_forest(small_spec(1))
ctx = L.Context(device)
dev = L.Forest(ctx, f)
plan = L.Plan(dev, make_params(coverage=25.0, purity=0.7, sequencer=2, error_rate=0.02, insert_size_mean=180,
                               shard_rank=1, shard_count=2))
base = ctx.shared_open(handle)
st = plan.accumulate(base, base + 4 * n_depth)
ctx.shared_close(base)
print("child_reads", st.n_reads)
"""


def test_two_processes_accumulate_into_one_table(tmp_path):
    import torch
    f = synth_forest(small_spec(1))
    ctx = L.Context(0)
    dev = L.Forest(ctx, f)
    kw = dict(coverage=25.0, purity=0.7, sequencer=A.PCS_SEQ_BASIC_RANDOM, error_rate=0.02, insert_size_mean=180)
    want_occ, want_cov, want_st = dev.simulate(make_params(**kw))
    plan = L.Plan(dev, make_params(shard_rank=0, shard_count=2, **kw))
    S, M, Lc = plan.info.n_out_samples, plan.info.n_mut, plan.info.n_loci
    base, handle = ctx.shared_alloc(S * Lc + 2 * S * M)
    depth_p, occ_p, cov_p = base, base + 4 * S * Lc, base + 4 * (S * Lc + S * M)
    ctx.memset_u32(base, S * Lc + 2 * S * M)
    mine = plan.accumulate(depth_p, occ_p)  # also synchronises the memset
    second = 1 if torch.cuda.device_count() > 1 else 0
    script = tmp_path / "child.py"
    script.write_text(CHILD.format(root=ROOT))
    out = subprocess.run([sys.executable, str(script), handle.hex(), str(S * Lc), str(second)], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    child_reads = int(out.stdout.strip().split()[-1])
    fin = plan.finalize(depth_p, occ_p, cov_p)
    occ = ctx.to_host(occ_p, S * M).reshape(S, M)
    cov = ctx.to_host(cov_p, S * M).reshape(S, M)
    assert mine.n_reads + child_reads == want_st.n_reads
    assert np.array_equal(occ, want_occ) and np.array_equal(cov, want_cov)
    assert fin.sum_occurrences == want_st.sum_occurrences and fin.sum_depth == want_st.sum_depth
    ctx.shared_free(base)
    plan.close()
    dev.close()
    ctx.close()


@pytest.mark.parametrize("n_ctx", [2, 3])
@pytest.mark.parametrize("links", ["1", "force"])
def test_one_process_several_contexts(n_ctx, links, monkeypatch):
    """pcs_simulate_multi: one host process, one context per device (or several on the one GPU there is),
    shard i on context i, every sampler flushing into the first context's tables; the tables go home over the first
    device's link (PCS_MULTI_LINKS=1) or in row slices over every device's (forced here: the tables are small)."""
    import torch
    monkeypatch.setenv("PCS_MULTI_LINKS", links)
    f = synth_forest(small_spec(2))
    n_dev = torch.cuda.device_count()
    ctxs = [L.Context(i % n_dev) for i in range(n_ctx)]
    first = L.Forest(ctxs[0], f)
    forests = [first] + [L.replicate(first, c) for c in ctxs[1:]]
    P = make_params(coverage=30.0, purity=0.6, sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=0.01)
    want_occ, want_cov, want = first.simulate(P)
    occ, cov, st = L.simulate_multi(forests, P)
    assert np.array_equal(occ, want_occ) and np.array_equal(cov, want_cov)
    assert st.n_reads == want.n_reads and st.sum_occurrences == want.sum_occurrences
    # groups set on the first forest are what the replicas see (shared host view)
    groups = (np.arange(f.n_leaves) % 2).astype(np.uint32)
    first.set_groups(groups, 2)
    want_occ, want_cov, _ = first.simulate(P)
    occ, cov, _ = L.simulate_multi(forests, P)
    assert occ.shape[0] == 3 and np.array_equal(occ, want_occ) and np.array_equal(cov, want_cov)
    for fo in forests:
        fo.close()
    for c in ctxs:
        c.close()


def test_coverage_gather_on_a_side_stream_equals_the_plain_finalize():
    """pcs_plan_finalize_stream: the owner of shared tables gathers the coverage on a stream of its own, beside the
    next step's sampler (bench.py at N > 1).  Same coverage table as pcs_plan_finalize on the context's stream."""
    import torch
    f = synth_forest(small_spec(3))
    stream, side = torch.cuda.Stream(), torch.cuda.Stream()
    ctx = L.Context(0, stream.cuda_stream)
    dev = L.Forest(ctx, f)
    plan = L.Plan(dev, make_params(coverage=20.0, purity=0.8, sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=0.01))
    want_occ, want_cov, want = plan.run()
    S, M, Lc = plan.info.n_out_samples, plan.info.n_mut, plan.info.n_loci
    base, _ = ctx.shared_alloc(S * Lc + 2 * S * M)
    depth_p, occ_p = base, base + 4 * S * Lc
    cov_a = torch.zeros((S, M), dtype=torch.int32, device="cuda")
    cov_b = torch.full((S, M), -1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    plan.counters()  # read and reset what plan.run() counted
    ctx.memset_u32(base, S * Lc + 2 * S * M)
    plan.accumulate(depth_p, occ_p, wait=False)
    plan.finalize(depth_p, occ_p, cov_a.data_ptr(), wait=False)   # on the context's stream
    side.wait_stream(stream)
    plan.finalize_on(depth_p, cov_b.data_ptr(), side.cuda_stream)  # on the side stream, behind the sampler
    torch.cuda.synchronize()
    got = plan.counters()
    assert got.n_reads == want.n_reads
    assert np.array_equal(cov_a.cpu().numpy().astype(np.uint32), want_cov)
    assert np.array_equal(cov_b.cpu().numpy().astype(np.uint32), want_cov)
    assert np.array_equal(ctx.to_host(occ_p, S * M).reshape(S, M), want_occ)
    ctx.shared_free(base)
    plan.close()
    dev.close()
    ctx.close()
