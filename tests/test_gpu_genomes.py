"""The forest uploaded as explicit per-cell genomes (pcs_forest_upload_genomes: what the reference's seam hands over,
/root/reference/src/seq_simulation.cpp:566-575) must count and sample like the same forest uploaded as an
event-labelled tree: identical tables for identical reads, the same sampling law."""
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


@pytest.mark.parametrize("seed", [0, 3])
def test_injected_reads_and_own_reads_count_bit_exact(ctx, seed):
    f = synth_forest(small_spec(seed))
    g = oracle.cell_genomes(f)
    P = make_params(coverage=12.0, purity=0.7, sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=0.02,
                    preneoplastic_in_normal=1, seed=4)
    ref = oracle.simulate(f, P, trace_cap=600_000, trace_masks=True)
    dev = L.Forest(ctx, g)
    occ, cov, st = dev.count_injected(dev.n_out_samples(P), P.read_size, ref["trace"], ref["masks"])
    assert np.array_equal(occ, ref["occ"]) and np.array_equal(cov, ref["cov"])
    # the sampler on the genome-built view: the oracle's recount of the reads it placed
    plan = L.Plan(dev, P)
    occ, cov, st = plan.run()
    rec, masks = plan.trace(cap=int(st.n_reads) + 16, with_masks=True)
    occ2, cov2 = oracle.count_injected(f, plan.info.n_out_samples, P.read_size, rec, masks)
    assert st.n_reads > 10_000 and np.array_equal(occ, occ2) and np.array_equal(cov, cov2)
    # same job as the event-labelled upload: same number of reads (same tile grid and template counts)
    dev_e = L.Forest(ctx, f)
    _, _, st_e = dev_e.simulate(P)
    assert st_e.n_templates == st.n_templates
    dev_e.close()
    plan.close()
    dev.close()


def test_sampling_law_on_the_genome_built_view(ctx):
    import closed_form as CF
    f = synth_forest(CF.snv_only_spec())
    g = oracle.cell_genomes(f)
    coverage, R, purity = 3000.0, 100, 0.8
    e_cov, e_occ = CF.expected_tables(f, coverage, purity, R)
    dev = L.Forest(ctx, g)
    occ, cov, st = dev.simulate(make_params(coverage=coverage, purity=purity, read_size=R, seed=13))
    dev.close()
    for obs, exp in ((cov, e_cov), (occ, e_occ)):
        z, impossible = CF.z_scores(obs, exp)
        assert impossible == 0
        assert len(z) > 1500 and abs(z.mean()) < 0.1 and 0.9 < z.std() < 1.1 and np.abs(z).max() < 5.5


@pytest.mark.parametrize("seed", [0, 3, 7])
@pytest.mark.parametrize("as_genomes", [False, True])
def test_device_built_instances_equal_the_host_table(ctx, seed, as_genomes, monkeypatch):
    """An uploaded forest builds its instance table on the device from one germline mask byte per row
    (kernels.cu: build_instances_kernel); the host flattener's table (pcs_flat_create; also PCS_DEVICE_INSTANCES=0)
    is the same table, bit for bit -- rows in order, inside a row the somatic placements then the germline one."""
    f = synth_forest(small_spec(seed))
    src = oracle.cell_genomes(f) if as_genomes else f
    inst_h, off_h = L.Flat(src).instances()
    assert len(inst_h) > 1000 and (inst_h[:, 3] != 0x0101).any()  # there are indels among them
    monkeypatch.setenv("PCS_DEVICE_INSTANCES", "1")
    dev = L.Forest(ctx, src)
    inst_d, off_d = dev.instances()
    assert np.array_equal(off_d, off_h) and np.array_equal(inst_d, inst_h)
    # the rows carried by sequenced cells are computed on the host from a copy fetched on first use
    P = make_params(coverage=5.0, seed=1)
    res = dev.simulate_result(P, include_non_sequenced=1)[0]
    r1 = res.fetch()[0].copy()
    res.close()
    assert len(r1) > 1000
    dev.close()
    monkeypatch.setenv("PCS_DEVICE_INSTANCES", "0")
    dev0 = L.Forest(ctx, src)
    inst_0, off_0 = dev0.instances()
    assert np.array_equal(off_0, off_h) and np.array_equal(inst_0, inst_h)
    res0 = dev0.simulate_result(P, include_non_sequenced=1)[0]
    assert np.array_equal(r1, res0.fetch()[0])
    res0.close()
    dev0.close()
