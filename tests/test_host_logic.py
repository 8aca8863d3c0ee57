"""Host side of libpcs_seq without a GPU: the flattened (haplotype-interval) view
against explicit per-cell genomes, and the planner (tile grid, multinomial, shards)."""
import numpy as np
import pytest

import oracle
from process_b200 import _abi as A
from process_b200 import _lib as L
from process_b200.synth import synth_forest

from conftest import make_params, small_spec
from golden import micro_forest as MF


def check_flat_against_explicit_genomes(f, flat_of=None):
    """flat_of: what the flat view is built from (default: the event-labelled forest f itself)"""
    fl = L.Flat(f if flat_of is None else flat_of)
    germ = {}
    for m, mask in zip(f.germ_mut, f.germ_allele_mask):
        germ.setdefault(int(f.mut_chr[m]), []).append((int(m), int(mask)))
    n_roots = int((f.node_parent < 0).sum())
    cells = [(A.PCS_PLACE_TUMOUR, l) for l in range(f.n_leaves)] + [(A.PCS_PLACE_NORMAL_PLAIN, 0)] + \
            [(A.PCS_PLACE_NORMAL_PRENEO, r) for r in range(n_roots)]
    checked = 0
    for c in range(f.n_chr):
        for kind, cell in cells:
            frags, sids = oracle.cell_genome(f, kind, cell, c)
            haps = {a: (h, fs) for a, h, fs in fl.cell_haps(kind, cell, c)}
            alleles = {}
            for a, o, b, e in frags:
                d = alleles.setdefault(a, dict(origin=o, frags=[], rows=set()))
                if b > 0:
                    d["frags"].append((b, e))
            for a, r in sids:
                alleles[a]["rows"].add(r)
            assert set(haps) <= set(alleles)
            for a, d in alleles.items():
                if not d["frags"]:
                    assert a not in haps  # an allele with no DNA left cannot be read
                    continue
                h, fs = haps[a]
                assert fl.fragset(fs) == sorted(d["frags"])
                inside = lambda m: any(b <= int(f.mut_pos[m]) <= e for b, e in d["frags"])  # noqa: E731
                want = set(d["rows"]) | {m for m, mask in germ.get(c, []) if mask & (1 << d["origin"]) and inside(m)}
                got = {int(m) for m in fl.hap_rows(c, h) if inside(m)}
                assert got == want, (kind, cell, c, a)
                checked += 1
    return checked


def test_flat_view_of_the_micro_forest():
    f = MF.forest()
    assert check_flat_against_explicit_genomes(f) == 2 + 3 + 2 + 2
    info = L.Flat(f).info()
    assert info["n_loci"] == 8 and info["n_haplotypes"] == 2 + 3 + 2 + 2


@pytest.mark.parametrize("seed", range(6))
def test_flat_view_matches_explicit_genomes(seed):
    f = synth_forest(small_spec(seed))
    assert (f.ev_kind == A.PCS_EV_WGD).sum() >= 1 and (f.ev_kind == A.PCS_EV_CNA_AMP).sum() >= 3
    assert check_flat_against_explicit_genomes(f) > 100


def test_flat_view_with_two_roots_and_nested_copies():
    """amplify, then WGD the copy, then delete inside the copy of the copy"""
    F = dict(MF.FOREST)
    F = {k: list(v) if isinstance(v, list) else v for k, v in F.items()}
    # second root (node 3) with one sampled child (node 4)
    F["node_parent"] = [-1, 0, 0, -1, 3]
    F["leaf_node"] = [1, 2, 4]
    F["leaf_sample"] = [0, 1, 1]
    extra = [  # kind chr pos len allele dest mut nature
        (0, 0, 0, 0, 0, 0, 4, 3),      # root 3: pre-neoplastic SID row 4 (pos 400) on allele 0
        (1, 0, 350, 300, 0, 2, 0, 0),  # node 4: AMP [350,649] a0 -> a2
        (3, 0, 0, 0, 0, 0, 0, 0),      #         WGD: a0->a3, a1->a4, a2->a5
        (2, 0, 380, 40, 5, 0, 0, 0),   #         DEL [380,419] of a5
        (0, 0, 0, 0, 5, 0, 5, 1),      #         SID row 5 (pos 600) on a5
    ]
    for i, k in enumerate(["ev_kind", "ev_chr", "ev_pos", "ev_len", "ev_allele", "ev_dest", "ev_mut", "ev_nature"]):
        F[k] = F[k] + [e[i] for e in extra]
    F["node_event_off"] = [0, 1, 2, 6, 7, 11]
    from process_b200.forest import PhylogeneticForest
    f = PhylogeneticForest(**{k: (v if k in ("chr_names", "sample_names") else np.asarray(v)) for k, v in F.items()}).normalise()
    frags, sids = oracle.cell_genome(f, A.PCS_PLACE_TUMOUR, 2, 0)
    assert (5, 0, 350, 379) in frags and (5, 0, 420, 649) in frags and (2, 0, 350, 649) in frags
    # row 4 (pos 400) is copied to a2, a3 and a5, then deleted from a5 together with [380,419]
    assert sorted(sids) == [(0, 4), (2, 4), (3, 4), (5, 5)]
    assert check_flat_against_explicit_genomes(f) >= 10


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_flat_view_from_explicit_per_cell_genomes(seed):
    """pcs_forest_upload_genomes / pcs_flat_create_genomes: the forest handed over the way the reference's seam does
    it -- per-cell genomes, chromosome -> allele -> fragment -> SID, no tree
    (/root/reference/src/seq_simulation.cpp:566-575, src/phylogenetic_forest.cpp:279-290) -- must flatten to a view
    in which every cell's every allele has its fragments and carries exactly its SIDs"""
    f = synth_forest(small_spec(seed))
    g = oracle.cell_genomes(f)
    assert check_flat_against_explicit_genomes(f, flat_of=g) > 50
    fl_e, fl_g = L.Flat(f), L.Flat(g)
    ie, ig = fl_e.info(), fl_g.info()
    # same loci, same haplotypes; placements: one per run of carriers -- the tree-less ordering keeps clades together,
    # so there are about as many as events placed them (a SID a later deletion took from part of its clade splits)
    assert ig["n_loci"] == ie["n_loci"] and ig["n_haplotypes"] == ie["n_haplotypes"]
    assert ig["n_instances"] <= 1.2 * ie["n_instances"] + 16
    # the planner sees the same job: same tile grid, same templates per tile (the RNG streams are keyed by
    # (seed, sample, chromosome), the weights by fragment lengths and head counts)
    P = make_params(coverage=20.0, purity=0.7)
    info_e, t_e = fl_e.plan(P)
    info_g, t_g = fl_g.plan(P)
    assert info_e.n_templates_total == info_g.n_templates_total and info_e.n_tiles_total == info_g.n_tiles_total
    oe, og = np.argsort(t_e["id"]), np.argsort(t_g["id"])
    for k in ("id", "templates", "sample", "chr", "begin", "len"):
        assert np.array_equal(t_e[k][oe], t_g[k][og]), k


def test_explicit_genomes_are_validated():
    f = MF.forest()
    g = oracle.cell_genomes(f)
    check_flat_against_explicit_genomes(f, flat_of=g)
    bad = oracle.cell_genomes(f)
    bad.sid_row = bad.sid_row.copy()
    bad.sid_row[0] = 0  # row 0 is a germline SID of allele 0 ... on whatever allele: outside or duplicate
    with pytest.raises(L.PcsError):
        L.Flat(bad)
    bad = oracle.cell_genomes(f)
    bad.frag_end = bad.frag_end.copy()
    bad.frag_end[0] = 5000  # past the chromosome
    with pytest.raises(L.PcsError):
        L.Flat(bad)
    # without the normal cells that carry the pre-neoplastic SIDs, preneoplastic_in_normal is refused, nothing else
    g0 = oracle.cell_genomes(f, with_preneo=False)
    fl = L.Flat(g0)
    fl.plan(make_params(coverage=5.0))
    with pytest.raises(L.PcsError, match="pre-neoplastic"):
        fl.plan(make_params(coverage=5.0, preneoplastic_in_normal=1))


def _all_hap_rows(f):
    fl = L.Flat(f)
    out = []
    n_roots = int((f.node_parent < 0).sum())
    cells = [(A.PCS_PLACE_TUMOUR, l) for l in range(f.n_leaves)] + [(A.PCS_PLACE_NORMAL_PLAIN, 0)] + \
            [(A.PCS_PLACE_NORMAL_PRENEO, r) for r in range(n_roots)]
    for c in range(f.n_chr):
        for kind, cell in cells:
            for a, h, fs in fl.cell_haps(kind, cell, c):
                out.append((c, kind, cell, a, h, fs, fl.hap_rows(c, h).tolist()))
    return fl.info(), out


def test_germline_list_order_does_not_matter():
    """the germline list is taken in any order: sorted by row (scattered where it lies), shuffled (partitioned by
    row range first), and with a row listed twice (sorted copy): same view"""
    f = synth_forest(small_spec(3))
    rng = np.random.default_rng(5)
    order = np.argsort(f.germ_mut, kind="stable")
    base_mut, base_mask = f.germ_mut.copy(), f.germ_allele_mask.copy()
    f.germ_mut, f.germ_allele_mask = base_mut[order].copy(), base_mask[order].copy()
    info_sorted, rows_sorted = _all_hap_rows(f)
    p = rng.permutation(len(base_mut))
    f.germ_mut, f.germ_allele_mask = base_mut[p].copy(), base_mask[p].copy()
    info_shuffled, rows_shuffled = _all_hap_rows(f)
    assert info_sorted == info_shuffled and rows_sorted == rows_shuffled
    assert check_flat_against_explicit_genomes(f) > 100
    # a heterozygous SID listed once per allele instead of once with both bits: two instances, same carriers
    het = np.flatnonzero(base_mask == 3)[:5]
    assert len(het) == 5
    mut = np.concatenate([base_mut, base_mut[het]])
    mask = np.concatenate([base_mask, np.full(5, 2, np.uint8)])
    mask[het] = 1
    p = rng.permutation(len(mut))
    f.germ_mut, f.germ_allele_mask = mut[p].astype(np.uint32), mask[p].astype(np.uint8)
    info_dup, rows_dup = _all_hap_rows(f)
    assert info_dup["n_instances"] == info_sorted["n_instances"] + 5
    assert [(r[:6], sorted(set(r[6]))) for r in rows_dup] == [(r[:6], sorted(set(r[6]))) for r in rows_sorted]


@pytest.mark.parametrize("regroup", [False, True])
def test_sampling_lists_partition_the_haplotypes(regroup):
    """every haplotype is in exactly the list of (its output group | normal kind, its fragment set); lists are in
    increasing haplotype order and tile the device's hap_list in (group, fragment set) order, no gaps"""
    f = synth_forest(small_spec(4))
    fl = L.Flat(f)
    n_groups = f.n_samples
    leaf_group = np.asarray(f.leaf_sample, np.uint32)
    if regroup:
        n_groups = 4
        leaf_group = (np.arange(f.n_leaves) % 4).astype(np.uint32)
        fl.set_groups(leaf_group, n_groups)
    info = fl.info()
    n_roots = int((f.node_parent < 0).sum())
    want = {}
    cells = [(A.PCS_PLACE_TUMOUR, l) for l in range(f.n_leaves)] + [(A.PCS_PLACE_NORMAL_PLAIN, 0)] + \
            [(A.PCS_PLACE_NORMAL_PRENEO, r) for r in range(n_roots)]
    for c in range(f.n_chr):
        for kind, cell in cells:
            g = int(leaf_group[cell]) if kind == A.PCS_PLACE_TUMOUR else n_groups + (kind == A.PCS_PLACE_NORMAL_PRENEO)
            for a, h, fs in fl.cell_haps(kind, cell, c):
                want.setdefault((g, fs), []).append(h)
    assert sum(len(v) for v in want.values()) == info["n_haplotypes"]
    at = 0
    for g in range(n_groups + 2):
        for fs in range(info["n_fragment_sets"]):
            off, haps = fl.group_list(g, fs)
            assert haps.tolist() == sorted(want.get((g, fs), []))
            if len(haps):
                assert off == at
                at += len(haps)
    assert at == info["n_haplotypes"]


def test_malformed_forests_are_refused():
    f = MF.forest()
    f.mut_pos = f.mut_pos[::-1].copy()
    with pytest.raises(L.PcsError):
        L.Flat(f)
    f = MF.forest()
    f.node_parent = np.asarray([1, -1, 0], np.int32)
    with pytest.raises(L.PcsError):
        L.Flat(f)
    f = MF.forest()
    f.germ_allele_mask = np.asarray([1, 3, 4, 1], np.uint8)
    with pytest.raises(L.PcsError):
        L.Flat(f)


def _micro_with_events(extra, node):
    """the micro forest with `extra` events (kind, pos, len, allele, dest, mut, nature) appended to `node`"""
    d = {k: (list(v) if isinstance(v, list) else v) for k, v in MF.FOREST.items()}
    at = d["node_event_off"][node + 1]
    for j, (kind, pos, ln, allele, dest, mut, nature) in enumerate(extra):
        for key, val in (("ev_kind", kind), ("ev_chr", 0), ("ev_pos", pos), ("ev_len", ln), ("ev_allele", allele),
                         ("ev_dest", dest), ("ev_mut", mut), ("ev_nature", nature)):
            d[key].insert(at + j, val)
    d["node_event_off"] = [o + (len(extra) if i > node else 0) for i, o in enumerate(d["node_event_off"])]
    from process_b200.forest import PhylogeneticForest
    kw = {k: (v if k in ("chr_names", "sample_names") else np.asarray(v)) for k, v in d.items()}
    return PhylogeneticForest(**kw).normalise()


def test_a_haplotype_carries_at_most_one_sid_per_position():
    """the kernels' walk counts one SID per position of a haplotype; the flattener must refuse anything else,
    exactly where the oracle does (its explicit genomes are position-keyed maps)"""
    SID, AMP = A.PCS_EV_SID, A.PCS_EV_CNA_AMP
    refused = {
        "the root's SID again in a child, same allele": ([(SID, 0, 0, 1, 0, 2, 1)], 1),
        "the same SID twice in one node": ([(SID, 0, 0, 0, 0, 4, 1)], 1),
        "a somatic SID on a germline SID of the allele": ([(SID, 0, 0, 0, 0, 0, 1)], 1),
        "inherited through an amplified copy": ([(AMP, 350, 100, 0, 3, 0, 1), (SID, 0, 0, 3, 0, 4, 1)], 1),
        "amplification into an allele id the lineage has (no WGD anywhere)": ([(AMP, 350, 100, 0, 1, 0, 1)], 1),
    }
    for why, (extra, node) in refused.items():
        f = _micro_with_events(extra, node)
        with pytest.raises(L.PcsError):
            L.Flat(f)
        with pytest.raises(oracle.OracleError):
            oracle.cell_genome(f, A.PCS_PLACE_TUMOUR, 0, 0)
    accepted = {
        "the same position on the other allele": [(SID, 0, 0, 1, 0, 4, 1)],
        "a germline position of the OTHER allele": [(SID, 0, 0, 1, 0, 0, 1)],
        "a position the allele has lost": [(A.PCS_EV_CNA_DEL, 380, 40, 0, 0, 0, 1), (SID, 0, 0, 0, 0, 4, 1)],
    }
    for why, extra in accepted.items():
        f = _micro_with_events(extra, 2 if "lost" not in why else 1)
        check_flat_against_explicit_genomes(f)
    # a germline SID listed twice for one allele
    f = MF.forest()
    f.germ_mut = np.asarray([0, 1, 3, 7, 0], np.uint32)
    f.germ_allele_mask = np.asarray([1, 3, 2, 1, 1], np.uint8)
    with pytest.raises(L.PcsError):
        L.Flat(f)


def test_planner_tile_grid_and_template_counts():
    f = synth_forest(small_spec(1))
    fl = L.Flat(f)
    P = make_params(coverage=40.0, purity=0.8)
    info, t = fl.plan(P)
    assert info.n_out_samples == f.n_samples + 1 and info.n_tiles == len(t["id"])
    # A9: round(coverage * chr_len / read_size) templates per (sample, chromosome)
    for s in range(info.n_out_samples):
        for c in range(f.n_chr):
            got = int(t["templates"][(t["sample"] == s) & (t["chr"] == c)].sum())
            assert got == int(round(40.0 * int(f.chr_len[c]) / 150))
    # heaviest first, tiles inside their chromosome, distinct ids
    assert np.all(np.diff(t["templates"].astype(np.int64)) <= 0)
    assert np.all(t["begin"] + t["len"] - 1 <= f.chr_len[t["chr"]])
    assert len(np.unique(t["id"])) == len(t["id"])
    # same seed -> same plan; other seed -> other counts
    _, t2 = fl.plan(make_params(coverage=40.0, purity=0.8))
    assert all(np.array_equal(t[k], t2[k]) for k in t)
    _, t3 = fl.plan(make_params(coverage=40.0, purity=0.8, seed=8))
    assert not np.array_equal(t["templates"], t3["templates"])
    # paired reads: half as many templates
    info_p, _ = fl.plan(make_params(coverage=40.0, insert_size_mean=300))
    assert info_p.reads_per_template == 2
    assert abs(info_p.n_templates_total * 2 - info.n_templates_total) <= info.n_out_samples * f.n_chr


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_planner_shards_partition_the_tile_grid(shards):
    f = synth_forest(small_spec(2))
    fl = L.Flat(f)
    info, t = fl.plan(make_params(coverage=25.0, purity=0.7))
    seen = {}
    loads = []
    for r in range(shards):
        info_r, tr = fl.plan(make_params(coverage=25.0, purity=0.7, shard_rank=r, shard_count=shards))
        assert info_r.n_templates_total == info.n_templates_total and info_r.n_tiles_total == info.n_tiles_total
        for i, n in zip(tr["id"], tr["templates"]):
            assert i not in seen
            seen[int(i)] = int(n)
        loads.append(int(tr["templates"].sum()))
        assert loads[-1] == info_r.n_templates
    assert seen == {int(i): int(n) for i, n in zip(t["id"], t["templates"])}
    # LPT balances the tiles' COST (templates x a locus-density term), so template loads agree only roughly
    assert max(loads) / min(loads) < 1.15
    with pytest.raises(L.PcsError):
        fl.plan(make_params(shard_rank=shards, shard_count=shards))


def test_parameter_validation_messages():
    fl = L.Flat(MF.forest())
    with pytest.raises(L.PcsError, match="must be greater than or equal to its variance"):
        fl.plan(make_params(insert_size_mean=50, insert_size_stddev=10))
    with pytest.raises(L.PcsError, match="purity"):
        fl.plan(make_params(purity=1.5))
    with pytest.raises(L.PcsError, match="Unsupported sequencer type"):
        fl.plan(make_params(sequencer=9))
    info, _ = fl.plan(make_params(normal_only=1, purity=7.0))  # purity is ignored for the normal sample
    assert info.n_out_samples == 1
    # the insert law is tabulated on its support only: sd^2 close to the mean (billions of Binomial trials) is as
    # quick as sd = 10, and sd^2 == mean (p = 0: the insert is always 0) is a table of one column
    import time
    t0 = time.perf_counter()
    for mean, sd in ((100_000, 316), (500_000_000, 20_000), (289, 17), (300, 10)):
        info, _ = fl.plan(make_params(insert_size_mean=mean, insert_size_stddev=sd))
        assert info.reads_per_template == 2
    assert time.perf_counter() - t0 < 5.0
    # 32-bit positions on the device: a template that could wrap them is refused, not sampled
    with pytest.raises(L.PcsError, match="shorter than 2\\^30 bases"):
        fl.plan(make_params(insert_size_mean=4_000_000_000, insert_size_stddev=0))
    f = MF.forest()
    f.chr_len = f.chr_len.copy()
    f.chr_len[0] = 2**31
    with pytest.raises(L.PcsError, match="below 2\\^31"):
        L.Flat(f)


_THREAD_PROBE = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from conftest import make_params, small_spec
from process_b200 import _abi as A, _lib as L
from process_b200.synth import synth_forest
# > 65 536 rows: the row passes of the flattener run in several chunks
f = synth_forest(small_spec(2, chr_len=[30_000_000, 20_000_000, 12_000_000], germline_density=3e-3, sample_cells=[10, 12, 9],
                            node_snv_mean=40, cna_len=(100000, 5000000)))
fl = L.Flat(f)
h = hashlib.sha256(repr(sorted(fl.info().items())).encode())
rng = np.random.default_rng(0)
for c in range(f.n_chr):
    for l in rng.choice(f.n_leaves, 4, replace=False):
        for a, hp, fs in fl.cell_haps(A.PCS_PLACE_TUMOUR, int(l), c):
            h.update(np.asarray(fl.hap_rows(c, hp)).tobytes())
            h.update(repr((a, hp, fs, fl.fragset(fs))).encode())
info, tiles = fl.plan(make_params(coverage=20.0), cap=1 << 22)
for k in sorted(tiles):
    h.update(tiles[k].tobytes())
print(f.n_mut, info.n_tiles_total, info.n_templates_total, h.hexdigest())
"""


def test_flat_view_and_plan_do_not_depend_on_the_number_of_host_threads():
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for nt in ("1", "3", "8"):
        env = dict(os.environ, PCS_HOST_THREADS=nt)
        r = subprocess.run([sys.executable, "-c", _THREAD_PROBE, root], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.strip())
    assert int(outs[0].split()[0]) > 130_000 and outs[0] == outs[1] == outs[2], outs


def test_thinned_tiles_count_the_useful_offsets_exactly():
    """Single-end reads, staged tiles: the sampler draws only the templates whose read can span a locus (or run past
    its fragment).  The planner's count of such start offsets (Tile.u_len: the device recomputes it and traps on a
    mismatch) against a brute-force union of the loci's windows; n_useful ~ Binomial(templates, u_len / len)."""
    f = synth_forest(small_spec(3))
    fl = L.Flat(f)
    for R in (150, 37, 1):
        P = make_params(coverage=40.0, purity=0.8, read_size=R)
        info, t = fl.plan(P)
        th = fl.plan_thinning(P)
        by_id = {int(i): k for k, i in enumerate(t["id"])}
        assert set(by_id) == {int(i) for i in th["id"]} and th["thin"].sum() > 0.9 * len(th["id"])
        z = []
        for k in range(len(th["id"])):
            q = by_id[int(th["id"][k])]
            begin, ln, c, n_t = int(t["begin"][q]), int(t["len"][q]), int(t["chr"][q]), int(t["templates"][q])
            if not th["thin"][k]:
                assert th["n_useful"][k] == n_t
                continue
            pos = f.mut_pos[f.mut_chr == c].astype(np.int64)
            useful = np.zeros(ln, bool)
            for p in np.unique(pos[(pos >= begin) & (pos <= begin + ln + R - 2)]):
                lo, hi = max(p - R + 1, begin), min(p, begin + ln - 1)
                useful[lo - begin:hi - begin + 1] = True
            useful[int(th["tail_off"][k]):] = True  # the tail zone: a read from there may run past its fragment
            assert int(th["u_len"][k]) == int(useful.sum()), (R, k)
            assert th["n_useful"][k] <= n_t
            if n_t > 200 and 0 < useful.sum() < ln:
                pr = useful.sum() / ln
                z.append((th["n_useful"][k] - n_t * pr) / np.sqrt(n_t * pr * (1 - pr)))
        z = np.asarray(z)
        assert len(z) > 30 and abs(z.mean()) < 0.5 and 0.7 < z.std() < 1.3
    # paired reads and PCS_THIN=0 draw every template
    th = fl.plan_thinning(make_params(coverage=40.0, insert_size_mean=300))
    assert th["thin"].sum() == 0


def test_planner_binomial_sampler_matches_scipy_pmf():
    """csrc/plan_rng.hpp: the planner's Binomial(n, p) draws (inversion below a mean of 10, BTRS above, p > 1/2
    mirrored) against scipy's pmf -- chi-square over the bins with an expectation >= 10, mean and variance."""
    from scipy import stats
    cases = [(5, 0.3), (100, 0.05), (100, 0.5), (1000, 0.2), (139000, 0.2), (139000, 0.8), (8_900_000, 1 / 64),
             (10 ** 9, 0.37), (10 ** 12, 1e-11), (50, 0.999), (40, 0.25), (2000, 0.006)]
    n_draws = 400_000
    pvals = []
    for i, (n, p) in enumerate(cases):
        x = L.host_binomial(1000 + i, n, p, n_draws)
        assert int(x.max()) <= n
        lo, hi = int(x.min()), int(x.max())
        obs = np.bincount((x - lo).astype(np.int64), minlength=hi - lo + 1).astype(float)
        exp = stats.binom.pmf(np.arange(lo, hi + 1), n, p) * n_draws
        keep = exp >= 10
        o = np.append(obs[keep], obs[~keep].sum())
        e = np.append(exp[keep], n_draws - exp[keep].sum())
        if e[-1] < 5:
            o[-2] += o[-1]; e[-2] += e[-1]; o, e = o[:-1], e[:-1]
        pvals.append(stats.chi2.sf(((o - e) ** 2 / e).sum(), len(o) - 1))
        mean, var = n * p, n * p * (1 - p)
        assert abs(x.mean() - mean) < 5 * np.sqrt(var / n_draws), (n, p)
        assert abs(x.var() / var - 1) < 0.02, (n, p)
    # twelve independent p-values: none absurdly small, and not all small
    assert min(pvals) > 1e-4, pvals
    assert np.median(pvals) > 0.1, pvals
    # degenerate arguments
    assert L.host_binomial(1, 0, 0.5, 3).tolist() == [0, 0, 0]
    assert L.host_binomial(1, 17, 0.0, 2).tolist() == [0, 0] and L.host_binomial(1, 17, 1.0, 2).tolist() == [17, 17]


def test_kept_tile_geometry_is_keyed_by_everything_it_depends_on():
    """The forest keeps the tile geometry of its last call (pcs_seq.cpp: GridCache).  A call with another read size,
    insert law, chromosome mask or sample grouping must not see the kept grid; the first call repeated must."""
    f = synth_forest(small_spec(5))
    fl = L.Flat(f)

    def snapshot(**kw):
        P = make_params(**{**dict(coverage=30.0, seed=3), **kw})
        info, tiles = fl.plan(P)
        th = fl.plan_thinning(P)
        order = np.argsort(tiles["id"])
        o2 = np.argsort(th["id"])
        return (info.n_tiles_total, info.n_templates_total, tuple(tiles["begin"][order]), tuple(tiles["len"][order]),
                tuple(tiles["templates"][order]), tuple(th["u_len"][o2]), tuple(th["tail_off"][o2]), tuple(th["n_useful"][o2]))

    fresh = {}
    variants = [dict(), dict(read_size=100), dict(insert_size_mean=300), dict(chr_mask=[1, 0, 1]), dict(read_size=37)]
    for i, kw in enumerate(variants):  # every variant on a forest that has kept nothing
        fl_i = L.Flat(f)
        P = make_params(**{**dict(coverage=30.0, seed=3), **kw})
        info, tiles = fl_i.plan(P)
        th = fl_i.plan_thinning(P)
        order, o2 = np.argsort(tiles["id"]), np.argsort(th["id"])
        fresh[i] = (info.n_tiles_total, info.n_templates_total, tuple(tiles["begin"][order]), tuple(tiles["len"][order]),
                    tuple(tiles["templates"][order]), tuple(th["u_len"][o2]), tuple(th["tail_off"][o2]), tuple(th["n_useful"][o2]))
    # the same variants one after the other, forwards and backwards, on ONE forest: what it kept must never leak
    for i in list(range(len(variants))) + list(reversed(range(len(variants)))):
        assert snapshot(**variants[i]) == fresh[i], variants[i]
    assert fresh[0] != fresh[1] and fresh[0] != fresh[3]
    # regrouped samples: another number of output samples, possibly another tile width
    groups = (np.arange(f.n_leaves) % 2).astype(np.uint32)
    fl.set_groups(groups, 2)
    a = snapshot()
    fl2 = L.Flat(f)
    fl2.set_groups(groups, 2)
    P = make_params(coverage=30.0, seed=3)
    info, tiles = fl2.plan(P)
    assert a[0] == info.n_tiles_total and a[1] == info.n_templates_total
    fl.set_groups(None, 0)
    assert snapshot() == fresh[0]


def test_planner_template_counts_follow_the_multinomial_law():
    """A9: the templates of a (sample, chromosome) are a multinomial over its tiles with weights length x cell weight,
    drawn as two levels of binomial chains (plan_rng.hpp).  Normal sample only: every tile's weight is its length,
    so tile i of chromosome c gets Binomial(N_c, len_i / chr_len_c) templates -- z-scores over 300 seeds."""
    f = synth_forest(small_spec(6))
    fl = L.Flat(f)
    R, cov = 100, 400.0
    z, totals_ok = [], True
    for seed in range(300):
        P = make_params(coverage=cov, read_size=R, normal_only=1, with_normal_sample=0, seed=seed)
        info, t = fl.plan(P)
        for c in range(f.n_chr):
            sel = t["chr"] == c
            n_c = round(cov * int(f.chr_len[c]) / R)
            totals_ok &= int(t["templates"][sel].sum()) == n_c and int(t["len"][sel].sum()) == int(f.chr_len[c])
            p = t["len"][sel] / float(f.chr_len[c])
            z.extend((t["templates"][sel] - n_c * p) / np.sqrt(n_c * p * (1 - p)))
    z = np.asarray(z)
    assert totals_ok
    assert len(z) > 3000 and abs(z.mean()) < 0.06 and abs(z.std() - 1) < 0.05 and np.abs(z).max() < 5.5, (z.mean(), z.std(), np.abs(z).max())
    # different seeds give different splits, the same seed the same split
    a = fl.plan(make_params(coverage=cov, read_size=R, normal_only=1, with_normal_sample=0, seed=1))[1]["templates"]
    b = fl.plan(make_params(coverage=cov, read_size=R, normal_only=1, with_normal_sample=0, seed=1))[1]["templates"]
    c = fl.plan(make_params(coverage=cov, read_size=R, normal_only=1, with_normal_sample=0, seed=2))[1]["templates"]
    assert np.array_equal(a, b) and not np.array_equal(a, c)
