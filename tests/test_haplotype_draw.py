"""The sampler's haplotype draw, evaluated on the host (pcs_flat_draw runs dev.hpp::exact_leaf, the very inline
function the kernels compile): inside a sampling entry every haplotype must own floor(width / n) or ceil(width / n)
of the entry's draw words -- every cell / allele of a class equiprobable, whatever the purity
(/root/reference/src/sequencing.cpp:155-163, wiring src/seq_simulation.cpp:572-578).

Round 1 mapped the draw with a 32-bit scale, umulhi(u - base, floor(n * 2^32 / width)): exact only when an entry
owned the whole draw range; at purity 0.8 the contaminant's two germline alleles were drawn 55.6 : 44.4."""
import numpy as np
import pytest

from process_b200 import _lib as L
from process_b200.synth import synth_forest

from conftest import make_params, small_spec


@pytest.mark.parametrize("purity", [0.5, 0.7, 0.8, 0.9, 1.0])
@pytest.mark.parametrize("preneo", [0, 1])
def test_every_haplotype_of_an_entry_owns_the_same_number_of_draw_words(purity, preneo):
    f = synth_forest(small_spec(3))
    flat = L.Flat(f)
    P = make_params(coverage=30.0, purity=purity, preneoplastic_in_normal=preneo)
    info, tiles = flat.plan(P)
    rng = np.random.default_rng(1)
    # tumour samples: a tile in a plain stretch and tiles inside CNA pieces (most entries)
    n_ent = {}
    for tid in tiles["id"][tiles["sample"] < f.n_samples][:400]:
        n_ent[int(tid)] = len(flat.tile_entries(P, int(tid))["thr"])
    picks = sorted(n_ent, key=lambda t: -n_ent[t])[:3] + sorted(n_ent, key=lambda t: n_ent[t])[:1]
    checked = 0
    for tid in picks:
        E = flat.tile_entries(P, tid)
        base = 0
        for e in range(len(E["thr"])):
            lo, hi, n = base, int(E["thr"][e]), int(E["list_n"][e])
            width = hi - lo + 1
            base = hi + 1
            members = flat.hap_list(int(E["list_off"][e]), n)
            # (1) the map is floor(x * n / width) exactly: random words, the ends, and both sides of every boundary
            x = np.unique(np.concatenate([
                rng.integers(0, width, 4096), [0, width - 1],
                np.clip(np.concatenate([(np.arange(1, n) * width + n - 1) // n + d for d in (-1, 0)]), 0, width - 1)]))
            hap, ent = flat.draw(P, tid, (x + lo).astype(np.uint32))
            assert (ent == e).all()
            want = members[(x.astype(object) * n // width).astype(np.int64)]
            assert np.array_equal(hap, want), (purity, tid, e)
            # (2) so every leaf owns floor or ceil of width / n words
            first = (np.arange(0, n + 1).astype(object) * width + n - 1) // n  # first x with floor(x n / width) = k
            owned = np.diff(first.astype(np.int64))
            assert owned.min() >= width // n and owned.max() <= -(-width // n)
            checked += 1
    assert checked >= 4


def test_entry_widths_follow_the_purity_weights():
    """share of the draw range of the contaminant class = (1 - purity) * normal DNA / all DNA of the piece"""
    f = synth_forest(small_spec(3, clone_cna=0, wgd_clones=0))  # every haplotype whole: weights are head counts
    flat = L.Flat(f)
    for purity in (0.3, 0.8):
        P = make_params(coverage=30.0, purity=purity)
        info, tiles = flat.plan(P)
        tid = int(tiles["id"][(tiles["sample"] == 0) & (tiles["chr"] == 0)][0])
        E = flat.tile_entries(P, tid)
        assert len(E["thr"]) == 2
        n_t = int((f.leaf_sample == 0).sum())
        # tumour cells: 2 alleles each, weight purity / n_t; the normal cell: 2 alleles, weight 1 - purity
        share_t = (int(E["thr"][0]) + 1) / 2.0**32
        assert int(E["list_n"][0]) == 2 * n_t and int(E["list_n"][1]) == 2
        assert abs(share_t - purity) < 1e-9
