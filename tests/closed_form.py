"""Closed-form expectations of the count tables, from explicit per-cell genomes (test infrastructure).

For an SNV-only forest and single-end reads the rules of DESIGN.md section 3 (A9-A12) give, for sample s, chromosome
c and a locus at position x:

    N    = round(coverage * chr_len / R)                       templates of (s, c)
    W    = sum over molecules j of w_j * len_j                 DNA of the sample, weighted
    E[depth(x)] = N / W * sum_j w_j * #{starts p in [b_j, e_j] : p <= x <= p + R - 1 and p + R - 1 <= e_j}
    E[occ(m)]   = the same sum over the molecules that carry row m

with one molecule per (cell, allele, fragment [b, e]); w = purity / n_tumour_cells for a tumour cell of the sample
and (1 - purity) for the normal cell (germline alleles, whole).  Paired reads (A4, A16): N = round(coverage *
chr_len / (2 R)) templates of 2 R + k bases, k ~ Binomial(t = floor(mean / p), p = 1 - sd^2 / mean); the second mate
starts R + k bases after the first, both mates count, and the template must fit the fragment.  Nothing here shares code with the oracle's
sampler or the product: it uses only oracle.cell_genome() (the explicit genomes) and numpy."""
import numpy as np

import oracle
from process_b200 import _abi as A


def insert_law(mean, sd):
    """support and probabilities of the insert size, src/seq_simulation.cpp:431-451"""
    from scipy import stats
    p = 1 - sd * sd / mean
    t = int(mean / p)
    k = np.arange(0, t + 1)
    pk = stats.binom.pmf(k, t, p)
    keep = pk > 1e-15
    return k[keep], pk[keep] / pk[keep].sum()


def expected_tables(f, coverage, purity, R, with_normal=True, insert=None, preneoplastic_in_normal=False):
    """insert = (mean, sd) for paired reads, None for single reads; preneoplastic_in_normal: the normal cells are
    one per root, each with its root's pre-neoplastic SIDs (A10)"""
    assert ((f.mut_ref_len == 1) & (f.mut_alt_len == 1)).all(), "closed form is written for SNV-only forests"
    mates = 2 if insert else 1
    ks, pk = insert_law(*insert) if insert else (np.zeros(1, np.int64), np.ones(1))
    n_s = f.n_samples
    S = n_s + (1 if with_normal else 0)
    e_cov = np.zeros((S, f.n_mut))
    e_occ = np.zeros((S, f.n_mut))
    germ = {int(m): int(mask) for m, mask in zip(f.germ_mut, f.germ_allele_mask)}
    for c in range(f.n_chr):
        rows = np.flatnonzero(f.mut_chr == c)
        x = f.mut_pos[rows].astype(np.int64)
        N = int(np.floor(coverage * int(f.chr_len[c]) / (R * mates) + 0.5))

        def molecules(kind, cell):
            frags, sids = oracle.cell_genome(f, kind, cell, c)
            carried = {}
            for a, r in sids:
                carried.setdefault(a, set()).add(int(r))
            return [(o, b, e, carried.get(a, set())) for a, o, b, e in frags if b > 0]

        if preneoplastic_in_normal:
            n_roots = int((f.node_parent < 0).sum())
            normal = [(1.0 / n_roots,) + m for r in range(n_roots) for m in molecules(A.PCS_PLACE_NORMAL_PRENEO, r)]
        else:
            normal = [(1.0,) + m for m in molecules(A.PCS_PLACE_NORMAL_PLAIN, 0)]
        for s in range(S):
            cells = [] if s >= n_s else [l for l in range(f.n_leaves) if f.leaf_sample[l] == s]
            p = purity if cells else 0.0
            mol = []
            if p > 0:
                for l in cells:
                    mol += [(p / len(cells),) + m for m in molecules(A.PCS_PLACE_TUMOUR, l)]
            if p < 1:
                mol += [((1 - p) * m[0],) + m[1:] for m in normal]
            W = sum(w * (e - b + 1) for w, o, b, e, car in mol)
            for w, o, b, e, car in mol:
                n_starts = np.zeros(len(x))
                for k, pr in zip(ks, pk):
                    tlen = R if not insert else 2 * R + int(k)
                    last = e - tlen + 1  # last start whose template fits the fragment
                    n_starts += pr * np.maximum(0, np.minimum(x, last) - np.maximum(b, x - R + 1) + 1)
                    if insert:  # second mate: starts R + k after the first
                        n_starts += pr * np.maximum(0, np.minimum(x - R - k, last) - np.maximum(b, x - 2 * R - k + 1) + 1)
                has = np.array([(int(r) in car) or (((germ.get(int(r), 0) >> o) & 1) == 1 and b <= xx <= e)
                                for r, xx in zip(rows, x)], bool)
                e_cov[s, rows] += N / W * w * n_starts
                e_occ[s, rows] += N / W * w * n_starts * has
    return e_cov, e_occ


def expected_haplotype_reads(f, coverage, purity, R, with_normal=True, preneoplastic_in_normal=False):
    """single-end: E[reads placed] of every (sample, chromosome, PCS_PLACE_* kind, cell, allele) -- N / W * w *
    (fragment length - R + 1)+ summed over the allele's fragments (a start is uniform over the fragment and the read
    is dropped when it does not fit, A11).  Inside a class (the tumour cells of a sample; the normal cells) every
    cell weighs the same: src/sequencing.cpp:155-163."""
    n_s = f.n_samples
    S = n_s + (1 if with_normal else 0)
    out = {}
    for c in range(f.n_chr):
        N = int(np.floor(coverage * int(f.chr_len[c]) / R + 0.5))

        def molecules(kind, cell, weight):
            frags, _ = oracle.cell_genome(f, kind, cell, c)
            return [(weight, kind, cell, a, b, e) for a, o, b, e in frags if b > 0]

        if preneoplastic_in_normal:
            n_roots = int((f.node_parent < 0).sum())
            normal = [m for r in range(n_roots) for m in molecules(A.PCS_PLACE_NORMAL_PRENEO, r, 1.0 / n_roots)]
        else:
            normal = molecules(A.PCS_PLACE_NORMAL_PLAIN, 0, 1.0)
        for s in range(S):
            cells = [] if s >= n_s else [l for l in range(f.n_leaves) if f.leaf_sample[l] == s]
            p = purity if cells else 0.0
            mol = []
            if p > 0:
                for l in cells:
                    mol += molecules(A.PCS_PLACE_TUMOUR, l, p / len(cells))
            if p < 1:
                mol += [((1 - p) * m[0],) + m[1:] for m in normal]
            W = sum(w * (e - b + 1) for w, kind, cell, a, b, e in mol)
            for w, kind, cell, a, b, e in mol:
                key = (s, c, kind, cell, a)
                out[key] = out.get(key, 0.0) + N / W * w * max(0, e - b + 1 - R + 1)
    return out


def z_scores(obs, exp, min_expected=20.0):
    """Poisson z-scores of the cells whose expectation is large enough for the normal approximation; and the
    number of cells that are non-zero although their expectation is exactly zero"""
    ok = exp > min_expected
    z = (obs[ok].astype(np.float64) - exp[ok]) / np.sqrt(exp[ok])
    return z, int((obs[exp == 0] != 0).sum())


def snv_only_spec(seed=5, **kw):
    from conftest import small_spec
    d = dict(chr_names=["1", "X"], chr_len=[120_000, 80_000], chr_n_alleles=[2, 1], sample_cells=[5, 7],
             germline_density=3e-3, germline_indel_frac=0.0, n_preneo_snv=20, n_preneo_indel=0, indel_frac=0.0,
             node_snv_mean=6, n_clones=2, clone_cna=3, wgd_clones=1, cna_len=(3000, 40000))
    d.update(kw)
    return small_spec(seed, **d)


def cna_dense_spec(seed=9):
    """many short CNAs: most tiles draw from several sampling entries even at purity 1"""
    return snv_only_spec(seed, sample_cells=[6, 9], n_clones=3, clone_cna=12, wgd_clones=1, cna_len=(1500, 15000))
