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


# ----------------------------------------------------------------------------------------------------------------
# The general form: SNVs AND indels, both error models, paired reads, regrouped samples (FACS), normal_only.
#
# A haplotype fragment [b, e] with its carried SIDs is a sequence of TOKENS: a reference position that is present,
# or an inserted base.  A SID at p with (ref_len r, alt_len a) is its anchor token (position p, the occurrence),
# a - 1 inserted tokens, and the next token is position p + r.  A read starting at a
# present reference position x is the R tokens from x's token on (fewer if the fragment ends); it covers the
# reference positions among them; it carries a SID iff the anchor token is among them, and under an error model
# the occurrence survives iff none of the SID's tokens the read holds is a sequencing error (A14).  A start inside
# the stretch a carried deletion removed does not see that deletion (the read begins after its anchor): those few
# starts are walked one by one.  Expected counts are sums over all starts of (expected reads starting there) x
# (what such a read adds) -- exhaustive, no sampling, nothing shared with the oracle's walk or the kernels'.
def _ramp(i, R):
    return 0.5 + i / (R - 1) if R > 1 else np.ones_like(i, dtype=np.float64)


def _base_error(i, R, error_rate, random_quality):
    """P(read base i is a sequencing error), A14: constant model error_rate; random-quality model
    E[min(1, error_rate * ramp(i) * LogNormal(-sigma^2/2, sigma))] = error_rate * ramp(i) while the cap is out of reach"""
    i = np.asarray(i, np.float64)
    if not random_quality:
        return np.full(i.shape, float(error_rate))
    p = error_rate * _ramp(i, R)
    assert (p < 0.2).all(), "closed form ignores the cap at 1: keep error_rate small"
    return p


def _walk_one(x, R, e_frag, sids, pos_of):
    """reference positions covered and (row, first offset, tokens held) of the SIDs carried by a read of R bases from
    reference position x on a fragment ending at e_frag; sids: sorted [(pos, row, ref_len, alt_len)] with pos >= x"""
    covered, carried = [], []
    q, rem = x, R
    for pos, row, rl, al in sids:
        if pos < q:
            continue
        if rem == 0 or pos - q >= rem or pos > e_frag:
            break
        covered += list(range(q, pos + 1))
        rem -= pos - q
        held = min(al, rem)
        carried.append((row, R - rem, held))
        rem -= held
        q = pos + rl
    if rem > 0 and q <= e_frag:
        covered += list(range(q, min(q + rem - 1, e_frag) + 1))
    return covered, carried


def expected_tables_general(f, coverage, purity, R, with_normal=True, insert=None, preneoplastic_in_normal=False,
                            error_rate=0.0, random_quality=False, leaf_group=None, n_groups=None, normal_only=False):
    """E[coverage], E[occurrences] of every (output sample, row).  leaf_group / n_groups: the FACS repartition of
    the sampled cells (default: the forest's samples); normal_only: simulate_normal_seq (one sample, purity ignored)"""
    mates = 2 if insert else 1
    ks, pk = insert_law(*insert) if insert else (np.zeros(1, np.int64), np.ones(1))
    group = np.asarray(f.leaf_sample if leaf_group is None else leaf_group)
    n_s = f.n_samples if leaf_group is None else n_groups
    samples = [] if normal_only else [("tumour", s) for s in range(n_s)]
    if normal_only or with_normal:
        samples.append(("normal", None))
    S = len(samples)
    e_cov = np.zeros((S, f.n_mut))
    e_occ = np.zeros((S, f.n_mut))
    germ = {int(m): int(mask) for m, mask in zip(f.germ_mut, f.germ_allele_mask)}
    pos_of, rl_of, al_of = f.mut_pos.astype(np.int64), f.mut_ref_len.astype(np.int64), f.mut_alt_len.astype(np.int64)
    n_roots = int((f.node_parent < 0).sum())
    for c in range(f.n_chr):
        rows_c = np.flatnonzero(f.mut_chr == c)
        N = int(np.floor(coverage * int(f.chr_len[c]) / (R * mates) + 0.5))
        clen = int(f.chr_len[c])
        cache = {}

        def molecules(kind, cell):
            if (kind, cell) not in cache:
                frags, sids = oracle.cell_genome(f, kind, cell, c)
                som = {}
                for a, r in sids:
                    som.setdefault(a, []).append(int(r))
                out = []
                for a, o, b, e in frags:
                    if b == 0:
                        continue
                    rows = [r for r in som.get(a, []) if b <= pos_of[r] <= e]
                    rows += [int(r) for r in rows_c if ((germ.get(int(r), 0) >> o) & 1) and b <= pos_of[r] <= e]
                    out.append((b, e, sorted(set(rows), key=lambda r: pos_of[r])))
                cache[(kind, cell)] = out
            return cache[(kind, cell)]

        def contribution(b, e, rows):
            """per unit of (reads per start): arrays over rows_c of depth and occurrences from all starts of this molecule"""
            depth = np.zeros(clen + 2)          # by reference position
            occ = {}
            sids = [(int(pos_of[r]), r, int(rl_of[r]), int(al_of[r])) for r in rows]
            # expected reads per start position x, per unit: first mates and second mates
            x = np.arange(b, e + 1)
            s = np.zeros(len(x))
            for k, pr in zip(ks, pk):
                tlen = R if not insert else 2 * R + int(k)
                s += pr * (x <= e - tlen + 1)
                if insert:
                    first = x - R - int(k)
                    s += pr * ((first >= b) & (first <= e - tlen + 1))
            # tokens of the fragment as a read that started before every SID sees it.  A SID inside the stretch an
            # earlier carried deletion removed is not part of it (only reads starting inside the stretch meet it)
            deleted = np.zeros(e - b + 2, bool)  # reference positions a carried deletion removed (index x - b)
            pieces, anchor, at = [], {}, b
            for p, r, rl, al in sids:
                if p < at:
                    continue
                pieces.append(np.arange(at, p + 1))
                anchor[r] = sum(len(q) for q in pieces) - 1
                if al > 1:
                    pieces.append(np.full(al - 1, -1))
                deleted[p - b + 1:min(p + rl, e + 1) - b] = True
                at = p + rl
            if at <= e:
                pieces.append(np.arange(at, e + 1))
            tok = np.concatenate(pieces) if pieces else np.zeros(0, np.int64)
            is_ref = tok >= 0
            idx_of = np.full(e - b + 2, -1)
            idx_of[tok[is_ref] - b] = np.flatnonzero(is_ref)
            # starts on present positions: weight per token
            w_tok = np.zeros(len(tok))
            on = idx_of[x - b] >= 0
            w_tok[idx_of[x[on] - b]] = s[on]
            cum = np.concatenate([[0.0], np.cumsum(w_tok)])
            t = np.arange(len(tok))
            reads_over = cum[t + 1] - cum[np.maximum(t - R + 1, 0)]   # reads whose R tokens include token t
            depth[tok[is_ref]] += reads_over[is_ref]
            for p, r, rl, al in sids:
                if r not in anchor:
                    continue
                ta = anchor[r]
                lo = max(ta - R + 1, 0)
                offs = ta - np.arange(lo, ta + 1)                    # read offset of the anchor for a start at token lo..ta
                surv = np.ones(len(offs))
                if error_rate > 0:
                    for kk in range(al):
                        held = offs + kk < R
                        surv *= np.where(held, 1 - _base_error(np.minimum(offs + kk, R - 1), R, error_rate, random_quality), 1.0)
                occ[r] = occ.get(r, 0.0) + float((w_tok[lo:ta + 1] * surv).sum())
            # starts inside a deleted stretch: the read does not see that deletion
            for xx in x[~on]:
                wgt = s[xx - b]
                if wgt == 0:
                    continue
                covered, carried = _walk_one(int(xx), R, e, [q for q in sids if q[0] >= xx], pos_of)
                for y in covered:
                    depth[y] += wgt
                for r, off0, held in carried:
                    sv = 1.0
                    if error_rate > 0:
                        sv = float(np.prod(1 - _base_error(np.arange(off0, off0 + held), R, error_rate, random_quality)))
                    occ[r] = occ.get(r, 0.0) + wgt * sv
            return depth, occ

        for si, (what, s_id) in enumerate(samples):
            cells = [] if what == "normal" else [l for l in range(f.n_leaves) if group[l] == s_id]
            p = purity if cells else 0.0
            mol = []
            if p > 0:
                for l in cells:
                    mol += [(p / len(cells),) + m for m in molecules(A.PCS_PLACE_TUMOUR, l)]
            if p < 1:
                if preneoplastic_in_normal:
                    for r in range(n_roots):
                        mol += [((1 - p) / n_roots,) + m for m in molecules(A.PCS_PLACE_NORMAL_PRENEO, r)]
                else:
                    mol += [((1 - p),) + m for m in molecules(A.PCS_PLACE_NORMAL_PLAIN, 0)]
            W = sum(w * (e - b + 1) for w, b, e, rows in mol)
            memo = {}
            for w, b, e, rows in mol:
                key = (b, e, tuple(rows))
                if key not in memo:
                    memo[key] = contribution(b, e, rows)
                depth, occ = memo[key]
                e_cov[si, rows_c] += N / W * w * depth[pos_of[rows_c]]
                for r, v in occ.items():
                    e_occ[si, r] += N / W * w * v
    return e_cov, e_occ


def indel_spec(seed=6, **kw):
    """SNVs and indels, germline and somatic, CNAs and a WGD: the forests the general closed form is for"""
    from conftest import small_spec
    d = dict(chr_names=["1", "X"], chr_len=[60_000, 40_000], chr_n_alleles=[2, 1], sample_cells=[5, 6],
             germline_density=4e-3, germline_indel_frac=0.3, n_preneo_snv=15, n_preneo_indel=15, indel_frac=0.3,
             node_snv_mean=6, n_clones=2, clone_cna=3, wgd_clones=1, cna_len=(3000, 30000))
    d.update(kw)
    return small_spec(seed, **d)
