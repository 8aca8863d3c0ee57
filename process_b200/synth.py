"""Seeded synthetic phylogenetic forests shaped like the BASELINE.json configs.

No RACES, no downloads: a Yule cell tree, clone founders carrying driver SNVs /
CNAs / WGD, Poisson passenger SNVs and indels on every branch, a germline with
het/hom SNPs, pre-neoplastic SIDs on the root (vignettes/mutations.Rmd:106-148,
345-347 give the shape of the demo; SURVEY.md 8(d) gives the sizes).
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field

import numpy as np

from . import _abi as A
from .forest import PhylogeneticForest

GRCH38_LEN = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973,
              145138636, 138394717, 133797422, 135086622, 133275309, 114364328, 107043718,
              101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468,
              156040895, 57227415]
GRCH38_NAMES = [str(i) for i in range(1, 23)] + ["X", "Y"]
CHR22_GRCH37_LEN = 51304566  # demo reference (src/mutation_engine.cpp:80)


@dataclass
class SynthSpec:
    chr_names: list = field(default_factory=lambda: ["22"])
    chr_len: list = field(default_factory=lambda: [CHR22_GRCH37_LEN])
    chr_n_alleles: list | None = None
    sample_cells: list = field(default_factory=lambda: [100, 100, 560, 560])
    sample_names: list | None = None
    germline_density: float = 1e-3
    germline_hom_frac: float = 0.4
    germline_indel_frac: float = 0.05
    n_preneo_snv: int = 1000
    n_preneo_indel: int = 500
    trunk_snv: int = 0              # extra clonal passengers on the root
    node_snv_mean: float = 8.0      # Poisson mean of passenger SIDs per non-root node
    indel_frac: float = 0.1
    n_clones: int = 2
    clone_cna: int = 2              # passenger CNAs per clone
    wgd_clones: int = 1
    cna_len: tuple = (100_000, 2_000_000)
    seed: int = 0


def _yule_tree(n_leaves: int, rng) -> np.ndarray:
    parent = [-1]
    leaves = [0]
    for _ in range(n_leaves - 1):
        i = int(rng.integers(len(leaves)))
        v = leaves[i]
        a = len(parent)
        parent.append(v)
        parent.append(v)
        leaves[i] = a
        leaves.append(a + 1)
    return np.asarray(parent, dtype=np.int32)


def _dfs(parent: np.ndarray):
    n = len(parent)
    children = [[] for _ in range(n)]
    roots = []
    for v in range(n):
        p = int(parent[v])
        (roots if p < 0 else children[p]).append(v)
    tin = np.zeros(n, np.int64)
    tout = np.zeros(n, np.int64)
    leaf_order = []
    t = 0
    for r in roots:
        stack = [(r, 0)]
        while stack:
            v, i = stack.pop()
            if i == 0:
                tin[v] = t
                t += 1
                if not children[v]:
                    leaf_order.append(v)
            if i < len(children[v]):
                stack.append((v, i + 1))
                stack.append((children[v][i], 0))
            else:
                tout[v] = t
    return children, roots, tin, tout, np.asarray(leaf_order)


class _Karyo:
    """alleles and their fragments per chromosome, as the CNA events of a lineage leave them."""

    def __init__(self, chr_len, chr_n_alleles):
        self.frags = [{a: [(1, int(L))] for a in range(int(n))} for L, n in zip(chr_len, chr_n_alleles)]
        self.next_id = [int(n) for n in chr_n_alleles]

    def clone(self):
        return copy.deepcopy(self)

    @staticmethod
    def _clip(fr, lo, hi):
        return [(max(b, lo), min(e, hi)) for b, e in fr if not (e < lo or b > hi)]

    def covering(self, c, lo, hi):
        return [a for a, fr in self.frags[c].items() if self._clip(fr, lo, hi)]

    def amp(self, c, lo, hi, src):
        d = self.next_id[c]
        self.next_id[c] += 1
        self.frags[c][d] = self._clip(self.frags[c][src], lo, hi)
        return d

    def dele(self, c, lo, hi, a, clen):
        fr = self.frags[c][a]
        self.frags[c][a] = (self._clip(fr, 1, lo - 1) if lo > 1 else []) + \
                           (self._clip(fr, hi + 1, clen) if hi < clen else [])

    def wgd(self):
        for c in range(len(self.frags)):
            for a in sorted(self.frags[c].keys()):
                self.frags[c][self.next_id[c]] = list(self.frags[c][a])
                self.next_id[c] += 1


def synth_forest(spec: SynthSpec) -> PhylogeneticForest:
    rng = np.random.default_rng(spec.seed)
    chr_len = np.asarray(spec.chr_len, dtype=np.int64)
    n_chr = len(chr_len)
    n_all = np.asarray(spec.chr_n_alleles if spec.chr_n_alleles is not None else [2] * n_chr, dtype=np.uint8)
    cum = np.concatenate([[0], np.cumsum(chr_len)])
    total_len = int(cum[-1])

    n_leaves = int(sum(spec.sample_cells))
    parent = _yule_tree(n_leaves, rng)
    n_nodes = len(parent)
    children, roots, tin, tout, leaf_order = _dfs(parent)
    root = roots[0]

    # samples: contiguous DFS blocks (clade-enriched, like spatial boxes) with 20% mixing
    lab = np.repeat(np.arange(len(spec.sample_cells)), spec.sample_cells)
    mix = rng.random(n_leaves) < 0.2
    lab_mix = lab[mix]
    rng.shuffle(lab_mix)
    lab[mix] = lab_mix
    leaf_node = leaf_order.astype(np.uint32)
    leaf_sample = lab.astype(np.uint32)

    # ---- structural events (drivers, CNAs, WGD) on clone founders and their descendants
    sub_leaves = np.zeros(n_nodes, np.int64)
    for v in range(n_nodes - 1, -1, -1):
        if not children[v]:
            sub_leaves[v] = 1
        if parent[v] >= 0:
            sub_leaves[parent[v]] += sub_leaves[v]
    cand = np.where((sub_leaves >= max(2, n_leaves // 20)) & (sub_leaves <= max(2, n_leaves // 2)) &
                    (np.arange(n_nodes) != root))[0]
    if len(cand) == 0:
        cand = np.asarray([c for c in children[root]] or [root])
    founders = rng.choice(cand, size=min(spec.n_clones, len(cand)), replace=False) if spec.n_clones else []
    by_tin = np.argsort(tin)
    struct = {}  # node -> list of ("amp"|"del"|"wgd"|"drv")
    for k, f in enumerate(founders):
        f = int(f)
        ev = struct.setdefault(f, [])
        ev += ["drv", "amp", "del"]
        if k < spec.wgd_clones:
            ev.append("wgd")
        desc = by_tin[tin[f]:tout[f]]  # the founder and all its descendants
        for v in rng.choice(desc, size=spec.clone_cna, replace=True) if spec.clone_cna else []:
            struct.setdefault(int(v), []).append("amp" if rng.random() < 0.5 else "del")

    # events are collected as rows [node, class, seq, kind, chr, pos, len, allele, dest, key, nature]
    E = []
    states = [_Karyo(chr_len, n_all)]
    state_of = np.zeros(n_nodes, np.int64)
    drv_keys = []
    chr_p = chr_len / total_len
    for v in range(n_nodes):
        p = int(parent[v])
        sid = state_of[p] if p >= 0 else 0
        if v in struct:
            st = states[sid].clone()
            seq = 0
            for what in struct[v]:
                if what == "wgd":
                    st.wgd()
                    E.append([v, 1, seq, A.PCS_EV_WGD, 0, 0, 0, 0, 0, -1, A.PCS_NATURE_DRIVER])
                elif what == "drv":
                    c = int(rng.choice(n_chr, p=chr_p))
                    pos = int(rng.integers(1, chr_len[c] + 1))
                    ids = st.covering(c, pos, pos)
                    if not ids:
                        continue
                    a = int(rng.choice(ids))
                    key = int(cum[c]) + pos - 1
                    drv_keys.append(key)
                    E.append([v, 1, seq, A.PCS_EV_SID, c, 0, 0, a, 0, key, A.PCS_NATURE_DRIVER])
                else:
                    c = int(rng.choice(n_chr, p=chr_p))
                    L = int(min(rng.integers(spec.cna_len[0], spec.cna_len[1] + 1), chr_len[c] // 4))
                    L = max(L, 1)
                    pos = int(rng.integers(1, chr_len[c] - L + 2))
                    ids = st.covering(c, pos, pos + L - 1)
                    if not ids:
                        continue
                    a = int(rng.choice(ids))
                    nat = A.PCS_NATURE_DRIVER if v in set(int(x) for x in founders) else A.PCS_NATURE_PASSENGER
                    if what == "amp":
                        d = st.amp(c, pos, pos + L - 1, a)
                        E.append([v, 1, seq, A.PCS_EV_CNA_AMP, c, pos, L, a, d, -1, nat])
                    else:
                        st.dele(c, pos, pos + L - 1, a, int(chr_len[c]))
                        E.append([v, 1, seq, A.PCS_EV_CNA_DEL, c, pos, L, a, 0, -1, nat])
                seq += 1
            states.append(st)
            sid = len(states) - 1
        state_of[v] = sid

    # ---- how many SIDs of each kind
    n_germ = int(round(spec.germline_density * total_len))
    cnt = rng.poisson(spec.node_snv_mean, n_nodes).astype(np.int64)
    cnt[root] = spec.trunk_snv
    n_pass = int(cnt.sum())
    n_pre = spec.n_preneo_snv + spec.n_preneo_indel
    need = n_germ + n_pass + n_pre
    # distinct genome-wide positions (also distinct from the driver SNVs)
    keys = np.unique(rng.integers(0, total_len, size=int(need * 1.02) + 64, dtype=np.int64))
    if drv_keys:
        keys = np.setdiff1d(keys, np.asarray(drv_keys, dtype=np.int64))
    if len(keys) < need:
        raise ValueError("genome too small for the requested number of distinct SID positions")
    keys = rng.permutation(keys)[:need]
    k_germ, k_pass, k_pre = keys[:n_germ], keys[n_germ:n_germ + n_pass], keys[n_germ + n_pass:]

    # passengers: node, chromosome, allele (uniform over the alleles of the node's karyotype)
    p_node = np.repeat(np.arange(n_nodes), cnt)
    p_chr = np.searchsorted(cum, k_pass, side="right") - 1
    p_allele = np.zeros(n_pass, np.int64)
    grp = state_of[p_node] * n_chr + p_chr
    order = np.argsort(grp, kind="stable")
    gs = grp[order]
    bounds = np.flatnonzero(np.diff(gs)) + 1
    for lo, hi in zip(np.concatenate([[0], bounds]), np.concatenate([bounds, [len(gs)]])):
        if hi <= lo:
            continue
        g = int(gs[lo])
        ids = sorted(states[g // n_chr].frags[g % n_chr].keys())
        p_allele[order[lo:hi]] = rng.choice(ids, size=hi - lo)
    pre_chr = np.searchsorted(cum, k_pre, side="right") - 1
    pre_allele = np.array([rng.integers(0, n_all[c]) for c in pre_chr], dtype=np.int64) if n_pre else np.zeros(0, np.int64)

    # ---- mutation table: every distinct key, sorted
    all_keys = np.concatenate([k_germ, k_pass, k_pre, np.asarray(drv_keys, dtype=np.int64)])
    kind_of = np.concatenate([np.full(n_germ, 0), np.full(n_pass, 1), np.full(n_pre, 2),
                              np.full(len(drv_keys), 3)])
    srt = np.argsort(all_keys, kind="stable")
    row_key = all_keys[srt]
    n_mut = len(row_key)
    rank = np.empty(n_mut, np.int64)
    rank[srt] = np.arange(n_mut)
    mut_chr = (np.searchsorted(cum, row_key, side="right") - 1).astype(np.uint16)
    mut_pos = (row_key - cum[mut_chr] + 1).astype(np.uint32)
    row_kind = kind_of[srt]

    is_indel = np.zeros(n_mut, bool)
    u = rng.random(n_mut)
    is_indel |= (row_kind == 0) & (u < spec.germline_indel_frac)
    is_indel |= (row_kind == 1) & (u < spec.indel_frac)
    pre_rows = rank[n_germ + n_pass:n_germ + n_pass + n_pre]
    is_indel[pre_rows[spec.n_preneo_snv:]] = True
    k_len = np.minimum(50, rng.geometric(0.3, n_mut)).astype(np.uint8)
    is_del = rng.random(n_mut) < 0.5
    ref_len = np.where(is_indel & is_del, 1 + k_len, 1).astype(np.uint8)
    alt_len = np.where(is_indel & ~is_del, 1 + k_len, 1).astype(np.uint8)
    # keep deletions inside the chromosome
    over = mut_pos.astype(np.int64) + ref_len - 1 > chr_len[mut_chr]
    ref_len[over] = 1
    ref_code = rng.integers(0, 4, n_mut).astype(np.uint8)
    alt_code = ((ref_code + rng.integers(1, 4, n_mut)) & 3).astype(np.uint8)

    nature_bit = np.array([A.PCS_NATURE_GERMINAL, A.PCS_NATURE_PASSENGER, A.PCS_NATURE_PRENEOPLASTIC,
                           A.PCS_NATURE_DRIVER])
    mut_nature_mask = (1 << nature_bit[row_kind]).astype(np.uint8)
    cause_names = ["SBS1", "SBS13", "SBS5", "ID2", "ID13"]
    mut_cause = np.full(n_mut, -1, np.int16)
    som = (row_kind == 1) | (row_kind == 2)
    mut_cause[som & ~is_indel] = rng.integers(0, 3, int((som & ~is_indel).sum()))
    mut_cause[som & is_indel] = rng.integers(3, 5, int((som & is_indel).sum()))

    # ---- germline
    germ_mut = rank[:n_germ].astype(np.uint32)
    gchr = mut_chr[germ_mut]
    hom = rng.random(n_germ) < spec.germline_hom_frac
    het_allele = rng.integers(0, 2, n_germ)
    germ_mask = np.where(n_all[gchr] == 1, 1, np.where(hom, 3, 1 << het_allele)).astype(np.uint8)
    # listed in row order, one entry per row: what a shim gets when it walks its std::map<SID, ...> mutation table,
    # and the flattener's fast path (any order is accepted; tests/test_host_logic.py shuffles it)
    by_row = np.argsort(germ_mut, kind="stable")
    germ_mut, germ_mask = germ_mut[by_row], germ_mask[by_row]

    # ---- event table
    cols = 11
    blocks = []
    if E:
        Es = np.asarray(E, dtype=np.int64)
        drv = Es[:, 3] == A.PCS_EV_SID
        if drv.any():
            Es[drv, 9] = np.searchsorted(row_key, Es[drv, 9])
        blocks.append(Es)
    if n_pass:
        P = np.zeros((n_pass, cols), np.int64)
        P[:, 0] = p_node
        P[:, 1] = 2
        P[:, 2] = np.arange(n_pass)
        P[:, 3] = A.PCS_EV_SID
        P[:, 4] = p_chr
        P[:, 7] = p_allele
        P[:, 9] = rank[n_germ:n_germ + n_pass]
        P[:, 10] = A.PCS_NATURE_PASSENGER
        blocks.append(P)
    if n_pre:
        Q = np.zeros((n_pre, cols), np.int64)
        Q[:, 0] = root
        Q[:, 1] = 0
        Q[:, 2] = np.arange(n_pre)
        Q[:, 3] = A.PCS_EV_SID
        Q[:, 4] = pre_chr
        Q[:, 7] = pre_allele
        Q[:, 9] = pre_rows
        Q[:, 10] = A.PCS_NATURE_PRENEOPLASTIC
        blocks.append(Q)
    ev = np.concatenate(blocks) if blocks else np.zeros((0, cols), np.int64)
    ev = ev[np.lexsort((ev[:, 2], ev[:, 1], ev[:, 0]))]
    node_event_off = np.zeros(n_nodes + 1, np.uint64)
    np.add.at(node_event_off, ev[:, 0] + 1, 1)
    node_event_off = np.cumsum(node_event_off).astype(np.uint64)

    f = PhylogeneticForest(
        chr_names=list(spec.chr_names), chr_len=chr_len.astype(np.uint32), chr_n_alleles=n_all,
        node_parent=parent,
        sample_names=list(spec.sample_names) if spec.sample_names else [f"S_{i + 1}" for i in range(len(spec.sample_cells))],
        leaf_node=leaf_node, leaf_sample=leaf_sample, node_event_off=node_event_off,
        ev_kind=ev[:, 3], ev_chr=ev[:, 4], ev_pos=ev[:, 5], ev_len=ev[:, 6], ev_allele=ev[:, 7],
        ev_dest=ev[:, 8], ev_mut=np.maximum(ev[:, 9], 0), ev_nature=ev[:, 10],
        mut_chr=mut_chr, mut_pos=mut_pos, mut_ref_len=ref_len, mut_alt_len=alt_len,
        germ_mut=germ_mut, germ_allele_mask=germ_mask,
        mut_ref_code=ref_code, mut_alt_code=alt_code, mut_cause=mut_cause,
        mut_nature_mask=mut_nature_mask, cause_names=cause_names)
    return f.normalise()


# --------------------------------------------------------------------- configs
def config_spec(name: str, seed: int = 0, scale: float = 1.0) -> SynthSpec:
    """SynthSpec of a BASELINE.json config ("C1".."C5").  `scale` < 1 shrinks the
    genome (chromosome lengths) for CPU-sized parity cases; cells and densities stay."""
    def L(x):
        return [max(10_000, int(v * scale)) for v in x]
    male = [2] * 22 + [1, 1]
    if name in ("C1", "C2"):
        return SynthSpec(chr_names=["22"], chr_len=L([CHR22_GRCH37_LEN]), sample_cells=[100, 100, 560, 560],
                         sample_names=["S_1_1", "S_1_2", "S_2_1", "S_2_2"],
                         germline_density=1e-3, n_preneo_snv=1000, n_preneo_indel=500, node_snv_mean=6.0,
                         n_clones=2, clone_cna=1, wgd_clones=1, cna_len=(200_000, 200_000), seed=seed)
    if name == "C3":
        return SynthSpec(chr_names=GRCH38_NAMES, chr_len=L(GRCH38_LEN), chr_n_alleles=male,
                         sample_cells=[1000, 1000, 1000], germline_density=4.5e6 / 3088269832,
                         n_preneo_snv=1000, n_preneo_indel=500, node_snv_mean=8.0, n_clones=3,
                         clone_cna=20, wgd_clones=0, cna_len=(100_000, 20_000_000), seed=seed)
    if name == "C4":
        return SynthSpec(chr_names=GRCH38_NAMES, chr_len=L(GRCH38_LEN), chr_n_alleles=male,
                         sample_cells=[5000] * 8, germline_density=4.5e6 / 3088269832,
                         n_preneo_snv=1000, n_preneo_indel=500, node_snv_mean=4.0, n_clones=6,
                         clone_cna=20, wgd_clones=1, cna_len=(100_000, 20_000_000), seed=seed)
    if name == "C5":
        return SynthSpec(chr_names=GRCH38_NAMES, chr_len=L(GRCH38_LEN), chr_n_alleles=male,
                         sample_cells=[100_000], germline_density=4.5e6 / 3088269832,
                         n_preneo_snv=1000, n_preneo_indel=500, trunk_snv=900_000, node_snv_mean=5.0,
                         n_clones=8, clone_cna=25, wgd_clones=8, cna_len=(100_000, 20_000_000), seed=seed)
    raise ValueError(f"unknown config {name}")


def write_reference_fasta(forest: PhylogeneticForest, path: str, seed: int = 0, line: int = 60) -> str:
    """a random reference genome with the forest's chromosome names and lengths (there is no network for the
    real FASTA); also sets forest.reference_path."""
    rng = np.random.default_rng(seed)
    with open(path, "w") as fh:
        for name, n in zip(forest.chr_names, forest.chr_len):
            seq = np.frombuffer(b"ACGT", dtype="S1")[rng.integers(0, 4, int(n))].tobytes().decode()
            fh.write(f">{name} synthetic\n")
            for i in range(0, len(seq), line):
                fh.write(seq[i:i + line] + "\n")
    forest.reference_path = path
    return path


def read_fasta(path: str) -> dict:
    out, name, parts = {}, None, []
    with open(path) as fh:
        for ln in fh:
            ln = ln.strip()
            if ln.startswith(">"):
                if name is not None:
                    out[name] = "".join(parts)
                name, parts = ln[1:].split()[0], []
            else:
                parts.append(ln)
    if name is not None:
        out[name] = "".join(parts)
    return out
