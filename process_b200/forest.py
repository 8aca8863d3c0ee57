"""Flat, event-labelled phylogenetic forest: the host-side view handed to the C ABI.

This is what the Rcpp shim would build from `PhylogeneticForest` by walking the
forest nodes (reference: src/seq_simulation.cpp:566-575 obtains the per-cell
genomes; src/phylogenetic_forest.cpp:279-376 shows the chromosome -> allele ->
fragment -> SID traversal).  It holds numpy arrays in exactly the layout of
`pcs_forest_desc` (include/pcs_seq.h) plus the strings that never leave the host.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _abi as A

_BASES = np.array(list("ACGT"))


@dataclass(eq=False)
class PhylogeneticForest:
    chr_names: list
    chr_len: np.ndarray          # u32 [n_chr]
    chr_n_alleles: np.ndarray    # u8  [n_chr]
    node_parent: np.ndarray      # i32 [n_nodes]
    sample_names: list
    leaf_node: np.ndarray        # u32 [n_leaves]
    leaf_sample: np.ndarray      # u32 [n_leaves]
    node_event_off: np.ndarray   # u64 [n_nodes+1]
    ev_kind: np.ndarray          # u8
    ev_chr: np.ndarray           # u16
    ev_pos: np.ndarray           # u32
    ev_len: np.ndarray           # u32
    ev_allele: np.ndarray        # u16
    ev_dest: np.ndarray          # u16
    ev_mut: np.ndarray           # u32
    ev_nature: np.ndarray        # u8
    mut_chr: np.ndarray          # u16 [n_mut], sorted by (chr, pos)
    mut_pos: np.ndarray          # u32
    mut_ref_len: np.ndarray      # u8
    mut_alt_len: np.ndarray      # u8
    germ_mut: np.ndarray         # u32
    germ_allele_mask: np.ndarray # u8
    # host-only annotation of the rows
    mut_ref_code: np.ndarray = None    # u8: first base of ref
    mut_alt_code: np.ndarray = None    # u8: first base of alt (SNV) / repeated base (indel)
    mut_cause: np.ndarray = None       # i16 index into cause_names, -1 = NA
    mut_nature_mask: np.ndarray = None # u8 bit set over PCS_NATURE_*
    cause_names: list = field(default_factory=list)
    reference_path: str | None = None
    leaf_attrs: dict = field(default_factory=dict)   # per sampled cell: epistate, mutant, species, birth_time
    _keep: list = field(default_factory=list, repr=False)

    # ------------------------------------------------------------------ sizes
    @property
    def n_chr(self): return len(self.chr_len)
    @property
    def n_nodes(self): return len(self.node_parent)
    @property
    def n_leaves(self): return len(self.leaf_node)
    @property
    def n_samples(self): return len(self.sample_names)
    @property
    def n_mut(self): return len(self.mut_pos)
    @property
    def n_events(self): return len(self.ev_kind)

    def normalise(self):
        """cast every array to the ABI dtype, C-contiguous."""
        spec = dict(chr_len="<u4", chr_n_alleles="u1", node_parent="<i4", leaf_node="<u4",
                    leaf_sample="<u4", node_event_off="<u8", ev_kind="u1", ev_chr="<u2",
                    ev_pos="<u4", ev_len="<u4", ev_allele="<u2", ev_dest="<u2", ev_mut="<u4",
                    ev_nature="u1", mut_chr="<u2", mut_pos="<u4", mut_ref_len="u1",
                    mut_alt_len="u1", germ_mut="<u4", germ_allele_mask="u1")
        for k, dt in spec.items():
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=dt))
        n = self.n_mut
        if self.mut_ref_code is None:
            self.mut_ref_code = np.zeros(n, "u1")
        if self.mut_alt_code is None:
            self.mut_alt_code = np.ones(n, "u1")
        if self.mut_cause is None:
            self.mut_cause = np.full(n, -1, "<i2")
        if self.mut_nature_mask is None:
            self.mut_nature_mask = np.zeros(n, "u1")
        return self

    def host_bytes(self) -> int:
        """bytes of the arrays that cross the C ABI (what an upload copies from)."""
        names = ["chr_len", "chr_n_alleles", "node_parent", "leaf_node", "leaf_sample", "node_event_off",
                 "ev_kind", "ev_chr", "ev_pos", "ev_len", "ev_allele", "ev_dest", "ev_mut", "ev_nature",
                 "mut_chr", "mut_pos", "mut_ref_len", "mut_alt_len", "germ_mut", "germ_allele_mask"]
        return int(sum(getattr(self, k).nbytes for k in names))

    def as_desc(self) -> A.ForestDesc:
        self.normalise()
        d = A.ForestDesc()
        d.n_chr = self.n_chr
        d.chr_len = A.ptr(self.chr_len, C.c_uint32)
        d.chr_n_alleles = A.ptr(self.chr_n_alleles, C.c_uint8)
        d.n_nodes = self.n_nodes
        d.node_parent = A.ptr(self.node_parent, C.c_int32)
        d.n_samples = self.n_samples
        d.n_leaves = self.n_leaves
        d.leaf_node = A.ptr(self.leaf_node, C.c_uint32)
        d.leaf_sample = A.ptr(self.leaf_sample, C.c_uint32)
        d.n_events = self.n_events
        d.node_event_off = A.ptr(self.node_event_off, C.c_uint64)
        d.ev_kind = A.ptr(self.ev_kind, C.c_uint8)
        d.ev_chr = A.ptr(self.ev_chr, C.c_uint16)
        d.ev_pos = A.ptr(self.ev_pos, C.c_uint32)
        d.ev_len = A.ptr(self.ev_len, C.c_uint32)
        d.ev_allele = A.ptr(self.ev_allele, C.c_uint16)
        d.ev_dest = A.ptr(self.ev_dest, C.c_uint16)
        d.ev_mut = A.ptr(self.ev_mut, C.c_uint32)
        d.ev_nature = A.ptr(self.ev_nature, C.c_uint8)
        d.n_mut = self.n_mut
        d.mut_chr = A.ptr(self.mut_chr, C.c_uint16)
        d.mut_pos = A.ptr(self.mut_pos, C.c_uint32)
        d.mut_ref_len = A.ptr(self.mut_ref_len, C.c_uint8)
        d.mut_alt_len = A.ptr(self.mut_alt_len, C.c_uint8)
        d.n_germline = len(self.germ_mut)
        d.germ_mut = A.ptr(self.germ_mut, C.c_uint32)
        d.germ_allele_mask = A.ptr(self.germ_allele_mask, C.c_uint8)
        return d

    # ---------------------------------------------------- row annotation (host)
    # vectorised: a WGS result has millions of rows (src/seq_simulation.cpp:52-90 builds them one by one)
    def row_strings(self, rows: np.ndarray):
        """(ref, alt) strings of the given rows.  SNV: one base each.  Deletion:
        anchor + run, alt = anchor.  Insertion: ref = anchor, alt = anchor + run."""
        rows = np.asarray(rows)
        rc, ac = self.mut_ref_code[rows] & 3, self.mut_alt_code[rows] & 3
        rl, al = self.mut_ref_len[rows].astype(np.int64), self.mut_alt_len[rows].astype(np.int64)
        ref = _BASES[rc].astype(object)
        alt = _BASES[ac].astype(object)
        for i in np.flatnonzero((rl != 1) | (al != 1)):  # indels are a small minority
            a, b = _BASES[rc[i]], _BASES[ac[i]]
            ref[i] = a + b * (int(rl[i]) - 1)
            alt[i] = a + b * (int(al[i]) - 1)
        return ref, alt

    def row_string_codes(self, rows: np.ndarray):
        """The same strings as row_strings(), dictionary-encoded: (ref_codes, ref_table, alt_codes, alt_table) with
        ref = ref_table[ref_codes], alt = alt_table[alt_codes].  A row's strings depend on (first base, repeated
        base, length) only, so the tables are tiny and nothing is done per row in Python."""
        rows = np.asarray(rows)
        rl, al = self.mut_ref_len[rows], self.mut_alt_len[rows]
        base = ((self.mut_ref_code[rows] & 3) | ((self.mut_alt_code[rows] & 3) << 2)).astype(np.uint16)
        base |= ((rl != 1) | (al != 1)).astype(np.uint16) << 4
        out = []
        for length in (rl, al):
            key = base | (length.astype(np.uint16) << 5)   # < 2^13: lengths are u8
            present = np.flatnonzero(np.bincount(key, minlength=1 << 13))
            index = np.zeros(1 << 13, np.int32)
            index[present] = np.arange(len(present), dtype=np.int32)
            table = np.empty(len(present), dtype=object)
            for j, k in enumerate(present):
                a, b, is_indel, n = _BASES[k & 3], _BASES[(k >> 2) & 3], (k >> 4) & 1, int(k >> 5)
                if is_indel:
                    table[j] = str(a) + str(b) * (n - 1)
                else:  # SNV: ref is the first base, alt the other one
                    table[j] = str(a) if length is rl else str(b)
            out += [index[key], table]
        return tuple(out)

    def alt_table(self):
        """(alt_off [n_mut+1] uint32, alt_bytes) of every row, the layout pcs_forest_set_alt takes."""
        _, alt = self.row_strings(np.arange(self.n_mut))
        lens = self.mut_alt_len.astype(np.int64)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
        return off, "".join(alt).encode()

    def row_causes(self, rows: np.ndarray):
        c = self.mut_cause[np.asarray(rows)]
        table = np.asarray(list(self.cause_names) + [None], dtype=object)
        return table[np.where(c < 0, len(self.cause_names), c)]

    def row_classes(self, rows: np.ndarray):
        table = np.asarray([";".join(sorted(A.NATURE_DESCRIPTIONS[b] for b in range(4) if (m >> b) & 1))
                            for m in range(16)], dtype=object)
        return table[self.mut_nature_mask[np.asarray(rows)] & 15]
