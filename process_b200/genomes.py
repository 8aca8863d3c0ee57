"""Explicit per-cell genomes: the form in which the reference's seam hands a forest to the sequencing simulator
(forest.get_sample_mutations_list() / get_normal_sample(), src/seq_simulation.cpp:566-575; every genome is walked
chromosome -> allele -> fragment -> SID, src/phylogenetic_forest.cpp:279-290).  Arrays in the layout of
`pcs_cell_genomes_desc` (include/pcs_seq.h); the mutation table, the germline and the row annotation come from
`source` (a PhylogeneticForest or anything with the same attributes)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi as A


@dataclass(eq=False)
class CellGenomes:
    source: object               # chr_names, chr_len, chr_n_alleles, sample_names, mutation table, germline, annotation
    n_cells: int
    cell_sample: np.ndarray      # u32 [n_cells]
    n_normal_preneo: int
    allele_cell: np.ndarray      # u32 [n_alleles]
    allele_chr: np.ndarray       # u16
    allele_id: np.ndarray        # u16
    allele_origin: np.ndarray    # u8
    allele_frag_off: np.ndarray  # u64 [n_alleles+1]
    frag_begin: np.ndarray       # u32
    frag_end: np.ndarray         # u32
    allele_sid_off: np.ndarray   # u64 [n_alleles+1]
    sid_row: np.ndarray          # u32

    def __getattr__(self, name):
        # sizes, names, the mutation table and the row annotation are the source's
        if name in ("source", "__setstate__"):
            raise AttributeError(name)
        return getattr(self.source, name)

    @property
    def n_leaves(self):
        return self.n_cells

    @property
    def leaf_sample(self):
        return self.cell_sample

    def host_bytes(self) -> int:
        names = ["cell_sample", "allele_cell", "allele_chr", "allele_id", "allele_origin", "allele_frag_off", "frag_begin",
                 "frag_end", "allele_sid_off", "sid_row"]
        src = ["chr_len", "chr_n_alleles", "mut_chr", "mut_pos", "mut_ref_len", "mut_alt_len", "germ_mut", "germ_allele_mask"]
        return int(sum(getattr(self, k).nbytes for k in names) + sum(getattr(self.source, k).nbytes for k in src))

    def as_genomes_desc(self) -> A.CellGenomesDesc:
        s = self.source
        s.normalise()
        spec = dict(cell_sample="<u4", allele_cell="<u4", allele_chr="<u2", allele_id="<u2", allele_origin="u1",
                    allele_frag_off="<u8", frag_begin="<u4", frag_end="<u4", allele_sid_off="<u8", sid_row="<u4")
        for k, dt in spec.items():
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=dt))
        d = A.CellGenomesDesc()
        d.n_chr = s.n_chr
        d.chr_len = A.ptr(s.chr_len, C.c_uint32)
        d.chr_n_alleles = A.ptr(s.chr_n_alleles, C.c_uint8)
        d.n_samples = s.n_samples
        d.n_cells = self.n_cells
        d.cell_sample = A.ptr(self.cell_sample, C.c_uint32)
        d.n_normal_preneo = self.n_normal_preneo
        d.n_alleles = len(self.allele_cell)
        d.allele_cell = A.ptr(self.allele_cell, C.c_uint32)
        d.allele_chr = A.ptr(self.allele_chr, C.c_uint16)
        d.allele_id = A.ptr(self.allele_id, C.c_uint16)
        d.allele_origin = A.ptr(self.allele_origin, C.c_uint8)
        d.allele_frag_off = A.ptr(self.allele_frag_off, C.c_uint64)
        d.frag_begin = A.ptr(self.frag_begin, C.c_uint32)
        d.frag_end = A.ptr(self.frag_end, C.c_uint32)
        d.allele_sid_off = A.ptr(self.allele_sid_off, C.c_uint64)
        d.sid_row = A.ptr(self.sid_row, C.c_uint32)
        d.n_mut = s.n_mut
        d.mut_chr = A.ptr(s.mut_chr, C.c_uint16)
        d.mut_pos = A.ptr(s.mut_pos, C.c_uint32)
        d.mut_ref_len = A.ptr(s.mut_ref_len, C.c_uint8)
        d.mut_alt_len = A.ptr(s.mut_alt_len, C.c_uint8)
        d.n_germline = len(s.germ_mut)
        d.germ_mut = A.ptr(s.germ_mut, C.c_uint32)
        d.germ_allele_mask = A.ptr(s.germ_allele_mask, C.c_uint8)
        return d
