"""Hand-computed micro forest: the known-answer test that pins the semantics
(DESIGN.md "Semantics") for BOTH the CPU oracle and the CUDA path.

One chromosome of 1000 bp, two germline alleles, root + two sampled cells.

rows (sorted by position)
  0  pos 100  SNV        germline, allele 0
  1  pos 200  SNV        germline, both alleles
  2  pos 300  deletion   ref 5 bases -> alt 1 base; pre-neoplastic on the root, allele 1
  3  pos 302  SNV        germline, allele 1   (inside the reference span row 2 deletes)
  4  pos 400  SNV        passenger in cell 0, allele 0
  5  pos 600  insertion  ref 1 base -> alt 4 bases; passenger in cell 1, allele 1
  6  pos 800  SNV        passenger in cell 1 on allele 2 (the amplified copy)
  7  pos 905  SNV        germline, allele 0

events
  root   : SID row 2 on allele 1 (pre-neoplastic)
  cell 0 : SID row 4 on allele 0
  cell 1 : AMP [700,899] of allele 0 -> allele 2 ; SID row 6 on allele 2 ;
           SID row 5 on allele 1 ; DEL [1,150] of allele 0

explicit genomes this gives
  cell 0  allele 0 [1,1000]   somatic {400: row 4}           germline rows 0, 1, 7
          allele 1 [1,1000]   somatic {300: row 2}           germline rows 1, 3
  cell 1  allele 0 [151,1000] somatic {}                     germline rows 1, 7 (row 0 deleted)
          allele 1 [1,1000]   somatic {300: row 2, 600: row 5}  germline rows 1, 3
          allele 2 [700,899]  somatic {800: row 6}           germline (origin 0): none in range
  normal (plain)   alleles 0, 1 whole, germline only
  normal (preneo)  allele 1 additionally carries row 2

reads (read_size 10), their sample, and what each must add:
  A s0 cell0 a0 @95   spans 95..104            depth[100]            occ row0
  B s0 cell0 a1 @95   spans 95..104            depth[100]
  C s0 cell0 a1 @296  296..300 then deletion -> 305..309
                                               depth[300]            occ row2   (302 is deleted: no depth)
  D s0 cell0 a0 @296  spans 296..305           depth[300], depth[302]
  K s0 cell0 a0 @395  spans 395..404, error at read offset 5 (= position 400)
                                               depth[400]            (occurrence removed by the error)
  L s0 cell0 a0 @395  error at offset 4        depth[400]            occ row4
  E s1 cell1 a1 @595  5 ref bases, then 4 inserted bases, then 1 ref base (601)
                                               depth[600]            occ row5
  F s1 cell1 a1 @598  2 ref bases, 4 inserted, 4 ref bases
                                               depth[600]            occ row5
  G s1 cell1 a2 @795  spans 795..804           depth[800]            occ row6
  H s1 cell1 a2 @895  clamped at the fragment end 899: nothing
  I s1 cell1 a0 @897  spans 897..906           depth[905]            occ row7
  J s1 cell1 a2 @897  clamped at 899: nothing (905 is past the fragment)
  M s2 normal plain  a1 @296  spans 296..305   depth[300], depth[302]   occ row3
  N s2 normal preneo a1 @296  like C           depth[300]            occ row2
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

FOREST = dict(
    chr_names=["1"], chr_len=[1000], chr_n_alleles=[2],
    node_parent=[-1, 0, 0], sample_names=["s0", "s1"], leaf_node=[1, 2], leaf_sample=[0, 1],
    node_event_off=[0, 1, 2, 6],
    #         root  c0   c1:AMP  SID6  SID5  DEL
    ev_kind=[0,    0,   1,      0,    0,    2],
    ev_chr=[0, 0, 0, 0, 0, 0],
    ev_pos=[0, 0, 700, 0, 0, 1],
    ev_len=[0, 0, 200, 0, 0, 150],
    ev_allele=[1, 0, 0, 2, 1, 0],
    ev_dest=[0, 0, 2, 0, 0, 0],
    ev_mut=[2, 4, 0, 6, 5, 0],
    ev_nature=[3, 1, 1, 1, 1, 1],
    mut_chr=[0] * 8, mut_pos=[100, 200, 300, 302, 400, 600, 800, 905],
    mut_ref_len=[1, 1, 5, 1, 1, 1, 1, 1], mut_alt_len=[1, 1, 1, 1, 1, 4, 1, 1],
    germ_mut=[0, 1, 3, 7], germ_allele_mask=[1, 3, 2, 1],
)

# (cell, start, chr, allele, sample, flags, error offsets)
READS = [
    (0, 95, 0, 0, 0, 0, []), (0, 95, 0, 1, 0, 0, []), (0, 296, 0, 1, 0, 0, []), (0, 296, 0, 0, 0, 0, []),
    (0, 395, 0, 0, 0, 0, [5]), (0, 395, 0, 0, 0, 0, [4]),
    (1, 595, 0, 1, 1, 0, []), (1, 598, 0, 1, 1, 0, []), (1, 795, 0, 2, 1, 0, []), (1, 895, 0, 2, 1, 0, []),
    (1, 897, 0, 0, 1, 0, []), (1, 897, 0, 2, 1, 0, []),
    (0, 296, 0, 1, 2, 1, []), (0, 296, 0, 1, 2, 2, []),
]
READ_SIZE = 10

#                 row: 0  1  2  3  4  5  6  7
EXPECTED_COV = [[2, 0, 2, 1, 2, 0, 0, 0],
                [0, 0, 0, 0, 0, 2, 1, 1],
                [0, 0, 2, 1, 0, 0, 0, 0]]
EXPECTED_OCC = [[1, 0, 1, 0, 1, 0, 0, 0],
                [0, 0, 0, 0, 0, 2, 1, 1],
                [0, 0, 1, 1, 0, 0, 0, 0]]

# explicit genomes: (which, cell) -> fragments [(allele, origin, begin, end)], somatic SIDs [(allele, row)]
EXPECTED_GENOMES = {
    "tumour:0": dict(frags=[(0, 0, 1, 1000), (1, 1, 1, 1000)], sids=[(0, 4), (1, 2)]),
    "tumour:1": dict(frags=[(0, 0, 151, 1000), (1, 1, 1, 1000), (2, 0, 700, 899)], sids=[(1, 2), (1, 5), (2, 6)]),
    "plain:0": dict(frags=[(0, 0, 1, 1000), (1, 1, 1, 1000)], sids=[]),
    "preneo:0": dict(frags=[(0, 0, 1, 1000), (1, 1, 1, 1000)], sids=[(1, 2)]),
}


# Reference for the SAM content checks: "ACGT" repeated (position p holds "ACGT"[(p-1) % 4]); alt strings below.
REF = ("ACGT" * 250)
ALT = ["C", "G", "A", "T", "G", "ATTT", "C", "G"]   # rows 0..7 (row 2: deletion AC..->A; row 5: insertion A->ATTT)
# read C: cell 0, allele 1, start 296 (read_size 10): ref 296..299 = "TACG", the deletion's anchor alt "A" (5M),
#         4 deleted bases (4D), ref 305..309 = "ACGTA" (5M)
# read E: cell 1, allele 1, start 595: ref 595..599 = "GTACG", insertion alt "ATTT" = 1M + 3I, ref 601 = "A"
# read A: cell 0, allele 0, start 95: ref 95..99 = "GTACG", SNV alt "C" at 100, ref 101..104 = "ACGT"
# read L: cell 0, allele 0, start 395, error at offset 4 (position 399, "G" -> "ACGT"[(2+1+4%3)&3] = "A"),
#         SNV alt "G" at 400
EXPECTED_SAM = {
    0: ("GTACGCACGT", "10M"),        # read A
    2: ("TACGAACGTA", "5M4D5M"),     # read C
    5: ("GTACAGACGT", "10M"),        # read L
    6: ("GTACGATTTA", "6M3I1M"),     # read E
}


def forest():
    from process_b200.forest import PhylogeneticForest
    kw = {k: (v if k in ("chr_names", "sample_names") else np.asarray(v)) for k, v in FOREST.items()}
    return PhylogeneticForest(**kw).normalise()


def placements():
    from process_b200 import _abi as A
    rec = np.zeros(len(READS), A.PLACEMENT_DTYPE)
    masks = np.zeros((len(READS), A.PCS_ERRMASK_WORDS), np.uint32)
    for i, (cell, start, chrom, allele, sample, flags, errs) in enumerate(READS):
        rec[i] = (cell, start, chrom, allele, sample, flags)
        for e in errs:
            masks[i, e >> 5] |= np.uint32(1 << (e & 31))
    return rec, masks


if __name__ == "__main__":
    with open(os.path.join(HERE, "micro_forest.json"), "w") as fh:
        json.dump(dict(forest=FOREST, reads=READS, read_size=READ_SIZE, expected_cov=EXPECTED_COV,
                       expected_occ=EXPECTED_OCC, expected_genomes=EXPECTED_GENOMES), fh, indent=1)
    print("wrote micro_forest.json")
