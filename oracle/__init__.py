"""ctypes binding of the CPU oracle.  TEST INFRASTRUCTURE: import only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

from process_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_last_error.restype = C.c_char_p
    return _LIB


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise OracleError(lib().oracle_last_error().decode())


def set_placement_rule(rule):
    """A11 (UNPINNED): 0 = start uniform over the fragment, templates past its end dropped (the product's rule);
    1 = start uniform over the valid starts.  Returns the previous rule."""
    return int(lib().oracle_set_placement_rule(C.c_int(rule)))


def n_out_samples(forest, params: A.SeqParams, n_groups=None):
    if params.normal_only:
        return 1
    g = forest.n_samples if n_groups is None else n_groups
    return g + (1 if params.with_normal_sample else 0)


def simulate(forest, params: A.SeqParams, leaf_group=None, n_groups=None, n_threads=1, trace_cap=0,
             trace_masks=False):
    """returns dict(occ, cov [n_out, n_mut], n_reads, trace, masks)."""
    d = forest.as_desc()
    n_out = n_out_samples(forest, params, n_groups)
    occ = np.zeros((n_out, forest.n_mut), np.uint32)
    cov = np.zeros((n_out, forest.n_mut), np.uint32)
    rec = np.zeros(trace_cap, A.PLACEMENT_DTYPE) if trace_cap else None
    masks = np.zeros((trace_cap, A.PCS_ERRMASK_WORDS), np.uint32) if (trace_cap and trace_masks) else None
    tn = C.c_uint64(0)
    nr = C.c_uint64(0)
    lg = None if leaf_group is None else np.ascontiguousarray(leaf_group, dtype=np.uint32)
    _check(lib().oracle_simulate(
        C.byref(d), C.byref(params), A.ptr(lg, C.c_uint32), C.c_uint32(n_groups or 0), C.c_uint32(n_threads),
        A.ptr(occ, C.c_uint32), A.ptr(cov, C.c_uint32),
        C.cast(rec.ctypes.data if rec is not None else None, C.c_void_p),
        A.ptr(masks, C.c_uint32), C.c_uint64(trace_cap), C.byref(tn), C.byref(nr)))
    if rec is not None and tn.value > trace_cap:
        raise OracleError(f"trace capacity {trace_cap} < {tn.value} reads")
    tm = (C.c_double * 4)()
    lib().oracle_last_timing(tm)
    return dict(occ=occ, cov=cov, n_reads=nr.value,
                timing=dict(genomes_cpu_s=tm[0], fixed_cpu_s=tm[1], read_loop_cpu_s=tm[2], wall_s=tm[3]),
                trace=None if rec is None else rec[:tn.value],
                masks=None if masks is None else masks[:tn.value])


def count_injected(forest, n_out, read_size, placements, err_masks=None):
    d = forest.as_desc()
    placements = np.ascontiguousarray(placements, dtype=A.PLACEMENT_DTYPE)
    occ = np.zeros((n_out, forest.n_mut), np.uint32)
    cov = np.zeros((n_out, forest.n_mut), np.uint32)
    em = None if err_masks is None else np.ascontiguousarray(err_masks, dtype=np.uint32)
    _check(lib().oracle_count_injected(
        C.byref(d), C.c_uint32(n_out), C.c_uint32(read_size), C.c_void_p(placements.ctypes.data),
        A.ptr(em, C.c_uint32), C.c_uint64(len(placements)), A.ptr(occ, C.c_uint32), A.ptr(cov, C.c_uint32)))
    return occ, cov


def cell_genome(forest, which, cell, chrom, cap=1 << 16):
    """explicit genome of a cell on one chromosome:
    (fragments [(allele, origin, begin, end)], somatic SIDs [(allele, row)])."""
    d = forest.as_desc()
    fa = np.zeros(cap, np.uint16); fo = np.zeros(cap, np.uint16)
    fb = np.zeros(cap, np.uint32); fe = np.zeros(cap, np.uint32)
    sa = np.zeros(cap, np.uint16); sr = np.zeros(cap, np.uint32)
    nf = C.c_uint32(0); ns = C.c_uint32(0)
    _check(lib().oracle_cell_genome(
        C.byref(d), C.c_uint32(which), C.c_uint32(cell), C.c_uint32(chrom), C.c_uint32(cap),
        A.ptr(fa, C.c_uint16), A.ptr(fo, C.c_uint16), A.ptr(fb, C.c_uint32), A.ptr(fe, C.c_uint32),
        C.byref(nf), A.ptr(sa, C.c_uint16), A.ptr(sr, C.c_uint32), C.byref(ns)))
    if nf.value > cap or ns.value > cap:
        raise OracleError("cell_genome capacity exceeded")
    frags = [(int(fa[i]), int(fo[i]), int(fb[i]), int(fe[i])) for i in range(nf.value)]
    sids = [(int(sa[i]), int(sr[i])) for i in range(ns.value)]
    return frags, sids


def materialize(forest, ref_off, ref_bases: bytes, alt_off, alt_bytes: bytes, read_size, sequencer, error_rate,
                placements, err_masks=None):
    """SAM content of a placement list: (seq [n, R] uint8, qual, cigar [n, 16], n_cigar, lengths)."""
    d = forest.as_desc()
    placements = np.ascontiguousarray(placements, dtype=A.PLACEMENT_DTYPE)
    n = len(placements)
    seq = np.zeros((n, read_size), np.uint8); qual = np.zeros((n, read_size), np.uint8)
    cigar = np.zeros((n, 16), np.uint32); nc = np.zeros(n, np.uint32); ln = np.zeros(n, np.uint32)
    ro = np.ascontiguousarray(ref_off, dtype=np.uint64); ao = np.ascontiguousarray(alt_off, dtype=np.uint32)
    em = None if err_masks is None else np.ascontiguousarray(err_masks, dtype=np.uint32)
    _check(lib().oracle_materialize(
        C.byref(d), A.ptr(ro, C.c_uint64), C.c_char_p(ref_bases), A.ptr(ao, C.c_uint32), C.c_char_p(alt_bytes),
        C.c_uint32(read_size), C.c_uint32(sequencer), C.c_double(error_rate), C.c_void_p(placements.ctypes.data),
        A.ptr(em, C.c_uint32), C.c_uint64(n), A.ptr(seq, C.c_uint8), A.ptr(qual, C.c_uint8), A.ptr(cigar, C.c_uint32),
        A.ptr(nc, C.c_uint32), A.ptr(ln, C.c_uint32)))
    return seq, qual, cigar, nc, ln


def chr_genomes(forest, chrom, with_preneo=True):
    """every explicit genome of one chromosome: dict of CSR arrays (allele_cell, allele_id, allele_origin,
    allele_frag_off, frag_begin, frag_end, allele_sid_off, sid_row); cells >= n_leaves are the normal cells with the
    pre-neoplastic SIDs (one per root)"""
    d = forest.as_desc()
    n = (C.c_uint64 * 3)()
    null = lambda ct: C.cast(None, C.POINTER(ct))
    _check(lib().oracle_chr_genomes(C.byref(d), C.c_uint32(chrom), C.c_int(1 if with_preneo else 0), C.c_uint64(0),
                                    C.c_uint64(0), C.c_uint64(0), null(C.c_uint32), null(C.c_uint16), null(C.c_uint8),
                                    null(C.c_uint64), null(C.c_uint32), null(C.c_uint32), null(C.c_uint64),
                                    null(C.c_uint32), n))
    na, nf, ns = int(n[0]), int(n[1]), int(n[2])
    out = dict(allele_cell=np.zeros(na, np.uint32), allele_id=np.zeros(na, np.uint16), allele_origin=np.zeros(na, np.uint8),
               allele_frag_off=np.zeros(na + 1, np.uint64), frag_begin=np.zeros(max(nf, 1), np.uint32),
               frag_end=np.zeros(max(nf, 1), np.uint32), allele_sid_off=np.zeros(na + 1, np.uint64),
               sid_row=np.zeros(max(ns, 1), np.uint32))
    _check(lib().oracle_chr_genomes(C.byref(d), C.c_uint32(chrom), C.c_int(1 if with_preneo else 0), C.c_uint64(max(na, 1)),
                                    C.c_uint64(max(nf, 1)), C.c_uint64(max(ns, 1)), A.ptr(out["allele_cell"], C.c_uint32),
                                    A.ptr(out["allele_id"], C.c_uint16), A.ptr(out["allele_origin"], C.c_uint8),
                                    A.ptr(out["allele_frag_off"], C.c_uint64), A.ptr(out["frag_begin"], C.c_uint32),
                                    A.ptr(out["frag_end"], C.c_uint32), A.ptr(out["allele_sid_off"], C.c_uint64),
                                    A.ptr(out["sid_row"], C.c_uint32), n))
    out["frag_begin"], out["frag_end"], out["sid_row"] = out["frag_begin"][:nf], out["frag_end"][:nf], out["sid_row"][:ns]
    return out


def cell_genomes(forest, with_preneo=True):
    """the forest as explicit per-cell genomes (process_b200.genomes.CellGenomes): what the reference's
    get_sample_mutations_list() / get_normal_sample() hand to the sequencing simulator"""
    from process_b200.genomes import CellGenomes
    parts = [chr_genomes(forest, c, with_preneo) for c in range(forest.n_chr)]
    cat = lambda k, dt: np.concatenate([p[k] for p in parts]).astype(dt)
    frag_off = np.concatenate([[0]] + [p["allele_frag_off"][1:] + sum(len(q["frag_begin"]) for q in parts[:i])
                                       for i, p in enumerate(parts)]).astype(np.uint64)
    sid_off = np.concatenate([[0]] + [p["allele_sid_off"][1:] + sum(len(q["sid_row"]) for q in parts[:i])
                                      for i, p in enumerate(parts)]).astype(np.uint64)
    n_roots = int((np.asarray(forest.node_parent) < 0).sum())
    return CellGenomes(
        source=forest, n_cells=forest.n_leaves, cell_sample=np.asarray(forest.leaf_sample, np.uint32),
        n_normal_preneo=n_roots if with_preneo else 0,
        allele_cell=cat("allele_cell", np.uint32),
        allele_chr=np.concatenate([np.full(len(p["allele_cell"]), c, np.uint16) for c, p in enumerate(parts)]),
        allele_id=cat("allele_id", np.uint16), allele_origin=cat("allele_origin", np.uint8),
        allele_frag_off=frag_off, frag_begin=cat("frag_begin", np.uint32), frag_end=cat("frag_end", np.uint32),
        allele_sid_off=sid_off, sid_row=cat("sid_row", np.uint32))
