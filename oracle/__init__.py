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
    return dict(occ=occ, cov=cov, n_reads=nr.value,
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
