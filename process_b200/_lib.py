"""ctypes loader and thin object wrappers of libpcs_seq.so (the C ABI of include/pcs_seq.h).

There is no CPU fallback: every compute entry point needs a B200 and raises
PcsError otherwise.  Only the pcs_flat_* introspection calls run without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCS_LIB: another build of the library (kernel experiments: process_b200/csrc/Makefile, `make variants`)
LIB_PATH = os.environ.get("PCS_LIB") or os.path.join(_HERE, "libpcs_seq.so")
_LIB = None

EXPORTS = [
    "pcs_abi_version", "pcs_last_error", "pcs_create", "pcs_destroy", "pcs_device_name",
    "pcs_forest_upload", "pcs_forest_free", "pcs_forest_set_groups", "pcs_forest_info",
    "pcs_plan_create", "pcs_plan_info_get", "pcs_plan_free", "pcs_plan_run", "pcs_plan_trace",
    "pcs_simulate", "pcs_count_injected", "pcs_active_rows",
    "pcs_shared_alloc", "pcs_shared_free", "pcs_shared_open", "pcs_shared_close", "pcs_enable_peer",
    "pcs_memset_u32", "pcs_memcpy_d2h", "pcs_plan_accumulate", "pcs_plan_finalize",
    "pcs_forest_replicate", "pcs_simulate_multi",
    "pcs_forest_set_reference", "pcs_forest_load_fasta", "pcs_forest_set_alt", "pcs_plan_write_sam",
    "pcs_plan_materialize",
    "pcs_flat_create", "pcs_flat_free", "pcs_flat_set_groups", "pcs_flat_info", "pcs_flat_cell_haps",
    "pcs_flat_fragset", "pcs_flat_hap_rows", "pcs_flat_plan", "pcs_flat_group_list",
    "pcs_flat_tile_entries", "pcs_flat_draw", "pcs_flat_hap_list",
    "pcs_host_gather", "pcs_host_string_column", "pcs_plan_counters", "pcs_memset_u32_stream",
    "pcs_forest_upload_genomes", "pcs_flat_create_genomes", "pcs_plan_coverage_track", "pcs_flat_plan_thinning",
    "pcs_simulate_result", "pcs_plan_result", "pcs_result_info", "pcs_result_fetch", "pcs_result_free",
    "pcs_host_binomial", "pcs_plan_finalize_stream", "pcs_forest_instances", "pcs_flat_instances",
]


class PcsError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libpcs_seq status {status}: {msg}")
        self.status = status
        self.message = msg


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C process_b200/csrc` "
                "(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        _LIB = C.CDLL(LIB_PATH)
        _LIB.pcs_last_error.restype = C.c_char_p
        if _LIB.pcs_abi_version() != A.PCS_ABI_VERSION:
            raise ImportError("libpcs_seq.so ABI version mismatch; rebuild it")
    return _LIB


def _ok(rc):
    if rc != 0:
        raise PcsError(rc, lib().pcs_last_error().decode())


def _u32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint32)


def host_binomial(seed, n, p, count):
    """`count` draws of the planner's Binomial(n, p) sampler (csrc/plan_rng.hpp); host only."""
    out = np.empty(count, np.uint64)
    _ok(lib().pcs_host_binomial(C.c_uint32(seed), C.c_uint64(n), C.c_double(p), C.c_uint64(count), A.ptr(out, C.c_uint64)))
    return out


# --------------------------------------------------------------------------- host-side column builders
def host_gather(rows, src):
    """src[rows] on all host cores (rows: uint32)."""
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    src = np.ascontiguousarray(src)
    dst = np.empty(len(rows), src.dtype)
    _ok(lib().pcs_host_gather(A.ptr(rows, C.c_uint32), C.c_uint64(len(rows)), C.c_void_p(src.ctypes.data),
                              C.c_uint32(src.dtype.itemsize), C.c_void_p(dst.ctypes.data)))
    return dst


def host_string_column(rows, codes, table):
    """the strings table[codes[rows]] (None in the table = NA) as Arrow large_string buffers:
    (offsets int64 [n+1], data uint8 [len], validity uint8 [ceil(n/8)] or None, null_count)."""
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    codes = np.ascontiguousarray(codes, dtype=np.uint16)
    n = len(rows)
    enc = [None if x is None else str(x).encode() for x in table]
    tab = (C.c_char_p * max(1, len(enc)))(*enc)
    offsets = np.empty(n + 1, np.int64)
    has_na = any(x is None for x in enc)
    validity = np.zeros((n + 7) // 8, np.uint8) if has_na else None
    length, nulls = C.c_uint64(0), C.c_uint64(0)
    args = lambda data, cap: (A.ptr(rows, C.c_uint32), C.c_uint64(n), A.ptr(codes, C.c_uint16), tab, C.c_uint32(len(enc)),
                              offsets.ctypes.data_as(C.POINTER(C.c_int64)), data, C.c_uint64(cap),
                              A.ptr(validity, C.c_uint8), C.byref(length), C.byref(nulls))
    bound = n * max([1] + [len(x) for x in enc if x is not None])
    if bound <= (1 << 31):  # one call: pages of `data` past the real length are never touched
        data = np.empty(max(1, bound), np.uint8)
        _ok(lib().pcs_host_string_column(*args(C.c_void_p(data.ctypes.data), len(data))))
    else:  # size it first
        _ok(lib().pcs_host_string_column(*args(None, 0)))
        data = np.empty(max(1, length.value), np.uint8)
        _ok(lib().pcs_host_string_column(*args(C.c_void_p(data.ctypes.data), len(data))))
    return offsets, data[:length.value], validity, nulls.value


# --------------------------------------------------------------------------- host-only view
class Flat:
    """flattened (haplotype-interval) view of a forest, host side only."""

    def __init__(self, forest):
        self.forest = forest
        self._h = C.c_void_p()
        if hasattr(forest, "as_genomes_desc"):  # explicit per-cell genomes (process_b200.genomes.CellGenomes)
            d = forest.as_genomes_desc()
            _ok(lib().pcs_flat_create_genomes(C.byref(d), C.byref(self._h)))
        else:
            d = forest.as_desc()
            _ok(lib().pcs_flat_create(C.byref(d), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().pcs_flat_free(self._h)
            self._h = None

    def set_groups(self, leaf_group, n_groups):
        lg = _u32(leaf_group)
        _ok(lib().pcs_flat_set_groups(self._h, A.ptr(lg, C.c_uint32), C.c_uint32(n_groups)))

    def info(self):
        out = (C.c_uint64 * 6)()
        _ok(lib().pcs_flat_info(self._h, out))
        return dict(n_loci=out[0], n_instances=out[1], n_haplotypes=out[2], n_fragment_sets=out[3], n_pieces=out[4])

    def instances(self):
        """(inst [n_instances, 4], locus_inst_off [n_loci + 1]) as the host flattener builds them"""
        i = self.info()
        inst = np.zeros((i["n_instances"], 4), np.uint32)
        off = np.zeros(i["n_loci"] + 1, np.uint32)
        _ok(lib().pcs_flat_instances(self._h, A.ptr(inst, C.c_uint32), A.ptr(off, C.c_uint32)))
        return inst, off

    def cell_haps(self, kind, cell, chrom, cap=4096):
        al = np.zeros(cap, np.uint16); hp = np.zeros(cap, np.uint32); fs = np.zeros(cap, np.uint32)
        n = C.c_uint32(0)
        _ok(lib().pcs_flat_cell_haps(self._h, C.c_uint32(kind), C.c_uint32(cell), C.c_uint32(chrom), C.c_uint32(cap),
                                     A.ptr(al, C.c_uint16), A.ptr(hp, C.c_uint32), A.ptr(fs, C.c_uint32), C.byref(n)))
        return [(int(al[i]), int(hp[i]), int(fs[i])) for i in range(n.value)]

    def fragset(self, fs, cap=4096):
        b = np.zeros(cap, np.uint32); e = np.zeros(cap, np.uint32); n = C.c_uint32(0)
        _ok(lib().pcs_flat_fragset(self._h, C.c_uint32(fs), C.c_uint32(cap), A.ptr(b, C.c_uint32),
                                   A.ptr(e, C.c_uint32), C.byref(n)))
        return [(int(b[i]), int(e[i])) for i in range(n.value)]

    def hap_rows(self, chrom, hap):
        cap = max(1, self.forest.n_mut)
        rows = np.zeros(cap, np.uint32); n = C.c_uint32(0)
        _ok(lib().pcs_flat_hap_rows(self._h, C.c_uint32(chrom), C.c_uint32(hap), C.c_uint32(cap),
                                    A.ptr(rows, C.c_uint32), C.byref(n)))
        return rows[:n.value].copy()

    def group_list(self, group, fragset, cap=1 << 16):
        """(offset in hap_list, haplotype indices) of the sampling list of (group, fragment set)"""
        h = np.zeros(cap, np.uint32); off = C.c_uint32(0); n = C.c_uint32(0)
        _ok(lib().pcs_flat_group_list(self._h, C.c_uint32(group), C.c_uint32(fragset), C.c_uint32(cap),
                                      A.ptr(h, C.c_uint32), C.byref(off), C.byref(n)))
        assert n.value <= cap
        return int(off.value), h[:n.value].copy()

    def plan(self, params: A.SeqParams, cap=1 << 22):
        info = A.PlanInfo()
        arrs = [np.zeros(cap, np.uint32) for _ in range(6)]
        _ok(lib().pcs_flat_plan(self._h, C.byref(params), C.byref(info), C.c_uint64(cap),
                                *[A.ptr(a, C.c_uint32) for a in arrs]))
        n = int(min(cap, info.n_tiles))
        keys = ["id", "templates", "sample", "chr", "begin", "len"]
        return info, {k: a[:n] for k, a in zip(keys, arrs)}


    def plan_thinning(self, params: A.SeqParams, cap=1 << 22):
        """per tile of the plan (keyed by tile id): thin, u_len, tail_off, n_useful (dev.hpp: Tile)"""
        arrs = [np.zeros(cap, np.uint32) for _ in range(5)]
        _ok(lib().pcs_flat_plan_thinning(self._h, C.byref(params), C.c_uint64(cap), *[A.ptr(a, C.c_uint32) for a in arrs]))
        info, _ = self.plan(params, cap=1)
        n = int(min(cap, info.n_tiles))
        return {k: a[:n] for k, a in zip(["id", "thin", "u_len", "tail_off", "n_useful"], arrs)}

    def hap_list(self, offset, n):
        h = np.zeros(n, np.uint32)
        _ok(lib().pcs_flat_hap_list(self._h, C.c_uint32(offset), C.c_uint32(n), A.ptr(h, C.c_uint32)))
        return h

    def tile_entries(self, params: A.SeqParams, tile_id, cap=64):
        """sampling entries of one tile: dict of thr, list_off, list_n, frag_end (uint32 arrays)"""
        arrs = [np.zeros(cap, np.uint32) for _ in range(4)]
        n = C.c_uint32(0)
        _ok(lib().pcs_flat_tile_entries(self._h, C.byref(params), C.c_uint32(tile_id), C.c_uint32(cap),
                                        *[A.ptr(a, C.c_uint32) for a in arrs], C.byref(n)))
        assert n.value <= cap
        return {k: a[:n.value] for k, a in zip(["thr", "list_off", "list_n", "frag_end"], arrs)}

    def draw(self, params: A.SeqParams, tile_id, u):
        """(haplotype index, entry index) the sampler's haplotype draw maps each 32-bit word of u to"""
        u = _u32(u)
        hap = np.zeros(len(u), np.uint32); ent = np.zeros(len(u), np.uint32)
        _ok(lib().pcs_flat_draw(self._h, C.byref(params), C.c_uint32(tile_id), C.c_uint64(len(u)),
                                A.ptr(u, C.c_uint32), A.ptr(hap, C.c_uint32), A.ptr(ent, C.c_uint32)))
        return hap, ent


# --------------------------------------------------------------------------- device objects
class Context:
    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        _ok(lib().pcs_create(C.byref(self._h), C.c_int(device), C.c_void_p(stream)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and _LIB is not None:  # _LIB is gone at interpreter shutdown
            _LIB.pcs_destroy(self._h)
            self._h = None

    __del__ = close

    # ---- tables other GPUs / processes can accumulate into over NVLink
    def shared_alloc(self, n_words):
        """(device pointer, 64-byte IPC handle) of n_words uint32 of plain cudaMalloc memory."""
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        _ok(lib().pcs_shared_alloc(self._h, C.c_size_t(4 * n_words), C.byref(ptr), handle))
        return ptr.value, bytes(handle)

    def shared_free(self, ptr):
        _ok(lib().pcs_shared_free(self._h, C.c_void_p(ptr)))

    def shared_open(self, handle: bytes):
        ptr = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        _ok(lib().pcs_shared_open(self._h, buf, C.byref(ptr)))
        return ptr.value

    def shared_close(self, ptr):
        _ok(lib().pcs_shared_close(self._h, C.c_void_p(ptr)))

    def enable_peer(self, device):
        _ok(lib().pcs_enable_peer(self._h, C.c_int(device)))

    def memset_u32(self, ptr, n_words, stream=None):
        if stream is None:
            _ok(lib().pcs_memset_u32(self._h, C.c_void_p(ptr), C.c_size_t(n_words)))
        else:
            _ok(lib().pcs_memset_u32_stream(self._h, C.c_void_p(ptr), C.c_size_t(n_words), C.c_void_p(stream)))

    def to_host(self, ptr, n_words):
        out = np.zeros(n_words, np.uint32)
        _ok(lib().pcs_memcpy_d2h(self._h, C.c_void_p(out.ctypes.data), C.c_void_p(ptr), C.c_size_t(4 * n_words)))
        return out

    def device_name(self):
        buf = C.create_string_buffer(256)
        _ok(lib().pcs_device_name(self._h, buf, C.c_size_t(256)))
        return buf.value.decode()


class Result:
    """compact, column-major result of one call, resident on the device until fetched
    (pcs_simulate_result / pcs_plan_result: the data frame's rows, assembled on the GPU)."""

    def __init__(self, handle):
        self._h = handle
        n, s, v = C.c_uint32(0), C.c_uint32(0), C.c_int(0)
        _ok(lib().pcs_result_info(self._h, C.byref(n), C.byref(s), C.byref(v)))
        self.n_rows, self.n_samples, self.has_vaf = n.value, s.value, bool(v.value)
        self.d2h_bytes = 0

    def fetch(self, vaf=True):
        """(rows uint32 [n], occurrences int32 [S, n], coverage int32 [S, n], VAF float64 [S, n] or None)"""
        n, S = self.n_rows, self.n_samples
        rows = np.empty(n, np.uint32)
        occ = np.empty((S, n), np.int32)
        cov = np.empty((S, n), np.int32)
        vf = np.empty((S, n), np.float64) if (vaf and self.has_vaf) else None
        ptrs = lambda a, ct: (C.POINTER(ct) * S)(*[a[s].ctypes.data_as(C.POINTER(ct)) for s in range(S)])
        b = C.c_uint64(0)
        _ok(lib().pcs_result_fetch(self._h, A.ptr(rows, C.c_uint32), ptrs(occ, C.c_int32), ptrs(cov, C.c_int32),
                                   ptrs(vf, C.c_double) if vf is not None else None, C.byref(b)))
        self.d2h_bytes = b.value
        return rows, occ, cov, vf

    def close(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.pcs_result_free(self._h)
            self._h = None

    __del__ = close


class Forest:
    """a forest flattened and resident in HBM."""

    def __init__(self, ctx: Context, forest):
        self.ctx = ctx
        self.forest = forest
        self._h = C.c_void_p()
        if hasattr(forest, "as_genomes_desc"):  # explicit per-cell genomes (process_b200.genomes.CellGenomes)
            d = forest.as_genomes_desc()
            _ok(lib().pcs_forest_upload_genomes(ctx._h, C.byref(d), C.byref(self._h)))
        else:
            d = forest.as_desc()
            _ok(lib().pcs_forest_upload(ctx._h, C.byref(d), C.byref(self._h)))
        self.n_groups = forest.n_samples

    def close(self):
        if getattr(self, "_h", None) and _LIB is not None:  # _LIB is gone at interpreter shutdown
            _LIB.pcs_forest_free(self._h)
            self._h = None

    __del__ = close

    def set_groups(self, leaf_group, n_groups):
        lg = _u32(leaf_group)
        _ok(lib().pcs_forest_set_groups(self._h, A.ptr(lg, C.c_uint32), C.c_uint32(n_groups or 0)))
        self.n_groups = n_groups if leaf_group is not None else self.forest.n_samples

    def info(self):
        out = (C.c_uint64 * 6)()
        _ok(lib().pcs_forest_info(self._h, out))
        return dict(n_loci=out[0], n_instances=out[1], n_haplotypes=out[2], n_fragment_sets=out[3],
                    n_pieces=out[4], device_bytes=out[5])

    def instances(self):
        """(inst [n_instances, 4], locus_inst_off [n_loci + 1]) as they lie on the device"""
        i = self.info()
        inst = np.zeros((i["n_instances"], 4), np.uint32)
        off = np.zeros(i["n_loci"] + 1, np.uint32)
        _ok(lib().pcs_forest_instances(self._h, A.ptr(inst, C.c_uint32), A.ptr(off, C.c_uint32)))
        return inst, off

    # ---- sequences (only needed to write SAM)
    def set_reference(self, chrom, bases: bytes):
        _ok(lib().pcs_forest_set_reference(self._h, C.c_uint32(chrom), C.c_char_p(bases), C.c_uint64(len(bases))))

    def load_fasta(self, path):
        names = (C.c_char_p * self.forest.n_chr)(*[n.encode() for n in self.forest.chr_names])
        n = C.c_uint32(0)
        _ok(lib().pcs_forest_load_fasta(self._h, C.c_char_p(os.fsencode(path)), names, C.byref(n)))
        return n.value

    def set_alt(self, alt_off, alt_bytes: bytes):
        ao = _u32(alt_off)
        _ok(lib().pcs_forest_set_alt(self._h, A.ptr(ao, C.c_uint32), C.c_char_p(alt_bytes)))

    def n_out_samples(self, params: A.SeqParams):
        if params.normal_only:
            return 1
        return self.n_groups + (1 if params.with_normal_sample else 0)

    def simulate(self, params: A.SeqParams):
        """plan + run + free with host outputs (the call the Rcpp shim makes)."""
        n_out = self.n_out_samples(params)
        occ = np.zeros((n_out, self.forest.n_mut), np.uint32)
        cov = np.zeros((n_out, self.forest.n_mut), np.uint32)
        st = A.RunStats()
        _ok(lib().pcs_simulate(self._h, C.byref(params), A.ptr(occ, C.c_uint32), A.ptr(cov, C.c_uint32), C.byref(st)))
        return occ, cov, st

    def simulate_result(self, params: A.SeqParams, include_non_sequenced=False, with_vaf=True):
        """plan + sample + assemble the data frame's rows on the device; returns (Result, stats)."""
        h = C.c_void_p()
        st = A.RunStats()
        _ok(lib().pcs_simulate_result(self._h, C.byref(params), C.c_int(1 if include_non_sequenced else 0),
                                      C.c_int(1 if with_vaf else 0), C.byref(h), C.byref(st)))
        return Result(h), st

    def count_injected(self, n_out, read_size, placements, err_masks=None):
        placements = np.ascontiguousarray(placements, dtype=A.PLACEMENT_DTYPE)
        occ = np.zeros((n_out, self.forest.n_mut), np.uint32)
        cov = np.zeros((n_out, self.forest.n_mut), np.uint32)
        em = _u32(err_masks)
        st = A.RunStats()
        _ok(lib().pcs_count_injected(self._h, C.c_uint32(n_out), C.c_uint32(read_size),
                                     C.c_void_p(placements.ctypes.data), A.ptr(em, C.c_uint32),
                                     C.c_uint64(len(placements)), A.ptr(occ, C.c_uint32), A.ptr(cov, C.c_uint32),
                                     C.byref(st)))
        return occ, cov, st

    def active_rows(self, occ, include_non_sequenced=False, params=None):
        occ = np.ascontiguousarray(occ, dtype=np.uint32)
        rows = np.zeros(max(1, self.forest.n_mut), np.uint32)
        n = C.c_uint32(0)
        _ok(lib().pcs_active_rows(self._h, A.ptr(occ, C.c_uint32), C.c_uint32(occ.shape[0]),
                                  C.c_int(1 if include_non_sequenced else 0),
                                  C.byref(params) if params is not None else None, A.ptr(rows, C.c_uint32), C.byref(n)))
        return rows[:n.value].copy()


def replicate(src: "Forest", ctx: Context) -> "Forest":
    """a copy of an uploaded forest on another context's device, sharing the flattened host view."""
    f = Forest.__new__(Forest)
    f.ctx, f.forest, f.n_groups = ctx, src.forest, src.n_groups
    f._h = C.c_void_p()
    _ok(lib().pcs_forest_replicate(src._h, ctx._h, C.byref(f._h)))
    return f


def simulate_multi(forests, params: A.SeqParams):
    """one process driving len(forests) devices: shard i runs on forests[i], tables land on forests[0]'s GPU."""
    n_out = forests[0].n_out_samples(params)
    n_mut = forests[0].forest.n_mut
    occ = np.zeros((n_out, n_mut), np.uint32)
    cov = np.zeros((n_out, n_mut), np.uint32)
    arr = (C.c_void_p * len(forests))(*[f._h for f in forests])
    st = A.RunStats()
    _ok(lib().pcs_simulate_multi(arr, C.c_uint32(len(forests)), C.byref(params), A.ptr(occ, C.c_uint32),
                                 A.ptr(cov, C.c_uint32), C.byref(st)))
    return occ, cov, st


class Plan:
    """tile grid + sampling tables of one simulate call, resident in HBM."""

    def __init__(self, forest: Forest, params: A.SeqParams):
        self.forest = forest
        self.params = params
        self._h = C.c_void_p()
        _ok(lib().pcs_plan_create(forest._h, C.byref(params), C.byref(self._h)))
        self.info = A.PlanInfo()
        _ok(lib().pcs_plan_info_get(self._h, C.byref(self.info)))

    def close(self):
        if getattr(self, "_h", None) and _LIB is not None:  # _LIB is gone at interpreter shutdown
            _LIB.pcs_plan_free(self._h)
            self._h = None

    __del__ = close

    def run(self):
        n_out, n_mut = self.info.n_out_samples, self.info.n_mut
        occ = np.zeros((n_out, n_mut), np.uint32)
        cov = np.zeros((n_out, n_mut), np.uint32)
        st = A.RunStats()
        _ok(lib().pcs_plan_run(self._h, C.c_int(A.PCS_RUN_HOST_OUTPUT), A.ptr(occ, C.c_uint32),
                               A.ptr(cov, C.c_uint32), C.byref(st)))
        return occ, cov, st

    def result(self, include_non_sequenced=False, with_vaf=True):
        """assemble the tables the last run() left on the device (pcs_plan_result)."""
        h = C.c_void_p()
        _ok(lib().pcs_plan_result(self._h, C.c_int(1 if include_non_sequenced else 0), C.byref(self.params),
                                  C.c_int(1 if with_vaf else 0), C.byref(h)))
        return Result(h)

    def run_device(self, occ_ptr: int, cov_ptr: int, checksums=True, wait=True):
        """occ_ptr / cov_ptr: device addresses of uint32 [n_out_samples, n_mut] buffers.  wait=False: the call
        returns once the kernels are queued (no stats; counters() reads what accumulated)."""
        flags = A.PCS_RUN_DEVICE_OUTPUT | (0 if checksums else A.PCS_RUN_NO_CHECKSUMS) | (0 if wait else A.PCS_RUN_ASYNC)
        st = A.RunStats()
        _ok(lib().pcs_plan_run(self._h, C.c_int(flags), C.c_void_p(occ_ptr), C.c_void_p(cov_ptr), C.byref(st)))
        return st if wait else None

    def accumulate(self, depth_ptr: int, occ_ptr: int, wait=True):
        """this shard's sampler adding into depth [S, n_loci] / occ [S, n_mut] (device or peer pointers)."""
        st = A.RunStats()
        _ok(lib().pcs_plan_accumulate(self._h, C.c_void_p(depth_ptr), C.c_void_p(occ_ptr), C.byref(st) if wait else None))
        return st if wait else None

    def finalize(self, depth_ptr: int, occ_ptr: int, cov_ptr: int, wait=True):
        st = A.RunStats()
        _ok(lib().pcs_plan_finalize(self._h, C.c_void_p(depth_ptr), C.c_void_p(occ_ptr), C.c_void_p(cov_ptr),
                                    C.byref(st) if wait else None))
        return st if wait else None

    def finalize_on(self, depth_ptr: int, cov_ptr: int, stream: int):
        """the coverage gather queued on `stream` (a cudaStream_t handle)"""
        _ok(lib().pcs_plan_finalize_stream(self._h, C.c_void_p(depth_ptr), C.c_void_p(cov_ptr), C.c_void_p(stream)))

    def counters(self):
        """wait for the queued work; n_reads / checksums accumulated by the wait=False calls since the last read"""
        st = A.RunStats()
        _ok(lib().pcs_plan_counters(self._h, C.byref(st)))
        return st

    def materialize(self, cap):
        """the plan's reads as binary records: placements, error masks, seq, qual, cigar, n_cigar, lengths."""
        R = self.info.read_size
        rec = np.zeros(cap, A.PLACEMENT_DTYPE)
        masks = np.zeros((cap, A.PCS_ERRMASK_WORDS), np.uint32)
        seq = np.zeros((cap, R), np.uint8); qual = np.zeros((cap, R), np.uint8)
        cigar = np.zeros((cap, A.PCS_MAX_CIGAR), np.uint32)
        nc = np.zeros(cap, np.uint32); ln = np.zeros(cap, np.uint32)
        n = C.c_uint64(0)
        _ok(lib().pcs_plan_materialize(self._h, C.c_uint64(cap), C.c_void_p(rec.ctypes.data), A.ptr(masks, C.c_uint32),
                                       A.ptr(seq, C.c_uint8), A.ptr(qual, C.c_uint8), A.ptr(cigar, C.c_uint32),
                                       A.ptr(nc, C.c_uint32), A.ptr(ln, C.c_uint32), C.byref(n)))
        k = n.value
        return rec[:k], masks[:k], seq[:k], qual[:k], cigar[:k], nc[:k], ln[:k]

    def write_sam(self, output_dir, sample_names, filename_prefix="chr_", template_name_prefix="r", update=False):
        f = self.forest.forest
        chr_names = (C.c_char_p * f.n_chr)(*[n.encode() for n in f.chr_names])
        samples = (C.c_char_p * len(sample_names))(*[n.encode() for n in sample_names])
        opt = A.SamOptions(os.fsencode(output_dir), filename_prefix.encode(), template_name_prefix.encode(),
                           chr_names, samples, 1 if update else 0)
        n = C.c_uint64(0)
        _ok(lib().pcs_plan_write_sam(self._h, C.byref(opt), C.byref(n)))
        return n.value

    def coverage_track(self, bin_bp=4096):
        """binned depth of the plan's reads: (chr_bin_off uint64 [n_chr+1], track uint32 [n_out_samples, n_bins]);
        track[s, chr_bin_off[c] + pos // bin_bp] = reference bases the reads of sample s lay on that bin"""
        n_chr = self.forest.forest.n_chr
        off = np.zeros(n_chr + 1, np.uint64)
        n = C.c_uint64(0)
        _ok(lib().pcs_plan_coverage_track(self._h, C.c_uint32(bin_bp), A.ptr(off, C.c_uint64), None, C.c_uint64(0), C.byref(n)))
        track = np.zeros((self.info.n_out_samples, n.value), np.uint32)
        _ok(lib().pcs_plan_coverage_track(self._h, C.c_uint32(bin_bp), A.ptr(off, C.c_uint64), A.ptr(track, C.c_uint32),
                                          C.c_uint64(n.value), C.byref(n)))
        return off, track

    def trace(self, cap, with_masks=False):
        rec = np.zeros(cap, A.PLACEMENT_DTYPE)
        masks = np.zeros((cap, A.PCS_ERRMASK_WORDS), np.uint32) if with_masks else None
        n = C.c_uint64(0)
        _ok(lib().pcs_plan_trace(self._h, C.c_void_p(rec.ctypes.data), A.ptr(masks, C.c_uint32),
                                 C.c_uint64(cap), C.byref(n)))
        return rec[:n.value], (None if masks is None else masks[:n.value])
