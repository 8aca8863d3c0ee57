"""ctypes declarations mirroring include/pcs_seq.h (declarations only; nothing is loaded here)."""
import ctypes as C

import numpy as np

PCS_ABI_VERSION = 1

PCS_EV_SID, PCS_EV_CNA_AMP, PCS_EV_CNA_DEL, PCS_EV_WGD = 0, 1, 2, 3
PCS_NATURE_DRIVER, PCS_NATURE_PASSENGER, PCS_NATURE_GERMINAL, PCS_NATURE_PRENEOPLASTIC = 0, 1, 2, 3
NATURE_DESCRIPTIONS = ("driver", "passenger", "germinal", "preneoplastic")
PCS_SEQ_ERRORLESS, PCS_SEQ_BASIC_CONSTANT, PCS_SEQ_BASIC_RANDOM = 0, 1, 2
PCS_PLACE_TUMOUR, PCS_PLACE_NORMAL_PLAIN, PCS_PLACE_NORMAL_PRENEO = 0, 1, 2
PCS_ERRMASK_WORDS = 8
PCS_RUN_HOST_OUTPUT, PCS_RUN_DEVICE_OUTPUT, PCS_RUN_NO_CHECKSUMS, PCS_RUN_ASYNC = 0, 1, 2, 4

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)


class ForestDesc(C.Structure):
    _fields_ = [
        ("n_chr", C.c_uint32), ("chr_len", _u32p), ("chr_n_alleles", _u8p),
        ("n_nodes", C.c_uint32), ("node_parent", _i32p),
        ("n_samples", C.c_uint32), ("n_leaves", C.c_uint32), ("leaf_node", _u32p), ("leaf_sample", _u32p),
        ("n_events", C.c_uint64), ("node_event_off", _u64p),
        ("ev_kind", _u8p), ("ev_chr", _u16p), ("ev_pos", _u32p), ("ev_len", _u32p),
        ("ev_allele", _u16p), ("ev_dest", _u16p), ("ev_mut", _u32p), ("ev_nature", _u8p),
        ("n_mut", C.c_uint32), ("mut_chr", _u16p), ("mut_pos", _u32p),
        ("mut_ref_len", _u8p), ("mut_alt_len", _u8p),
        ("n_germline", C.c_uint64), ("germ_mut", _u32p), ("germ_allele_mask", _u8p),
    ]


class CellGenomesDesc(C.Structure):
    _fields_ = [
        ("n_chr", C.c_uint32), ("chr_len", _u32p), ("chr_n_alleles", _u8p),
        ("n_samples", C.c_uint32), ("n_cells", C.c_uint32), ("cell_sample", _u32p), ("n_normal_preneo", C.c_uint32),
        ("n_alleles", C.c_uint64), ("allele_cell", _u32p), ("allele_chr", _u16p), ("allele_id", _u16p),
        ("allele_origin", _u8p), ("allele_frag_off", _u64p), ("frag_begin", _u32p), ("frag_end", _u32p),
        ("allele_sid_off", _u64p), ("sid_row", _u32p),
        ("n_mut", C.c_uint32), ("mut_chr", _u16p), ("mut_pos", _u32p), ("mut_ref_len", _u8p), ("mut_alt_len", _u8p),
        ("n_germline", C.c_uint64), ("germ_mut", _u32p), ("germ_allele_mask", _u8p),
    ]


class SeqParams(C.Structure):
    _fields_ = [
        ("seed", C.c_int32), ("coverage", C.c_double), ("purity", C.c_double),
        ("read_size", C.c_uint32), ("insert_size_mean", C.c_uint32), ("insert_size_stddev", C.c_uint32),
        ("sequencer", C.c_uint32), ("error_rate", C.c_double),
        ("with_normal_sample", C.c_uint8), ("preneoplastic_in_normal", C.c_uint8),
        ("normal_only", C.c_uint8), ("reserved0", C.c_uint8),
        ("chr_mask", _u8p), ("shard_rank", C.c_uint32), ("shard_count", C.c_uint32),
    ]


class ReadPlacement(C.Structure):
    _fields_ = [("cell", C.c_uint32), ("start", C.c_uint32), ("chr", C.c_uint16),
                ("allele", C.c_uint16), ("sample", C.c_uint16), ("flags", C.c_uint16)]


PLACEMENT_DTYPE = np.dtype([("cell", "<u4"), ("start", "<u4"), ("chr", "<u2"),
                            ("allele", "<u2"), ("sample", "<u2"), ("flags", "<u2")])
assert PLACEMENT_DTYPE.itemsize == C.sizeof(ReadPlacement) == 16


class PlanInfo(C.Structure):
    _fields_ = [("n_out_samples", C.c_uint32), ("n_mut", C.c_uint32), ("n_loci", C.c_uint32),
                ("n_tiles", C.c_uint64), ("n_tiles_total", C.c_uint64),
                ("n_templates", C.c_uint64), ("n_templates_total", C.c_uint64),
                ("reads_per_template", C.c_uint32), ("read_size", C.c_uint32)]


class RunStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_ms", C.c_double), ("kernel_launches", C.c_uint64),
                ("n_templates", C.c_uint64), ("n_reads", C.c_uint64), ("sum_depth", C.c_uint64),
                ("sum_occurrences", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def ptr(arr, ctype):
    """pointer to a C-contiguous numpy array (None -> NULL)."""
    if arr is None:
        return C.cast(None, C.POINTER(ctype))
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(C.POINTER(ctype))


class SamOptions(C.Structure):
    _fields_ = [("output_dir", C.c_char_p), ("filename_prefix", C.c_char_p), ("template_name_prefix", C.c_char_p),
                ("chr_names", C.POINTER(C.c_char_p)), ("sample_names", C.POINTER(C.c_char_p)), ("update", C.c_uint8)]


PCS_MAX_CIGAR = 16
