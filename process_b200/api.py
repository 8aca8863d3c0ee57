"""Host-side mirror of the reference's R/Rcpp interface for the sequencing path.

Same names, argument order, defaults and error behaviour as
  simulate_seq()          src/sequencing.cpp:193-209, src/seq_simulation.cpp:517-601
  simulate_normal_seq()   src/sequencing.cpp:268-282, src/seq_simulation.cpp:603-679
  BasicIlluminaSequencer / ErrorlessIlluminaSequencer
                          src/sequencers.hpp:25-72, src/sequencers.cpp:80-106
(R is not available in this image; INTEGRATION.md shows the Rcpp shim that makes
the same C-ABI calls.)  Everything below the argument handling runs on the GPU
through libpcs_seq.so; there is no CPU path.
"""
from __future__ import annotations

import os
import atexit
import weakref
from dataclasses import dataclass

import numpy as np
import pandas as pd

from . import _abi as A
from . import _lib as L
from .forest import PhylogeneticForest

NORMAL_SAMPLE_NAME = "normal_sample"  # src/seq_simulation.cpp:572, 651


# ------------------------------------------------------------------ sequencers
class ErrorlessIlluminaSequencer:
    """src/sequencers.hpp:25-38"""

    @property
    def error_rate(self):
        return 0.0

    def __repr__(self):
        return 'Errorless Illumina (platform: "ILLUMINA")'


class BasicIlluminaSequencer:
    """src/sequencers.hpp:40-72; validation as build_sequencer(), src/sequencers.cpp:80-106"""

    def __init__(self, error_rate, random_quality_scores=True):
        if isinstance(error_rate, bool) or not isinstance(error_rate, (int, float, np.integer, np.floating)):
            raise ValueError('The parameter "error_rate" must be a positive real number.')
        if error_rate < 0:
            raise ValueError('The parameter "error_rate" must be a positive real number.')
        if not isinstance(random_quality_scores, (bool, np.bool_)):
            raise ValueError('The parameter "random_quality_scores" must be a Boolean value.')
        self.error_rate = float(error_rate)
        self.random_quality_scores = bool(random_quality_scores)

    def __repr__(self):
        kind = "random quality scores" if self.random_quality_scores else "constant quality scores"
        return f'Basic Illumina (platform: "ILLUMINA" error rate: {self.error_rate:f} {kind})'


def _sequencer_model(sequencer):
    """dispatch of src/seq_simulation.cpp:375-429"""
    if sequencer is None or isinstance(sequencer, ErrorlessIlluminaSequencer):
        return A.PCS_SEQ_ERRORLESS, 0.0
    if isinstance(sequencer, BasicIlluminaSequencer):
        kind = A.PCS_SEQ_BASIC_RANDOM if sequencer.random_quality_scores else A.PCS_SEQ_BASIC_CONSTANT
        return kind, sequencer.error_rate
    raise ValueError("Unsupported sequencer type")


def _sequencer_data(sequencer):
    """get_sequencer_data(), src/seq_simulation.cpp:453-515"""
    if sequencer is None:
        return None
    if isinstance(sequencer, BasicIlluminaSequencer):
        return dict(name="BasicIlluminaSequencer", error_rate=sequencer.error_rate,
                    random_quality_scores=sequencer.random_quality_scores)
    return dict(name="ErrorlessIlluminaSequencer", error_rate=0.0)


# ------------------------------------------------------------------ cell labelling
@dataclass
class SampledCell:
    """const view of a sampled cell for the labelling callback (src/sampled_cell.hpp:27-56)"""
    cell_id: int
    sample: str
    epistate: str = ""
    mutant: str = ""
    species: str = ""
    birth_time: float = 0.0


class VectorisedLabelling:
    """A cell_labelling evaluated ONCE over all sampled cells instead of once per cell
    (the reference calls an R closure per cell, src/seq_simulation.cpp:196-201; at 1e5 sampled
    cells that loop is the host bottleneck -- SURVEY.md 8 f4).  `fn` receives a dict of
    equal-length columns (cell_id, sample and the forest's per-leaf attributes: epistate,
    mutant, species, birth_time when present) and returns one string label per cell.
    Same sample names, same order of first appearance as the per-cell form."""

    def __init__(self, fn):
        if not callable(fn):
            raise ValueError("The FACs_labelling_function must be a function.")
        self.fn = fn


def _vectorised_labels(forest: PhylogeneticForest, labelling: VectorisedLabelling):
    attrs = getattr(forest, "leaf_attrs", None) or {}
    sample_names = np.asarray(forest.sample_names, dtype=object)
    cols = {"cell_id": np.asarray(forest.leaf_node).astype(np.int64), "sample": sample_names[forest.leaf_sample]}
    cols.update({k: np.asarray(v) for k, v in attrs.items()})
    labels = np.asarray(labelling.fn(cols), dtype=object)
    if labels.shape != (forest.n_leaves,):
        raise ValueError("The vectorised labelling function must return one label per sampled cell.")
    if not all(isinstance(x, str) for x in labels):
        raise ValueError("The labelling function must return a string.")
    # order of first appearance when cells are visited sample by sample, leaves in order within a sample
    visit = np.argsort(forest.leaf_sample, kind="stable")
    key = np.char.add(np.char.add(forest.leaf_sample[visit].astype(str), "\x1f"), labels[visit].astype(str))
    code, uniq = pd.factorize(key)
    group = np.zeros(forest.n_leaves, np.uint32)
    group[visit] = code.astype(np.uint32)
    names = []
    for u in uniq:
        s, label = u.split("\x1f", 1)
        names.append(forest.sample_names[int(s)] + ("_" + label if label != "" else ""))
    return group, names


def _apply_FACS_labels(forest: PhylogeneticForest, labelling):
    """apply_FACS_labels()/split_by_labels(), src/seq_simulation.cpp:183-243: one
    callback per sampled cell; cells of sample S labelled L go to sample "S_L"
    (or "S" for the empty label), new samples in order of first appearance."""
    if labelling is None:
        return None, list(forest.sample_names)
    if isinstance(labelling, VectorisedLabelling):
        return _vectorised_labels(forest, labelling)
    if not callable(labelling):
        raise ValueError("The FACs_labelling_function must be a function.")
    attrs = getattr(forest, "leaf_attrs", None) or {}
    names, index = [], {}
    group = np.zeros(forest.n_leaves, np.uint32)
    for s in range(forest.n_samples):
        for l in np.flatnonzero(forest.leaf_sample == s):
            cell = SampledCell(cell_id=int(forest.leaf_node[l]), sample=forest.sample_names[s],
                               **{k: (v[l].item() if hasattr(v[l], "item") else v[l]) for k, v in attrs.items()})
            label = labelling(cell)
            if not isinstance(label, str):
                raise ValueError("The labelling function must return a string.")
            name = forest.sample_names[s] + ("_" + label if label != "" else "")
            key = (s, label)
            if key not in index:
                index[key] = len(names)
                names.append(name)
            group[l] = index[key]
    return group, names


# ------------------------------------------------------------------ argument handling
def _reference_genome(forest, reference_genome):
    """get_reference_genome(), src/seq_simulation.cpp:245-280"""
    if reference_genome is None:
        path = forest.reference_path
        if path is None or not os.path.exists(path):
            raise RuntimeError(f'The reference genome file "{path}" does not exists anymore. Please, re-build '
                               'the mutation engine or use the parameter "reference_genome".')
        return path
    if isinstance(reference_genome, str):
        if not os.path.exists(reference_genome):
            raise RuntimeError(f'The reference genome file "{reference_genome}" does not exists.')
        return reference_genome
    raise ValueError('The parameter "reference_genome" must be either NULL or a string.')


def _ordinal(i):
    suf = "th" if 10 <= i % 100 <= 20 else {1: "st", 2: "nd", 3: "rd"}.get(i % 10, "th")
    return f"{i}{suf}"


def _chr_mask(forest, chromosomes):
    """get_relevant_chr_set(), src/seq_simulation.cpp:299-352"""
    if chromosomes is None:
        return None
    if isinstance(chromosomes, str):
        chromosomes = [chromosomes]
    if not isinstance(chromosomes, (list, tuple, np.ndarray)):
        raise ValueError("Unsupported chromosome list type")
    mask = np.zeros(forest.n_chr, np.uint8)
    for i, name in enumerate(chromosomes, 1):
        if not isinstance(name, (str, np.str_)):
            raise ValueError(f"Expected a list of string: the {_ordinal(i)} element of the list is not a string.")
        if name not in forest.chr_names:
            raise ValueError(f'Unknown chromosome "{name}"')
        mask[forest.chr_names.index(name)] = 1
    return mask


def _resolve_seed(seed):
    """get_random_seed<int>(), src/utility.hpp:41-64: a number is cast; NULL draws runif(INT_MIN, INT_MAX) from the
    session's RNG state, so that set.seed() governs the run -- here NumPy's global state (np.random.seed())."""
    if seed is None:
        return int(np.random.uniform(-2.0**31, 2.0**31 - 1))
    if isinstance(seed, bool) or not isinstance(seed, (int, float, np.integer, np.floating)):
        raise ValueError("The seed must be either a number or NILL.")
    return int(seed)


def _agree_over_ranks(c_seed, group, names):
    """More than one rank (torch.distributed initialised): every rank must plan THE SAME call, or their tile sets
    are not a partition of the job -- the per-tile template counts and the shard assignment both follow the seed.
    seed=NULL is resolved on rank 0 and broadcast (each rank would otherwise draw its own from its own RNG state);
    the sample partition (names and the cells of every group) must be identical, else the call is refused."""
    rank, world = _shard()
    if world == 1:
        return c_seed
    import hashlib
    import torch.distributed as dist
    box = [c_seed]
    dist.broadcast_object_list(box, src=0)
    digest = hashlib.sha256(repr(list(names)).encode() + (b"" if group is None else np.ascontiguousarray(group).tobytes())).hexdigest()
    seen = [None] * world
    dist.all_gather_object(seen, digest)
    if len(set(seen)) != 1:
        raise ValueError("simulate_seq over several ranks: the ranks' sample partitions (cell_labelling) differ")
    return int(box[0])


_ctx_cache = {}
_forest_cache = weakref.WeakKeyDictionary()


def _device_forest(forest: PhylogeneticForest, device: int, cache: bool):
    ctx = _ctx_cache.get(device)
    if ctx is None:
        ctx = _ctx_cache[device] = L.Context(device)
    if not cache:
        return L.Forest(ctx, forest), True
    key = (device,)
    slot = _forest_cache.setdefault(forest, {})
    if key not in slot:
        slot[key] = L.Forest(ctx, forest)
    return slot[key], False


def release_device_cache():
    for slot in list(_forest_cache.values()):
        for f in slot.values():
            f.close()
    _forest_cache.clear()
    for c in _ctx_cache.values():
        c.close()
    _ctx_cache.clear()


# forests before their contexts, also when the interpreter exits with the cache still populated
# (include/pcs_seq.h: lifetime rule of pcs_destroy)
atexit.register(release_device_cache)


def _string_array(arrow_array):
    """a pyarrow large_string array as the column pandas would infer from Python strings, or None when this pandas
    does not keep strings in Arrow memory"""
    import pandas as pd
    dtype = pd.Series(["x"]).dtype  # str (pyarrow-backed) from pandas 3 on
    if getattr(dtype, "storage", None) != "pyarrow":
        return None
    return dtype.construct_array_type()(arrow_array, dtype=dtype)


def _string_column(rows, codes, table):
    """The column table[codes[rows]] (None in the table = NA) as pandas would infer it from Python strings, but built
    from the dictionary encoding: millions of rows draw from a handful of distinct strings (chromosome names,
    bases, signature names, class names), so no Python object is made per row -- the library's host threads write
    the Arrow buffers (offsets + bytes + validity) in one parallel pass and pandas wraps them without a copy
    (the data frame, not the sampler, is what a 5 M row result waits for)."""
    table = list(table)
    plain = lambda: np.asarray(table + [None], dtype=object)[:-1][np.asarray(codes)[np.asarray(rows, dtype=np.int64)]]
    if len(rows) == 0:
        return plain()
    try:
        import pyarrow as pa
        offsets, data, validity, nulls = L.host_string_column(rows, codes, table)
        if nulls == len(rows):
            return plain()  # nothing to infer a string type from: the object column pandas makes of it
        arr = pa.LargeStringArray.from_buffers(len(rows), pa.py_buffer(offsets), pa.py_buffer(data),
                                               pa.py_buffer(validity) if validity is not None else None, int(nulls))
        col = _string_array(arr)
        return plain() if col is None else col
    except ImportError:  # no pyarrow: plain object strings, the same values
        return plain()


_CLASS_TABLE = [";".join(sorted(A.NATURE_DESCRIPTIONS[b] for b in range(4) if (m >> b) & 1)) for m in range(16)]


def _row_codes(forest):
    """dictionary codes of the annotation columns for every row of the forest's mutation table, made once per forest
    (src/seq_simulation.cpp:52-90 builds these strings row by row on every call)"""
    cache = getattr(forest, "_row_code_cache", None)
    if cache is None or cache["n_mut"] != forest.n_mut:
        ref_codes, ref_table, alt_codes, alt_table = forest.row_string_codes(np.arange(forest.n_mut))
        cause = np.where(forest.mut_cause < 0, len(forest.cause_names), forest.mut_cause)
        cache = dict(n_mut=forest.n_mut,
                     chr=(forest.mut_chr.astype(np.uint16), list(forest.chr_names)),
                     ref=(ref_codes.astype(np.uint16), list(ref_table)), alt=(alt_codes.astype(np.uint16), list(alt_table)),
                     causes=(cause.astype(np.uint16), list(forest.cause_names) + [None]),
                     classes=((forest.mut_nature_mask & 15).astype(np.uint16), _CLASS_TABLE))
        forest._row_code_cache = cache
    return cache


def _frame(forest, names, rows, occ_cols, cov_cols, vaf_cols):
    """get_result_dataframe()/add_sample_statistics(), src/seq_simulation.cpp:52-181, from the compact columns the
    device assembled: `rows` (ascending = SID order) and per sample occurrences / coverage / VAF of those rows.
    Sample columns in name order (std::map iteration)."""
    import pandas as pd
    codes = _row_codes(forest)
    cols = {"chr": _string_column(rows, *codes["chr"]),
            "chr_pos": L.host_gather(rows, forest.mut_pos).view(np.int32),
            "ref": _string_column(rows, *codes["ref"]), "alt": _string_column(rows, *codes["alt"]),
            "causes": _string_column(rows, *codes["causes"]), "classes": _string_column(rows, *codes["classes"])}
    for s in sorted(range(len(names)), key=lambda i: names[i]):
        cols[f"{names[s]}.occurrences"] = occ_cols[s]
        cols[f"{names[s]}.coverage"] = cov_cols[s]
        cols[f"{names[s]}.VAF"] = vaf_cols[s]
    return pd.DataFrame(cols, copy=False)


def _frame_from_tables(forest, dev, occ, cov, names, include_non_sequenced, params=None):
    """the same frame from full host tables (several ranks: the tables are summed over ranks first): active rows by
    pcs_active_rows, columns gathered by the library's host threads"""
    rows = dev.active_rows(occ, include_non_sequenced, params)
    occ_cols = [L.host_gather(rows, occ[s]).view(np.int32) for s in range(len(names))]
    cov_cols = [L.host_gather(rows, cov[s]).view(np.int32) for s in range(len(names))]
    # VAF = occurrences / coverage; a row the sample never covered has no occurrence: VAF 0, as the rows the
    # reference does not find in the sample's data (src/seq_simulation.cpp:121-131)
    vaf_cols = [np.divide(o, c, out=np.zeros(len(o), np.float64), where=c != 0) for o, c in zip(occ_cols, cov_cols)]
    return _frame(forest, names, rows, occ_cols, cov_cols, vaf_cols)


def _run(forest, sequencer, reference_genome, chromosomes, coverage, read_size, insert_size_mean,
         insert_size_stddev, write_SAM, group, names, purity, with_normal_sample, preneo, normal_only,
         include_non_sequenced, c_seed, device, cache, shard, sam=None):
    if not isinstance(forest, PhylogeneticForest):
        raise TypeError("phylo_forest must be a PhylogeneticForest")
    ref_path = _reference_genome(forest, reference_genome)
    kind, rate = _sequencer_model(sequencer)
    mask = _chr_mask(forest, chromosomes)
    if not normal_only and not (0 <= purity <= 1):
        raise ValueError("The purity must belong to the interval [0,1].")
    if write_SAM and shard[1] > 1:
        raise NotImplementedError("write_SAM with more than one rank: every rank would write the same files")
    P = A.SeqParams(seed=c_seed, coverage=float(coverage), purity=float(purity), read_size=int(read_size),
                    insert_size_mean=int(insert_size_mean), insert_size_stddev=int(insert_size_stddev),
                    sequencer=kind, error_rate=rate, with_normal_sample=int(bool(with_normal_sample)),
                    preneoplastic_in_normal=int(bool(preneo)), normal_only=int(bool(normal_only)),
                    shard_rank=shard[0], shard_count=shard[1])
    if mask is not None:
        import ctypes as C
        P.chr_mask = A.ptr(mask, C.c_uint8)
    dev, owned = _device_forest(forest, device, cache)
    try:
        if not normal_only:
            dev.set_groups(group, len(names) if group is not None else None)
        out_names = [NORMAL_SAMPLE_NAME] if normal_only else list(names) + ([NORMAL_SAMPLE_NAME] if with_normal_sample else [])
        if write_SAM:
            # the counting kernels and the SAM records of one plan draw the same reads (same Philox counters)
            output_dir, filename_prefix, template_name_prefix, update_SAM = sam
            if update_SAM is False and os.path.exists(output_dir):
                raise ValueError(f'The output directory "{output_dir}" already exists: use update_SAM=TRUE '
                                 "to add files to it.")
            if getattr(dev, "_fasta", None) != ref_path:
                dev.load_fasta(ref_path)
                dev._fasta = ref_path
            if not getattr(dev, "_alt_set", False):
                dev.set_alt(*forest.alt_table())
                dev._alt_set = True
            plan = L.Plan(dev, P)
            try:
                occ, cov, st = plan.run()
                plan.write_sam(output_dir, out_names, filename_prefix, template_name_prefix, update_SAM)
                res = plan.result(include_non_sequenced)
            finally:
                plan.close()
        elif shard[1] > 1:
            occ, cov, st = dev.simulate(P)
            occ, cov = _reduce_over_ranks(occ, cov)
            res = None
        else:
            res, st = dev.simulate_result(P, include_non_sequenced)
        if res is None:
            df = _frame_from_tables(forest, dev, occ, cov, out_names, include_non_sequenced, P)
        else:
            try:  # the data frame's rows were assembled on the device: only they cross the link
                rows, occ_c, cov_c, vaf_c = res.fetch()
            finally:
                res.close()
            df = _frame(forest, out_names, rows, occ_c, cov_c, vaf_c)
    finally:
        if owned:
            dev.close()
    return df, st


def _reduce_over_ranks(occ, cov):
    """sum the per-rank count tables (torch.distributed: NCCL on GPU, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    out = []
    for a in (occ, cov):
        t = torch.from_numpy(a.astype(np.int32)).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out.append(t.cpu().numpy().astype(np.uint32))
    return out


def _shard():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def simulate_seq(phylo_forest, sequencer=None, reference_genome=None, chromosomes=None, coverage=10,
                 read_size=150, insert_size_mean=0, insert_size_stddev=10, output_dir="ProCESS_SAM",
                 write_SAM=False, update_SAM=False, cell_labelling=None, purity=1, with_normal_sample=True,
                 preneoplastic_in_normal=False, filename_prefix="chr_", template_name_prefix="r",
                 include_non_sequenced_mutations=False, seed=None, *, device=0, cache=True):
    """Simulate the sequencing of the samples in a phylogenetic forest.

    Returns {"mutations": DataFrame, "parameters": dict} with the reference's schema
    (src/seq_simulation.cpp:84-89, 137-139, 586-600)."""
    c_seed = _resolve_seed(seed)
    group, names = _apply_FACS_labels(phylo_forest, cell_labelling)
    c_seed = _agree_over_ranks(c_seed, group, names)
    df, st = _run(phylo_forest, sequencer, reference_genome, chromosomes, coverage, read_size, insert_size_mean,
                  insert_size_stddev, write_SAM, group, names, purity, with_normal_sample,
                  preneoplastic_in_normal, False, include_non_sequenced_mutations, c_seed, device, cache, _shard(),
                  sam=(output_dir, filename_prefix, template_name_prefix, update_SAM))
    parameters = dict(sequencer=_sequencer_data(sequencer), reference_genome=reference_genome,
                      chromosomes=chromosomes, coverage=coverage, read_size=read_size,
                      insert_size_mean=insert_size_mean, insert_size_stddev=insert_size_stddev,
                      output_dir=output_dir, write_SAM=write_SAM, update_SAM=update_SAM,
                      cell_labelling=cell_labelling, purity=purity, with_normal_sample=with_normal_sample,
                      filename_prefix=filename_prefix, template_name_prefix=template_name_prefix,
                      include_non_sequenced_mutations=include_non_sequenced_mutations, seed=c_seed)
    return {"mutations": df, "parameters": parameters, "_stats": st.as_dict()}


def simulate_normal_seq(phylo_forest, sequencer=None, reference_genome=None, chromosomes=None, coverage=10,
                        read_size=150, insert_size_mean=0, insert_size_stddev=10,
                        output_dir="ProCESS_normal_SAM", write_SAM=True, update_SAM=False,
                        with_preneoplastic=False, filename_prefix="chr_", template_name_prefix="r",
                        include_non_sequenced_mutations=False, seed=None, *, device=0, cache=True):
    """Simulate the sequencing of a normal sample (purity forced to 1,
    src/seq_simulation.cpp:650-657)."""
    c_seed = _agree_over_ranks(_resolve_seed(seed), None, [])
    df, st = _run(phylo_forest, sequencer, reference_genome, chromosomes, coverage, read_size, insert_size_mean,
                  insert_size_stddev, write_SAM, None, [], 1.0, False, with_preneoplastic, True,
                  include_non_sequenced_mutations, c_seed, device, cache, _shard(),
                  sam=(output_dir, filename_prefix, template_name_prefix, update_SAM))
    parameters = dict(sequencer=_sequencer_data(sequencer), reference_genome=reference_genome,
                      chromosomes=chromosomes, coverage=coverage, read_size=read_size,
                      insert_size_mean=insert_size_mean, insert_size_stddev=insert_size_stddev,
                      output_dir=output_dir, write_SAM=write_SAM, update_SAM=update_SAM,
                      with_preneoplastic=with_preneoplastic, filename_prefix=filename_prefix,
                      template_name_prefix=template_name_prefix,
                      include_non_sequenced_mutations=include_non_sequenced_mutations, seed=c_seed)
    return {"mutations": df, "parameters": parameters, "_stats": st.as_dict()}


# --------------------------------------------------------------------------- downstream data formats
# The R helpers that take simulate_seq()'s result apart.  They never touch the sampler: they are
# here so that the column contract of the result ("<sample>.occurrences/.coverage/.VAF", row identity
# (chr, chr_pos, ref, alt)) is exercised by the same code a user of the reference would run next.

def _mutations_frame(seq_results):
    """a result list is reduced to its "mutations" field (R/seq_to_long.R:34-41)"""
    if isinstance(seq_results, dict) and "mutations" in seq_results:
        return seq_results["mutations"]
    return seq_results


def seq_to_long(seq_results):
    """Wide result -> long format, one block of rows per sample in column order (R/seq_to_long.R:31-67).

    Samples are the columns ending in ".VAF"; per sample the three columns "<sample>.occurrences",
    "<sample>.coverage", "<sample>.VAF" become NV, DP, VAF.  Columns of the result: chr, from, ref, alt, causes,
    classes, NV, DP, VAF, sample_name, to (to == from)."""
    import pandas as pd
    df = _mutations_frame(seq_results)
    samples = [c[:-len(".VAF")] for c in df.columns if c.endswith(".VAF")]
    fixed = ["chr", "chr_pos", "ref", "alt", "causes", "classes"]
    blocks = []
    for sn in samples:
        cols = [f"{sn}.occurrences", f"{sn}.coverage", f"{sn}.VAF"]
        missing = [c for c in fixed + cols if c not in df.columns]
        if missing:
            raise ValueError(f"seq_to_long: column(s) {', '.join(missing)} missing from the sequencing result")
        b = df[fixed + cols].copy()
        b.columns = ["chr", "from", "ref", "alt", "causes", "classes", "NV", "DP", "VAF"]
        b["sample_name"] = sn
        blocks.append(b)
    if not blocks:
        out = pd.DataFrame({k: [] for k in ["chr", "from", "ref", "alt", "causes", "classes", "NV", "DP", "VAF", "sample_name"]})
    else:
        out = pd.concat(blocks, ignore_index=True)
    out["to"] = out["from"]
    return out


def _validate_chromosomes(df, chromosomes):
    """R/ggplot_config.R:75-94"""
    present = list(dict.fromkeys(df["chr"].tolist()))
    if chromosomes is None:
        return present
    chromosomes = [chromosomes] if isinstance(chromosomes, str) else list(chromosomes)
    unknown = [c for c in chromosomes if c not in present]
    if unknown:
        names = ", ".join(map(str, unknown))
        msg = f"The chromosomes {names} are" if len(unknown) > 1 else f"The chromosome {names} is"
        raise ValueError(msg + " not present in the sequence reference data.")
    return chromosomes


def get_seq_data(seq_res, sample, chromosomes=None):
    """{"tumour": long rows of `sample`, "normal": long rows of "normal_sample"}, restricted to `chromosomes`
    (R/plot_genome_wide_mutations.R:3-19)."""
    df = _mutations_frame(seq_res)
    chromosomes = _validate_chromosomes(df, chromosomes)
    data = seq_to_long(df)
    data = data[data["chr"].isin(chromosomes)]
    available = list(dict.fromkeys(data["sample_name"].tolist()))
    if sample not in available:
        raise ValueError("Inserted 'sample' is not available, available samples are: " + ", ".join(available))
    return {"tumour": data[data["sample_name"] == sample].reset_index(drop=True),
            "normal": data[data["sample_name"] == "normal_sample"].reset_index(drop=True)}


def depth_ratio(seq_result, sample, chromosomes=None):
    """The table behind plot_DR(): tumour rows joined with the normal sample on (chr, from, ref, alt),
    DR = DP.tumour / DP.normal (R/plot_genome_wide_mutations.R:75-108; the subsampling to N points and the
    plot itself are not rebuilt)."""
    df = _mutations_frame(seq_result)
    need = ["normal_sample.occurrences", "normal_sample.coverage", "normal_sample.VAF"]
    if not all(c in df.columns for c in need):
        raise ValueError('The parameter "seq_result" does not contain the mandatory normal sample "normal_sample".')
    data = get_seq_data(df, sample, chromosomes)
    d = data["tumour"].merge(data["normal"], on=["chr", "from", "ref", "alt"], how="left", suffixes=(".tumour", ".normal"))
    with np.errstate(divide="ignore", invalid="ignore"):
        d["DR"] = d["DP.tumour"].to_numpy(np.float64) / d["DP.normal"].to_numpy(np.float64)
    return d
