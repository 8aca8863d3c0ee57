"""Argument handling of the Python mirror of the Rcpp interface (no GPU needed):
the error behaviour of src/seq_simulation.cpp and src/sequencers.cpp."""
import numpy as np
import pytest

from process_b200 import api
from process_b200 import _abi as A

from golden import micro_forest as MF


def test_sequencer_constructors_validate_like_build_sequencer():
    s = api.BasicIlluminaSequencer(4e-3)
    assert s.error_rate == 4e-3 and s.random_quality_scores is True
    assert api._sequencer_model(s) == (A.PCS_SEQ_BASIC_RANDOM, 4e-3)
    assert api._sequencer_model(api.BasicIlluminaSequencer(1e-3, False)) == (A.PCS_SEQ_BASIC_CONSTANT, 1e-3)
    assert api._sequencer_model(None) == api._sequencer_model(api.ErrorlessIlluminaSequencer()) == (A.PCS_SEQ_ERRORLESS, 0.0)
    for bad in (-1, "x", None, True):
        with pytest.raises(ValueError, match="error_rate"):
            api.BasicIlluminaSequencer(bad)
    with pytest.raises(ValueError, match="random_quality_scores"):
        api.BasicIlluminaSequencer(1e-3, "yes")
    with pytest.raises(ValueError, match="Unsupported sequencer type"):
        api._sequencer_model(object())
    assert api._sequencer_data(None) is None
    assert api._sequencer_data(s) == dict(name="BasicIlluminaSequencer", error_rate=4e-3, random_quality_scores=True)
    assert api._sequencer_data(api.ErrorlessIlluminaSequencer()) == dict(name="ErrorlessIlluminaSequencer", error_rate=0.0)


def test_reference_genome_must_exist(tmp_path):
    f = MF.forest()
    with pytest.raises(RuntimeError, match="does not exists anymore"):
        api.simulate_seq(f)
    with pytest.raises(RuntimeError, match="does not exists"):
        api.simulate_seq(f, reference_genome=str(tmp_path / "nope.fa"))
    with pytest.raises(ValueError, match="either NULL or a string"):
        api.simulate_seq(f, reference_genome=3)


def test_chromosome_list_handling():
    f = MF.forest()
    assert api._chr_mask(f, None) is None
    assert api._chr_mask(f, ["1"]).tolist() == [1]
    with pytest.raises(ValueError, match="2nd element of the list is not a string"):
        api._chr_mask(f, ["1", 2])
    with pytest.raises(ValueError, match="Unsupported chromosome list type"):
        api._chr_mask(f, 22)


def test_cell_labelling_splits_samples_in_order_of_first_appearance():
    f = MF.forest()
    group, names = api._apply_FACS_labels(f, lambda c: "" if c.cell_id == 1 else "B")
    assert names == ["s0", "s1_B"] and group.tolist() == [0, 1]
    group, names = api._apply_FACS_labels(f, None)
    assert group is None and names == ["s0", "s1"]
    with pytest.raises(ValueError, match="must be a function"):
        api._apply_FACS_labels(f, 3)


def test_seed_resolution():
    assert api._resolve_seed(7) == 7 and api._resolve_seed(7.0) == 7
    assert -2**31 <= api._resolve_seed(None) < 2**31
    # NULL: drawn from the session's RNG state, so seeding the session governs the run (src/utility.hpp:50-57)
    np.random.seed(3)
    a = [api._resolve_seed(None) for _ in range(3)]
    np.random.seed(3)
    assert a == [api._resolve_seed(None) for _ in range(3)] and len(set(a)) == 3
    with pytest.raises(ValueError, match="The seed must be either a number or NILL."):
        api._resolve_seed("x")


def test_vectorised_labelling_equals_per_cell_labelling():
    """one evaluation over all sampled cells gives the samples (names, order of first
    appearance, membership) of the per-cell callback of src/seq_simulation.cpp:183-243."""
    import numpy as np
    from conftest import small_spec
    from process_b200.synth import synth_forest
    f = synth_forest(small_spec(2, sample_cells=[40, 35, 50]))
    rng = np.random.default_rng(0)
    f.leaf_attrs = {"epistate": rng.choice(np.array(["+", "-", ""], dtype=object), f.n_leaves),
                    "mutant": rng.choice(np.array(["A", "B"], dtype=object), f.n_leaves)}
    per_cell = lambda c: c.epistate if c.mutant == "A" else ""
    g0, n0 = api._apply_FACS_labels(f, per_cell)
    vec = api.VectorisedLabelling(lambda cols: np.where(cols["mutant"] == "A", cols["epistate"], ""))
    g1, n1 = api._apply_FACS_labels(f, vec)
    assert n0 == n1 and len(n0) > f.n_samples
    assert np.array_equal(g0, g1)
    with pytest.raises(ValueError, match="one label per sampled cell"):
        api._apply_FACS_labels(f, api.VectorisedLabelling(lambda cols: ["x"]))
    with pytest.raises(ValueError, match="must return a string"):
        api._apply_FACS_labels(f, api.VectorisedLabelling(lambda cols: np.arange(f.n_leaves)))
    with pytest.raises(ValueError, match="must be a function"):
        api.VectorisedLabelling(3)


def _wide_example():
    """the data frame of the reference's own example, R/seq_to_long.R:14-26"""
    import pandas as pd
    return pd.DataFrame({
        "chr": ["chr1", "chr2"], "chr_pos": [100, 200], "ref": ["A", "C"], "alt": ["T", "G"],
        "causes": ["SBS5", "SBS1"], "classes": ["germinal", "passneger"],
        "Sample.A.occurrences": [10, 90], "Sample.A.coverage": [100, 100], "Sample.A.VAF": [0.1, 0.9],
        "normal_sample.occurrences": [45, 52], "normal_sample.coverage": [100, 100], "normal_sample.VAF": [0.45, 0.52]})


def test_seq_to_long_on_the_reference_example():
    """wide -> long as R/seq_to_long.R:31-67 does it: samples found by the ".VAF" suffix, one block per sample in
    column order, occurrences/coverage renamed NV/DP, chr_pos renamed from, to == from"""
    wide = _wide_example()
    long = api.seq_to_long(wide)
    assert list(long.columns) == ["chr", "from", "ref", "alt", "causes", "classes", "NV", "DP", "VAF", "sample_name", "to"]
    assert long["sample_name"].tolist() == ["Sample.A", "Sample.A", "normal_sample", "normal_sample"]
    assert long["NV"].tolist() == [10, 90, 45, 52] and long["DP"].tolist() == [100] * 4
    assert long["VAF"].tolist() == [0.1, 0.9, 0.45, 0.52]
    assert long["from"].tolist() == [100, 200, 100, 200] == long["to"].tolist()
    assert long["chr"].tolist() == ["chr1", "chr2"] * 2 and long["classes"].tolist() == ["germinal", "passneger"] * 2
    # a result list is reduced to its "mutations" field
    assert api.seq_to_long({"mutations": wide, "parameters": {}}).equals(long)
    with pytest.raises(ValueError, match="Sample.A.coverage"):
        api.seq_to_long(wide.drop(columns=["Sample.A.coverage"]))
    assert len(api.seq_to_long(wide[["chr", "chr_pos", "ref", "alt", "causes", "classes"]])) == 0


def test_get_seq_data_and_depth_ratio_follow_the_plot_helpers():
    """R/plot_genome_wide_mutations.R:3-19, 90-108 and R/ggplot_config.R:75-94"""
    wide = _wide_example()
    d = api.get_seq_data(wide, "Sample.A")
    assert d["tumour"]["NV"].tolist() == [10, 90] and d["normal"]["NV"].tolist() == [45, 52]
    d = api.get_seq_data(wide, "Sample.A", chromosomes=["chr2"])
    assert d["tumour"]["from"].tolist() == [200] and d["normal"]["from"].tolist() == [200]
    with pytest.raises(ValueError, match="The chromosome chr9 is not present in the sequence reference data."):
        api.get_seq_data(wide, "Sample.A", chromosomes=["chr1", "chr9"])
    with pytest.raises(ValueError, match="The chromosomes chr8, chr9 are not present"):
        api.get_seq_data(wide, "Sample.A", chromosomes=["chr8", "chr9"])
    with pytest.raises(ValueError, match="available samples are: Sample.A, normal_sample"):
        api.get_seq_data(wide, "Sample.B")
    wide["Sample.A.coverage"] = [150, 50]
    dr = api.depth_ratio({"mutations": wide}, "Sample.A")
    assert dr["DR"].tolist() == [1.5, 0.5]
    assert {"DP.tumour", "DP.normal", "VAF.tumour", "VAF.normal", "NV.tumour", "NV.normal"} <= set(dr.columns)
    with pytest.raises(ValueError, match='mandatory normal sample "normal_sample"'):
        api.depth_ratio(wide.drop(columns=["normal_sample.VAF"]), "Sample.A")


def test_result_dataframe_from_dictionary_codes_equals_per_row_strings():
    """the string columns are built from their dictionary encoding (no Python object per row);
    the frame must be the one the per-row construction gives -- values, NA positions, dtypes -- also when there
    are no rows and when no row has a cause"""
    import pandas as pd
    from conftest import small_spec
    from process_b200.synth import synth_forest

    class FakeDevice:
        def __init__(self, rows):
            self.rows = rows

        def active_rows(self, occ, include_non_sequenced, params):
            return self.rows

    def per_row(forest, rows, occ, cov, names):
        ref, alt = forest.row_strings(rows)
        cols = {"chr": np.asarray(forest.chr_names, dtype=object)[forest.mut_chr[rows]],
                "chr_pos": forest.mut_pos[rows].astype(np.int32), "ref": ref, "alt": alt,
                "causes": forest.row_causes(rows), "classes": forest.row_classes(rows)}
        for s in sorted(range(len(names)), key=lambda i: names[i]):
            o, c = occ[s, rows].astype(np.int32), cov[s, rows].astype(np.int32)
            # a row the sample never covered: VAF 0 (/root/reference/src/seq_simulation.cpp:129-131)
            vaf = np.asarray([oo / cc if cc else 0.0 for oo, cc in zip(o.tolist(), c.tolist())], np.float64)
            cols[f"{names[s]}.occurrences"], cols[f"{names[s]}.coverage"], cols[f"{names[s]}.VAF"] = o, c, vaf
        return pd.DataFrame(cols)

    for f in (MF.forest(), synth_forest(small_spec(0)), synth_forest(small_spec(4))):
        names = list(f.sample_names) + ["normal_sample"]
        rng = np.random.default_rng(1)
        cov = rng.poisson(30, (len(names), f.n_mut)).astype(np.uint32)
        occ = rng.integers(0, 20, (len(names), f.n_mut)).astype(np.uint32)
        cov[:, ::7] = 0
        occ[:, ::7] = 0
        indels = np.flatnonzero((f.mut_ref_len != 1) | (f.mut_alt_len != 1)).astype(np.uint32)
        for rows in (np.arange(f.n_mut, dtype=np.uint32), np.zeros(0, np.uint32),
                     np.arange(0, f.n_mut, 3, dtype=np.uint32), indels):
            a = api._frame_from_tables(f, FakeDevice(rows), occ, cov, names, False)
            b = per_row(f, rows, occ, cov, names)
            assert list(a.columns) == list(b.columns) and (a.dtypes == b.dtypes).all() and a.equals(b)
    ref_codes, ref_table, alt_codes, alt_table = f.row_string_codes(indels)
    ref, alt = f.row_strings(indels)
    assert len(indels) > 5 and list(ref_table[ref_codes]) == list(ref) and list(alt_table[alt_codes]) == list(alt)
