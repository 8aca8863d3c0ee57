"""The C++ host mirror (process_b200/csrc/process_seq.hpp): error behaviour on CPU, and on the
GPU the same tables as the Python mirror for the same seed (both sit on the same C ABI)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = tmp_path_factory.mktemp("cpp") / "host_mirror_test"
    lib_dir = os.path.join(ROOT, "process_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"),
                           "-o", str(exe), "-L", lib_dir, "-lpcs_seq", f"-Wl,-rpath,{lib_dir}"])
    return str(exe)


def test_cpp_mirror_argument_handling(driver):
    out = subprocess.run([driver, "validate"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


@pytest.mark.gpu
def test_cpp_mirror_matches_python_mirror(driver, tmp_path):
    from golden import micro_forest as MF
    from process_b200 import api
    ref = tmp_path / "ref.fa"
    ref.write_text(">1 micro\n" + MF.REF + "\n")
    sam_dir = tmp_path / "normal_sam"
    out = subprocess.run([driver, "run", str(ref), str(sam_dir)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().split("\n")
    header = lines[0].split("\t")
    rows = [l.split("\t") for l in lines[1:] if not l.startswith("#")]
    f = MF.forest()
    f.reference_path = str(ref)
    f.mut_nature_mask = np.asarray([4, 4, 8, 4, 2, 2, 2, 4], np.uint8)
    r = api.simulate_seq(f, sequencer=api.BasicIlluminaSequencer(1e-2, False), chromosomes=["1"], coverage=400.0,
                         read_size=20, purity=0.8, preneoplastic_in_normal=True, seed=7)
    df = r["mutations"]
    assert header[:6] == list(df.columns[:6]) and header == list(df.columns)
    assert len(rows) == len(df) > 0
    for i, row in enumerate(rows):
        assert int(row[1]) == int(df["chr_pos"].iloc[i])
        assert row[5] == df["classes"].iloc[i]
        for j in range(6, len(header)):
            a, b = float(row[j]), float(df.iloc[i, j])
            assert (np.isnan(a) and np.isnan(b)) or abs(a - b) < 1e-6, (i, header[j], a, b)
    meta = [l for l in lines if l.startswith("#seed")][0].split("\t")
    assert meta[1] == "7" and meta[3] == "BasicIlluminaSequencer" and int(meta[5]) == r["_stats"]["n_reads"]
    normal = [l for l in lines if l.startswith("#normal")][0].split("\t")
    assert normal[1] == "1" and normal[2] == "normal_sample"
    # simulate_normal_seq wrote its SAM file (write_SAM = TRUE is the reference's default for it)
    sam = (sam_dir / "chr_1.sam").read_text().split("\n")
    assert sam[0].startswith("@HD") and any(l.startswith("@RG\tID:normal_sample") for l in sam)
    reads = [l.split("\t") for l in sam if l and not l.startswith("@")]
    assert len(reads) > 10_000 and all(r[5] in ("20M",) or "D" in r[5] or "I" in r[5] for r in reads[:2000])
    api.release_device_cache()
