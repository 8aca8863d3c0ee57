"""The C-ABI shared library: it loads, exports every symbol include/pcs_seq.h declares,
agrees with the header on struct layout, and refuses to compute without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from process_b200 import _abi as A
from process_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcs_seq.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcs_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = L.lib()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pcs_seq.h but not exported"
    assert sorted(L.EXPORTS) == names
    assert lib.pcs_abi_version() == A.PCS_ABI_VERSION


def test_struct_layouts_match_the_header(tmp_path):
    prog = tmp_path / "sizes.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pcs_seq.h"\nint main(void){'
                    'printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pcs_forest_desc), sizeof(pcs_seq_params),'
                    'sizeof(pcs_read_placement), sizeof(pcs_plan_info), sizeof(pcs_run_stats),'
                    'offsetof(pcs_seq_params, chr_mask), offsetof(pcs_seq_params, shard_rank),'
                    'offsetof(pcs_forest_desc, germ_allele_mask));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(A.ForestDesc), C.sizeof(A.SeqParams), C.sizeof(A.ReadPlacement), C.sizeof(A.PlanInfo),
            C.sizeof(A.RunStats), A.SeqParams.chr_mask.offset, A.SeqParams.shard_rank.offset,
            A.ForestDesc.germ_allele_mask.offset]
    assert got == want


def test_header_is_plain_c(tmp_path):
    prog = tmp_path / "c89.c"
    prog.write_text('#include "pcs_seq.h"\nint main(void){return pcs_abi_version == 0;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Werror", "-fsyntax-only", "-I",
                           os.path.join(ROOT, "include"), str(prog)])


def test_compute_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(L.PcsError) as e:
        L.Context(0)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may use oracle/."""
    pkg = os.path.join(ROOT, "process_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cpp", ".cu", ".hpp", ".h")) or fn == "Makefile":
                text = open(os.path.join(dp, fn)).read()
                assert "oracle" not in text.lower() or fn in ("flatten.cpp",), fn
