"""Offline fuzz of the flattener (not collected by pytest): random forests, flat view against the explicit
genomes of every cell.  usage: python tests/fuzz/fuzz_flattener.py SEED N_FORESTS"""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import numpy as np
from conftest import small_spec
from process_b200.synth import synth_forest
from test_host_logic import check_flat_against_explicit_genomes
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
t0 = time.time()
for it in range(n):
    seed = int(rng.integers(1 << 30))
    nchr = int(rng.integers(1, 5))
    kw = dict(chr_names=[str(i + 1) for i in range(nchr)], chr_len=[int(rng.integers(20_000, 200_000)) for _ in range(nchr)],
              chr_n_alleles=[int(rng.integers(1, 3)) for _ in range(nchr)],
              sample_cells=[int(rng.integers(1, 9)) for _ in range(int(rng.integers(1, 4)))],
              germline_density=float(rng.choice([0.0, 5e-4, 3e-3])), n_preneo_snv=int(rng.integers(0, 30)),
              n_preneo_indel=int(rng.integers(0, 12)), node_snv_mean=float(rng.choice([0.0, 2.0, 8.0])),
              n_clones=int(rng.integers(1, 5)), clone_cna=int(rng.integers(0, 8)), wgd_clones=int(rng.integers(0, 3)),
              cna_len=(int(rng.integers(500, 3000)), int(rng.integers(5000, 90000))))
    kw["wgd_clones"] = min(kw["wgd_clones"], kw["n_clones"])
    try:
        f = synth_forest(small_spec(seed, **kw))
        k = check_flat_against_explicit_genomes(f); tot = globals().get("tot", 0) + k
    except Exception as e:
        print("FAIL", seed, kw, repr(e)); raise
print("checked", tot); print("ok", n, "forests", round(time.time() - t0, 1), "s")
