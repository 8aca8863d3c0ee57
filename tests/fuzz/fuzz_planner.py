"""Offline fuzz of the planner (not collected by pytest): random forests x random parameters, template totals,
tile bounds, shard partitions.  usage: python tests/fuzz/fuzz_planner.py SEED N_FORESTS [tree root]"""
import sys, time
import os
ROOT = sys.argv[3] if len(sys.argv) > 3 else os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import numpy as np
from conftest import small_spec, make_params
from process_b200.synth import synth_forest
from process_b200 import _lib as L, _abi as A
rng = np.random.default_rng(int(sys.argv[1])); n = int(sys.argv[2])
t0 = time.time(); nplans = 0
for it in range(n):
    nchr = int(rng.integers(1, 5))
    kw = dict(chr_names=[str(i + 1) for i in range(nchr)], chr_len=[int(rng.integers(2_000, 400_000)) for _ in range(nchr)],
              chr_n_alleles=[int(rng.integers(1, 3)) for _ in range(nchr)],
              sample_cells=[int(rng.integers(1, 9)) for _ in range(int(rng.integers(1, 4)))],
              germline_density=float(rng.choice([0.0, 5e-4, 2e-2])), n_preneo_snv=int(rng.integers(0, 30)),
              n_preneo_indel=int(rng.integers(0, 12)), node_snv_mean=float(rng.choice([0.0, 2.0, 8.0])),
              n_clones=int(rng.integers(1, 5)), clone_cna=int(rng.integers(0, 8)), wgd_clones=int(rng.integers(0, 3)),
              cna_len=(int(rng.integers(200, 1000)), int(rng.integers(1500, 90000))))
    kw["wgd_clones"] = min(kw["wgd_clones"], kw["n_clones"])
    f = synth_forest(small_spec(int(rng.integers(1 << 30)), **kw))
    fl = L.Flat(f)
    if rng.random() < 0.3:
        ng = int(rng.integers(1, 5)); fl.set_groups(rng.integers(0, ng, f.n_leaves).astype(np.uint32), ng)
    else:
        ng = f.n_samples
    for _ in range(3):
        R = int(rng.choice([1, 20, 150, 300]))
        ins = int(rng.choice([0, 0, 120, 400]))
        pk = dict(coverage=float(rng.choice([0.0, 0.3, 5.0, 60.0])), purity=float(rng.choice([0.0, 0.35, 1.0])), read_size=R,
                  insert_size_mean=ins, insert_size_stddev=int(rng.integers(0, 10)), seed=int(rng.integers(1 << 31)),
                  with_normal_sample=int(rng.integers(0, 2)), preneoplastic_in_normal=int(rng.integers(0, 2)),
                  normal_only=int(rng.random() < 0.15))
        if rng.random() < 0.3:
            m = rng.integers(0, 2, nchr).astype(np.uint8)
            pk["chr_mask"] = m
        else:
            m = np.ones(nchr, np.uint8)
        try:
            info, t = fl.plan(make_params(**pk))
        except L.PcsError as e:
            assert "insert" in str(e) or "sample" in str(e).lower() or "nothing" in str(e).lower(), (str(e), pk)
            continue
        nplans += 1
        S = info.n_out_samples
        assert S == (1 if pk["normal_only"] else ng + pk["with_normal_sample"]), (S, pk, ng)
        mates = 2 if ins else 1
        assert info.reads_per_template == mates
        want = sum(int(np.floor(pk["coverage"] * int(f.chr_len[c]) / (R * mates) + 0.5)) for c in range(nchr) if m[c]) * S
        got = int(t["templates"].astype(np.int64).sum())
        # (sample, chromosome) pairs with no DNA at all (everything deleted) draw nothing
        assert got <= want and info.n_templates_total == got, (got, want, pk)
        if got < want:
            assert kw["clone_cna"] > 0
        assert np.all(t["begin"] >= 1) and np.all(t["begin"].astype(np.int64) + t["len"] - 1 <= f.chr_len[t["chr"]])
        assert np.all(m[t["chr"]] == 1) and np.all(t["sample"] < S) and len(np.unique(t["id"])) == len(t["id"])
        sh = int(rng.integers(2, 9)); tot = 0; ids = []
        for r in range(sh):
            ir, tr = fl.plan(make_params(shard_rank=r, shard_count=sh, **pk))
            tot += int(tr["templates"].astype(np.int64).sum()); ids += tr["id"].tolist()
        assert tot == got and sorted(ids) == sorted(t["id"].tolist())
print("ok", n, "forests", nplans, "plans", round(time.time() - t0, 1), "s")
