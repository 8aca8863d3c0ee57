#!/usr/bin/env python
"""All five BASELINE.json configurations at full size, sampler kernel only (device-resident outputs).

    python profiles/all_configs.py gpurun_out/configs.json      # on a B200

Per configuration: 2 warm-up passes, 5 timed passes of pcs_plan_run with device outputs; kernel time is the
library's CUDA-event bracket around the sampler launch (stats.kernel_ms)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402  (device memory for the output tables)

from bench import make_params  # noqa: E402
from process_b200 import _abi as A  # noqa: E402
from process_b200 import _lib as L  # noqa: E402
from process_b200.synth import config_spec, synth_forest  # noqa: E402

CASES = [
    ("C1 demo chr22 50x errorless", "C1", dict(coverage=50.0)),
    ("C2 chr22 200x BasicIllumina(1e-3, random) purity 0.8", "C1",
     dict(coverage=200.0, purity=0.8, sequencer=A.PCS_SEQ_BASIC_RANDOM, error_rate=1e-3)),
    ("C2 simulate_normal_seq 200x", "C1", dict(coverage=200.0, normal_only=1, with_normal_sample=0)),
    ("C3 WGS 80x errorless", "C3", dict(coverage=80.0)),
    ("C3 WGS 80x BasicIllumina(1e-3, constant)", "C3",
     dict(coverage=80.0, sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=1e-3)),
    ("C3 WGS 80x BasicIllumina(1e-3, random)", "C3",
     dict(coverage=80.0, sequencer=A.PCS_SEQ_BASIC_RANDOM, error_rate=1e-3)),
    ("C3 WGS 80x paired-end insert 300", "C3", dict(coverage=80.0, insert_size_mean=300)),
    ("C4 8x5000 cells 200x", "C4", dict(coverage=200.0, purity=0.9)),
    ("C5 1e5 cells WGD 300x purity 0.9 BasicIllumina(1e-3, constant)", "C5",
     dict(coverage=300.0, purity=0.9, sequencer=A.PCS_SEQ_BASIC_CONSTANT, error_rate=1e-3)),
]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "configs.json"
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = L.Context(0, stream.cuda_stream)
    forests, out = {}, []
    for name, cfg, kw in CASES:
        if cfg not in forests:
            forests.clear()  # one forest resident at a time
            t0 = time.perf_counter()
            f = synth_forest(config_spec(cfg, seed=0))
            t1 = time.perf_counter()
            dev = L.Forest(ctx, f)
            forests[cfg] = (f, dev, t1 - t0, time.perf_counter() - t1)
        f, dev, synth_s, upload_s = forests[cfg]
        P = make_params(**kw)
        t0 = time.perf_counter()
        plan = L.Plan(dev, P)
        plan_s = time.perf_counter() - t0
        S, M = plan.info.n_out_samples, plan.info.n_mut
        occ = torch.zeros((S, M), dtype=torch.int32, device="cuda")
        cov = torch.zeros((S, M), dtype=torch.int32, device="cuda")
        for _ in range(2):
            plan.run_device(occ.data_ptr(), cov.data_ptr())
        stats = [plan.run_device(occ.data_ptr(), cov.data_ptr()) for _ in range(5)]
        ms = float(np.mean([s.kernel_ms for s in stats]))
        info = dev.info()
        row = {"config": name, "cells": int(f.n_leaves), "rows": int(M), "samples": int(S),
               "reads": int(stats[-1].n_reads), "kernel_ms": round(ms, 3),
               "gbases_per_s": round(stats[-1].n_reads * plan.info.read_size / (ms * 1e-3) / 1e9),
               # A11 (unpinned): templates dropped because they ran past their fragment's end, i.e. the coverage
               # deficit of "uniform over the fragment + drop" against "uniform over the valid starts"
               "a11_dropped_fraction": 1.0 - stats[-1].n_reads / float(stats[-1].n_templates * plan.info.reads_per_template),
               "flatten_upload_s": round(upload_s, 3), "plan_s": round(plan_s, 3),
               "device_MB": int(info["device_bytes"] // 1_000_000), "haplotypes": int(info["n_haplotypes"]),
               "tiles": int(plan.info.n_tiles)}
        print(json.dumps(row), flush=True)
        out.append(row)
        plan.close()
    with open(out_path, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
