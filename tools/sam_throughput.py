#!/usr/bin/env python
"""SAM output throughput: simulate_normal_seq(write_SAM=TRUE) -- the reference's default for the
normal sample (src/sequencing.cpp:268-282) -- on the C2 forest (demo chr22).  Prints one JSON line: reads/s, Gbases/s
and GB/s of SAM text, split into GPU materialisation (kernel + records back), host formatting and file write.
    PCS_TIMING=1 python tools/sam_throughput.py [coverage]"""
import json, os, re, shutil, subprocess, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if os.environ.get("PCS_SAM_CHILD") != "1":
    env = dict(os.environ, PCS_SAM_CHILD="1", PCS_TIMING="1")
    p = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, capture_output=True, text=True)
    m = re.findall(r"\[pcs sam\] reads (\d+) text_bytes (\d+) gpu_materialise_ms ([\d.]+) host_format_ms ([\d.]+) file_write_ms ([\d.]+)", p.stderr)
    line = json.loads(p.stdout.strip().split("\n")[-1]) if p.returncode == 0 else {"error": p.stderr[-2000:]}
    if m:
        reads, nbytes, g, f, w = m[-1]
        line["split_ms"] = {"gpu_materialise": float(g), "host_format": float(f), "file_write": float(w)}
        line["text_gb_per_s_of_host_formatting_alone"] = int(nbytes) / 1e9 / (float(f) / 1e3)
    print(json.dumps(line))
    sys.exit(p.returncode)

import numpy as np
from process_b200 import api
from process_b200.synth import config_spec, synth_forest

coverage = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
f = synth_forest(config_spec("C1", seed=0))
tmp = tempfile.mkdtemp(prefix="pcs_sam_")
ref = os.path.join(tmp, "ref.fa")
rng = np.random.default_rng(0)
with open(ref, "w") as fh:
    for name, n in zip(f.chr_names, f.chr_len):
        fh.write(f">{name}\n")
        fh.write("".join(np.asarray(list("ACGT"))[rng.integers(0, 4, int(n))]) + "\n")
f.reference_path = ref
out = os.path.join(tmp, "sam")
api.simulate_normal_seq(f, coverage=0.5, output_dir=out + "_warm", seed=1)   # FASTA load, contexts, pinned buffers
t0 = time.perf_counter()
r = api.simulate_normal_seq(f, sequencer=api.BasicIlluminaSequencer(1e-3, True), coverage=coverage, output_dir=out, seed=2)
dt = time.perf_counter() - t0
size = sum(os.path.getsize(os.path.join(out, x)) for x in os.listdir(out))
reads = r["_stats"]["n_reads"]
print(json.dumps({"what": "simulate_normal_seq(write_SAM=TRUE), C2 forest (chr22), BasicIllumina(1e-3, random quality)",
                  "coverage": coverage, "reads": int(reads), "seconds": dt, "reads_per_s": reads / dt,
                  "gbases_per_s": reads * 150 / dt / 1e9, "sam_bytes": size, "sam_gb_per_s": size / dt / 1e9}))
shutil.rmtree(tmp, ignore_errors=True)
