#!/usr/bin/env python
"""Turn an ncu report of the sampler kernel into the small files kept under profiles/:
   python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep r01_v4 [--no-traffic]
   python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep r02_v3 --key C3/errorless/single
writes profiles/<tag>_sampler_ncu_raw.csv, <tag>_sampler_summary.md and the record `--key` (workload / sequencer /
single|paired) of sampler_traffic.json, which bench.py reads for roofline.traffic and roofline.issue."""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__t_sectors.sum", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_shared_atom.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
]


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    open(os.path.join(HERE, f"{tag}_sampler_ncu_raw.csv"), "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    name = m.get("Kernel Name", ("?", ""))[0]
    lines = [f"# ncu --set full --clock-control none, kernel `{name}`", "",
             "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in m:
            lines.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            lines.append(f"| {h} | {m[h][0]} | {m[h][1]} |")
    open(os.path.join(HERE, f"{tag}_sampler_summary.md"), "w").write("\n".join(lines) + "\n")

    def to_bytes(key):
        v, u = m[key]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        return float(v.replace(",", "")) * scale
    # sampler_traffic.json: per-launch counters bench.py multiplies with its live kernel time, one record per
    # workload, stamped with the SHA of the kernel sources (bench.py flags a record whose sources have changed)
    key = "C3/errorless/single"
    if "--key" in sys.argv:
        key = sys.argv[sys.argv.index("--key") + 1]
    sys.path.insert(0, os.path.dirname(HERE))
    import bench
    num = lambda k: float(m[k][0].replace(",", ""))
    t_s = num("gpu__time_duration.sum") * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[m["gpu__time_duration.sum"][1]]
    l2_bytes = num("lts__t_sectors.sum") * 32.0
    l2_pct = num("lts__t_sectors.sum.pct_of_peak_sustained_elapsed")
    rec = {"kernel": name, "kernel_source_sha": bench.kernel_source_sha(), "capture": f"profiles/{tag}_sampler_ncu_raw.csv",
           "kernel_ms_under_ncu": t_s * 1e3,
           "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
           "warp_instructions_per_launch": num("smsp__inst_executed.sum"),
           "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "lanes_per_instruction": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
           "l2_bytes_per_launch": l2_bytes,
           "l2_peak_bytes_per_s": l2_bytes / t_s / (l2_pct / 100.0) if l2_pct > 0 else None,
           "l2_hit_rate_pct": num("lts__t_sector_hit_rate.pct"),
           "shared_atomics_per_launch": num("smsp__inst_executed_op_shared_atom.sum"),
           "global_reds_per_launch": num("smsp__inst_executed_op_global_red.sum"),
           "shared_bank_conflicts_per_launch": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
           "command": "ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged -s 3 -c 1 "
                      "python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e [--workload ... --sequencer ...]"}
    path = os.path.join(HERE, "sampler_traffic.json")
    book = {"captures": {}}
    if os.path.exists(path):
        old = json.load(open(path))
        if "captures" in old:
            book = old
    book["captures"][key] = rec
    json.dump(book, open(path, "w"), indent=1)
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    main()
