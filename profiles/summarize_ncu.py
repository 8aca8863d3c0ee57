#!/usr/bin/env python
"""Turn an ncu report of the sampler kernel into the small files kept under profiles/:
   python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep r01_v4 [--no-traffic]
writes profiles/<tag>_sampler_ncu_raw.csv, <tag>_sampler_summary.md and (unless --no-traffic: captures of
other sequencer models than the headline one) sampler_traffic.json, which bench.py reads."""
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
    if "--no-traffic" in sys.argv:
        print("\n".join(lines[:14]))
        return
    traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    inst = float(m["smsp__inst_executed.sum"][0].replace(",", ""))
    json.dump({"kernel": name, "dram_bytes_per_launch": traffic, "warp_instructions_per_launch": inst,
               "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
               "source": f"profiles/{tag}_sampler_ncu_raw.csv",
               "command": "ncu --set full --clock-control none --import-source on -k regex:sample_tiles_staged "
                          "-s 3 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"},
              open(os.path.join(HERE, "sampler_traffic.json"), "w"), indent=1)
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    main()
