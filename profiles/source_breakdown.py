#!/usr/bin/env python
"""Where the warp instructions of the sampler kernel go, by source function.

    python profiles/source_breakdown.py gpurun_out/prof.ncu-rep r01_v17 [reads_per_launch]

Reads the source page of an `ncu --set full --import-source on` capture (kernels are built with
-lineinfo) in the cuda,sass view, sums "Instructions Executed" and the stall samples per source line of
process_b200/csrc/kernels.cu, maps the lines to the device functions that hold them and writes
profiles/<tag>_source_breakdown.md.  Counts are quoted per 32 reads (one read per lane)."""
import csv
import io
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "process_b200", "csrc", "kernels.cu")


def function_ranges():
    """(first line, name) of every function / lambda worth naming, in file order."""
    marks = []
    pat = re.compile(r"^(?:template <[^>]*>\s*)?(?:__device__|__global__|static|struct)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*(?:\(|\{|$)")
    lam = re.compile(r"^\s*auto ([a-z_]+) = \[&\]")
    with open(SRC) as fh:
        for no, line in enumerate(fh, 1):
            m = lam.match(line)
            if m:
                marks.append((no, "kernel: " + m.group(1)))
                continue
            if line.startswith(("__device__", "__global__", "struct ", "sample_tiles_staged_kernel")):
                m = pat.match(line)
                name = m.group(1) if m else line.split("(")[0].split()[-1]
                if line.startswith("sample_tiles_staged_kernel"):
                    name = "kernel: staging + flush"
                marks.append((no, name))
            if "// Thinned tile: only the templates whose read can span a locus are drawn" in line:
                marks.append((no, "kernel: thinned-tile loop (draw -> window -> start)"))
            if "// Philox block j:" in line:
                marks.append((no, "kernel: draw + probe loop"))
            if "// ---- flush:" in line:
                marks.append((no, "kernel: staging + flush"))
    return marks


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    reads = float(sys.argv[3]) if len(sys.argv) > 3 else 6588301254.0
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                                  text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    per_line = {}
    in_file = False
    for r in rows:
        if len(r) >= 2 and r[0] in ("File Path", "File Name"):
            in_file = r[1].endswith("kernels.cu")
            continue
        if not in_file or len(r) <= i_inst or not r[0].isdigit():
            continue
        try:
            old = per_line.get(int(r[0]), (0.0, 0.0))
            per_line[int(r[0])] = (old[0] + float(r[i_inst]), old[1] + float(r[i_samp]))
        except ValueError:
            pass
    marks = function_ranges()
    agg = {}
    for no, (inst, samp) in per_line.items():
        name = "?"
        for first, nm in marks:
            if first <= no:
                name = nm
            else:
                break
        a = agg.setdefault(name, [0.0, 0.0])
        a[0] += inst
        a[1] += samp
    tot_i = sum(a[0] for a in agg.values())
    tot_s = sum(a[1] for a in agg.values()) or 1.0
    unit = reads / 32.0
    lines = [f"# warp instructions of the sampler kernel by source function ({os.path.basename(rep)})", "",
             f"{tot_i:.4g} warp instructions per launch, {reads:.4g} reads: {tot_i / unit:.1f} per 32 reads.", "",
             "| source function (process_b200/csrc/kernels.cu) | warp instr / 32 reads | share | stall samples |",
             "|---|---|---|---|"]
    for name, (inst, samp) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if inst / tot_i < 0.002:
            continue
        lines.append(f"| `{name}` | {inst / unit:.2f} | {100 * inst / tot_i:.1f} % | {100 * samp / tot_s:.1f} % |")
    out = os.path.join(HERE, f"{tag}_source_breakdown.md")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
