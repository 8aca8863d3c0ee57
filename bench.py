#!/usr/bin/env python
"""bench.py -- simulated Gbases/s of the simulate_seq() read sampler on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C3]

A step = one full pass of the hot path over the workload: every read of every sample of the synthetic forest is
placed and counted.  Default workload C3 = BASELINE.json configs[2] (GRCh38-length genome, 3 tumour samples x 1000
cells + normal sample, 80x WGS, read size 150, errorless Illumina); C1, C2, C4, C5 are the other configs at full size.
N > 1: one process per GPU under torchrun, the tile grid is sharded over ranks (strong scaling of the same job)
and every rank's sampler adds straight into rank 0's tables over NVLink peer memory.

Prints ONE JSON line (rank 0):
  value      device-timed whole-job throughput, forest and plan resident in HBM
  e2e        the same metric through the C ABI with HOST buffers: flatten + upload + plan + kernels + tables back
             (N > 1: one process driving N GPUs, pcs_simulate_multi -- the call an R session makes)
  roofline   the sampler kernel: the SURVEY 8(d) byte model against HBM peak (a model, not a bound: the kernel
             keeps reads in registers) and, as `issue`, the warp-instruction issue rate that actually bounds it
  cpu_baseline  the CPU oracle on one host thread, >= 2e8 reads over all chromosomes
`--impl reference` times the CPU oracle (the reference's RACES code is not in the image and ProCESS does not compile
without Rcpp) with all host threads on bounded samples of the same workload.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from process_b200 import _abi as A  # noqa: E402
from process_b200.synth import config_spec, synth_forest  # noqa: E402

METRIC = "simulated_gbases_per_s"
UNIT = "Gbases/s"
SEQ = {"errorless": A.PCS_SEQ_ERRORLESS, "constant": A.PCS_SEQ_BASIC_CONSTANT, "random": A.PCS_SEQ_BASIC_RANDOM}
WORKLOADS = {
    # name: (synth config, genome scale, simulate_seq arguments, BASELINE.json config it is)
    "C1": ("C1", 1.0, dict(coverage=50.0, purity=1.0, read_size=150), "configs[0]: demo forest, chr22, 50x, errorless"),
    "C2": ("C1", 1.0, dict(coverage=200.0, purity=0.8, read_size=150, sequencer="random", error_rate=1e-3),
           "configs[1]: demo forest, chr22, 200x, BasicIllumina(1e-3), purity 0.8"),
    "C3": ("C3", 1.0, dict(coverage=80.0, purity=1.0, read_size=150),
           "configs[2]: GRCh38-length genome, 3 tumour samples x 1000 cells + normal, 80x WGS, errorless"),
    "C4": ("C4", 1.0, dict(coverage=200.0, purity=0.9, read_size=150),
           "configs[3]: 8 samples x 5000 cells + normal, 200x WGS, purity 0.9, errorless"),
    "C5": ("C5", 1.0, dict(coverage=300.0, purity=0.9, read_size=150, sequencer="constant", error_rate=1e-3),
           "configs[4]: 1e5 cells, WGD + dense CNAs, 1e6 SNVs/genome, tumour + normal 300x, BasicIllumina(1e-3)"),
    "C3-small": ("C3", 0.02, dict(coverage=80.0, purity=1.0, read_size=150), "C3 on 2 % of the genome (smoke runs)"),
}


def make_params(shard_rank=0, shard_count=1, chr_mask=None, **kw):
    d = dict(seed=0, coverage=80.0, purity=1.0, read_size=150, insert_size_mean=0, insert_size_stddev=10,
             sequencer=A.PCS_SEQ_ERRORLESS, error_rate=0.0, with_normal_sample=1, preneoplastic_in_normal=0,
             normal_only=0, shard_rank=shard_rank, shard_count=shard_count)
    d.update(kw)
    if isinstance(d["sequencer"], str):
        d["sequencer"] = SEQ[d["sequencer"]]
    p = A.SeqParams(**d)
    if chr_mask is not None:
        import ctypes as C
        p._keep = np.ascontiguousarray(chr_mask, dtype=np.uint8)
        p.chr_mask = A.ptr(p._keep, C.c_uint8)
    return p


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy burst)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha():
    """what the ncu figures were captured from: the sampler's sources"""
    h = hashlib.sha256()
    for name in ("kernels.cu", "dev.hpp"):
        with open(os.path.join(ROOT, "process_b200", "csrc", name), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_counters(workload, sequencer, insert):
    """per-launch ncu counters of the sampler kernel for this workload from the committed capture
    (profiles/sampler_traffic.json, written by profiles/summarize_ncu.py), with `stale` = the kernel sources
    have changed since the capture.  None if there is no capture of this workload."""
    path = os.path.join(ROOT, "profiles", "sampler_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        book = json.load(fh)
    key = f"{workload}/{sequencer}/{'paired' if insert else 'single'}"
    rec = book.get("captures", {}).get(key)
    if rec is None:
        return None
    rec = dict(rec)
    rec["stale"] = rec.get("kernel_source_sha") != kernel_source_sha()
    return rec


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, p[2:]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(args, wl_params):
    """config of the JSON line: the same object in both arms"""
    return {"workload": f"{args.workload}: {WORKLOADS[args.workload][3]}",
            "coverage": wl_params["coverage"], "purity": wl_params["purity"], "read_size": wl_params["read_size"],
            "sequencer": wl_params.get("sequencer", "errorless"), "error_rate": wl_params.get("error_rate", 0.0),
            "insert_size_mean": wl_params.get("insert_size_mean", 0)}


def cpu_sample(forest, wl_params, budget_reads, threads):
    """a bounded sample of the workload for the CPU legs -- the whole genome, every sample,
    every chromosome in proportion (the coverage is scaled down), about `budget_reads` reads.  Returns the figures
    of the run: throughput of the read loop alone (what a full-size run converges to: the oracle's fixed work --
    explicit genomes, zeroing the per-base coverage vectors -- is 1 % of a full run but a third of a sample), the
    plain wall-clock throughput of the sample, and the full job extrapolated linearly in the number of reads."""
    import oracle
    n_out = forest.n_samples + 1
    R = wl_params["read_size"]
    mates = 2 if wl_params.get("insert_size_mean", 0) else 1
    reads_full = sum(round(wl_params["coverage"] * int(n) / (R * mates)) * mates for n in forest.chr_len) * n_out
    frac = min(1.0, budget_reads / reads_full)
    kw = dict(wl_params)
    kw["coverage"] = wl_params["coverage"] * frac
    t0 = time.perf_counter()
    r = oracle.simulate(forest, make_params(**kw), n_threads=threads)
    wall = time.perf_counter() - t0
    tm = r["timing"]
    workers = min(threads, forest.n_chr) if threads > 1 and forest.n_chr >= 2 else 1
    loop_s = tm["read_loop_cpu_s"] / workers                      # wall-clock of the read loop on `workers` cores
    fixed_s = (tm["genomes_cpu_s"] + tm["fixed_cpu_s"]) / workers
    reads = r["n_reads"]
    full_s = fixed_s + loop_s * reads_full / max(1, reads)
    return {"reads": int(reads), "wall_s": wall, "loop_gbases_per_s": reads * R / loop_s / 1e9,
            "wall_gbases_per_s": reads * R / wall / 1e9, "full_job_reads": int(reads_full),
            "full_job_extrapolated_s": full_s, "full_job_extrapolated_gbases_per_s": reads_full * R / full_s / 1e9,
            "cores": workers,
            "sample": (f"all {forest.n_chr} chromosomes in proportion, {n_out} samples, coverage {kw['coverage']:.3f}x of "
                       f"{wl_params['coverage']:g}x: {reads} reads measured in {wall:.1f} s on {workers} thread(s) "
                       f"(read loop {loop_s:.1f} s, explicit genomes + per-base vectors {fixed_s:.1f} s); value = read "
                       f"loop only; full job ({reads_full} reads) extrapolated linearly: {full_s:.0f} s")}


def run_reference(args, forest, wl_params):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # every step is a bounded sample; the timed steps together hold >= 2e8 reads
    per_step = max(2.5e7, 2.0e8 / max(1, args.steps))
    vals, walls, samples = [], [], []
    for i in range(args.warmup + args.steps):
        s = cpu_sample(forest, wl_params, per_step, threads)
        if i >= args.warmup:
            vals.append(s["loop_gbases_per_s"])
            walls.append(s["wall_s"])
            samples.append(s)
    val = float(np.mean(vals))
    last = samples[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(walls)) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload_config(args, wl_params),
        "note": "CPU oracle (restatement of the RACES@1142937 semantics, not RACES itself: the reference cannot be built "
                "here), bounded samples of the workload",
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": last["cores"], "kind": "port",
                         "sample": f"{args.steps} timed steps x [{last['sample']}]",
                         "wall_gbases_per_s": float(np.mean([s["wall_gbases_per_s"] for s in samples])),
                         "full_job_extrapolated_gbases_per_s": last["full_job_extrapolated_gbases_per_s"],
                         "reads_measured": int(sum(s["reads"] for s in samples))},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    """the ONE JSON line of the contract, on the process's original stdout"""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    # Libraries write to fd 1 behind Python's back (NCCL prints its version there when NCCL_DEBUG is set):
    # everything but the JSON line goes to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-reads", type=float, default=2.0e8, help="reads of the one-thread CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sequencer", default=None, choices=["errorless", "constant", "random"],
                    help="override the workload's sequencer model (profiling)")
    ap.add_argument("--error-rate", type=float, default=1e-3)
    ap.add_argument("--insert-size", type=int, default=0, help="paired-end reads with this mean insert (sd 10)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: 'peer' = every rank's sampler flushes into rank 0's tables over NVLink peer memory "
                         "(reduction fused into the kernel), 'nccl' = local tables + NCCL reduce")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    cfg, scale, wl_params, _ = WORKLOADS[args.workload]
    wl_params = dict(wl_params)
    if args.sequencer is not None:
        wl_params.update(sequencer=args.sequencer, error_rate=0.0 if args.sequencer == "errorless" else args.error_rate)
    if args.insert_size:
        wl_params.update(insert_size_mean=args.insert_size, insert_size_stddev=10)
    seq_name = wl_params.get("sequencer", "errorless")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference" and rank != 0:
        return
    forest = synth_forest(config_spec(cfg, seed=0, scale=scale))
    if args.impl == "reference":
        if args.workload == "C5":
            emit({"impl": "reference", "unavailable": "C5: explicit per-cell genomes of 1e5 cells x 1e6 SNVs do not fit "
                  "in memory -- the CPU oracle (like the reference's own data model) cannot hold this forest"})
            return
        run_reference(args, forest, wl_params)
        return

    import torch
    import torch.distributed as dist
    from process_b200 import _lib as L

    torch.cuda.set_device(local)
    cores = os.cpu_count() or 1
    # the ranks share the host's cores: split them, or every rank's flattener oversubscribes the box
    os.environ.setdefault("PCS_HOST_THREADS", str(max(1, min(32, cores // world))))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # one explicit stream for everything: the library launches on it, torch events are recorded on it and
    # NCCL orders its collectives against it (torch's default stream has handle 0, which pcs_create reads
    # as "make your own stream" -- that would leave the sampler and the collectives unordered)
    stream = torch.cuda.Stream()
    side = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = L.Context(local, stream.cuda_stream)
    dev = L.Forest(ctx, forest)
    P = make_params(shard_rank=rank, shard_count=world, **wl_params)
    plan = L.Plan(dev, P)
    S, M, Lc = plan.info.n_out_samples, plan.info.n_mut, plan.info.n_loci
    occ = torch.zeros((S, M), dtype=torch.int32, device="cuda")
    cov = torch.zeros((S, M), dtype=torch.int32, device="cuda")

    # what the shards count on their own (untimed, with checksums): the reduced tables must add up to exactly this
    st0 = plan.run_device(occ.data_ptr(), cov.data_ptr())
    expect = torch.tensor([float(st0.sum_occurrences), float(st0.sum_depth)], dtype=torch.float64, device="cuda")
    rank_kernel_ms = torch.tensor([st0.kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(expect, op=dist.ReduceOp.SUM)
        all_kernel_ms = [torch.zeros_like(rank_kernel_ms) for _ in range(world)]
        dist.all_gather(all_kernel_ms, rank_kernel_ms)
        all_kernel_ms = [float(t.item()) for t in all_kernel_ms]
    else:
        all_kernel_ms = [float(st0.kernel_ms)]

    exchange = args.exchange if world > 1 else "none"
    ring = None
    words = S * Lc + S * M
    if exchange == "peer":
        # two table sets (depth + occurrences) on rank 0, mapped by every other rank through CUDA IPC
        handles = [None, None]
        try:
            if rank == 0:
                owned = [ctx.shared_alloc(words) for _ in range(2)]
                handles = [h for _, h in owned]
            dist.broadcast_object_list(handles, src=0)
            ring = [p for p, _ in owned] if rank == 0 else [ctx.shared_open(h) for h in handles]
            ok = torch.ones(1, device="cuda")
        except L.PcsError as e:  # no peer access between these GPUs
            print(f"rank {rank}: peer tables unavailable ({e}); using NCCL reduce", file=sys.stderr)
            ok = torch.zeros(1, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            exchange, ring = "nccl", None
        elif rank == 0:
            ctx.memset_u32(ring[0], words)
            ctx.memset_u32(ring[1], words)
            torch.cuda.synchronize()
        dist.barrier()

    step_no = [0]
    token = torch.zeros(1, dtype=torch.int32, device="cuda")

    def step():
        """one pass of the hot path, queued on the stream: no host synchronisation inside the timed region"""
        if exchange == "peer":
            b = step_no[0] & 1
            step_no[0] += 1
            plan.accumulate(ring[b], ring[b] + 4 * S * Lc, wait=False)
            if rank == 0:
                # the side stream's work of the step before (below) is complete before rank 0 joins this step's
                # collective -- and no rank's next sampler, which adds into the table set zeroed there, starts
                # before the collective is through
                stream.wait_stream(side)
            # every rank's flush has landed in rank 0's tables once this (on-stream) collective is through
            dist.all_reduce(token)
            if rank == 0:
                # coverage gather of this step and zeroing of its table set for the step after next: on a side
                # stream, beside the next step's sampler instead of in front of it
                side.wait_stream(stream)
                plan.finalize_on(ring[b], cov.data_ptr(), side.cuda_stream)
                ctx.memset_u32(ring[b], words, stream=side.cuda_stream)
            return
        plan.run_device(occ.data_ptr(), cov.data_ptr(), checksums=False, wait=False)
        if world > 1:  # 'nccl': sum the per-sample count tables on rank 0
            dist.reduce(occ, dst=0, op=dist.ReduceOp.SUM)
            dist.reduce(cov, dst=0, op=dist.ReduceOp.SUM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    plan.counters()  # reset what the warm-up accumulated
    clocks = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t1 = time.time()
    timed = plan.counters()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    reads = torch.tensor([float(timed.n_reads)], dtype=torch.float64, device="cuda")
    launches = torch.tensor([float(timed.kernel_launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(reads, op=dist.ReduceOp.SUM)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    total_ms = float(ms.item())
    total_reads = float(reads.item())
    clk = clocks.stop(t0, t1) if clocks else None
    R = plan.info.read_size
    value = total_reads * R / (total_ms * 1e-3) / 1e9

    # one more step, synchronous and with checksums (untimed): the reduced tables must hold exactly what the ranks
    # counted on their own
    tables_ok = True
    if exchange == "peer":
        b = step_no[0] & 1
        step_no[0] += 1
        barrier()
        plan.accumulate(ring[b], ring[b] + 4 * S * Lc)
        barrier()
        if rank == 0:
            fin = plan.finalize(ring[b], ring[b] + 4 * S * Lc, cov.data_ptr())
            tables_ok = fin.sum_occurrences == int(expect[0].item()) and fin.sum_depth == int(expect[1].item())
    else:
        plan.run_device(occ.data_ptr(), cov.data_ptr())
        if world > 1:
            dist.reduce(occ, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            tables_ok = int(occ.sum(dtype=torch.int64).item()) == int(expect[0].item())
    barrier()

    # ---- roofline of the sampler kernel on this rank
    kernel_ms = float(np.mean([plan.run_device(occ.data_ptr(), cov.data_ptr(), checksums=False).kernel_ms for _ in range(3)]))
    if world > 1:  # every rank's sampler kernel, warm: their spread is the imbalance of the shards
        mine = torch.tensor([kernel_ms], dtype=torch.float64, device="cuda")
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        all_kernel_ms = [float(t.item()) for t in gathered]
    else:
        all_kernel_ms = [kernel_ms]
    reads_per_step = total_reads / args.steps
    rank_reads = float(timed.n_reads) / args.steps
    kbar = float(expect[1].item()) / max(1.0, reads_per_step)   # job-wide: every shard sees the same mix
    kalt = float(expect[0].item()) / max(1.0, reads_per_step)
    b_read = 24.0 + 20.0 * kbar + 8.0 * kalt
    achieved = rank_reads * b_read / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = peaks()
    cnt = ncu_counters(args.workload, seq_name, args.insert_size) if world == 1 else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": cnt["dram_bytes_per_launch"] if cnt else None, "peak_source": peak_src,
                "kernel": "pcs::sample_tiles_staged_kernel", "kernel_ms": kernel_ms, "reads_per_launch": int(rank_reads),
                "bytes_per_read": b_read, "k_bar": kbar, "k_alt": kalt, "algorithmic_bytes_per_launch": rank_reads * b_read,
                "note": "MODEL bytes (SURVEY.md 8d: 24 B of read descriptor + 20 B per covered locus + 8 B per occurrence), "
                        "not a physical bound: the kernel keeps reads in registers and stages loci in shared memory, so its "
                        "DRAM traffic is a fraction of a percent of the model; what bounds it is `issue`"}
    if cnt:
        sm_hz = (clk["sm_mhz"] if clk and clk.get("sm_mhz") else 1965.0) * 1e6
        peak_issue = 148 * 4 * sm_hz          # warp instructions per second: 148 SMs x 4 schedulers x SM clock
        inst = cnt["warp_instructions_per_launch"]
        roofline["issue"] = {
            "bound": "warp-instruction issue slots (148 SMs x 4 schedulers x SM clock)",
            "achieved_ginst_per_s": inst / (kernel_ms * 1e-3) / 1e9, "peak_ginst_per_s": peak_issue / 1e9,
            "frac": inst / (kernel_ms * 1e-3) / peak_issue, "warp_instructions_per_32_reads": inst * 32 / max(1.0, rank_reads),
            "l2_bytes_per_launch": cnt.get("l2_bytes_per_launch"),
            "l2_frac_of_peak": (cnt["l2_bytes_per_launch"] / (kernel_ms * 1e-3) / cnt["l2_peak_bytes_per_s"]
                                if cnt.get("l2_bytes_per_launch") and cnt.get("l2_peak_bytes_per_s") else None),
            "shared_atomics_per_launch": cnt.get("shared_atomics_per_launch"),
            "shared_atomic_frac_of_peak": (cnt["shared_atomics_per_launch"] / (kernel_ms * 1e-3) / (148 * sm_hz)
                                           if cnt.get("shared_atomics_per_launch") else None),
            "global_reds_per_launch": cnt.get("global_reds_per_launch"),
            "dram_frac_of_peak": cnt["dram_bytes_per_launch"] / (kernel_ms * 1e-3) / (peak * 1e9),
            "source": "instruction / byte counts per launch: ncu capture of this workload (" + str(cnt.get("capture")) +
                      "); time: measured live",
            "stale": cnt["stale"],
            "stale_note": ("the kernel sources changed since the ncu capture: counts are from the older build"
                           if cnt["stale"] else None)}

    # ---- end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(args, forest, wl_params, ctx, L, world, rank, local, cores, R, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload != "C5":
        s = cpu_sample(forest, wl_params, args.cpu_reads, 1)
        cpu = {"value": s["loop_gbases_per_s"], "unit": UNIT, "cores": 1, "kind": "port",
               "sample": s["sample"] + "; one thread: the reference is single-threaded",
               "wall_gbases_per_s": s["wall_gbases_per_s"],
               "full_job_extrapolated_gbases_per_s": s["full_job_extrapolated_gbases_per_s"],
               "full_job_extrapolated_s": s["full_job_extrapolated_s"], "reads_measured": s["reads"]}
    elif rank == 0 and world == 1 and args.workload == "C5":
        cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "none: explicit per-cell genomes of 1e5 cells x 1e6 SNVs do not fit in memory"}

    if rank == 0:
        info = dev.info()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(args, wl_params),
            "detail": {"samples": S, "rows": M, "reads_per_step": reads_per_step,
                       "tiles_this_rank": int(plan.info.n_tiles), "parallelism": f"tile-sharded x{world}",
                       "exchange": {"none": "single GPU", "peer": "sampler flush adds into rank 0's tables over NVLink "
                                    "peer memory (CUDA IPC); one on-stream collective per step, no host synchronisation",
                                    "nccl": "local tables + NCCL reduce to rank 0"}[exchange],
                       "sampler_kernel_ms_per_rank": all_kernel_ms,
                       "l2": f"working set {(info['device_bytes'] + 3 * S * M * 4) / 1e6:.0f} MB > 126 MB L2; "
                             "count tables re-zeroed every step; no explicit flush"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches.item()),
            "checks": {"reduced_tables_equal_sum_of_rank_counts": bool(tables_ok)},
            "clocks": clk,
        }
        emit(line)
    if ring is not None:  # mappings first, then the owner frees
        barrier()
        if rank != 0:
            for p_ in ring:
                ctx.shared_close(p_)
        barrier()
        if rank == 0:
            for p_ in ring:
                ctx.shared_free(p_)
    plan.close()
    dev.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def measure_e2e(args, forest, wl_params, ctx, L, world, rank, local, cores, R, barrier):
    """The call a user makes, with host buffers, every step: pcs_forest_upload (flatten + H2D) + pcs_simulate
    (plan, kernels, D2H of the tables).  N > 1: ONE process drives the N GPUs (pcs_forest_replicate +
    pcs_simulate_multi, what a single-threaded R session does): rank 0 runs it on all devices of the job while
    the other ranks wait without spinning."""
    import torch.distributed as dist
    e2e_steps = max(1, min(args.steps, 3))
    P1 = lambda: make_params(**wl_params)
    if world > 1:
        store = dist.distributed_c10d._get_default_store()
        barrier()
        if rank != 0:
            store.wait(["pcs_e2e_done"])  # blocks on the socket: the rank's core is free for rank 0's host threads
            barrier()
            return None
    os.environ["PCS_HOST_THREADS"] = str(max(1, min(32, cores)))
    devices = list(range(world)) if world > 1 else [local]
    ctxs = [ctx] + [L.Context(d) for d in devices[1:]]
    out = {"unit": UNIT}
    h2d = d2h = 0
    reads = 0
    for it in range(1 + e2e_steps):  # one untimed pass first: it pins the library's staging memory
        if it == 1:
            t_e = time.perf_counter()
            h2d = d2h = 0
            reads = 0
        d0 = L.Forest(ctxs[0], forest)
        if world == 1:
            o, c, s2 = d0.simulate(P1())
            h2d += forest.host_bytes() + s2.h2d_bytes
            d2h += s2.d2h_bytes
        else:
            reps = [d0] + [L.replicate(d0, cx) for cx in ctxs[1:]]
            o, c, s2 = L.simulate_multi(reps, P1())
            h2d += forest.host_bytes() + s2.h2d_bytes
            d2h += s2.d2h_bytes
            for r_ in reps[1:]:
                r_.close()
        reads += s2.n_reads
        d0.close()
    dt = time.perf_counter() - t_e
    out.update({"value": reads * R / dt / 1e9, "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
                "ms_per_step": dt * 1e3 / e2e_steps,
                "h2d_note": "bytes of the forest description the call reads from host memory + the plan's tables; fewer "
                            "cross the link: the flattened tables are 3-13 bytes per row and the 16-byte instances are "
                            "built on the device",
                "what": ("pcs_forest_upload (flatten + H2D) + pcs_simulate (plan, kernels, D2H of the tables), host buffers"
                         if world == 1 else
                         f"one process, {world} GPUs: pcs_forest_upload + pcs_forest_replicate x{world - 1} + pcs_simulate_multi "
                         "(plan, kernels on every GPU flushing into GPU 0 over NVLink, D2H of the tables), host buffers")})
    try:
        # later calls on the same forest (other coverage, purity, sequencer): the forest stays on the device(s)
        d0 = L.Forest(ctxs[0], forest)
        reps = [d0] + [L.replicate(d0, cx) for cx in ctxs[1:]]
        run = (lambda: d0.simulate(P1())) if world == 1 else (lambda: L.simulate_multi(reps, P1()))
        run()
        t_w = time.perf_counter()
        warm_reads = 0
        for _ in range(e2e_steps):
            warm_reads += run()[2].n_reads
        dt_w = time.perf_counter() - t_w
        out["forest_resident"] = {"value": warm_reads * R / dt_w / 1e9, "unit": UNIT, "ms_per_step": dt_w * 1e3 / e2e_steps,
                                  "what": "pcs_simulate on a forest uploaded by an earlier call"}
        if world == 1:
            # the data frame's rows assembled on the device (pcs_simulate_result + pcs_result_fetch): only the
            # active rows' columns (occurrences, coverage, VAF) cross the link
            def result_pass(with_vaf):
                res, s3 = d0.simulate_result(P1(), with_vaf=with_vaf)
                res.fetch()
                nb, nr = res.d2h_bytes, res.n_rows
                res.close()
                return s3, nb, nr
            out["device_result"] = {"unit": UNIT, "what": "pcs_simulate_result + pcs_result_fetch on a resident forest: "
                                    "active rows compacted (and VAF computed) on the device"}
            for name, with_vaf in (("with_vaf", True), ("without_vaf", False)):
                result_pass(with_vaf)  # untimed: sizes the pinned staging area
                t_r = time.perf_counter()
                for _ in range(e2e_steps):
                    s3, nb, nr = result_pass(with_vaf)
                dt_r = time.perf_counter() - t_r
                out["device_result"][name] = {"ms_per_step": dt_r * 1e3 / e2e_steps,
                                              "value": s3.n_reads * R * e2e_steps / dt_r / 1e9,
                                              "active_rows": int(nr), "d2h_bytes_per_step": int(nb)}
        for r_ in reps[1:]:
            r_.close()
        d0.close()
    except Exception as ex:  # never lose the line over the extra figures
        out["forest_resident"] = {"error": str(ex)}
    if world == 1:
        try:
            # the reference-facing call itself, data frame included (what an R user waits for)
            from process_b200 import api
            forest.reference_path = forest.reference_path or __file__
            kw = dict(coverage=wl_params["coverage"], purity=wl_params["purity"], read_size=wl_params["read_size"],
                      insert_size_mean=wl_params.get("insert_size_mean", 0), seed=0, device=local)
            sq = wl_params.get("sequencer", "errorless")
            if sq != "errorless":
                kw["sequencer"] = api.BasicIlluminaSequencer(wl_params["error_rate"], sq == "random")
            t_a = time.perf_counter()
            api.simulate_seq(forest, **kw)
            first = time.perf_counter() - t_a
            t_a = time.perf_counter()
            for _ in range(e2e_steps):
                r = api.simulate_seq(forest, **kw)
            dt_a = (time.perf_counter() - t_a) / e2e_steps
            out["api_ms"] = dt_a * 1e3
            out["api"] = {"ms_per_call": dt_a * 1e3, "first_call_ms": first * 1e3, "rows": int(len(r["mutations"])),
                          "what": "process_b200.api.simulate_seq() -> {mutations: DataFrame, parameters}: resident forest, "
                                  "device result assembly, native column builders; first_call_ms adds flatten + upload + "
                                  "the per-forest string codes"}
            api.release_device_cache()
        except Exception as ex:
            out["api"] = {"error": str(ex)}
    for cx in ctxs[1:]:
        cx.close()
    if world > 1:
        store.set("pcs_e2e_done", "1")
        barrier()
    return out


if __name__ == "__main__":
    main()
