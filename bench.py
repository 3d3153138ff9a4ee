#!/usr/bin/env python
"""bench.py -- the abyss-bloom-dbg hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R]

One "step" = one complete pass of the hot path over the synthetic read set:
    pass 1  ntHash of every k-mer of every read + ordered counting-Bloom insert
    pass 2  classify reads, Bloom-backed unitig extension, ordered replay -> unitig FASTA
on the workload BASELINE.json's metric is quoted on (configs[1]): 50 M x 150 bp synthetic
paired reads (187.5 Mbp uniform-random genome, 40x, 0.5 % substitutions), k=64, kc=3, H=4,
-b 8 GiB (7 635 497 472 one-byte counters + 954 437 184 B assembled-k-mer bit filter), 1 x B200.
k-mers/s = (sum over reads of len-k+1) / step time, counted once per input k-mer.

`value`  : reads resident in HBM when the timed region starts (abb_*_dev entry points).
`e2e`    : same job through the host-buffer C-ABI calls (abb_insert_reads +
           abb_assembler_process_reads): pinned host reads are copied to the device inside the
           timed region in both passes, the unitigs come back to the host.
`roofline`: the Bloom-insert commit kernel (k_commit): algorithmic bytes (64*H B of 32-byte
           sectors per inserted k-mer + L/(L-k+1) B of read bases) / its summed
           CUDA-event launch time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline`: the UNMODIFIED reference (oracle/_ref/abyss-bloom-dbg-ref, built by
           oracle/Makefile) on a bounded sample of the same reads with all host threads.
`--impl reference` times that reference binary as the step (bounded sample per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, KC, H, L = 64, 3, 4, 150
BLOOM_BYTES = 8 << 30
N_READS = 50_000_000
GENOME = 187_500_000
ERR = 0.005
SEED = 2
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "abyss-bloom-dbg-ref")
ALG_BYTES_PER_KMER = 64 * H + L / (L - K + 1)  # 257.7 B
# dram__bytes_read.sum + dram__bytes_write.sum of one k_commit launch over a 2^19-slot window of this filter
# (ncu --set full, cold caches; profiles/r01_k_commit_w19_full_raw.csv): 387.9 MB + 87.0 MB
NCU_TRAFFIC_PER_COMMIT_LAUNCH = 474.9e6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.stop_flag, self.index = [], False, index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        self.t.join(timeout=3)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples if len(s) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(float(self.samples[0][1])), "reasons": reasons}


def write_sample_fastq(rs, n, path):
    rs.write_fastq(path, 0, n)


def run_reference(fq, threads, out_fa):
    """abyss-bloom-dbg (unmodified reference) wall time on fq with `threads` OpenMP threads"""
    cmd = f"ulimit -s 65536; exec {REF_BIN} -k{K} --kc={KC} -b{BLOOM_BYTES} -H{H} -j{threads} {fq} > {out_fa}"
    t0 = time.perf_counter()
    r = subprocess.run(["bash", "-c", cmd], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference run failed: " + r.stderr[-500:])
    return dt


def reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation on this box's host cores"""
    if rank != 0:
        return
    from abyss_b200.synth import ReadSet
    cores = os.cpu_count() or 1
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/abyss-bloom-dbg-ref not built (make -C oracle ref)"}))
        return
    sample = args.ref_reads
    rs = ReadSet(SEED, GENOME, N_READS, L, ERR, paired=True)
    tmp = tempfile.mkdtemp(prefix="abyss_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fq = os.path.join(tmp, "sample.fq")
    write_sample_fastq(rs, sample, fq)
    kmers = sample * (L - K + 1)
    times = []
    for i in range(args.warmup + args.steps):
        dt = run_reference(fq, cores, os.path.join(tmp, "ref.fa"))
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = kmers / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": "k-mers/sec (Bloom insert + unitig extend)", "value": val, "unit": "k-mers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, extra={"sample_reads": sample}),
        "cpu_baseline": {"value": val, "unit": "k-mers/s", "cores": cores, "kind": "reference",
                         "sample": f"first {sample} reads of the workload, abyss-bloom-dbg -j{cores}, files on tmpfs"},
        "e2e": {"value": val, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, extra=None):
    c = {"workload": f"{args.reads} x {L} bp synthetic paired reads, {GENOME} bp random genome, err {ERR}, k={K} kc={KC} H={H} -b 8GiB "
                     "(BASELINE.json configs[1])",
         "reads": args.reads, "read_len": L, "k": K, "kc": KC, "num_hashes": H, "bloom_bytes": BLOOM_BYTES,
         "l2_policy": "inputs (7.5 GB reads) and filters (8.6 GB) far exceed the 126 MB L2; no flush needed",
         "parallelism": (f"pass 1 sharded by k-mer hash range over {args.gpus} GPUs (NCCL all-to-all + all-reduce max), read classification "
                         "sharded by reads (all-gather), rest of pass 2 replicated"
                         if args.gpus > 1 else "1 GPU")}
    if extra:
        c.update(extra)
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=N_READS, help="number of reads of the workload (default: the BASELINE config)")
    ap.add_argument("--ref-reads", type=int, default=2_000_000, help="bounded CPU sample (reads)")
    ap.add_argument("--window", type=int, default=0, help="ordered-insert window in k-mer slots (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from abyss_b200 import build as abb_build
    if rank == 0:
        abb_build.build()
    from abyss_b200 import capi
    from abyss_b200.synth import ReadSet
    from abyss_b200.synth_torch import TorchReadSet

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    capi.load()
    dev = torch.device("cuda", local_rank)

    # ---- synthetic reads, generated on the device
    rs = ReadSet(SEED, GENOME, args.reads, L, ERR, paired=True)
    trs = TorchReadSet(rs, dev)
    bases = torch.empty(rs.n * L, dtype=torch.uint8, device=dev)
    step_reads = 1 << 21
    for s in range(0, rs.n, step_reads):
        e = min(rs.n, s + step_reads)
        bases[s * L:e * L] = trs.ascii(s, e).reshape(-1)
    offs = torch.arange(rs.n + 1, dtype=torch.int64, device=dev) * L
    del trs
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    n_kmers_expected = rs.n * (L - K + 1)
    counters = capi.counters_for_budget(BLOOM_BYTES)

    filt = capi.Filter.counting(counters, H, K, KC, device=local_rank)
    filt.set_profiling(True)
    if args.window:
        filt.set_window(args.window)
    ext = torch.cuda.ExternalStream(filt.stream(), device=dev)

    # N > 1: pass 1 is sharded by k-mer hash range (all-to-all of hashes, ordered insert of the owned
    # k-mers, all-reduce(max) union over NVLink); pass 2 runs replicated on every rank
    lo, up = rank * rs.n // world, (rank + 1) * rs.n // world
    offs_slice = torch.arange(up - lo + 1, dtype=torch.int64, device=dev) * L

    asm = capi.Assembler(filt)  # one handle for all steps: device buffers are allocated once, state is reset per step
    asm.raw_results = True      # the unitig sequences are copied to the host by the library; no Python string per unitig

    def one_step(host=None):
        """returns (n_kmers, contigs, assembler stats, insert stats)"""
        filt.clear()
        filt.stats(reset=True)
        asm.reset()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(ext)
        if world > 1:
            from abyss_b200 import multigpu
            multigpu.sharded_insert(filt, bases[lo * L:up * L], offs_slice, up - lo)
            nk = n_kmers_expected
            p1.record(ext)
            codes = multigpu.sharded_classify(asm, bases[lo * L:up * L], offs_slice, up - lo, rs.n)
            contigs = asm.process_reads_dev(bases.data_ptr(), offs.data_ptr(), rs.n)
            del codes
        elif host is None:
            nk = filt.insert_reads_dev(bases.data_ptr(), offs.data_ptr(), rs.n, bases.numel())
            p1.record(ext)
            contigs = asm.process_reads_dev(bases.data_ptr(), offs.data_ptr(), rs.n)
        else:
            nk = filt.insert_reads(host)
            p1.record(ext)
            contigs = asm.process_reads(host)
        torch.cuda.synchronize()
        ast, ist, cnt = asm.stats(), filt.stats(), asm.counters()
        ist.ms_pass1 = p0.elapsed_time(p1)
        return nk, contigs, ast, ist, cnt

    def timed(n_steps, host=None):
        res = []
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(n_steps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ev0.record(ext)
            out = one_step(host)
            ev1.record(ext)
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            res.append((ms, out))
        return res

    for _ in range(args.warmup):
        timed(1)
    with ClockSampler(local_rank) as clk:
        runs = timed(args.steps)
    ms_step = sum(r[0] for r in runs) / len(runs)
    nk, contigs, ast, ist, cnt = runs[-1][1]
    assert nk == n_kmers_expected, (nk, n_kmers_expected)
    total_kmers = nk  # one job, counted once (strong scaling)
    value = total_kmers / (ms_step * 1e-3)

    # ---- e2e through host buffers
    e2e = None
    if not args.no_e2e and world == 1:
        hb = torch.empty(bases.numel(), dtype=torch.uint8, pin_memory=True)
        hb.copy_(bases)
        ho = offs.cpu().numpy().astype(np.uint64)
        host = (hb.numpy(), ho)
        timed(1, host)
        eruns = timed(max(1, args.steps), host)
        ems = sum(r[0] for r in eruns) / len(eruns)
        econt = eruns[-1][1][1]
        d2h = sum(c[1] for c in econt) + 24 * len(econt) + rs.n  # unitig bases + records + per-read codes
        e2e = {"value": total_kmers / (ems * 1e-3), "unit": "k-mers/s", "ms_per_step": ems,
               "h2d_bytes_per_step": 2 * (int(hb.numel()) + int(ho.nbytes)), "d2h_bytes_per_step": int(d2h)}
        del hb

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    commit_ms = ist.ms_commit / max(1, ist.commit_launches)
    kmers_per_launch = ist.kmers / max(1, ist.commit_launches)  # k-mers this rank inserted per k_commit launch
    achieved = ALG_BYTES_PER_KMER * kmers_per_launch / (commit_ms * 1e-3) / 1e9 if commit_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_commit (ordered counting-Bloom min-increment)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC_PER_COMMIT_LAUNCH if (args.window in (0, 1 << 19) and args.reads == N_READS) else None,
                "traffic_note": "bytes per launch, ncu cold-cache capture profiles/r01_k_commit_w19_full_raw.csv; algorithmic bytes per launch = "
                                f"{ALG_BYTES_PER_KMER * kmers_per_launch:.3e}",
                "peak_source": peak_src,
                "alg_bytes_per_kmer": ALG_BYTES_PER_KMER, "launches": int(ist.commit_launches),
                "avg_launch_ms": commit_ms, "share_of_step": ist.ms_commit / ms_step}

    cpu = None
    if not args.no_cpu_baseline and os.path.exists(REF_BIN):
        cores = os.cpu_count() or 1
        tmp = tempfile.mkdtemp(prefix="abyss_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        fq = os.path.join(tmp, "sample.fq")
        write_sample_fastq(rs, min(args.ref_reads, rs.n), fq)
        dt = run_reference(fq, cores, os.path.join(tmp, "ref.fa"))
        sample = min(args.ref_reads, rs.n)
        cpu = {"value": sample * (L - K + 1) / dt, "unit": "k-mers/s", "cores": cores, "kind": "reference",
               "sample": f"first {sample} reads of the workload, abyss-bloom-dbg -j{cores} (unmodified reference), {dt:.1f} s wall"}

    line = {
        "metric": "k-mers/sec (Bloom insert + unitig extend)", "value": value, "unit": "k-mers/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args), "clocks": clk.summary(), "e2e": e2e,
        "gpu_launches": int(ist.launches + ast.launches), "roofline": roofline, "cpu_baseline": cpu,
        "phases_ms": {"hash": ist.ms_hash, "insert": ist.ms_insert, "classify": ast.ms_classify, "visited": ast.ms_visited,
                      "tiles": ast.ms_tiles, "extend": ast.ms_extend, "extend_walk": ast.ms_walk, "extend_stage": ast.ms_stage,
                      "extend_repeat_check": ast.ms_repeat, "replay": ast.ms_replay, "pass2_wall": ast.ms_total,
                      "candidate_list_host": ast.ms_cand},
        "pass1_ms": ist.ms_pass1, "insert_kmers_per_s": nk / (ist.ms_pass1 * 1e-3),
        "extend_kmers_per_s": nk / ((ast.ms_classify + ast.ms_tiles + ast.ms_visited + ast.ms_extend + ast.ms_replay) * 1e-3),
        "unitigs": int(cnt.contig_id), "bases_assembled": int(cnt.bases_assembled),
        "speculation": {"rounds": int(ast.rounds), "speculated": int(ast.speculated_reads), "wasted": int(ast.wasted_reads),
                        "markers": int(ast.markers), "tiles": int(ast.tiles), "serial_fallbacks": int(ast.serial_fallbacks)},
        "deferred_inserts": int(ist.deferred),
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
