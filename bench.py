#!/usr/bin/env python3
"""bench.py -- k-mer insert+lookup throughput of the B200 engine on BASELINE.json config 2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): synthetic 150 bp SE reads from a 100 Mbp random genome, 0.5 % substitutions,
`-gs 100` k-mer lengths (prefix 12, p17/s20/b24), original order, one reference worker (`-t 1`).
A STEP is one reads_block of that stream (16 MiB of FASTQ = 51 k reads, reads_block.h:121-169) pushed through the hot path
exactly as the reference schedules it: block g is cut into calc_no_synchronizations(g)+1 sync segments
(application.h:85-92) and every segment is one fqsk_segment + fqsk_sync.  Steps continue the same job, so the default
W + K = 196 blocks is the whole 10 M-read configuration.

value : bases/s with the reads already resident in HBM (fqsk_segment_device), timed with CUDA events on the engine's stream.
e2e   : the same steps through the host-buffer C-ABI call (fqsk_segment): H2D of the reads and D2H of every per-base record
        inside the timed region.
roofline / cpu_baseline: see DESIGN.md section 7.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from fqsqueezer_b200 import schedule as S  # noqa: E402
from fqsqueezer_b200 import synth  # noqa: E402

METRIC = "kmer_insert_lookup_bases_per_s"
UNIT = "bases/s"
L = 150
GENOME = 100_000_000
SEED = 43
GS = 100
READS_PER_BLOCK = 51_000            # 16 MiB / ~325 B per record, minus the 100 KiB margin (reads_block.h:25, 137)
B_ALG = 251.5                       # algorithmic bytes per base, SURVEY.md section 8d (config 2)
PHASE_ALG = {"lookup": 127 * 32 / 150.0, "sync_apply": (127 + 131) * 64 / 150.0, "sync_siv": 134 * 128 / 150.0}
PHASE_KERNEL = {"prep": "k_prep", "lookup": "k_lookup", "partial": "k_partial", "walk": "k_walk", "compact": "k_compact2+scans",
                "sort": "cub radix sort", "local": "k_local", "rough": "k_rough", "fold": "k_fold", "sync_locate": "k_locate_heads",
                "sync_apply": "k_apply_keys", "sync_siv": "k_siv_increment", "mt": "k_mt_extend"}


def workload_config(extra=None):
    c = {"workload": "BASELINE config 2: 150bp SE reads, 100 Mbp random genome, 0.5% subs, -gs 100 (p17/s20/b24), -om o, -t 1 sync schedule",
         "reads_per_step": READS_PER_BLOCK, "read_len": L, "l2": "tables (4 GiB p-mer array + GiB-scale b/s tables) are far larger than L2; no flush needed"}
    if extra:
        c.update(extra)
    return c


def block_codes(genome, g, rank=0):
    codes, _ = synth.make_reads(genome, READS_PER_BLOCK, L=L, seed=1000 * (rank + 1) + g)
    return codes


def codes_to_slab(codes):
    """DNA lines only ('ACGT...\\n'): ids and qualities never reach the k-mer path."""
    n = codes.shape[0]
    slab = np.empty((n, L + 1), np.uint8)
    slab[:, :L] = synth.codes_to_ascii(codes)
    slab[:, L] = 10
    off = (np.arange(n, dtype=np.uint64) * np.uint64(L + 1))
    ln = np.full(n, L, np.uint32)
    return slab.reshape(-1), off, ln


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = max(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows if len(r) > 3 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified fqs-1.1 compiled from /root/reference (oracle/_ref/fqs-1.1)
# ------------------------------------------------------------------------------------------------------------------
def run_reference(genome, n_reads, threads, seed_off=0):
    """Times the unmodified fqs-1.1 on a bounded sample.  Its 'Processing time' starts with the construction of the CPU tables
    (4 GiB p-mer array + 5 M sub-tables at -gs 100: seconds, independent of the input), which a 10 M-read job amortises and a
    bounded sample does not: the same binary is therefore also timed on an 8-read file and that start-up is reported and taken off."""
    from oracle import oracle as O          # checker side only: this function never touches the product path
    if not os.path.exists(O.REF_BIN):
        return None

    def one(codes, err, tmp, name):
        fq = os.path.join(tmp, name + ".fastq")
        nbytes = synth.write_fastq(fq, codes, err, seed=1)
        cmd = [O.REF_BIN, "e", "-s", "-om", "o", "-qm", "o", "-im", "o", "-gs", str(GS), "-t", str(threads), "-v", "0", "-out", os.path.join(tmp, name + ".fqs"), fq]
        t = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
        wall = time.time() - t
        m = re.search(r"Processing time:\s*([0-9.eE+-]+)", r.stdout + r.stderr)
        return (float(m.group(1)) if m else wall), nbytes

    codes, err = synth.make_reads(genome, n_reads, L=L, seed=777 + seed_off)
    with tempfile.TemporaryDirectory() as tmp:
        startup, _ = one(codes[:8], err[:8], tmp, "tiny")
        secs, nbytes = one(codes, err, tmp, "s")
    net = max(secs - startup, 0.05 * secs)
    return {"bases": n_reads * L, "seconds": net, "seconds_total": secs, "startup_seconds": startup, "fastq_bytes": nbytes}


def run_compress_e2e(genome, n_reads):
    """BASELINE.json's second metric, `e2e compress MB/s`: the reference compressor with its k-mer engine bound to libfqsk.so
    (host/_bin/fqs-1.1-fqsk: host/build_host.py + host/fqsk_live.h, the integration of INTEGRATION.md) against the unmodified
    fqs-1.1 at -t 1 -- the parity configuration -- on the same FASTQ: FASTQ bytes / 'Processing time', and whether the two .fqs
    files are byte-identical.  Bounded sample; everything outside the k-mer engine is the reference's own single-threaded host code."""
    from oracle import oracle as O          # only to locate the unmodified reference binary (the baseline of this leg)
    live = os.path.join(ROOT, "host", "_bin", "fqs-1.1-fqsk")
    lib = os.path.join(ROOT, "fqsqueezer_b200", "libfqsk.so")
    if not (os.path.exists(live) and os.path.exists(O.REF_BIN)):
        return {"unavailable": "host/_bin/fqs-1.1-fqsk or oracle/_ref/fqs-1.1 not built"}
    codes, err = synth.make_reads(genome, n_reads, L=L, seed=4242)
    with tempfile.TemporaryDirectory() as tmp:
        fq = os.path.join(tmp, "s.fastq")
        nbytes = synth.write_fastq(fq, codes, err, seed=1)
        base = ["e", "-s", "-om", "o", "-qm", "o", "-im", "o", "-gs", str(GS), "-t", "1", "-v", "0"]

        def one(exe, out, env=None):
            t = time.time()
            r = subprocess.run([exe, *base, "-out", out, fq], capture_output=True, text=True, cwd=tmp, env=env, timeout=300)
            wall = time.time() - t
            if r.returncode != 0:
                raise RuntimeError(f"{os.path.basename(exe)} exit {r.returncode}: {r.stderr[-300:]}")
            m = re.search(r"Processing time:\s*([0-9.eE+-]+)", r.stdout + r.stderr)
            return (float(m.group(1)) if m else wall), wall, r.stderr

        env = dict(os.environ, FQSK_LIB=lib, FQSK_VERBOSE="1")
        one(live, os.path.join(tmp, "w.fqs"), env)                     # warm-up: CUDA context, module load, page cache
        t_live, wall_live, log = one(live, os.path.join(tmp, "a.fqs"), env)
        t_ref, wall_ref, _ = one(O.REF_BIN, os.path.join(tmp, "b.fqs"))
        same = open(os.path.join(tmp, "a.fqs"), "rb").read() == open(os.path.join(tmp, "b.fqs"), "rb").read()
        fqs_bytes = os.path.getsize(os.path.join(tmp, "a.fqs"))
    m = re.search(r"([0-9.]+) s inside the engine calls, (\d+) kernel launches", log)
    return {"value": nbytes / 1e6 / t_live, "unit": "MB/s", "reference_t1": nbytes / 1e6 / t_ref, "speedup_vs_t1": t_ref / t_live,
            "byte_identical": bool(same), "fastq_bytes": nbytes, "fqs_bytes": fqs_bytes, "seconds": t_live, "reference_seconds": t_ref,
            "engine_call_seconds": float(m.group(1)) if m else None, "gpu_launches": int(m.group(2)) if m else None,
            "sample": f"{n_reads} reads of a config-2 stream ({nbytes / 1e6:.1f} MB FASTQ), e -s -om o -qm o -im o -gs {GS} -t 1: fqs-1.1-fqsk (reference host code, k-mer engine on the GPU through the C-ABI, blocking fqsk_segment + fqsk_sync) vs the unmodified fqs-1.1; 'Processing time' of each, start-up included (reference: construction of its CPU tables, several seconds; ours: CUDA context + HBM tables)"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = min(os.cpu_count() or 1, 64)
    genome = synth.make_genome(GENOME, SEED)
    per_step = max(200, 120_000 // max(args.steps, 1))     # K steps of a bounded sample: ~120 k reads in total (tens of seconds beyond the table construction)
    if args.warmup:
        run_reference(genome, min(per_step * args.warmup, 10_000), threads, seed_off=1)
    res = run_reference(genome, per_step * args.steps, threads)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/fqs-1.1 not built"}))
        return 0
    v = res["bases"] / res["seconds"]
    sample = (f"{per_step * args.steps} reads ({args.steps} steps x {per_step}) of the config-2 stream, fqs-1.1 e -s -om o -qm o -im o -gs {GS} -t {threads}: "
              f"'Processing time' {res['seconds_total']:.2f} s minus {res['startup_seconds']:.2f} s of table construction (same binary on an 8-read file)")
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * res["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "impl": "reference", "config": workload_config({"reads_per_step": per_step}),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------------
# sharded arm: ONE config-2 job over N GPUs, reference -t N semantics (tables sharded by owner key, rows routed at every sync)
# ------------------------------------------------------------------------------------------------------------------
def sharded_arm(args, rank, world, local_rank, dist, dev):
    import torch
    from fqsqueezer_b200 import engine as E
    from fqsqueezer_b200 import sharded

    pref, p, s, b = E.kmer_params(GS)
    genome = synth.make_genome(GENOME, SEED)
    n_blocks = args.warmup + args.steps
    eng = sharded.ShardedKmerEngine(p, s, b, pref, rank, world, device=local_rank, dist=dist, expected_kmers=(1 << 29) // world,
                                    reserve_reads=READS_PER_BLOCK // world + 16, reserve_bytes=(READS_PER_BLOCK // world + 16) * L)
    segs, d_blocks = [], []
    off_np = (np.arange(READS_PER_BLOCK, dtype=np.int64) * L)
    d_off = torch.from_numpy(off_np).to(dev)
    d_len = torch.full((READS_PER_BLOCK,), L, dtype=torch.int32, device=dev)
    for g in range(n_blocks):
        sg = S.worker_segments(0, READS_PER_BLOCK, g, world, rank)
        segs.append(sg)
        lo, hi = sg[0][0], sg[-1][1]
        codes = block_codes(genome, g, 0)[lo:hi]               # every rank cuts ITS slice out of the same block
        d_blocks.append((lo, torch.from_numpy(synth.codes_to_ascii(codes).reshape(-1)).to(dev)))
    torch.cuda.synchronize()

    def run_block(g):
        eng.block_start()
        lo, t = d_blocks[g]
        for a, bb in segs[g]:
            n = bb - a
            eng.segment_device(t.data_ptr() + (a - lo) * L, n * L, d_off.data_ptr(), d_len.data_ptr(), n, want_n_recs=False)
            eng.sync()

    for g in range(args.warmup):
        run_block(g)
    st0 = eng.stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    eng.timer_begin()
    t0 = time.time()
    for g in range(args.warmup, n_blocks):
        run_block(g)
    dev_ms = eng.timer_end()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    wall_ms = (time.time() - t0) * 1e3
    sampler.stop_flag = True
    st1 = eng.stats()
    ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = torch.tensor([st1["kernel_launches"] - st0["kernel_launches"]], dtype=torch.int64, device=dev)
    dist.all_reduce(launches)
    dev_ms_max = float(ms.item())
    bases = args.steps * READS_PER_BLOCK * L               # the whole job, all ranks together
    eng.close()
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = B_ALG * bases / (dev_ms_max / 1e3) / 1e9
        line = {"metric": METRIC, "value": bases / (dev_ms_max / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": workload_config({"parallelism": f"ONE job, tables hash-sharded over {world} GPUs by the reference's owner keys (-t {world} semantics): lookups through NVLink peer "
                                                          f"mappings, exchange rows by peer stores into the owners' inboxes, NCCL barrier + all-reduce per sync",
                                           "blocks": f"{args.warmup}..{n_blocks - 1} of the job", "segments_timed": (st1["n_segments"] - st0["n_segments"])}),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world), "traffic": None, "peak_source": peak_src,
                             "kernel": "whole step over all ranks, 251.5 algorithmic B/base"},
                "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches.item()), "clocks": sampler.summary(), "wall_ms_per_step": wall_ms / args.steps}
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=190)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", action="store_true", help="N > 1: ONE job with hash-sharded tables (reference -t N semantics, strong scaling) instead of N independent replicas")
    ap.add_argument("--no-phase-events", action="store_true", help="do not bracket the internal phases with CUDA events (fewer host API calls per segment)")
    ap.add_argument("--no-box-warmup", action="store_true", help="skip the throw-away engine that warms the box up before the W warm-up steps")
    ap.add_argument("--cpu-sample-reads", type=int, default=120_000, help="bounded sample of the cpu_baseline leg: large enough that the reference spends tens of seconds beyond its table construction")
    ap.add_argument("--no-compress-e2e", action="store_true", help="skip the whole-compressor leg (fqs-1.1-fqsk vs fqs-1.1 -t 1)")
    ap.add_argument("--compress-sample-reads", type=int, default=60_000)
    ap.add_argument("--trace-blocks", type=int, default=0, help="print per-block phase ms every N blocks to stderr")
    ap.add_argument("--profile-block", type=int, default=-1, help="cudaProfilerStart/Stop around this block (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    from fqsqueezer_b200 import engine as E

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the k-mer path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    if args.shard and world > 1:
        return sharded_arm(args, rank, world, local_rank, dist, dev)

    pref, p, s, b = E.kmer_params(GS)
    genome = synth.make_genome(GENOME, SEED)
    n_blocks = args.warmup + args.steps
    # every block's schedule: (#syncs, list of segments)
    sched = []
    for g in range(n_blocks):
        ns = S.calc_no_synchronizations(g, READS_PER_BLOCK, 1)
        sched.append(list(S.segments(0, READS_PER_BLOCK, ns)))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: reads resident in HBM ----------------
    # Timed pass: no per-phase event brackets (they cost ~25 % in host API calls on the small early segments).  The phase
    # shares come from a second, identical pass with the brackets on (below).
    eng = E.KmerEngine(p, s, b, pref, device=local_rank, expected_kmers=1 << 29, profile=bool(args.trace_blocks), reserve_reads=READS_PER_BLOCK, reserve_bytes=READS_PER_BLOCK * L)
    d_blocks = []
    off_np = (np.arange(READS_PER_BLOCK, dtype=np.int64) * L)
    d_off = torch.from_numpy(off_np).to(dev)                      # same offsets for every block
    d_len = torch.full((READS_PER_BLOCK,), L, dtype=torch.int32, device=dev)
    for g in range(n_blocks):
        codes = block_codes(genome, g, rank)
        d_blocks.append(torch.from_numpy(synth.codes_to_ascii(codes).reshape(-1)).to(dev))
    torch.cuda.synchronize()

    def run_block_device(g, eng):
        eng.block_start()
        base = d_blocks[g].data_ptr()
        for a, bb in sched[g]:
            n = bb - a
            eng.segment_device(base + a * L, n * L, d_off.data_ptr(), d_len.data_ptr(), n, want_n_recs=False)   # offsets are relative to the segment's first read
            eng.sync()

    # Box warm-up (untimed, besides the W warm-up steps below): a fresh box starts with idle clocks and a cold driver, and this
    # job is a chain of thousands of host-observed segments, so its first seconds run 20-30 % slower than the same blocks a few
    # seconds later (measured: 609 vs 750 Mbases/s for the whole job).  A throw-away engine runs the first blocks of the job once.
    if not args.no_box_warmup:
        scratch = E.KmerEngine(p, s, b, pref, device=local_rank, expected_kmers=1 << 29, reserve_reads=READS_PER_BLOCK, reserve_bytes=READS_PER_BLOCK * L)
        for g in range(min(24, n_blocks)):
            run_block_device(g, scratch)
        scratch.close()
        del scratch
    for g in range(args.warmup):
        run_block_device(g, eng)
    st0 = eng.stats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    eng.timer_begin()
    t_wall = time.time()
    prev_prof, t_prev = eng.profile(), time.time()
    for g in range(args.warmup, n_blocks):
        if g == args.profile_block:
            torch.cuda.cudart().cudaProfilerStart()
        run_block_device(g, eng)
        if g == args.profile_block:
            torch.cuda.cudart().cudaProfilerStop()
        if args.trace_blocks and (g % args.trace_blocks == 0 or g == n_blocks - 1):
            pr = eng.profile()
            print(f"block {g} ({len(sched[g])} segments): wall {1e3 * (time.time() - t_prev):.1f} ms " +
                  str({k: round(pr[k] - prev_prof[k], 2) for k in pr}) + " " +
                  str({k: v for k, v in eng.stats().items() if k in ("n_hot_segments", "n_replays", "bmer_buckets", "smer_buckets", "bmer_stash_used", "n_bmers", "n_smers", "draws_b")}),
                  file=sys.stderr, flush=True)
        if args.trace_blocks:
            prev_prof, t_prev = eng.profile(), time.time()
    dev_ms = eng.timer_end()
    barrier()
    wall_ms = (time.time() - t_wall) * 1e3
    sampler.stop_flag = True
    st1 = eng.stats()
    ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dev_ms_max = float(ms.item())
    bases_rank = args.steps * READS_PER_BLOCK * L
    value = world * bases_rank / (dev_ms_max / 1e3)
    n_seg = st1["n_segments"] - st0["n_segments"]
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    eng.close()

    # ---------------- phase shares: the same pass again with CUDA-event brackets around the internal phases ----------------
    phases = None
    if not args.no_phase_events:
        engp = E.KmerEngine(p, s, b, pref, device=local_rank, expected_kmers=1 << 29, profile=True, reserve_reads=READS_PER_BLOCK, reserve_bytes=READS_PER_BLOCK * L)
        for g in range(args.warmup):
            run_block_device(g, engp)
        prof0 = engp.profile()
        for g in range(args.warmup, n_blocks):
            run_block_device(g, engp)
        prof1 = engp.profile()
        phases = {k: prof1[k] - prof0[k] for k in prof1}
        engp.close()
    del d_blocks
    torch.cuda.empty_cache()

    # ---------------- e2e: host buffers through fqsk_segment ----------------
    e2e = None
    if not args.no_e2e:
        eng2 = E.KmerEngine(p, s, b, pref, device=local_rank, expected_kmers=1 << 29, reserve_reads=READS_PER_BLOCK, reserve_bytes=READS_PER_BLOCK * (L + 1))
        slabs = []
        for g in range(n_blocks):
            slabs.append(codes_to_slab(block_codes(genome, g, rank)))

        def run_block_host(g, pend):
            """fqsk_submit / fqsk_collect, two segments in flight: segment n + 1 is submitted before n is collected -- the point where
            the reference's host-side coder would consume the records of n."""
            slab, off, ln = slabs[g]
            eng2.block_start()
            nb = 0
            for a, bb in sched[g]:
                t = eng2.submit(slab, off[a:bb], ln[a:bb])
                if pend is not None:
                    recs, dup, _ = eng2.collect(pend)
                    nb += recs.nbytes + dup.nbytes
                pend = t
            return nb, pend

        pend = None
        for g in range(args.warmup):
            _, pend = run_block_host(g, pend)
        if pend is not None:
            eng2.collect(pend)
            pend = None
        barrier()
        t0 = time.time()
        d2h = 0
        for g in range(args.warmup, n_blocks):
            nb, pend = run_block_host(g, pend)
            d2h += nb
        if pend is not None:
            recs, dup, _ = eng2.collect(pend)
            d2h += recs.nbytes + dup.nbytes
        barrier()
        e2e_s = time.time() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * bases_rank / float(t.item()), "unit": UNIT,
               "h2d_bytes_per_step": READS_PER_BLOCK * (L + 12), "d2h_bytes_per_step": d2h // args.steps,
               "note": "fqsk_submit / fqsk_collect with host slab + read descriptors (H2D inside), every per-base record (28 B) copied back into page-locked host memory on a second stream, two segments in flight; wall clock incl. ctypes/numpy host code"}
        eng2.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---------------- roofline ----------------
    peak, peak_src = measured_peak()
    step_achieved = B_ALG * bases_rank / (dev_ms / 1e3) / 1e9
    roof = {"bound": "hbm", "achieved": step_achieved, "peak": peak, "unit": "GB/s", "frac": step_achieved / peak, "traffic": None,
            "peak_source": peak_src,
            "kernel": "whole step (all kernels of fqsk_segment + fqsk_sync), 251.5 algorithmic B/base",
            "traffic_note": "ncu of one steady-state step (profiles/r01d_steady_block_summary.md): 8.5 GB of DRAM traffic per 7.65 Mbase segment = 4.4x the algorithmic bytes; the average step of the job mixes 1..95 segments, so no single per-launch figure exists"}
    if phases:
        dom = max(phases, key=lambda k: phases[k])
        dom_share = phases[dom] / max(sum(phases.values()), 1e-9)
        dom_alg = PHASE_ALG.get(dom, 0.0) * bases_rank
        roof["dominant_kernel"] = {"name": PHASE_KERNEL.get(dom, dom), "share_of_device_time": dom_share, "ms_per_step": phases[dom] / args.steps,
                                   "algorithmic_GBps": (dom_alg / (phases[dom] / 1e3) / 1e9) if phases[dom] > 0 else None,
                                   "note": "phases without algorithmic bytes (walk, local, sort, fold, mt ...) are the cost of making the parallel order exact"}
        roof["phase_ms_per_step"] = {k: round(v / args.steps, 4) for k, v in phases.items()}
        roof["phase_note"] = "measured in a second identical pass with CUDA-event brackets on the engine's stream (the brackets are off in the timed pass)"

    # ---------------- cpu baseline (bounded sample, rank 0, N = 1 only) ----------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = min(os.cpu_count() or 1, 64)
        res = run_reference(genome, args.cpu_sample_reads, threads)
        if res is not None:
            cpu = {"value": res["bases"] / res["seconds"], "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"{args.cpu_sample_reads} reads of a config-2 stream ({res['bases'] / 1e6:.1f} Mbases), fqs-1.1 e -s -om o -qm o -im o -gs {GS} -t {threads}: 'Processing time' {res['seconds_total']:.2f} s minus {res['startup_seconds']:.2f} s of table construction (same binary on an 8-read file) = {res['seconds']:.2f} s (whole compressor, the k-mer engine is ~75% of it)"}

    # ---------------- whole compressor through the drop-in (bounded sample, rank 0, N = 1 only) ----------------
    compress = None
    if not args.no_compress_e2e and world == 1:
        try:
            compress = run_compress_e2e(genome, args.compress_sample_reads)
        except Exception as ex:              # never at the expense of the line itself
            compress = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": workload_config({"parallelism": "1 engine per GPU" + ("" if world == 1 else f" x {world} independent replicas (ONE job over hash-sharded tables: --shard)"),
                                       "blocks": f"{args.warmup}..{n_blocks - 1} of the job", "segments_timed": n_seg}),
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "compress_e2e": compress, "gpu_launches": launches,
            "clocks": sampler.summary(), "wall_ms_per_step": wall_ms / args.steps}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints its version
    # there when the first communicator is created) are sent to stderr, the JSON line goes to the saved descriptor
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    try:
        rc = main()
    finally:
        _real_stdout.flush()
    sys.exit(rc)
