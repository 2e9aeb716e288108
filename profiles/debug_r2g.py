"""Debug run of round 2 (under gpurun): (1) bench parity check of the first blocks with the side streams on / off, (2) per-segment wall
clock and phase brackets of an early block, (3) the compiled drop-in in paired-end sorted order with a launch trace when it hangs."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
from fqsqueezer_b200 import engine as E, schedule as S, synth

what = sys.argv[1:] or ["parity", "timing", "live"]
if os.environ.get("FQSK_LIB_PATH"):
    E.LIB_PATH = os.environ["FQSK_LIB_PATH"]
dev = torch.device("cuda", 0)
pref, p, s, b = E.kmer_params(B.GS)


def mk(flags=0, profile=False):
    return E.KmerEngine(p, s, b, pref, expected_kmers=1 << 29, reserve_reads=B.RESERVE_READS, reserve_bytes=B.RESERVE_READS * B.L, profile=profile, flags=flags)


if "parity" in what or "timing" in what:
    reads = B.JobReads(synth.make_genome(B.GENOME, B.SEED))
    blocks = B.job_blocks()
    z = np.load(B.GOLDEN_SUMS)

def run_blocks(eng, g0, g1):
    for g in range(g0, g1):
        f, l = blocks[g]
        t = torch.from_numpy(synth.codes_to_ascii(reads.codes(f, l)).reshape(-1)).to(dev)
        d_off = torch.arange(l - f, dtype=torch.int64, device=dev) * B.L
        d_len = torch.full((l - f,), B.L, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        eng.block_start()
        for a, bb in S.segments(0, l - f, S.calc_no_synchronizations(g, l - f, 1)):
            eng.segment_device(t.data_ptr() + a * B.L, (bb - a) * B.L, d_off.data_ptr(), d_len.data_ptr(), bb - a, want_n_recs=False)
            eng.sync()


if "parity" in what:
    NBLK = int(os.environ.get("NBLK", "3"))
    variants = (("default", 0), ("serial", E.F_SERIAL), ("profile", E.F_PROFILE), ("trace", E.F_TRACE_LAUNCH), ("dirty ff", 0), ("dirty a5", 0), ("after a whole engine", 0))
    pick = os.environ.get("VARIANTS", "default,serial").split(",")
    for name, flags in [v for v in variants if v[0] in pick]:
        if name.startswith("dirty"):      # the engine's allocations land on memory another allocation has left behind
            junk = torch.full((40 << 30,), 0xFF if name.endswith("ff") else 0xA5, dtype=torch.uint8, device=dev)
            torch.cuda.synchronize(); del junk; torch.cuda.empty_cache()
        if name.startswith("after"):
            e0 = mk(0); run_blocks(e0, 0, 4); run_blocks(e0, 96, 104); e0.close(); del e0
        eng = mk(flags)
        seg, bad, t0 = 0, [], time.time()
        for g in range(NBLK):
            f, l = blocks[g]
            t = torch.from_numpy(synth.codes_to_ascii(reads.codes(f, l)).reshape(-1)).to(dev)
            d_off = torch.arange(l - f, dtype=torch.int64, device=dev) * B.L
            d_len = torch.full((l - f,), B.L, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()      # the engine reads these on its own non-blocking stream
            eng.block_start()
            segs = list(S.segments(0, l - f, S.calc_no_synchronizations(g, l - f, 1)))
            for k, (a, bb) in enumerate(segs):
                eng.segment_device(t.data_ptr() + a * B.L, (bb - a) * B.L, d_off.data_ptr(), d_len.data_ptr(), bb - a, want_n_recs=False)
                if k + 1 < len(segs) and not os.environ.get("NO_ANNOUNCE"):
                    a2, b2 = segs[k + 1]
                    eng.announce_device(t.data_ptr() + a2 * B.L, (b2 - a2) * B.L, d_off.data_ptr(), d_len.data_ptr(), b2 - a2)
                cs, n = eng.recs_checksum()
                if n != int(z["seg_nrecs"][seg]) or cs != int(z["seg_sum"][seg]):
                    bad.append((seg, n, int(z["seg_nrecs"][seg]), bb - a))
                seg += 1
                eng.sync()
        st = eng.stats()
        print(f"[parity {name}] {seg} segments, {len(bad)} bad (seg, n, golden n, reads): {bad[:12]}  {time.time() - t0:.1f}s  replays {st['n_replays']} launches {st['kernel_launches']}", flush=True)
        eng.close()

if "timing" in what:
    for name, flags, prof in (("default", 0, False), ("serial", E.F_SERIAL, False), ("profile", 0, True)):
        eng = mk(flags, prof)
        for g in (0, 1, 2):
            f, l = blocks[g]
            t = torch.from_numpy(synth.codes_to_ascii(reads.codes(f, l)).reshape(-1)).to(dev)
            d_off = torch.arange(l - f, dtype=torch.int64, device=dev) * B.L
            d_len = torch.full((l - f,), B.L, dtype=torch.int32, device=dev)
            sched = list(S.segments(0, l - f, S.calc_no_synchronizations(g, l - f, 1)))
            torch.cuda.synchronize()
            p0 = eng.profile()
            eng.block_start()
            t0 = time.perf_counter()
            for k, (a, bb) in enumerate(sched):
                eng.segment_device(t.data_ptr() + a * B.L, (bb - a) * B.L, d_off.data_ptr(), d_len.data_ptr(), bb - a, want_n_recs=False)
                if k + 1 < len(sched) and not os.environ.get("NO_ANNOUNCE"):
                    a2, b2 = sched[k + 1]
                    eng.announce_device(t.data_ptr() + a2 * B.L, (b2 - a2) * B.L, d_off.data_ptr(), d_len.data_ptr(), b2 - a2)
                eng.sync()
            tot = time.perf_counter() - t0
            p1 = eng.profile()
            print(f"[timing {name}] block {g}: {len(sched)} segments, {1e3 * tot:.2f} ms, {1e6 * tot / len(sched):.0f} us / segment" +
                  (" " + str({k: round(p1[k] - p0[k], 2) for k in p1}) if prof else ""), flush=True)
        eng.close()

if "live" in what:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests import test_live_host as T
    from oracle import oracle as O
    for case in T.PAIRED_SORTED[:1]:
        gs, G, n, L, seed = case
        with tempfile.TemporaryDirectory() as tmp:
            files = list(T._fastq_pe(tmp, gs, G, n, L, seed))
            base = ["e", "-p", "-om", "s", "-qm", "o", "-im", "o", "-t", "1", "-gs", str(gs), "-v", "0"]
            plain = os.path.join(tmp, "plain.fqs")
            subprocess.run([O.REF_BIN, *base, "-out", plain, *files], check=True, cwd=tmp, stdout=subprocess.DEVNULL)
            for name, env in (("default", {}), ("ctx off", {"FQSK_CTX": "0"}), ("serial", {"FQSK_FLAGS": "16"}), ("trace", {"FQSK_FLAGS": "8"})):
                ours = os.path.join(tmp, "ours.fqs")
                if os.path.exists(ours):
                    os.unlink(ours)
                try:
                    r = subprocess.run([T.LIVE_BIN, *base, "-out", ours, *files], cwd=tmp, env=dict(os.environ, FQSK_LIB=T.REAL_LIB, FQSK_VERBOSE="1", **env),
                                       capture_output=True, text=True, timeout=90)
                    same = os.path.exists(ours) and open(ours, "rb").read() == open(plain, "rb").read()
                    print(f"[live {name}] rc {r.returncode}, identical {same}; stderr tail: {r.stderr[-400:]!r}", flush=True)
                except subprocess.TimeoutExpired as ex:
                    err = ex.stderr.decode() if isinstance(ex.stderr, bytes) else (ex.stderr or "")
                    print(f"[live {name}] TIMEOUT; stderr tail: {err[-1500:]!r}", flush=True)

if "timeline" in what:      # fork / join structure of three consecutive segments of an early block, side streams on
    reads = B.JobReads(synth.make_genome(B.GENOME, B.SEED))
    blocks = B.job_blocks()
    eng = mk(0)
    G = int(os.environ.get("TL_BLOCK", "10"))
    for g in range(G + 1):
        f, l = blocks[g]
        t = torch.from_numpy(synth.codes_to_ascii(reads.codes(f, l)).reshape(-1)).to(dev)
        d_off = torch.arange(l - f, dtype=torch.int64, device=dev) * B.L
        d_len = torch.full((l - f,), B.L, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        eng.block_start()
        segs = list(S.segments(0, l - f, S.calc_no_synchronizations(g, l - f, 1)))
        for k, (a, bb) in enumerate(segs):
            if g == G and k == int(os.environ.get("TL_K0", "40")):
                eng.timeline(True)
            eng.segment_device(t.data_ptr() + a * B.L, (bb - a) * B.L, d_off.data_ptr(), d_len.data_ptr(), bb - a, want_n_recs=False)
            if k + 1 < len(segs) and not os.environ.get("NO_ANNOUNCE"):
                a2, b2 = segs[k + 1]
                eng.announce_device(t.data_ptr() + a2 * B.L, (b2 - a2) * B.L, d_off.data_ptr(), d_len.data_ptr(), b2 - a2)
            eng.sync()
            if g == G and k == int(os.environ.get("TL_K1", "42")):
                eng.timeline(False)
    eng.close()

if "blocks" in what:      # wall clock per block over the early regime (tables fill up as in the job)
    reads = B.JobReads(synth.make_genome(B.GENOME, B.SEED))
    blocks = B.job_blocks()
    eng = mk(E.F_SERIAL if os.environ.get("SERIAL") else 0)
    for g in range(int(os.environ.get("NBLK", "104"))):
        f, l = blocks[g]
        t = torch.from_numpy(synth.codes_to_ascii(reads.codes(f, l)).reshape(-1)).to(dev)
        d_off = torch.arange(l - f, dtype=torch.int64, device=dev) * B.L
        d_len = torch.full((l - f,), B.L, dtype=torch.int32, device=dev)
        segs = list(S.segments(0, l - f, S.calc_no_synchronizations(g, l - f, 1)))
        torch.cuda.synchronize()
        st0 = eng.stats()
        eng.block_start()
        t0 = time.perf_counter()
        for k, (a, bb) in enumerate(segs):
            eng.segment_device(t.data_ptr() + a * B.L, (bb - a) * B.L, d_off.data_ptr(), d_len.data_ptr(), bb - a, want_n_recs=False)
            if k + 1 < len(segs) and not os.environ.get("NO_ANNOUNCE"):
                a2, b2 = segs[k + 1]
                eng.announce_device(t.data_ptr() + a2 * B.L, (b2 - a2) * B.L, d_off.data_ptr(), d_len.data_ptr(), b2 - a2)
            eng.sync()
        tot = time.perf_counter() - t0
        st1 = eng.stats()
        if g % 5 == 0 or g >= 95:
            print(f"[blocks] block {g}: {len(segs)} segments of {segs[0][1] - segs[0][0]} reads, {1e3 * tot:.2f} ms, {1e6 * tot / len(segs):.0f} us / segment, "
                  f"{(st1['kernel_launches'] - st0['kernel_launches']) / len(segs):.1f} launches / segment, {(st1['n_replays'] - st0['n_replays']) / len(segs):.2f} walks / segment, hot {st1['n_hot_segments'] - st0['n_hot_segments']}", flush=True)
    eng.close()
