#!/usr/bin/env python3
"""Golden per-segment checksums of the BASELINE config-2 job, produced by the REAL reference (tapped fqs-1.1, -t 1).

bench.py times the k-mer engine on the config-2 read stream (fqsqueezer_b200/synth.py: job_chunk) cut into reads_blocks and sync
segments exactly as the reference cuts the corresponding FASTQ.  The CPU oracle and the reference itself need minutes per block of
that stream (at 100 Mbp the tables start empty and nearly every base runs the rough searches), far too long for a bench run.  So
the reference runs ONCE, here: this script writes the first N blocks of the job as a FASTQ, runs oracle/_ref/fqs-1.1-tap on it at
-t 1 with the tap going to a FIFO, and keeps -- per sync segment -- the number of per-base records and their order-sensitive
checksum (fqsk_recs_checksum's formula, fqsqueezer_b200/engine.py: recs_checksum_host).  bench.py's `parity_check` compares the
device-side checksums of the same segments with these values: count vectors, levels, rough flags and cor_pos of every coded base
of the first N blocks of the very job it times, bit for bit, against the reference's own output.

    python oracle/make_bench_golden.py [N_BLOCKS]        -> tests/golden/bench_config2_ref_checksums.npz  (re-saved after every block)

TEST INFRASTRUCTURE: never imported by the product package."""
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:] = [x for x in sys.path if os.path.abspath(x or ".") != HERE]
sys.path.insert(0, ROOT)
from fqsqueezer_b200 import engine as E  # noqa: E402  (only recs_checksum_host and REC_DTYPE: host-side numpy)
from fqsqueezer_b200 import schedule as S  # noqa: E402
from fqsqueezer_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

GENOME, SEED, GS, L, JOB_READS = 100_000_000, 43, 100, 150, 10_000_000
OUT = os.path.join(ROOT, "tests", "golden", "bench_config2_ref_checksums.npz")
POS_SYNC = 0xFFFFFFFE


def main():
    n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    blocks = S.split_blocks(synth.fastq_record_sizes(1, JOB_READS, L))[:n_blocks]
    n_reads = blocks[-1][1]
    genome = synth.make_genome(GENOME, SEED)
    tmp = tempfile.mkdtemp(prefix="fqs_bench_golden_")
    fq, fifo = os.path.join(tmp, "job.fastq"), os.path.join(tmp, "tap.fifo")
    t0 = time.time()
    at = 0
    for c in range((n_reads + synth.JOB_CHUNK - 1) // synth.JOB_CHUNK):
        codes, err = synth.job_chunk(genome, c, L)
        take = min(synth.JOB_CHUNK, n_reads - at)
        synth.write_fastq(fq, codes[:take], np.zeros((take, L), bool), seed=1, first_id=at + 1, append=c > 0)
        at += take
    print(f"[bench golden] {n_reads} reads = blocks 0..{n_blocks - 1} of the job written in {time.time() - t0:.0f} s ({os.path.getsize(fq) / 1e6:.0f} MB)", flush=True)
    expect = []      # segments per block, as the schedule predicts them (cross-check of fqsqueezer_b200/schedule.py against the reference)
    for g, (f, l) in enumerate(blocks):
        expect.append(len(list(S.segments(f, l, S.calc_no_synchronizations(g, l - f, 1)))))
    os.mkfifo(fifo)
    seg_n, seg_sum = [], []

    def save(done_segments):
        # whole blocks only
        nb, acc = 0, 0
        while nb < len(expect) and acc + expect[nb] <= done_segments:
            acc += expect[nb]; nb += 1
        if nb:
            np.savez_compressed(OUT, seg_nrecs=np.array(seg_n[:acc], np.uint64), seg_sum=np.array(seg_sum[:acc], np.uint64), segs_per_block=np.array(expect[:nb], np.uint32),
                                block_first=np.array([b[0] for b in blocks[:nb]], np.uint64), block_last=np.array([b[1] for b in blocks[:nb]], np.uint64),
                                gs=np.int64(GS), genome=np.int64(GENOME), seed=np.int64(SEED), read_len=np.int64(L))
        return nb

    def reader():
        item = E.REC_DTYPE.itemsize
        buf = b""
        cur = []
        saved = 0
        with open(fifo, "rb") as f:
            while True:
                chunk = f.read(item * (1 << 18))
                if not chunk:
                    break
                buf += chunk
                n = len(buf) // item
                recs = np.frombuffer(buf[: n * item], dtype=E.REC_DTYPE)
                buf = buf[n * item:]
                syncs = np.flatnonzero(recs["pos"] == POS_SYNC)
                a = 0
                for sidx in syncs:
                    part = recs[a:sidx]
                    cur.append(part[part["pos"] < 0xFFFFFFF0])
                    seg = np.concatenate(cur) if len(cur) > 1 else cur[0]
                    seg_n.append(len(seg)); seg_sum.append(E.recs_checksum_host(seg))
                    cur = []
                    a = sidx + 1
                    nb = save(len(seg_n))
                    if nb > saved:
                        saved = nb
                        print(f"[bench golden] block {nb - 1} done at {time.time() - t0:.0f} s ({len(seg_n)} segments)", flush=True)
                part = recs[a:]
                cur.append(part[part["pos"] < 0xFFFFFFF0].copy())
        assert sum(len(x) for x in cur) == 0, "records after the last sync"

    th = threading.Thread(target=reader)
    th.start()
    r = subprocess.run([O.REF_TAP_BIN, "e", "-s", "-om", "o", "-qm", "o", "-im", "o", "-gs", str(GS), "-t", "1", "-v", "0", "-out", os.path.join(tmp, "job.fqs"), fq],
                       env=dict(os.environ, FQS_TAP=fifo), capture_output=True, text=True, cwd=tmp)
    th.join()
    assert r.returncode == 0, r.stderr[-500:]
    assert len(seg_n) == sum(expect), (len(seg_n), sum(expect))
    save(len(seg_n))
    print(f"[bench golden] {len(seg_n)} segments, {sum(seg_n)} records, reference {r.stdout.strip()[-40:]}; total {time.time() - t0:.0f} s -> {OUT}", flush=True)
    for x in (fq, fifo, os.path.join(tmp, "job.fqs")):
        if os.path.exists(x):
            os.remove(x)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
