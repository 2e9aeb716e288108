#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REAL reference (tapped fqs-1.1 built by oracle/build_ref.py).

Run in the build container (needs /root/reference for build_ref.py; afterwards only oracle/_ref binaries).
Each fixture holds: the FASTQ bytes, the reference options, the per-base tap records, the sync positions and the
final table dumps.  Also asserts that the tapped binary's .fqs equals the untapped one and that it decodes.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:] = [x for x in sys.path if os.path.abspath(x or '.') != HERE]
sys.path.insert(0, ROOT)
from fqsqueezer_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def run_ref(fastq_path, gs, extra, tmp, threads=1):
    plain = os.path.join(tmp, "plain.fqs")
    tapd = os.path.join(tmp, "tap.fqs")
    tap = os.path.join(tmp, "tap.bin")
    dump = os.path.join(tmp, "dump.bin")
    base = ["e", "-s", "-qm", "o", "-im", "o", "-t", str(threads), "-gs", str(gs), "-v", "0", *extra]
    subprocess.run([O.REF_BIN, *base, "-out", plain, fastq_path], check=True, cwd=tmp, stdout=subprocess.DEVNULL)
    env = dict(os.environ, FQS_TAP=tap, FQS_TAP_DUMP=dump, FQS_TAP_CTX=os.path.join(tmp, "ctx.bin"))
    subprocess.run([O.REF_TAP_BIN, *base, "-out", tapd, fastq_path], check=True, cwd=tmp, env=env, stdout=subprocess.DEVNULL)
    assert open(plain, "rb").read() == open(tapd, "rb").read(), "tap changed the .fqs bytes"
    dec = os.path.join(tmp, "dec.fastq")
    subprocess.run([O.REF_BIN, "d", "-out", dec, plain], check=True, cwd=tmp, stdout=subprocess.DEVNULL)
    recs = np.fromfile(tap, dtype=O.REC_DTYPE)
    if threads > 1:     # one tap file per worker thread ($FQS_TAP, $FQS_TAP.t1, ...)
        recs = [recs] + [np.fromfile(tap + ".t%d" % i, dtype=O.REC_DTYPE) for i in range(1, threads)]
    d = np.fromfile(dump, dtype="<u8").reshape(-1, 3)
    return recs, d, open(plain, "rb").read(), open(dec, "rb").read()


def mixed_unc_reads(seed):
    """Reads that reach the two branches of find_counts no random genome reaches at fixture size, -gs 1 lengths (b = 19):
    * `mixed` (dna.cpp:470-478): two 120 bp haplotypes that differ in ONE base, covered ~6 000 x each -- the two sibling b-mers that end
      at that base both run into the 6-bit counter's ceiling (63 = ~3 200 occurrences), so the b-mer lookup shows two saturated counters;
    * the bmer_unc revert (dna.cpp:697-705): per locus, reads A = S CTX x1 (they END at x1), reads C = CTX[1:] x2 W (they never hold the
      b-mer CTX x2) and, after a sync, ONE read B = S CTX x2 W: at x2 the counts say x1 (> 3) and x2 is unseen -> repair_kmers_existing puts
      x1 into the corrected registers; at the next base the corrected context CTX[1:] x1 is in no table (the A reads end there) while the
      uncorrected CTX[1:] x2 is (from the C reads) -> counts_level bmer_unc -> the corrected registers are dropped."""
    rng = np.random.default_rng(seed)
    L = 60
    hap = rng.integers(0, 4, 120, dtype=np.uint8)
    hap2 = hap.copy(); hap2[60] = (hap2[60] + 1 + rng.integers(0, 3)) & 3
    n_main = 16000
    which = rng.integers(0, 2, n_main)
    starts = rng.integers(0, 120 - L + 1, n_main)
    strand = rng.integers(0, 2, n_main).astype(bool)
    idx = starts[:, None] + np.arange(L)[None, :]
    codes = np.where(which[:, None] == 0, hap[idx], hap2[idx]).astype(np.uint8)
    codes = np.where(strand[:, None], (3 - codes)[:, ::-1], codes).astype(np.uint8)
    err = rng.random(codes.shape) < 0.004
    codes = np.where(err, (codes + rng.integers(1, 4, codes.shape, dtype=np.uint8)) & 3, codes).astype(np.uint8)
    reads = [codes[i] for i in range(n_main)]
    errs = [err[i] for i in range(n_main)]
    head, late = [], []
    for locus in range(6):
        S, CTX, W = rng.integers(0, 4, 22, dtype=np.uint8), rng.integers(0, 4, 18, dtype=np.uint8), rng.integers(0, 4, 30, dtype=np.uint8)
        x1 = np.uint8(rng.integers(0, 4)); x2 = np.uint8((x1 + 1 + rng.integers(0, 3)) & 3)
        A = np.concatenate([S, CTX, [x1]]).astype(np.uint8)
        Cr = np.concatenate([CTX[1:], [x2], W]).astype(np.uint8)
        B = np.concatenate([S, CTX, [x2], W]).astype(np.uint8)
        for _ in range(6):
            head += [A, Cr]
        late.append(B)
    out = head + reads[:400]
    for k, B in enumerate(late):            # one B per locus, each after at least one sync (block 0 syncs every ~160 reads)
        out += reads[400 + 300 * k: 400 + 300 * (k + 1)] + [B]
    out += reads[400 + 300 * len(late):]
    e_out = [np.zeros(len(r), bool) for r in out]
    return out, e_out


def write_fastq_ragged(path, reads, errs, seed):
    """synth.write_fastq for reads of unequal length (same record format)."""
    rng = np.random.default_rng(seed + 3)
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    with open(path, "wb") as f:
        for i, (c, e) in enumerate(zip(reads, errs)):
            q = np.where(rng.random(len(c)) < 0.1, ord("F"), ord("I")).astype(np.uint8)
            q = np.where(e, ord("#"), q).astype(np.uint8)
            f.write(b"@SIM.%d %d/1\n" % (i + 1, i + 1) + lut[c].tobytes() + b"\n+\n" + q.tobytes() + b"\n")


def make_case(name, G, n_reads, L, gs, seed, n_frac=0.0, dup_frac=0.0, extra=("-om", "o"), repeats=False, threads=1, custom=None, keep_fqs=False, keep_ctx=False):
    genome = synth.make_genome(G, seed)
    if repeats:
        # low-complexity stretches (homopolymers, di-/tri-nucleotide repeats, a tandem duplication): k-mers that occur far more
        # than 8 times inside one sync segment, so the thread-local counters run on the cinc_lb / cinc_ls PRNG streams
        rng = np.random.default_rng(seed + 99)
        for a in range(200, G - 400, 900):
            kind = rng.integers(0, 4)
            unit = [np.array([rng.integers(0, 4)]), rng.integers(0, 4, 2), rng.integers(0, 4, 3), rng.integers(0, 4, 31)][kind]
            ln = int(rng.integers(60, 220))
            genome[a:a + ln] = np.resize(unit, ln)
    if custom is None:
        codes, err = synth.make_reads(genome, n_reads, L=L, seed=seed, n_frac=n_frac, dup_frac=dup_frac)
    with tempfile.TemporaryDirectory() as tmp:
        fq = os.path.join(tmp, "in.fastq")
        if custom is None:
            synth.write_fastq(fq, codes, err, seed=seed)
        else:
            write_fastq_ragged(fq, custom[0], custom[1], seed)
            n_reads = len(custom[0])
        recs, d, fqs, dec = run_ref(fq, gs, list(extra), tmp, threads)
        ctx_ids = np.fromfile(os.path.join(tmp, "ctx.bin"), dtype="<u8").reshape(-1, 8) if keep_ctx else None
        fastq = np.fromfile(fq, dtype=np.uint8)
        if tuple(extra) == ("-om", "o"):
            assert dec == fastq.tobytes(), "reference round trip failed"
        else:
            # sorted order: the decoder writes the reads in the order they were coded -> that IS the engine's input order
            assert sorted(dec.split(b"\n")[1::4]) == sorted(fastq.tobytes().split(b"\n")[1::4]), "reference round trip lost reads"
            fastq = np.frombuffer(dec, dtype=np.uint8).copy()
    dumps = {}
    for tag, nm in ((0, "siv"), (1, "smer"), (2, "bmer"), (3, "pair")):
        x = d[d[:, 0] == tag]
        o = np.lexsort((x[:, 2], x[:, 1]))
        dumps[nm + "_keys"] = x[o, 1]
        dumps[nm + "_vals"] = x[o, 2]
    stat = d[d[:, 0] == 4][0]
    out = os.path.join(GOLD, name + ".npz")
    if threads > 1:
        dumps.update({"recs_t%d" % i: r for i, r in enumerate(recs)})
        recs = recs[0]
    if keep_ctx:      # per base coded with counts: the 7 context ids + the coded rank, as the reference's coder saw them
        dumps["ctx_ids"] = ctx_ids
    if keep_fqs:      # the reference's .fqs itself: the live-host test on the GPU box compares with it instead of running a 45 GB reference there
        dumps["fqs"] = np.frombuffer(fqs, dtype=np.uint8)
    np.savez_compressed(out, fastq=fastq, gs=np.int64(gs), extra=np.array(list(extra)), recs=recs, threads=np.int64(threads),
                        siv_no_filled=stat[1], siv_no_updates=stat[2], fqs_size=np.int64(len(fqs)), **dumps)
    lv = np.bincount(recs[recs["pos"] < 0xFFFFFFF0]["level"], minlength=6)
    print(name, "reads", n_reads, "records", len(recs), "levels", lv.tolist(), "file", os.path.getsize(out))


def make_case_pe(name, G, n_pairs, L, gs, seed, threads=1, order="o", keep_fqs=False, n_frac=0.0):
    """Paired-end (-p), original order (-om o) or the reference's default, sorted order (-om s: pairs binned and sorted by mate 1,
    application.cpp:415-506, io.h:499-528; mate 1 goes through CompressSorted, dna.cpp:1793-1796): two FASTQ files; the fixture keeps them
    interleaved (mate 1, mate 2, ...), which is how the reference lays the pairs out inside a reads_block (reads_block.h:144-169) --
    in sorted order in the order the reference coded them (its decoder returns exactly that order)."""
    genome = synth.make_genome(G, seed)
    c1, e1, c2, e2 = synth.make_pairs(genome, n_pairs, L=L, seed=seed, ins_mean=2.2 * L, ins_sd=0.2 * L)
    if n_frac > 0:
        rng = np.random.default_rng(seed + 5)
        c1 = np.where(rng.random(c1.shape) < n_frac, 4, c1).astype(np.uint8)
        c2 = np.where(rng.random(c2.shape) < n_frac, 4, c2).astype(np.uint8)
        for i in np.flatnonzero(rng.random(n_pairs) < 0.01):       # a few duplicated first mates (the duplicate flag of a sorted pair)
            if i:
                c1[i] = c1[i - 1]; e1[i] = e1[i - 1]
    with tempfile.TemporaryDirectory() as tmp:
        f1, f2 = os.path.join(tmp, "in_1.fastq"), os.path.join(tmp, "in_2.fastq")
        synth.write_fastq(f1, c1, e1, mate=1, seed=seed)
        synth.write_fastq(f2, c2, e2, mate=2, seed=seed + 1)
        plain, tapd, tap, dump = (os.path.join(tmp, x) for x in ("plain.fqs", "tap.fqs", "tap.bin", "dump.bin"))
        base = ["e", "-p", "-om", order, "-qm", "o", "-im", "o", "-t", str(threads), "-gs", str(gs), "-v", "0"]
        subprocess.run([O.REF_BIN, *base, "-out", plain, f1, f2], check=True, cwd=tmp, stdout=subprocess.DEVNULL)
        subprocess.run([O.REF_TAP_BIN, *base, "-out", tapd, f1, f2], check=True, cwd=tmp, env=dict(os.environ, FQS_TAP=tap, FQS_TAP_DUMP=dump), stdout=subprocess.DEVNULL)
        assert open(plain, "rb").read() == open(tapd, "rb").read(), "tap changed the .fqs bytes"
        d1, d2 = os.path.join(tmp, "d1.fastq"), os.path.join(tmp, "d2.fastq")
        subprocess.run([O.REF_BIN, "d", "-out", d1, "-out2", d2, plain], check=True, cwd=tmp, stdout=subprocess.DEVNULL)
        if order == "o":
            assert open(d1, "rb").read() == open(f1, "rb").read() and open(d2, "rb").read() == open(f2, "rb").read(), "reference PE round trip failed"
        else:
            def pairs_of(a, b):
                x, y = open(a, "rb").read().split(b"\n"), open(b, "rb").read().split(b"\n")
                return sorted((x[i + 1], y[i + 1]) for i in range(0, len(x) - 1, 4))
            assert pairs_of(d1, d2) == pairs_of(f1, f2), "reference PE round trip lost pairs"
            f1, f2 = d1, d2            # the order the reference coded the pairs in
        recs = np.fromfile(tap, dtype=O.REC_DTYPE)
        per_thread = [recs] + [np.fromfile(tap + ".t%d" % i, dtype=O.REC_DTYPE) for i in range(1, threads)]   # one tap file per worker thread
        d = np.fromfile(dump, dtype="<u8").reshape(-1, 3)
        r1 = open(f1, "rb").read().split(b"\n")
        r2 = open(f2, "rb").read().split(b"\n")
        inter = []
        for i in range(n_pairs):
            inter += r1[4 * i:4 * i + 4] + r2[4 * i:4 * i + 4]
        fastq = np.frombuffer(b"\n".join(inter) + b"\n", dtype=np.uint8).copy()
        fqs_size = os.path.getsize(plain)
        fqs_bytes = open(plain, "rb").read()
    dumps = {}
    for tag, nm in ((0, "siv"), (1, "smer"), (2, "bmer"), (3, "pair")):
        x = d[d[:, 0] == tag]
        o = np.lexsort((x[:, 2], x[:, 1]))
        dumps[nm + "_keys"] = x[o, 1]
        dumps[nm + "_vals"] = x[o, 2]
    stat = d[d[:, 0] == 4][0]
    out = os.path.join(GOLD, name + ".npz")
    if threads > 1:
        dumps.update({"recs_t%d" % i: r for i, r in enumerate(per_thread)})
    if keep_fqs:
        dumps["fqs"] = np.frombuffer(fqs_bytes, dtype=np.uint8)
    np.savez_compressed(out, fastq=fastq, gs=np.int64(gs), extra=np.array(["-p", "-om", order]), recs=recs, threads=np.int64(threads),
                        siv_no_filled=stat[1], siv_no_updates=stat[2], fqs_size=np.int64(fqs_size), **dumps)
    pi = recs[recs["pos"] == 0xFFFFFFFB]
    print(name, "pairs", n_pairs, "records", len(recs), "pairs with a minimizer hit", int((pi["c"][:, 0] == 1).sum()), "coded from a minimizer", int(((pi["c"][:, 0] == 1) & (pi["c"][:, 1] < 15)).sum()),
          "file", os.path.getsize(out))


def main():
    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[1:]
    global make_case
    if only:
        _mk = make_case
        make_case = lambda name, **kw: _mk(name, **kw) if name in only else None
    # SE original order, tiny k (gs 1: prefix 9, p14/s17/b19), high coverage so that repairs, rough searches,
    # probabilistic counters (> 7) and the avg_filling_factor >= 7 gate all fire; Ns and duplicate reads included.
    make_case("se_orig_gs1", G=6000, n_reads=1500, L=80, gs=1, seed=7, n_frac=0.002, dup_frac=0.01)
    # same options as BASELINE config 2 (-gs 100: prefix 12, p17/s20/b24), small input
    make_case("se_orig_gs100", G=20000, n_reads=1200, L=100, gs=100, seed=43)
    # repeats: thread-local counters above the deterministic range (cinc_lb / cinc_ls draws) inside a segment
    make_case("se_orig_repeats_gs1", G=8000, n_reads=1600, L=90, gs=1, seed=51, n_frac=0.001, repeats=True)
    # sorted order (-om s): bins by 4-symbol prefix, std::sort inside a bin, sorted-prefix coding (flag / dif) + suffix from p_len
    make_case("se_sorted_gs1", G=5000, n_reads=2500, L=70, gs=1, seed=45, n_frac=0.002, dup_frac=0.01, extra=("-om", "s"))
    # two / three worker threads (-t 2, -t 3): per-worker PRNG streams, owner-routed exchange rows, global gate statistics
    make_case("se_orig_gs1_t2", G=6000, n_reads=1800, L=80, gs=1, seed=61, n_frac=0.002, dup_frac=0.01, threads=2)
    make_case("se_orig_gs16_t3", G=9000, n_reads=1500, L=100, gs=16, seed=62, threads=3)
    # paired end, original order: pair table, minimizer candidates, mate 2 coded from a shared minimizer (forward + reversed part)
    if not only or "pe_orig_gs1" in only:
        make_case_pe("pe_orig_gs1", G=5000, n_pairs=1200, L=80, gs=1, seed=71)
    # paired end at -t 2: pair-table owners by (fmix64(key) >> 48) % T (dna.cpp:1076-1081), pair triples in the exchange matrix
    if not only or "pe_orig_gs1_t2" in only:
        make_case_pe("pe_orig_gs1_t2", G=5000, n_pairs=1400, L=80, gs=1, seed=72, threads=2)
    # the reference's DEFAULT k-mer lengths (-gs 3100: prefix 13, p18/s21/b27; BASELINE configs 1, 4, 5): front-truncated b-mer lookups
    # with up to 5 missing symbols (1 364 trials per read start), a 16 GiB p-mer array, prefix sums over 4^(p - prefix - 1) fields.
    # The reference needs 45 GB and a minute per run at this setting; its .fqs rides in the fixture for the live-host test on the GPU box.
    make_case("se_orig_gs3100", G=30000, n_reads=3000, L=150, gs=3100, seed=81, n_frac=0.001, dup_frac=0.005, keep_fqs=True)
    if not only or "pe_orig_gs3100" in only:
        make_case_pe("pe_orig_gs3100", G=30000, n_pairs=1400, L=150, gs=3100, seed=82, keep_fqs=True)
    # `mixed` (two saturated 6-bit counters in one context) and the bmer_unc revert: crafted reads, see mixed_unc_reads
    if not only or "se_mixed_unc_gs1" in only:
        make_case("se_mixed_unc_gs1", G=1000, n_reads=0, L=60, gs=1, seed=91, custom=mixed_unc_reads(91))
    # device-side context ids (row f1): the 7 ids of determine_ctx_codes + the coded rank per base, tapped from the reference's coder
    make_case("se_ctx_gs1", G=4000, n_reads=700, L=80, gs=1, seed=95, n_frac=0.004, dup_frac=0.01, keep_ctx=True)
    # paired end in the reference's default order (-p with -om s): mate 1 through the sorted prefix, bins and sort by mate 1
    if not only or "pe_sorted_gs1" in only:
        make_case_pe("pe_sorted_gs1", G=5000, n_pairs=1500, L=80, gs=1, seed=73, order="s", n_frac=0.002)
    # the reference's default order at -t 2 (sorted bins, every reads_block split between two workers): SE and PE
    make_case("se_sorted_gs1_t2", G=5000, n_reads=2600, L=70, gs=1, seed=46, n_frac=0.002, dup_frac=0.01, extra=("-om", "s"), threads=2)
    if not only or "pe_sorted_gs1_t2" in only:
        make_case_pe("pe_sorted_gs1_t2", G=5000, n_pairs=1500, L=80, gs=1, seed=74, order="s", n_frac=0.002, threads=2)


if __name__ == "__main__":
    main()
