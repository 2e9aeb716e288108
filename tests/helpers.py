"""Shared test drivers: run any engine (CPU oracle or the CUDA engine -- same method names) over a FASTQ slab
following the reference's block / sync schedule (fqsqueezer_b200/schedule.py)."""
import os

import numpy as np

from fqsqueezer_b200 import schedule as S
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


def run_se(engine, slab, kind=0, max_blocks=None):
    """Returns the concatenated record stream (sync markers included where the engine emits them are dropped)."""
    off, ln, roff, rsz = S.parse_fastq(slab)
    out = []
    for gen, (f, l) in enumerate(S.split_blocks(rsz)):
        if max_blocks is not None and gen >= max_blocks:
            break
        ns = S.calc_no_synchronizations(gen, l - f, 1)
        engine.block_start()
        for a, b in S.segments(f, l, ns):
            recs, dup = engine.segment(slab, off[a:b], ln[a:b], kind)
            out.append(recs)
            engine.sync()
    recs = np.concatenate(out) if out else np.zeros(0, O.REC_DTYPE)
    return recs[recs["pos"] != O.POS_SYNC]


def run_se_workers(group, slab, n_workers, kind=0):
    """The reference at -t T (application.cpp:575-671): every block is partitioned among the workers (reads_block.h:197-214),
    all of them run the same number of sync segments and meet at every sync.  `group.workers[i]` codes worker i's reads,
    `group.sync()` is the exchange + table update of all workers.  Returns one record stream per worker."""
    off, ln, roff, rsz = S.parse_fastq(slab)
    out = [[] for _ in range(n_workers)]
    for gen, (f, l) in enumerate(S.split_blocks(rsz)):
        ns = S.calc_no_synchronizations(gen, l - f, n_workers)
        ranges = S.partition_for_workers(l - f, n_workers)
        segs = [list(S.segments(f + a, f + b, ns)) for a, b in ranges]
        assert len({len(x) for x in segs}) == 1, "workers disagree on the number of syncs"
        for w in group.workers:
            w.block_start()
        for q in range(len(segs[0])):
            for wi, w in enumerate(group.workers):
                a, b = segs[wi][q]
                recs, dup = w.segment(slab, off[a:b], ln[a:b], kind)
                out[wi].append(recs)
            group.sync()
    res = []
    for wi in range(n_workers):
        r = np.concatenate(out[wi]) if out[wi] else np.zeros(0, O.REC_DTYPE)
        res.append(r[r["pos"] < 0xFFFFFFF0])
    return res


def run_async(engine, slab, paired=False):
    """Same schedule through fqsk_submit / fqsk_collect, two segments in flight: segment n + 1 is submitted before n is collected
    (the host-side coder would consume n meanwhile).  Returns (records, dup flags per read[, pair_info])."""
    off, ln, roff, rsz = S.parse_fastq(slab)
    out, dups, info = [], [], []
    pend = None

    def take(t):
        recs, dup, rec_off = engine.collect(t)
        assert int(rec_off[len(dup)]) == len(recs)
        out.append(recs.copy()); dups.append(dup.copy())
        if paired:
            info.append(engine.pair_info(len(dup) // 2))      # decisions of the segment just collected

    for gen, (f, l) in enumerate(S.split_blocks(rsz, paired=paired)):
        ns = S.calc_no_synchronizations(gen, l - f, 1)
        engine.block_start()
        for a, b in S.segments(f, l, ns, paired=paired):
            t = engine.submit(slab, off[a:b], ln[a:b])
            if pend is not None:
                take(pend)
            pend = t
    if pend is not None:
        take(pend)
    recs = np.concatenate(out) if out else np.zeros(0, O.REC_DTYPE)
    dup = np.concatenate(dups) if dups else np.zeros(0, np.uint8)
    if paired:
        return recs, dup, (np.concatenate(info) if info else np.zeros((0, 3), np.uint32))
    return recs, dup


def run_pe_workers(group, slab, n_workers):
    """Paired-end at -t T (application.cpp:1020-1331): blocks close on pair boundaries, PartitionForWorkers keeps pairs together,
    every worker steps through its pairs with `i >= next_synchro`.  Returns per worker (records, pair_info)."""
    off, ln, roff, rsz = S.parse_fastq(slab)
    out = [[] for _ in range(n_workers)]
    for gen, (f, l) in enumerate(S.split_blocks(rsz, paired=True)):
        segs = [S.worker_segments(f, l, gen, n_workers, w, paired=True) for w in range(n_workers)]
        assert len({len(x) for x in segs}) == 1, "workers disagree on the number of syncs"
        for w in group.workers:
            w.block_start()
        for q in range(len(segs[0])):
            for wi, w in enumerate(group.workers):
                a, b = segs[wi][q]
                recs, dup = w.segment(slab, off[a:b], ln[a:b], 3)
                out[wi].append(recs)
            group.sync()
    res = []
    for wi in range(n_workers):
        r = np.concatenate(out[wi]) if out[wi] else np.zeros(0, O.REC_DTYPE)
        pi = r[r["pos"] == POS_PAIR]
        res.append((r[r["pos"] < 0xFFFFFFF0], pi["c"][:, :3].astype(np.uint32)))
    return res


POS_PAIR = 0xFFFFFFFB    # per pair: c0 = a candidate list exists, c1 = minimizer id (0..14, 15 = none usable), c2 = its position in mate 2


def run_pe(engine, slab, is_gpu=False):
    """Paired-end driver (application.cpp:1020-1331): the slab holds the pairs interleaved (mate 1, mate 2, ...); blocks close
    on pair boundaries with twice the margin, syncs step by pairs (`i >= next_synchro`).  Returns (records, pair_info[n, 3])."""
    off, ln, roff, rsz = S.parse_fastq(slab)
    out, info = [], []
    for gen, (f, l) in enumerate(S.split_blocks(rsz, paired=True)):
        ns = S.calc_no_synchronizations(gen, l - f, 1)
        engine.block_start()
        for a, b in S.segments(f, l, ns, paired=True):
            recs, dup = engine.segment(slab, off[a:b], ln[a:b], 3)
            out.append(recs)
            if is_gpu:
                info.append(engine.pair_info((b - a) // 2))
            engine.sync()
    recs = np.concatenate(out) if out else np.zeros(0, O.REC_DTYPE)
    if not is_gpu:
        pi = recs[recs["pos"] == POS_PAIR]
        info = [pi["c"][:, :3].astype(np.uint32)]
    return recs[recs["pos"] < 0xFFFFFFF0], (np.concatenate(info) if info else np.zeros((0, 3), np.uint32))


def run_pe_sorted(engine, slab, is_gpu=False):
    """Paired-end in the reference's default order (-p -om s): preprocess_pe bins the pairs by the first 4 symbols of MATE 1 (N -> T,
    application.cpp:415-506), each bin is one pair of temp files read through CSortedFASTQFile (pairs sorted by mate 1, io.h:499-528), so
    reads_blocks never span bins while the block generation keeps counting (application.cpp:1048-1104).  `slab` holds the pairs
    interleaved in the order the reference coded them.  Returns (records, pair_info[n, 3], flags, difs); the (flag, dif) of
    compress_prefix_sorted exist for the first mate of every non-duplicate pair."""
    off, ln, roff, rsz = S.parse_fastq(slab)
    lut = np.full(256, 3, np.int64)
    lut[ord("A")], lut[ord("C")], lut[ord("G")], lut[ord("T")] = 0, 1, 2, 3
    o1 = off[0::2].astype(np.int64)
    first4 = np.stack([lut[slab[o1 + k]] for k in range(4)], axis=1)
    bins = first4[:, 0] * 64 + first4[:, 1] * 16 + first4[:, 2] * 4 + first4[:, 3]
    cuts = [0] + list(2 * (np.flatnonzero(np.diff(bins)) + 1)) + [len(off)]
    out, info, flags, difs = [], [], [], []
    gen = 0
    for a0, a1 in zip(cuts[:-1], cuts[1:]):
        for f, l in S.split_blocks(rsz[a0:a1], paired=True):
            f += a0; l += a0
            ns = S.calc_no_synchronizations(gen, l - f, 1)
            engine.block_start()
            for a, b in S.segments(f, l, ns, paired=True):
                recs, dup = engine.segment(slab, off[a:b], ln[a:b], 3)
                out.append(recs)
                if is_gpu:
                    info.append(engine.pair_info((b - a) // 2))
                    fl, df = engine.sorted_prefix(b - a)
                    keep = (dup == 0) & (np.arange(b - a) % 2 == 0)
                    flags.append(fl[keep]); difs.append(df[keep])
                engine.sync()
            gen += 1
    recs = np.concatenate(out) if out else np.zeros(0, O.REC_DTYPE)
    if not is_gpu:
        pi = recs[recs["pos"] == POS_PAIR]
        info = [pi["c"][:, :3].astype(np.uint32)]
        sp = recs[recs["pos"] == O.POS_SORTED]
        flags = [sp["c"][:, 0].astype(np.uint32)]
        difs = [sp["c"][:, 1].astype(np.uint64) | (sp["c"][:, 2].astype(np.uint64) << np.uint64(32))]
    cat = lambda v, dt, shape=(0,): np.concatenate(v) if v else np.zeros(shape, dt)
    return recs[recs["pos"] < 0xFFFFFFF0], cat(info, np.uint32, (0, 3)), cat(flags, np.uint32), cat(difs, np.uint64)


def golden_pe_sorted_expect(gold):
    recs, info = golden_pe_expect(gold)
    _, flags, difs = golden_sorted_expect(gold)
    return recs, info, flags, difs


def golden_pe_expect(gold):
    r = gold["recs"]
    pi = r[r["pos"] == POS_PAIR]
    return r[r["pos"] < 0xFFFFFFF0], pi["c"][:, :3].astype(np.uint32)


def assert_recs_equal(got, want):
    want = want[want["pos"] != O.POS_SYNC]
    assert len(got) == len(want), (len(got), len(want))
    if len(got) == 0:
        return
    a = np.ascontiguousarray(got).view(np.uint8).reshape(len(got), -1)
    b = np.ascontiguousarray(want).view(np.uint8).reshape(len(want), -1)
    bad = np.flatnonzero((a != b).any(axis=1))
    assert len(bad) == 0, f"{len(bad)} records differ, first at {bad[0]}: got {got[bad[0]]} want {want[bad[0]]}"


def assert_dump_equal(engine, gold, pairs=False):
    for which, nm in ((0, "siv"), (1, "smer"), (2, "bmer")) + (((3, "pair"),) if pairs else ()):
        k, v = engine.dump(which)
        assert np.array_equal(k, gold[nm + "_keys"]), nm
        assert np.array_equal(v, gold[nm + "_vals"]), nm
    st = engine.stats()
    assert st["siv_no_filled"] == int(gold["siv_no_filled"])
    assert st["siv_no_updates"] == int(gold["siv_no_updates"])


def run_sorted(engine, slab, is_gpu):
    """Sorted-order driver: the reference feeds one temp file per 4-symbol bin (N -> T) through the same block / sync
    loop (application.cpp:198-219, 349-412, 538-570), so blocks never span bins and the block generation keeps counting.
    `slab` must already be in the reference's processing order (its decoder writes reads in that order).
    Returns (records without markers, flags, difs) -- for the oracle the (flag, dif) values ride in 0xFFFFFFFC records."""
    off, ln, roff, rsz = S.parse_fastq(slab)
    lut = np.full(256, 3, np.int64)
    lut[ord("A")], lut[ord("C")], lut[ord("G")], lut[ord("T")] = 0, 1, 2, 3
    first4 = np.stack([lut[slab[off.astype(np.int64) + k]] for k in range(4)], axis=1)
    bins = first4[:, 0] * 64 + first4[:, 1] * 16 + first4[:, 2] * 4 + first4[:, 3]
    cuts = [0] + list(np.flatnonzero(np.diff(bins)) + 1) + [len(off)]
    out, flags, difs = [], [], []
    gen = 0
    for a0, a1 in zip(cuts[:-1], cuts[1:]):
        for f, l in S.split_blocks(rsz[a0:a1]):
            f += a0; l += a0
            ns = S.calc_no_synchronizations(gen, l - f, 1)
            engine.block_start()
            for a, b in S.segments(f, l, ns):
                recs, dup = engine.segment(slab, off[a:b], ln[a:b], 2)
                if is_gpu:
                    fl, df = engine.sorted_prefix(b - a)
                    keep = dup == 0
                    flags.append(fl[keep]); difs.append(df[keep])
                out.append(recs)
                engine.sync()
            gen += 1
    recs = np.concatenate(out) if out else np.zeros(0, O.REC_DTYPE)
    if not is_gpu:
        sp = recs[recs["pos"] == O.POS_SORTED]
        flags = [sp["c"][:, 0].astype(np.uint32)]
        difs = [sp["c"][:, 1].astype(np.uint64) | (sp["c"][:, 2].astype(np.uint64) << np.uint64(32))]
    recs = recs[recs["pos"] < 0xFFFFFFF0]
    return recs, np.concatenate(flags) if flags else np.zeros(0, np.uint32), np.concatenate(difs) if difs else np.zeros(0, np.uint64)


def golden_sorted_expect(gold):
    r = gold["recs"]
    sp = r[r["pos"] == O.POS_SORTED]
    flags = sp["c"][:, 0].astype(np.uint32)
    difs = sp["c"][:, 1].astype(np.uint64) | (sp["c"][:, 2].astype(np.uint64) << np.uint64(32))
    return r[r["pos"] < 0xFFFFFFF0], flags, difs


def sorted_blocks(slab, off, rsz, paired=False):
    """(generation, first, last) of every reads_block of a sorted-order run: one temp file per 4-symbol bin (of mate 1 when paired; N -> T)
    goes through the block loop, so blocks never span bins while the generation keeps counting (application.cpp:349-412, 415-506, 538-570)."""
    lut = np.full(256, 3, np.int64)
    lut[ord("A")], lut[ord("C")], lut[ord("G")], lut[ord("T")] = 0, 1, 2, 3
    step = 2 if paired else 1
    o1 = off[0::step].astype(np.int64)
    first4 = np.stack([lut[slab[o1 + k]] for k in range(4)], axis=1)
    bins = first4[:, 0] * 64 + first4[:, 1] * 16 + first4[:, 2] * 4 + first4[:, 3]
    cuts = [0] + list(step * (np.flatnonzero(np.diff(bins)) + 1)) + [len(off)]
    gen = 0
    for a0, a1 in zip(cuts[:-1], cuts[1:]):
        for f, l in S.split_blocks(rsz[a0:a1], paired=paired):
            yield gen, f + a0, l + a0
            gen += 1


def split_sorted_markers(recs, paired):
    """(records, flags, difs[, pair_info]) of an oracle / tap record stream of a sorted-order run."""
    sp = recs[recs["pos"] == O.POS_SORTED]
    flags = sp["c"][:, 0].astype(np.uint32)
    difs = sp["c"][:, 1].astype(np.uint64) | (sp["c"][:, 2].astype(np.uint64) << np.uint64(32))
    out = [recs[recs["pos"] < 0xFFFFFFF0], flags, difs]
    if paired:
        out.append(recs[recs["pos"] == POS_PAIR]["c"][:, :3].astype(np.uint32))
    return out


def run_sorted_workers(group, slab, n_workers, paired=False):
    """The reference's default order at -t T (-s / -p with -om s): the blocks of run_sorted / run_pe_sorted, each partitioned among the
    workers as in run_se_workers / run_pe_workers.  Returns per worker the raw record stream (markers included: split_sorted_markers)."""
    off, ln, roff, rsz = S.parse_fastq(slab)
    out = [[] for _ in range(n_workers)]
    for gen, f, l in sorted_blocks(slab, off, rsz, paired):
        segs = [S.worker_segments(f, l, gen, n_workers, w, paired=paired) for w in range(n_workers)]
        assert len({len(x) for x in segs}) == 1, "workers disagree on the number of syncs"
        for w in group.workers:
            w.block_start()
        for q in range(len(segs[0])):
            for wi, w in enumerate(group.workers):
                a, b = segs[wi][q]
                recs, dup = w.segment(slab, off[a:b], ln[a:b], 3 if paired else 2)
                out[wi].append(recs)
            group.sync()
    return [np.concatenate(x) if x else np.zeros(0, O.REC_DTYPE) for x in out]
