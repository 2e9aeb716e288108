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


def assert_recs_equal(got, want):
    want = want[want["pos"] != O.POS_SYNC]
    assert len(got) == len(want), (len(got), len(want))
    if len(got) == 0:
        return
    a = np.ascontiguousarray(got).view(np.uint8).reshape(len(got), -1)
    b = np.ascontiguousarray(want).view(np.uint8).reshape(len(want), -1)
    bad = np.flatnonzero((a != b).any(axis=1))
    assert len(bad) == 0, f"{len(bad)} records differ, first at {bad[0]}: got {got[bad[0]]} want {want[bad[0]]}"


def assert_dump_equal(engine, gold):
    for which, nm in ((0, "siv"), (1, "smer"), (2, "bmer")):
        k, v = engine.dump(which)
        assert np.array_equal(k, gold[nm + "_keys"]), nm
        assert np.array_equal(v, gold[nm + "_vals"]), nm
    st = engine.stats()
    assert st["siv_no_filled"] == int(gold["siv_no_filled"])
    assert st["siv_no_updates"] == int(gold["siv_no_updates"])
