"""Device-side context ids: include/fqsk_ctx.h restates cor_zone, determine_ctx_codes, rank and the
recent-rank history (dna.cpp:739-774, code_ctx.cpp:257-338) for the CUDA kernel and the reference-side binding.  Here the same
functions run on the CPU over the records tapped from the real reference and must reproduce, for every base coded with counts, the 7
context ids and the rank the reference's coder used (second tap of oracle/build_ref.py, fixture se_ctx_gs1)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from fqsqueezer_b200 import schedule as S
from oracle import oracle as O
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CTX_REC = np.dtype([("a", "<u8"), ("b", "<u8")])


def build_harness(tmp):
    so = os.path.join(str(tmp), "libctx_harness.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", os.path.join(ROOT, "tests", "ctx_harness.cpp"), "-I", os.path.join(ROOT, "include"), "-o", so], check=True)
    lib = C.CDLL(so)
    lib.ctx_from_tap.restype = C.c_uint64
    lib.ctx_from_tap.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
    return lib


def ctx_expected_from_tap(lib, g):
    """(ids[n, 8] computed by fqsk_ctx.h from the tapped records, raw 16-byte records per base)."""
    pref, p, s, b = O.kmer_params(int(g["gs"]))
    slab = np.ascontiguousarray(g["fastq"])
    off, ln, _, _ = S.parse_fastq(slab)
    recs = np.ascontiguousarray(g["recs"])
    n_base = int((recs["pos"] < 0xFFFFFFF0).sum())
    out = np.zeros((n_base, 8), np.uint64)
    raw = np.zeros(n_base, CTX_REC)
    off = np.ascontiguousarray(off, np.uint64); ln = np.ascontiguousarray(ln, np.uint32)
    n = lib.ctx_from_tap(recs.ctypes.data, len(recs), slab.ctypes.data, off.ctypes.data, ln.ctypes.data, len(off), p, s, b, pref, out.ctypes.data, n_base, raw.ctypes.data)
    assert n != 2 ** 64 - 1, "record stream and reads disagree"
    return out[:n], raw


def test_ctx_records_reproduce_the_reference_coder_ids(tmp_path):
    lib = build_harness(tmp_path)
    g = H.load_golden("se_ctx_gs1")
    got, raw = ctx_expected_from_tap(lib, g)
    want = g["ctx_ids"]
    assert len(got) == len(want) and len(want) > 30000, (len(got), len(want))
    bad = np.flatnonzero((got != want).any(axis=1))
    assert len(bad) == 0, (len(bad), bad[:3], [hex(int(x)) for x in got[bad[0]]], [hex(int(x)) for x in want[bad[0]]])
    # the fixture must reach every ingredient: all four count tables, correction zones, rough results, N runs, ranks 0..4
    recs = g["recs"][g["recs"]["pos"] < 0xFFFFFFF0]
    assert set(np.unique(recs["level"])) >= {0, 1, 2, 3} and (recs["rough"] == 1).sum() > 100
    assert set(np.unique(want[:, 7])) >= {0, 1, 2, 3}
    assert ((raw["a"] >> np.uint64(58)) & np.uint64(1)).sum() == len(want)


@pytest.mark.gpu
def test_ctx_records_from_the_device(tmp_path):
    """fqsk_submit_ctx (k_ctx_codes): the engine's 16-byte context records for every base of the fixture, through the reference's block
    and sync schedule, against the records include/fqsk_ctx.h produces on the CPU from the reference's own tap -- which the test above
    pins to the ids of the reference's coder.  Bit-exact, every base (coded with counts or not)."""
    from fqsqueezer_b200 import engine as E
    lib = build_harness(tmp_path)
    g = H.load_golden("se_ctx_gs1")
    _, want = ctx_expected_from_tap(lib, g)
    pref, p, s, b = E.kmer_params(int(g["gs"]))
    e = E.KmerEngine(p, s, b, pref)
    slab = g["fastq"]
    off, ln, roff, rsz = S.parse_fastq(slab)
    out, pend = [], None
    for gen, (f, l) in enumerate(S.split_blocks(rsz)):
        e.block_start()
        for a, bb in S.segments(f, l, S.calc_no_synchronizations(gen, l - f, 1)):
            t = e.submit(slab, off[a:bb], ln[a:bb], ctx=True)
            if pend is not None:
                out.append(e.collect(pend)[0].copy())
            pend = t
    out.append(e.collect(pend)[0].copy())
    got = np.concatenate(out)
    assert got.dtype == E.CTX_REC_DTYPE and len(got) == len(want), (len(got), len(want))
    bad = np.flatnonzero((got["a"] != want["a"]) | (got["b"] != want["b"]))
    assert len(bad) == 0, (len(bad), bad[:3], hex(int(got["a"][bad[0]])), hex(int(want["a"][bad[0]])), hex(int(got["b"][bad[0]])), hex(int(want["b"][bad[0]])))
    H.assert_dump_equal(e, g)
    assert e.stats()["kernel_launches"] > 0
    e.close()
