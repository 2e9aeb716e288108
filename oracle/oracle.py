"""ctypes loaders for the CPU oracle (oracle/libfqs_oracle.so) and, when present, the real-reference harness
(oracle/_ref/libfqs_ref.so).  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The product package never does."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libfqs_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libfqs_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "fqs-1.1")
REF_TAP_BIN = os.path.join(HERE, "_ref", "fqs-1.1-tap")

# one record per coded base -- same layout as the reference tap (oracle/build_ref.py) and as fqsk_base_rec
REC_DTYPE = np.dtype([("pos", "<u4"), ("c", "<u4", 4), ("cor_pos", "<u4"), ("level", "u1"), ("rough", "u1"), ("pad", "<u2")])
POS_READ, POS_SYNC, POS_DUP, POS_SORTED = 0xFFFFFFFF, 0xFFFFFFFE, 0xFFFFFFFD, 0xFFFFFFFC

_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "fqs_oracle.cpp")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", src, "-o", ORACLE_SO], check=True)
    return ORACLE_SO


def _bind_common(lib, pfx):
    g = lambda n: getattr(lib, pfx + n)
    g("mt_stream").argtypes = [C.c_uint32, C.c_uint64, _u32p]
    g("cinc_new").restype = C.c_void_p
    g("cinc_new").argtypes = [C.c_uint32] * 3
    g("cinc_free").argtypes = [C.c_void_p]
    g("cinc_inc1").argtypes = [C.c_void_p, _u32p, C.c_uint64, _u32p]
    g("cinc_incn").argtypes = [C.c_void_p, _u32p, _u32p, C.c_uint64, _u32p]
    g("ht_new").restype = C.c_void_p
    g("ht_free").argtypes = [C.c_void_p]
    g("ht_insert").argtypes = [C.c_void_p, C.c_void_p, _u64p, C.c_uint64]
    g("ht_find").argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, _u64p, _u64p, _u32p, C.c_uint64, _u32p, _u8p]
    g("ht_count").argtypes = [C.c_void_p, _u64p, C.c_uint64, _u32p]
    g("ht_dump").restype = C.c_uint64
    g("ht_dump").argtypes = [C.c_void_p, _u64p, _u32p, C.c_uint64]
    g("kmer_script").argtypes = [C.c_uint32, _u32p, C.c_uint64, _u64p, _u32p]
    g("siv_new").restype = C.c_void_p
    g("siv_new").argtypes = [C.c_uint32]
    g("siv_free").argtypes = [C.c_void_p]
    g("siv_increment").restype = C.c_uint64
    g("siv_increment").argtypes = [C.c_void_p, _u64p, C.c_uint64]
    g("siv_test").argtypes = [C.c_void_p, _u64p, C.c_uint64, _u32p]
    g("siv_counts").argtypes = [C.c_void_p, _u64p, C.c_uint64, _u32p]
    g("siv_test_shorter").argtypes = [C.c_void_p, _u64p, _u32p, C.c_uint64, _u64p]
    g("pair_new").restype = C.c_void_p
    g("pair_free").argtypes = [C.c_void_p]
    g("pair_insert").argtypes = [C.c_void_p, _u64p, _u64p, _u64p, C.c_uint64]
    g("pair_find").restype = C.c_uint64
    g("pair_find").argtypes = [C.c_void_p, C.c_uint64, _u64p, C.c_uint64]
    g("pair_count").argtypes = [C.c_void_p, _u64p, _u64p, C.c_uint64, _u64p]


class _Units:
    """Unit-level view (same method names for the oracle and for the real reference)."""

    def __init__(self, lib, pfx, is_ref):
        self.lib, self.pfx, self.is_ref = lib, pfx, is_ref

    def _f(self, n):
        return getattr(self.lib, self.pfx + n)

    def mt_stream(self, seed, n):
        out = np.empty(n, np.uint32)
        self._f("mt_stream")(seed, n, out)
        return out

    def cinc_new(self, thr, mult, top):
        return self._f("cinc_new")(thr, mult, top)

    def cinc_free(self, c):
        self._f("cinc_free")(c)

    def cinc_inc1(self, c, cnt):
        cnt = np.ascontiguousarray(cnt, np.uint32)
        out = np.empty_like(cnt)
        self._f("cinc_inc1")(c, cnt, len(cnt), out)
        return out

    def cinc_incn(self, c, cnt, inc):
        cnt = np.ascontiguousarray(cnt, np.uint32)
        inc = np.ascontiguousarray(inc, np.uint32)
        out = np.empty_like(cnt)
        self._f("cinc_incn")(c, cnt, inc, len(cnt), out)
        return out

    def ht_new(self, k, counter_bits, item_bytes=4):
        if self.is_ref:
            self.lib.ref_ht_new.argtypes = [C.c_uint32] * 3
            return self.lib.ref_ht_new(k, counter_bits, item_bytes)
        self.lib.fqso_ht_new.argtypes = [C.c_uint32] * 2
        return self.lib.fqso_ht_new(k, counter_bits)

    def ht_free(self, h):
        self._f("ht_free")(h)

    def ht_insert(self, h, cinc, kmers):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        self._f("ht_insert")(h, cinc, kmers, len(kmers))

    def ht_find(self, h, cinc, k, d, rc, cur):
        d = np.ascontiguousarray(d, np.uint64)
        rc = np.ascontiguousarray(rc, np.uint64)
        cur = np.ascontiguousarray(cur, np.uint32)
        out = np.empty(4 * len(d), np.uint32)
        found = np.empty(len(d), np.uint8)
        self._f("ht_find")(h, cinc, k, d, rc, cur, len(d), out, found)
        return out.reshape(-1, 4), found

    def ht_count(self, h, kmers):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        out = np.empty(len(kmers), np.uint32)
        self._f("ht_count")(h, kmers, len(kmers), out)
        return out

    def ht_clear(self, h):
        fn = self.lib.ref_ht_clear0 if self.is_ref else self.lib.fqso_ht_clear
        fn.argtypes = [C.c_void_p]
        fn(h)

    def ht_dump(self, h, cap=1 << 22):
        k = np.empty(cap, np.uint64)
        c = np.empty(cap, np.uint32)
        n = self._f("ht_dump")(h, k, c, cap)
        assert n <= cap
        o = np.argsort(k[:n], kind="stable")
        return k[:n][o], c[:n][o]

    def kmer_script(self, k, ops):
        ops = np.ascontiguousarray(ops, np.uint32).reshape(-1, 3)
        o64 = np.empty((len(ops), 6), np.uint64)
        o32 = np.empty((len(ops), 3), np.uint32)
        self._f("kmer_script")(k, ops.reshape(-1), len(ops), o64.reshape(-1), o32.reshape(-1))
        return o64, o32

    def siv_new(self, key_bits):
        return self._f("siv_new")(key_bits)

    def siv_free(self, s):
        self._f("siv_free")(s)

    def siv_increment(self, s, idx):
        idx = np.ascontiguousarray(idx, np.uint64)
        return int(self._f("siv_increment")(s, idx, len(idx)))

    def siv_test(self, s, idx):
        idx = np.ascontiguousarray(idx, np.uint64)
        out = np.empty(len(idx), np.uint32)
        self._f("siv_test")(s, idx, len(idx), out)
        return out

    def siv_counts(self, s, idx):
        idx = np.ascontiguousarray(idx, np.uint64)
        out = np.empty(4 * len(idx), np.uint32)
        self._f("siv_counts")(s, idx, len(idx), out)
        return out.reshape(-1, 4)

    def siv_test_shorter(self, s, idx, size_bits):
        idx = np.ascontiguousarray(idx, np.uint64)
        size_bits = np.ascontiguousarray(size_bits, np.uint32)
        out = np.empty(len(idx), np.uint64)
        self._f("siv_test_shorter")(s, idx, size_bits, len(idx), out)
        return out

    def pair_new(self, k, parts=1):
        if self.is_ref:
            self.lib.ref_pair_new.argtypes = [C.c_uint32, C.c_uint64]
            return self.lib.ref_pair_new(k, parts)
        self.lib.fqso_pair_new.argtypes = [C.c_uint32]
        return self.lib.fqso_pair_new(k)

    def pair_free(self, p):
        self._f("pair_free")(p)

    def pair_insert(self, p, key, val, cnt):
        key = np.ascontiguousarray(key, np.uint64)
        val = np.ascontiguousarray(val, np.uint64)
        cnt = np.ascontiguousarray(cnt, np.uint64)
        self._f("pair_insert")(p, key, val, cnt, len(key))

    def pair_find(self, p, key, cap=4096):
        out = np.empty(cap, np.uint64)
        n = int(self._f("pair_find")(p, int(key), out, cap))
        return np.sort(out[:min(n, cap)])

    def pair_count(self, p, key, val):
        key = np.ascontiguousarray(key, np.uint64)
        val = np.ascontiguousarray(val, np.uint64)
        out = np.empty(len(key), np.uint64)
        self._f("pair_count")(p, key, val, len(key), out)
        return out


_oracle_lib = None
_ref_lib = None


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is None:
        build_oracle()
        lib = C.CDLL(ORACLE_SO)
        _bind_common(lib, "fqso_")
        lib.fqso_create.restype = C.c_void_p
        lib.fqso_create.argtypes = [C.c_uint32] * 5
        lib.fqso_destroy.argtypes = [C.c_void_p]
        lib.fqso_block_start.argtypes = [C.c_void_p]
        lib.fqso_segment.restype = C.c_uint64
        lib.fqso_segment.argtypes = [C.c_void_p, _u8p, _u64p, _u32p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, _u8p]
        lib.fqso_sync.argtypes = [C.c_void_p]
        lib.fqso_dump.restype = C.c_uint64
        lib.fqso_dump.argtypes = [C.c_void_p, C.c_uint32, _u64p, _u64p, C.c_uint64]
        lib.fqso_stats.argtypes = [C.c_void_p, _u64p]
        lib.fqso_pending.restype = C.c_uint64
        lib.fqso_pending.argtypes = [C.c_void_p, C.c_uint32, _u64p, C.c_uint64]
        lib.fqso_create_worker.restype = C.c_void_p
        lib.fqso_create_worker.argtypes = [C.c_void_p]
        lib.fqso_sync_group.argtypes = [C.POINTER(C.c_void_p), C.c_uint32]
        _oracle_lib = lib
    return _oracle_lib


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        lib = C.CDLL(REF_SO)
        _bind_common(lib, "ref_")
        _ref_lib = lib
    return _ref_lib


def sort_ranks(slab: np.ndarray, off: np.ndarray, length: np.ndarray) -> np.ndarray:
    """rank[i] = number of reads the comparator of CSortedFASTQFile::sort_reads (io.h:499-528) puts strictly before read i."""
    lib = oracle_lib()
    lib.fqso_sort_ranks.argtypes = [_u8p, _u64p, _u32p, C.c_uint32, _u32p]
    lib.fqso_sort_ranks.restype = None
    n = len(off)
    rank = np.zeros(max(n, 1), np.uint32)
    lib.fqso_sort_ranks(np.ascontiguousarray(slab, np.uint8), np.ascontiguousarray(off, np.uint64), np.ascontiguousarray(length, np.uint32), n, rank)
    return rank[:n]


def oracle_units() -> _Units:
    return _Units(oracle_lib(), "fqso_", False)


def ref_units() -> _Units:
    return _Units(ref_lib(), "ref_", True)


STAT_NAMES = ["siv_no_filled", "siv_no_updates", "n_smers", "n_bmers", "draws_b", "draws_s", "draws_lb", "draws_ls",
              "rough_b", "rough_s", "rough_p", "repair_existing", "repair_missing", "local_hits", "probe_runs_s", "probe_runs_b", "mixed", "unc_reverts"]


class OracleEngine:
    """Sequential CPU model of one reference worker (-t 1): CDNACompressor's k-mer half + the global tables."""

    def __init__(self, p, s, b, prefix_len, mode=0):
        self.lib = oracle_lib()
        self.h = self.lib.fqso_create(p, s, b, prefix_len, mode)
        self.p, self.s, self.b, self.prefix_len, self.mode = p, s, b, prefix_len, mode

    def close(self):
        if self.h:
            self.lib.fqso_destroy(self.h)
            self.h = None

    __del__ = close

    def block_start(self):
        self.lib.fqso_block_start(self.h)

    def segment(self, slab: np.ndarray, off: np.ndarray, length: np.ndarray, kind: int = 0):
        slab = np.ascontiguousarray(slab, np.uint8)
        off = np.ascontiguousarray(off, np.uint64)
        length = np.ascontiguousarray(length, np.uint32)
        n = len(off)
        cap = int(length.sum()) + 3 * n + 16
        recs = np.zeros(cap, REC_DTYPE)
        dup = np.zeros(max(n, 1), np.uint8)
        m = self.lib.fqso_segment(self.h, slab, off, length, n, kind, recs.ctypes.data, cap, dup)
        assert m <= cap
        return recs[:m], dup[:n]

    def sync(self):
        self.lib.fqso_sync(self.h)

    def dump(self, which, cap=1 << 24):
        k = np.empty(cap, np.uint64)
        v = np.empty(cap, np.uint64)
        n = self.lib.fqso_dump(self.h, which, k, v, cap)
        assert n <= cap
        return k[:n].copy(), v[:n].copy()

    def stats(self):
        o = np.zeros(len(STAT_NAMES), np.uint64)
        self.lib.fqso_stats(self.h, o)
        return dict(zip(STAT_NAMES, (int(x) for x in o)))

    def pending(self, which, cap=1 << 24):
        o = np.empty(cap, np.uint64)
        n = self.lib.fqso_pending(self.h, which, o, cap)
        assert n <= cap
        return o[:n].copy()


class OracleGroup:
    """Sequential CPU model of the reference at -t T: T workers (one CDNACompressor each: own PRNG streams, thread-local tables,
    s_letters, read_prev) on shared global tables; `sync` is InsertKmersToHT + ClearKmersToHT of all of them with the
    reference's owner routing (dna.cpp:2393-2488).  workers[i] has the OracleEngine interface for that worker's reads."""

    def __init__(self, p, s, b, prefix_len, n_workers, mode=0):
        self.workers = [OracleEngine(p, s, b, prefix_len, mode)]
        lib = self.workers[0].lib
        for _ in range(1, n_workers):
            w = OracleEngine.__new__(OracleEngine)
            w.lib = lib
            w.h = lib.fqso_create_worker(self.workers[0].h)
            w.p, w.s, w.b, w.prefix_len, w.mode = p, s, b, prefix_len, mode
            self.workers.append(w)
        self.lib = lib

    def sync(self):
        arr = (C.c_void_p * len(self.workers))(*[w.h for w in self.workers])
        self.lib.fqso_sync_group(arr, len(self.workers))

    def dump(self, which):
        return self.workers[0].dump(which)

    def stats(self):
        return self.workers[0].stats()

    def close(self):
        for w in reversed(self.workers):     # worker 0 owns the shared tables: free it last
            w.close()


def kmer_params(genome_size_mb: int):
    """-gs -> (prefix_len, pmer, smer, bmer): params.h:131-155."""
    table = [(1, 9, 14, 17, 19), (4, 9, 15, 18, 20), (16, 10, 15, 18, 21), (64, 11, 16, 18, 23), (256, 12, 17, 20, 24),
             (1024, 12, 17, 21, 26), (4096, 13, 18, 21, 27), (16384, 14, 18, 22, 27), (65536, 15, 18, 22, 27)]
    for gs, pref, p, s, b in table:
        if genome_size_mb <= gs:
            return pref, p, s, b
    return 14, 13, 15, 26  # CParams defaults (params.h:68-75) when -gs exceeds the table
