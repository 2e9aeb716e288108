"""ctypes host mirror of include/fqsk.h.

`KmerEngine` keeps the vocabulary of the reference's host code for this path: block_start (application.cpp:624),
segment (the k-mer half of CDNACompressor::CompressDirect / CompressSorted for the reads between two syncs),
sync (InsertKmersToHT + ClearKmersToHT, dna.cpp:2393-2488), dump / stats, and the table-level batch mirrors of
CHT_kmer<T> and TSmallIntVector<2>.  Everything computes on the GPU; failures raise FqskError."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfqsk.so")

REC_DTYPE = np.dtype([("pos", "<u4"), ("c", "<u4", 4), ("cor_pos", "<u4"), ("level", "u1"), ("rough", "u1"), ("pad", "<u2")])
READ_DESC_DTYPE = np.dtype([("dna_off", "<u8"), ("dna_len", "<u4"), ("flags", "<u4")])
CTX_REC_DTYPE = np.dtype([("a", "<u8"), ("b", "<u8")])      # fqsk_ctx_rec (include/fqsk_ctx.h)
TABLE_SIV, TABLE_SMER, TABLE_BMER, TABLE_PAIR = 0, 1, 2, 3
MODE_SE_ORIGINAL, MODE_SE_SORTED, MODE_PE_ORIGINAL, MODE_PE_SORTED = 0, 1, 2, 3
F_PROFILE, F_TRACE_ALLOC, F_TEST_HOOKS, F_TRACE_LAUNCH, F_SERIAL, F_TEST_CROWD = 1, 2, 4, 8, 16, 32
PHASES = ["prep", "lookup", "partial", "walk", "compact", "sort", "local", "rough", "fold", "sync_locate", "sync_apply", "sync_siv", "mt"]


class FqskError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fqsk error {code}: {msg}")
        self.code = code


class _Params(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("pmer_len", C.c_uint32), ("smer_len", C.c_uint32), ("bmer_len", C.c_uint32),
                ("prefix_len", C.c_uint32), ("smer_counter_bits", C.c_uint32), ("bmer_counter_bits", C.c_uint32), ("mode", C.c_uint32),
                ("n_workers", C.c_uint32), ("device", C.c_int32), ("bmer_log2_buckets", C.c_uint32), ("smer_log2_buckets", C.c_uint32),
                ("expected_kmers", C.c_uint64), ("world_size", C.c_uint32), ("rank", C.c_uint32), ("max_iterations", C.c_uint32),
                ("flags", C.c_uint32), ("reserve_reads", C.c_uint32), ("reserve_bytes", C.c_uint32), ("pair_log2_slots", C.c_uint32), ("test_hooks", C.c_uint32)]


class _Stats(C.Structure):
    _fields_ = [("siv_no_filled", C.c_uint64), ("siv_no_updates", C.c_uint64), ("n_smers", C.c_uint64), ("n_bmers", C.c_uint64),
                ("draws", C.c_uint64 * 4), ("n_segments", C.c_uint64), ("n_syncs", C.c_uint64), ("n_replays", C.c_uint64),
                ("n_bases", C.c_uint64), ("n_reads", C.c_uint64), ("kernel_launches", C.c_uint64), ("bmer_buckets", C.c_uint64),
                ("smer_buckets", C.c_uint64), ("bmer_stash_used", C.c_uint64), ("smer_stash_used", C.c_uint64), ("n_hot_segments", C.c_uint64),
                ("n_filtered_segments", C.c_uint64), ("n_looks", C.c_uint64), ("look_wait_ns", C.c_uint64), ("api_ns", C.c_uint64),
                ("n_table_growths", C.c_uint64)]


EXPORTS = ["fqsk_create", "fqsk_destroy", "fqsk_last_error", "fqsk_block_start", "fqsk_segment", "fqsk_segment_device", "fqsk_announce_device",
           "fqsk_device_recs", "fqsk_recs_checksum", "fqsk_sorted_prefix", "fqsk_pair_info", "fqsk_submit", "fqsk_submit_ctx", "fqsk_collect", "fqsk_block_host", "fqsk_block_stream", "fqsk_sync", "fqsk_dump", "fqsk_stats_get", "fqsk_profile", "fqsk_ht_insert", "fqsk_ht_find",
           "fqsk_ht_count", "fqsk_timeline", "fqsk_timer_begin", "fqsk_timer_end", "fqsk_siv_increment", "fqsk_siv_test", "fqsk_siv_counts", "fqsk_siv_test_shorter", "fqsk_mt_stream", "fqsk_host_alloc", "fqsk_host_free",
           "fqsk_sort_ranks", "fqsk_shard_export", "fqsk_shard_attach", "fqsk_sync_route", "fqsk_sync_apply", "fqsk_sync_finish", "fqsk_sync_device",
           "fqsk_shard_grow_request", "fqsk_shard_grow", "fqsk_shard_attach_local"]

_lib = None


def load_library():
    """Loads libfqsk.so (built in-tree by fqsqueezer_b200/build.py).  Raises if it is missing: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FqskError(-2, f"{LIB_PATH} not built (run `python -m fqsqueezer_b200.build`); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, u64p, u32p, u8p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    lib.fqsk_create.argtypes = [C.POINTER(_Params), C.POINTER(vp)]
    lib.fqsk_destroy.argtypes = [vp]
    lib.fqsk_destroy.restype = None
    lib.fqsk_last_error.argtypes = [vp]
    lib.fqsk_last_error.restype = C.c_char_p
    lib.fqsk_block_start.argtypes = [vp]
    lib.fqsk_segment.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint32, vp, C.c_uint64, u64p, vp, vp]
    lib.fqsk_segment_device.argtypes = [vp, vp, C.c_uint64, vp, vp, C.c_uint32, u64p]
    lib.fqsk_announce_device.argtypes = [vp, vp, C.c_uint64, vp, vp, C.c_uint32]
    lib.fqsk_device_recs.argtypes = [vp, C.POINTER(vp), u64p]
    lib.fqsk_recs_checksum.argtypes = [vp, u64p, u64p]
    lib.fqsk_sorted_prefix.argtypes = [vp, vp, vp, C.c_uint32]
    lib.fqsk_pair_info.argtypes = [vp, vp, C.c_uint32]
    lib.fqsk_submit.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint32, vp, C.c_uint64, vp, vp, u64p]
    lib.fqsk_submit_ctx.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint32, vp, C.c_uint64, vp, vp, u64p]
    lib.fqsk_collect.argtypes = [vp, C.c_uint64, u64p]
    lib.fqsk_block_host.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint64, vp, vp, vp]
    lib.fqsk_block_stream.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint32, vp, C.c_uint32, vp, vp, C.c_uint64, vp, vp, vp, u64p, u64p]
    lib.fqsk_sync.argtypes = [vp]
    lib.fqsk_dump.argtypes = [vp, C.c_int, vp, vp, C.c_uint64, u64p]
    lib.fqsk_stats_get.argtypes = [vp, C.POINTER(_Stats)]
    lib.fqsk_profile.argtypes = [vp, C.POINTER(C.c_double), C.c_uint32]
    lib.fqsk_timer_begin.argtypes = [vp]
    lib.fqsk_timeline.argtypes = [vp, C.c_int]
    lib.fqsk_timer_end.argtypes = [vp, C.POINTER(C.c_double)]
    lib.fqsk_ht_insert.argtypes = [vp, C.c_int, vp, C.c_uint64]
    lib.fqsk_ht_find.argtypes = [vp, C.c_int, vp, vp, vp, C.c_uint64, vp]
    lib.fqsk_ht_count.argtypes = [vp, C.c_int, vp, C.c_uint64, vp]
    lib.fqsk_siv_increment.argtypes = [vp, vp, C.c_uint64, u64p]
    lib.fqsk_siv_test.argtypes = [vp, vp, C.c_uint64, vp]
    lib.fqsk_siv_counts.argtypes = [vp, vp, C.c_uint64, vp]
    lib.fqsk_siv_test_shorter.argtypes = [vp, vp, vp, C.c_uint64, vp]
    lib.fqsk_mt_stream.argtypes = [vp, C.c_uint64, vp]
    lib.fqsk_host_alloc.argtypes = [C.c_uint64, C.POINTER(vp)]
    lib.fqsk_sort_ranks.argtypes = [C.c_int, vp, C.c_uint64, vp, C.c_uint32, vp]
    lib.fqsk_host_free.argtypes = [vp]
    lib.fqsk_host_free.restype = None
    for n in EXPORTS:
        if n not in ("fqsk_destroy", "fqsk_last_error", "fqsk_host_free"):
            getattr(lib, n).restype = C.c_int
    _lib = lib
    return lib


def kmer_params(genome_size_mb: int):
    """-gs -> (prefix_len, pmer, smer, bmer), CParams::adjust_kmer_sizes (params.h:131-155)."""
    table = [(1, 9, 14, 17, 19), (4, 9, 15, 18, 20), (16, 10, 15, 18, 21), (64, 11, 16, 18, 23), (256, 12, 17, 20, 24),
             (1024, 12, 17, 21, 26), (4096, 13, 18, 21, 27), (16384, 14, 18, 22, 27), (65536, 15, 18, 22, 27)]
    for gs, pref, p, s, b in table:
        if genome_size_mb <= gs:
            return pref, p, s, b
    return 14, 13, 15, 26


def recs_checksum_host(recs: np.ndarray) -> int:
    """fqsk_recs_checksum over a host-side record array (the oracle's records in bench.py's parity_check and in the tests)."""
    n = len(recs)
    if n == 0:
        return 0
    w = np.ascontiguousarray(recs).view(np.uint32).reshape(n, 7).astype(np.uint64)
    m1, m2 = np.uint64(0xff51afd7ed558ccd), np.uint64(0xc4ceb9fe1a85ec53)

    def fmix(x):
        x = x ^ (x >> np.uint64(33)); x = x * m1; x = x ^ (x >> np.uint64(33)); x = x * m2; return x ^ (x >> np.uint64(33))

    a, b, c, d = w[:, 0] | (w[:, 1] << np.uint64(32)), w[:, 2] | (w[:, 3] << np.uint64(32)), w[:, 4] | (w[:, 5] << np.uint64(32)), w[:, 6]
    with np.errstate(over="ignore"):
        h = fmix(a ^ fmix(b ^ fmix(c ^ fmix(d + np.arange(n, dtype=np.uint64)))))
        return int(h.sum(dtype=np.uint64))


def sort_ranks(slab: np.ndarray, off: np.ndarray, length: np.ndarray, device: int = 0) -> np.ndarray:
    """fqsk_sort_ranks: per read an integer order-isomorphic to the comparator of CSortedFASTQFile::sort_reads (io.h:499-528)."""
    lib = load_library()
    slab = np.ascontiguousarray(slab, np.uint8)
    n = len(off)
    desc = np.zeros(n, READ_DESC_DTYPE)
    desc["dna_off"] = off
    desc["dna_len"] = length
    rank = np.zeros(n, np.uint32)
    rc = lib.fqsk_sort_ranks(device, _ptr(slab), slab.size, _ptr(desc), n, _ptr(rank))
    if rc != 0:
        raise FqskError(rc, lib.fqsk_last_error(None).decode())
    return rank


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class KmerEngine:
    def __init__(self, p, s, b, prefix_len, mode=MODE_SE_ORIGINAL, device=0, expected_kmers=0, bmer_log2_buckets=0,
                 smer_log2_buckets=0, profile=False, max_iterations=0, reserve_reads=0, reserve_bytes=0, test_fail_every=0, test_retry_every=0, flags=0):
        self.lib = load_library()
        hooks = (test_fail_every & 0xFFFF) | ((test_retry_every & 0xFFFF) << 16)      # fault injection of the recovery paths (tests only)
        prm = _Params(abi_version=1, pmer_len=p, smer_len=s, bmer_len=b, prefix_len=prefix_len, smer_counter_bits=12,
                      bmer_counter_bits=6, mode=mode, n_workers=1, device=device, bmer_log2_buckets=bmer_log2_buckets,
                      smer_log2_buckets=smer_log2_buckets, expected_kmers=expected_kmers, world_size=1, rank=0,
                      max_iterations=max_iterations, flags=(F_PROFILE if profile else 0) | (F_TEST_HOOKS if hooks else 0) | flags, reserve_reads=reserve_reads, reserve_bytes=reserve_bytes,
                      test_hooks=hooks)
        h = C.c_void_p()
        rc = self.lib.fqsk_create(C.byref(prm), C.byref(h))
        if rc != 0:
            raise FqskError(rc, self.lib.fqsk_last_error(None).decode())
        self.h = h
        self.p, self.s, self.b, self.prefix_len, self.mode = p, s, b, prefix_len, mode
        self.reserve_reads, self.reserve_bytes = reserve_reads, reserve_bytes

    def close(self):
        for ptr, _ in getattr(self, "_pinned", {}).values():
            self.lib.fqsk_host_free(ptr)
        self._pinned = {}
        self.__dict__.pop("_submit_cache", None)
        self.__dict__.pop("_submit_cache_ctx", None)
        self.__dict__.pop("_stream_state", None)
        if getattr(self, "h", None):
            self.lib.fqsk_destroy(self.h)
            self.h = None

    def _pinned_array(self, name: str, nbytes: int) -> np.ndarray:
        """A reusable page-locked byte buffer (fqsk_host_alloc) of at least nbytes, as a numpy uint8 view."""
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        cur = self._pinned.get(name)
        if cur is None or cur[1] < nbytes:
            if cur is not None:
                self.lib.fqsk_host_free(cur[0])
            cap = max(int(nbytes * 1.25) + 4096, 1 << 16)
            ptr = C.c_void_p()
            self._ck(self.lib.fqsk_host_alloc(cap, C.byref(ptr)))
            self._pinned[name] = cur = (ptr, cap)
        return np.ctypeslib.as_array(C.cast(cur[0], C.POINTER(C.c_uint8)), shape=(cur[1],))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise FqskError(rc, self.lib.fqsk_last_error(self.h).decode())

    # ---- segment-level API ----
    def block_start(self):
        self._ck(self.lib.fqsk_block_start(self.h))

    def segment(self, slab: np.ndarray, off: np.ndarray, length: np.ndarray, kind: int = 0, want_rec_off=False, pinned=False):
        """One sync segment from host memory.  pinned=True returns views into reusable page-locked buffers owned by the engine
        (valid until the next call): the records then come back by DMA without a staging copy."""
        slab = np.ascontiguousarray(slab, np.uint8)
        n = len(off)
        desc = np.zeros(n, READ_DESC_DTYPE)
        desc["dna_off"] = off
        desc["dna_len"] = length
        cap = int(np.asarray(length, np.int64).sum()) + 16
        if pinned:
            recs = self._pinned_array("recs", cap * REC_DTYPE.itemsize)[: cap * REC_DTYPE.itemsize].view(REC_DTYPE)
            dup = self._pinned_array("dup", max(n, 1))[: max(n, 1)]
            rec_off = self._pinned_array("rec_off", (n + 1) * 8)[: (n + 1) * 8].view(np.uint64)
        else:
            recs = np.zeros(cap, REC_DTYPE)
            dup = np.zeros(max(n, 1), np.uint8)
            rec_off = np.zeros(n + 1, np.uint64)
        n_recs = C.c_uint64(0)
        self._ck(self.lib.fqsk_segment(self.h, _ptr(slab), slab.size, _ptr(desc), n, _ptr(recs), cap, C.byref(n_recs), _ptr(dup), _ptr(rec_off)))
        out = recs[: n_recs.value]
        if want_rec_off:
            return out, dup[:n], rec_off
        return out, dup[:n]

    def submit(self, slab: np.ndarray, off: np.ndarray, length: np.ndarray, first: int | None = None, ctx: bool = False):
        """Asynchronous segment + sync (fqsk_submit): returns a ticket; collect(ticket) -> (records, dup, rec_off).  The output
        buffers are page-locked and owned by the engine, two sets used alternately: a ticket's arrays stay valid until the
        second submit after it.  ctx=True: fqsk_submit_ctx -- the records are the 16-byte context records of include/fqsk_ctx.h
        (CTX_REC_DTYPE) built on the device instead of the 28-byte per-base records."""
        if slab.dtype != np.uint8 or not slab.flags.c_contiguous:
            slab = np.ascontiguousarray(slab, np.uint8)
        n = len(off)
        slot = getattr(self, "_submit_slot", 0) ^ 1
        self._submit_slot = slot
        dt = CTX_REC_DTYPE if ctx else REC_DTYPE
        # descriptor array and page-locked output buffers are kept per slot and reused (sized once for the largest announced
        # segment: reallocating 200 MB of pinned memory costs ~0.1 s, building numpy views ~10 us each)
        cache = self.__dict__.setdefault("_submit_cache_ctx" if ctx else "_submit_cache", {})
        ent = cache.get(slot)
        cap = int(length.sum(dtype=np.int64)) + 16
        n_cap = max(n, self.reserve_reads, 1)
        r_cap = max(cap, self.reserve_bytes + 16)
        if ent is None or ent[0] < n_cap or ent[1] < r_cap:
            ent = (n_cap, r_cap, np.zeros(n_cap, READ_DESC_DTYPE),
                   self._pinned_array(f"srecs{slot}{'c' if ctx else ''}", r_cap * dt.itemsize)[: r_cap * dt.itemsize].view(dt),
                   self._pinned_array(f"sdup{slot}", n_cap), self._pinned_array(f"soff{slot}", (n_cap + 1) * 8)[: (n_cap + 1) * 8].view(np.uint64))
            cache[slot] = ent
        desc = ent[2]
        desc["dna_off"][:n] = off
        desc["dna_len"][:n] = length
        recs, dup, rec_off = ent[3][:cap], ent[4][: max(n, 1)], ent[5][: n + 1]
        t = C.c_uint64(0)
        fn = self.lib.fqsk_submit_ctx if ctx else self.lib.fqsk_submit
        self._ck(fn(self.h, _ptr(slab), slab.size, _ptr(desc), n, _ptr(recs), cap, _ptr(dup), _ptr(rec_off), C.byref(t)))
        if not hasattr(self, "_tickets"):
            self._tickets = {}
        self._tickets[t.value] = (recs, dup, rec_off, n, (slab, desc))      # keep the inputs alive until the copy is staged (it is, on return)
        return t.value

    def collect(self, ticket: int):
        recs, dup, rec_off, n, _ = self._tickets.pop(ticket)
        n_recs = C.c_uint64(0)
        self._ck(self.lib.fqsk_collect(self.h, ticket, C.byref(n_recs)))
        return recs[: n_recs.value], dup[:n], rec_off

    def block_host(self, slab: np.ndarray, off: np.ndarray, length: np.ndarray, seg_end):
        """One reads_block through fqsk_block_host: the worker loop of the reference (block start, then segment k + 1 submitted before
        segment k is collected) in one call.  seg_end: exclusive end (read index) of every sync segment.  Returns (records buffer,
        duplicate flags, first record of every segment, records of every segment); the buffers are page-locked, owned by the engine and
        valid until the next block_host call."""
        if slab.dtype != np.uint8 or not slab.flags.c_contiguous:
            slab = np.ascontiguousarray(slab, np.uint8)
        n = len(off)
        seg_end = np.ascontiguousarray(seg_end, np.uint32)
        ns = len(seg_end)
        cap = int(length.sum(dtype=np.int64)) + 16
        ent = self.__dict__.get("_block_cache")
        n_cap, r_cap = max(n, self.reserve_reads, 1), max(cap, self.reserve_bytes + 16)
        if ent is None or ent[0] < n_cap or ent[1] < r_cap:
            ent = (n_cap, r_cap, np.zeros(n_cap, READ_DESC_DTYPE),
                   self._pinned_array("brecs", r_cap * REC_DTYPE.itemsize)[: r_cap * REC_DTYPE.itemsize].view(REC_DTYPE), self._pinned_array("bdup", n_cap))
            self._block_cache = ent
        desc = ent[2]
        desc["dna_off"][:n] = off
        desc["dna_len"][:n] = length
        seg_off, seg_n = np.zeros(ns, np.uint64), np.zeros(ns, np.uint64)
        self._ck(self.lib.fqsk_block_host(self.h, _ptr(slab), slab.size, _ptr(desc), n, _ptr(seg_end), ns, _ptr(ent[3]), r_cap, _ptr(ent[4]), _ptr(seg_off), _ptr(seg_n)))
        return ent[3], ent[4][: max(n, 1)], seg_off, seg_n

    def block_stream(self, slab: np.ndarray, off: np.ndarray, length: np.ndarray, seg_end, ctx: bool = False):
        """One reads_block of a run of consecutive blocks through fqsk_block_stream: like block_host, but the block's last segment stays in
        flight and is collected by the next block_stream call (or by stream_finish after the last block).  Returns a list of record views
        in stream order that became complete during this call: the previous block's last segment (if any), then this block's segments
        0 .. n_segs - 2.  Page-locked buffers owned by the engine, two sets used alternately: the views stay valid until the second
        block_stream call after this one."""
        if slab.dtype != np.uint8 or not slab.flags.c_contiguous:
            slab = np.ascontiguousarray(slab, np.uint8)
        n = len(off)
        seg_end = np.ascontiguousarray(seg_end, np.uint32)
        ns = len(seg_end)
        dt = CTX_REC_DTYPE if ctx else REC_DTYPE
        cap = int(length.sum(dtype=np.int64)) + 16
        st = self.__dict__.setdefault("_stream_state", {"slot": 0, "ticket": C.c_uint64(0), "tail": None, "cache": {}})
        slot = st["slot"] ^ 1
        st["slot"] = slot
        key = (slot, bool(ctx))
        ent = st["cache"].get(key)
        n_cap, r_cap = max(n, self.reserve_reads, 1), max(cap, self.reserve_bytes + 16)
        if ent is None or ent[0] < n_cap or ent[1] < r_cap:
            ent = (n_cap, r_cap, np.zeros(n_cap, READ_DESC_DTYPE),
                   self._pinned_array(f"trecs{slot}{'c' if ctx else ''}", r_cap * dt.itemsize)[: r_cap * dt.itemsize].view(dt), self._pinned_array(f"tdup{slot}", n_cap))
            st["cache"][key] = ent
        desc = ent[2]
        desc["dna_off"][:n] = off
        desc["dna_len"][:n] = length
        seg_off, seg_n = np.zeros(ns, np.uint64), np.zeros(ns, np.uint64)
        carry_n = C.c_uint64(0)
        had = st["ticket"].value != 0
        self._ck(self.lib.fqsk_block_stream(self.h, _ptr(slab), slab.size, _ptr(desc), n, _ptr(seg_end), ns, None if ctx else _ptr(ent[3]), _ptr(ent[3]) if ctx else None, r_cap,
                                            _ptr(ent[4]), _ptr(seg_off), _ptr(seg_n), C.byref(st["ticket"]), C.byref(carry_n)))
        out = []
        if had:
            buf, o = st["tail"]
            out.append(buf[o:o + carry_n.value])
        for k in range(ns - 1):
            out.append(ent[3][int(seg_off[k]): int(seg_off[k]) + int(seg_n[k])])
        st["tail"] = (ent[3], int(seg_off[ns - 1]))
        st["keep"] = (slab, desc)
        return out

    def stream_finish(self):
        """Collects the segment the last block_stream call left in flight; returns its records (or None when there is none)."""
        st = self.__dict__.get("_stream_state")
        if not st or not st["ticket"].value:
            return None
        n_recs = C.c_uint64(0)
        self._ck(self.lib.fqsk_collect(self.h, st["ticket"].value, C.byref(n_recs)))
        st["ticket"] = C.c_uint64(0)
        buf, o = st["tail"]
        return buf[o:o + n_recs.value]

    def segment_device(self, d_dna_ptr: int, dna_bytes: int, d_off_ptr: int, d_len_ptr: int, n_reads: int, want_n_recs: bool = True):
        """Reads already in HBM.  want_n_recs=False only enqueues the segment: the library looks at the outcome when the
        following sync (or device_recs / stats) needs it, which saves one host round trip per segment."""
        if not want_n_recs:
            self._ck(self.lib.fqsk_segment_device(self.h, C.c_void_p(d_dna_ptr), dna_bytes, C.c_void_p(d_off_ptr), C.c_void_p(d_len_ptr), n_reads, None))
            return None
        n_recs = C.c_uint64(0)
        self._ck(self.lib.fqsk_segment_device(self.h, C.c_void_p(d_dna_ptr), dna_bytes, C.c_void_p(d_off_ptr), C.c_void_p(d_len_ptr), n_reads, C.byref(n_recs)))
        return n_recs.value

    def announce_device(self, d_dna_ptr: int, dna_bytes: int, d_off_ptr: int, d_len_ptr: int, n_reads: int):
        """Hint between segment_device and its sync: the arguments of the NEXT segment_device call of the same reads_block (its
        read-only preparation then runs next to the segment in flight)."""
        self._ck(self.lib.fqsk_announce_device(self.h, C.c_void_p(d_dna_ptr), dna_bytes, C.c_void_p(d_off_ptr), C.c_void_p(d_len_ptr), n_reads))

    def device_recs(self):
        p = C.c_void_p()
        n = C.c_uint64(0)
        self._ck(self.lib.fqsk_device_recs(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def recs_checksum(self):
        """(checksum, number of records) of the last segment, computed on the device (fqsk_recs_checksum)."""
        s, n = C.c_uint64(0), C.c_uint64(0)
        self._ck(self.lib.fqsk_recs_checksum(self.h, C.byref(s), C.byref(n)))
        return s.value, n.value

    def sorted_prefix(self, n_reads):
        """(flag, dif) per read of the last segment -- compress_prefix_sorted, dna.cpp:589-605."""
        flag = np.zeros(max(n_reads, 1), np.uint32)
        dif = np.zeros(max(n_reads, 1), np.uint64)
        self._ck(self.lib.fqsk_sorted_prefix(self.h, _ptr(flag), _ptr(dif), n_reads))
        return flag[:n_reads], dif[:n_reads]

    def pair_info(self, n_pairs):
        """(found, minim2_id, minim2_pos) per pair of the last segment -- what CompressPE codes, dna.cpp:1790-1880."""
        out = np.zeros((max(n_pairs, 1), 3), np.uint32)
        self._ck(self.lib.fqsk_pair_info(self.h, _ptr(out), n_pairs))
        return out[:n_pairs]

    def sync(self):
        self._ck(self.lib.fqsk_sync(self.h))

    def dump(self, which):
        n = C.c_uint64(0)
        self._ck(self.lib.fqsk_dump(self.h, which, None, None, 0, C.byref(n)))
        k = np.zeros(max(n.value, 1), np.uint64)
        v = np.zeros(max(n.value, 1), np.uint64)
        self._ck(self.lib.fqsk_dump(self.h, which, _ptr(k), _ptr(v), n.value, C.byref(n)))
        return k[: n.value], v[: n.value]

    def stats(self):
        st = _Stats()
        self._ck(self.lib.fqsk_stats_get(self.h, C.byref(st)))
        d = {f: getattr(st, f) for f, _ in _Stats._fields_ if f != "draws"}
        d.update(draws_b=st.draws[0], draws_s=st.draws[1], draws_lb=st.draws[2], draws_ls=st.draws[3])
        return d

    def timeline(self, on: bool):
        """Measurement aid: bracket every launch with events on its own stream (on) / print the collected launches to stderr (off)."""
        self._ck(self.lib.fqsk_timeline(self.h, 1 if on else 0))

    def timer_begin(self):
        self._ck(self.lib.fqsk_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_double(0)
        self._ck(self.lib.fqsk_timer_end(self.h, C.byref(ms)))
        return ms.value

    def profile(self):
        ms = (C.c_double * len(PHASES))()
        self._ck(self.lib.fqsk_profile(self.h, ms, len(PHASES)))
        return dict(zip(PHASES, list(ms)))

    # ---- table-level batch mirrors (CHT_kmer<T>, TSmallIntVector<2>) ----
    def ht_insert(self, table, kmers):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        self._ck(self.lib.fqsk_ht_insert(self.h, table, _ptr(kmers), len(kmers)))

    def ht_find(self, table, d, rc, cur):
        d = np.ascontiguousarray(d, np.uint64)
        rc = np.ascontiguousarray(rc, np.uint64)
        cur = np.ascontiguousarray(cur, np.uint32)
        out = np.zeros(4 * max(len(d), 1), np.uint32)
        self._ck(self.lib.fqsk_ht_find(self.h, table, _ptr(d), _ptr(rc), _ptr(cur), len(d), _ptr(out)))
        return out[: 4 * len(d)].reshape(-1, 4)

    def ht_count(self, table, kmers):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        out = np.zeros(max(len(kmers), 1), np.uint32)
        self._ck(self.lib.fqsk_ht_count(self.h, table, _ptr(kmers), len(kmers), _ptr(out)))
        return out[: len(kmers)]

    def siv_increment(self, idx):
        idx = np.ascontiguousarray(idx, np.uint64)
        nn = C.c_uint64(0)
        self._ck(self.lib.fqsk_siv_increment(self.h, _ptr(idx), len(idx), C.byref(nn)))
        return nn.value

    def siv_test(self, idx):
        idx = np.ascontiguousarray(idx, np.uint64)
        out = np.zeros(max(len(idx), 1), np.uint32)
        self._ck(self.lib.fqsk_siv_test(self.h, _ptr(idx), len(idx), _ptr(out)))
        return out[: len(idx)]

    def siv_counts(self, idx):
        idx = np.ascontiguousarray(idx, np.uint64)
        out = np.zeros(4 * max(len(idx), 1), np.uint32)
        self._ck(self.lib.fqsk_siv_counts(self.h, _ptr(idx), len(idx), _ptr(out)))
        return out[: 4 * len(idx)].reshape(-1, 4)

    def siv_test_shorter(self, idx, size_bits):
        idx = np.ascontiguousarray(idx, np.uint64)
        size_bits = np.ascontiguousarray(size_bits, np.uint32)
        out = np.zeros(max(len(idx), 1), np.uint64)
        self._ck(self.lib.fqsk_siv_test_shorter(self.h, _ptr(idx), _ptr(size_bits), len(idx), _ptr(out)))
        return out[: len(idx)]

    def mt_stream(self, n):
        out = np.zeros(n, np.uint32)
        self._ck(self.lib.fqsk_mt_stream(self.h, n, _ptr(out)))
        return out
