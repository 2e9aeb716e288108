"""Sharded operation over the GPUs of one box: one process per GPU, reference `-t N` semantics.

Rank r is the reference's worker thread r (application.cpp:575-671): it codes its own slice of every reads_block with its
own PRNG streams and thread-local tables, and it OWNS the k-mers whose routing key maps to it (dna.cpp:825, 836, 845,
2381-2388).  Lookups read every rank's shard through NVLink peer mappings (CUDA IPC, set up once here); at a sync the
exchange matrices X_to_add[src][dst] are written straight into the owners' inboxes by the routing kernel (peer stores) together
with the number of the sync, for which the owners wait on the device; the second barrier ("every owner has inserted") is a sequence
number as well and the global p-mer statistics are NVLink atomics into every rank's accumulators (fqsk_sync_device): torch.distributed
(NCCL) only carries the set-up (descriptor exchange) and, on request (host_collective=True), the three-step form with an all-reduce.

    grp = ShardedKmerEngine(p, s, b, prefix_len, rank, world, device=local_rank)   # after dist.init_process_group
    grp.block_start(); recs, dup = grp.segment(slab, off, ln); grp.sync()
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import engine as E


class _ShardDesc(C.Structure):
    _fields_ = [("rank", C.c_uint32), ("world_size", C.c_uint32), ("geometry", C.c_uint32 * 6), ("inbox_cap", C.c_uint64), ("ipc", (C.c_uint8 * 64) * 8)]


def owner_of_kmer(x: np.ndarray, world: int) -> np.ndarray:
    """Owner of normalised s-/b-mers: ((x >> 46) & 0x3fff) % T (dna.cpp:825, 836, 2382-2388)."""
    return ((np.asarray(x, np.uint64) >> np.uint64(46)) & np.uint64(0x3FFF)) % np.uint64(world)


def owner_of_pmer(idx: np.ndarray, pmer_len: int, world: int) -> np.ndarray:
    """Owner of aligned p-mers: (x >> (2p - 12)) % T (dna.cpp:658, 845, 2381)."""
    return (np.asarray(idx, np.uint64) >> np.uint64(2 * pmer_len - 12)) % np.uint64(world)


def gather_descs(dist, my_blob: bytes, world: int):
    """Every rank's shard descriptor, in rank order (the descriptors are plain bytes: any transport would do)."""
    blobs = [None] * world
    dist.all_gather_object(blobs, my_blob)
    assert all(isinstance(x, (bytes, bytearray)) and len(x) == C.sizeof(_ShardDesc) for x in blobs)
    return [_ShardDesc.from_buffer_copy(x) for x in blobs]


RESHARD = 1      # FQSK_RESHARD (include/fqsk.h): the sync is complete, table shards double before the next segment


def allreduce_stats(dist, fresh: int, updates: int, grow_request: int = 0):
    """Sum over the ranks of (fresh p-mer fields, p-mer updates): TSmallIntVector's global atomics (bit_vec.h:212-220) -- and, in the same
    all-reduce, how many ranks ask for a doubling of the s-mer / b-mer / pair table (fqsk_shard_grow_request): returned as the OR of the requests."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([fresh, updates] + [(grow_request >> i) & 1 for i in range(3)], dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    a, b, g0, g1, g2 = t.tolist()
    return int(a), int(b), (1 if g0 else 0) | (2 if g1 else 0) | (4 if g2 else 0)


class ShardedKmerEngine(E.KmerEngine):
    """KmerEngine of one rank of a sharded group.  `dist` is torch.distributed (initialised by the caller, NCCL on GPU boxes)."""

    def __init__(self, p, s, b, prefix_len, rank, world, device=0, dist=None, expected_kmers=0, reserve_reads=0, reserve_bytes=0, mode=E.MODE_SE_ORIGINAL, **kw):
        self.rank, self.world, self.dist = rank, world, dist
        self.host_collective = bool(kw.get("host_collective", False))      # True: the three-step form with the caller's all-reduce (NCCL / gloo)
        self.lib = E.load_library()
        self._bind()
        prm = E._Params(abi_version=1, pmer_len=p, smer_len=s, bmer_len=b, prefix_len=prefix_len, smer_counter_bits=12, bmer_counter_bits=6,
                        mode=mode, n_workers=world, device=device, bmer_log2_buckets=kw.get("bmer_log2_buckets", 0),
                        smer_log2_buckets=kw.get("smer_log2_buckets", 0), expected_kmers=expected_kmers, world_size=world, rank=rank,
                        max_iterations=0, flags=(E.F_PROFILE if kw.get("profile") else 0) | kw.get("flags", 0), reserve_reads=reserve_reads, reserve_bytes=reserve_bytes,
                        pair_log2_slots=kw.get("pair_log2_slots", 0))
        h = C.c_void_p()
        rc = self.lib.fqsk_create(C.byref(prm), C.byref(h))
        if rc != 0:
            raise E.FqskError(rc, self.lib.fqsk_last_error(None).decode())
        self.h = h
        self.p, self.s, self.b, self.prefix_len, self.mode = p, s, b, prefix_len, mode
        self.reserve_reads, self.reserve_bytes = reserve_reads, reserve_bytes
        if world > 1:
            self._attach_peers()

    def _bind(self):
        lib = self.lib
        vp = C.c_void_p
        lib.fqsk_shard_export.argtypes = [vp, C.POINTER(_ShardDesc)]
        lib.fqsk_shard_attach.argtypes = [vp, C.POINTER(_ShardDesc)]
        lib.fqsk_sync_route.argtypes = [vp]
        lib.fqsk_sync_apply.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.fqsk_sync_finish.argtypes = [vp, C.c_uint64, C.c_uint64]
        lib.fqsk_sync_device.argtypes = [vp]
        lib.fqsk_shard_grow_request.argtypes = [vp, C.POINTER(C.c_uint32)]
        lib.fqsk_shard_grow.argtypes = [vp, C.c_uint32]
        for n in ("fqsk_shard_export", "fqsk_shard_attach", "fqsk_sync_route", "fqsk_sync_apply", "fqsk_sync_finish", "fqsk_sync_device", "fqsk_shard_grow_request", "fqsk_shard_grow"):
            getattr(lib, n).restype = C.c_int

    def _attach_peers(self):
        d = _ShardDesc()
        self._ck(self.lib.fqsk_shard_export(self.h, C.byref(d)))
        for r, peer in enumerate(gather_descs(self.dist, bytes(d), self.world)):
            if r != self.rank:
                self._ck(self.lib.fqsk_shard_attach(self.h, C.byref(peer)))
        self.dist.barrier()

    def sync(self):
        """InsertKmersToHT + ClearKmersToHT of all workers (dna.cpp:2393-2488) around the reference's barriers."""
        if self.world == 1:
            return super().sync()
        if not self.host_collective:
            # rows into the owners' inboxes, both barriers as sequence numbers the ranks post in each other's inbox headers and wait for on
            # the device, global p-mer statistics as NVLink atomics: no collective library on the path of a sync
            rc = self.lib.fqsk_sync_device(self.h)
        else:
            self._ck(self.lib.fqsk_sync_route(self.h))              # rows [rank][*] into the owners' inboxes + this sync's number posted there
            fresh, upd = C.c_uint64(0), C.c_uint64(0)                 # (the owners wait for their sources on the device: no host barrier here)
            self._ck(self.lib.fqsk_sync_apply(self.h, C.byref(fresh), C.byref(upd)))
            req = C.c_uint32(0)
            self._ck(self.lib.fqsk_shard_grow_request(self.h, C.byref(req)))
            # global p-mer statistics; also the second barrier: every owner has finished its inserts before anybody looks up again
            f_all, u_all, grow = allreduce_stats(self.dist, fresh.value, upd.value, req.value)
            if grow:
                self._ck(self.lib.fqsk_shard_grow(self.h, grow))
            rc = self.lib.fqsk_sync_finish(self.h, f_all, u_all)
        if rc == RESHARD:
            # CHT_kmer::restruct (ht_kmer.h:88-112) for shards: some rank's shard is past half full, all shards of that table double.  Every
            # rank got the same verdict; nobody frees a table before all peers have closed their mappings of it (the barrier), the doubling
            # happens inside fqsk_shard_export, then the new descriptors are exchanged and attached as at set-up.
            self.reshards = getattr(self, "reshards", 0) + 1
            self.dist.barrier()
            self._attach_peers()
            return
        self._ck(rc)

    def dump_all(self, which):
        """Sorted contents of the whole (sharded) table, gathered on every rank -- parity check 1."""
        k, v = self.dump(which)
        if self.world == 1:
            return k, v
        parts = [None] * self.world
        self.dist.all_gather_object(parts, (k, v))
        kk = np.concatenate([x[0] for x in parts]); vv = np.concatenate([x[1] for x in parts])
        o = np.lexsort((vv, kk))         # by key, then value (the pair table holds several values per key)
        return kk[o], vv[o]
