"""Host-side logic of the sharded (N > 1) path on CPU: owner routing keys, per-worker schedule, descriptor exchange and the
statistics all-reduce over torch.distributed with the gloo backend (world_size 2).  No GPU, no compute calls into libfqsk."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from fqsqueezer_b200 import schedule as S
from fqsqueezer_b200 import sharded
from oracle import oracle as O
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_owner_keys_reproduce_the_reference_exchange_matrix():
    """Route every worker's pending rows with sharded.owner_of_* and apply them owner by owner, source by source, through the
    table-level oracle: the result must equal OracleGroup.sync (pinned against fqs-1.1 -t 2 by test_golden_oracle)."""
    g = H.load_golden("se_orig_gs1_t2")
    T = int(g["threads"])
    pref, p, s, b = O.kmer_params(int(g["gs"]))
    slab = g["fastq"]
    off, ln, roff, rsz = S.parse_fastq(slab)
    grp = O.OracleGroup(p, s, b, pref, T)
    f, l = S.split_blocks(rsz)[0]
    segs = [S.worker_segments(f, l, 0, T, w) for w in range(T)]
    assert len({len(x) for x in segs}) == 1
    for w in grp.workers:
        w.block_start()
    rows = []
    for wi, w in enumerate(grp.workers):
        a, bb = segs[wi][0]
        w.segment(slab, off[a:bb], ln[a:bb])
        rows.append({k: w.pending(i) for i, k in enumerate("psb")})
    # matrix [src][dst] in push order
    mat = {k: [[rows[i][k][(sharded.owner_of_pmer(rows[i][k], p, T) if k == "p" else sharded.owner_of_kmer(rows[i][k], T)) == j] for j in range(T)] for i in range(T)] for k in "psb"}
    for k in "psb":
        assert sum(len(mat[k][i][j]) for i in range(T) for j in range(T)) == sum(len(rows[i][k]) for i in range(T))
        assert all(len(mat[k][i][j]) > 0 for i in range(T) for j in range(T)), "fixture too small to exercise every [src][dst] cell"
    # owners apply their column in source order to fresh tables with fresh incrementers (first sync: both start empty)
    U = O.oracle_units()
    want_b = {}
    for j in range(T):
        t, ci = U.ht_new(b, 6), U.cinc_new(7, 2, 63)
        U.ht_insert(t, ci, np.concatenate([mat["b"][i][j] for i in range(T)]))
        kk, vv = U.ht_dump(t)
        want_b.update(zip(kk.tolist(), vv.tolist()))
    grp.sync()
    kg, vg = grp.dump(2)
    assert dict(zip(kg.tolist(), vg.tolist())) == want_b
    grp.close()


def test_worker_schedule_tiles_the_block():
    for n_reads, T, gen in [(51000, 2, 0), (51000, 8, 3), (1001, 3, 99), (640, 4, 150), (7, 2, 0)]:
        segs = [S.worker_segments(100, 100 + n_reads, gen, T, w) for w in range(T)]
        assert len({len(x) for x in segs}) == 1
        flat = [x for w in segs for x in w]
        assert flat[0][0] == 100 and flat[-1][1] == 100 + n_reads
        pos = 100
        for w in segs:
            for a, b in w:
                assert a == pos and b >= a
                pos = b
        assert len(segs[0]) == S.calc_no_synchronizations(gen, n_reads, T) + 1


def _gloo_rank(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = sharded._ShardDesc()
        d.rank, d.world_size, d.inbox_cap = rank, world, 1234 + rank
        for qi in range(6):
            for k in range(64):
                d.ipc[qi][k] = (17 * rank + 3 * qi + k) & 0xFF
        got = sharded.gather_descs(dist, bytes(d), world)
        ok = all(x.rank == r and x.inbox_cap == 1234 + r and x.ipc[5][63] == (17 * r + 15 + 63) & 0xFF for r, x in enumerate(got))
        tot = sharded.allreduce_stats(dist, 10 + rank, 100 * (rank + 1))
        ok = ok and tot == (sum(10 + r for r in range(world)), sum(100 * (r + 1) for r in range(world)), 0)
        # growth requests ride in the same all-reduce and come back as the OR over the ranks: rank 0 asks for the b-mer table, rank 1 for the pair table
        tot = sharded.allreduce_stats(dist, 1, 1, 2 if rank == 0 else 4 if rank == 1 else 0)
        ok = ok and tot == (world, world, 6)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_descriptor_exchange_and_stat_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_gloo_rank, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for pr in procs:
        pr.join(60)
    assert res == [(0, True), (1, True)]
