"""One rank of the sharded GPU parity run (launched by tests/test_gpu_sharded.py through torch.distributed.run):
rank r plays the reference's worker thread r on fixture <name> (produced by the tapped reference at -t world) and checks
its per-base records, and -- merged over the ranks -- the final tables, bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fqsqueezer_b200 import engine as E  # noqa: E402
from fqsqueezer_b200 import schedule as S  # noqa: E402
from fqsqueezer_b200 import sharded  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    name = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = H.load_golden(name)
    assert int(g["threads"]) == world, (int(g["threads"]), world)
    pref, p, s, b = E.kmer_params(int(g["gs"]))
    slab = g["fastq"]
    paired = "-p" in [str(x) for x in g["extra"]]
    sorted_order = [str(x) for x in g["extra"]][-2:] == ["-om", "s"]      # the reference's default order: sorted bins, sorted-prefix coding
    mode = (E.MODE_PE_SORTED if sorted_order else E.MODE_PE_ORIGINAL) if paired else (E.MODE_SE_SORTED if sorted_order else E.MODE_SE_ORIGINAL)
    grow = "grow" in sys.argv[2:]      # smallest legal tables + FQSK_F_TEST_CROWD: the shards have to double, together, several times during the run
    kw = dict(bmer_log2_buckets=1, smer_log2_buckets=1, pair_log2_slots=10, flags=E.F_TEST_HOOKS | E.F_TEST_CROWD) if grow else {}
    eng = sharded.ShardedKmerEngine(p, s, b, pref, rank, world, device=local, dist=dist, reserve_bytes=1 << 20, reserve_reads=1 << 14,
                                    mode=mode, host_collective="host" in sys.argv[2:], **kw)
    off, ln, roff, rsz = S.parse_fastq(slab)
    out, info, flags, difs = [], [], [], []
    blocks = H.sorted_blocks(slab, off, rsz, paired) if sorted_order else ((gen, f, l) for gen, (f, l) in enumerate(S.split_blocks(rsz, paired=paired)))
    for gen, f, l in blocks:
        eng.block_start()
        for a, bb in S.worker_segments(f, l, gen, world, rank, paired=paired):
            recs, dup = eng.segment(slab, off[a:bb], ln[a:bb])
            out.append(recs)
            if paired:
                info.append(eng.pair_info((bb - a) // 2))
            if sorted_order:      # what compress_prefix_sorted codes per read (per first mate) that is not a duplicate
                fl, df = eng.sorted_prefix(bb - a)
                keep = (dup == 0) & ((np.arange(bb - a) % 2 == 0) if paired else True)
                flags.append(fl[keep]); difs.append(df[keep])
            eng.sync()
    recs = np.concatenate(out)
    want = g["recs_t%d" % rank]
    H.assert_recs_equal(recs, want[want["pos"] < 0xFFFFFFF0])
    if paired:      # what CompressPE codes per pair: (candidate list exists, minimizer id, its position in mate 2)
        winfo = want[want["pos"] == H.POS_PAIR]["c"][:, :3].astype(np.uint32)
        assert np.array_equal(np.concatenate(info), winfo)
    if sorted_order:
        _, wflags, wdifs = H.split_sorted_markers(want, False)
        assert np.array_equal(np.concatenate(flags), wflags) and np.array_equal(np.concatenate(difs), wdifs)
        assert (wdifs > 0).sum() > 50
    for which, nm in ((0, "siv"), (1, "smer"), (2, "bmer")) + (((3, "pair"),) if paired else ()):
        k, v = eng.dump_all(which)
        assert np.array_equal(k, g[nm + "_keys"]), (nm, len(k), len(g[nm + "_keys"]))
        assert np.array_equal(v, g[nm + "_vals"]), nm
    st = eng.stats()
    assert st["siv_no_filled"] == int(g["siv_no_filled"]) and st["siv_no_updates"] == int(g["siv_no_updates"]), (st["siv_no_filled"], st["siv_no_updates"])
    assert st["kernel_launches"] > 0
    if grow:      # CHT_kmer::restruct for shards: FQSK_RESHARD -> barrier -> doubling inside fqsk_shard_export -> descriptors exchanged and attached again
        assert getattr(eng, "reshards", 0) >= 2 and st["n_table_growths"] >= 4, (getattr(eng, "reshards", 0), st["n_table_growths"])
    print(f"rank {rank}/{world}: {len(recs)} records and the merged tables bit-exact vs fqs-1.1 -t {world} ({st['kernel_launches']} launches, {getattr(eng, 'reshards', 0)} coordinated table doublings)", flush=True)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
