"""Timeline (fqsk_timeline) of two consecutive sync segments of the sharded job on rank 0; run under torch.distributed.run with N ranks."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
from fqsqueezer_b200 import engine as E, schedule as S, synth, sharded as SH
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
pref, p, s, b = E.kmer_params(B.GS)
reads = B.JobReads(synth.make_genome(B.GENOME, B.SEED))
blocks = B.job_blocks()
rr = B.RESERVE_READS
eng = SH.ShardedKmerEngine(p, s, b, pref, rank, world, device=local, dist=dist, expected_kmers=(1 << 29) // world, reserve_reads=rr // world + 16, reserve_bytes=(rr // world + 16) * (B.L + 1))
G = int(os.environ.get("TL_BLOCK", "2"))
d_off = torch.arange(rr + 16, dtype=torch.int64, device=dev) * B.L
d_len = torch.full((rr + 16,), B.L, dtype=torch.int32, device=dev)
for g in range(G + 1):
    f, l = blocks[g]
    sched = S.worker_segments(0, l - f, g, world, rank)
    lo, hi = sched[0][0], sched[-1][1]
    t = torch.from_numpy(synth.codes_to_ascii(reads.codes(f + lo, f + hi)).reshape(-1)).to(dev)
    torch.cuda.synchronize()
    eng.block_start()
    for k, (a, bb) in enumerate(sched):
        if g == G and k == 40 and rank == 0:
            eng.timeline(True)
        eng.segment_device(t.data_ptr() + (a - lo) * B.L, (bb - a) * B.L, d_off.data_ptr(), d_len.data_ptr(), bb - a, want_n_recs=False)
        eng.sync()
        if g == G and k == 41 and rank == 0:
            eng.timeline(False)
eng.close()
dist.barrier()
dist.destroy_process_group()
