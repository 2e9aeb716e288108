"""Measurement helper: wall clock per block / per segment of the device-resident path (fqsk_segment_device + fqsk_sync) for a
few block generations of config 2 -- 100 syncs per block, ..., one 51 000-read segment.  Tables stay young in this short run, so the
large segments do more rough searches than in the real job; use bench.py --trace-blocks for the steady state."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from fqsqueezer_b200 import engine as E, schedule as S, synth
pref, p, s, b = E.kmer_params(B.GS)
genome = synth.make_genome(B.GENOME, B.SEED)
dev = torch.device("cuda", 0)
eng = E.KmerEngine(p, s, b, pref, expected_kmers=1 << 29, reserve_reads=B.READS_PER_BLOCK, reserve_bytes=B.READS_PER_BLOCK * B.L)
d_off = torch.from_numpy(np.arange(B.READS_PER_BLOCK, dtype=np.int64) * B.L).to(dev)
d_len = torch.full((B.READS_PER_BLOCK,), B.L, dtype=torch.int32, device=dev)
for g in [0, 1, 60, 80, 90, 95, 97, 120, 121, 122, 123, 124, 125]:
    blk = torch.from_numpy(synth.codes_to_ascii(B.block_codes(genome, g, 0)).reshape(-1)).to(dev)
    sched = list(S.segments(0, B.READS_PER_BLOCK, S.calc_no_synchronizations(g, B.READS_PER_BLOCK, 1)))
    torch.cuda.synchronize()
    eng.block_start()
    t0 = time.perf_counter()
    for a, bb in sched:
        n = bb - a
        eng.segment_device(blk.data_ptr() + a * B.L, n * B.L, d_off.data_ptr(), d_len.data_ptr(), n, want_n_recs=False)
        eng.sync()
    tot = time.perf_counter() - t0
    print(f"block {g}: {len(sched)} segs of {sched[0][1]-sched[0][0]} reads, {1e3*tot:.2f} ms, per segment {1e6*tot/len(sched):.0f} us", flush=True)
