"""Measurement helper (not part of the product): wall clock per C-ABI call in the small-segment regime of config 2, for the
device-resident path (fqsk_segment_device + fqsk_sync) and the host-buffer path (fqsk_submit / fqsk_collect), plus a few
full-size segments.  Run on the GPU box: `python profiles/probe_latency.py [n_blocks]`.  The numbers quoted in
profiles/r01c_early_block_warm_summary.md and DESIGN.md section 6 come from it."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from fqsqueezer_b200 import engine as E, schedule as S, synth

pref, p, s, b = E.kmer_params(B.GS)
genome = synth.make_genome(B.GENOME, B.SEED)
dev = torch.device("cuda", 0)
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 24
eng = E.KmerEngine(p, s, b, pref, expected_kmers=1 << 29, reserve_reads=B.READS_PER_BLOCK, reserve_bytes=B.READS_PER_BLOCK * B.L)
d_off = torch.from_numpy(np.arange(B.READS_PER_BLOCK, dtype=np.int64) * B.L).to(dev)
d_len = torch.full((B.READS_PER_BLOCK,), B.L, dtype=torch.int32, device=dev)
blocks = [torch.from_numpy(synth.codes_to_ascii(B.block_codes(genome, g, 0)).reshape(-1)).to(dev) for g in range(NB)]
torch.cuda.synchronize()
for g in range(NB):
    sched = list(S.segments(0, B.READS_PER_BLOCK, S.calc_no_synchronizations(g, B.READS_PER_BLOCK, 1)))
    eng.block_start()
    base = blocks[g].data_ptr()
    t_seg = t_sync = 0.0
    t0 = time.perf_counter()
    for a, bb in sched:
        n = bb - a
        t1 = time.perf_counter()
        eng.segment_device(base + a * B.L, n * B.L, d_off.data_ptr(), d_len.data_ptr(), n, want_n_recs=False)
        t2 = time.perf_counter()
        eng.sync()
        t3 = time.perf_counter()
        t_seg += t2 - t1; t_sync += t3 - t2
    tot = time.perf_counter() - t0
    if g >= 4:
        print(f"block {g}: {len(sched)} segs, per segment: total {1e6 * tot / len(sched):.0f} us, segment_device (enqueue only) {1e6 * t_seg / len(sched):.0f} us, sync (enqueue + look) {1e6 * t_sync / len(sched):.0f} us")
st = eng.stats()
print({k: st[k] for k in ("n_segments", "n_replays", "kernel_launches", "n_hot_segments")})

# ---- e2e path (host buffers through fqsk_submit / fqsk_collect) ----
eng.close()
eng2 = E.KmerEngine(p, s, b, pref, expected_kmers=1 << 29, reserve_reads=B.READS_PER_BLOCK, reserve_bytes=B.READS_PER_BLOCK * (B.L + 1))
pend = None
for g in list(range(NB)) + [NB, NB + 1, NB + 2, NB + 3, NB + 4]:
    steady = g >= NB
    gg = 120 + g - NB if steady else g
    slab, off, ln = B.codes_to_slab(B.block_codes(genome, gg, 0))
    sched = [(0, B.READS_PER_BLOCK)] if steady else list(S.segments(0, B.READS_PER_BLOCK, S.calc_no_synchronizations(g, B.READS_PER_BLOCK, 1)))
    eng2.block_start()
    t_sub = t_col = 0.0
    t0 = time.perf_counter()
    for a, bb in sched:
        t1 = time.perf_counter()
        t = eng2.submit(slab, off[a:bb], ln[a:bb])
        t2 = time.perf_counter()
        if pend is not None:
            eng2.collect(pend)
        t3 = time.perf_counter()
        pend = t
        t_sub += t2 - t1; t_col += t3 - t2
    tot = time.perf_counter() - t0
    if g >= NB - 3:
        print(f"e2e block {gg}: {len(sched)} segs, per segment: total {1e6 * tot / len(sched):.0f} us, submit {1e6 * t_sub / len(sched):.0f} us, collect {1e6 * t_col / len(sched):.0f} us")
eng2.collect(pend)
