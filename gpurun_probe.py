import sys, time, numpy as np
sys.path.insert(0, '.')
from fqsqueezer_b200 import engine as E, synth, schedule as S
pref,p,s,b = E.kmer_params(100)
G = int(sys.argv[1]) if len(sys.argv)>1 else 100_000_000
nblocks = int(sys.argv[2]) if len(sys.argv)>2 else 6
steady_after = int(sys.argv[3]) if len(sys.argv)>3 else 10**9
every = int(sys.argv[4]) if len(sys.argv)>4 else 1
t=time.time(); genome = synth.make_genome(G, 43); print('genome', time.time()-t, flush=True)
e = E.KmerEngine(p,s,b,pref, expected_kmers=1<<27, profile=True)
print('engine', time.time()-t, flush=True)
L=150; per_block=51000
from tests.test_gpu_segment import _fastq_slab
tot=0; t0=time.time(); prev={}
for g in range(nblocks):
    codes,_ = synth.make_reads(genome, per_block, L=L, seed=1000+g)
    slab = _fastq_slab(codes)
    off,ln,roff,rsz = S.parse_fastq(slab)
    ns = S.calc_no_synchronizations(g if g < steady_after else 100, per_block, 1)
    e.block_start()
    tb=time.time()
    for a,bb in S.segments(0, per_block, ns):
        e.segment(slab, off[a:bb], ln[a:bb]); e.sync()
    dt=time.time()-tb
    st=e.stats()
    if g % every == 0 or g == nblocks-1: print(f'block {g}: {dt*1e3:.1f} ms  {per_block*L/dt/1e6:.1f} Mbases/s  replays/seg={st["n_replays"]/st["n_segments"]:.2f} launches={st["kernel_launches"]} bmers={st["n_bmers"]} stash={st["bmer_stash_used"]}', flush=True)
    pr=e.profile()
    if g % every == 0 or g == nblocks-1: print('   ', {k: round(v-prev.get(k,0),1) for k,v in pr.items()}, flush=True)
    prev=pr
