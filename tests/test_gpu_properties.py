"""Size-independent properties at BASELINE config-2 scale (p17/s20/b24, 51 000-read blocks, the reference's sync schedule), where
the CPU oracle is too slow to be the checker: (1) the ways into the engine -- reads resident in HBM (fqsk_segment_device, alone and with
the next segment announced by fqsk_announce_device), host buffers blocking (fqsk_segment + fqsk_sync), host buffers asynchronous
(fqsk_submit / fqsk_collect), a whole reads_block per call (fqsk_block_host) and a run of blocks with the last segment carried into the
next call (fqsk_block_stream) -- must produce the same records, tables and PRNG positions; (2) a run is a pure function of its input (two runs, identical checksums);
(3) conservation: every coded base yields exactly one record and the p-mer update count equals pushes + hidden updates."""
import hashlib

import numpy as np
import pytest

from fqsqueezer_b200 import engine as E
from fqsqueezer_b200 import schedule as S
from fqsqueezer_b200 import synth

pytestmark = pytest.mark.gpu

GS, GENOME, L, READS = 100, 2_000_000, 150, 51_000
BLOCKS = [0, 1, 97, 98, 150]      # block generations: 100 and 99 syncs per block, then 3, 2 and a single 51 000-read segment


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8)).hexdigest()


def _tables(e):
    out = {}
    for which in (0, 1, 2):
        k, v = e.dump(which)
        out[which] = (len(k), _digest(k), _digest(v))
    st = e.stats()
    return out, {k: st[k] for k in ("siv_no_filled", "siv_no_updates", "n_smers", "n_bmers", "draws_b", "draws_s", "draws_lb", "draws_ls")}


def _slab(codes):
    n = codes.shape[0]
    slab = np.empty((n, L + 1), np.uint8)
    slab[:, :L] = synth.codes_to_ascii(codes)
    slab[:, L] = 10
    return slab.reshape(-1), np.arange(n, dtype=np.uint64) * np.uint64(L + 1), np.full(n, L, np.uint32)


def _run(mode, blocks):
    import torch
    pref, p, s, b = E.kmer_params(GS)
    e = E.KmerEngine(p, s, b, pref, expected_kmers=1 << 24, reserve_reads=READS, reserve_bytes=READS * (L + 1))
    h = hashlib.sha256()
    n_recs = 0
    pend = None
    for g, codes in zip(BLOCKS, blocks):
        sched = list(S.segments(0, READS, S.calc_no_synchronizations(g, READS, 1)))
        slab, off, ln = _slab(codes)
        e.block_start()
        if mode in ("device", "device_announce"):
            d = torch.from_numpy(synth.codes_to_ascii(codes).reshape(-1)).cuda()
            d_off = torch.from_numpy(np.arange(READS, dtype=np.int64) * L).cuda()
            d_len = torch.full((READS,), L, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()      # the engine reads them on its own non-blocking stream (include/fqsk.h)
        if mode == "block":
            recs, dup, seg_off, seg_n = e.block_host(slab, off, ln, np.array([bb for _, bb in sched], np.uint32))
            for o, n in zip(seg_off.tolist(), seg_n.tolist()):
                h.update(np.ascontiguousarray(recs[o:o + n]).view(np.uint8)); n_recs += n
            continue
        if mode in ("stream", "stream_ctx"):
            for r2 in e.block_stream(slab, off, ln, np.array([bb for _, bb in sched], np.uint32), ctx=mode == "stream_ctx"):
                h.update(np.ascontiguousarray(r2).view(np.uint8)); n_recs += len(r2)
            continue
        for k, (a, bb) in enumerate(sched):
            if mode in ("device", "device_announce"):
                if mode == "device":
                    n = e.segment_device(d.data_ptr() + a * L, (bb - a) * L, d_off.data_ptr(), d_len.data_ptr(), bb - a)
                else:
                    e.segment_device(d.data_ptr() + a * L, (bb - a) * L, d_off.data_ptr(), d_len.data_ptr(), bb - a, want_n_recs=False)
                    if k + 1 < len(sched):
                        a2, b2 = sched[k + 1]
                        e.announce_device(d.data_ptr() + a2 * L, (b2 - a2) * L, d_off.data_ptr(), d_len.data_ptr(), b2 - a2)
                ptr, n2 = e.device_recs()
                if mode == "device":
                    assert n == n2
                n = n2
                buf = torch.empty(n * E.REC_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
                E.C.cdll.LoadLibrary("libcudart.so").cudaMemcpy(E.C.c_void_p(buf.data_ptr()), E.C.c_void_p(ptr), E.C.c_size_t(buf.numel()), 3)
                recs = buf.cpu().numpy().view(E.REC_DTYPE)
                e.sync()
            elif mode == "blocking":
                recs, dup = e.segment(slab, off[a:bb], ln[a:bb])
                e.sync()
            else:
                t = e.submit(slab, off[a:bb], ln[a:bb], ctx=mode == "async_ctx")
                if pend is not None:
                    r2, _, _ = e.collect(pend)
                    h.update(np.ascontiguousarray(r2).view(np.uint8)); n_recs += len(r2)
                pend = t
                continue
            h.update(np.ascontiguousarray(recs).view(np.uint8)); n_recs += len(recs)
    if pend is not None:
        r2, _, _ = e.collect(pend)
        h.update(np.ascontiguousarray(r2).view(np.uint8)); n_recs += len(r2)
    if mode in ("stream", "stream_ctx"):
        r2 = e.stream_finish()
        h.update(np.ascontiguousarray(r2).view(np.uint8)); n_recs += len(r2)
    tabs, st = _tables(e)
    e.close()
    return h.hexdigest(), n_recs, tabs, st


def test_entry_paths_agree_and_runs_are_deterministic():
    genome = synth.make_genome(GENOME, 5)
    blocks = [synth.make_reads(genome, READS, L=L, seed=100 + i)[0] for i in range(len(BLOCKS))]
    pref, p, s, b = E.kmer_params(GS)
    res = {m: _run(m, blocks) for m in ("device", "device_announce", "blocking", "async", "block", "stream", "async_ctx", "stream_ctx")}
    again = _run("async", blocks)
    assert res["device"] == res["device_announce"] == res["blocking"] == res["async"] == res["block"] == res["stream"] == again
    assert res["async_ctx"] == res["stream_ctx"] and res["async_ctx"][1:] == res["async"][1:]      # context records: same count, tables and PRNG positions
    digest, n_recs, tabs, st = res["async"]
    assert n_recs == len(BLOCKS) * READS * (L - pref)               # no duplicates in this stream: one record per coded suffix base
    assert st["siv_no_filled"] == tabs[0][0] and st["n_smers"] == tabs[1][0] and st["n_bmers"] == tabs[2][0]
    assert st["draws_b"] > 0 and tabs[2][0] > 1_000_000
