"""GPU parity, segment level: per-base (counts, level, rough, cor_pos) records, duplicate flags and final tables of the
CUDA engine against (1) golden fixtures produced by the real reference (tapped fqs-1.1) and (2) the CPU oracle on seeded
synthetic reads.  Bit-exact.  All calls go through the C-ABI (fqsk_segment / fqsk_sync / fqsk_dump)."""
import numpy as np
import pytest

from fqsqueezer_b200 import engine as E
from fqsqueezer_b200 import synth
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["se_orig_gs1", "se_orig_gs100", "se_orig_repeats_gs1", "se_mixed_unc_gs1", "se_orig_gs3100"])
def test_engine_matches_reference_golden(name):
    """se_mixed_unc_gs1: the fixture in which counts_level `mixed` (dna.cpp:470-478; > 5 000 records) and the bmer_unc revert
    (dna.cpp:697-705) occur; se_orig_gs3100: the reference's default k-mer lengths (p18/s21/b27, prefix 13; BASELINE configs 1, 4, 5)."""
    g = H.load_golden(name)
    pref, p, s, b = E.kmer_params(int(g["gs"]))
    e = E.KmerEngine(p, s, b, pref)
    recs = H.run_se(e, g["fastq"])
    want = g["recs"]
    want = want[want["pos"] < 0xFFFFFFF0]
    H.assert_recs_equal(recs, want)
    H.assert_dump_equal(e, g)
    if name == "se_orig_repeats_gs1":    # low-complexity reads: the thread-local PRNG streams (cinc_lb) must have been used
        st = e.stats()
        assert st["n_hot_segments"] > 0 and st["draws_lb"] > 0
    if name == "se_mixed_unc_gs1":
        assert (recs["level"] == 4).sum() > 1000
    e.close()


def _fastq_slab(codes):
    """Minimal FASTQ slab for codes[n, L] (ids and qualities are irrelevant to the k-mer path)."""
    n, L = codes.shape
    seq = synth.codes_to_ascii(codes)
    rec = np.empty((n, 3 + L + 3 + L + 1), np.uint8)
    rec[:, 0] = ord("@"); rec[:, 1] = ord("r"); rec[:, 2] = 10
    rec[:, 3:3 + L] = seq
    rec[:, 3 + L] = 10; rec[:, 4 + L] = ord("+"); rec[:, 5 + L] = 10
    rec[:, 6 + L:6 + 2 * L] = ord("I")
    rec[:, 6 + 2 * L] = 10
    return rec.reshape(-1)


@pytest.mark.parametrize("gs,G,n_reads,L,seed,nfrac,dupfrac", [
    (1, 5000, 3000, 60, 11, 0.003, 0.02),      # tiny k, heavy coverage: counters > thr, repairs, rough, avg_filling_factor gate
    (16, 60000, 6000, 100, 12, 0.0, 0.0),
    (100, 30000, 2500, 150, 13, 0.001, 0.0),   # BASELINE config-2 k-mer lengths
])
def test_engine_matches_oracle(gs, G, n_reads, L, seed, nfrac, dupfrac):
    genome = synth.make_genome(G, seed)
    codes, _ = synth.make_reads(genome, n_reads, L=L, seed=seed, n_frac=nfrac, dup_frac=dupfrac)
    slab = _fastq_slab(codes)
    pref, p, s, b = E.kmer_params(gs)
    e = E.KmerEngine(p, s, b, pref, bmer_log2_buckets=12, smer_log2_buckets=12)   # small tables: growth + stash are exercised
    o = O.OracleEngine(p, s, b, pref)
    got = H.run_se(e, slab)
    want = H.run_se(o, slab)
    want = want[want["pos"] < 0xFFFFFFF0]
    H.assert_recs_equal(got, want)
    for which in (0, 1, 2):
        kg, vg = e.dump(which)
        ko, vo = o.dump(which)
        assert np.array_equal(kg, ko) and np.array_equal(vg, vo), which
    sg, so = e.stats(), o.stats()
    for key in ("siv_no_filled", "siv_no_updates", "n_smers", "n_bmers", "draws_b", "draws_s", "draws_lb", "draws_ls"):
        assert sg[key] == so[key], key
    e.close(); o.close()


def test_edge_cases_empty_and_ragged():
    """Empty segments, reads shorter than the directly coded prefix / than k, a block made only of duplicates."""
    pref, p, s, b = E.kmer_params(1)
    e = E.KmerEngine(p, s, b, pref)
    o = O.OracleEngine(p, s, b, pref)
    rng = np.random.default_rng(3)
    lens = [0, 1, 5, 9, 10, 13, 14, 17, 18, 19, 20, 25, 40, 40, 40, 7, 80]
    parts, off, ln = [], [], []
    at = 0
    for i, L in enumerate(lens):
        seq = rng.integers(0, 4, L)
        if i in (13, 14):
            seq = prev
        prev = seq
        body = b"@r\n" + bytes(b"ACGT"[c] for c in seq) + b"\n+\n" + b"I" * L + b"\n"
        off.append(at + 3); ln.append(L)
        parts.append(body); at += len(body)
    slab = np.frombuffer(b"".join(parts) + b"\n" * 64, np.uint8)
    off = np.array(off, np.uint64); ln = np.array(ln, np.uint32)
    for eng in (e, o):
        eng.block_start()
    for a, bb in ((0, 0), (0, 7), (7, 7), (7, len(lens))):
        rg, dg = e.segment(slab, off[a:bb], ln[a:bb])
        ro, do = o.segment(slab, off[a:bb], ln[a:bb])
        ro = ro[ro["pos"] < 0xFFFFFFF0]
        H.assert_recs_equal(rg, ro)
        assert np.array_equal(dg, do)
        e.sync(); o.sync()
    for which in (0, 1, 2):
        kg, vg = e.dump(which)
        ko, vo = o.dump(which)
        assert np.array_equal(kg, ko) and np.array_equal(vg, vo)
    e.close(); o.close()


def test_engine_matches_reference_golden_sorted_order():
    """-om s (row a15): compress_prefix_sorted's siv test + linear scan, p-mer pushes of the prefix, suffix from p_len."""
    g = H.load_golden("se_sorted_gs1")
    pref, p, s, b = E.kmer_params(int(g["gs"]))
    e = E.KmerEngine(p, s, b, pref, mode=E.MODE_SE_SORTED)
    recs, flags, difs = H.run_sorted(e, g["fastq"], is_gpu=True)
    want, wflags, wdifs = H.golden_sorted_expect(g)
    H.assert_recs_equal(recs, want)
    assert np.array_equal(flags, wflags) and np.array_equal(difs, wdifs)
    H.assert_dump_equal(e, g)
    e.close()


def test_engine_matches_reference_golden_paired_end():
    """-p -om o (rows a10, a17): pair table, window minimizers, candidate ranking, mate 2 coded from a shared minimizer
    (forward part + reversed left part) -- records, per-pair (found, id, pos) and all four tables vs the tapped reference."""
    g = H.load_golden("pe_orig_gs1")
    pref, p, s, b = E.kmer_params(int(g["gs"]))
    e = E.KmerEngine(p, s, b, pref, mode=E.MODE_PE_ORIGINAL)
    recs, info = H.run_pe(e, g["fastq"], is_gpu=True)
    want, winfo = H.golden_pe_expect(g)
    assert np.array_equal(info, winfo), np.flatnonzero((info != winfo).any(axis=1))[:5]
    H.assert_recs_equal(recs, want)
    H.assert_dump_equal(e, g, pairs=True)
    e.close()


def test_engine_matches_reference_golden_paired_end_default_kmer_lengths():
    """-p -om o at -gs 3100 (p18/s21/b27, prefix 13): BASELINE config 5's k-mer lengths."""
    g = H.load_golden("pe_orig_gs3100")
    pref, p, s, b = E.kmer_params(int(g["gs"]))
    e = E.KmerEngine(p, s, b, pref, mode=E.MODE_PE_ORIGINAL)
    recs, info = H.run_pe(e, g["fastq"], is_gpu=True)
    want, winfo = H.golden_pe_expect(g)
    assert np.array_equal(info, winfo), np.flatnonzero((info != winfo).any(axis=1))[:5]
    H.assert_recs_equal(recs, want)
    H.assert_dump_equal(e, g, pairs=True)
    e.close()


def test_engine_matches_reference_golden_paired_end_sorted_order():
    """-p in the reference's default order (-om s; BASELINE configs 3 and 5 as written): mate 1 through the sorted prefix (flag / dif,
    suffix from p_len: dna.cpp:1793-1796), mate 2 as in original order; pairs binned and sorted by mate 1."""
    g = H.load_golden("pe_sorted_gs1")
    pref, p, s, b = E.kmer_params(int(g["gs"]))
    e = E.KmerEngine(p, s, b, pref, mode=E.MODE_PE_SORTED)
    recs, info, flags, difs = H.run_pe_sorted(e, g["fastq"], is_gpu=True)
    want, winfo, wflags, wdifs = H.golden_pe_sorted_expect(g)
    assert np.array_equal(info, winfo), np.flatnonzero((info != winfo).any(axis=1))[:5]
    assert np.array_equal(flags, wflags) and np.array_equal(difs, wdifs)
    H.assert_recs_equal(recs, want)
    H.assert_dump_equal(e, g, pairs=True)
    e.close()


def _pe_slab(genome, n_pairs, L, seed, nfrac=0.0, dup_every=0):
    """Interleaved FASTQ slab of synthetic pairs: mate 2 = reverse complement of the fragment end."""
    rng = np.random.default_rng(seed)
    G = len(genome)
    frag = np.clip(rng.normal(2.6 * L, 0.25 * L, n_pairs).astype(np.int64), L, G - 1)
    start = rng.integers(0, G - frag)
    idx = np.arange(L)
    m1 = genome[start[:, None] + idx[None, :]]
    m2 = 3 - genome[(start + frag)[:, None] - 1 - idx[None, :]]
    for m in (m1, m2):
        err = rng.random(m.shape) < 0.005
        m[err] = (m[err] + rng.integers(1, 4, int(err.sum()))) & 3
    codes = np.empty((2 * n_pairs, L), np.int64)
    codes[0::2] = m1; codes[1::2] = m2
    if nfrac:
        codes[rng.random(codes.shape) < nfrac] = 4
    if dup_every:
        codes[2 * dup_every::2 * dup_every] = codes[2 * dup_every - 2:-2:2 * dup_every]    # mate 1 equal to the previous mate 1
    return _fastq_slab(codes)


@pytest.mark.parametrize("gs,G,n_pairs,L,seed,nfrac,dup_every", [
    (1, 6000, 2500, 70, 21, 0.002, 40),       # heavy coverage: long candidate lists, saturating pair counters, duplicates, Ns
    (100, 40000, 1500, 150, 22, 0.0, 0),      # BASELINE config-3 k-mer lengths
])
def test_engine_matches_oracle_paired_end(gs, G, n_pairs, L, seed, nfrac, dup_every):
    genome = synth.make_genome(G, seed)
    slab = _pe_slab(genome, n_pairs, L, seed, nfrac, dup_every)
    pref, p, s, b = E.kmer_params(gs)
    e = E.KmerEngine(p, s, b, pref, mode=E.MODE_PE_ORIGINAL, bmer_log2_buckets=12, smer_log2_buckets=12)
    o = O.OracleEngine(p, s, b, pref, mode=2)
    got, ginfo = H.run_pe(e, slab, is_gpu=True)
    want, winfo = H.run_pe(o, slab)
    assert np.array_equal(ginfo, winfo), np.flatnonzero((ginfo != winfo).any(axis=1))[:5]
    assert (winfo[:, 0] == 1).sum() > n_pairs // 4 and ((winfo[:, 0] == 1) & (winfo[:, 1] < 15)).sum() > n_pairs // 8
    H.assert_recs_equal(got, want)
    for which in (0, 1, 2, 3):
        kg, vg = e.dump(which)
        ko, vo = o.dump(which)
        assert np.array_equal(kg, ko) and np.array_equal(vg, vo), which
    sg, so = e.stats(), o.stats()
    for key in ("siv_no_filled", "siv_no_updates", "n_smers", "n_bmers", "draws_b", "draws_s", "draws_lb", "draws_ls"):
        assert sg[key] == so[key], key
    e.close(); o.close()


@pytest.mark.parametrize("name", ["se_orig_gs1", "se_orig_repeats_gs1"])
def test_async_submit_collect_matches_reference_golden(name):
    """fqsk_submit / fqsk_collect (double-buffered records on a copy stream, sync enqueued behind the segment, two tickets in
    flight) must give the same bytes as the blocking calls: records, duplicate flags and tables vs the tapped reference.
    The repeats fixture forces segments that need more than their first pass (records are copied a second time)."""
    g = H.load_golden(name)
    pref, p, s, b = E.kmer_params(int(g["gs"]))
    e = E.KmerEngine(p, s, b, pref)
    recs, dup = H.run_async(e, g["fastq"])
    want = g["recs"]
    n_dup_want = int((want["pos"] == O.POS_DUP).sum())
    want = want[want["pos"] < 0xFFFFFFF0]
    H.assert_recs_equal(recs, want)
    assert int(dup.sum()) == n_dup_want
    H.assert_dump_equal(e, g)
    e.close()


def test_async_submit_collect_paired_end():
    g = H.load_golden("pe_orig_gs1")
    pref, p, s, b = E.kmer_params(int(g["gs"]))
    e = E.KmerEngine(p, s, b, pref, mode=E.MODE_PE_ORIGINAL)
    recs, dup, info = H.run_async(e, g["fastq"], paired=True)
    want, winfo = H.golden_pe_expect(g)
    assert np.array_equal(info, winfo)
    H.assert_recs_equal(recs, want)
    assert dup[1::2].sum() == 0
    H.assert_dump_equal(e, g, pairs=True)
    e.close()


@pytest.mark.parametrize("G,expect_hot", [(2_000_000, False), (300_000, True)])
def test_large_segments_filtered_delta_matches_oracle(G, expect_hot):
    """Segments above 1 MiB of DNA build their thread-local delta through the core filter (only pushes a lookup can ask for are
    inserted) and sync through the radix-sort path: records, tables and PRNG positions vs the oracle, on reads with thread-local
    hits, repairs, draws, duplicates and Ns.  The small genome pushes k-mers more than thr + 1 times inside one segment: the
    filtered attempt must hand over to the ordered thread-local evaluator (full delta) without leaving a trace."""
    genome = synth.make_genome(G, 31)
    codes, _ = synth.make_reads(genome, 24_600, L=150, seed=31, n_frac=0.0005, dup_frac=0.002)
    slab = _fastq_slab(codes)
    off, ln, _, _ = __import__("fqsqueezer_b200.schedule", fromlist=["x"]).parse_fastq(slab)
    pref, p, s, b = E.kmer_params(16)
    e = E.KmerEngine(p, s, b, pref, expected_kmers=1 << 21)
    o = O.OracleEngine(p, s, b, pref)
    for eng in (e, o):
        eng.block_start()
    for a, bb in ((0, 8200), (8200, 16400), (16400, 24600)):
        rg, dg = e.segment(slab, off[a:bb], ln[a:bb])
        ro, do = o.segment(slab, off[a:bb], ln[a:bb])
        ro = ro[ro["pos"] < 0xFFFFFFF0]
        H.assert_recs_equal(rg, ro)
        assert np.array_equal(dg, do)
        e.sync(); o.sync()
    for which in (0, 1, 2):
        kg, vg = e.dump(which)
        ko, vo = o.dump(which)
        assert np.array_equal(kg, ko) and np.array_equal(vg, vo), which
    sg, so = e.stats(), o.stats()
    for key in ("siv_no_filled", "siv_no_updates", "n_smers", "n_bmers", "draws_b", "draws_s", "draws_lb", "draws_ls"):
        assert sg[key] == so[key], key
    assert sg["n_filtered_segments"] == 3 and so["local_hits"] > 100
    assert (sg["n_hot_segments"] > 0) == expect_hot
    e.close(); o.close()


@pytest.mark.parametrize("var", ["test_fail_every", "test_retry_every"])
@pytest.mark.parametrize("mode", ["blocking", "async"])
def test_recovery_paths_of_the_enqueued_sync(var, mode):
    """Fault injection: every 3rd segment has its first-pass verdict forced to 'not settled' AFTER the early grouping half of its
    b-mer sync has run (claimed slots must be released, the sync redone the plain way), resp. is evaluated again from scratch as
    after a capacity overflow (claimed slots must be released BEFORE the tables are read again).  Output must not change."""
    g = H.load_golden("se_orig_gs1")
    pref, p, s, b = E.kmer_params(int(g["gs"]))
    e = E.KmerEngine(p, s, b, pref, **{var: 3})      # FQSK_F_TEST_HOOKS + fqsk_params.test_hooks
    if mode == "async":
        recs, _ = H.run_async(e, g["fastq"])
    else:
        recs = H.run_se(e, g["fastq"])
    want = g["recs"]
    H.assert_recs_equal(recs, want[want["pos"] < 0xFFFFFFF0])
    H.assert_dump_equal(e, g)
    e.close()


@pytest.mark.parametrize("name,mode", [("se_orig_gs1", "blocking"), ("se_orig_gs1", "async"), ("se_orig_repeats_gs1", "blocking")])
def test_tables_double_between_segments(name, mode):
    """CHT_kmer::restruct (ht_kmer.h:88-112, ht_kmer.cpp:50-75) in the middle of a run: with FQSK_F_TEST_CROWD the s-mer and b-mer tables
    count as crowded at 1/64 of the usual load, so they double several times between the sync segments of a small fixture (dump,
    2x buckets, re-insert); records, table contents and PRNG positions must not notice."""
    g = H.load_golden(name)
    pref, p, s, b = E.kmer_params(int(g["gs"]))
    e = E.KmerEngine(p, s, b, pref, bmer_log2_buckets=1, smer_log2_buckets=1, flags=E.F_TEST_HOOKS | E.F_TEST_CROWD)      # (log2 buckets are raised to the smallest legal geometry)
    if mode == "async":
        recs, _ = H.run_async(e, g["fastq"])
    else:
        recs = H.run_se(e, g["fastq"])
    want = g["recs"]
    H.assert_recs_equal(recs, want[want["pos"] < 0xFFFFFFF0])
    H.assert_dump_equal(e, g)
    st = e.stats()
    assert st["n_table_growths"] >= 4, st["n_table_growths"]
    e.close()


def test_async_path_matches_oracle_config2_parameters_deep_coverage():
    """BASELINE config-2 k-mer lengths (p17/s20/b24) at 20x coverage through the reference's schedule (100 sync segments), via
    fqsk_submit / fqsk_collect: side streams, early grouping, enqueued syncs, saturating counters, the avg_filling_factor gate --
    records, tables and all four PRNG positions against the oracle."""
    genome = synth.make_genome(300_000, 77)
    codes, _ = synth.make_reads(genome, 40_000, L=150, seed=77, n_frac=0.0005, dup_frac=0.001)
    slab = _fastq_slab(codes)
    pref, p, s, b = E.kmer_params(100)
    e = E.KmerEngine(p, s, b, pref, expected_kmers=1 << 22)
    o = O.OracleEngine(p, s, b, pref)
    got, dup = H.run_async(e, slab)
    want = H.run_se(o, slab)
    n_dup = int((want["pos"] == O.POS_DUP).sum())
    H.assert_recs_equal(got, want[want["pos"] < 0xFFFFFFF0])
    assert int(dup.sum()) == n_dup and n_dup > 0
    for which in (0, 1, 2):
        kg, vg = e.dump(which)
        ko, vo = o.dump(which)
        assert np.array_equal(kg, ko) and np.array_equal(vg, vo), which
    sg, so = e.stats(), o.stats()
    for key in ("siv_no_filled", "siv_no_updates", "n_smers", "n_bmers", "draws_b", "draws_s", "draws_lb", "draws_ls"):
        assert sg[key] == so[key], key
    assert so["repair_missing"] > 0 and so["draws_b"] > 100_000 and so["local_hits"] > 1000
    e.close(); o.close()


def test_large_paired_end_segments_match_oracle():
    """Paired-end segments above 1 MiB of DNA (work items through the filtered delta, radix-sort sync, pair-table growth): records,
    pair decisions and all four tables vs the oracle."""
    genome = synth.make_genome(1_500_000, 41)
    slab = _pe_slab(genome, 8_400, 150, 41, nfrac=0.0005, dup_every=500)
    S_ = __import__("fqsqueezer_b200.schedule", fromlist=["x"])
    off, ln, _, _ = S_.parse_fastq(slab)
    pref, p, s, b = E.kmer_params(16)
    e = E.KmerEngine(p, s, b, pref, mode=E.MODE_PE_ORIGINAL, expected_kmers=1 << 21)
    o = O.OracleEngine(p, s, b, pref, mode=2)
    for eng in (e, o):
        eng.block_start()
    for a, bb in ((0, 8400), (8400, 16800)):
        rg, dg = e.segment(slab, off[a:bb], ln[a:bb])
        ig = e.pair_info((bb - a) // 2)
        ro, do = o.segment(slab, off[a:bb], ln[a:bb], 3)
        io = ro[ro["pos"] == H.POS_PAIR]["c"][:, :3].astype(np.uint32)
        assert np.array_equal(ig, io)
        H.assert_recs_equal(rg, ro[ro["pos"] < 0xFFFFFFF0])
        assert np.array_equal(dg, do)
        e.sync(); o.sync()
    for which in (0, 1, 2, 3):
        kg, vg = e.dump(which)
        ko, vo = o.dump(which)
        assert np.array_equal(kg, ko) and np.array_equal(vg, vo), which
    sg, so = e.stats(), o.stats()
    for key in ("siv_no_filled", "siv_no_updates", "n_smers", "n_bmers", "draws_b", "draws_s", "draws_lb", "draws_ls"):
        assert sg[key] == so[key], key
    assert sg["n_filtered_segments"] == 2 and (io[:, 0] == 1).sum() > 100
    e.close(); o.close()


def test_front_truncated_thread_local_lookup_with_hundreds_of_completions():
    """Low-complexity input: inside ONE segment a fixed 30-mer is preceded by every possible 4-symbol prefix (256 distinct s-mers share
    their last 13 symbols), the b-mers inside it are pushed hundreds of times (thread-local counters above thr: the ordered evaluator
    runs, cinc_lb draws) and later reads START with the 30-mer -- their front-truncated thread-local s-mer lookups match 256 completions,
    twice what the evaluator's in-register list holds (it then switches to its direct-indexed table).  Records, tables and all four PRNG
    positions against the oracle."""
    rng = np.random.default_rng(77)
    L = 80
    fixed = rng.integers(0, 4, 30).astype(np.uint8)
    reads = []
    for q in range(256):      # every 4-symbol prefix in front of the fixed 30-mer
        pre = np.array([(q >> 6) & 3, (q >> 4) & 3, (q >> 2) & 3, q & 3], np.uint8)
        reads.append(np.concatenate((rng.integers(0, 4, 16).astype(np.uint8), pre, fixed, rng.integers(0, 4, L - 50).astype(np.uint8))))
    for _ in range(120):      # reads that start with the fixed 30-mer
        reads.append(np.concatenate((fixed, rng.integers(0, 4, L - 30).astype(np.uint8))))
    for _ in range(200):
        reads.append(rng.integers(0, 4, L).astype(np.uint8))
    codes = np.stack(reads)
    slab = _fastq_slab(codes)
    off, ln, _, _ = __import__("fqsqueezer_b200.schedule", fromlist=["x"]).parse_fastq(slab)
    pref, p, s, b = E.kmer_params(1)
    assert s - p + 1 == 4      # the front-truncated s-mer lookup misses up to 4 symbols: 4^4 completions x 1 next symbol
    e = E.KmerEngine(p, s, b, pref)
    o = O.OracleEngine(p, s, b, pref)
    for eng in (e, o):
        eng.block_start()
    for a, bb in ((0, 376), (376, len(codes))):      # the second segment re-reads everything from the global tables
        rg, dg = e.segment(slab, off[a:bb], ln[a:bb])
        ro, do = o.segment(slab, off[a:bb], ln[a:bb])
        H.assert_recs_equal(rg, ro[ro["pos"] < 0xFFFFFFF0])
        assert np.array_equal(dg, do)
        e.sync(); o.sync()
    for which in (0, 1, 2):
        kg, vg = e.dump(which)
        ko, vo = o.dump(which)
        assert np.array_equal(kg, ko) and np.array_equal(vg, vo), which
    sg, so = e.stats(), o.stats()
    for key in ("siv_no_filled", "siv_no_updates", "n_smers", "n_bmers", "draws_b", "draws_s", "draws_lb", "draws_ls"):
        assert sg[key] == so[key], key
    assert sg["n_hot_segments"] >= 1 and sg["draws_lb"] > 0
    e.close(); o.close()
