"""Pins the CPU oracle against fixtures produced by the REAL reference (tapped fqs-1.1, oracle/make_golden.py):
per-base (counts, level, rough, cor_pos) records and the final k-mer tables must be bit-exact."""
import pytest

from oracle import oracle as O
from tests import helpers as H


@pytest.mark.parametrize("name", ["se_orig_gs1", "se_orig_gs100", "se_orig_repeats_gs1", "se_mixed_unc_gs1", "se_orig_gs3100"])
def test_oracle_matches_reference_tap(name):
    g = H.load_golden(name)
    pref, p, s, b = O.kmer_params(int(g["gs"]))
    e = O.OracleEngine(p, s, b, pref)
    recs = H.run_se(e, g["fastq"])
    H.assert_recs_equal(recs, g["recs"])
    H.assert_dump_equal(e, g)
    st = e.stats()
    if name == "se_orig_gs1":   # the fixture must exercise the hard paths, otherwise it pins nothing
        assert st["draws_b"] > 1000 and st["rough_b"] > 100 and st["repair_existing"] > 10 and st["local_hits"] > 10
        assert (recs["pos"] == O.POS_DUP).sum() > 0
    if name == "se_orig_repeats_gs1":
        assert st["draws_lb"] > 1000    # thread-local counters above thr: cinc_lb in use
    if name == "se_mixed_unc_gs1":      # `mixed` (dna.cpp:470-478) and the bmer_unc revert (dna.cpp:697-705) must both occur
        assert (g["recs"]["level"][g["recs"]["pos"] < 0xFFFFFFF0] == 4).sum() > 1000 and st["mixed"] > 1000
        assert st["unc_reverts"] >= 3
    if name == "se_orig_gs3100":        # the reference's default k-mer lengths: front-truncated b-mer lookups with up to 5 missing symbols
        assert (e.p, e.s, e.b, e.prefix_len) == (18, 21, 27, 13)
        assert st["draws_b"] > 1000 and st["rough_b"] > 100 and st["repair_existing"] > 10 and st["probe_runs_b"] > 1364 * 1000
    e.close()


def test_oracle_matches_reference_tap_sorted_order():
    """-om s: sorted-prefix (flag, dif) values, suffix records from p_len and final tables vs the tapped reference."""
    g = H.load_golden("se_sorted_gs1")
    pref, p, s, b = O.kmer_params(int(g["gs"]))
    e = O.OracleEngine(p, s, b, pref, mode=1)
    recs, flags, difs = H.run_sorted(e, g["fastq"], is_gpu=False)
    want, wflags, wdifs = H.golden_sorted_expect(g)
    H.assert_recs_equal(recs, want)
    import numpy as np
    assert np.array_equal(flags, wflags) and np.array_equal(difs, wdifs)
    assert (wflags == 4).sum() > 0 and (wdifs > 0).sum() > 100
    H.assert_dump_equal(e, g)
    e.close()


@pytest.mark.parametrize("name", ["se_orig_gs1_t2", "se_orig_gs16_t3"])
def test_oracle_group_matches_reference_tap_multithread(name):
    """-t 2 / -t 3: per-worker record streams (each worker has its own PRNG streams and thread-local tables), owner-routed
    exchange rows and the shared tables after all syncs vs the tapped reference run with that many threads."""
    g = H.load_golden(name)
    T = int(g["threads"])
    pref, p, s, b = O.kmer_params(int(g["gs"]))
    grp = O.OracleGroup(p, s, b, pref, T)
    per_worker = H.run_se_workers(grp, g["fastq"], T)
    for w in range(T):
        want = g["recs_t%d" % w]
        H.assert_recs_equal(per_worker[w], want[want["pos"] < 0xFFFFFFF0])
    H.assert_dump_equal(grp, g)
    grp.close()


def test_oracle_matches_reference_tap_paired_end():
    """-p -om o: mate 1 as a single read, minimizer candidates from the pair tables (global + thread-local), mate 2 coded
    forward + reversed from a shared minimizer, 14 pair-table pushes per pair -- records, per-pair decisions and all four tables."""
    import numpy as np
    g = H.load_golden("pe_orig_gs1")
    pref, p, s, b = O.kmer_params(int(g["gs"]))
    e = O.OracleEngine(p, s, b, pref, mode=2)
    recs, info = H.run_pe(e, g["fastq"])
    want, winfo = H.golden_pe_expect(g)
    assert np.array_equal(info, winfo), np.flatnonzero((info != winfo).any(axis=1))[:5]
    H.assert_recs_equal(recs, want)
    assert (winfo[:, 0] == 1).sum() > 500 and ((winfo[:, 0] == 1) & (winfo[:, 1] < 15)).sum() > 500 and (winfo[:, 1] == 15).sum() > 0
    H.assert_dump_equal(e, g, pairs=True)
    e.close()


def test_oracle_matches_reference_tap_paired_end_default_kmer_lengths():
    """-p -om o at the reference's default -gs 3100 (p18/s21/b27, prefix 13: BASELINE config 5's lengths)."""
    import numpy as np
    g = H.load_golden("pe_orig_gs3100")
    pref, p, s, b = O.kmer_params(int(g["gs"]))
    assert (pref, p, s, b) == (13, 18, 21, 27)
    e = O.OracleEngine(p, s, b, pref, mode=2)
    recs, info = H.run_pe(e, g["fastq"])
    want, winfo = H.golden_pe_expect(g)
    assert np.array_equal(info, winfo)
    H.assert_recs_equal(recs, want)
    assert ((winfo[:, 0] == 1) & (winfo[:, 1] < 15)).sum() > 300
    H.assert_dump_equal(e, g, pairs=True)
    e.close()


def test_oracle_matches_reference_tap_paired_end_sorted_order():
    """-p in the reference's DEFAULT order (-om s; params.h:60, BASELINE configs 3 and 5 as written): pairs binned and sorted by mate 1,
    mate 1 through CompressSorted (dna.cpp:1793-1796: sorted prefix flag / dif, suffix from p_len), mate 2 as in original order."""
    import numpy as np
    g = H.load_golden("pe_sorted_gs1")
    pref, p, s, b = O.kmer_params(int(g["gs"]))
    e = O.OracleEngine(p, s, b, pref, mode=3)
    recs, info, flags, difs = H.run_pe_sorted(e, g["fastq"])
    want, winfo, wflags, wdifs = H.golden_pe_sorted_expect(g)
    assert np.array_equal(info, winfo)
    assert np.array_equal(flags, wflags) and np.array_equal(difs, wdifs)
    H.assert_recs_equal(recs, want)
    assert (wdifs > 0).sum() > 100 and ((winfo[:, 0] == 1) & (winfo[:, 1] < 15)).sum() > 500
    H.assert_dump_equal(e, g, pairs=True)
    e.close()


def test_oracle_group_matches_reference_tap_paired_end_two_threads():
    """-p -t 2: per-worker records and pair decisions, pair triples routed to their owners ((fmix64(key) >> 48) % T), all four
    shared tables after all syncs vs the tapped reference run with two threads."""
    import numpy as np
    g = H.load_golden("pe_orig_gs1_t2")
    T = int(g["threads"])
    pref, p, s, b = O.kmer_params(int(g["gs"]))
    grp = O.OracleGroup(p, s, b, pref, T, mode=2)
    per_worker = H.run_pe_workers(grp, g["fastq"], T)
    for w in range(T):
        want = g["recs_t%d" % w]
        winfo = want[want["pos"] == H.POS_PAIR]["c"][:, :3].astype(np.uint32)
        assert np.array_equal(per_worker[w][1], winfo), w
        H.assert_recs_equal(per_worker[w][0], want[want["pos"] < 0xFFFFFFF0])
    H.assert_dump_equal(grp, g, pairs=True)
    grp.close()


@pytest.mark.parametrize("name,paired", [("se_sorted_gs1_t2", False), ("pe_sorted_gs1_t2", True)])
def test_oracle_group_matches_reference_tap_sorted_order_two_threads(name, paired):
    """The reference's default order at -t 2 (-s -om s / -p -om s): sorted bins, every reads_block split between two workers; per worker the
    records, the (flag, dif) of compress_prefix_sorted and the pair decisions, and the shared tables after all syncs."""
    import numpy as np
    g = H.load_golden(name)
    T = int(g["threads"])
    pref, p, s, b = O.kmer_params(int(g["gs"]))
    grp = O.OracleGroup(p, s, b, pref, T, mode=3 if paired else 1)
    per_worker = H.run_sorted_workers(grp, g["fastq"], T, paired=paired)
    n_dif = 0
    for w in range(T):
        got, want = H.split_sorted_markers(per_worker[w], paired), H.split_sorted_markers(g["recs_t%d" % w], paired)
        for x, y in zip(got[1:], want[1:]):
            assert np.array_equal(x, y), w
        H.assert_recs_equal(got[0], want[0])
        n_dif += int((want[2] > 0).sum())
    assert n_dif > 100
    H.assert_dump_equal(grp, g, pairs=paired)
    grp.close()
