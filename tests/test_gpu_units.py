"""GPU parity, table level: the CUDA mirrors of CHT_kmer<T> / TSmallIntVector<2> / CCounterIncrementer / mt19937 against
the CPU oracle on the same seeded inputs.  Bit-exact (integer work).  Everything goes through the C-ABI."""
import numpy as np
import pytest

from fqsqueezer_b200 import engine as E
from oracle import oracle as O
from tests.test_oracle_vs_ref import _normalize, _rand_regs

pytestmark = pytest.mark.gpu


def _engine(p=14, s=17, b=19, pref=9, **kw):
    return E.KmerEngine(p, s, b, pref, **kw)


def test_mt19937_device_stream():
    e = _engine()
    got = e.mt_stream(700000)
    want = O.oracle_units().mt_stream(5481, 700000)
    assert np.array_equal(got, want)
    e.close()


def test_mt19937_device_stream_parallel_chunks():
    """Long extensions run as 32 chunks side by side, each started from a jumped-ahead state (k_mt_jump: x^(j chunk) mod the
    characteristic polynomial, fqsk_mtjump.h): 30 M outputs = three parallel launches + a sequential tail, against the sequential
    generator of the oracle (= std::mt19937 seeded 5481, utils.h:298)."""
    e = _engine()
    n = 30_000_000
    got = e.mt_stream(n)
    want = O.oracle_units().mt_stream(5481, n)
    bad = np.flatnonzero(got != want)
    assert len(bad) == 0, (len(bad), bad[:4])
    assert e.stats()["kernel_launches"] >= 6
    e.close()


@pytest.mark.parametrize("table,k,cbits,log2b", [(E.TABLE_BMER, 19, 6, 0), (E.TABLE_BMER, 24, 6, 0), (E.TABLE_SMER, 20, 12, 0),
                                                 (E.TABLE_BMER, 19, 6, 12), (E.TABLE_BMER, 27, 6, 0)])
def test_insert_dump_find_count(table, k, cbits, log2b):
    """Ordered inserts with the probabilistic counter (> thr draws mt19937 in push order), table growth and the stash."""
    rng = np.random.default_rng(31 * k + log2b)
    p, s, b = (14, 17, k) if table == E.TABLE_BMER else (14, k, k + 4)
    e = _engine(p, s, b, 9, bmer_log2_buckets=log2b, smer_log2_buckets=log2b)
    ou = O.oracle_units()
    thr, mult, top = (7, 2, 63) if cbits == 6 else (2047, 1, 4095)
    to, co = ou.ht_new(k, cbits), ou.cinc_new(thr, mult, top)
    d, rc, cur, _ = _rand_regs(rng, k, 5000)
    universe = _normalize(d, rc, k)
    for rnd in range(4):
        stream = universe[rng.integers(0, 600 if rnd < 3 else len(universe), 60000)]
        stream = np.concatenate([stream, np.repeat(universe[:3], 2000)])
        rng.shuffle(stream)
        e.ht_insert(table, stream)
        ou.ht_insert(to, co, stream)
        kg, vg = e.dump(table)
        ko, vo = ou.ht_dump(to)
        assert np.array_equal(kg, ko)
        assert np.array_equal(vg.astype(np.uint32), vo)
        qd, qrc, qcur, _ = _rand_regs(rng, k, 2000)
        qd = np.concatenate([d[:2000], qd]); qrc = np.concatenate([rc[:2000], qrc]); qcur = np.concatenate([cur[:2000], qcur])
        fo, _ = ou.ht_find(to, co, k, qd, qrc, qcur)
        fg = e.ht_find(table, qd, qrc, qcur)
        assert np.array_equal(fg, fo)
        assert np.array_equal(e.ht_count(table, universe[:3000]), ou.ht_count(to, universe[:3000]))
    assert vo.max() > thr or cbits == 12
    st = e.stats()
    assert st["draws_b" if table == E.TABLE_BMER else "draws_s"] == O.oracle_lib().fqso_cinc_draws(co)
    e.close()


@pytest.mark.parametrize("table,k,cbits,log2b,saturate", [(E.TABLE_BMER, 24, 6, 0, False), (E.TABLE_BMER, 19, 6, 18, False), (E.TABLE_BMER, 24, 6, 0, True),
                                                          (E.TABLE_SMER, 20, 12, 0, False)])
def test_large_row_insert_bucket_grouped(table, k, cbits, log2b, saturate):
    """Rows of millions of k-mers (a steady-state sync): the engine's own radix partition by table bucket (fqsk_sort.cuh) + the
    bucket-grouped ordered insert (k_bucket_flags / k_scan_u8 / k_bucket_apply); log2b = 18 crowds the buckets (1.5 distinct k-mers per bucket on average: stash, growth, runs
    with many distinct k-mers).  saturate: one k-mer repeated far beyond the counter's ceiling -> the row must fall back to the
    sorted-by-k-mer path (48-bit radix sort + k_locate_heads / k_apply_keys with their verifying passes).  Contents, counters and
    the PRNG position against the oracle."""
    rng = np.random.default_rng(77 * k + log2b + saturate)
    p, s, b = (14, 17, k) if table == E.TABLE_BMER else (14, k, k + 4)
    e = _engine(p, s, b, 9, bmer_log2_buckets=log2b, smer_log2_buckets=log2b)
    ou = O.oracle_units()
    thr, mult, top = (7, 2, 63) if cbits == 6 else (2047, 1, 4095)
    to, co = ou.ht_new(k, cbits), ou.cinc_new(thr, mult, top)
    d, rc, cur, _ = _rand_regs(rng, k, 400_000)
    universe = np.unique(_normalize(d, rc, k))
    for rnd in range(3):
        # most k-mers once or twice, a hot core of 20 000 k-mers ~30 times each: counters far above thr, draws in push order
        stream = np.concatenate([universe[rng.integers(0, len(universe), 2_000_000)], universe[rng.integers(0, 20_000, 600_000)]])
        if saturate:
            stream = np.concatenate([stream, np.repeat(universe[5:7], 6000)])
        rng.shuffle(stream)
        e.ht_insert(table, stream)
        ou.ht_insert(to, co, stream)
        kg, vg = e.dump(table)
        ko, vo = ou.ht_dump(to, cap=1 << 22)
        assert np.array_equal(kg, ko)
        bad = np.flatnonzero(vg.astype(np.uint32) != vo)
        assert len(bad) == 0, (rnd, len(bad), kg[bad[:3]], vg[bad[:3]], vo[bad[:3]])
    if saturate:
        assert vo.max() == top
    st = e.stats()
    assert st["draws_b" if table == E.TABLE_BMER else "draws_s"] == O.oracle_lib().fqso_cinc_draws(co)
    e.close()


@pytest.mark.parametrize("k,missing", [(19, 1), (19, 2), (24, 3), (27, 5)])
def test_find_partial_ordered_merge(k, missing):
    """Front-truncated lookups: 4^m completions, merged in order with the PRNG-aware addition; draw offsets across
    queries are resolved by the device-side scan + replay fix point."""
    rng = np.random.default_rng(7 * k + missing)
    e = _engine(14, 17, k, 9)
    ou = O.oracle_units()
    to, co = ou.ht_new(k, 6), ou.cinc_new(7, 2, 63)
    n_suf = 300
    _, _, _, suf = _rand_regs(rng, k, n_suf, cur=k - missing)
    fronts = rng.integers(0, 4, (n_suf, 6, missing), dtype=np.uint64)
    regs = np.array([np.concatenate([f, suf[i]]) for i in range(n_suf) for f in fronts[i]], np.uint64)
    d = np.zeros(len(regs), np.uint64); rc = np.zeros(len(regs), np.uint64)
    for i in range(k):
        d |= regs[:, i] << np.uint64(62 - 2 * i)
        rc |= (np.uint64(3) - regs[:, k - 1 - i]) << np.uint64(62 - 2 * i)
    keys = _normalize(d, rc, k)
    stream = keys[rng.integers(0, len(keys), 30000)]
    e.ht_insert(E.TABLE_BMER, stream); ou.ht_insert(to, co, stream)
    cur = k - missing
    qd = np.zeros(n_suf, np.uint64); qrc = np.zeros(n_suf, np.uint64)
    for i in range(cur):
        qd |= suf[:, i] << np.uint64(62 - 2 * i)
        qrc |= (np.uint64(3) - suf[:, cur - 1 - i]) << np.uint64(62 - 2 * i)
    qcur = np.full(n_suf, cur, np.uint32)
    fo, _ = ou.ht_find(to, co, k, qd, qrc, qcur)
    fg = e.ht_find(E.TABLE_BMER, qd, qrc, qcur)
    assert np.array_equal(fg, fo)
    assert fo.max() > 7
    assert e.stats()["draws_b"] == O.oracle_lib().fqso_cinc_draws(co)
    e.close()


def test_small_int_vector():
    rng = np.random.default_rng(5)
    e = _engine(12, 17, 19, 9)       # p = 12 -> 24 key bits
    ou = O.oracle_units()
    so = ou.siv_new(24)
    idx = rng.integers(0, 1 << 24, 400000).astype(np.uint64)
    idx = np.concatenate([idx, np.repeat(idx[:100], 5)])
    assert e.siv_increment(idx) == ou.siv_increment(so, idx)
    q = np.concatenate([rng.integers(0, 1 << 24, 5000).astype(np.uint64), idx[:5000]])
    assert np.array_equal(e.siv_test(q), ou.siv_test(so, q))
    assert np.array_equal(e.siv_counts(q), ou.siv_counts(so, q))
    for size in (24, 22, 18, 14, 10, 8):
        pre = (q >> np.uint64(24 - size)).astype(np.uint64)
        bits = np.full(len(pre), size, np.uint32)
        assert np.array_equal(e.siv_test_shorter(pre, bits), ou.siv_test_shorter(so, pre, bits))
    kg, vg = e.dump(E.TABLE_SIV)
    assert len(kg) == e.stats()["siv_no_filled"]
    assert np.array_equal(vg.astype(np.uint32), ou.siv_test(so, kg))
    e.close()
