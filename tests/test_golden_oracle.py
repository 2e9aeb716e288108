"""Pins the CPU oracle against fixtures produced by the REAL reference (tapped fqs-1.1, oracle/make_golden.py):
per-base (counts, level, rough, cor_pos) records and the final k-mer tables must be bit-exact."""
import pytest

from oracle import oracle as O
from tests import helpers as H


@pytest.mark.parametrize("name", ["se_orig_gs1", "se_orig_gs100"])
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
    e.close()
