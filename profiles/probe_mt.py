"""Measurement helper: throughput of the device mt19937 generator (k_mt_extend) through fqsk_mt_stream, incl. ring allocation and D2H."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fqsqueezer_b200 import engine as E
pref, p, s, b = E.kmer_params(100)
e = E.KmerEngine(p, s, b, pref, expected_kmers=1 << 20)
for n in (1 << 20, 1 << 24, 1 << 24):
    t = time.perf_counter(); x = e.mt_stream(n); dt = time.perf_counter() - t
    print(f"mt_stream({n}): {dt*1e3:.1f} ms -> {n/dt/1e6:.0f} M outputs/s (incl. D2H)")
