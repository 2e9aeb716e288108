"""Synthetic Illumina-like reads (SURVEY.md §8d / BASELINE.json `configs`).

genome = default_rng(seed).integers(0, 4, G); read start uniform in [0, G-L); strand uniform;
L = 150; i.i.d. substitution errors p = 0.005 to one of the 3 other bases; no N;
qualities 'I' (10 % 'F', '#' at errors); ids '@SIM.<n> <n>/1'.
PE: insert ~ N(400, 40) clipped to [L, 2000]; mate 2 = reverse complement of the fragment end.

Everything is numpy and deterministic in `seed`; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import numpy as np

_ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_genome(G: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 4, G, dtype=np.uint8)


def _mutate(rng, codes: np.ndarray, p_err: float):
    err = rng.random(codes.shape) < p_err
    shift = rng.integers(1, 4, codes.shape, dtype=np.uint8)
    out = np.where(err, (codes + shift) & 3, codes).astype(np.uint8)
    return out, err


def make_reads(genome: np.ndarray, n_reads: int, L: int = 150, p_err: float = 0.005,
               seed: int = 0, n_frac: float = 0.0, dup_frac: float = 0.0):
    """Returns (codes[n_reads, L] uint8 in 0..3 (4 = N), err_mask[n_reads, L] bool)."""
    rng = np.random.default_rng(seed + 1)
    G = genome.shape[0]
    starts = rng.integers(0, G - L, n_reads)
    strand = rng.integers(0, 2, n_reads).astype(bool)
    idx = starts[:, None] + np.arange(L)[None, :]
    codes = genome[idx]
    rc = (3 - codes)[:, ::-1]
    codes = np.where(strand[:, None], rc, codes).astype(np.uint8)
    codes, err = _mutate(rng, codes, p_err)
    if n_frac > 0:
        isn = rng.random(codes.shape) < n_frac
        codes = np.where(isn, 4, codes).astype(np.uint8)
    if dup_frac > 0 and n_reads > 1:
        dup = rng.random(n_reads) < dup_frac
        dup[0] = False
        for i in np.nonzero(dup)[0]:
            codes[i] = codes[i - 1]
            err[i] = err[i - 1]
    return codes, err


def make_pairs(genome: np.ndarray, n_pairs: int, L: int = 150, p_err: float = 0.005, seed: int = 0, ins_mean: float = 400, ins_sd: float = 40):
    """Returns (codes1, err1, codes2, err2); mate 2 is the reverse complement of the fragment end."""
    rng = np.random.default_rng(seed + 2)
    G = genome.shape[0]
    ins = np.clip(np.rint(rng.normal(ins_mean, ins_sd, n_pairs)), L, 2000).astype(np.int64)
    starts = rng.integers(0, G - min(2000, G // 2), n_pairs)
    strand = rng.integers(0, 2, n_pairs).astype(bool)
    ar = np.arange(L)[None, :]
    left = genome[starts[:, None] + ar]
    right = genome[(starts + ins - L)[:, None] + ar]
    right_rc = (3 - right)[:, ::-1]
    left_rc = (3 - left)[:, ::-1]
    m1 = np.where(strand[:, None], right_rc, left).astype(np.uint8)
    m2 = np.where(strand[:, None], left, right_rc).astype(np.uint8)
    m1, e1 = _mutate(rng, m1, p_err)
    m2, e2 = _mutate(rng, m2, p_err)
    return m1, e1, m2, e2


def codes_to_ascii(codes: np.ndarray) -> np.ndarray:
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    return lut[codes]


# ---- the BASELINE config-2 job as ONE stream of reads (bench.py, oracle/make_bench_golden.py) -------------------------------
JOB_CHUNK = 51_000          # reads are generated in chunks of this many, chunk c with seed 1000 + c; the job is their concatenation


def job_chunk(genome: np.ndarray, c: int, L: int = 150):
    """(codes, err) of chunk c of the job's read stream."""
    return make_reads(genome, JOB_CHUNK, L=L, seed=1000 + c)


def fastq_record_sizes(first_id: int, n: int, L: int = 150) -> np.ndarray:
    """Bytes of the FASTQ records write_fastq produces for reads first_id .. first_id + n - 1 (ids count from 1):
    '@SIM.<n> <n>/1' + DNA + '+' + qualities, four newlines.  The reference cuts its 16 MiB reads_blocks by these sizes
    (reads_block.h:121-169), so they decide which reads share a block -- and with that the whole sync schedule."""
    ids = np.arange(first_id, first_id + n, dtype=np.int64)
    digits = np.floor(np.log10(ids)).astype(np.int64) + 1
    return (9 + 2 * digits + 2 * (L + 1) + 2).astype(np.uint32)


def write_fastq(path: str, codes: np.ndarray, err: np.ndarray, mate: int = 1, seed: int = 0, first_id: int = 1, append: bool = False) -> int:
    """Writes a FASTQ file in the §8d format; returns the number of bytes written."""
    rng = np.random.default_rng(seed + 3)
    n, L = codes.shape
    seq = codes_to_ascii(codes)
    q = np.where(rng.random(codes.shape) < 0.1, ord("F"), ord("I")).astype(np.uint8)
    q = np.where(err, ord("#"), q).astype(np.uint8)
    total = 0
    with open(path, "ab" if append else "wb") as f:
        chunk = 100000
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            parts = []
            for i in range(a, b):
                parts.append(b"@SIM.%d %d/%d\n" % (i + first_id, i + first_id, mate))
                parts.append(seq[i].tobytes())
                parts.append(b"\n+\n")
                parts.append(q[i].tobytes())
                parts.append(b"\n")
            blob = b"".join(parts)
            f.write(blob)
            total += len(blob)
    return total
