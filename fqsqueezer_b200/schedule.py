"""Host-side control that decides WHEN the k-mer tables change.

Mirrors, for the hot path only:
  * CReadsBlock::Read            reads_block.h:121-169   (16 MiB slabs, closed when < 100 KiB / 200 KiB remain)
  * calc_no_synchronizations     application.h:85-92
  * the worker loop's next_synchro arithmetic   application.cpp:617-655 (SE: i == next), 1145-1179 (PE: i >= next, step 2)
  * PartitionForWorkers          reads_block.h:197-214
Pure integer host logic; no GPU, no oracle.
"""
from __future__ import annotations

from typing import Iterator, List, Tuple

import numpy as np

READS_BLOCK_SIZE = 16 << 20          # application.h:34
BLOCK_SIZE_MARGIN = 102400           # reads_block.h:25


def parse_fastq(slab: np.ndarray):
    """Returns (dna_off[u64], dna_len[u32], rec_off[u64], rec_size[u32]) for a FASTQ byte slab (4-line records)."""
    nl = np.flatnonzero(slab == 10)
    n = len(nl) // 4
    nl = nl[: 4 * n].reshape(n, 4)
    rec_off = np.concatenate(([0], nl[:-1, 3] + 1)).astype(np.uint64) if n else np.zeros(0, np.uint64)
    dna_off = (nl[:, 0] + 1).astype(np.uint64)
    dna_len = (nl[:, 1] - nl[:, 0] - 1).astype(np.uint32)
    rec_size = (nl[:, 3] + 1 - rec_off.astype(np.int64)).astype(np.uint32)
    return dna_off, dna_len, rec_off, rec_size


def split_blocks(rec_size: np.ndarray, paired: bool = False) -> List[Tuple[int, int]]:
    """Read-index ranges [first, last) of consecutive reads_blocks (reads_block.h:121-169)."""
    margin = BLOCK_SIZE_MARGIN * (2 if paired else 1)
    step = 2 if paired else 1
    csum = np.concatenate(([0], np.cumsum(rec_size.astype(np.int64))))
    n = len(rec_size)
    blocks = []
    first = 0
    while first < n:
        # smallest last such that BLOCK - (csum[last] - csum[first]) < margin, in steps of `step`
        target = csum[first] + READS_BLOCK_SIZE - margin
        last = int(np.searchsorted(csum, target, side="right"))   # csum[last] > target
        if paired and (last - first) % 2:
            last += 1
        last = min(max(last, first + step), n)
        blocks.append((first, last))
        first = last
    return blocks


def calc_no_synchronizations(generation: int, n_reads: int, n_threads: int) -> int:
    """application.h:85-92."""
    r = 100 - generation if generation < 100 else 0
    r = max(0, min(r, n_reads // n_threads // 2))
    return r - 1 if r else 0


def partition_for_workers(n_reads: int, n_workers: int) -> List[Tuple[int, int]]:
    """reads_block.h:197-214."""
    out, lower = [], 0
    for i in range(n_workers):
        upper = (i + 1) * n_reads // n_workers
        if i < n_workers - 1:
            upper &= ~1
        out.append((lower, upper))
        lower = upper
    return out


def segments(first: int, last: int, n_sync: int, paired: bool = False) -> Iterator[Tuple[int, int]]:
    """Yields the read ranges [a, b) between consecutive table updates for one worker of one block.
    A sync happens after coding read `next` (SE: i == next; PE: i >= next with i stepping by 2) and once more at
    the end of the block (application.cpp:643-662, 1170-1193)."""
    step = 2 if paired else 1
    gen = 0
    nxt = (gen + 1) * (last - first) // (n_sync + 1) + first
    a = first
    i = first
    while i < last:
        hit = (i >= nxt) if paired else (i == nxt)
        if hit:
            yield (a, i + step)
            a = i + step
            gen += 1
            nxt = (gen + 1) * (last - first) // (n_sync + 1) + first
        i += step
    yield (a, last)       # end-of-block sync (possibly over an empty range)


def worker_segments(first: int, last: int, generation: int, n_workers: int, worker: int, paired: bool = False) -> List[Tuple[int, int]]:
    """Sync segments of worker `worker` for the block [first, last) with generation number `generation` at -t n_workers:
    PartitionForWorkers (reads_block.h:197-214) + the per-worker next_synchro arithmetic (application.cpp:617-662).
    Every worker gets the same number of segments (they all meet at every sync)."""
    ns = calc_no_synchronizations(generation, last - first, n_workers)
    a, b = partition_for_workers(last - first, n_workers)[worker]
    return list(segments(first + a, first + b, ns, paired=paired))
