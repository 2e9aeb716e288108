// fqsk_front.cuh -- sorted-mode front end: the order CSortedFASTQFile::sort_reads gives the reads of a bin
// (io.h:499-528).  The reference sorts read descriptors with std::sort and a comparator that walks two reads symbol by symbol:
//   (1) the symbols with N (and anything else that is not A, C, G) read as T, over the shorter length (dna_convert_NT, io.h:552-563);
//   (2) the shorter read first;  (3) the raw bytes (distinguishes N from T).
// std::sort is not stable, so the order of reads that compare equal depends on the sorting algorithm itself.  To stay byte-identical the
// host keeps ITS std::sort call and swaps the comparator for `rank[x] < rank[y]`, where rank is computed here: an integer per read that
// is order-isomorphic to the comparator (equal exactly for reads the comparator calls equivalent).  Every comparison then has the same
// outcome as before, hence the same sequence of swaps and the same final order, ties included -- but it is one integer compare.
//
//   k_front_key    thread per read: the first 32 symbols (N -> T) as a big-endian 64-bit key (shorter reads padded with A)
//   radix sort     (key, read) pairs by all 64 bits: reads that differ within their first 32 symbols are in final order
//   k_front_group  the head of every run of equal keys writes the run's bounds to its members
//   k_front_rank   thread per read: rank = start of its run + the number of members the comparator puts strictly before it
//                  (runs are short: reads that share 32 symbols start at the same genome position or are duplicates)
#pragma once
#include "fqsk_sort.cuh"

namespace fqsk {

__device__ __forceinline__ uint32_t front_nt(uint8_t ch) { return ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : 3u; }      // dna_convert_NT (io.h:552-563)

// the reference's comparator (io.h:501-527): does x sort strictly before y?
__device__ bool front_less(const uint8_t *x, uint32_t xl, const uint8_t *y, uint32_t yl) {
	const uint32_t ml = xl < yl ? xl : yl;
	for (uint32_t i = 0; i < ml; ++i) {
		const uint32_t a = front_nt(x[i]), b = front_nt(y[i]);
		if (a != b) return a < b;
	}
	if (xl != yl) return xl < yl;
	for (uint32_t i = 0; i < ml; ++i) if (x[i] != y[i]) return x[i] < y[i];
	return false;
}

__global__ void k_front_key(const uint8_t *slab, const fqsk_read_desc *reads, uint32_t n, unsigned long long *key) { pdl_enter();
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const uint8_t *p = slab + reads[r].dna_off;
	const uint32_t len = reads[r].dna_len < 32 ? reads[r].dna_len : 32;
	unsigned long long k = 0;
	for (uint32_t i = 0; i < len; ++i) k |= (unsigned long long) front_nt(p[i]) << (62 - 2 * i);
	key[r] = k;
}

// skey: sorted keys.  The head of a run of equal keys walks the run and leaves [start, end) at every member.
__global__ void k_front_group(const unsigned long long *skey, uint32_t n, uint32_t *gstart, uint32_t *gend, uint32_t *max_run) { pdl_enter();
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const unsigned long long k = skey[i];
	if (i > 0 && skey[i - 1] == k) return;
	uint32_t e = i + 1;
	while (e < n && skey[e] == k) ++e;
	for (uint32_t j = i; j < e; ++j) { gstart[j] = i; gend[j] = e; }
	if (e - i > 1) atomicMax(max_run, e - i);
}

__global__ void k_front_rank(const uint8_t *slab, const fqsk_read_desc *reads, const uint32_t *sidx, const uint32_t *gstart, const uint32_t *gend, uint32_t n, uint32_t *rank) { pdl_enter();
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t me = sidx[i], s = gstart[i], e = gend[i];
	uint32_t r = s;
	if (e - s > 1) {
		const uint8_t *x = slab + reads[me].dna_off;
		const uint32_t xl = reads[me].dna_len;
		for (uint32_t j = s; j < e; ++j) {
			if (j == i) continue;
			const uint32_t o = sidx[j];
			if (front_less(slab + reads[o].dna_off, reads[o].dna_len, x, xl)) ++r;
		}
	}
	rank[me] = r;
}

}  // namespace fqsk
