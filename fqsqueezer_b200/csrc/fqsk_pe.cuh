// fqsk_pe.cuh -- paired-end front end: CHT_pair_kmers on the device, window minimizers, the
// candidate ranking of find_minim_cand / merge_minim_results and the choice of the shared minimizer of mate 2
// (dna.cpp:880-1136, 1757-1880; ht_kmer.h:559-663, ht_kmer.cpp:17-230).
//
// Why this is parallel although the reference interleaves it with the coding of the reads: the 14 (key, value, weight)
// triples a pair pushes (append_pe_mers3, dna.cpp:1058-1136) are a pure function of the two mates, the global pair table is
// frozen between syncs, and the thread-local pair table at pair r holds exactly the triples of pairs < r of the segment.  So
// every pair's decision (found, minimizer id, position) can be taken up front: the triples of the whole segment are sorted by
// (key, value) -- stable, so pair order survives inside a group -- and a thread-local `find` is a binary search plus a walk
// over the entries with an earlier pair index.  The same sorted list is what the sync inserts into the global table.
// The decisions turn each pair into three work ITEMS for the segment pipeline: mate 1, mate 2 (whole, or the part right of the
// minimizer), and the reverse complement of the part left of it (CompressDirectWithMinim, dna.cpp:1559-1638).
#pragma once
#include "fqsk_pipeline.cuh"

namespace fqsk {

struct PairDev {                     // CHT_pair_kmers by contents: open addressing on fmix64(key), item = (key, value | count << 2b)
	unsigned long long *keys, *vcs;  // empty: key == PAIR_EMPTY, vc == ~0 (its value part is value_mask, which is never stored)
	unsigned long long mask;         // slots - 1
	unsigned long long vm, top;      // value_mask = 4^b - 1 (ht_kmer.cpp:25), counter ceiling = ~0 >> 2b
	uint32_t b;
	// sharded operation: the pair table is split by the reference's own owner key (fmix64(key) >> 48) % world (ht_kmer.h:599-602,
	// dna.cpp:1076-1081); keys / vcs are THIS rank's shard, peer_* every rank's (NVLink peer mappings, [rank] = own) for lookups.
	// All shards have the same size.
	const unsigned long long *peer_keys[8], *peer_vcs[8];
	uint32_t world;
};
FQSK_HD uint32_t pair_owner(uint64_t key, uint32_t world) { return world > 1 ? (uint32_t) ((fmix64(key) >> 48) % world) : 0u; }
static const unsigned long long PAIR_EMPTY = ~0ull;


// b-mer starting at p[j] as an aligned (right-justified) value; false when the window holds an N (the reference restarts its
// rolling register at every N: dna.cpp:985-986, 1013-1015, 1042-1044)
FQSK_DEV bool bmer_fwd(const uint8_t *p, uint32_t j, uint32_t b, unsigned long long &v) {
	v = 0;
	for (uint32_t t = 0; t < b; ++t) { uint32_t c = dna_code(p[j + t]); if (c == 4) return false; v = (v << 2) | c; }
	return true;
}
FQSK_DEV unsigned long long bmer_rev_nocheck(const uint8_t *p, uint32_t j, uint32_t b) {   // symbols entering from the END of the window (find_maximizer, dna.cpp:1031-1055)
	unsigned long long v = 0;
	for (uint32_t t = 0; t < b; ++t) v = (v << 2) | (dna_code(p[j + b - 1 - t]) & 3u);
	return v;
}
FQSK_DEV bool valid_minimizer(unsigned long long x, uint32_t b) { unsigned long long f = x >> (2 * b - 6); return f != 0 && f != 1; }        // not AAA*, AAC* (dna.cpp:880-891)
FQSK_DEV bool valid_maximizer(unsigned long long x, uint32_t b) { unsigned long long f = x >> (2 * b - 6); return f != 0x3e && f != 0x3f; }  // not TTG*, TTT* (dna.cpp:894-905)
FQSK_DEV unsigned long long warp_min64(unsigned long long x) {
	for (int o = 16; o; o >>= 1) { unsigned long long y = __shfl_xor_sync(0xffffffffu, x, o); x = y < x ? y : x; }
	return x;
}
FQSK_DEV unsigned long long warp_max64(unsigned long long x) {
	for (int o = 16; o; o >>= 1) { unsigned long long y = __shfl_xor_sync(0xffffffffu, x, o); x = y > x ? y : x; }
	return x;
}

// append_pe_mers3 (dna.cpp:1091-1115): which of {m11, m12, m13, m21, m22, m23, x1, x2} is key / value of push u, and its weight
__constant__ uint8_t PE_TRI_K[14] = {0, 0, 0, 1, 1, 2, 2, 3, 3, 3, 4, 4, 5, 5};
__constant__ uint8_t PE_TRI_V[14] = {3, 5, 6, 3, 5, 3, 5, 0, 2, 7, 0, 2, 0, 2};
__constant__ uint8_t PE_TRI_C[14] = {2, 4, 1, 3, 3, 4, 2, 2, 4, 1, 3, 4, 4, 2};

// ------------------------------------------------------------------------------------------------------------------
// k_pe_minim: one warp per pair.  Window minimizers of both mates: thirds (the pushes, dna.cpp:1064-1090), quarters of mate 1
// (the look-up keys, dna.cpp:1762-1769), the maximizer of the second half of mate 1 and the minimizer of the second half of
// mate 2 (dna.cpp:1085-1090).  Writes the 14 triples of the pair (invalid ones -- a missing minimizer, ht_kmer.cpp:123-124 --
// get a key above every real key) and the 4 look-up keys.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_pe_minim(const uint8_t *dna, const unsigned long long *off, const uint32_t *len, uint32_t n_pairs, uint32_t b,
                                                  unsigned long long *tri_key, unsigned long long *tri_val, unsigned long long *qkeys) { pdl_enter();
	const uint32_t pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (pair >= n_pairs) return;
	__shared__ unsigned long long vals_all[4][8];
	unsigned long long *vals = vals_all[threadIdx.x >> 5];
	const unsigned long long vm = (1ull << (2 * b)) - 1;
	const uint8_t *p1 = dna + off[2 * pair], *p2 = dna + off[2 * pair + 1];
	const uint32_t L1 = len[2 * pair], L2 = len[2 * pair + 1];
	unsigned long long t0 = vm, t1 = vm, t2 = vm, q0 = vm, q1 = vm, q2 = vm, q3 = vm, x1 = 0, u0 = vm, u1 = vm, u2 = vm, x2 = vm;
	if (L1 >= b) {
		const uint32_t mss = L1 - b + 1, sp1 = mss / 3, sp2 = 2 * mss / 3, s1 = mss / 4, s2 = 2 * mss / 4, s3 = 3 * mss / 4, xs = (L1 + b) / 2 - b + 1;
		for (uint32_t j = lane; j < mss; j += 32) {
			unsigned long long v;
			if (!bmer_fwd(p1, j, b, v)) continue;
			if (valid_minimizer(v, b)) {
				if (j < sp1) t0 = v < t0 ? v : t0; else if (j < sp2) t1 = v < t1 ? v : t1; else t2 = v < t2 ? v : t2;
				if (j < s1) q0 = v < q0 ? v : q0; else if (j < s2) q1 = v < q1 ? v : q1; else if (j < s3) q2 = v < q2 ? v : q2; else q3 = v < q3 ? v : q3;
			}
			if (j >= xs) { unsigned long long vr = bmer_rev_nocheck(p1, j, b); if (valid_maximizer(vr, b) && vr > x1) x1 = vr; }
		}
	}
	if (L2 >= b) {
		const uint32_t mss = L2 - b + 1, sp1 = mss / 3, sp2 = 2 * mss / 3, xs = (L2 + b) / 2 - b + 1;
		for (uint32_t j = lane; j < mss; j += 32) {
			unsigned long long v;
			if (!bmer_fwd(p2, j, b, v) || !valid_minimizer(v, b)) continue;
			if (j < sp1) u0 = v < u0 ? v : u0; else if (j < sp2) u1 = v < u1 ? v : u1; else u2 = v < u2 ? v : u2;
			if (j >= xs) x2 = v < x2 ? v : x2;      // sic: a MINimizer for the second mate (dna.cpp:1089)
		}
	}
	t0 = warp_min64(t0); t1 = warp_min64(t1); t2 = warp_min64(t2); q0 = warp_min64(q0); q1 = warp_min64(q1); q2 = warp_min64(q2); q3 = warp_min64(q3);
	u0 = warp_min64(u0); u1 = warp_min64(u1); u2 = warp_min64(u2); x2 = warp_min64(x2); x1 = warp_max64(x1);
	if (lane == 0) {
		vals[0] = t0; vals[1] = t1; vals[2] = t2; vals[3] = u0; vals[4] = u1; vals[5] = u2; vals[6] = (~x1) & vm; vals[7] = (~x2) & vm;
		qkeys[4ull * pair + 0] = q0; qkeys[4ull * pair + 1] = q1; qkeys[4ull * pair + 2] = q2; qkeys[4ull * pair + 3] = q3;
	}
	__syncwarp();
	if (lane < 14) {
		unsigned long long k = vals[PE_TRI_K[lane]], v = vals[PE_TRI_V[lane]];
		if (k == vm || v == vm) { k = vm + 1; v = 0; }
		tri_key[14ull * pair + lane] = k; tri_val[14ull * pair + lane] = v;
	}
}

__global__ void k_pe_gather(const unsigned long long *src, const uint32_t *idx, unsigned long long *dst, uint32_t n) { pdl_enter();
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[i] = src[idx[i]];
}

FQSK_DEV uint32_t pe_lower_bound(const unsigned long long *a, uint32_t n, unsigned long long x) {
	uint32_t lo = 0, hi = n;
	while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (a[m] < x) lo = m + 1; else hi = m; }
	return lo;
}

struct PeSeg {                   // the segment's triples, sorted by (key, value); ties in push order
	const unsigned long long *skey, *sval; const uint32_t *sidx; uint32_t n;   // sidx = 14 * pair + push number
};

// candidates of one look-up key from one source (0: global table, 1: thread-local entries of earlier pairs); out == nullptr counts
FQSK_DEV uint32_t pe_collect(const PairDev &G, const PeSeg &L, uint32_t src, unsigned long long key, uint32_t pair, unsigned long long *out) {
	uint32_t n = 0;
	if (key >= G.vm) return 0;                    // "no minimizer" is never stored (ht_kmer.cpp:123-124)
	if (src == 0) {
		const uint32_t o = pair_owner(key, G.world);
		const unsigned long long *K = G.peer_keys[o], *V = G.peer_vcs[o];
		for (unsigned long long s = fmix64(key) & G.mask;; s = (s + 1) & G.mask) {
			const unsigned long long k = K[s];
			if (k == PAIR_EMPTY) break;
			if (k == key) { if (out) out[n] = V[s]; ++n; }
		}
		return n;
	}
	uint32_t i = pe_lower_bound(L.skey, L.n, key);
	while (i < L.n && L.skey[i] == key) {
		const unsigned long long v = L.sval[i];
		unsigned long long c = 0;
		uint32_t j = i;
		for (; j < L.n && L.skey[j] == key && L.sval[j] == v; ++j) if (L.sidx[j] / 14 < pair) c += PE_TRI_C[L.sidx[j] % 14];
		if (c) { if (out) out[n] = v | ((c < G.top ? c : G.top) << (2 * G.b)); ++n; }   // saturating counter of the local table (ht_kmer.cpp:144-160)
		i = j;
	}
	return n;
}

struct PeItems {                 // work items of the segment pipeline, 3 per pair
	unsigned long long *src; uint32_t *len, *bytes, *first, *bias, *dup_prev; uint8_t *flags;
};

// by_count of merge_minim_results (dna.cpp:925-932): larger counter first, then smaller value
FQSK_DEV bool pe_before(unsigned long long x, unsigned long long y, uint32_t sh, unsigned long long vm) {
	unsigned long long cx = x >> sh, cy = y >> sh;
	if (cx != cy) return cx > cy;
	return (x & vm) < (y & vm);
}

// ------------------------------------------------------------------------------------------------------------------
// k_pe_decide: one warp per pair.  find_minim_cand (dna.cpp:1757-1787): 4 finds in the global and 4 in the thread-local pair
// table; merge_minim_results (906-971); then CompressPE's search of the ranked candidates among the valid b-mers of mate 2
// (1805-1822, generate_read_bmers 974-999): first candidate (of the first 15) present in mate 2, at its first position.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_pe_decide(PairDev G, PeSeg L, const unsigned long long *qkeys, const uint8_t *dna, const unsigned long long *off, const uint32_t *len,
                                                   uint32_t n_pairs, uint32_t prefix_len, uint32_t first1, unsigned long long *pool, uint32_t *pool_used, uint32_t pool_cap, int *overflow,
                                                   uint32_t *info, PeItems I) { pdl_enter();
	const uint32_t pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (pair >= n_pairs) return;
	__shared__ unsigned long long top_all[4][48];
	unsigned long long *top = top_all[threadIdx.x >> 5];
	const uint32_t b = G.b, sh = 2 * b;
	const unsigned long long vm = G.vm;
	const unsigned long long key = lane < 8 ? qkeys[4ull * pair + (lane & 3)] : vm;
	uint32_t mine = lane < 8 ? pe_collect(G, L, lane >> 2, key, pair, nullptr) : 0;
	uint32_t pre = mine;
	for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= (uint32_t) o) pre += y; }
	const uint32_t total = __shfl_sync(0xffffffffu, pre, 31);
	pre -= mine;
	uint32_t base = 0;
	if (lane == 0 && total) base = atomicAdd(pool_used, total);
	base = __shfl_sync(0xffffffffu, base, 0);
	if (total && (unsigned long long) base + total > pool_cap) { if (lane == 0) *overflow = 1; return; }
	if (mine) pe_collect(G, L, lane >> 2, key, pair, pool + base + pre);
	__syncwarp();
	uint32_t ntop = 0;
	if (lane == 0 && total) {
		unsigned long long *c = pool + base;
		uint32_t n = total;
		if (n == 1) { top[0] = c[0]; ntop = 1; }
		else {
			if (n > 48) {          // partial_sort to the best 3 * no_examined_pe_minim (dna.cpp:922-934, dna.h:84)
				for (uint32_t i = 0; i < 48; ++i) {
					uint32_t best = i;
					for (uint32_t j = i + 1; j < n; ++j) if (pe_before(c[j], c[best], sh, vm)) best = j;
					unsigned long long t = c[i]; c[i] = c[best]; c[best] = t;
				}
				n = 48;
			}
			for (uint32_t i = 0; i < n; ++i) {       // sort by value (936-938)
				unsigned long long x = c[i]; uint32_t j = i;
				while (j > 0 && (top[j - 1] & vm) > (x & vm)) { top[j] = top[j - 1]; --j; }
				top[j] = x;
			}
			uint32_t m = 0;                            // merge equal values, saturating (944-957)
			for (uint32_t i = 1; i < n; ++i) {
				if ((top[m] & vm) != (top[i] & vm)) top[++m] = top[i];
				else { unsigned long long cx = top[m] >> sh, cy = top[i] >> sh; if (cx + cy > G.top) cy = G.top - cx; top[m] += cy << sh; }
			}
			++m;
			for (uint32_t i = 1; i < m; ++i) {       // order by counter (959-970); only the first 15 are ever told apart
				unsigned long long x = top[i]; uint32_t j = i;
				while (j > 0 && pe_before(x, top[j - 1], sh, vm)) { top[j] = top[j - 1]; --j; }
				top[j] = x;
			}
			ntop = m;
		}
	}
	ntop = __shfl_sync(0xffffffffu, ntop, 0);
	__syncwarp();
	const uint8_t *p2 = dna + off[2 * pair + 1];
	const uint32_t L1 = len[2 * pair], L2 = len[2 * pair + 1];
	const uint32_t lim = ntop < 15 ? ntop : 15;
	uint32_t best = lim, pos = 0;
	if (L2 >= b) {
		const uint32_t mss = L2 - b + 1;
		for (uint32_t j0 = 0; j0 < mss && best > 0; j0 += 32) {
			const uint32_t j = j0 + lane;
			unsigned long long v = 0;
			const bool ok = j < mss && bmer_fwd(p2, j, b, v) && valid_minimizer(v, b);
			for (uint32_t i = 0; i < best; ++i) {
				const unsigned m = __ballot_sync(0xffffffffu, ok && v == (top[i] & vm));
				if (m) { best = i; pos = j0 + (uint32_t) __ffs(m) - 1; break; }
			}
		}
	}
	if (lane != 0) return;
	const uint32_t found = total ? 1 : 0;
	const uint32_t id = found ? (best < lim ? best : 15) : 0;
	const bool split = found && id < 15;
	info[3ull * pair] = found; info[3ull * pair + 1] = id; info[3ull * pair + 2] = split ? pos : 0;
	const uint32_t a = 3 * pair;
	// mate 1: CompressDirect, or -- in sorted order -- CompressSorted, which codes from p_len (first1 = p_len then; dna.cpp:1793-1796)
	I.src[a] = off[2 * pair]; I.len[a] = L1; I.bytes[a] = L1 > first1 ? L1 : first1; I.first[a] = first1; I.bias[a] = 0; I.flags[a] = first1 != prefix_len ? IF_SORTED : 0;
	I.dup_prev[a] = pair ? a - 3 : 0xFFFFFFFFu;
	I.dup_prev[a + 1] = I.dup_prev[a + 2] = 0xFFFFFFFFu;
	if (split) {
		I.src[a + 1] = off[2 * pair + 1] + pos; I.len[a + 1] = L2 - pos; I.bytes[a + 1] = L2 - pos; I.first[a + 1] = b; I.bias[a + 1] = pos;
		I.flags[a + 1] = IF_NO_DUPCHECK | IF_SEEDED | IF_NO_LETTERS;
		I.src[a + 2] = off[2 * pair + 1]; I.len[a + 2] = pos + b; I.bytes[a + 2] = pos + b; I.first[a + 2] = b; I.bias[a + 2] = 0;
		I.flags[a + 2] = IF_NO_DUPCHECK | IF_SEEDED | IF_LETTERS_PREV | IF_REVCOMP;
	} else {
		I.src[a + 1] = off[2 * pair + 1]; I.len[a + 1] = L2; I.bytes[a + 1] = L2 > prefix_len ? L2 : prefix_len; I.first[a + 1] = prefix_len; I.bias[a + 1] = 0;
		I.flags[a + 1] = IF_NO_DUPCHECK;
		I.src[a + 2] = 0; I.len[a + 2] = 0; I.bytes[a + 2] = 0; I.first[a + 2] = 0; I.bias[a + 2] = 0; I.flags[a + 2] = IF_SKIP | IF_NO_DUPCHECK;
	}
}

// item texts, one warp per item: a copy of the mate (or of its right part), or the reverse complement of its left part
// (dna.cpp:1598-1602; reverse_complement_alhpa, utils.h:105-116)
__global__ void __launch_bounds__(128) k_pe_fill(const uint8_t *dna, PeItems I, const uint32_t *off32, uint32_t n_items, uint8_t *out, unsigned long long *off64) { pdl_enter();
	const uint32_t it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (it > n_items) return;
	if (lane == 0) off64[it] = off32[it];
	if (it == n_items) return;
	const uint8_t *s = dna + I.src[it];
	uint8_t *d = out + off32[it];
	const uint32_t n = I.bytes[it];
	if (I.flags[it] & IF_REVCOMP) {
		for (uint32_t t = lane; t < n; t += 32) { uint8_t ch = s[n - 1 - t]; d[t] = ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : 'N'; }
	} else for (uint32_t t = lane; t < n; t += 32) d[t] = s[t];
}

// ------------------------------------------------------------------------------------------------------------------
// sync: CHT_pair_kmers::insert for every pushed triple (dna.cpp:2448-2468, ht_kmer.cpp:121-188).  The saturating addition is
// commutative, so the triples of one (key, value) are summed first (they are adjacent in the sorted list) and every distinct
// pair is inserted once: no two threads ever work on the same item.  A slot claimed by another new pair shows an unwritten
// value part (value_mask) until its owner stores it, which no real value equals -- so it is skipped, as it must be.
// ------------------------------------------------------------------------------------------------------------------
FQSK_DEV void pair_insert_one(const PairDev &G, unsigned long long key, unsigned long long val, unsigned long long cnt, unsigned long long *n_items) {
	const uint32_t sh = 2 * G.b;
	for (unsigned long long s = fmix64(key) & G.mask;; s = (s + 1) & G.mask) {
		unsigned long long k = G.keys[s];
		if (k == PAIR_EMPTY) {
			k = atomicCAS(G.keys + s, PAIR_EMPTY, key);
			if (k == PAIR_EMPTY) {
				atomicExch(G.vcs + s, val | ((cnt < G.top ? cnt : G.top) << sh));
				atomicAdd(n_items, 1ull);
				return;
			}
		}
		if (k != key) continue;
		const unsigned long long vc = *(volatile unsigned long long *) (G.vcs + s);
		if ((vc & G.vm) != val) continue;
		const unsigned long long c = vc >> sh;
		G.vcs[s] = vc + (((c + cnt < G.top) ? cnt : G.top - c) << sh);
		return;
	}
}
__global__ void k_pair_insert(PairDev G, PeSeg L, unsigned long long *n_items) { pdl_enter();
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= L.n) return;
	const unsigned long long key = L.skey[i], val = L.sval[i];
	if (key > G.vm) return;
	if (i > 0 && L.skey[i - 1] == key && L.sval[i - 1] == val) return;
	unsigned long long cnt = 0;
	for (uint32_t j = i; j < L.n && L.skey[j] == key && L.sval[j] == val; ++j) cnt += PE_TRI_C[L.sidx[j] % 14];
	pair_insert_one(G, key, val, cnt, n_items);
}
// sharded sync, source side: the distinct (key, value) pairs of the segment with their summed weights, compacted (order is
// irrelevant: the insertion is commutative) -- the rows [rank][*] of the reference's pe_mers_to_add matrix before routing
__global__ void k_pair_heads(PeSeg L, unsigned long long vm, unsigned long long *okey, unsigned long long *oval, unsigned long long *ocnt, uint32_t *n_out) { pdl_enter();
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= L.n) return;
	const unsigned long long key = L.skey[i], val = L.sval[i];
	if (key > vm) return;
	if (i > 0 && L.skey[i - 1] == key && L.sval[i - 1] == val) return;
	unsigned long long cnt = 0;
	for (uint32_t j = i; j < L.n && L.skey[j] == key && L.sval[j] == val; ++j) cnt += PE_TRI_C[L.sidx[j] % 14];
	const uint32_t o = atomicAdd(n_out, 1u);
	okey[o] = key; oval[o] = val; ocnt[o] = cnt;
}
__global__ void k_pair_owner_keys(const unsigned long long *keys, uint32_t n, uint32_t world, uint8_t *okeys, uint32_t *hist) { pdl_enter();
	__shared__ uint32_t sh[8];
	if (threadIdx.x < 8) sh[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { const uint32_t o = pair_owner(keys[i], world); okeys[i] = (uint8_t) o; atomicAdd(sh + o, 1u); }
	__syncthreads();
	if (threadIdx.x < 8 && sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, sh[threadIdx.x]);
}
// sharded sync, owner side: one source's row (distinct pairs) into this rank's shard
__global__ void k_pair_insert_list(PairDev G, const unsigned long long *keys, const unsigned long long *vals, const unsigned long long *cnts, uint32_t n, unsigned long long *n_items) { pdl_enter();
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) pair_insert_one(G, keys[i], vals[i], cnts[i], n_items);
}

__global__ void k_pair_rehash(PairDev old_t, PairDev new_t) { pdl_enter();
	const unsigned long long slots = old_t.mask + 1;
	for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (unsigned long long) gridDim.x * blockDim.x) {
		const unsigned long long key = old_t.keys[i];
		if (key == PAIR_EMPTY) continue;
		for (unsigned long long s = fmix64(key) & new_t.mask;; s = (s + 1) & new_t.mask) {
			if (new_t.keys[s] != PAIR_EMPTY) continue;
			if (atomicCAS(new_t.keys + s, PAIR_EMPTY, key) == PAIR_EMPTY) { new_t.vcs[s] = old_t.vcs[i]; break; }
		}
	}
}

}  // namespace fqsk
