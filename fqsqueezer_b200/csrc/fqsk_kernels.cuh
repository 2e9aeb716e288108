// fqsk_kernels.cuh -- the kernels of the k-mer statistics engine (sm_100a).  Host orchestration lives in fqsk.cu.
//
// Kernel inventory:
//   k_mt_extend        mt19937 block recurrence, one CTA, state in shared memory        (utils.h:257, 298)
//   k_prep             per read: duplicate flag, A/C/G/T totals, number of coded bases  (dna.cpp:1521-1533, 2047-2057)
//   k_replay           per read: the k-mer half of CompressDirect/CompressSorted        (dna.cpp:457-877, 1517-1754)
//   k_compact          per-read push regions -> contiguous to_add rows in push order    (dna.cpp:822-873)
//   k_locate/k_apply/k_commit   InsertKmersToHT for s-/b-mers with ordered PRNG draws   (dna.cpp:2420-2446, ht_kmer.h:420-438)
//   k_siv_increment    InsertKmersToHT for p-mers                                        (dna.cpp:2401-2418)
//   k_find / k_count / k_siv_*   table-level batch mirrors                               (ht_kmer.h:441-510, bit_vec.h:53-123)
#pragma once
#include "fqsk_dev.cuh"
#include "../../include/fqsk.h"

namespace fqsk {

struct U64x4 {
	unsigned long long v[4];
	__host__ __device__ U64x4 operator+(const U64x4 &o) const { U64x4 r; for (int i = 0; i < 4; ++i) r.v[i] = v[i] + o.v[i]; return r; }
};

// ------------------------------------------------------------------------------------------------------------------
// mt19937: extend a stream by n_blocks x 624 tempered outputs.  state[0..623] in HBM between launches.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t far) {
	uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
	return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__global__ void __launch_bounds__(256) k_mt_extend(uint32_t *state, uint32_t *out, uint32_t n_blocks) {
	__shared__ uint32_t st[624];
	const int t = threadIdx.x;
	for (int i = t; i < 624; i += 256) st[i] = state[i];
	__syncthreads();
	for (uint32_t blk = 0; blk < n_blocks; ++blk) {
		uint32_t v;
		// phase A: i in [0, 227) uses old st[i], st[i+1], st[i+397]
		if (t < 227) v = mt_twist(st[t], st[t + 1], st[t + 397]);
		__syncthreads();
		if (t < 227) st[t] = v;
		__syncthreads();
		// phase B: i in [227, 454) uses old st[i], st[i+1] and NEW st[i-227]
		if (t < 227) v = mt_twist(st[t + 227], st[t + 228], st[t]);
		__syncthreads();
		if (t < 227) st[t + 227] = v;
		__syncthreads();
		// phase C: i in [454, 624): i < 623 uses old st[i], st[i+1], new st[i-227]; i = 623 wraps to new st[0]
		if (t < 170) { int i = t + 454; v = mt_twist(st[i], i == 623 ? st[0] : st[i + 1], st[i - 227]); }
		__syncthreads();
		if (t < 170) st[t + 454] = v;
		__syncthreads();
		for (int i = t; i < 624; i += 256) {
			uint32_t y = st[i];
			y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
			out[(uint64_t) blk * 624 + i] = y;
		}
	}
	__syncthreads();
	for (int i = t; i < 624; i += 256) state[i] = st[i];
}

// ------------------------------------------------------------------------------------------------------------------
// engine-wide device view
// ------------------------------------------------------------------------------------------------------------------
struct EngineDev {
	HtDev hb, hs;
	SivDev siv;
	CIncP cib, cis;                 // cinc_b / cinc_s; the local incrementers have the same parameters (dna.cpp:162-165)
	uint32_t p, s, b, prefix_len, sorted;
	uint32_t gate_missing;          // siv_pmer->avg_filling_factor() >= 7.0 (dna.cpp:376), constant inside a segment
	const uint32_t *draws[4];       // streams cinc_b, cinc_s, cinc_lb, cinc_ls at their consumed position
	unsigned long long avail[4];
	int *flags;                     // [0] draw window overflow, [1] unsupported path, [2] changed, [3] scratch
};

struct SegDev {
	const uint8_t *dna; const unsigned long long *off; const uint32_t *len; uint32_t n_reads;
	const uint8_t *prev_read; uint32_t prev_len;      // last read of the previous segment of this block (read_prev, dna.cpp:1550)
	unsigned long long pprev_dir; uint32_t pprev_valid; // pmer_can_prev carried across segments (dna.cpp:167, 655)
	uint8_t *dup; uint32_t *n_coded; U64x4 *letters;  // k_prep outputs
	const unsigned long long *rec_off; const U64x4 *sl_prefix; U64x4 sl_base;
	fqsk_base_rec *recs;
	unsigned long long *push_b, *push_s, *push_p;     // per-read regions: b at 2*off, s at off, p at 2*off + 2*r
	uint32_t *cnt_b, *cnt_s, *cnt_p, *hidden;
	U64x4 *draw_cnt; const U64x4 *draw_guess;
	const uint32_t *base_b, *base_s;                  // push-index base of every read in the delta's numbering
	DeltaDev delta_b, delta_s;
	uint32_t *sorted_flag; unsigned long long *sorted_dif;
};

__device__ __forceinline__ uint32_t dna_code(uint8_t ch) {  // dna.cpp:18-23
	return ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 4u;
}

__global__ void k_prep(SegDev S, uint32_t first_len_bytes) {
	uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= S.n_reads) return;
	const uint8_t *p = S.dna + S.off[r];
	uint32_t n = S.len[r];
	bool same;
	if (r == 0) {
		same = (S.prev_len == n);
		for (uint32_t i = 0; same && i < n; ++i) same = S.prev_read[i] == p[i];
	} else {
		const uint8_t *q = S.dna + S.off[r - 1];
		same = (S.len[r - 1] == n);
		for (uint32_t i = 0; same && i < n; ++i) same = q[i] == p[i];
	}
	S.dup[r] = same;
	U64x4 L; L.v[0] = L.v[1] = L.v[2] = L.v[3] = 0;
	uint32_t coded = 0;
	if (!same) {
		for (uint32_t i = 0; i < n; ++i) { uint32_t c = dna_code(p[i]); if (c < 4) { ++L.v[c]; ++L.v[3 - c]; } }
		coded = n > first_len_bytes ? n - first_len_bytes : 0;
	}
	S.letters[r] = L;
	S.n_coded[r] = coded;
}

// ------------------------------------------------------------------------------------------------------------------
// CHT_kmer::find (ht_kmer.h:504-510): full context -> one sector; front-truncated context -> 4^m completions in odometer
// order merged with the PRNG-aware addition (ht_kmer.h:266-327).  Lookups are issued 4 at a time for memory parallelism.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool any4(const uint32_t c[4]) { return (c[0] | c[1] | c[2] | c[3]) != 0; }

__device__ __forceinline__ KReg partial_trial(const KReg &r, uint32_t k, uint32_t m, uint32_t n) {
	uint32_t rev = 0;
	for (uint32_t j = 0; j < m; ++j) rev |= ((n >> (2 * j)) & 3u) << (2 * (m - 1 - j));
	KReg t;
	t.dir = (r.dir >> (2 * m)) | ((uint64_t) rev << (64 - 2 * m));
	t.rc = r.rc + ((uint64_t) (((1u << (2 * m)) - 1u) - n) << (64 - 2 * k));
	return t;
}

__device__ bool ht_find(const HtDev &t, const CIncP &ci, const KReg &r, uint32_t cur, uint32_t c[4], DrawCursor &dc) {
	c[0] = c[1] = c[2] = c[3] = 0;
	if (cur >= t.k) {
		bool d = kr_is_dir(r, t.k);
		ht_ctx_counts(t, d ? r.dir : r.rc, d, c);
		return any4(c);
	}
	uint32_t m = t.k - cur, trials = 1u << (2 * m);
	for (uint32_t n0 = 0; n0 < trials; n0 += 4) {
		HtKey key[4]; Bucket bk[4]; bool isd[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			KReg tr = partial_trial(r, t.k, m, n0 + u);
			isd[u] = kr_is_dir(tr, t.k);
			key[u] = ht_key(t, isd[u] ? tr.dir : tr.rc);
			bk[u] = ht_load_bucket(t, key[u].bucket);
		}
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			uint32_t loc[4] = {0, 0, 0, 0};
			ht_ctx_counts_from(t, key[u], isd[u], bk[u], loc);
			for (int i = 0; i < 4; ++i) if (loc[i]) c[i] = ci_plus(ci, c[i], loc[i], dc);
		}
	}
	return any4(c);
}

// find_counts_rough_{s,b} (dna.cpp:257-330): every single substitution at positions 0..k-2 (the original included once per
// position); a non-empty neighbour is merged for ALL four symbols.
__device__ bool rough_ht(const HtDev &t, const CIncP &ci, const KReg &base, uint32_t c[4], DrawCursor &dc) {
	c[0] = c[1] = c[2] = c[3] = 0;
	for (uint32_t i = 0; i + 1 < t.k; ++i) {
		HtKey key[4]; Bucket bk[4]; bool isd[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			KReg tr = base;
			kr_set(tr, t.k, j, i);
			isd[j] = kr_is_dir(tr, t.k);
			key[j] = ht_key(t, isd[j] ? tr.dir : tr.rc);
			bk[j] = ht_load_bucket(t, key[j].bucket);
		}
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			uint32_t loc[4] = {0, 0, 0, 0};
			ht_ctx_counts_from(t, key[j], isd[j], bk[j], loc);
			if (any4(loc)) for (int q = 0; q < 4; ++q) c[q] = ci_plus(ci, c[q], loc[q], dc);
		}
	}
	return any4(c);
}

// thread-local table lookups (dna.cpp:485, 495) against the segment delta
__device__ bool local_find(const DeltaDev &d, uint32_t k, const CIncP &ci, const KReg &r, uint32_t cur, uint32_t T, uint32_t c[4], DrawCursor &dc, int *unsupported) {
	c[0] = c[1] = c[2] = c[3] = 0;
	if (d.n == 0) return false;
	if (cur >= k) {
		bool dd = kr_is_dir(r, k);
		delta_ctx_counts(d, k, dd ? r.dir : r.rc, dd, T, c, unsupported);
		return any4(c);
	}
	uint32_t m = k - cur, trials = 1u << (2 * m);
	for (uint32_t n = 0; n < trials; ++n) {
		KReg tr = partial_trial(r, k, m, n);
		bool dd = kr_is_dir(tr, k);
		uint32_t loc[4] = {0, 0, 0, 0};
		delta_ctx_counts(d, k, dd ? tr.dir : tr.rc, dd, T, loc, unsupported);
		for (int i = 0; i < 4; ++i) if (loc[i]) c[i] = ci_plus(ci, c[i], loc[i], dc);
	}
	return any4(c);
}

// ------------------------------------------------------------------------------------------------------------------
// k_replay: one thread replays one read through the reference's per-base state machine.
// ------------------------------------------------------------------------------------------------------------------
struct ReadState {
	KReg pc, sc, bc, pu, su, bu;
	uint32_t n;          // symbols pushed since the registers were reset
	uint32_t cor_pos, n_run;
};

__device__ __forceinline__ void rs_push_all(ReadState &R, const EngineDev &E, uint64_t s) {
	kr_push(R.pc, E.p, R.n, s); kr_push(R.sc, E.s, R.n, s); kr_push(R.bc, E.b, R.n, s);
	kr_push(R.pu, E.p, R.n, s); kr_push(R.su, E.s, R.n, s); kr_push(R.bu, E.b, R.n, s);
	++R.n;
}
__device__ __forceinline__ uint32_t cur_of(uint32_t k, uint32_t n) { return n < k ? n : k; }

__global__ void __launch_bounds__(128) k_replay(EngineDev E, SegDev S) {
	uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= S.n_reads) return;
	U64x4 zero4; zero4.v[0] = zero4.v[1] = zero4.v[2] = zero4.v[3] = 0;
	if (S.dup[r]) { S.cnt_b[r] = S.cnt_s[r] = S.cnt_p[r] = S.hidden[r] = 0; S.draw_cnt[r] = zero4; if (E.sorted) { S.sorted_flag[r] = 0; S.sorted_dif[r] = 0; } return; }
	const uint8_t *p = S.dna + S.off[r];
	const uint32_t size = S.len[r];
	unsigned long long *out_b = S.push_b + 2 * S.off[r];
	unsigned long long *out_s = S.push_s + S.off[r];
	unsigned long long *out_p = S.push_p + 2 * S.off[r] + 2ull * r;
	fqsk_base_rec *rec = S.recs + S.rec_off[r];
	uint32_t nb = 0, ns = 0, np = 0, hidden = 0;
	unsigned long long sl[4];
	for (int i = 0; i < 4; ++i) sl[i] = S.sl_base.v[i] + S.sl_prefix[r].v[i];
	DrawCursor dc[4];
	for (int i = 0; i < 4; ++i) { dc[i].buf = E.draws[i]; dc[i].avail = E.avail[i]; dc[i].base = S.draw_guess[r].v[i]; dc[i].used = 0; dc[i].overflow = E.flags + 0; }
	int *unsupported = E.flags + 1;
	const uint32_t Tb0 = S.base_b ? S.base_b[r] : 0, Ts0 = S.base_s ? S.base_s[r] : 0;
	const CIncP cil_b = E.cib, cil_s = E.cis;

	ReadState R;
	R.pc = R.sc = R.bc = R.pu = R.su = R.bu = KReg{0, 0};
	R.n = 0; R.cor_pos = 0; R.n_run = 0;

	uint32_t start;
	if (!E.sorted) {
		// compress_prefix_direct, register half (dna.cpp:518-545)
		for (uint32_t i = 0; i < E.prefix_len; ++i) {
			uint32_t sym = dna_code(p[i]);
			if (sym == 4) { sym = 0; R.cor_pos = i; }
			rs_push_all(R, E, sym);
		}
		start = E.prefix_len;
	} else {
		// compress_prefix_sorted, siv half (dna.cpp:555-605, 655-660)
		for (uint32_t i = 0; i < E.p; ++i) {
			uint32_t sym = dna_code(p[i]);
			if (sym == 4) { sym = 3; ++R.n_run; } else R.n_run = 0;
			rs_push_all(R, E, sym);
		}
		unsigned long long prev_dir; bool prev_valid;
		if (r == 0) { prev_dir = S.pprev_dir; prev_valid = S.pprev_valid != 0; }
		else {
			const uint8_t *q = S.dna + S.off[r - 1];
			prev_dir = 0;
			for (uint32_t i = 0; i < E.p; ++i) { uint32_t sym = dna_code(q[i]); if (sym == 4) sym = 3; prev_dir |= (unsigned long long) sym << (62 - 2 * i); }
			prev_valid = true;
		}
		uint64_t cur_al = R.pc.dir >> (64 - 2 * E.p);
		uint64_t prev_al = prev_valid ? prev_dir >> (64 - 2 * E.p) : 0;
		uint32_t flag; unsigned long long dif = 0;
		if (R.pc.dir == prev_dir) flag = 4;   // an unset pmer_can_prev has kmer_dir == 0 (kmer.h:233-237), same comparison
		else flag = siv_test(E.siv, cur_al);
		if (flag < 4) for (uint64_t i = prev_al + 1; i < cur_al; ++i) dif += siv_test(E.siv, i) == flag;
		S.sorted_flag[r] = flag; S.sorted_dif[r] = dif;
		out_p[np++] = cur_al;
		out_p[np++] = R.pc.rc >> (64 - 2 * E.p);
		start = E.p;
	}

	const uint32_t b_margin = E.b - E.s - 1, s_margin = E.s - E.p + 1;
	uint32_t c[4] = {0, 0, 0, 0};
	for (uint32_t i = start; i < size; ++i) {
		const uint32_t sym = dna_code(p[i]);
		const uint64_t ks = sym == 4 ? 0 : sym;
		rs_push_all(R, E, 0);   // placeholder (dna.cpp:687-693)
		const uint32_t cb = cur_of(E.b, R.n), cs = cur_of(E.s, R.n), cp = cur_of(E.p, R.n);

		// ---- find_counts (dna.cpp:457-502)
		uint32_t lev = FQSK_LEVEL_NONE;
		c[0] = c[1] = c[2] = c[3] = 0;
		bool done = false;
		if (cb + b_margin >= E.b) {
			if (ht_find(E.hb, E.cib, R.bc, cb, c, dc[0])) {
				int sat = (c[0] == E.hb.top) + (c[1] == E.hb.top) + (c[2] == E.hb.top) + (c[3] == E.hb.top);
				if (sat > 1) {
					uint32_t c2[4];
					ht_find(E.hs, E.cis, R.sc, cs, c2, dc[1]);
					for (int q = 0; q < 4; ++q) c[q] += c2[q];
					lev = FQSK_LEVEL_MIXED;
				} else lev = FQSK_LEVEL_BMER;
				done = true;
			} else if (local_find(S.delta_b, E.b, cil_b, R.bc, cb, Tb0 + nb, c, dc[2], unsupported)) { lev = FQSK_LEVEL_BMER; done = true; }
			else if (R.bc.dir != R.bu.dir && ht_find(E.hb, E.cib, R.bu, cb, c, dc[0])) { lev = FQSK_LEVEL_BMER_UNC; done = true; }
		}
		if (!done) {
			if (cs + s_margin >= E.s) {
				if (ht_find(E.hs, E.cis, R.sc, cs, c, dc[1])) lev = FQSK_LEVEL_SMER;
				else if (local_find(S.delta_s, E.s, cil_s, R.sc, cs, Ts0 + ns, c, dc[3], unsupported)) lev = FQSK_LEVEL_SMER;
			} else {
				// find_counts_p (dna.cpp:210-226)
				if (cp < E.p) {
					for (uint64_t j = 0; j < 4; ++j) {
						KReg t = R.pc;
						kr_set_last(t, cp, j);
						c[j] = (uint32_t) siv_prefix_sum(E.siv, t.rc >> (64 - 2 * cp), 2 * cp);
					}
				} else siv_counts(E.siv, R.pc.dir >> (64 - 2 * E.p), c, false);
				if (any4(c)) lev = FQSK_LEVEL_PMER;
			}
		}
		if (lev == FQSK_LEVEL_BMER_UNC) { R.bc = R.bu; R.sc = R.su; R.pc = R.pu; R.cor_pos = 0; lev = FQSK_LEVEL_BMER; }  // dna.cpp:697-705
		uint32_t rough = 0;
		if (lev == FQSK_LEVEL_NONE) {   // dna.cpp:709-735
			if (cb == E.b) { if (rough_ht(E.hb, E.cib, R.bc, c, dc[0])) { lev = FQSK_LEVEL_PMER; rough = 1; } }
			else if (cs == E.s) { if (rough_ht(E.hs, E.cis, R.sc, c, dc[1])) { lev = FQSK_LEVEL_PMER; rough = 1; } }
			else if (cp == E.p) {
				c[0] = c[1] = c[2] = c[3] = 0;   // find_counts_rough_p (dna.cpp:229-254)
				for (uint32_t q = 0; q + 1 < E.p; ++q)
					for (uint64_t j = 0; j < 4; ++j) { KReg t = R.pc; kr_set(t, E.p, j, q); siv_counts(E.siv, t.dir >> (64 - 2 * E.p), c, true); }
				if (any4(c)) { lev = FQSK_LEVEL_PMER; rough = 1; }
			}
		}
		{
			fqsk_base_rec o;
			o.pos = i; o.counts[0] = c[0]; o.counts[1] = c[1]; o.counts[2] = c[2]; o.counts[3] = c[3];
			o.cor_pos = R.cor_pos; o.level = (uint8_t) lev; o.rough = (uint8_t) rough; o.pad = 0;
			rec[i - start] = o;
		}
		R.n_run = sym == 4 ? R.n_run + 1 : 0;
		kr_set_last(R.pc, cp, ks); kr_set_last(R.sc, cs, ks); kr_set_last(R.bc, cb, ks);
		kr_set_last(R.pu, cp, ks); kr_set_last(R.su, cs, ks); kr_set_last(R.bu, cb, ks);
		if (sym < 4) {   // dna.cpp:818-852
			bool p_insert = true;
			if (cb == E.b) {
				out_b[nb++] = kr_norm(R.bc, E.b);
				if ((lev == FQSK_LEVEL_SMER || lev == FQSK_LEVEL_BMER || lev == FQSK_LEVEL_MIXED) && c[sym] >= 3) p_insert = false;
			}
			if (cs == E.s) out_s[ns++] = kr_norm(R.sc, E.s);
			if (cp == E.p && i - R.cor_pos >= E.p - 1) {
				if (p_insert) { out_p[np++] = R.pc.dir >> (64 - 2 * E.p); out_p[np++] = R.pc.rc >> (64 - 2 * E.p); }
				else hidden += 2;
			}
		}
		if (cb == E.b) {   // dna.cpp:854-875
			bool repaired = false;
			if (lev == FQSK_LEVEL_BMER || lev == FQSK_LEVEL_MIXED) {
				// repair_kmers_existing (dna.cpp:333-370)
				uint32_t best = 0;
				for (uint32_t q = 1; q < 4; ++q) if (c[q] > c[best] || (c[q] == c[best] && sl[q] > sl[best])) best = q;
				bool ok = true;
				if (sym != 4) ok = best != sym && c[sym] == 0 && c[best] > 3;
				if (ok) { kr_set_last(R.pc, cp, best); kr_set_last(R.sc, cs, best); kr_set_last(R.bc, cb, best); R.cor_pos = i; repaired = true; }
			} else if ((lev == FQSK_LEVEL_NONE || lev == FQSK_LEVEL_PMER) && E.gate_missing) {
				// repair_kmers_missing (dna.cpp:374-454)
				int best_c = 4, best_count = 0, best_j = 0;
				for (int j = 1; j < 6; ++j) {
					uint32_t cnts[4];
					uint64_t orig = kr_sym(R.bc, cb - 1 - j);
#pragma unroll
					for (uint64_t q = 0; q < 4; ++q) {
						cnts[q] = 0;
						if (q == orig) continue;
						KReg t = R.bc;
						kr_set(t, cb, q, cb - 1 - j);
						cnts[q] = ht_count(E.hb, kr_norm(t, E.b));
					}
					for (int q = 0; q < 4; ++q) {
						if ((uint64_t) q == orig) continue;
						int cnt = (int) cnts[q];
						if (cnt >= best_count && cnt >= 2) { best_c = q; best_count = cnt; best_j = j; }
					}
				}
				if (best_j) {
					kr_set(R.bc, cb, best_c, cb - 1 - best_j);
					if (best_j < (int) cs) kr_set(R.sc, cs, best_c, cs - 1 - best_j);
					if (best_j < (int) cp) kr_set(R.pc, cp, best_c, cp - 1 - best_j);
					uint32_t np2 = i - (uint32_t) best_j;
					R.cor_pos = R.cor_pos > np2 ? R.cor_pos : np2;
					repaired = true;
				}
			}
			if (repaired) out_b[nb++] = kr_norm(R.bc, E.b);
		}
	}
	S.cnt_b[r] = nb; S.cnt_s[r] = ns; S.cnt_p[r] = np; S.hidden[r] = hidden;
	U64x4 d; for (int i = 0; i < 4; ++i) d.v[i] = dc[i].used;
	S.draw_cnt[r] = d;
}

// per-read regions -> contiguous rows (push order == read order, then base order)
__global__ void k_compact(SegDev S, const uint32_t *off_b, const uint32_t *off_s, const uint32_t *off_p,
                          unsigned long long *row_b, unsigned long long *row_s, unsigned long long *row_p) {
	uint32_t r = blockIdx.x;
	if (r >= S.n_reads) return;
	const unsigned long long *sb = S.push_b + 2 * S.off[r], *ss = S.push_s + S.off[r], *sp = S.push_p + 2 * S.off[r] + 2ull * r;
	for (uint32_t i = threadIdx.x; i < S.cnt_b[r]; i += blockDim.x) row_b[off_b[r] + i] = sb[i];
	for (uint32_t i = threadIdx.x; i < S.cnt_s[r]; i += blockDim.x) row_s[off_s[r] + i] = ss[i];
	for (uint32_t i = threadIdx.x; i < S.cnt_p[r]; i += blockDim.x) row_p[off_p[r] + i] = sp[i];
}

__global__ void k_compare_u64(const unsigned long long *a, const unsigned long long *b, uint64_t n, int *changed) {
	uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && a[i] != b[i]) *changed = 1;
}
__global__ void k_compare_u32(const uint32_t *a, const uint32_t *b, uint64_t n, int *changed) {
	uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && a[i] != b[i]) *changed = 1;
}
__global__ void k_iota(uint32_t *a, uint32_t n) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = i; }

// ------------------------------------------------------------------------------------------------------------------
// sync step for one hash table: InsertKmersToHT (dna.cpp:2420-2446) == for every pushed k-mer, in push order,
// CHT_kmer::insert(x, cinc) (ht_kmer.h:420-438).
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_locate(HtDev t, const unsigned long long *kmers, uint32_t n, uint32_t *slot, uint32_t *val) {
	uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	bool created;
	uint64_t s = ht_locate(t, kmers[j], created);
	slot[j] = (uint32_t) s;
	val[j] = j | (created ? 0x80000000u : 0u);
}
// One thread per group of equal slots (sorted stably, so occurrences are in push order).  flag[j] says whether occurrence j
// consumes a draw; the kernel recomputes that from the counter it sees and reports a change (fix point over the ordered
// draw indices; changes can only come from counters saturating inside the batch).
__global__ void k_apply(HtDev t, CIncP ci, const uint32_t *slot_sorted, const uint32_t *val_sorted, uint32_t n,
                        uint8_t *flag, const uint32_t *draw_off, const uint32_t *draws, unsigned long long avail,
                        uint32_t *final_cnt, int *flags) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t sl = slot_sorted[i];
	if (i > 0 && slot_sorted[i - 1] == sl) return;   // not a group head
	uint32_t e = i;
	uint32_t created = 0;
	while (e < n && slot_sorted[e] == sl) { created |= val_sorted[e] >> 31; ++e; }
	uint32_t c = ht_slot_get(t, sl) - created;       // counter before this sync (a fresh slot was claimed with 1)
	for (uint32_t q = i; q < e; ++q) {
		uint32_t j = val_sorted[q] & 0x7fffffffu;
		if (c >= t.top) { if (flag[j]) { flag[j] = 0; flags[2] = 1; } continue; }   // ht_kmer.h:435: cnt < counter_max
		if (c <= ci.thr) { if (flag[j]) { flag[j] = 0; flags[2] = 1; } ++c; continue; }
		if (!flag[j]) { flag[j] = 1; flags[2] = 1; continue; }                      // needs a draw it was not given yet
		uint32_t di = draw_off[j];
		if (di >= avail) { flags[0] = 1; continue; }
		if (draws[di] % (ci.mult * (c - ci.thr)) == 0) ++c;
	}
	final_cnt[i] = c;
}
__global__ void k_commit(HtDev t, const uint32_t *slot_sorted, uint32_t n, const uint32_t *final_cnt) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t sl = slot_sorted[i];
	if (i > 0 && slot_sorted[i - 1] == sl) return;
	ht_slot_set(t, sl, final_cnt[i]);
}

__global__ void k_siv_increment(SivDev s, const unsigned long long *idx, uint64_t n, unsigned long long *n_new) {
	uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t fresh = 0;
	if (i < n) fresh = siv_increment(s, idx[i]);
	unsigned mask = __ballot_sync(0xffffffffu, fresh);
	if ((threadIdx.x & 31) == 0 && mask) atomicAdd(n_new, (unsigned long long) __popc(mask));
}

// ------------------------------------------------------------------------------------------------------------------
// table-level batch mirrors
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_find(HtDev t, CIncP ci, const unsigned long long *dir, const unsigned long long *rc, const uint32_t *cur, uint32_t n,
                       uint32_t *counts, const uint32_t *draws, unsigned long long avail, const unsigned long long *guess, uint32_t *used, int *flags) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	DrawCursor dc; dc.buf = draws; dc.avail = avail; dc.base = guess ? guess[i] : 0; dc.used = 0; dc.overflow = flags;
	KReg r{dir[i], rc[i]};
	uint32_t c[4];
	ht_find(t, ci, r, cur[i], c, dc);
	for (int q = 0; q < 4; ++q) counts[4 * i + q] = c[q];
	used[i] = dc.used;
}
__global__ void k_count(HtDev t, const unsigned long long *kmers, uint32_t n, uint32_t *out) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = ht_count(t, kmers[i]);
}
__global__ void k_siv_test(SivDev s, const unsigned long long *idx, uint32_t n, uint32_t *out) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = siv_test(s, idx[i]);
}
__global__ void k_siv_counts(SivDev s, const unsigned long long *idx, uint32_t n, uint32_t *out) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { uint32_t c[4]; siv_counts(s, idx[i], c, false); for (int q = 0; q < 4; ++q) out[4 * i + q] = c[q]; }
}
__global__ void k_siv_prefix(SivDev s, const unsigned long long *idx, const uint32_t *bits, uint32_t n, unsigned long long *out) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = siv_prefix_sum(s, idx[i], bits[i]);
}

// ------------------------------------------------------------------------------------------------------------------
// dumps (parity check 1) and table growth
// ------------------------------------------------------------------------------------------------------------------
struct MixInv { uint64_t inv1, inv2; };
__device__ __forceinline__ uint64_t ht_unmix(const HtDev &t, const MixInv &mi, uint64_t x) {
	x ^= x >> t.mix_sh;                 // mix_sh >= W/2 -> the xor-shift is an involution
	x = (x * mi.inv2) & t.maskW;
	x ^= x >> t.mix_sh;
	x = (x * mi.inv1) & t.maskW;
	return x;
}
__global__ void k_dump_ht(HtDev t, MixInv mi, unsigned long long *keys, unsigned long long *vals, unsigned long long cap, unsigned long long *n_out) {
	uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t nm = 8ull << t.B, ns = 1ull << t.stash_log2;
	if (i >= nm + ns) return;
	uint64_t x, cnt;
	if (i < nm) {
		uint32_t it = t.main[i];
		if (!it) return;
		uint64_t rem = it >> (8 + t.cbits);
		uint64_t h = ((i >> 3) << t.rem_bits) | rem;
		uint64_t kernel = ht_unmix(t, mi, h);
		uint64_t ends = (it >> t.cbits) & 0xFF;
		x = ((ends >> 4) << 60) | (kernel << (64 - 2 * t.k + 4)) | ((ends & 0xF) << (64 - 2 * t.k));
		cnt = it & t.top;
	} else {
		unsigned long long it = t.stash[i - nm];
		if (!it) return;
		x = (it >> t.cbits) << (64 - 2 * t.k);
		cnt = it & t.top;
	}
	unsigned long long o = atomicAdd(n_out, 1ull);
	if (o < cap) { keys[o] = x; vals[o] = cnt; }
}
__global__ void k_dump_siv(SivDev s, unsigned long long *keys, unsigned long long *vals, unsigned long long cap, unsigned long long *n_out) {
	uint64_t w = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t nw = (1ull << s.key_bits) >> 4;
	if (w >= nw) return;
	uint32_t d = s.w[w];
	if (!d) return;
	for (uint32_t r = 0; r < 16; ++r) {
		uint32_t f = (d >> (2 * r)) & 3;
		if (f) { unsigned long long o = atomicAdd(n_out, 1ull); if (o < cap) { keys[o] = w * 16 + r; vals[o] = f; } }
	}
}
// re-insert dumped (k-mer, counter) pairs into a fresh, larger table
__global__ void k_reinsert(HtDev t, const unsigned long long *keys, const unsigned long long *vals, uint64_t n) {
	uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	bool created;
	uint64_t s = ht_locate(t, keys[i], created);
	ht_slot_set(t, s, (uint32_t) vals[i]);
}

}  // namespace fqsk
