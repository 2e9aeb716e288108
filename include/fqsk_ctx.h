/* fqsk_ctx.h -- device-side context ids of the DNA stream: one 16-byte record per coded base.
 *
 * What the reference computes on the host for every base it codes with counts (dna.cpp:737-774):
 *   cor_zone (dna.cpp:739-744), CCodeContext::determine_ctx_codes (code_ctx.cpp:257-324: 7 nested 64-bit context ids built from the
 *   4 sorted counts quantised at four resolutions, counts_level, position, correction zone, recent-rank history and let_max),
 *   rank (dna.cpp:177-193), update_ctx_r_sym (dna.cpp:664-671).
 * All of it is a pure function of values the engine has on the device (counts, level, rough flag, cor_pos, the read, the running
 * A/C/G/T totals), so with fqsk_submit_ctx the engine ships THIS record instead of the 28-byte fqsk_base_rec; the host expands it to
 * the 7 ids with a few masks (fqsk_ctx_expand) and goes straight to find_rc_code_context + Encode.
 *
 * Plain C, also compiled by nvcc (__host__ __device__): the same functions run in the CUDA kernel (k_ctx_codes), in the reference-side
 * binding (host/fqsk_live.h) and in the CPU test harness that pins them against ids tapped from the real reference
 * (tests/test_ctx_codes.py, tests/golden/se_ctx_gs1.npz).  `reference:` citations are relative to /root/reference/fqs/.
 */
#ifndef FQSK_CTX_H
#define FQSK_CTX_H

#include <stdint.h>

#ifdef __CUDACC__
#define FQSK_CTX_FN __host__ __device__ static inline
#else
#define FQSK_CTX_FN static inline
#endif

/* field layout of a context id, reference: code_ctx.h:29-60 */
enum {
	FQSK_CTX_SHIFT_POS = 0, FQSK_CTX_SHIFT_LEVEL = 14, FQSK_CTX_SHIFT_COUNTS = 17 /* 4 fields of 7 bits */, FQSK_CTX_SHIFT_RSYM = 45,
	FQSK_CTX_SHIFT_LETMAX = 49, FQSK_CTX_SHIFT_CORZONE = 52, FQSK_CTX_BITS = 55, FQSK_CTX_EOR = 5 /* eor_size, code_ctx.h:73 */
};
#define FQSK_CTX_MASK_POS 0x3FFFull
#define FQSK_CTX_EN_POS (0x3FFFull << FQSK_CTX_SHIFT_POS)
#define FQSK_CTX_EN_COUNT(i) (0x7Full << (FQSK_CTX_SHIFT_COUNTS + 7 * (i)))
#define FQSK_CTX_EN_COUNTS_ALL (0xFFFFFFFull << FQSK_CTX_SHIFT_COUNTS)
#define FQSK_CTX_EN_RSYM (0xFull << FQSK_CTX_SHIFT_RSYM)
#define FQSK_CTX_EN_LETMAX (0x7ull << FQSK_CTX_SHIFT_LETMAX)
#define FQSK_CTX_EN_CORZONE (0x7ull << FQSK_CTX_SHIFT_CORZONE)
#define FQSK_CTX_MAX_READ 65000u     /* positions beyond this would carry out of the 14-bit position field (code_ctx.cpp:281, 322) */

/* One coded base.  a: the level-6 context id in bits 0..54 (bits 55..63 of every id are ones), bits 55..57 = rank of the true symbol
 * among the counts (r_sym, 4 for N), bit 58 = 1 when the base is coded with counts (counts_level != none && N_run_len < 2, dna.cpp:737;
 * 0: plain-letter coding, the rest of the record is zero).  b: what the coarser levels need besides masks -- bits 0..6 / 7..13 the two
 * largest counts at resolution 1, 14..20 / 21..27 the two smallest at resolution 0, 28..34 / 35..41 the two largest at resolution 2,
 * 42..55 the position field of levels 1-5. */
typedef struct fqsk_ctx_rec { uint64_t a, b; } fqsk_ctx_rec;

/* convert_lev_{1,2_4,3}_count (code_ctx.cpp:26-239): identity below `ident`, then one code per upper bound, then a last code */
FQSK_CTX_FN uint32_t fqsk_ctx_quant(uint32_t count, uint32_t ident, const uint16_t *ub, int n_ub) {
	if (count < ident) return count;
	for (int i = 0; i < n_ub; ++i) if (count < ub[i]) return ident + (uint32_t) i;
	return ident + (uint32_t) n_ub;
}
/* convert_count (code_ctx.cpp:15-23): table by counts_level (pmer -> lev_1, bmer -> lev_3, everything else -> lev_2_4), resolution cnt_lev */
FQSK_CTX_FN uint32_t fqsk_ctx_convert(uint32_t count, uint32_t level, uint32_t cnt_lev) {
	const uint32_t flag = cnt_lev << 5;
	if (level == 1) {   /* code_ctx.cpp:26-88 */
		const uint16_t u0[4] = {8, 16, 32, 64};
		const uint16_t u1[21] = {16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 176, 192, 224, 288, 384, 512, 1024, 2048};
		return flag + (cnt_lev == 0 ? fqsk_ctx_quant(count, 5, u0, 4) : fqsk_ctx_quant(count, 8, u1, 21));
	}
	const uint16_t a0[1] = {5}, a1[4] = {8, 13, 20, 30};
	if (cnt_lev == 0) return flag + fqsk_ctx_quant(count, 3, a0, 1);
	if (cnt_lev == 1) return flag + fqsk_ctx_quant(count, 5, a1, 4);
	if (level == 3) {   /* code_ctx.cpp:172-239 */
		const uint16_t b2[4] = {13, 20, 30, 50};
		const uint16_t b3[14] = {18, 20, 25, 30, 40, 50, 60, 64, 68, 72, 76, 80, 84, 88};
		return flag + (cnt_lev == 2 ? fqsk_ctx_quant(count, 10, b2, 4) : fqsk_ctx_quant(count, 15, b3, 14));
	}
	/* code_ctx.cpp:91-169 */
	const uint16_t c2[4] = {15, 20, 30, 50};
	const uint16_t c3[19] = {16, 24, 32, 48, 64, 128, 256, 512, 1024, 2048, 2080, 2112, 2176, 2240, 2304, 2432, 2560, 2816, 3072};
	return flag + (cnt_lev == 2 ? fqsk_ctx_quant(count, 10, c2, 4) : fqsk_ctx_quant(count, 10, c3, 19));
}
/* rank (dna.cpp:177-193): position of the true symbol when the counts are ordered, ties by the running letter totals, then by symbol */
FQSK_CTX_FN uint32_t fqsk_ctx_rank(const uint32_t c[4], const uint64_t sl[4], uint32_t sym) {
	if (sym == 4) return 4;
	uint32_t r = 0;
	for (uint32_t i = 0; i < 4; ++i) {
		if (c[i] != c[sym]) r += c[sym] < c[i];
		else if (sl[i] != sl[sym]) r += sl[sym] < sl[i];
		else r += sym > i;
	}
	return r;
}
/* position field (code_ctx.cpp:276-281 for fine == 0, 317-322 for fine == 1) */
FQSK_CTX_FN uint64_t fqsk_ctx_pos_field(uint32_t pos, uint32_t read_len, uint32_t limit, int fine) {
	if (pos < limit) return (uint64_t) pos + (fine ? (1u << 13) : 0u);
	if (pos + FQSK_CTX_EOR >= read_len) return FQSK_CTX_MASK_POS - (uint64_t) (read_len - pos);
	return (uint64_t) limit + pos / (fine ? 8u : 16u) + (fine ? (1u << 13) : 0u);
}

/* The record of one base.  c, level, rough, cor_pos: as in fqsk_base_rec; i = the position the coder's loop is at (record.pos), pos /
 * read_len = what compress_suffix passes to determine_ctx_codes (i and size, or size - i - 1 and ~0u for the reversed part of a mate,
 * dna.cpp:747-752); n_run = N_run_len; r_hist = ctx_r_sym (8 bits, newest base in bit 0); sl = s_letters; p / s / b = k-mer lengths. */
FQSK_CTX_FN fqsk_ctx_rec fqsk_ctx_make(const uint32_t c[4], uint32_t level, uint32_t rough, uint32_t cor_pos, uint32_t i, uint32_t pos, uint32_t read_len,
                                        uint32_t sym, uint32_t n_run, uint32_t r_hist, const uint64_t sl[4], uint32_t p_len, uint32_t s_len, uint32_t b_len) {
	fqsk_ctx_rec o; o.a = 0; o.b = 0;
	if (level == 0 || n_run >= 2) return o;                                  /* dna.cpp:737 */
	const int cor_dist = level == 1 ? (int) p_len : level == 2 ? (int) s_len : (int) b_len;     /* dna.cpp:739 */
	const int d = (int) i - (int) cor_pos;
	uint32_t cor_zone = d < cor_dist ? (uint32_t) (1 + 2 * (cor_dist - d) / cor_dist) : 0u;  /* dna.cpp:741 */
	if (rough) cor_zone = 3;
	uint32_t srt[4] = {c[0], c[1], c[2], c[3]};                               /* sort_copy_stats (utils.cpp:109-126): descending */
	for (int x = 0; x < 3; ++x) for (int y = 0; y < 3 - x; ++y) if (srt[y] < srt[y + 1]) { uint32_t t = srt[y]; srt[y] = srt[y + 1]; srt[y + 1] = t; }
	const uint32_t limit = level == 1 ? p_len : level == 2 ? s_len : b_len;  /* pos_limit (code_ctx.cpp:263) */
	uint32_t lm = 0;                                                          /* let_max_element (code_ctx.cpp:327-338) */
	for (uint32_t q = 1; q < 4; ++q) if (c[q] > c[lm] || (c[q] == c[lm] && sl[q] > sl[lm])) lm = q;
	uint32_t pc = 0;
	for (uint32_t q = r_hist & 0xFFu; q; q &= q - 1) ++pc;                    /* transform_r_sym = popcnt (code_ctx.cpp:371-373) */
	uint64_t id = (uint64_t) level << FQSK_CTX_SHIFT_LEVEL;
	id += (uint64_t) fqsk_ctx_convert(srt[0], level, 3) << (FQSK_CTX_SHIFT_COUNTS + 0);
	id += (uint64_t) fqsk_ctx_convert(srt[1], level, 3) << (FQSK_CTX_SHIFT_COUNTS + 7);
	id += (uint64_t) fqsk_ctx_convert(srt[2], level, 1) << (FQSK_CTX_SHIFT_COUNTS + 14);
	id += (uint64_t) fqsk_ctx_convert(srt[3], level, 1) << (FQSK_CTX_SHIFT_COUNTS + 21);
	id += (uint64_t) pc << FQSK_CTX_SHIFT_RSYM;
	id += (uint64_t) lm << FQSK_CTX_SHIFT_LETMAX;
	id += (uint64_t) cor_zone << FQSK_CTX_SHIFT_CORZONE;
	id += fqsk_ctx_pos_field(pos, read_len, limit, 1) << FQSK_CTX_SHIFT_POS;
	o.a = id | ((uint64_t) fqsk_ctx_rank(c, sl, sym) << 55) | (1ull << 58);
	o.b = (uint64_t) fqsk_ctx_convert(srt[0], level, 1) | ((uint64_t) fqsk_ctx_convert(srt[1], level, 1) << 7) | ((uint64_t) fqsk_ctx_convert(srt[2], level, 0) << 14) |
	      ((uint64_t) fqsk_ctx_convert(srt[3], level, 0) << 21) | ((uint64_t) fqsk_ctx_convert(srt[0], level, 2) << 28) | ((uint64_t) fqsk_ctx_convert(srt[1], level, 2) << 35) |
	      (fqsk_ctx_pos_field(pos, read_len, limit, 0) << 42);
	return o;
}

FQSK_CTX_FN int fqsk_ctx_coded(const fqsk_ctx_rec *r) { return (int) ((r->a >> 58) & 1u); }
FQSK_CTX_FN uint32_t fqsk_ctx_rsym(const fqsk_ctx_rec *r) { return (uint32_t) ((r->a >> 55) & 7u); }
/* the 7 ids a_code_ctx[0..6] of determine_ctx_codes (code_ctx.cpp:268-323) from the record */
FQSK_CTX_FN void fqsk_ctx_expand(const fqsk_ctx_rec *r, uint64_t out[7]) {
	const uint64_t ones = ~0ull << FQSK_CTX_BITS;
	const uint64_t c6 = (r->a & ~ones) | ones, b = r->b;
	const uint64_t c5 = (c6 & ~FQSK_CTX_EN_POS) | (((b >> 42) & FQSK_CTX_MASK_POS) << FQSK_CTX_SHIFT_POS);
	const uint64_t c4 = c5 | FQSK_CTX_EN_LETMAX;
	const uint64_t c3 = (c4 & ~FQSK_CTX_EN_COUNTS_ALL) | (((b >> 28) & 0x7F) << (FQSK_CTX_SHIFT_COUNTS + 0)) | (((b >> 35) & 0x7F) << (FQSK_CTX_SHIFT_COUNTS + 7)) |
	                    (((b >> 14) & 0x7F) << (FQSK_CTX_SHIFT_COUNTS + 14)) | (((b >> 21) & 0x7F) << (FQSK_CTX_SHIFT_COUNTS + 21));
	const uint64_t c2 = (c3 & ~(FQSK_CTX_EN_COUNT(0) | FQSK_CTX_EN_COUNT(1))) | ((b & 0x7F) << (FQSK_CTX_SHIFT_COUNTS + 0)) | (((b >> 7) & 0x7F) << (FQSK_CTX_SHIFT_COUNTS + 7));
	out[6] = c6; out[5] = c5; out[4] = c4; out[3] = c3; out[2] = c2;
	out[1] = c2 | FQSK_CTX_EN_CORZONE | FQSK_CTX_EN_RSYM;
	out[0] = ~0ull;
}

#endif /* FQSK_CTX_H */
