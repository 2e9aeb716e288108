// TEST INFRASTRUCTURE: drives include/fqsk_ctx.h (the functions the CUDA kernel k_ctx_codes and the reference-side binding share) over
// a tapped record stream on the CPU, so that tests/test_ctx_codes.py can pin them against context ids tapped from the real reference.
// Single-end, original order: markers 0xFFFFFFFF (read, c0 = size), 0xFFFFFFFD (duplicate), 0xFFFFFFFE (sync); everything else is a base.
#include <cstdint>
#include <cstring>

#include "fqsk.h"
#include "fqsk_ctx.h"

static inline uint32_t code_of(uint8_t ch) { return ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 4u; }

extern "C" uint64_t ctx_from_tap(const fqsk_base_rec *recs, uint64_t n, const uint8_t *slab, const uint64_t *off, const uint32_t *len, uint32_t n_reads,
                                 uint32_t p_len, uint32_t s_len, uint32_t b_len, uint32_t prefix_len, uint64_t *out /* [cap][8] */, uint64_t cap, fqsk_ctx_rec *raw /* per base, may be null */) {
	uint64_t sl[4] = {0, 0, 0, 0};
	uint64_t n_out = 0, n_base = 0;
	int64_t r = -1;
	bool dup = false;
	uint32_t n_run = 0, r_hist = 0;
	auto finish_read = [&]() {      // update_s_letters (dna.cpp:2047-2057) unless the read was a duplicate (dna.cpp:1532)
		if (r < 0 || dup) return;
		for (uint32_t i = 0; i < len[r]; ++i) { uint32_t c = code_of(slab[off[r] + i]); if (c < 4) { ++sl[c]; ++sl[3 - c]; } }
	};
	for (uint64_t g = 0; g < n; ++g) {
		const fqsk_base_rec &q = recs[g];
		if (q.pos == 0xFFFFFFFFu) { finish_read(); ++r; dup = false; n_run = 0; r_hist = 0; if (r >= (int64_t) n_reads || q.counts[0] != len[r]) return ~0ull; continue; }
		if (q.pos == 0xFFFFFFFDu) { dup = true; continue; }
		if (q.pos >= 0xFFFFFFF0u) continue;
		(void) prefix_len;
		const uint32_t sym = code_of(slab[off[r] + q.pos]);
		fqsk_ctx_rec c = fqsk_ctx_make(q.counts, q.level, q.rough, q.cor_pos, q.pos, q.pos, len[r], sym, n_run, r_hist, sl, p_len, s_len, b_len);
		if (raw) raw[n_base] = c;
		++n_base;
		const int coded = fqsk_ctx_coded(&c);
		if (coded) {
			if (n_out < cap) { fqsk_ctx_expand(&c, out + 8 * n_out); out[8 * n_out + 7] = fqsk_ctx_rsym(&c); }
			++n_out;
		}
		r_hist = ((r_hist << 1) | (uint32_t) (coded && fqsk_ctx_rsym(&c) == 0)) & 0xFFu;      // update_ctx_r_sym (dna.cpp:664-671)
		n_run = sym == 4 ? n_run + 1 : 0;                                                     // dna.cpp:805-808
	}
	return n_out;
}
