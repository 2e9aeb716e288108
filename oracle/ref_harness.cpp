// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" driver over the REFERENCE's own headers (kmer.h, ht_kmer.h, bit_vec.h, utils.h of
// refresh-bio/fqsqueezer 1.1), compiled by oracle/build_ref.py against a scratch copy of those headers
// (plus read-only visitors) into oracle/_ref/libfqs_ref.so.  It exists so that the CPU restatement in
// oracle/fqs_oracle.cpp can be pinned against the real classes at unit level (tests/test_oracle_vs_ref.py).
// No reference source is stored in this repository; this file only calls the reference's public methods.
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>
#include <tuple>
#include <mutex>
#include <iostream>
#include <chrono>
#include <atomic>
#include <algorithm>

#include "defs.h"
#include "utils.h"
#include "kmer.h"
#include "ht_kmer.h"
#include "bit_vec.h"

namespace {
struct HtBase {
	virtual ~HtBase() {}
	virtual void insert(uint64_t x, CCounterIncrementer *c) = 0;
	virtual bool find(const CKmer &k, stats_t &s, CCounterIncrementer &c) = 0;
	virtual int count(uint64_t x) = 0;
	virtual void clear0() = 0;
	virtual uint64_t dump(uint64_t *k, uint32_t *c, uint64_t cap) = 0;
	virtual uint64_t no_kmers() = 0;
	virtual uint32_t prefix_len() = 0;
};
template <typename T> struct Ht : HtBase {
	CHT_kmer<T> ht;
	Ht(uint32_t k, uint32_t cb) : ht(k, cb, 0.8) {}
	void insert(uint64_t x, CCounterIncrementer *c) override { ht.insert(x, c); }
	bool find(const CKmer &k, stats_t &s, CCounterIncrementer &c) override { return ht.find(k, s, c); }
	int count(uint64_t x) override { return ht.count(x); }
	void clear0() override { ht.clear(0); }
	uint64_t dump(uint64_t *k, uint32_t *c, uint64_t cap) override {
		uint64_t n = 0;
		ht.tap_for_each([&](uint64_t km, uint32_t cnt) { if (n < cap) { k[n] = km; c[n] = cnt; } ++n; });
		return n;
	}
	uint64_t no_kmers() override { return ht.get_no_kmers(); }
	uint32_t prefix_len() override { return ht.get_ht_prefix_len(); }
};
}  // namespace

extern "C" {

// ---- CCounterIncrementer (utils.h:256-335) ----
void *ref_cinc_new(uint32_t thr, uint32_t mult, uint32_t max_val) { auto *c = new CCounterIncrementer(); c->Reset(thr, mult, max_val); return c; }
void ref_cinc_free(void *c) { delete (CCounterIncrementer *) c; }
void ref_cinc_inc1(void *c, const uint32_t *cnt, uint64_t n, uint32_t *out) { for (uint64_t i = 0; i < n; ++i) out[i] = ((CCounterIncrementer *) c)->Increment(cnt[i]); }
void ref_cinc_incn(void *c, const uint32_t *cnt, const uint32_t *inc, uint64_t n, uint32_t *out) { for (uint64_t i = 0; i < n; ++i) out[i] = ((CCounterIncrementer *) c)->Increment(cnt[i], inc[i]); }

// ---- std::mt19937 seeded as the reference seeds it (utils.h:298) ----
void ref_mt_stream(uint32_t seed, uint64_t n, uint32_t *out) { std::mt19937 mt; mt.seed(seed); for (uint64_t i = 0; i < n; ++i) out[i] = (uint32_t) mt(); }

// ---- CHT_kmer<T> (ht_kmer.h:29-554) ----
void *ref_ht_new(uint32_t k, uint32_t counter_bits, uint32_t item_bytes) { return item_bytes == 4 ? (HtBase *) new Ht<uint32_t>(k, counter_bits) : (HtBase *) new Ht<uint64_t>(k, counter_bits); }
void ref_ht_free(void *h) { delete (HtBase *) h; }
void ref_ht_insert(void *h, void *cinc, const uint64_t *kmers, uint64_t n) { for (uint64_t i = 0; i < n; ++i) ((HtBase *) h)->insert(kmers[i], (CCounterIncrementer *) cinc); }
// find(): one CKmer per query, rebuilt from (dir, rc, cur_size); results 4 x u32 per query, processed in order
void ref_ht_find(void *h, void *cinc, uint32_t k, const uint64_t *dir, const uint64_t *rc, const uint32_t *cur, uint64_t n, uint32_t *counts, uint8_t *found) {
	for (uint64_t i = 0; i < n; ++i) {
		CKmer km(dir[i], rc[i], k, cur[i], kmer_mode_t::canonical);
		stats_t s;
		bool f = ((HtBase *) h)->find(km, s, *(CCounterIncrementer *) cinc);
		for (int j = 0; j < 4; ++j) counts[4 * i + j] = s[j];
		if (found) found[i] = f;
	}
}
void ref_ht_count(void *h, const uint64_t *kmers, uint64_t n, uint32_t *out) { for (uint64_t i = 0; i < n; ++i) out[i] = (uint32_t) ((HtBase *) h)->count(kmers[i]); }
void ref_ht_clear0(void *h) { ((HtBase *) h)->clear0(); }
uint64_t ref_ht_dump(void *h, uint64_t *kmers, uint32_t *counts, uint64_t cap) { return ((HtBase *) h)->dump(kmers, counts, cap); }
uint64_t ref_ht_no_kmers(void *h) { return ((HtBase *) h)->no_kmers(); }
uint32_t ref_ht_prefix_len(void *h) { return ((HtBase *) h)->prefix_len(); }

// ---- CKmer (kmer.h:18-540): replay a script of register operations on one canonical register ----
// op: 0 reset, 1 insert(a), 2 insert_zero, 3 replace_last(a), 4 replace(a, b), 5 insert_front(a), 6 shorten(a)
// out per op: dir, rc, normalized, aligned_dir, aligned_rc, kernel_canonical (u64 x 6), then is_dir / cur / full (u32 x 3)
void ref_kmer_script(uint32_t k, const uint32_t *ops, uint64_t n_ops, uint64_t *out64, uint32_t *out32) {
	CKmer km(k, kmer_mode_t::canonical);
	for (uint64_t i = 0; i < n_ops; ++i) {
		uint32_t op = ops[3 * i], a = ops[3 * i + 1], b = ops[3 * i + 2];
		switch (op) {
		case 0: km.Reset(); break;
		case 1: km.insert(a); break;
		case 2: km.insert_zero(); break;
		case 3: km.replace_last(a); break;
		case 4: km.replace(a, b); break;
		case 5: km.insert_front(a); break;
		case 6: km.shorten(a); break;
		}
		out64[6 * i + 0] = km.data_dir();
		out64[6 * i + 1] = km.data_rc();
		out64[6 * i + 2] = km.data_normalized();
		out64[6 * i + 3] = km.get_cur_size() ? km.data_aligned_dir() : 0;
		out64[6 * i + 4] = km.get_cur_size() ? km.data_aligned_rc() : 0;
		out64[6 * i + 5] = km.kernel_canonical();
		out32[3 * i + 0] = km.is_normalized_dir();
		out32[3 * i + 1] = km.get_cur_size();
		out32[3 * i + 2] = km.is_full();
	}
}

// ---- TSmallIntVector<2> (bit_vec.h:17-231) ----
void *ref_siv_new(uint32_t key_size) { return new TSmallIntVector<2>(key_size); }
void ref_siv_free(void *s) { delete (TSmallIntVector<2> *) s; }
uint64_t ref_siv_increment(void *s, const uint64_t *idx, uint64_t n) { uint64_t nf = 0; for (uint64_t i = 0; i < n; ++i) nf += ((TSmallIntVector<2> *) s)->increment(idx[i]); return nf; }
void ref_siv_test(void *s, const uint64_t *idx, uint64_t n, uint32_t *out) { for (uint64_t i = 0; i < n; ++i) out[i] = (uint32_t) ((TSmallIntVector<2> *) s)->test(idx[i]); }
void ref_siv_counts(void *s, const uint64_t *idx, uint64_t n, uint32_t *out) { for (uint64_t i = 0; i < n; ++i) { stats_t c; ((TSmallIntVector<2> *) s)->counts(idx[i], c); for (int j = 0; j < 4; ++j) out[4 * i + j] = c[j]; } }
void ref_siv_test_shorter(void *s, const uint64_t *idx, const uint32_t *size_bits, uint64_t n, uint64_t *out) { for (uint64_t i = 0; i < n; ++i) out[i] = ((TSmallIntVector<2> *) s)->test_shorter(idx[i], size_bits[i]); }
uint64_t ref_siv_dump(void *s, uint64_t *idx, uint32_t *val, uint64_t cap) { uint64_t n = 0; ((TSmallIntVector<2> *) s)->tap_for_each([&](uint64_t i, uint64_t v) { if (n < cap) { idx[n] = i; val[n] = (uint32_t) v; } ++n; }); return n; }

// ---- CHT_pair_kmers (ht_kmer.h:559-663, ht_kmer.cpp:17-230) ----
void *ref_pair_new(uint32_t k, uint64_t parts) { return new CHT_pair_kmers(k, parts); }
void ref_pair_free(void *p) { delete (CHT_pair_kmers *) p; }
void ref_pair_insert(void *p, const uint64_t *key, const uint64_t *val, const uint64_t *cnt, uint64_t n) { for (uint64_t i = 0; i < n; ++i) ((CHT_pair_kmers *) p)->insert(key[i], val[i], cnt[i]); }
uint64_t ref_pair_find(void *p, uint64_t key, uint64_t *out, uint64_t cap) { std::vector<uint64_t> v; ((CHT_pair_kmers *) p)->find(key, v); for (uint64_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i]; return v.size(); }
void ref_pair_count(void *p, const uint64_t *key, const uint64_t *val, uint64_t n, uint64_t *out) { for (uint64_t i = 0; i < n; ++i) out[i] = ((CHT_pair_kmers *) p)->count(key[i], val[i]); }
uint64_t ref_pair_part_id(void *p, uint64_t key) { return ((CHT_pair_kmers *) p)->get_part_id(key); }
uint64_t ref_pair_dump(void *p, uint64_t *key, uint64_t *vc, uint64_t cap) { uint64_t n = 0; ((CHT_pair_kmers *) p)->tap_for_each([&](uint64_t k, uint64_t v) { if (n < cap) { key[n] = k; vc[n] = v; } ++n; }); return n; }
void ref_pair_clear(void *p) { ((CHT_pair_kmers *) p)->clear(); }

}  // extern "C"
