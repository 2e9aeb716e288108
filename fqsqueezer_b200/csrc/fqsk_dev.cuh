// fqsk_dev.cuh -- device-side primitives of the B200 k-mer statistics engine (sm_100a).
//
// Data layout in HBM:
//   * b-/s-mer tables: 2^B buckets of 8 x u32 (one 32-byte sector). A k-mer lives in the bucket chosen by a bijective
//     mix of its KERNEL (symbols s2..s(k-3), kmer.h:199-202) so the 4 next-symbol siblings of a context -- which differ
//     only in s(k-1) (dir-oriented) or s0 (rc-oriented) -- share one sector, and one sector read answers
//     CHT_kmer::find_full (ht_kmer.h:205-263).  item = [1 | rem | s0 s1 s(k-2) s(k-1) | counter]; 0 = empty (bit 31 marks an
//     occupied slot, so an item stays distinguishable from an empty slot whatever its counter is).
//     Buckets only hold native items; a k-mer whose bucket is full goes to a small u64 open-addressing stash.
//     Reference lookups depend only on table CONTENTS, so this layout is free to differ from ht_kmer.h:49-76.
//   * p-mer array: 2-bit saturating fields, 16 per u32, direct-addressed (bit_vec.h:17-231).
// `reference:` citations are relative to /root/reference/fqs/.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define FQSK_DEV __device__ __forceinline__
#define FQSK_HD __host__ __device__ __forceinline__

namespace fqsk {

// ------------------------------------------------------------------------------------------------------------------
// canonical rolling register (reference: kmer.h:18-540).  cur (symbols held) is tracked by the caller: all six
// registers of a read advance in lock-step (dna.cpp:687-693, 810-816), so cur = min(k, symbols pushed).
// ------------------------------------------------------------------------------------------------------------------
// programmatic dependent launch (see pdl() in fqsk.cu): let the next kernel of the stream be scheduled now, then wait until the
// previous one has completed and its writes are visible.  First statement of every kernel; a no-op for plain launches.
FQSK_DEV void pdl_enter() {
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

struct KReg { uint64_t dir, rc; };

// ---- 2-bit packed symbols: 32 per 64-bit word, symbol j of a word at bits 63-2j .. 62-2j (the layout of a left-aligned CKmer) ----
// word of 32 symbols from two warp ballots (bit 1 / bit 0 of every lane's symbol); lane 0 = first symbol
FQSK_DEV uint64_t pk_spread(uint32_t x) {            // bit j -> bit 2j
	uint64_t v = x;
	v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
	v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
	v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
	v = (v | (v << 2)) & 0x3333333333333333ull;
	v = (v | (v << 1)) & 0x5555555555555555ull;
	return v;
}
FQSK_DEV uint64_t pk_from_ballots(unsigned b1, unsigned b0) { return (pk_spread(__brev(b1)) << 1) | pk_spread(__brev(b0)); }
// len (< 32) symbols starting at symbol a of the sequence whose words are w[0], w[1], ...: hi = w[a >> 5], lo = w[(a >> 5) + 1]
// (the sequence needs one word of slack); left-aligned, rest zero
FQSK_DEV uint64_t pk_window(uint64_t hi, uint64_t lo, uint32_t a, uint32_t len) {
	const uint32_t sh = 2 * (a & 31);
	uint64_t x = sh ? (hi << sh) | (lo >> (64 - sh)) : hi;
	return len ? x & (~0ull << (64 - 2 * len)) : 0ull;
}

FQSK_HD uint64_t kr_top_mask(uint32_t k) { return ~0ull << (64 - 2 * k); }
FQSK_HD uint64_t kr_kernel_mask(uint32_t k) { return ((1ull << (2 * k - 8)) - 1ull) << (64 - 2 * k + 4); }
// kmer.h:73-108 (insert_zero == insert of symbol 0: the rc side receives T)
FQSK_HD void kr_push(KReg &r, uint32_t k, uint32_t cur_before, uint64_t s) {
	r.rc = ((r.rc >> 2) + ((3 - s) << 62)) & kr_top_mask(k);
	if (cur_before >= k) r.dir = (r.dir << 2) + (s << (64 - 2 * k));
	else r.dir += s << (62 - 2 * cur_before);
}
FQSK_HD void kr_set_last(KReg &r, uint32_t cur, uint64_t s) {  // kmer.h:153-174
	uint32_t sh = 64 - 2 * cur;
	r.dir = (r.dir & ~(3ull << sh)) + (s << sh);
	r.rc = ((r.rc << 2) >> 2) + ((3 - s) << 62);
}
FQSK_HD void kr_set(KReg &r, uint32_t cur, uint64_t s, uint32_t pos) {  // kmer.h:144-167
	uint32_t sh = 62 - 2 * pos;
	r.dir = (r.dir & ~(3ull << sh)) + (s << sh);
	sh = 64 - 2 * cur + 2 * pos;
	r.rc = (r.rc & ~(3ull << sh)) + ((3 - s) << sh);
}
FQSK_HD uint64_t kr_sym(const KReg &r, uint32_t pos) { return (r.dir >> (62 - 2 * pos)) & 3; }
FQSK_HD bool kr_is_dir(const KReg &r, uint32_t k) { uint64_t m = kr_kernel_mask(k); return (r.dir & m) < (r.rc & m); }  // kmer.h:380-385
FQSK_HD uint64_t kr_norm(const KReg &r, uint32_t k) { return kr_is_dir(r, k) ? r.dir : r.rc; }  // kmer.h:366-377

// ------------------------------------------------------------------------------------------------------------------
// approximate counters (reference: utils.h:256-335).  v_mapping has the closed form thr + mult * n(n+1)/2.
// ------------------------------------------------------------------------------------------------------------------
struct CIncP { uint32_t thr, mult, top; };

FQSK_HD uint32_t ci_real_of(const CIncP &p, uint32_t c) {
	if (c <= p.thr) return c;
	if (c > p.top) c = p.top;  // guard entry v_mapping[max+1] = v_mapping[max] (utils.h:312)
	uint32_t n = c - p.thr;
	return p.thr + p.mult * (n * (n + 1) / 2);
}
FQSK_HD uint32_t ci_to_real(const CIncP &p, uint32_t c) { return c <= p.thr ? c : (ci_real_of(p, c) + ci_real_of(p, c + 1)) / 2; }  // utils.h:264-270
// utils.h:272-291; *need_draw tells the caller whether one mt19937 output is consumed; `draw` is that output.
FQSK_HD uint32_t ci_code_floor(const CIncP &p, uint32_t r) {  // largest code in [thr, top] with real_of(code) <= r
	uint32_t d = (r - p.thr) / p.mult;  // n(n+1)/2 <= d
	uint32_t n = (uint32_t) ((sqrt(8.0 * (double) d + 1.0) - 1.0) * 0.5);
	while ((uint64_t) (n + 1) * (n + 2) / 2 <= d) ++n;
	while (n > 0 && (uint64_t) n * (n + 1) / 2 > d) --n;
	uint32_t code = p.thr + n;
	return code > p.top ? p.top : code;
}

// A cursor into one pre-generated mt19937 stream (tempered 32-bit outputs in HBM).  `next` is relative to the read's
// guessed starting offset; the fix point in fqsk.cu makes guess == truth before results are released.
struct DrawCursor {
	const uint32_t *ring;  // ring buffer of tempered outputs
	uint64_t mask;         // ring size - 1
	uint64_t pos0;         // absolute index of the stream's `consumed` position
	uint64_t avail;        // outputs generated beyond that position
	uint64_t base;         // guessed offset of this consumer
	uint32_t used;         // draws consumed by this consumer so far
	int *overflow;         // set when the pre-generated window is too short (host extends and replays)
	__device__ uint32_t next() {
		uint64_t i = base + used++;
		if (i >= avail) { *overflow = 1; return 0; }
		return ring[(pos0 + i) & mask];
	}
};

FQSK_DEV uint32_t ci_from_real(const CIncP &p, uint32_t r, DrawCursor &dc) {
	if (r <= p.thr) return r;
	uint32_t code = ci_code_floor(p, r);
	if (code >= p.top) return p.top;
	uint32_t rest = r - ci_real_of(p, code);
	uint32_t width = ci_real_of(p, code + 1) - ci_real_of(p, code);
	if (dc.next() % width < rest) ++code;
	return code;
}
FQSK_DEV uint32_t ci_plus(const CIncP &p, uint32_t c, uint32_t inc, DrawCursor &dc) { return ci_from_real(p, ci_to_real(p, c) + ci_to_real(p, inc), dc); }  // utils.h:328-334

// ------------------------------------------------------------------------------------------------------------------
// bucketed k-mer table
// ------------------------------------------------------------------------------------------------------------------
struct HtDev {
	uint32_t *main;                 // 8 << B items
	unsigned long long *stash;      // 1 << stash_log2 items: ((aligned k-mer + 1) << cbits) | counter, 0 = empty
	uint32_t k, cbits, W, B, rem_bits, top, stash_log2, mix_sh;
	uint64_t maskW;
	unsigned long long *n_items;    // device counters: [0] main items, [1] stash items
	int *err;                       // set when an insert finds the stash full (a single sync brought more new k-mers than the table can take before it
	                                // grows): the probe loops stay bounded, the host reports FQSK_E_CAPACITY at its next look
	// one bit per bucket, set when the first item of the bucket is created and never cleared (16 MiB for 2^27 buckets: L2-resident).
	// k_rough tests it before reading a neighbour's bucket while the table is sparse (occ_read != null): in the first blocks of a
	// file nearly all of the 4(k-1) trials of a rough search land on empty buckets (measured on block 10 of config 2: DRAM reads of
	// the kernel 345 -> 132 MB, 86.7 -> 75.6 us).  Other readers leave occ_read null: for them the extra dependent L2 access costs
	// more than the skipped DRAM access saves.
	uint32_t *occ; const uint32_t *occ_read;
	// hash sharding over the GPUs of one box: the table is split by the reference's own owner key of a k-mer,
	// ((x >> 46) & 0x3fff) % world (dna.cpp:825, 836, 2382-2388) -- bits of symbols s2..s8, i.e. of the kernel, so the 4
	// siblings of a context share the owner.  All shards have the same geometry; main / stash are THIS rank's shard (the only
	// one it ever writes), peer_* are every rank's shard (NVLink peer mappings, [rank] = own) for lookups.
	uint32_t world, rank;
	const uint32_t *peer_main[8];
	const unsigned long long *peer_stash[8];
};
static const uint32_t FQSK_MAX_WORLD = 8;

static const uint64_t MIX_C1 = 0x9E3779B97F4A7C15ull, MIX_C2 = 0xD6E8FEB86659FD93ull;

FQSK_HD uint64_t ht_mix(const HtDev &t, uint64_t x) {  // bijection on W-bit words
	x = (x * MIX_C1) & t.maskW; x ^= x >> t.mix_sh;
	x = (x * MIX_C2) & t.maskW; x ^= x >> t.mix_sh;
	return x;
}
FQSK_HD uint64_t ht_kernel(const HtDev &t, uint64_t x) { return (x >> (64 - 2 * t.k + 4)) & t.maskW; }
FQSK_HD uint32_t ht_ends(const HtDev &t, uint64_t x) { return (uint32_t) (((x >> 60) & 0xF) << 4 | ((x >> (64 - 2 * t.k)) & 0xF)); }
FQSK_HD uint64_t ht_slot_count() { return 8; }

struct HtKey { uint64_t bucket; uint32_t q; uint32_t owner; uint64_t kal; uint64_t h; };  // q = item without counter; owner = rank holding the k-mer
FQSK_HD uint32_t ht_owner(uint32_t world, uint64_t x) { return world > 1 ? (uint32_t) (((x >> 46) & 0x3fff) % world) : 0u; }
FQSK_HD HtKey ht_key(const HtDev &t, uint64_t x) {
	HtKey k;
	k.h = ht_mix(t, ht_kernel(t, x));
	k.bucket = k.h >> t.rem_bits;
	uint32_t rem = (uint32_t) (k.h & ((1ull << t.rem_bits) - 1));
	k.q = 0x80000000u | (rem << (8 + t.cbits)) | (ht_ends(t, x) << t.cbits);
	k.kal = x >> (64 - 2 * t.k);
	k.owner = ht_owner(t.world, x);
	return k;
}
FQSK_HD uint64_t ht_stash_pos(const HtDev &t, uint64_t h) { return (h * 0xC2B2AE3D27D4EB4Full) >> (64 - t.stash_log2); }

struct Bucket { uint4 lo, hi; };
FQSK_DEV Bucket ht_load_bucket(const HtDev &t, const HtKey &key) {
	const uint4 *p = reinterpret_cast<const uint4 *>(t.peer_main[key.owner] + key.bucket * 8);
	Bucket r;
	if (t.occ_read && !((__ldg(t.occ_read + (key.bucket >> 5)) >> (key.bucket & 31)) & 1u)) {
		r.lo = make_uint4(0, 0, 0, 0); r.hi = make_uint4(0, 0, 0, 0);
		return r;
	}
	r.lo = __ldg(p);
	r.hi = __ldg(p + 1);
	return r;
}
FQSK_DEV uint32_t bucket_item(const Bucket &b, int i) {
	switch (i) { case 0: return b.lo.x; case 1: return b.lo.y; case 2: return b.lo.z; case 3: return b.lo.w;
	             case 4: return b.hi.x; case 5: return b.hi.y; case 6: return b.hi.z; default: return b.hi.w; }
}

// The 4 next-symbol counters of a context (ht_kmer.h:205-263) from an already loaded bucket; falls through to the
// stash only when the bucket is full.  x = normalised register (context + placeholder), is_dir = its orientation.
FQSK_DEV void ht_ctx_counts_from(const HtDev &t, const HtKey &key, bool is_dir, const Bucket &bk, uint32_t c[4]) {
	uint32_t unk = is_dir ? (3u << t.cbits) : (3u << (t.cbits + 6));
	uint32_t mm = ~(t.top | unk);
	uint32_t sh = is_dir ? t.cbits : t.cbits + 6;
	bool full = true;
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		uint32_t it = bucket_item(bk, i);
		if (it == 0) { full = false; break; }
		if (((it ^ key.q) & mm) == 0) { uint32_t f = (it >> sh) & 3; c[is_dir ? f : 3 - f] += it & t.top; }
	}
	if (!full) return;
	uint64_t ush = is_dir ? 0 : 2 * t.k - 2;  // unknown symbol inside the aligned k-mer
	uint64_t m64 = ~(3ull << ush);
	uint64_t smask = (1ull << t.stash_log2) - 1;
	const unsigned long long *stash = t.peer_stash[key.owner];
	for (uint64_t p = ht_stash_pos(t, key.h), guard = 0; guard <= smask; p = (p + 1) & smask, ++guard) {      // bounded: a full stash never spins
		unsigned long long it = stash[p];
		if (it == 0) break;
		uint64_t kal = (it >> t.cbits) - 1;
		if (((kal ^ key.kal) & m64) == 0) { uint32_t f = (uint32_t) ((kal >> ush) & 3); c[is_dir ? f : 3 - f] += (uint32_t) (it & t.top); }
	}
}
FQSK_DEV void ht_ctx_counts(const HtDev &t, uint64_t x, bool is_dir, uint32_t c[4]) {
	HtKey key = ht_key(t, x);
	Bucket bk = ht_load_bucket(t, key);
	ht_ctx_counts_from(t, key, is_dir, bk, c);
}
// CHT_kmer::count(uint64_t): exact match (ht_kmer.h:441-454)
FQSK_DEV uint32_t ht_count(const HtDev &t, uint64_t x) {
	HtKey key = ht_key(t, x);
	Bucket bk = ht_load_bucket(t, key);
	bool full = true;
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		uint32_t it = bucket_item(bk, i);
		if (it == 0) { full = false; break; }
		if ((it & ~t.top) == key.q) return it & t.top;
	}
	if (!full) return 0;
	uint64_t smask = (1ull << t.stash_log2) - 1;
	const unsigned long long *stash = t.peer_stash[key.owner];
	for (uint64_t p = ht_stash_pos(t, key.h), guard = 0; guard <= smask; p = (p + 1) & smask, ++guard) {
		unsigned long long it = stash[p];
		if (it == 0) return 0;
		if ((it >> t.cbits) == key.kal + 1) return (uint32_t) (it & t.top);
	}
	return 0;
}
// find-or-create for the sync step (ht_kmer.h:330-362).  New slots are claimed with counter 1; the group pass of the sync
// step turns that into the reference's "start at 0, then Increment".  Returns a slot id (main: index, stash: 8<<B + index).
FQSK_DEV uint64_t ht_locate(const HtDev &t, uint64_t x, bool &created, uint32_t claim = 1u) {      // claim: counter a new slot starts with
	HtKey key = ht_key(t, x);
	uint32_t *bp = t.main + key.bucket * 8;
	created = false;
	for (int i = 0; i < 8; ++i) {
		uint32_t it = *((volatile uint32_t *) (bp + i));
		if (it == 0) {
			uint32_t old = atomicCAS(bp + i, 0u, key.q | claim);
			if (old == 0) {
				created = true; atomicAdd(t.n_items, 1ull);
				if (t.occ) { const uint32_t bit = 1u << (key.bucket & 31); if (!(t.occ[key.bucket >> 5] & bit)) atomicOr(t.occ + (key.bucket >> 5), bit); }
				return key.bucket * 8 + i;
			}
			it = old;
		}
		if ((it & ~t.top) == key.q) return key.bucket * 8 + i;
	}
	uint64_t smask = (1ull << t.stash_log2) - 1;
	unsigned long long fresh = ((key.kal + 1) << t.cbits) | (unsigned long long) claim;
	uint64_t probes = 0;
	for (uint64_t p = ht_stash_pos(t, key.h);; p = (p + 1) & smask) {
		if (++probes > smask) { if (t.err) *t.err = 1; return (8ull << t.B) + p; }      // stash full: give up (the host fails the call), never spin
		unsigned long long it = *((volatile unsigned long long *) (t.stash + p));
		if (it == 0) {
			unsigned long long old = atomicCAS(t.stash + p, 0ull, fresh);
			if (old == 0) { created = true; atomicAdd(t.n_items + 1, 1ull); return (8ull << t.B) + p; }
			it = old;
		}
		if ((it >> t.cbits) == key.kal + 1) return (8ull << t.B) + p;
	}
}
FQSK_DEV uint32_t ht_slot_get(const HtDev &t, uint64_t slot) {
	uint64_t nm = 8ull << t.B;
	return slot < nm ? (t.main[slot] & t.top) : (uint32_t) (t.stash[slot - nm] & t.top);
}
FQSK_DEV void ht_slot_set(const HtDev &t, uint64_t slot, uint32_t cnt) {
	uint64_t nm = 8ull << t.B;
	if (slot < nm) t.main[slot] = (t.main[slot] & ~t.top) | cnt;
	else t.stash[slot - nm] = (t.stash[slot - nm] & ~(unsigned long long) t.top) | cnt;
}

// ------------------------------------------------------------------------------------------------------------------
// p-mer array (reference: bit_vec.h:17-231), u32 words of 16 two-bit fields
// ------------------------------------------------------------------------------------------------------------------
// Sharding: the owner of a p-mer is (x >> (2p - 12)) % world (dna.cpp:658, 845, 2381), i.e. its top 12 bits pick the rank; a
// rank stores the fields of its own top values densely (local top = top / world).  w is THIS rank's shard, peer_w every rank's.
struct SivDev { uint32_t *w; uint32_t key_bits; uint32_t world, rank, top_shift; const uint32_t *peer_w[8]; };

FQSK_HD uint32_t siv_owner(const SivDev &s, uint64_t idx) { return s.world > 1 ? (uint32_t) ((idx >> s.top_shift) % s.world) : 0u; }
// owner's array for the field `idx`; idx becomes the index inside that shard
FQSK_DEV const uint32_t *siv_shard(const SivDev &s, uint64_t &idx) {
	if (s.world <= 1) return s.w;
	const uint64_t top = idx >> s.top_shift;
	const uint32_t owner = (uint32_t) (top % s.world);
	idx = ((top / s.world) << s.top_shift) | (idx & ((1ull << s.top_shift) - 1ull));
	return s.peer_w[owner];
}
FQSK_DEV uint32_t siv_test(const SivDev &s, uint64_t idx) { const uint32_t *w = siv_shard(s, idx); return (__ldg(w + (idx >> 4)) >> (2 * (idx & 15))) & 3; }  // bit_vec.h:69-81
FQSK_DEV void siv_counts(const SivDev &s, uint64_t idx, uint32_t c[4], bool accumulate) {  // bit_vec.h:83-111
	const uint32_t *w = siv_shard(s, idx);
	uint32_t d = __ldg(w + (idx >> 4)) >> (2 * ((idx & 15) & ~3ull));
#pragma unroll
	for (int i = 0; i < 4; ++i) { uint32_t v = (d >> (2 * i)) & 3; c[i] = accumulate ? c[i] + v : v; }
}
FQSK_DEV uint32_t siv_word_sum(uint32_t w) {
	uint32_t x = (w & 0x33333333u) + ((w >> 2) & 0x33333333u);
	x = (x + (x >> 4)) & 0x0F0F0F0Fu;
	return (x * 0x01010101u) >> 24;
}
FQSK_DEV uint64_t siv_prefix_sum(const SivDev &s, uint64_t prefix, uint32_t prefix_bits) {  // bit_vec.h:113-166
	uint32_t sh = s.key_bits - prefix_bits;
	uint64_t start = prefix << sh, n = 1ull << sh;
	if (n == 1) return siv_test(s, start);
	const uint32_t *w = siv_shard(s, start);      // prefixes are at least 12 bits long (prefix_len >= 9 symbols): one owner per range
	if (n < 16) { uint32_t d = __ldg(w + (start >> 4)) >> (2 * (start & 15)); return siv_word_sum(d & ((1u << (2 * n)) - 1)); }
	uint64_t r = 0;
	for (uint64_t wd = start >> 4, e = (start + n) >> 4; wd < e; ++wd) r += siv_word_sum(__ldg(w + wd));
	return r;
}
// bit_vec.h:53-67, order-independent form: saturate at 3, return 1 when the field was zero
FQSK_DEV uint32_t siv_increment(const SivDev &s, uint64_t idx) {   // idx is owned by this rank (routing happened before)
	siv_shard(s, idx);
	uint32_t *p = s.w + (idx >> 4);
	uint32_t sh = 2 * (idx & 15);
	uint32_t old = *((volatile uint32_t *) p);
	for (;;) {
		uint32_t f = (old >> sh) & 3;
		if (f == 3) return 0;
		uint32_t seen = atomicCAS(p, old, old + (1u << sh));
		if (seen == old) return f == 0;
		old = seen;
	}
}

// ------------------------------------------------------------------------------------------------------------------
// intra-segment delta (the reference's thread-local CHT_kmer<uint64_t>, dna.cpp:99-103, 826, 837, 862, 872): every push of
// the segment as a (k-mer, push time) entry of an open-addressing table whose probe run is chosen by the CANONICAL INNER
// CORE of the k-mer (symbols t..k-1-t, min with its reverse complement).  All 4^m front completions, both orientations
// and the 4 next-symbol siblings of a context share that core, so one probe run answers a thread-local find() -- full or
// front-truncated (ht_kmer.h:266-327 needs 4^m probe runs for the same answer).  A lookup at time T sees entries < T.
// ------------------------------------------------------------------------------------------------------------------
struct DeltaDev {
	const unsigned long long *keys; const uint32_t *times;
	uint32_t mask, n;         // slots - 1; number of pushes (0 = empty delta)
	uint32_t k, t;            // k-mer length; symbols trimmed on each side to get the core
	uint32_t exact_limit;     // thr + 1: highest counter value reachable without the thread-local PRNG
	// hot mode only (a k-mer is pushed more than exact_limit times inside the segment): per entry, its push-order rank among
	// equal k-mers, the previous equal entry, and the counter after this push as evaluated with the cinc_lb / cinc_ls stream
	uint32_t *rank_at, *prev_at, *cnt_at;
	// large segments: only the pushes somebody can ask for are inserted.  filter = one bit per hash of a canonical inner core, set
	// for every context the miss list (k_lookup / k_partial) or a repair window (k_walk) looks up; k_delta_build skips a push whose
	// bit is clear.  All k-mers a query can match share its core, so a query whose bit was set before the build sees exactly what
	// the full table would show it; a query whose bit was clear sets it and asks for one more pass (delta_note).
	uint32_t *filter; uint32_t fmask;
};
static const uint32_t DELTA_EMPTY = 0xFFFFFFFFu;

FQSK_HD uint64_t fmix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
FQSK_DEV uint64_t rc_kmer(uint64_t x, uint32_t k) {   // reverse complement of a left-aligned k-mer
	uint64_t y = __brevll(x);                                            // symbols reversed, bits inside each symbol swapped
	y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
	return (~(y << (64 - 2 * k))) & (~0ull << (64 - 2 * k));
}
FQSK_DEV uint64_t delta_slot_of_key(uint64_t x, uint32_t k, uint32_t t, uint32_t mask) {
	uint32_t cl = k - 2 * t;
	uint64_t c1 = (x << (2 * t)) >> (64 - 2 * cl);
	uint64_t c2 = (rc_kmer(x, k) << (2 * t)) >> (64 - 2 * cl);
	return fmix64(c1 < c2 ? c1 : c2) & mask;
}
FQSK_DEV uint32_t delta_fbit_of_key(uint64_t x, uint32_t k, uint32_t t, uint32_t fmask) {      // filter bit of a stored k-mer
	uint32_t cl = k - 2 * t;
	uint64_t c1 = (x << (2 * t)) >> (64 - 2 * cl);
	uint64_t c2 = (rc_kmer(x, k) << (2 * t)) >> (64 - 2 * cl);
	return (uint32_t) (fmix64(c1 < c2 ? c1 : c2) >> 34) & fmask;
}
FQSK_DEV uint32_t delta_fbit_of_query(const DeltaDev &D, const KReg &r, uint32_t cur) {            // ... of a context (cur symbols, placeholder last)
	const uint32_t k = D.k, m = k - cur, cl = k - 2 * D.t;
	uint64_t c1 = (r.dir << (2 * (D.t - m))) >> (64 - 2 * cl);
	uint64_t c2 = (r.rc << (2 * D.t)) >> (64 - 2 * cl);
	return (uint32_t) (fmix64(c1 < c2 ? c1 : c2) >> 34) & D.fmask;
}
// announce a thread-local lookup: sets the filter bit of the context; false when the bit was clear, i.e. the delta built so far
// may lack entries this lookup must see
FQSK_DEV bool delta_note(const DeltaDev &D, const KReg &r, uint32_t cur) {
	if (!D.filter) return true;
	const uint32_t fb = delta_fbit_of_query(D, r, cur), bit = 1u << (fb & 31);
	if (D.filter[fb >> 5] & bit) return true;
	return (atomicOr(D.filter + (fb >> 5), bit) & bit) != 0;
}
// Visits every entry older than T whose k-mer completes the context held in `r` (cur symbols, the last one is the
// placeholder; k - cur leading symbols unknown), exactly as the reference's trial loop would find it: a stored key X counts
// for the trial Y in {X, rc(X)} whose known symbols match and whose normalised form is X (kmer.h:366-385).
// f(Y, time, slot): Y = the trial k-mer (dir form, left-aligned): its last symbol is the next symbol, its first k - cur
// symbols are the front completion.
template <typename F>
FQSK_DEV void delta_scan(const DeltaDev &D, const KReg &r, uint32_t cur, uint32_t T, F &f) {
	if (D.n == 0) return;
	const uint32_t k = D.k, m = k - cur, cl = k - 2 * D.t;
	uint64_t c1 = (r.dir << (2 * (D.t - m))) >> (64 - 2 * cl);
	uint64_t c2 = (r.rc << (2 * D.t)) >> (64 - 2 * cl);
	const uint64_t qd = r.dir >> (2 * m);
	const uint64_t kmask = (cur >= 2 ? ((1ull << (2 * (cur - 1))) - 1ull) : 0ull) << (64 - 2 * k + 2);
	const uint64_t km = kr_kernel_mask(k);
	for (uint64_t slot = fmix64(c1 < c2 ? c1 : c2) & D.mask;; slot = (slot + 1) & D.mask) {
		uint32_t tm = D.times[slot];
		if (tm == DELTA_EMPTY) break;
		if (tm >= T) continue;
		uint64_t X = D.keys[slot];
		uint64_t Xr = rc_kmer(X, k);
		uint64_t kx = X & km, kxr = Xr & km;
		bool pal = X == Xr;
		if (((X ^ qd) & kmask) == 0 && (kx < kxr || pal)) f(X, tm, (uint32_t) slot);
		if (!pal && ((Xr ^ qd) & kmask) == 0 && !(kxr < kx)) f(Xr, tm, (uint32_t) slot);
	}
}
struct DeltaCount {      // plain occurrence counts per next symbol
	uint32_t c[4]; uint32_t lsh;
	FQSK_DEV void operator()(uint64_t Y, uint32_t, uint32_t) { ++c[(Y >> lsh) & 3]; }
};
struct DeltaLatest {     // latest entry per next symbol (full contexts: one k-mer per symbol)
	uint32_t t[4], s[4]; uint32_t lsh;
	FQSK_DEV void operator()(uint64_t Y, uint32_t tm, uint32_t slot) { uint32_t q = (Y >> lsh) & 3; if (s[q] == 0xFFFFFFFFu || tm > t[q]) { t[q] = tm; s[q] = slot; } }
};
FQSK_DEV uint32_t delta_count_at(const DeltaDev &D, uint32_t slot) {   // counter after the push stored at `slot` (hot mode)
	uint32_t rk = D.rank_at[slot];
	return rk < D.exact_limit ? rk + 1 : D.cnt_at[slot];
}
// thread-local find(): true when anything was found.
// Normal mode: counter values that would need the thread-local PRNG stream (cinc_lb / cinc_ls, dna.cpp:164-165) are reported
// through *unsupported (the host then redoes the segment in hot mode).
// Hot mode (D.rank_at != nullptr): full contexts read the evaluated counters; front-truncated contexts with any match are
// left to the ordered evaluator (*need_fold).
FQSK_DEV bool delta_find(const DeltaDev &D, const CIncP &ci, const KReg &r, uint32_t cur, uint32_t T, uint32_t c[4], int *unsupported, int *need_fold = nullptr) {
	c[0] = c[1] = c[2] = c[3] = 0;
	if (D.n == 0) return false;
	const uint32_t lsh = 64 - 2 * D.k;
	if (D.rank_at && cur >= D.k) {
		DeltaLatest L; L.lsh = lsh;
		for (int i = 0; i < 4; ++i) { L.t[i] = 0; L.s[i] = 0xFFFFFFFFu; }
		delta_scan(D, r, cur, T, L);
		for (int i = 0; i < 4; ++i) if (L.s[i] != 0xFFFFFFFFu) c[i] = delta_count_at(D, L.s[i]);
		return (c[0] | c[1] | c[2] | c[3]) != 0;
	}
	DeltaCount C; C.lsh = lsh; C.c[0] = C.c[1] = C.c[2] = C.c[3] = 0;
	delta_scan(D, r, cur, T, C);
	for (int i = 0; i < 4; ++i) c[i] = C.c[i];
	bool any = (c[0] | c[1] | c[2] | c[3]) != 0;
	if (D.rank_at) {   // hot mode, front-truncated context
		if (any && need_fold) *need_fold = 1;
		return any;
	}
	uint32_t lim = cur >= D.k ? D.exact_limit : ci.thr;   // a merge of several completions must stay in the exact range
	for (int i = 0; i < 4; ++i) if (c[i] > lim) { *unsupported = 1; c[i] = lim; }
	return any;
}

}  // namespace fqsk
