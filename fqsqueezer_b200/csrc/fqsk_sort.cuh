// fqsk_sort.cuh -- the engine's own stable radix partition and prefix scans (no library kernels on the path).
//
// What the sync step needs from a sort is narrow: group the pending b-mers of a large row by the BUCKET they live in (27 key
// bits at config-2 size, not the 48 bits of the k-mer) without disturbing push order inside a group, so that one thread can
// walk a bucket's occurrences in the reference's insert order (dna.cpp:2441-2446) against ONE 32-byte sector.  The same
// partition serves the paired-end triples, the owner routing of a sharded sync and the time order of the thread-local
// evaluator's events.
//
// One LSD pass = three launches:
//   k_rdx_hist     per tile of 2048 elements: digit counts -> hist[digit][tile]
//   k_rdx_rowscan  per digit: exclusive prefix over the tiles + the digit's total
//   k_rdx_scatter  per tile: stable ranks (warp match + per-warp private counters), elements staged in shared memory in their
//                  tile-local sorted order and written out run by run (coalesced per digit)
// Stability: a warp owns 256 consecutive elements and visits them in order; warps and tiles are offset in index order.
#pragma once
#include "fqsk_dev.cuh"

namespace fqsk {

static const uint32_t RDX_TILE = 2048, RDX_THREADS = 256, RDX_ROUNDS = RDX_TILE / RDX_THREADS;      // 8 rounds of 32 elements per warp

struct BitsOp {            // digit = bits [shift, shift + NBITS) of the key
	uint32_t shift, mask;
	FQSK_DEV uint32_t operator()(unsigned long long key) const { return (uint32_t) (key >> shift) & mask; }
};
struct BucketOp {          // digit = bits of the table bucket a normalised k-mer lives in (fqsk_dev.cuh: ht_key)
	HtDev t; uint32_t shift, mask;
	FQSK_DEV uint32_t operator()(unsigned long long key) const { return (uint32_t) ((ht_mix(t, ht_kernel(t, key)) >> t.rem_bits) >> shift) & mask; }
};
struct OwnerOp {           // digit = owning rank of a pending k-mer (kind 0: p-mer index, 1: normalised s-/b-mer, 2: pair key); dna.cpp:825, 836, 845, 1076-1081
	uint32_t kind, pshift, world;
	FQSK_DEV uint32_t operator()(unsigned long long key) const {
		return kind == 0 ? (uint32_t) ((key >> pshift) % world) : kind == 1 ? ht_owner(world, key) : (uint32_t) ((fmix64(key) >> 48) % world);
	}
};

template <int NBITS, class Op>
__global__ void __launch_bounds__(RDX_THREADS) k_rdx_hist(const unsigned long long *keys, uint32_t n, uint32_t n_tiles, uint32_t *hist, Op op) { pdl_enter();
	constexpr uint32_t NB = 1u << NBITS;
	__shared__ uint32_t cnt[NB];
	for (uint32_t d = threadIdx.x; d < NB; d += RDX_THREADS) cnt[d] = 0;
	__syncthreads();
	const uint32_t base = blockIdx.x * RDX_TILE;
	for (uint32_t r = 0; r < RDX_ROUNDS; ++r) {
		const uint32_t e = base + r * RDX_THREADS + threadIdx.x;
		if (e < n) atomicAdd(cnt + op(keys[e]), 1u);
	}
	__syncthreads();
	for (uint32_t d = threadIdx.x; d < NB; d += RDX_THREADS) hist[(size_t) d * n_tiles + blockIdx.x] = cnt[d];
}

// one CTA per digit: hist[d][*] -> exclusive prefix over the tiles (in place), totals[d] = the digit's count
__global__ void __launch_bounds__(256) k_rdx_rowscan(uint32_t *hist, uint32_t n_tiles, uint32_t *totals) { pdl_enter();
	__shared__ uint32_t wsum[8];
	__shared__ uint32_t carry;
	uint32_t *row = hist + (size_t) blockIdx.x * n_tiles;
	const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5;
	if (t == 0) carry = 0;
	__syncthreads();
	for (uint32_t base = 0; base < n_tiles; base += 256 * 4) {
		uint32_t v[4], run = 0;
		for (int e = 0; e < 4; ++e) { const uint32_t i = base + t * 4 + e; v[e] = i < n_tiles ? row[i] : 0; }
		for (int e = 0; e < 4; ++e) { const uint32_t x = v[e]; v[e] = run; run += x; }
		uint32_t inc = run;
		for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
		if (lane == 31) wsum[w] = inc;
		__syncthreads();
		uint32_t woff = 0, tot = 0;
		for (uint32_t q = 0; q < 8; ++q) { if (q < w) woff += wsum[q]; tot += wsum[q]; }
		const uint32_t off = carry + woff + inc - run;
		for (int e = 0; e < 4; ++e) { const uint32_t i = base + t * 4 + e; if (i < n_tiles) row[i] = off + v[e]; }
		__syncthreads();
		if (t == 0) carry += tot;
		__syncthreads();
	}
	if (t == 0) totals[blockIdx.x] = carry;
}

// vals_in == nullptr: the value of element e is e itself (first pass of a sort that wants the original index back)
template <int NBITS, class Op>
__global__ void __launch_bounds__(RDX_THREADS) k_rdx_scatter(const unsigned long long *keys_in, const uint32_t *vals_in, unsigned long long *keys_out, uint32_t *vals_out,
                                                             uint32_t n, uint32_t n_tiles, const uint32_t *hist, const uint32_t *totals, Op op) { pdl_enter();
	constexpr uint32_t NB = 1u << NBITS;
	__shared__ uint32_t wh[8][NB];          // per warp: digit counts, then running tile-local positions
	__shared__ uint32_t lds[NB + 1];        // tile-local start of every digit
	__shared__ uint32_t gb[NB];             // global position of the tile's first element of every digit
	__shared__ unsigned long long keyS[RDX_TILE];
	__shared__ uint32_t valS[RDX_TILE];
	__shared__ uint32_t scan_w[8];
	const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5;
	const uint32_t base = blockIdx.x * RDX_TILE + w * (RDX_TILE / 8);      // this warp's 256 consecutive elements
	for (uint32_t d = t; d < 8 * NB; d += RDX_THREADS) (&wh[0][0])[d] = 0;
	__syncthreads();
	unsigned long long key[RDX_ROUNDS]; uint32_t val[RDX_ROUNDS], dig[RDX_ROUNDS];
#pragma unroll
	for (uint32_t r = 0; r < RDX_ROUNDS; ++r) {
		const uint32_t e = base + r * 32 + lane;
		const bool ok = e < n;
		key[r] = ok ? keys_in[e] : 0ull;
		val[r] = ok ? (vals_in ? vals_in[e] : e) : 0u;
		dig[r] = ok ? op(key[r]) : 0xFFFFFFFFu;
		const unsigned m = __match_any_sync(0xffffffffu, dig[r]);
		if (ok && lane == (uint32_t) __ffs(m) - 1) wh[w][dig[r]] += __popc(m);
		__syncwarp();
	}
	__syncthreads();
	// digit totals of the tile -> tile-local starts (exclusive scan over the digits), per-warp starting positions, global bases
	{
		// digit starts over all tiles: exclusive scan of totals[]; every CTA redoes it (NB <= 512 values)
		uint32_t tsum[(NB + RDX_THREADS - 1) / RDX_THREADS], csum[(NB + RDX_THREADS - 1) / RDX_THREADS];
		uint32_t a = 0, b = 0;
		for (uint32_t q = 0; q < (NB + RDX_THREADS - 1) / RDX_THREADS; ++q) {
			const uint32_t d = t * ((NB + RDX_THREADS - 1) / RDX_THREADS) + q;
			uint32_t c = 0;
			if (d < NB) for (uint32_t ww = 0; ww < 8; ++ww) c += wh[ww][d];
			tsum[q] = d < NB ? totals[d] : 0; csum[q] = c;
			a += tsum[q]; b += csum[q];
		}
		uint32_t ia = a, ib = b;
		for (int o = 1; o < 32; o <<= 1) { const uint32_t ya = __shfl_up_sync(0xffffffffu, ia, o), yb = __shfl_up_sync(0xffffffffu, ib, o); if (lane >= (uint32_t) o) { ia += ya; ib += yb; } }
		__shared__ uint32_t wa[8], wb[8];
		if (lane == 31) { wa[w] = ia; wb[w] = ib; }
		__syncthreads();
		uint32_t oa = ia - a, ob = ib - b;
		for (uint32_t q = 0; q < w; ++q) { oa += wa[q]; ob += wb[q]; }
		for (uint32_t q = 0; q < (NB + RDX_THREADS - 1) / RDX_THREADS; ++q) {
			const uint32_t d = t * ((NB + RDX_THREADS - 1) / RDX_THREADS) + q;
			if (d < NB) {
				lds[d] = ob;
				gb[d] = oa + hist[(size_t) d * n_tiles + blockIdx.x];
				uint32_t run = ob;
				for (uint32_t ww = 0; ww < 8; ++ww) { const uint32_t c = wh[ww][d]; wh[ww][d] = run; run += c; }
			}
			oa += tsum[q]; ob += csum[q];
		}
		if (t == RDX_THREADS - 1) lds[NB] = ob;
		(void) scan_w;
	}
	__syncthreads();
	// stable ranks: a warp walks its rounds in order; equal digits inside a round keep lane order
#pragma unroll
	for (uint32_t r = 0; r < RDX_ROUNDS; ++r) {
		const bool ok = dig[r] != 0xFFFFFFFFu;
		const unsigned m = __match_any_sync(0xffffffffu, dig[r]);
		const uint32_t leader = (uint32_t) __ffs(m) - 1;
		uint32_t pos = 0;
		if (ok && lane == leader) { pos = wh[w][dig[r]]; wh[w][dig[r]] = pos + __popc(m); }
		pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(m & ((1u << lane) - 1u));
		if (ok) { keyS[pos] = key[r]; valS[pos] = val[r]; }
		__syncwarp();
	}
	__syncthreads();
	const uint32_t cnt = lds[NB];
	for (uint32_t i = t; i < cnt; i += RDX_THREADS) {
		const unsigned long long k = keyS[i];
		const uint32_t d = op(k);
		const uint32_t g = gb[d] + (i - lds[d]);
		keys_out[g] = k;
		if (vals_out) vals_out[g] = valS[i];
	}
}

// ------------------------------------------------------------------------------------------------------------------
// flags (u8) -> exclusive scan (u32), out[n] = total.  Chained: every CTA scans a tile of 4096 flags, publishes its sum tagged
// with the launch epoch and adds up the sums of the CTAs before it (CTAs are dispatched in index order, so the CTAs a waiting
// one depends on are resident or done).
// ------------------------------------------------------------------------------------------------------------------
static const uint32_t SCAN8_TILE = 4096;
__global__ void __launch_bounds__(256) k_scan_u8(const uint8_t *flag, uint32_t n, uint32_t *out, unsigned long long *partials, uint32_t epoch) { pdl_enter();
	const uint32_t b = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
	const uint32_t base = b * SCAN8_TILE;
	__shared__ uint32_t wsum[8];
	__shared__ uint32_t bprefix;
	const uint32_t r0 = base + t * 16;
	uint32_t v[16];
	if (r0 + 16 <= n && ((reinterpret_cast<uintptr_t>(flag) & 15) == 0)) {
		const uint4 q = *reinterpret_cast<const uint4 *>(flag + r0);
		const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
		for (int e = 0; e < 16; ++e) v[e] = (wd[e >> 2] >> (8 * (e & 3))) & 0xFF;
	} else {
#pragma unroll
		for (int e = 0; e < 16; ++e) v[e] = r0 + e < n ? flag[r0 + e] : 0;
	}
	uint32_t run = 0;
#pragma unroll
	for (int e = 0; e < 16; ++e) { const uint32_t x = v[e]; v[e] = run; run += x; }
	uint32_t inc = run;
	for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t) o) inc += y; }
	if (lane == 31) wsum[w] = inc;
	__syncthreads();
	uint32_t woff = 0, bsum = 0;
#pragma unroll
	for (int q = 0; q < 8; ++q) { if (q < (int) w) woff += wsum[q]; bsum += wsum[q]; }
	if (t == 0) { __threadfence(); atomicExch(partials + b, ((unsigned long long) epoch << 32) | bsum); }
	uint32_t pre = 0;
	for (uint32_t q = t; q < b; q += 256) {
		unsigned long long x;
		do { x = *((volatile unsigned long long *) (partials + q)); } while ((uint32_t) (x >> 32) != epoch);
		pre += (uint32_t) x;
	}
	for (int o = 16; o; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
	__syncthreads();
	if (lane == 0) wsum[w] = pre;
	__syncthreads();
	if (t == 0) { uint32_t x = 0; for (int q = 0; q < 8; ++q) x += wsum[q]; bprefix = x; }
	__syncthreads();
	const uint32_t off = bprefix + woff + inc - run;
#pragma unroll
	for (int e = 0; e < 16; ++e) if (r0 + e < n) out[r0 + e] = off + v[e];
	if (base + SCAN8_TILE >= n && t == 0) out[n] = bprefix + bsum;
}

}  // namespace fqsk
