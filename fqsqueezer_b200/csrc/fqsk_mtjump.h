// fqsk_mtjump.h -- jump-ahead for mt19937 (host side): the polynomials that let G thread blocks extend ONE stream concurrently.
//
// The reference's approximate counters draw from std::mt19937 seeded 5481 (utils.h:256-335); the engine needs that exact
// sequence, tens of millions of outputs per reads_block late in a file.  The block recurrence x[n + 624] = f(x[n], x[n + 1],
// x[n + 397]) has a dependency distance of 227 words, so one CTA tops out near 2 G outputs/s.  mt19937 is linear over GF(2):
// with phi(x) the characteristic polynomial of its transition (degree 19937) and g(x) = x^J mod phi(x),
//     x[J + t] = XOR over { i : g_i = 1 } of x[i + t]        (every bit of the state: t = 1..623, and the top bit of t = 0),
// i.e. the state J steps ahead is a fixed GF(2) combination of the next 19937 + 623 words of the sequence (Haramoto, Matsumoto,
// Nishimura, Panneton, L'Ecuyer: "Efficient jump ahead for F2-linear random number generators", 2008).  k_mt_jump
// (fqsk_kernels.cuh) evaluates that sum for J = j * chunk, j = 1..G-1, and k_mt_extend then runs G chunks side by side.
//
// Nothing here is a magic table: phi is found at run time by Berlekamp-Massey on 2 * 19937 output bits of the generator itself,
// x^J mod phi by square-and-multiply; tests/test_gpu_units.py compares millions of device outputs with the sequential generator.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace fqsk_mtjump {

static const int DEG = 19937;
static const int PW = 624;                  // 32-bit words per polynomial (19968 bits >= DEG + 1)
typedef std::vector<uint64_t> Poly;         // bit i = coefficient of x^i

inline bool pbit(const Poly &p, int i) { return (p[(size_t) i >> 6] >> (i & 63)) & 1ull; }
inline void pflip(Poly &p, int i) { p[(size_t) i >> 6] ^= 1ull << (i & 63); }

// the raw (untempered) word sequence of mt19937 from a given 624-word state
struct RawMt {
	uint32_t st[624]; int at = 0;
	void seed(uint32_t s) { st[0] = s; for (int i = 1; i < 624; ++i) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t) i; at = 0; }
	uint32_t next() {      // x[n + 624], replacing x[n]
		const int i = at, j = (at + 1) % 624, k = (at + 397) % 624;
		const uint32_t y = (st[i] & 0x80000000u) | (st[j] & 0x7fffffffu);
		st[i] = st[k] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
		at = j;
		return st[i];
	}
};

// Berlekamp-Massey over GF(2): connection polynomial C (C_0 = 1) with s[n] = XOR_{i=1..L} C_i s[n - i]
inline Poly berlekamp_massey(const std::vector<uint8_t> &s, int &L_out) {
	const size_t W = (s.size() + 64) / 64 + 1;
	Poly C(W, 0), B(W, 0), T;
	C[0] = B[0] = 1;
	int L = 0, m = 1;
	Poly win(W, 0);      // bit i = s[n - i]
	for (size_t n = 0; n < s.size(); ++n) {
		// shift the window left by one and put s[n] at bit 0
		uint64_t carry = s[n];
		for (size_t w = 0; w < W; ++w) { uint64_t nc = win[w] >> 63; win[w] = (win[w] << 1) | carry; carry = nc; }
		uint64_t d = 0;
		const size_t lw = (size_t) L / 64 + 1;
		for (size_t w = 0; w < lw && w < W; ++w) d ^= C[w] & win[w];
		if (!(__builtin_popcountll(d) & 1)) { ++m; continue; }
		T = C;
		// C ^= B << m
		const size_t ws = (size_t) m >> 6; const int bs = m & 63;
		for (size_t w = W; w-- > ws;) {
			uint64_t v = B[w - ws] << bs;
			if (bs && w > ws) v |= B[w - ws - 1] >> (64 - bs);
			C[w] ^= v;
		}
		if (2 * L <= (int) n) { L = (int) n + 1 - L; B = T; m = 1; } else ++m;
	}
	L_out = L;
	return C;
}

struct Field {
	Poly phi;            // characteristic polynomial, degree DEG
	bool ok = false;
	static const size_t W = (2 * DEG + 64) / 64 + 1;

	void init() {
		RawMt g; g.seed(5481);
		std::vector<uint8_t> s(2 * DEG + 64);
		for (auto &b : s) b = (uint8_t) (g.next() >> 31);      // the top bit of every word is a linear functional of the state
		int L = 0;
		Poly C = berlekamp_massey(s, L);
		if (L != DEG) return;
		phi.assign(W, 0);
		for (int i = 0; i <= DEG; ++i) if (pbit(C, i)) pflip(phi, DEG - i);      // reciprocal: phi(x) = x^L C(1 / x)
		ok = pbit(phi, DEG) && pbit(phi, 0);
	}
	void reduce(Poly &a) const {      // a (degree < 2 DEG) mod phi, in place
		for (int i = 2 * DEG - 1; i >= DEG; --i) {
			if (!pbit(a, i)) continue;
			const int sh = i - DEG; const size_t ws = (size_t) sh >> 6; const int bs = sh & 63;
			const size_t nw = (size_t) DEG / 64 + 1;
			for (size_t w = 0; w < nw; ++w) {
				a[w + ws] ^= phi[w] << bs;
				if (bs) a[w + ws + 1] ^= phi[w] >> (64 - bs);
			}
		}
	}
	Poly mul(const Poly &a, const Poly &b) const {
		Poly r(W + 2, 0);
		const size_t nw = (size_t) DEG / 64 + 1;
		for (int i = 0; i < DEG; ++i) {
			if (!pbit(a, i)) continue;
			const size_t ws = (size_t) i >> 6; const int bs = i & 63;
			for (size_t w = 0; w < nw; ++w) {
				r[w + ws] ^= b[w] << bs;
				if (bs) r[w + ws + 1] ^= b[w] >> (64 - bs);
			}
		}
		reduce(r);
		r.resize(W);
		return r;
	}
	Poly x_pow(uint64_t J) const {      // x^J mod phi
		Poly r(W, 0); r[0] = 1;
		int top = 63; while (top > 0 && !((J >> top) & 1)) --top;
		for (int b = top; b >= 0; --b) {
			r = mul(r, r);
			if ((J >> b) & 1) {      // times x
				uint64_t carry = 0;
				for (size_t w = 0; w < W; ++w) { uint64_t nc = r[w] >> 63; r[w] = (r[w] << 1) | carry; carry = nc; }
				if (pbit(r, DEG)) for (size_t w = 0; w < (size_t) DEG / 64 + 1; ++w) r[w] ^= phi[w];
			}
		}
		return r;
	}
};

// polys[j - 1] = x^(j * chunk) mod phi as PW 32-bit words (bit i of word w = coefficient of x^(32 w + i)), j = 1..n
inline bool jump_polys(uint64_t chunk, int n, std::vector<uint32_t> &out) {
	Field F; F.init();
	if (!F.ok) return false;
	Poly g1 = F.x_pow(chunk), g = g1;
	out.assign((size_t) n * PW, 0);
	for (int j = 1; j <= n; ++j) {
		for (int w = 0; w < PW; ++w) out[(size_t) (j - 1) * PW + w] = (uint32_t) (g[(size_t) w >> 1] >> (32 * (w & 1)));
		if (j < n) g = F.mul(g, g1);
	}
	return true;
}

// host reference of the jump (tests of this header itself): state J steps ahead of `st` by the polynomial g
inline void jump_host(const uint32_t *st, const uint32_t *g, uint32_t *out) {
	std::vector<uint32_t> X(624 + DEG + 8);
	RawMt m; memcpy(m.st, st, sizeof m.st); m.at = 0;
	for (int i = 0; i < 624; ++i) X[i] = st[i];
	for (size_t i = 624; i < X.size(); ++i) X[i] = m.next();
	for (int t = 0; t < 624; ++t) {
		uint32_t acc = 0;
		for (int i = 0; i < DEG; ++i) if ((g[i >> 5] >> (i & 31)) & 1u) acc ^= X[(size_t) i + t];
		out[t] = acc;
	}
}

}  // namespace fqsk_mtjump
