// TEST INFRASTRUCTURE, never shipped: a stand-in for libfqsk.so that serves the segment-level C-ABI (include/fqsk.h) from the
// CPU oracle (oracle/libfqs_oracle.so).  tests/test_live_host.py points host/_bin/fqs-1.1-fqsk at it ($FQSK_LIB) to check the
// HOST half of the integration -- the patched worker loop and compress_suffix of host/build_host.py -- without a GPU: with a
// correct record source the .fqs must be byte-identical to the plain reference's.  The GPU leg of the same test binds the real
// library.  Only the entry points host/fqsk_live.h binds are provided.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fqsk.h"

extern "C" {
void *fqso_create(uint32_t p, uint32_t s, uint32_t b, uint32_t prefix_len, uint32_t mode);
void fqso_destroy(void *h);
void fqso_block_start(void *h);
uint64_t fqso_segment(void *h, const uint8_t *slab, const uint64_t *off, const uint32_t *len, uint32_t n, uint32_t kind, fqsk_base_rec *recs, uint64_t cap, uint8_t *dup);
void fqso_sync(void *h);
void fqso_stats(void *h, uint64_t *o);
}

struct fqsk_handle { void *o; uint64_t n_segments = 0, n_syncs = 0; };

extern "C" {

int fqsk_create(const fqsk_params *p, fqsk_handle **out) {
	if (!p || !out || p->abi_version != FQSK_ABI_VERSION || p->mode != FQSK_MODE_SE_ORIGINAL || p->n_workers != 1) return FQSK_E_INVAL;
	fqsk_handle *h = new fqsk_handle();
	h->o = fqso_create(p->pmer_len, p->smer_len, p->bmer_len, p->prefix_len, 0);
	*out = h;
	return FQSK_OK;
}
void fqsk_destroy(fqsk_handle *h) { if (h) { fqso_destroy(h->o); delete h; } }
const char *fqsk_last_error(fqsk_handle *) { return "mock"; }
int fqsk_block_start(fqsk_handle *h) { fqso_block_start(h->o); return FQSK_OK; }

int fqsk_segment(fqsk_handle *h, const uint8_t *slab, uint64_t, const fqsk_read_desc *reads, uint32_t n_reads, fqsk_base_rec *recs, uint64_t rec_cap, uint64_t *n_recs, uint8_t *dup, uint64_t *) {
	std::vector<uint64_t> off(n_reads + 1);
	std::vector<uint32_t> len(n_reads + 1);
	uint64_t total = 0;
	for (uint32_t i = 0; i < n_reads; ++i) { off[i] = reads[i].dna_off; len[i] = reads[i].dna_len; total += len[i]; }
	std::vector<fqsk_base_rec> tmp(total + 3 * (uint64_t) n_reads + 16);     // the oracle also emits per-read / duplicate markers (pos >= 0xFFFFFFF0)
	std::vector<uint8_t> d(n_reads + 1);
	uint64_t m = fqso_segment(h->o, slab, off.data(), len.data(), n_reads, 0, tmp.data(), tmp.size(), d.data());
	if (m > tmp.size()) return FQSK_E_CAPACITY;
	uint64_t k = 0;
	for (uint64_t i = 0; i < m; ++i) if (tmp[i].pos < 0xFFFFFFF0u) { if (k >= rec_cap) return FQSK_E_CAPACITY; recs[k++] = tmp[i]; }
	*n_recs = k;
	if (dup) memcpy(dup, d.data(), n_reads);
	++h->n_segments;
	return FQSK_OK;
}
int fqsk_sync(fqsk_handle *h) { fqso_sync(h->o); ++h->n_syncs; return FQSK_OK; }
int fqsk_stats_get(fqsk_handle *h, fqsk_stats *out) {
	uint64_t o[16];
	fqso_stats(h->o, o);
	memset(out, 0, sizeof(*out));
	out->siv_no_filled = o[0]; out->siv_no_updates = o[1]; out->n_smers = o[2]; out->n_bmers = o[3];
	out->n_segments = h->n_segments; out->n_syncs = h->n_syncs;
	return FQSK_OK;
}
int fqsk_host_alloc(uint64_t bytes, void **out) { *out = malloc(bytes); return *out ? FQSK_OK : FQSK_E_NOMEM; }
void fqsk_host_free(void *p) { free(p); }

}
