// TEST INFRASTRUCTURE, never shipped: a stand-in for libfqsk.so that serves the segment-level C-ABI (include/fqsk.h) from the
// CPU oracle (oracle/libfqs_oracle.so).  tests/test_live_host.py points host/_bin/fqs-1.1-fqsk at it ($FQSK_LIB) to check the
// HOST half of the integration -- the patched worker loop and compress_suffix of host/build_host.py -- without a GPU: with a
// correct record source the .fqs must be byte-identical to the plain reference's.  The GPU leg of the same test binds the real
// library.  Only the entry points host/fqsk_live.h binds are provided.
#include <cstdint>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <vector>

#include "fqsk.h"

extern "C" {
void *fqso_create(uint32_t p, uint32_t s, uint32_t b, uint32_t prefix_len, uint32_t mode);
void fqso_destroy(void *h);
void fqso_block_start(void *h);
uint64_t fqso_segment(void *h, const uint8_t *slab, const uint64_t *off, const uint32_t *len, uint32_t n, uint32_t kind, fqsk_base_rec *recs, uint64_t cap, uint8_t *dup);
void fqso_sync(void *h);
void fqso_stats(void *h, uint64_t *o);
void *fqso_create_worker(void *first);
void fqso_sync_group(void **hs, uint32_t T);
}

// -t N (host/fqsk_live.h: one engine per worker thread): the workers of one group share the oracle's tables (fqso_create_worker) and meet
// in fqsk_sync_device, where the last one to arrive runs InsertKmersToHT + ClearKmersToHT of all of them (fqso_sync_group)
struct MockGroup {
	std::mutex mu; std::condition_variable cv;
	uint32_t world = 1, arrived = 0, alive = 0; uint64_t generation = 0; void *owner_o = nullptr;
	fqsk_handle *member[8] = {nullptr};
};
static std::mutex g_seg_mu;      // the oracle is sequential code: segments of different workers are served one at a time

struct fqsk_handle {
	void *o = nullptr; uint32_t mode = 0; uint64_t n_segments = 0, n_syncs = 0;
	uint32_t world = 1, rank = 0; MockGroup *grp = nullptr; fqsk_params P{};
	std::vector<uint32_t> s_flag, pair; std::vector<uint64_t> s_dif;      // the segment fqsk_sorted_prefix / fqsk_pair_info describe
	struct Ticket { uint64_t id = 0, n_recs = 0; std::vector<uint32_t> s_flag, pair; std::vector<uint64_t> s_dif; } tk[2];
	uint64_t next_ticket = 1;
};

extern "C" {

int fqsk_create(const fqsk_params *p, fqsk_handle **out) {
	const uint32_t world = p && p->world_size ? p->world_size : 1;
	if (!p || !out || p->abi_version != FQSK_ABI_VERSION || p->mode > FQSK_MODE_PE_SORTED || p->n_workers != world || world > 8 || p->rank >= world) return FQSK_E_INVAL;
	fqsk_handle *h = new fqsk_handle();
	h->mode = p->mode; h->world = world; h->rank = p->rank; h->P = *p;
	if (p->rank == 0) {      // worker 0 owns the shared tables; the others join it in fqsk_shard_attach_local
		h->o = fqso_create(p->pmer_len, p->smer_len, p->bmer_len, p->prefix_len, p->mode);   // dna_mode_t = FQSK_MODE_*
		if (world > 1) { h->grp = new MockGroup(); h->grp->world = world; h->grp->member[0] = h; h->grp->alive = 1; }
	}
	*out = h;
	return FQSK_OK;
}
// worker 0 owns the oracle's shared tables and must be released last; host/fqsk_live.h destroys in rank order, so its release is deferred
void fqsk_destroy(fqsk_handle *h) {
	if (!h) return;
	MockGroup *g = h->grp;
	if (!g) { if (h->o) fqso_destroy(h->o); delete h; return; }
	if (h->rank == 0) g->owner_o = h->o; else if (h->o) fqso_destroy(h->o);
	g->member[h->rank] = nullptr;
	if (--g->alive == 0) { if (g->owner_o) fqso_destroy(g->owner_o); delete g; }
	delete h;
}
int fqsk_shard_attach_local(fqsk_handle *h, fqsk_handle *peer) {
	if (!h || !peer || h->world <= 1 || peer->world != h->world) return FQSK_E_INVAL;
	if (h->rank != 0 && peer->rank == 0 && !h->o) {
		h->o = fqso_create_worker(peer->o);
		h->grp = peer->grp;
		h->grp->member[h->rank] = h; ++h->grp->alive;
	}
	return FQSK_OK;
}
int fqsk_shard_export(fqsk_handle *h, fqsk_shard_desc *out) { if (!h || !out) return FQSK_E_INVAL; memset(out, 0, sizeof *out); out->rank = h->rank; out->world_size = h->world; return FQSK_OK; }
int fqsk_sync_device(fqsk_handle *h) {
	if (!h || h->world <= 1 || !h->grp) return FQSK_E_INVAL;
	MockGroup *g = h->grp;
	std::unique_lock<std::mutex> lk(g->mu);
	const uint64_t gen = g->generation;
	++h->n_syncs;
	if (++g->arrived == g->world) {
		void *hs[8];
		for (uint32_t i = 0; i < g->world; ++i) { if (!g->member[i] || !g->member[i]->o) return FQSK_E_INVAL; hs[i] = g->member[i]->o; }
		fqso_sync_group(hs, g->world);
		g->arrived = 0; ++g->generation;
		g->cv.notify_all();
	} else g->cv.wait(lk, [&] { return g->generation != gen; });
	return FQSK_OK;
}
const char *fqsk_last_error(fqsk_handle *) { return "mock"; }
int fqsk_block_start(fqsk_handle *h) { if (!h->o) return FQSK_E_INVAL; fqso_block_start(h->o); return FQSK_OK; }

int fqsk_segment(fqsk_handle *h, const uint8_t *slab, uint64_t, const fqsk_read_desc *reads, uint32_t n_reads, fqsk_base_rec *recs, uint64_t rec_cap, uint64_t *n_recs, uint8_t *dup, uint64_t *) {
	std::vector<uint64_t> off(n_reads + 1);
	std::vector<uint32_t> len(n_reads + 1);
	uint64_t total = 0;
	for (uint32_t i = 0; i < n_reads; ++i) { off[i] = reads[i].dna_off; len[i] = reads[i].dna_len; total += len[i]; }
	std::vector<fqsk_base_rec> tmp(total + 3 * (uint64_t) n_reads + 16);     // the oracle also emits per-read / duplicate markers (pos >= 0xFFFFFFF0)
	std::vector<uint8_t> d(n_reads + 1);
	const uint32_t kind = h->mode == FQSK_MODE_SE_SORTED ? 2 : (h->mode == FQSK_MODE_PE_ORIGINAL || h->mode == FQSK_MODE_PE_SORTED) ? 3 : 0;     // oracle: 0 CompressDirect, 2 CompressSorted, 3 CompressPE
	if (!h->o) return FQSK_E_INVAL;
	uint64_t m;
	{ std::lock_guard<std::mutex> lk(g_seg_mu); m = fqso_segment(h->o, slab, off.data(), len.data(), n_reads, kind, tmp.data(), tmp.size(), d.data()); }
	if (m > tmp.size()) return FQSK_E_CAPACITY;
	h->s_flag.assign(n_reads + 1, 0); h->s_dif.assign(n_reads + 1, 0); h->pair.assign(3 * (n_reads / 2) + 3, 0);
	uint64_t k = 0, rd = 0, pr = 0;       // rd: reads started so far (one 0xFFFFFFFF marker each; a with-minimizer mate 2 emits kind 1), pr: pairs decided
	for (uint64_t i = 0; i < m; ++i) {
		const fqsk_base_rec &r = tmp[i];
		if (r.pos < 0xFFFFFFF0u) { if (k >= rec_cap) return FQSK_E_CAPACITY; recs[k++] = r; }
		else if (r.pos == 0xFFFFFFFFu) ++rd;
		else if (r.pos == 0xFFFFFFFCu && rd >= 1 && rd <= n_reads) { h->s_flag[rd - 1] = r.counts[0]; h->s_dif[rd - 1] = r.counts[1] | ((uint64_t) r.counts[2] << 32); }
		else if (r.pos == 0xFFFFFFFBu && pr < n_reads / 2) { for (int q = 0; q < 3; ++q) h->pair[3 * pr + q] = r.counts[q]; ++pr; }
	}
	*n_recs = k;
	if (dup) memcpy(dup, d.data(), n_reads);
	++h->n_segments;
	return FQSK_OK;
}
int fqsk_sync(fqsk_handle *h);
// fqsk_submit = segment + sync, served synchronously (the oracle has no stream to overlap with); the per-read extras of a ticket
// become visible to fqsk_sorted_prefix / fqsk_pair_info when it is collected, as in the real library
int fqsk_submit(fqsk_handle *h, const uint8_t *slab, uint64_t slab_size, const fqsk_read_desc *reads, uint32_t n_reads, fqsk_base_rec *recs, uint64_t rec_cap, uint8_t *dup, uint64_t *rec_off,
                uint64_t *ticket) {
	fqsk_handle::Ticket &T = h->tk[h->next_ticket & 1];
	std::vector<uint32_t> kf = h->s_flag, kp = h->pair; std::vector<uint64_t> kd = h->s_dif;
	int rc = fqsk_segment(h, slab, slab_size, reads, n_reads, recs, rec_cap, &T.n_recs, dup, rec_off);
	if (rc != FQSK_OK) return rc;
	T.s_flag.swap(h->s_flag); T.s_dif.swap(h->s_dif); T.pair.swap(h->pair);
	h->s_flag = kf; h->s_dif = kd; h->pair = kp;
	fqsk_sync(h);
	T.id = h->next_ticket++;
	*ticket = T.id;
	return FQSK_OK;
}
int fqsk_collect(fqsk_handle *h, uint64_t ticket, uint64_t *n_recs) {
	for (auto &T : h->tk) if (T.id == ticket) { h->s_flag = T.s_flag; h->s_dif = T.s_dif; h->pair = T.pair; if (n_recs) *n_recs = T.n_recs; T.id = 0; return FQSK_OK; }
	return FQSK_E_INVAL;
}
int fqsk_sorted_prefix(fqsk_handle *h, uint32_t *flag, uint64_t *dif, uint32_t n_reads) {
	for (uint32_t i = 0; i < n_reads; ++i) { flag[i] = h->s_flag[i]; dif[i] = h->s_dif[i]; }
	return FQSK_OK;
}
int fqsk_pair_info(fqsk_handle *h, uint32_t *info, uint32_t n_pairs) {
	for (uint32_t i = 0; i < 3 * n_pairs; ++i) info[i] = h->pair[i];
	return FQSK_OK;
}
int fqsk_sync(fqsk_handle *h) { fqso_sync(h->o); ++h->n_syncs; return FQSK_OK; }
int fqsk_stats_get(fqsk_handle *h, fqsk_stats *out) {
	uint64_t o[32];
	fqso_stats(h->o, o);
	memset(out, 0, sizeof(*out));
	out->siv_no_filled = o[0]; out->siv_no_updates = o[1]; out->n_smers = o[2]; out->n_bmers = o[3];
	out->n_segments = h->n_segments; out->n_syncs = h->n_syncs;
	return FQSK_OK;
}
int fqsk_host_alloc(uint64_t bytes, void **out) { *out = malloc(bytes); return *out ? FQSK_OK : FQSK_E_NOMEM; }
void fqsk_host_free(void *p) { free(p); }

}

// sorted-order front end: the oracle's restatement of the comparator of CSortedFASTQFile::sort_reads (io.h:499-528)
extern "C" void fqso_sort_ranks(const uint8_t *slab, const uint64_t *off, const uint32_t *len, uint32_t n, uint32_t *rank);
extern "C" int fqsk_sort_ranks(int, const uint8_t *slab, uint64_t, const fqsk_read_desc *reads, uint32_t n_reads, uint32_t *rank) {
	std::vector<uint64_t> off(n_reads + 1);
	std::vector<uint32_t> len(n_reads + 1);
	for (uint32_t i = 0; i < n_reads; ++i) { off[i] = reads[i].dna_off; len[i] = reads[i].dna_len; }
	fqso_sort_ranks(slab, off.data(), len.data(), n_reads, rank);
	return FQSK_OK;
}
