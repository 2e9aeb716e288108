/* fqsk.h -- C-ABI of the B200 k-mer statistics engine for FQSqueezer (libfqsk.so).
 *
 * Drop-in boundary for ONE path of refresh-bio/fqsqueezer 1.1: the k-mer statistics engine behind
 * CDNACompressor (dna.cpp) -- CKmer (kmer.h), CHT_kmer<T> (ht_kmer.h), TSmallIntVector<2> (bit_vec.h),
 * CCounterIncrementer (utils.h:256-335).  The reference calls those classes ~13 times per base; a per-call
 * mirror cannot be fast, so the boundary is lifted to what CDNACompressor::compress_suffix (dna.cpp:674-877)
 * needs per sync segment.  Everything below is plain C: pointers and sizes only.
 *
 * All `reference:` citations are relative to /root/reference/fqs/.
 * Every function returns 0 on success or a negative FQSK_E_* code (one positive, non-error code: FQSK_RESHARD); nothing throws across the ABI.
 * There is NO CPU fallback: without a CUDA device fqsk_create fails with FQSK_E_NO_DEVICE.
 */
#ifndef FQSK_H
#define FQSK_H

#include <stdint.h>

#include "fqsk_ctx.h"   /* fqsk_ctx_rec: the 16-byte per-base record of fqsk_submit_ctx, and the functions that build / expand it */

#ifdef __cplusplus
extern "C" {
#endif

#define FQSK_ABI_VERSION 1

enum {
	FQSK_OK = 0,
	FQSK_RESHARD = 1,         /* not an error -- sharded engines only: the sync is complete and table shards must double before the next segment (see fqsk_sync_finish) */
	FQSK_E_INVAL = -1,        /* bad argument */
	FQSK_E_NO_DEVICE = -2,    /* no usable CUDA device */
	FQSK_E_CUDA = -3,         /* CUDA runtime error, see fqsk_last_error() */
	FQSK_E_NOMEM = -4,        /* device allocation failed */
	FQSK_E_CAPACITY = -5,     /* caller's output buffer too small */
	FQSK_E_UNSUPPORTED = -6,  /* input needs a path that is not implemented yet (reported, never silently approximated) */
	FQSK_E_NO_CONVERGE = -7   /* ordered-replay fix point did not settle within the iteration limit */
};

/* dna_mode_t, reference: params.h:18 */
enum { FQSK_MODE_SE_ORIGINAL = 0, FQSK_MODE_SE_SORTED = 1, FQSK_MODE_PE_ORIGINAL = 2, FQSK_MODE_PE_SORTED = 3 };
/* counts_level_t, reference: defs.h:46 */
enum { FQSK_LEVEL_NONE = 0, FQSK_LEVEL_PMER = 1, FQSK_LEVEL_SMER = 2, FQSK_LEVEL_BMER = 3, FQSK_LEVEL_MIXED = 4, FQSK_LEVEL_BMER_UNC = 5 };
/* tables, for fqsk_dump / fqsk_ht_* */
enum { FQSK_TABLE_SIV = 0, FQSK_TABLE_SMER = 1, FQSK_TABLE_BMER = 2, FQSK_TABLE_PAIR = 3 };

/* Replaces the constructor arguments of the reference's tables (application.cpp:86-91, ht_kmer.h:366-399,
 * bit_vec.h:29-40) and the k-mer lengths chosen by CParams::adjust_kmer_sizes (params.h:131-155). */
typedef struct fqsk_params {
	uint32_t abi_version;        /* FQSK_ABI_VERSION */
	uint32_t pmer_len, smer_len, bmer_len, prefix_len;
	uint32_t smer_counter_bits;  /* 12, defs.h:26 */
	uint32_t bmer_counter_bits;  /* 6,  defs.h:27 */
	uint32_t mode;               /* FQSK_MODE_* */
	uint32_t n_workers;          /* reference -t: 1, or world_size (one reference worker thread per GPU; bit-exact with `fqs-1.1 -t N`) */
	int32_t device;              /* CUDA ordinal */
	uint32_t bmer_log2_buckets;  /* 0 = choose from expected_kmers; 8 slots of 4 bytes per bucket */
	uint32_t smer_log2_buckets;
	uint64_t expected_kmers;     /* hint for initial table sizes (distinct b-mers); tables grow when half full */
	/* hash sharding across the GPUs of one box, at most 8 ranks: this handle is reference worker `rank` of
	 * `world_size` and owns the k-mers whose owner key == rank -- ((x >> 46) & 0x3fff) % world_size for s-/b-mers
	 * (dna.cpp:825, 836), (x >> (2p - 12)) % world_size for p-mers (dna.cpp:845) -- exactly the rows [*][rank] of the
	 * reference's X_to_add exchange matrices (application.h:56-59) */
	uint32_t world_size, rank;
	uint32_t max_iterations;     /* fix-point limit per segment, 0 = default (16) */
	uint32_t flags;              /* FQSK_F_* */
	uint32_t reserve_reads;      /* optional: largest segment the caller will submit (reads / DNA bytes); scratch is allocated once */
	uint32_t reserve_bytes;
	uint32_t pair_log2_slots;    /* paired-end: initial size of the pair table (0 = default); it doubles when half full (sharded: all shards together, FQSK_RESHARD) */
	uint32_t test_hooks;         /* only with FQSK_F_TEST_HOOKS, else must be 0: bits 0-15 = N -> every N-th segment's first-pass verdict is forced to
	                              * "not settled"; bits 16-31 = M -> every M-th segment is evaluated a second time from scratch (fault injection) */
} fqsk_params;

#define FQSK_F_PROFILE 1u        /* record CUDA-event timings per internal phase (fqsk_profile) */
#define FQSK_F_TRACE_ALLOC 2u    /* print every device allocation of the segment scratch to stderr */
#define FQSK_F_TRACE_LAUNCH 8u   /* debugging: synchronise the device after every kernel launch and print its source line to stderr */
#define FQSK_F_SERIAL 16u        /* debugging / measurement: every kernel on the engine's one stream (no fork / join over side streams) */
#define FQSK_F_TEST_HOOKS 4u     /* honour fqsk_params.test_hooks (tests of the recovery paths; never set by a host) */
#define FQSK_F_TEST_CROWD 32u    /* with FQSK_F_TEST_HOOKS: the s-mer / b-mer tables double at 1/64 of the usual load (growth paths on small fixtures) */

typedef struct fqsk_handle fqsk_handle;

/* One read of a reads_block slab: replaces read_desc_t::dna / read_len() (defs.h:58-85). */
typedef struct fqsk_read_desc {
	uint64_t dna_off;            /* byte offset of the first DNA symbol inside the slab */
	uint32_t dna_len;            /* number of DNA symbols */
	uint32_t flags;              /* reserved (PE: bit0 = second mate) */
} fqsk_read_desc;

/* One record per coded suffix base, in the order compress_suffix visits them: exactly the values the host-side
 * context model consumes at dna.cpp:737-760 (counts, level after the bmer_unc->bmer and rough->pmer rewrites at
 * 697-735, the rough flag, and cor_pos as used by the cor_zone formula at 741).  28 bytes, little endian. */
typedef struct fqsk_base_rec {
	uint32_t pos;
	uint32_t counts[4];
	uint32_t cor_pos;
	uint8_t level;
	uint8_t rough;
	uint16_t pad;
} fqsk_base_rec;

typedef struct fqsk_stats {
	uint64_t siv_no_filled, siv_no_updates;  /* bit_vec.h:25-26, 204-220: feed avg_filling_factor (dna.cpp:376) */
	uint64_t n_smers, n_bmers;               /* ht_kmer.h:46, 531-534 */
	uint64_t draws[4];                       /* mt19937 outputs consumed so far: cinc_b, cinc_s, cinc_lb, cinc_ls (dna.cpp:162-165) */
	uint64_t n_segments, n_syncs, n_replays; /* engine counters: replays / segments = fix-point cost */
	uint64_t n_bases, n_reads;
	uint64_t kernel_launches;                /* launches of this library's own kernels */
	uint64_t bmer_buckets, smer_buckets, bmer_stash_used, smer_stash_used;
	uint64_t n_hot_segments;                 /* segments redone with the ordered thread-local evaluator (cinc_lb / cinc_ls in use) */
	uint64_t n_filtered_segments;            /* large segments whose thread-local delta held only the pushes a lookup could ask for */
	uint64_t n_looks, look_wait_ns;          /* host looks at the device and the time the host spent WAITING in them: api time minus this = the
	                                          * host's own work (launching); a job whose wait time is near zero is bound by its host, not the GPU */
	uint64_t api_ns;                         /* time inside the segment / sync / submit / collect calls */
	uint64_t n_table_growths;                /* doublings of the s-mer / b-mer / pair table (ht_kmer.h:88-112 restruct) so far */
} fqsk_stats;

/* Named phases of fqsk_profile(): device milliseconds accumulated since create (only with FQSK_F_PROFILE). */
enum { FQSK_PH_PREP = 0, FQSK_PH_LOOKUP, FQSK_PH_PARTIAL, FQSK_PH_WALK, FQSK_PH_COMPACT, FQSK_PH_SORT, FQSK_PH_LOCAL, FQSK_PH_ROUGH, FQSK_PH_FOLD,
       FQSK_PH_SYNC_LOCATE, FQSK_PH_SYNC_APPLY, FQSK_PH_SYNC_SIV, FQSK_PH_MT, FQSK_PH_COUNT };

int fqsk_create(const fqsk_params *p, fqsk_handle **out);
void fqsk_destroy(fqsk_handle *h);
const char *fqsk_last_error(fqsk_handle *h);   /* also valid with h == NULL after a failed create */

/* application.cpp:624 (ResetReadPrev at the start of every reads_block). */
int fqsk_block_start(fqsk_handle *h);

/* Replaces, for the reads [0, n_reads) of one sync segment (application.cpp:630-655), every k-mer-engine call made by
 * CDNACompressor::CompressDirect / CompressSorted (dna.cpp:1517-1556, 1716-1754): register updates, find_counts,
 * rough searches, repairs, pushes to the to_add rows and to the intra-segment delta tables.
 *   slab/reads : the CReadsBlock slab (reads_block.h:18-215) and the reads of this segment (host memory)
 *   recs       : out, capacity rec_cap records; *n_recs = records written (sum of coded suffix bases)
 *   dup        : out, n_reads bytes, 1 = duplicate of the previous read (coded as a flag, no k-mer work; dna.cpp:1523-1533)
 *   rec_off    : out (optional, may be NULL), n_reads + 1 entries: first record of every read
 * Side effect: the segment's pending table updates are kept on the device until fqsk_sync. */
int fqsk_segment(fqsk_handle *h, const uint8_t *slab, uint64_t slab_size, const fqsk_read_desc *reads, uint32_t n_reads,
                 fqsk_base_rec *recs, uint64_t rec_cap, uint64_t *n_recs, uint8_t *dup, uint64_t *rec_off);

/* Same, with the reads already resident in HBM (bench `value`): d_dna is a device pointer to the concatenated DNA bytes
 * (ASCII), d_off/d_len device arrays (u64 / u32); records stay on the device in the handle (fqsk_device_recs).  The engine reads the
 * three arrays on its own non-blocking stream: whatever filled them (a copy, a kernel on another stream) must have completed before
 * the call, and they must stay untouched until the fqsk_sync that follows. */
int fqsk_segment_device(fqsk_handle *h, const uint8_t *d_dna, uint64_t dna_bytes, const uint64_t *d_off, const uint32_t *d_len,
                        uint32_t n_reads, uint64_t *n_recs);
/* Optional hint between fqsk_segment_device and its fqsk_sync: the arguments the NEXT fqsk_segment_device call will have.  What depends on
 * the reads alone (duplicate flags -- dna.cpp:1521-1533 --, letter totals -- 2047-2057 --, record offsets) is then computed next to the
 * segment in flight instead of at the head of the next one.  Same contract for the arrays as fqsk_segment_device; a hint that is not
 * followed by the matching call is dropped.  The announced segment belongs to the same reads_block (no fqsk_block_start in between).
 * Ignored for paired-end modes. */
int fqsk_announce_device(fqsk_handle *h, const uint8_t *d_dna, uint64_t dna_bytes, const uint64_t *d_off, const uint32_t *d_len, uint32_t n_reads);
int fqsk_device_recs(fqsk_handle *h, const fqsk_base_rec **d_recs, uint64_t *n_recs);
/* Order-sensitive checksum of the last segment's records, computed on the device (no record leaves HBM): the sum over the records i of
 * fmix64(a ^ fmix64(b ^ fmix64(c ^ fmix64(d + i)))) with a, b, c = the first three little-endian 64-bit words of record i, d = its last
 * 32 bits and fmix64 = murmur3's finaliser.  bench.py compares it with the same sum over the oracle's records (BASELINE.md section 3:
 * "count vectors verified on device").  Replaces nothing in the reference. */
int fqsk_recs_checksum(fqsk_handle *h, uint64_t *sum, uint64_t *n_recs);
/* Sorted-order modes: what compress_prefix_sorted (dna.cpp:589-605) codes per read of the last segment -- flag = siv.test(p-mer)
 * or 4 when the p-mer equals the previous read's, dif = number of p-mers with that flag between the previous and this p-mer.
 * Paired end in sorted order (FQSK_MODE_PE_SORTED): only first mates go through CompressSorted (dna.cpp:1793-1796); the entries of
 * second mates are 0.  "The last segment" is the one of the most recent fqsk_segment / fqsk_segment_device call or -- with the
 * asynchronous API -- of the most recent fqsk_collect (fqsk_pair_info likewise). */
int fqsk_sorted_prefix(fqsk_handle *h, uint32_t *flag, uint64_t *dif, uint32_t n_reads);

/* Paired-end (FQSK_MODE_PE_ORIGINAL, FQSK_MODE_PE_SORTED): fqsk_segment / fqsk_segment_device take the pairs interleaved (mate 1, mate 2, ...), an even
 * number of reads; the records of a pair come in the order CompressPE codes them (dna.cpp:1790-1880): mate 1, then mate 2 --
 * either whole, or from the shared minimizer to the end followed by the reverse complement of the part left of it
 * (CompressDirectWithMinim, dna.cpp:1559-1638; `pos` of a record is the index inside the text compress_suffix was given).
 * fqsk_pair_info returns what CompressPE codes per pair of the last segment, 3 words each: [0] a candidate list exists
 * (find_minim_cand, dna.cpp:1757-1787), [1] minim2_id (0..14, 15 = no usable candidate; 0 when [0] == 0), [2] minim2_pos (0 unless
 * [1] < 15).  The 14 pushes of append_pe_mers3 (dna.cpp:1058-1136) go to the pair table at the next fqsk_sync. */
int fqsk_pair_info(fqsk_handle *h, uint32_t *info, uint32_t n_pairs);

/* Replaces CDNACompressor::InsertKmersToHT + ClearKmersToHT (dna.cpp:2393-2488) and the three barriers around them
 * (application.cpp:645-654): p-mers, then s-mers, then b-mers, in push order, with the reference's PRNG draw order. */
int fqsk_sync(fqsk_handle *h);

/* Asynchronous, double-buffered form of fqsk_segment + fqsk_sync -- the pair the reference's worker loop executes for every sync
 * segment (application.cpp:630-655).  fqsk_submit stages the reads, enqueues H2D, the segment, the copy of its records into the
 * caller's buffers (page-locked memory from fqsk_host_alloc, on a second CUDA stream) and the sync, and returns a ticket;
 * fqsk_collect blocks until that segment's records, duplicate flags and record offsets are in the caller's buffers.  Up to two
 * tickets may be open: submit segment n + 1, then collect n and hand it to the host-side coder while the GPU works on n + 1 and
 * the records of n + 1 travel behind it.  rec_cap must cover sum(max(dna_len - first coded position, 0)) of the segment (records
 * beyond *n_recs are undefined).  Buffers of a ticket belong to the library until it is collected; fqsk_segment / fqsk_sync /
 * fqsk_dump must not be called while a ticket is open. */
int fqsk_submit(fqsk_handle *h, const uint8_t *slab, uint64_t slab_size, const fqsk_read_desc *reads, uint32_t n_reads,
                fqsk_base_rec *recs, uint64_t rec_cap, uint8_t *dup, uint64_t *rec_off, uint64_t *ticket);
int fqsk_collect(fqsk_handle *h, uint64_t ticket, uint64_t *n_recs);
/* One reads_block through the asynchronous pair, i.e. the reference's worker loop (application.cpp:617-662) with the k-mer engine one segment
 * ahead of the consumer: fqsk_block_start, then fqsk_submit of segment k + 1 before fqsk_collect of segment k.  Segment k covers reads
 * [seg_end[k - 1], seg_end[k]) (seg_end[-1] = 0; the segment after the last sync of a block may be empty).  recs: page-locked (fqsk_host_alloc),
 * rec_cap >= sum of dna_len; the records of segment k start at recs[seg_rec_off[k]] and number seg_n_recs[k].  dup: n_reads bytes or NULL. */
int fqsk_block_host(fqsk_handle *h, const uint8_t *slab, uint64_t slab_size, const fqsk_read_desc *reads, uint32_t n_reads, const uint32_t *seg_end, uint32_t n_segs,
                    fqsk_base_rec *recs, uint64_t rec_cap, uint8_t *dup, uint64_t *seg_rec_off, uint64_t *seg_n_recs);
/* fqsk_block_host for a run of consecutive reads_blocks: the last segment of the block is left in flight -- *carry_ticket names it on
 * return -- and is collected by the next call right after that call's first submit (its record count goes to *carry_n_recs), so the engine
 * stays one segment ahead of the consumer across block boundaries too (draining at every block end would leave the GPU idle while the
 * last 51 000-read segment's records travel).  In: *carry_ticket = the ticket the previous call left, 0 on the first block; after the
 * last block the caller collects *carry_ticket with fqsk_collect.  seg_n_recs[n_segs - 1] is NOT written (it is the next call's
 * *carry_n_recs).  Exactly one of recs / ctx is given: 28-byte count records (fqsk_submit) or 16-byte context records (fqsk_submit_ctx).
 * The buffers of a block (recs / ctx, dup) belong to the library until the carried ticket is collected: alternate between two sets. */
int fqsk_block_stream(fqsk_handle *h, const uint8_t *slab, uint64_t slab_size, const fqsk_read_desc *reads, uint32_t n_reads, const uint32_t *seg_end, uint32_t n_segs,
                      fqsk_base_rec *recs, fqsk_ctx_rec *ctx, uint64_t rec_cap, uint8_t *dup, uint64_t *seg_rec_off, uint64_t *seg_n_recs,
                      uint64_t *carry_ticket, uint64_t *carry_n_recs);
/* fqsk_submit with the context ids of the DNA stream built on the device: instead of fqsk_base_rec the
 * caller receives one fqsk_ctx_rec (16 bytes, include/fqsk_ctx.h) per coded base -- what cor_zone (dna.cpp:739-744),
 * CCodeContext::determine_ctx_codes (code_ctx.cpp:257-324), rank (dna.cpp:177-193) and update_ctx_r_sym (dna.cpp:664-671) compute on
 * the host in the reference: fqsk_ctx_expand() gives the 7 context ids, fqsk_ctx_rsym() the rank to code, fqsk_ctx_coded() == 0 means
 * plain-letter coding (dna.cpp:776-801).  Collected with fqsk_collect like any ticket.  Reads of FQSK_CTX_MAX_READ symbols or more
 * are refused (FQSK_E_UNSUPPORTED): their positions would carry out of the 14-bit position field. */
int fqsk_submit_ctx(fqsk_handle *h, const uint8_t *slab, uint64_t slab_size, const fqsk_read_desc *reads, uint32_t n_reads,
                    fqsk_ctx_rec *ctx, uint64_t rec_cap, uint8_t *dup, uint64_t *rec_off, uint64_t *ticket);

/* Sorted-order front end.  Replaces the comparator of CSortedFASTQFile::sort_reads (io.h:499-528: symbols with
 * N read as T over the shorter length, then the shorter read first, then the raw bytes): rank[i] is an integer per read that is
 * order-isomorphic to it -- rank[x] < rank[y] exactly when the comparator puts x before y, equal ranks for reads it calls equivalent.
 * The host keeps its own std::sort call and compares ranks: the same comparison outcomes, hence the same order as the reference, ties of
 * the unstable sort included.  slab: the bin file (or any buffer the descriptors point into); no engine handle is needed; `device` is the
 * CUDA device to use.  More than 65 536 reads sharing their first 32 symbols: FQSK_E_UNSUPPORTED (message via fqsk_last_error(NULL)). */
int fqsk_sort_ranks(int device, const uint8_t *slab, uint64_t slab_size, const fqsk_read_desc *reads, uint32_t n_reads, uint32_t *rank);

/* ---- sharded operation (world_size > 1): one handle per GPU / process, reference `-t world_size` semantics --------------------
 * Replaces the shared-memory coupling of the reference's worker threads: global tables read by everybody between barriers
 * (application.cpp:645-654), X_to_add[src][dst] exchange matrices + InsertKmersToHT on the owner (dna.cpp:2393-2472).
 * Set-up: every rank calls fqsk_shard_export, the descriptors are exchanged by the caller (any transport) and every rank calls
 * fqsk_shard_attach for each peer: tables and inboxes become NVLink peer mappings (CUDA IPC).
 * Per sync, instead of fqsk_sync:  fqsk_sync_route  ->  fqsk_sync_apply  -> ALL-REDUCE(sum) of (fresh, updates) -> fqsk_sync_finish.
 * The reference's first barrier ("every row is filled") is a sequence number the routing step posts in the owners' inboxes and the
 * apply step waits for ON THE DEVICE; the all-reduce -- the second barrier: every owner has inserted before anybody looks up again --
 * is the caller's (NCCL in fqsqueezer_b200/sharded.py); the payload itself moves inside fqsk_sync_route as peer stores into the
 * owners' inboxes.  Table growth (CHT_kmer::restruct, ht_kmer.h:88-112, ht_kmer.cpp:50-75) is coordinated: all shards of a table have the
 * same geometry, so when ONE rank's shard of the s-mer, b-mer or pair table passes half full at a sync, fqsk_sync_finish / fqsk_sync_device
 * return FQSK_RESHARD on EVERY rank (the sync itself is complete, peer mappings of the tables concerned are already closed); the caller then
 * (1) synchronises all ranks -- nobody may free a table a peer still maps --, (2) calls fqsk_shard_export on every rank (the doubling
 * happens here: dump, 2x buckets / slots, re-insert), (3) exchanges the descriptors and calls fqsk_shard_attach for every peer, exactly as
 * at set-up.  Paired-end (FQSK_MODE_PE_ORIGINAL): the pair table is
 * sharded by (fmix64(key) >> 48) % world_size (ht_kmer.h:599-602, dna.cpp:1076-1081) and the distinct (key, value, weight)
 * triples of a segment travel with the same exchange step. */
typedef struct fqsk_shard_desc {
	uint32_t rank, world_size;
	uint32_t geometry[6];        /* b: log2 buckets, log2 stash; s: same; p-mer key bits; log2 pair-table slots (0: no pair table) -- must agree on all ranks */
	uint64_t inbox_cap;
	uint8_t ipc[8][64];          /* cudaIpcMemHandle_t of: b main, b stash, s main, s stash, p-mer shard, inbox, pair keys, pair values (paired-end only) */
} fqsk_shard_desc;
int fqsk_shard_export(fqsk_handle *h, fqsk_shard_desc *out);
int fqsk_shard_attach(fqsk_handle *h, const fqsk_shard_desc *peer);
/* fqsk_shard_attach for two handles that live in ONE process (one worker thread per GPU, as the reference's -t N worker threads of
 * application.cpp:575-671 do): no descriptor, no IPC -- peer access is enabled between the two devices and the peer's shards are read
 * directly.  Call for every ordered pair of handles after all of them are created (and again, after a barrier and every handle's
 * fqsk_shard_export, when a sync returned FQSK_RESHARD). */
int fqsk_shard_attach_local(fqsk_handle *h, fqsk_handle *peer);
/* Routes this worker's pending rows (p, s, b) to their owners: rows [rank][*] of the exchange matrices, written into the owners'
 * inboxes.  Call on every rank, then synchronise all ranks (barrier). */
int fqsk_sync_route(fqsk_handle *h);
/* Owner side of InsertKmersToHT: applies the rows [*][rank] in source order (p-mers, s-mers, b-mers; this rank's cinc_s / cinc_b).
 * fresh = p-mer fields that became non-zero here, updates = p-mers applied here + this worker's hidden updates (dna.cpp:2416-2418). */
int fqsk_sync_apply(fqsk_handle *h, uint64_t *fresh, uint64_t *updates);
/* Three-step form only, between fqsk_sync_apply and fqsk_sync_finish: *request = the tables of which THIS rank's shard is past half full
 * (bit 0 s-mers, bit 1 b-mers, bit 2 pairs).  The caller ORs the requests of all ranks -- they can travel with the all-reduce of the
 * statistics -- and hands the result to fqsk_shard_grow on every rank; fqsk_sync_finish then returns FQSK_RESHARD when it is non-zero.
 * (fqsk_sync_device needs neither call: the requests cross NVLink with the statistics.) */
int fqsk_shard_grow_request(fqsk_handle *h, uint32_t *request);
int fqsk_shard_grow(fqsk_handle *h, uint32_t request_all_ranks);
/* After the all-reduce: the global p-mer statistics (bit_vec.h:204-220 -- they gate repair_kmers_missing on every worker) and
 * ClearKmersToHT.  The call sequence must be completed on all ranks before any of them starts its next segment.
 * Returns FQSK_RESHARD instead of FQSK_OK when table shards have to double first (see above). */
int fqsk_sync_finish(fqsk_handle *h, uint64_t fresh_all_ranks, uint64_t updates_all_ranks);
/* The three steps in one call, with NO collective of the caller's: the second barrier is a sequence number too (every rank posts it in
 * every peer's inbox once its inserts are enqueued, and waits for all of them on the device), and the global p-mer statistics travel
 * as NVLink atomics into every rank's accumulators.  All ranks must call it for every sync, as with the three-step form. */
int fqsk_sync_device(fqsk_handle *h);

/* Call after fqsk_sync, not between a segment and its sync (the grouping half of a small segment's sync may already be enqueued:
 * FQSK_E_INVAL).
 * Sorted (key, value) contents: FQSK_TABLE_SIV -> (p-mer index, 2-bit field); SMER/BMER -> (normalised k-mer, counter);
 * PAIR -> (key minimizer, value minimizer | count << 2b), the items of CHT_pair_kmers (ht_kmer.h:566-571).
 * Call with keys == NULL to get the count in *n.  Replaces nothing in the reference; parity check 1. */
int fqsk_dump(fqsk_handle *h, int table, uint64_t *keys, uint64_t *vals, uint64_t cap, uint64_t *n);
int fqsk_stats_get(fqsk_handle *h, fqsk_stats *out);
int fqsk_profile(fqsk_handle *h, double *ms, uint32_t n);   /* n <= FQSK_PH_COUNT */
/* CUDA-event stopwatch on the engine's own stream (bench.py times steps on the device with it): begin records an event,
 * end records a second one, waits for it and returns the elapsed device milliseconds. */
/* Measurement aid: on != 0 starts bracketing every kernel launch of the process with timing events on its own stream; on == 0 prints one
 * line per launch (kernel, stream, begin / end in microseconds) to stderr.  Replaces nothing in the reference. */
int fqsk_timeline(fqsk_handle *h, int on);
int fqsk_timer_begin(fqsk_handle *h);
int fqsk_timer_end(fqsk_handle *h, double *ms);

/* ---- table-level batch mirrors (unit parity with the reference classes; also the multi-GPU exchange building blocks) ---- */
/* CHT_kmer<T>::insert for a list in order, with the table's CCounterIncrementer stream (ht_kmer.h:420-438). */
int fqsk_ht_insert(fqsk_handle *h, int table, const uint64_t *kmers_normalized, uint64_t n);
/* CHT_kmer<T>::find on n registers in order (ht_kmer.h:504-510): full registers -> find_full, front-truncated -> find_partial
 * with the ordered PRNG-aware merge.  counts: 4 x u32 per query. */
int fqsk_ht_find(fqsk_handle *h, int table, const uint64_t *kmer_dir, const uint64_t *kmer_rc, const uint32_t *cur_size, uint64_t n, uint32_t *counts);
/* CHT_kmer<T>::count(uint64_t) (ht_kmer.h:441-454). */
int fqsk_ht_count(fqsk_handle *h, int table, const uint64_t *kmers_normalized, uint64_t n, uint32_t *out);
/* TSmallIntVector<2>::increment / test / counts / test_shorter (bit_vec.h:53-123). */
int fqsk_siv_increment(fqsk_handle *h, const uint64_t *idx, uint64_t n, uint64_t *n_new);
int fqsk_siv_test(fqsk_handle *h, const uint64_t *idx, uint64_t n, uint32_t *out);
int fqsk_siv_counts(fqsk_handle *h, const uint64_t *idx, uint64_t n, uint32_t *out4);
int fqsk_siv_test_shorter(fqsk_handle *h, const uint64_t *idx, const uint32_t *size_bits, uint64_t n, uint64_t *out);
/* Page-locked host memory for the caller's slab / record buffers (fqsk_segment copies to and from them with DMA; pageable
 * buffers work too, only slower).  Replaces nothing in the reference: its buffers are plain heap memory. */
int fqsk_host_alloc(uint64_t bytes, void **out);
void fqsk_host_free(void *p);

/* First n outputs of std::mt19937 seeded 5481 as generated on the device (utils.h:298). */
int fqsk_mt_stream(fqsk_handle *h, uint64_t n, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* FQSK_H */
