// fqsk_live.h -- the reference-side binding of libfqsk.so: what a maintainer of refresh-bio/fqsqueezer adds to the tree so that
// CApplication's worker loop (application.cpp:575-671) and CDNACompressor::compress_suffix (dna.cpp:674-877) take the k-mer
// statistics of every coded base from the B200 engine instead of CHT_kmer / TSmallIntVector / CKmer.
//
// This file is OUR code (no reference source in it).  host/build_host.py compiles the reference's own sources -- where they lie,
// patched in a scratch copy -- with this header into host/_bin/fqs-1.1-fqsk.  INTEGRATION.md walks through the patch.
//
// Scope of the live host: single-end and paired-end, each in original and in sorted order (-s / -p with -om o / -om s; sorted is the
// reference's default, params.h:60), at -t 1 (the parity configuration of the north star: one engine, the asynchronous API, context
// records from the device) and at -t N <= 8 (one engine per worker thread = one GPU per worker, tables hash-sharded over the GPUs,
// fqsk_segment + fqsk_sync_device per sync segment: byte-identical to `fqs-1.1 -t N`).  Anything else stops with a message: there is no
// CPU fallback for the k-mer path.
//
// The library is bound with dlopen ($FQSK_LIB, default "libfqsk.so") so that the binary has no link-time CUDA dependency and the
// CPU test suite can point it at a mock built from the oracle (tests/mock_fqsk.cpp) to check the HOST half of the integration.
#pragma once
#include <dlfcn.h>
#include <time.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "fqsk.h"

struct fqs_rp_rec { uint32_t pos; uint32_t c[4]; uint32_t cor_pos; uint8_t level; uint8_t rough; uint16_t pad; };
static_assert(sizeof(fqs_rp_rec) == sizeof(fqsk_base_rec), "record layout");

// the library binding: one per process, shared by all workers
struct CFqskLib {
	void *lib = nullptr;
	decltype(&fqsk_create) p_create = nullptr;
	decltype(&fqsk_destroy) p_destroy = nullptr;
	decltype(&fqsk_last_error) p_last_error = nullptr;
	decltype(&fqsk_block_start) p_block_start = nullptr;
	decltype(&fqsk_submit) p_submit = nullptr;
	decltype(&fqsk_collect) p_collect = nullptr;
	decltype(&fqsk_submit_ctx) p_submit_ctx = nullptr;      // optional: context ids built on the device
	bool ctx_avail = false;
	decltype(&fqsk_stats_get) p_stats = nullptr;
	decltype(&fqsk_sorted_prefix) p_sorted_prefix = nullptr;
	decltype(&fqsk_pair_info) p_pair_info = nullptr;
	decltype(&fqsk_host_alloc) p_host_alloc = nullptr;
	decltype(&fqsk_host_free) p_host_free = nullptr;
	decltype(&fqsk_sort_ranks) p_sort_ranks = nullptr;      // sorted-order front end
	// -t N: one sharded engine per worker (bound on demand)
	decltype(&fqsk_segment) p_segment = nullptr;
	decltype(&fqsk_sync_device) p_sync_device = nullptr;
	decltype(&fqsk_shard_export) p_shard_export = nullptr;
	decltype(&fqsk_shard_attach_local) p_attach_local = nullptr;

	template <typename F> void sym(F &f, const char *name) {
		f = (F) dlsym(lib, name);
		if (!f) { fprintf(stderr, "fqsk: %s does not export %s\n", getenv("FQSK_LIB") ? getenv("FQSK_LIB") : "libfqsk.so", name); exit(3); }
	}
	// binds the library (once): create() needs it, and so does the sorted-order reader, which may open its first bin file earlier
	void load() {
		static std::mutex mu;
		std::lock_guard<std::mutex> lk(mu);
		if (lib) return;
		const char *path = getenv("FQSK_LIB");
		void *l = dlopen(path ? path : "libfqsk.so", RTLD_NOW | RTLD_LOCAL);
		if (!l) {   // the library needs the CUDA runtime: when the loader's search path does not have it, take it from $FQSK_CUDART or the toolkit
			const char *rt = getenv("FQSK_CUDART");
			if (dlopen(rt ? rt : "/usr/local/cuda/lib64/libcudart.so.12", RTLD_NOW | RTLD_GLOBAL)) l = dlopen(path ? path : "libfqsk.so", RTLD_NOW | RTLD_LOCAL);
		}
		if (!l) { fprintf(stderr, "fqsk: cannot load the k-mer engine (%s); set FQSK_LIB. There is no CPU fallback.\n", dlerror()); exit(3); }
		lib = l;
		sym(p_create, "fqsk_create"); sym(p_destroy, "fqsk_destroy"); sym(p_last_error, "fqsk_last_error"); sym(p_block_start, "fqsk_block_start");
		sym(p_submit, "fqsk_submit"); sym(p_collect, "fqsk_collect"); sym(p_stats, "fqsk_stats_get"); sym(p_host_alloc, "fqsk_host_alloc"); sym(p_host_free, "fqsk_host_free");
		sym(p_sorted_prefix, "fqsk_sorted_prefix"); sym(p_pair_info, "fqsk_pair_info");
		// with fqsk_submit_ctx the engine ships the 16-byte context records of include/fqsk_ctx.h instead of the 28-byte per-base records:
		// cor_zone, determine_ctx_codes and rank (dna.cpp:739-760) have run on the device.  FQSK_CTX=0 keeps the per-base records.
		p_submit_ctx = (decltype(p_submit_ctx)) dlsym(lib, "fqsk_submit_ctx");
		ctx_avail = p_submit_ctx && !(getenv("FQSK_CTX") && !strcmp(getenv("FQSK_CTX"), "0"));
		sym(p_sort_ranks, "fqsk_sort_ranks");
	}
	void load_sharded() {
		load();
		if (p_segment) return;
		sym(p_segment, "fqsk_segment"); sym(p_sync_device, "fqsk_sync_device"); sym(p_shard_export, "fqsk_shard_export"); sym(p_attach_local, "fqsk_shard_attach_local");
	}
	static CFqskLib &get() { static CFqskLib x; return x; }
};

class CFqskLive {
	CFqskLib &L = CFqskLib::get();
	bool ctx_on = false;
	uint32_t world = 1, rank = 0;      // -t N: this object is worker `rank` of `world` (one engine = one GPU each)

	fqsk_handle *h = nullptr;
	const uint8_t *slab = nullptr;
	uint64_t slab_size = 0;
	std::vector<fqsk_read_desc> descs;
	// two segments in flight (fqsk_submit / fqsk_collect): the engine works on segment n + 1 while the coder consumes segment n
	struct Slot { fqsk_base_rec *recs = nullptr; fqsk_ctx_rec *ctx = nullptr; uint64_t cap = 0; std::vector<uint8_t> dup; uint64_t ticket = 0; };
	Slot slot[2];
	struct Seg { uint64_t first, n; bool submitted; };     // the sync segments of the current reads_block, as the worker loop will cut them
	std::vector<Seg> plan;
	size_t plan_cur = 0;
	void *plan_reads = nullptr; size_t plan_stride = 0;
	fqsk_base_rec *recs = nullptr;
	fqsk_ctx_rec *ctxs = nullptr;
	uint64_t n_recs = 0, cursor = 0;
	const uint8_t *dup = nullptr;
	uint32_t mode = FQSK_MODE_SE_ORIGINAL;
	uint64_t n_seg_reads = 0, cur_read = 0;
	std::vector<uint32_t> s_flag, pair_words;      // per read: compress_prefix_sorted's flag; per pair: what CompressPE codes for mate 2
	std::vector<uint64_t> s_dif;
	uint64_t n_segments = 0, n_syncs = 0, n_bases = 0, n_reshards = 0;
	double engine_s = 0, wait_s = 0;

	[[noreturn]] void die(const char *what, int rc = 0) {
		fprintf(stderr, "fqsk: %s%s%s (rc %d)\n", what, (h || L.p_last_error) ? ": " : "", L.p_last_error ? L.p_last_error(h) : "", rc);
		exit(3);
	}
	static double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

public:
	// One object per worker thread of the reference (application.cpp:575-671).  The worker loop binds its object to its thread first
	// (bind_worker); every later CFqskLive::get() on that thread -- in the loop, in compress_suffix, in CompressPE -- is that worker's.
	// Threads that never bind (main, the readers) see worker 0: they only use the process-wide parts (create, sort_ranks, finish_all).
	static const uint32_t MAX_WORKERS = 8;
	static CFqskLive &worker(uint32_t t) { static CFqskLive x[MAX_WORKERS]; return x[t]; }
	static CFqskLive *&bound() { static thread_local CFqskLive *p = nullptr; return p; }
	static CFqskLive &get() { return bound() ? *bound() : worker(0); }
	static void bind_worker(uint32_t thread_id) { if (thread_id >= MAX_WORKERS) { fprintf(stderr, "fqsk: worker %u\n", thread_id); exit(3); } bound() = &worker(thread_id); }
	static uint32_t &n_workers() { static uint32_t n = 1; return n; }
	void load() { L.load(); }

	// io.h:499-528 (CSortedFASTQFile::sort_reads): the comparator's work -- two reads walked symbol by symbol per comparison -- is done once
	// per bin file on the GPU: rank[read_id] is order-isomorphic to it, std::sort then compares integers (same outcomes, same final order)
	template <typename V> std::vector<uint32_t> sort_ranks(const uint8_t *buffer, uint64_t size, V &v_reads) {
		load();
		const size_t n = v_reads.size();
		std::vector<fqsk_read_desc> d(n);
		for (size_t i = 0; i < n; ++i) {
			if (v_reads[i].second != i) { fprintf(stderr, "fqsk: sort_reads expects the reads in file order\n"); exit(3); }
			d[i].dna_off = (uint64_t) (v_reads[i].first.dna - buffer); d[i].dna_len = (uint32_t) v_reads[i].first.read_len(); d[i].flags = 0;
		}
		std::vector<uint32_t> rank(n);
		int rc = L.p_sort_ranks(getenv("FQSK_DEVICE") ? atoi(getenv("FQSK_DEVICE")) : 0, buffer, size, d.data(), (uint32_t) n, rank.data());
		if (rc != FQSK_OK) die("fqsk_sort_ranks", rc);
		return rank;
	}

	// application.cpp:86-91 (AdjustToParams): the engine takes the place of siv_pmer / ht_smer / ht_bmer
	// dna_mode: params.h:18 (0 se_original, 1 se_sorted, 2 pe_original, 3 pe_sorted) = FQSK_MODE_*
	// Called once, on the main thread (AdjustToParams): creates the engine of every worker thread.  -t 1: one unsharded engine.  -t N: worker
	// t is rank t of N on CUDA device $FQSK_DEVICE + t, the tables are hash-sharded by the reference's own owner keys (dna.cpp:825, 836,
	// 845, 1076-1081) and every engine reads its peers' shards directly (fqsk_shard_attach_local: NVLink peer access inside one process).
	static void create(uint32_t pmer_len, uint32_t smer_len, uint32_t bmer_len, uint32_t prefix_len, uint64_t genome_mbp, uint32_t dna_mode, uint32_t n_threads, bool dup_check) {
		if (dna_mode > FQSK_MODE_PE_SORTED || n_threads < 1 || n_threads > MAX_WORKERS || !dup_check) {
			fprintf(stderr, "fqsk: the live host covers -t 1 .. %u (one GPU per worker thread) with the duplicates check on (-s / -p, -om o / -om s); no CPU fallback for other modes\n", MAX_WORKERS);
			exit(3);
		}
		n_workers() = n_threads;
		for (uint32_t t = 0; t < n_threads; ++t) worker(t).create_one(pmer_len, smer_len, bmer_len, prefix_len, genome_mbp, dna_mode, n_threads, t);
		if (n_threads > 1) for (uint32_t t = 0; t < n_threads; ++t) worker(t).attach_peers();
	}
	void attach_peers() {
		for (uint32_t r = 0; r < world; ++r) if (r != rank) {
			int rc = L.p_attach_local(h, worker(r).h);
			if (rc != FQSK_OK) die("fqsk_shard_attach_local", rc);
		}
	}
	void create_one(uint32_t pmer_len, uint32_t smer_len, uint32_t bmer_len, uint32_t prefix_len, uint64_t genome_mbp, uint32_t dna_mode, uint32_t n_threads, uint32_t thread_id) {
		mode = dna_mode;
		world = n_threads; rank = thread_id;
		if (world > 1) L.load_sharded(); else L.load();
		ctx_on = L.ctx_avail && world == 1;      // a sharded engine is driven with the blocking pair (fqsk_segment + fqsk_sync_device): 28-byte count records
		fqsk_params P;
		memset(&P, 0, sizeof(P));
		P.abi_version = FQSK_ABI_VERSION;
		P.pmer_len = pmer_len; P.smer_len = smer_len; P.bmer_len = bmer_len; P.prefix_len = prefix_len;
		P.smer_counter_bits = 12; P.bmer_counter_bits = 6;                       // defs.h:26-27
		P.mode = mode;
		P.n_workers = world; P.world_size = world; P.rank = rank;
		P.device = (getenv("FQSK_DEVICE") ? atoi(getenv("FQSK_DEVICE")) : 0) + (int) rank;
		if (const char *e = getenv("FQSK_FLAGS")) P.flags = (uint32_t) strtoul(e, nullptr, 0) & (FQSK_F_PROFILE | FQSK_F_TRACE_ALLOC | FQSK_F_TRACE_LAUNCH | FQSK_F_SERIAL | FQSK_F_TEST_HOOKS | FQSK_F_TEST_CROWD);      // debugging aids of the library
		uint64_t expect = genome_mbp * 3000000ull;                               // genomic + error k-mers; the tables grow when half full
		if (const char *e = getenv("FQSK_EXPECTED_KMERS")) expect = strtoull(e, nullptr, 10);
		expect = expect < (1ull << 22) ? (1ull << 22) : expect > (1ull << 30) ? (1ull << 30) : expect;
		P.expected_kmers = world > 1 ? expect / world + (1ull << 20) : expect;      // per shard
		if (const char *e = getenv("FQSK_LOG2_BUCKETS")) P.bmer_log2_buckets = P.smer_log2_buckets = (uint32_t) atoi(e);      // tests: small tables that have to double
		if (const char *e = getenv("FQSK_PAIR_LOG2_SLOTS")) P.pair_log2_slots = (uint32_t) atoi(e);
		P.reserve_reads = 1u << 17; P.reserve_bytes = 16u << 20;                 // one reads_block (application.h:34) is the largest segment
		int rc = L.p_create(&P, &h);
		if (rc != FQSK_OK) { h = nullptr; die("fqsk_create", rc); }
	}

	// the workers of a sharded run meet here when tables have to double (FQSK_RESHARD): a plain generation barrier
	static void barrier() {
		static std::mutex mu; static std::condition_variable cv; static uint32_t arrived = 0; static uint64_t generation = 0;
		std::unique_lock<std::mutex> lk(mu);
		const uint64_t g = generation;
		if (++arrived == n_workers()) { arrived = 0; ++generation; cv.notify_all(); }
		else cv.wait(lk, [&] { return generation != g; });
	}

	// application.cpp:617-624: start of a reads_block for this worker.  The block's sync segments are known up front -- they follow from
	// my_first / my_last / no_synchronizations exactly as the worker loop computes next_synchro (SE: sync after read i == next_synchro,
	// application.cpp:643; PE: after the first pair with i >= next_synchro, 1170; plus the sync at the end of the block, 657-661 /
	// 1184-1190, possibly over an empty segment) -- so segment n + 1 can be handed to the engine before the coder starts on segment n.
	template <typename RD> void block_start(const uint8_t *input_FASTQ, uint64_t filled_size, RD *reads, uint64_t my_first, uint64_t my_last, uint64_t no_synchronizations, bool paired) {
		slab = input_FASTQ; slab_size = filled_size;
		int rc = L.p_block_start(h);
		if (rc != FQSK_OK) die("fqsk_block_start", rc);
		plan.clear(); plan_cur = 0;
		plan_reads = reads; plan_stride = sizeof(RD);
		const uint64_t step = paired ? 2 : 1;
		uint64_t gen = 0, a = my_first;
		uint64_t next = (gen + 1) * (my_last - my_first) / (no_synchronizations + 1) + my_first;
		for (uint64_t i = my_first; i < my_last; i += step)
			if (paired ? i >= next : i == next) {
				plan.push_back(Seg{a, i + step - a, false});
				a = i + step;
				++gen;
				next = (gen + 1) * (my_last - my_first) / (no_synchronizations + 1) + my_first;
			}
		plan.push_back(Seg{a, my_last - a, false});
	}

	template <typename RD> uint64_t fill_descs(const Seg &sg) {
		RD *reads = (RD *) plan_reads + sg.first;      // read_desc_t::read_len() is not const-qualified (defs.h:58-85)
		descs.resize(sg.n);
		uint64_t total = 0;
		for (uint64_t q = 0; q < sg.n; ++q) {
			descs[q].dna_off = (uint64_t) (reads[q].dna - slab);
			descs[q].dna_len = reads[q].read_len();
			descs[q].flags = 0;
			total += descs[q].dna_len;
		}
		return total;
	}
	void ensure_slot(Slot &sl, uint64_t total, uint64_t n) {
		if (total + 1 > sl.cap) {
			if (sl.recs) L.p_host_free(sl.recs);
			if (sl.ctx) L.p_host_free(sl.ctx);
			sl.recs = nullptr; sl.ctx = nullptr;
			sl.cap = total + total / 4 + 4096;
			void *p = nullptr;
			int rc = L.p_host_alloc(sl.cap * (ctx_on ? sizeof(fqsk_ctx_rec) : sizeof(fqsk_base_rec)), &p);
			if (rc != FQSK_OK) die("fqsk_host_alloc", rc);
			if (ctx_on) sl.ctx = (fqsk_ctx_rec *) p; else sl.recs = (fqsk_base_rec *) p;
		}
		sl.dup.resize(n + 1);
	}
	template <typename RD> void submit(size_t k) {
		Seg &sg = plan[k];
		Slot &sl = slot[k & 1];
		const uint64_t total = fill_descs<RD>(sg);
		ensure_slot(sl, total, sg.n);
		int rc = ctx_on ? L.p_submit_ctx(h, slab, slab_size, descs.data(), (uint32_t) sg.n, sl.ctx, sl.cap, sl.dup.data(), nullptr, &sl.ticket)
		                : L.p_submit(h, slab, slab_size, descs.data(), (uint32_t) sg.n, sl.recs, sl.cap, sl.dup.data(), nullptr, &sl.ticket);
		if (rc != FQSK_OK) die(ctx_on ? "fqsk_submit_ctx" : "fqsk_submit", rc);
		sg.submitted = true;
		n_bases += total;
	}

	// One sync segment: reads[0 .. n) of the current block, i.e. the reads the worker codes before the next InsertKmersToHT.
	// RD is the reference's read_desc_t (defs.h:58-85): only .dna and .read_len() are used.  The call must agree with the plan made at
	// the start of the block (it is the worker loop's own arithmetic, checked here).  fqsk_submit = the segment + the sync behind it.
	template <typename RD> void segment(RD *reads, uint64_t n) {
		if (plan_cur >= plan.size() || (RD *) plan_reads + plan[plan_cur].first != reads || plan[plan_cur].n != n) {
			fprintf(stderr, "fqsk: the worker loop's segment %llu (%llu reads) is not the planned one\n", (unsigned long long) plan_cur, (unsigned long long) n);
			exit(3);
		}
		const size_t k = plan_cur++;
		const double t0 = now();
		double t1 = t0;
		int rc;
		if (world > 1) {
			// -t N: the blocking call -- every worker's engine evaluates its segment against the shards of all of them; the tables change
			// only inside sync(), where the workers meet on the device (fqsk_sync_device)
			Slot &sl = slot[0];
			const uint64_t total = fill_descs<RD>(plan[k]);
			ensure_slot(sl, total, plan[k].n);
			rc = L.p_segment(h, slab, slab_size, descs.data(), (uint32_t) plan[k].n, sl.recs, sl.cap, &n_recs, sl.dup.data(), nullptr);
			if (rc != FQSK_OK) die("fqsk_segment", rc);
			plan[k].submitted = true;
			n_bases += total;
			recs = sl.recs; ctxs = nullptr; dup = sl.dup.data();
		} else {
			if (!plan[k].submitted) submit<RD>(k);
			if (k + 1 < plan.size() && !plan[k + 1].submitted) submit<RD>(k + 1);        // the engine runs ahead of the coder
			t1 = now();
			Slot &sl = slot[k & 1];
			rc = L.p_collect(h, sl.ticket, &n_recs);
			if (rc != FQSK_OK) die("fqsk_collect", rc);
			recs = sl.recs; ctxs = sl.ctx; dup = sl.dup.data();
		}
		if (mode == FQSK_MODE_SE_SORTED || mode == FQSK_MODE_PE_SORTED) {            // dna.cpp:589-605: (flag, dif) of every read's p-mer prefix (paired end: of the first mates)
			s_flag.resize(n + 1); s_dif.resize(n + 1);
			rc = L.p_sorted_prefix(h, s_flag.data(), s_dif.data(), (uint32_t) n);
			if (rc != FQSK_OK) die("fqsk_sorted_prefix", rc);
		}
		if (mode == FQSK_MODE_PE_ORIGINAL || mode == FQSK_MODE_PE_SORTED) {          // dna.cpp:1798-1838: the shared-minimizer decision of every pair
			if (n & 1) { fprintf(stderr, "fqsk: a paired-end segment with an odd number of reads\n"); exit(3); }
			pair_words.resize(3 * (n / 2) + 3);
			rc = L.p_pair_info(h, pair_words.data(), (uint32_t) (n / 2));
			if (rc != FQSK_OK) die("fqsk_pair_info", rc);
		}
		const double t2 = now();
		engine_s += t2 - t0; wait_s += t2 - t1;
		n_seg_reads = n; cur_read = 0;
		cursor = 0;
		++n_segments;
	}

	// the worker is about to code read `idx` of the segment (paired-end: its first mate)
	void set_read(uint64_t idx) {
		if (idx >= n_seg_reads) { fprintf(stderr, "fqsk: read %llu outside the segment of %llu reads\n", (unsigned long long) idx, (unsigned long long) n_seg_reads); exit(3); }
		cur_read = idx;
	}
	// dna.cpp:1523-1533, 1722-1732: the host keeps its own duplicate test (it owns read_prev); the engine must agree
	void check_dup(bool same_read) {
		if ((dup[cur_read] != 0) != same_read) { fprintf(stderr, "fqsk: duplicate flag of read %llu disagrees with the host\n", (unsigned long long) cur_read); exit(3); }
	}
	uint64_t sorted_flag() const { return s_flag[cur_read]; }
	uint64_t sorted_dif() const { return s_dif[cur_read]; }
	// dna.cpp:1798-1838: minim_found, minim2_id (-1 when no candidate list exists, else 0..15), minim2_pos
	void pair_decision(bool &minim_found, int &minim2_id, uint32_t &minim2_pos) const {
		const uint32_t *w = pair_words.data() + 3 * (cur_read / 2);
		minim_found = w[0] != 0;
		minim2_id = minim_found ? (int) w[1] : -1;
		minim2_pos = w[2];
	}

	bool ctx_mode() const { return ctx_on; }
	// dna.cpp:737-760 -- the context record of the base compress_suffix is about to code (ctx mode)
	const fqsk_ctx_rec &next_ctx() {
		if (cursor >= n_recs) { fprintf(stderr, "fqsk: record stream of the segment ended early\n"); exit(3); }
		return ctxs[cursor++];
	}
	// dna.cpp:695 -- the record of the base compress_suffix is about to code
	const fqsk_base_rec &next(uint32_t expect_pos) {
		if (cursor >= n_recs) { fprintf(stderr, "fqsk: record stream of the segment ended early (position %u)\n", expect_pos); exit(3); }
		const fqsk_base_rec &r = recs[cursor++];
		if (r.pos != expect_pos) { fprintf(stderr, "fqsk: record for position %u where %u is being coded\n", r.pos, expect_pos); exit(3); }
		return r;
	}

	// application.cpp:645-654 and 657-661: InsertKmersToHT + ClearKmersToHT between the barriers.  The engine has enqueued this sync right
	// behind the segment (fqsk_submit); what is left to do here is the coder's side of the contract: every record was consumed.
	// -t N: the sync is made here, by all workers (the reference's three barriers, application.cpp:645-654, become the two device-side
	// barriers of fqsk_sync_device).  FQSK_RESHARD: some shard is past half full and all shards of that table double -- the workers meet,
	// every engine doubles inside its fqsk_shard_export, they meet again and attach each other's new tables.
	void sync() {
		if (cursor != n_recs) { fprintf(stderr, "fqsk: %llu records of the segment were not consumed\n", (unsigned long long) (n_recs - cursor)); exit(3); }
		n_recs = cursor = 0;
		++n_syncs;
		if (world > 1) {
			const double t0 = now();
			int rc = L.p_sync_device(h);
			if (rc == FQSK_RESHARD) {
				barrier();
				fqsk_shard_desc d;
				rc = L.p_shard_export(h, &d);
				if (rc != FQSK_OK) die("fqsk_shard_export", rc);
				barrier();
				attach_peers();
				barrier();
				++n_reshards;
			} else if (rc != FQSK_OK) die("fqsk_sync_device", rc);
			engine_s += now() - t0;
		}
	}

	static void finish_all() { for (uint32_t t = 0; t < n_workers(); ++t) worker(t).finish(); }
	void finish() {
		if (!h) return;
		if (getenv("FQSK_VERBOSE")) {
			fqsk_stats st; memset(&st, 0, sizeof(st));
			L.p_stats(h, &st);
			const std::string who = world > 1 ? "worker " + std::to_string(rank) + "/" + std::to_string(world) + ": " : "";
			fprintf(stderr, "fqsk: %s%llu segments, %llu syncs, %llu bases, %.3f s inside the engine calls, %llu kernel launches, %llu s-mers, %llu b-mers; %.3f s waiting for the engine\n", who.c_str(),
			        (unsigned long long) n_segments, (unsigned long long) n_syncs, (unsigned long long) n_bases, engine_s,
			        (unsigned long long) st.kernel_launches, (unsigned long long) st.n_smers, (unsigned long long) st.n_bmers, wait_s);
			if (world > 1) fprintf(stderr, "fqsk: %s%llu coordinated table doublings (%llu tables doubled)\n", who.c_str(), (unsigned long long) n_reshards, (unsigned long long) st.n_table_growths);
		}
		if (getenv("FQSK_VERBOSE") && rank == 0) fprintf(stderr, "fqsk: per-base records: %s\n", ctx_on ? "16-byte context records built on the device (fqsk_submit_ctx)" : world > 1 ? "28-byte count records (fqsk_segment, sharded engines)" : "28-byte count records (fqsk_submit)");
		for (auto &sl : slot) { if (sl.recs) L.p_host_free(sl.recs); if (sl.ctx) L.p_host_free(sl.ctx); sl.recs = nullptr; sl.ctx = nullptr; }
		recs = nullptr;
		L.p_destroy(h);
		h = nullptr;
	}
};

// dna.cpp:695 as patched by host/build_host.py: the record of the base being coded, in the layout compress_suffix reads it
inline fqs_rp_rec fqs_rp_next(uint32_t expect_pos) {
	const fqsk_base_rec &k = CFqskLive::get().next(expect_pos);
	fqs_rp_rec r;
	memcpy(&r, &k, sizeof(r));
	return r;
}
