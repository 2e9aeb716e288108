#!/usr/bin/env python3
"""Builds the REAL reference (refresh-bio/fqsqueezer 1.1) into oracle/_ref/ -- test infrastructure only.

Nothing from /root/reference is copied into the repository: sources are compiled where they lie
(or from a scratch copy under /tmp when a tap has to be patched in), and only binaries land in
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun like our own .so files).

Products
  oracle/_ref/fqs-1.1          unmodified reference compressor (CPU baseline `kind: "reference"`,
                               .fqs byte-identity and decode checks)
  oracle/_ref/fqs-1.1-tap      same sources + a build-time tap: after the count-vector resolution of
                               every coded base (dna.cpp:736) it appends one 28-byte record to $FQS_TAP,
                               one marker per read and one per sync; at exit it dumps the k-mer tables to
                               $FQS_TAP_DUMP.  Its .fqs output must equal the untapped one (checked by
                               oracle/make_golden.py).
  oracle/_ref/fqs-1.1-replay   same sources, but compress_suffix (dna.cpp:674-877) takes counts / level / rough flag / cor_pos of
                               every coded base from the record stream in $FQS_REPLAY instead of calling its own k-mer engine
                               (find_counts, rough searches, pushes, repairs are all bypassed): the reference's untouched
                               context model + range coders consuming OUR records.  Its .fqs must be byte-identical to the
                               plain binary's -- north-star check 3 (tests/test_fqs_bytes.py).  Original order only.
  oracle/_ref/libfqs_ref.so    harness TU (oracle/ref_harness.cpp) over the reference's own
                               kmer.h / ht_kmer.h / bit_vec.h / utils.h for unit-level pinning.

The reference needs `-include cstdint` (defs.h:15 uses uint32_t without it).
This script is a no-op (exit 0) where /root/reference does not exist (the GPU box).
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FQS_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "fqs")
OUT = os.path.join(HERE, "_ref")
CXXFLAGS = ["-O3", "-m64", "-std=c++14", "-pthread", "-mavx", "-include", "cstdint", "-w"]


def run(cmd, **kw):
    subprocess.run(cmd, check=True, **kw)


def compile_dir(src_dir, build_dir, exe):
    os.makedirs(build_dir, exist_ok=True)
    cpps = sorted(f for f in os.listdir(src_dir) if f.endswith(".cpp"))

    def one(f):
        o = os.path.join(build_dir, f[:-4] + ".o")
        run(["g++", *CXXFLAGS, "-I", src_dir, "-c", os.path.join(src_dir, f), "-o", o])
        return o

    with ThreadPoolExecutor(8) as ex:
        objs = list(ex.map(one, cpps))
    run(["g++", "-O3", "-pthread", "-o", exe, *objs, "-lm"])


def patch(path, anchor, insertion, before=False, count=1):
    s = open(path, encoding='latin-1').read()
    assert s.count(anchor) >= 1, f"anchor not found in {path}: {anchor!r}"
    if before:
        s = s.replace(anchor, insertion + anchor, count)
    else:
        s = s.replace(anchor, anchor + insertion, count)
    open(path, 'w', encoding='latin-1').write(s)


TAP_H = r'''
#pragma once
// build-time tap (oracle/build_ref.py) -- not part of the reference
// One file per worker thread: $FQS_TAP for worker 0, $FQS_TAP.t<i> for worker i (the workers announce themselves with
// fqs_tap_set_thread; other threads never emit).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <string>
struct fqs_tap_rec { uint32_t pos; uint32_t c[4]; uint32_t cor_pos; uint8_t level; uint8_t rough; uint16_t pad; };
inline FILE **fqs_tap_files() { static FILE *f[64] = {nullptr}; return f; }
inline int &fqs_tap_tid() { static thread_local int tid = -1; return tid; }
inline void fqs_tap_set_thread(int tid) {
	if (tid < 0 || tid >= 64) return;
	fqs_tap_tid() = tid;
	const char *p = getenv("FQS_TAP");
	if (p && !fqs_tap_files()[tid]) { std::string n(p); if (tid) n += ".t" + std::to_string(tid); fqs_tap_files()[tid] = fopen(n.c_str(), "wb"); }
}
inline void fqs_tap_emit(uint32_t pos, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t cor_pos, uint8_t level, uint8_t rough) {
	int tid = fqs_tap_tid(); if (tid < 0) return;
	FILE *f = fqs_tap_files()[tid]; if (!f) return;
	fqs_tap_rec r{pos, {c0, c1, c2, c3}, cor_pos, level, rough, 0};
	fwrite(&r, sizeof(r), 1, f);
}
inline void fqs_tap_flush() { for (int i = 0; i < 64; ++i) if (fqs_tap_files()[i]) fflush(fqs_tap_files()[i]); }
// second tap ($FQS_TAP_CTX, worker 0 only): per base coded with counts, the 7 context ids of determine_ctx_codes and the rank of the
// true symbol (dna.cpp:748-760) -- 8 x u64
inline void fqs_tap_ctx(const uint64_t *ids, uint64_t r_sym) {
	static FILE *f = nullptr; static bool init = false;
	if (fqs_tap_tid() != 0) return;
	if (!init) { init = true; const char *p = getenv("FQS_TAP_CTX"); if (p) f = fopen(p, "wb"); }
	if (!f) return;
	uint64_t r[8]; for (int i = 0; i < 7; ++i) r[i] = ids[i]; r[7] = r_sym;
	fwrite(r, 8, 8, f);
	fflush(f);
}
'''

HT_FOREACH = r'''
	// build-time tap: visit every stored (k-mer, count)
	template<typename F> void tap_for_each(F f) const
	{
		for (uint64_t i = 0; i < tab_size; ++i)
			for (uint64_t j = 0; j < ht_desc[i].size; ++j)
				if (ht_desc[i].ht[j] != EMPTY)
					f(join_kmer((uint32_t) i, ht_desc[i].ht[j] >> (8 * sizeof(item_t) - 2 * (kmer_len - prefix_len))), get_count(ht_desc[i].ht[j]));
	}
'''

PAIR_FOREACH = r'''
	template<typename F> void tap_for_each(F f) const
	{
		for (uint64_t i = 0; i < no_parts; ++i)
			for (uint64_t j = 0; j < ht_desc[i].size; ++j)
				if (ht_desc[i].ht[j] != EMPTY)
					f(ht_desc[i].ht[j].first, ht_desc[i].ht[j].second);
	}
'''

SIV_FOREACH = r'''
	template<typename F> void tap_for_each(F f) const
	{
		for (uint64_t w = 0; w < data_size; ++w)
			if (data[w])
				for (uint64_t r = 0; r < divider; ++r)
				{
					uint64_t v = (data[w] >> (FIELD_SIZE * r)) & mask;
					if (v) f(w * divider + r, v);
				}
	}
'''

APP_DUMP = r'''
	// build-time tap: dump the k-mer tables (sorted by the test harness, not here)
	if (const char *dump_path = getenv("FQS_TAP_DUMP"))
	{
		FILE *df = fopen(dump_path, "wb");
		auto put = [&](uint64_t tag, uint64_t a, uint64_t b) { uint64_t r[3] = {tag, a, b}; fwrite(r, 8, 3, df); };
		siv_pmer->tap_for_each([&](uint64_t idx, uint64_t v) { put(0, idx, v); });
		ht_smer->tap_for_each([&](uint64_t k, uint32_t c) { put(1, k, c); });
		ht_bmer->tap_for_each([&](uint64_t k, uint32_t c) { put(2, k, c); });
		ht_pe_mers->tap_for_each([&](uint64_t k, uint64_t vc) { put(3, k, vc); });
		put(4, siv_pmer->get_no_pmers(), siv_pmer->get_no_updates());
		fclose(df);
	}
	fqs_tap_flush();
'''


def patch_headers(d):
    """Adds read-only visitors to the three table classes (scratch copy only)."""
    patch(os.path.join(d, "ht_kmer.h"),
          "public:\n\t// ************************************************************************************\n\tCHT_kmer(uint64_t _kmer_len",
          HT_FOREACH + "\n", before=True)
    patch(os.path.join(d, "ht_kmer.h"), "public:\n\tCHT_pair_kmers(uint32_t _kmer_size, uint64_t _no_parts);", PAIR_FOREACH)
    patch(os.path.join(d, "bit_vec.h"), "public:\n\tTSmallIntVector(const uint32_t _key_size)", SIV_FOREACH + "\n", before=True)
    # the visitor sits in the private section of CHT_kmer / TSmallIntVector -> re-open as public
    s = open(os.path.join(d, "ht_kmer.h"), encoding="latin-1").read()
    s = s.replace(HT_FOREACH, "public:" + HT_FOREACH + "private:\n", 1)
    open(os.path.join(d, "ht_kmer.h"), "w", encoding="latin-1").write(s)
    s = open(os.path.join(d, "bit_vec.h"), encoding="latin-1").read()
    s = s.replace(SIV_FOREACH, "public:" + SIV_FOREACH + "private:\n", 1)
    open(os.path.join(d, "bit_vec.h"), "w", encoding="latin-1").write(s)


def build_tap(scratch):
    d = os.path.join(scratch, "tap_src")
    if os.path.exists(d):
        shutil.rmtree(d)
    os.makedirs(d)
    for f in os.listdir(SRC):
        if f.endswith((".h", ".cpp")):
            shutil.copy(os.path.join(SRC, f), d)
    open(os.path.join(d, "fqs_tap.h"), "w").write(TAP_H)
    patch_headers(d)
    dna = os.path.join(d, "dna.cpp")
    patch(dna, '#include "dna.h"\n', '#include "fqs_tap.h"\n')
    # per coded base: after find_counts + bmer_unc rewrite + rough resolution (dna.cpp:736)
    patch(dna, "\t\tif (counts_level != counts_level_t::none && N_run_len < 2)\n\t\t{\n\t\t\tint cor_dist",
          "\t\tfqs_tap_emit(i, counts[0], counts[1], counts[2], counts[3], cor_pos, (uint8_t) counts_level, (uint8_t) rough_counts);\n",
          before=True)
    # the context ids the coder looks up for this base and the rank it codes (dna.cpp:748-760; the decoder repeats these lines: first hit only)
    patch(dna, "\t\t\tuint8_t r_sym = rank(counts, sym);\n\t\t\tp_rc5->Encode(r_sym);\n",
          "\t\t\tfqs_tap_ctx(ctx_lev_codes.data(), r_sym);\n")
    # per read markers (dna.cpp:1517, 1559, 1716): pos = 0xFFFFFFFF, c0 = size, c1 = kind
    patch(dna, "bool CDNACompressor::CompressDirect(uint8_t *p, uint32_t size, uint8_t *q, bool first_read_of_pair)\n{\n",
          "\tfqs_tap_emit(0xFFFFFFFFu, size, 0, first_read_of_pair, 0, 0, 0, 0);\n")
    patch(dna, "bool CDNACompressor::CompressDirectWithMinim(uint8_t *p, uint32_t size, uint32_t minim_pos)\n{\n",
          "\tfqs_tap_emit(0xFFFFFFFFu, size, 1, minim_pos, 0, 0, 0, 0);\n")
    patch(dna, "bool CDNACompressor::CompressSorted(uint8_t *p, uint32_t size, bool first_read_of_pair)\n{\n",
          "\tfqs_tap_emit(0xFFFFFFFFu, size, 2, first_read_of_pair, 0, 0, 0, 0);\n")
    # sorted-order prefix (dna.cpp:589-605): flag and, when present, the dif count -- pos = 0xFFFFFFFC, c0 = flag, c1/c2 = dif lo/hi
    patch(dna, "\tif (flag < max_prefix_sorted_flag_value)\n\t{\n\t\tdif = 0;", "\tif (flag >= max_prefix_sorted_flag_value) fqs_tap_emit(0xFFFFFFFCu, (uint32_t) flag, 0, 0, 0, 0, 0, 0);\n", before=True)
    patch(dna, "\t\tdif_no_bytes = no_bytes(dif);", "\t\tfqs_tap_emit(0xFFFFFFFCu, (uint32_t) flag, (uint32_t) dif, (uint32_t) (dif >> 32), 0, 0, 0, 0);\n", before=True)
    # duplicate flag: emitted right where the reference codes it
    patch(dna, "\t\tif (same_read)\n\t\t\treturn true;", "\t\tif (same_read) fqs_tap_emit(0xFFFFFFFDu, 0, 0, 0, 0, 0, 0, 0);\n", before=True, count=2)
    # per pair (dna.cpp:1848): what CompressPE decided for mate 2 -- pos = 0xFFFFFFFB, c0 = minim_found, c1 = minim2_id, c2 = minim2_pos
    patch(dna, "\tif (minim2_id < 0)\n\t\tCompressDirect(p2, size2, nullptr, false);",
          "\tfqs_tap_emit(0xFFFFFFFBu, (uint32_t) minim_found, minim_found ? (uint32_t) minim2_id : 0u, (minim_found && minim2_id < 15) ? minim2_pos : 0u, 0, 0, 0, 0);\n", before=True)
    # per sync marker (dna.cpp:2393)
    patch(dna, "void CDNACompressor::InsertKmersToHT()\n{\n", "\tfqs_tap_emit(0xFFFFFFFEu, 0, 0, 0, 0, 0, 0, 0);\n")
    app = os.path.join(d, "application.cpp")
    patch(app, '#include "application.h"\n', '#include "fqs_tap.h"\n')
    # dump the tables when the SE / PE compress loops are done (application.cpp:762, 1310)
    s = open(app, encoding="latin-1").read()
    # every worker thread announces its id to the tap (application.cpp:588, 1117: right where it picks its CDNACompressor)
    anchor = "\t\t\tCDNACompressor &dna_comp = v_dna_comp[thread_id];\n"
    assert s.count(anchor) >= 2, s.count(anchor)
    s = s.replace(anchor, anchor + "\t\t\tfqs_tap_set_thread((int) thread_id);\n")
    anchor = "\tv_thr_compress.clear();\n"
    assert s.count(anchor) == 2, s.count(anchor)
    s = s.replace(anchor, anchor + APP_DUMP)
    open(app, "w", encoding="latin-1").write(s)
    compile_dir(d, os.path.join(scratch, "tap_obj"), os.path.join(OUT, "fqs-1.1-tap"))


REPLAY_H = r'''
#pragma once
// build-time record replay (oracle/build_ref.py) -- not part of the reference
#include <cstdio>
#include <cstdlib>
#include <cstdint>
struct fqs_rp_rec { uint32_t pos; uint32_t c[4]; uint32_t cor_pos; uint8_t level; uint8_t rough; uint16_t pad; };
inline FILE *fqs_rp_file() {
	static FILE *f = nullptr; static bool init = false;
	if (!init) { init = true; const char *p = getenv("FQS_REPLAY"); if (p) { f = fopen(p, "rb"); if (!f) { fprintf(stderr, "replay: cannot open %s\n", p); exit(3); } } }
	return f;
}
inline bool fqs_rp_on() { return fqs_rp_file() != nullptr; }
inline fqs_rp_rec fqs_rp_next(uint32_t expect_pos) {
	fqs_rp_rec r;
	if (fread(&r, sizeof(r), 1, fqs_rp_file()) != 1) { fprintf(stderr, "replay: record stream ended early (position %u)\n", expect_pos); exit(3); }
	if (r.pos != expect_pos) { fprintf(stderr, "replay: record for position %u where %u is being coded\n", r.pos, expect_pos); exit(3); }
	return r;
}
'''


def build_replay(scratch):
    d = os.path.join(scratch, "replay_src")
    if os.path.exists(d):
        shutil.rmtree(d)
    os.makedirs(d)
    for f in os.listdir(SRC):
        if f.endswith((".h", ".cpp")):
            shutil.copy(os.path.join(SRC, f), d)
    open(os.path.join(d, "fqs_replay.h"), "w").write(REPLAY_H)
    dna = os.path.join(d, "dna.cpp")
    patch(dna, '#include "dna.h"\n', '#include "fqs_replay.h"\n')
    s = open(dna, encoding="latin-1").read()
    # (1) the count vector of a coded base comes from the record stream (dna.cpp:695)
    a = "\t\tcounts_level_t counts_level = find_counts(counts);\n"
    assert s.count(a) >= 1   # the decoder repeats these lines further down: the first hit is compress_suffix
    s = s.replace(a, "\t\tconst bool rp = fqs_rp_on();\n\t\tfqs_rp_rec rp_rec{};\n\t\tcounts_level_t counts_level;\n"
                     "\t\tif (rp) { rp_rec = fqs_rp_next(i); for (int q = 0; q < 4; ++q) counts[q] = rp_rec.c[q]; counts_level = (counts_level_t) rp_rec.level; cor_pos = rp_rec.cor_pos; }\n"
                     "\t\telse counts_level = find_counts(counts);\n", 1)
    # (2) no rough searches of its own (dna.cpp:709-735); the rough flag rides in the record
    a = "\t\tif (counts_level == counts_level_t::none)\n\t\t{\n\t\t\tif (bmer_can.is_full())\n\t\t\t{\n\t\t\t\tif (find_counts_rough_b(counts))"
    assert s.count(a) >= 1   # the decoder repeats these lines further down: the first hit is compress_suffix
    s = s.replace(a, a.replace("if (counts_level == counts_level_t::none)", "if (!rp && counts_level == counts_level_t::none)"), 1)
    a = "\t\tif (counts_level != counts_level_t::none && N_run_len < 2)\n\t\t{\n\t\t\tint cor_dist"
    assert s.count(a) >= 1   # the decoder repeats these lines further down: the first hit is compress_suffix
    s = s.replace(a, "\t\tif (rp) rough_counts = rp_rec.rough != 0;\n" + a, 1)
    # (3) no pushes, no thread-local inserts, no repairs (dna.cpp:810-876): the registers are not needed any more
    a = "\t\tpmer_can.replace_last(sym_to_kmers);\n\t\tsmer_can.replace_last(sym_to_kmers);\n\t\tbmer_can.replace_last(sym_to_kmers);\n"
    assert s.count(a) >= 1   # the decoder repeats these lines further down: the first hit is compress_suffix
    s = s.replace(a, "\t\tif (rp) continue;\n" + a, 1)
    open(dna, "w", encoding="latin-1").write(s)
    compile_dir(d, os.path.join(scratch, "replay_obj"), os.path.join(OUT, "fqs-1.1-replay"))


def build_harness(scratch):
    d = os.path.join(scratch, "hdr_src")
    if os.path.exists(d):
        shutil.rmtree(d)
    os.makedirs(d)
    for f in ("defs.h", "kmer.h", "ht_kmer.h", "ht_kmer.cpp", "bit_vec.h", "utils.h", "utils.cpp"):
        shutil.copy(os.path.join(SRC, f), d)
    patch_headers(d)
    run(["g++", *CXXFLAGS, "-fPIC", "-shared", "-I", d,
         os.path.join(HERE, "ref_harness.cpp"), os.path.join(d, "ht_kmer.cpp"), os.path.join(d, "utils.cpp"),
         "-o", os.path.join(OUT, "libfqs_ref.so")])


def main():
    if not os.path.isdir(SRC):
        print(f"[build_ref] {SRC} not present -- keeping prebuilt oracle/_ref (if any)")
        return 0
    os.makedirs(OUT, exist_ok=True)
    scratch = os.environ.get("FQS_REF_SCRATCH", "/tmp/fqs_ref_build")
    os.makedirs(scratch, exist_ok=True)
    what = sys.argv[1:] or ["plain", "tap", "replay", "harness"]
    if "plain" in what:
        compile_dir(SRC, os.path.join(scratch, "plain_obj"), os.path.join(OUT, "fqs-1.1"))
    if "tap" in what:
        build_tap(scratch)
    if "replay" in what:
        build_replay(scratch)
    if "harness" in what:
        build_harness(scratch)
    print("[build_ref] ok:", sorted(os.listdir(OUT)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
