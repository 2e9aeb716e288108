#!/usr/bin/env python3
"""Builds host/_bin/fqs-1.1-fqsk: the reference compressor (refresh-bio/fqsqueezer 1.1) with its k-mer engine replaced by
libfqsk.so -- the drop-in of INTEGRATION.md, compiled, not sketched.

The reference's own sources are compiled from a scratch copy under /tmp (nothing from /root/reference enters the repository;
host/_bin/ is git-ignored and travels to the GPU box like our own .so files).  The patch below is the whole reference-side change:

  application.cpp  AdjustToParams (86-91)      the engine is created in place of siv_pmer / ht_smer / ht_bmer (which shrink to stubs)
                   worker loops                fqsk_block_start per reads_block; ONE fqsk_segment before the worker codes the reads of a
                   (617-662, 1145-1193)        sync segment; fqsk_sync in place of InsertKmersToHT + ClearKmersToHT
  dna.cpp          compress_suffix (674-877)   counts / level / rough flag / cor_pos of every coded base come from the segment's records;
                                               find_counts, the rough searches, repairs, pushes, thread-local inserts and prefetches are gone
                   compress_prefix_sorted      (flag, dif) of the p-mer prefix from fqsk_sorted_prefix instead of siv_pmer->test + the linear
                   (549-661)                   scan; no p-mer pushes
                   CompressPE (1790-1880)      the shared-minimizer decision of mate 2 from fqsk_pair_info instead of find_minim_cand /
                                               generate_read_bmers / the candidate search; no append_pe_mers3
  io.h             sort_reads (499-528)        the comparator of the per-bin std::sort compares ranks computed on the GPU (fqsk_sort_ranks)
  everything else (context model, range coders, id / quality / meta streams, container, decompressor) is untouched.

host/fqsk_live.h (ours) holds the binding itself (dlopen of $FQSK_LIB, descriptors, record cursor).
Scope: -s and -p, each with -om o and -om s, at -t 1 (one engine) and -t N <= 8 (one sharded engine = one GPU per worker thread).  No-op (exit 0) where /root/reference does not exist (the GPU box uses the prebuilt binary).
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("FQS_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "fqs")
OUT = os.path.join(HERE, "_bin")
EXE = os.path.join(OUT, "fqs-1.1-fqsk")
CXXFLAGS = ["-O3", "-m64", "-std=c++14", "-pthread", "-mavx", "-include", "cstdint", "-w"]   # the reference's own flags (makefile) + cstdint (defs.h:15)


def replace_once(s, old, new, where):
    assert s.count(old) >= 1, f"anchor not found in {where}: {old!r}"
    return s.replace(old, new, 1)


def in_range(s, start_marker, end_marker, fn, where):
    """Applies fn to the text between two markers (one function of the reference) and splices the result back."""
    a0 = s.index(start_marker)
    a1 = s.index(end_marker, a0 + len(start_marker))
    return s[:a0] + fn(s[a0:a1]) + s[a1:]


def patch_worker_loop(f, first_read_anchor, paired, where):
    """The compress worker loop of compress_se_files (application.cpp:617-662) / compress_pe_files (1145-1193)."""
    step = 2 if paired else 1
    # start of a reads_block (application.cpp:624 / 1152)
    f = replace_once(f, "\t\t\t\tdna_comp.ResetReadPrev();\n",
                     "\t\t\t\tdna_comp.ResetReadPrev();\n"
                     "\t\t\t\tCFqskLive::get().block_start(reads_block_cur->input_FASTQ, reads_block_cur->filled_size, reads_block_cur->v_reads.data(), my_first, my_last, no_synchronizations, %s);\n"
                     "\t\t\t\tuint64_t fqsk_seg_begin = my_first;\n" % ("true" if paired else "false"), where)
    # first read of a sync segment: the engine resolves the whole segment -- the reads up to and including the one that triggers the
    # next sync (SE: i == next_synchro, application.cpp:643; PE: the first pair with i >= next_synchro, 1170) or the tail of the block
    if paired:
        seg = ("\t\t\t\t\t\tuint64_t fqsk_seg_last = next_synchro > i ? next_synchro : i;\n"
               "\t\t\t\t\t\tif ((fqsk_seg_last - my_first) & 1) ++fqsk_seg_last;\n"
               "\t\t\t\t\t\tif (fqsk_seg_last + 2 > my_last) fqsk_seg_last = my_last - 2;\n"
               "\t\t\t\t\t\tCFqskLive::get().segment(&reads_block_cur->v_reads[i], fqsk_seg_last + 2 - i);\n")
    else:
        seg = ("\t\t\t\t\t\tuint64_t fqsk_seg_last = (next_synchro >= i && next_synchro < my_last) ? next_synchro : my_last - 1;\n"
               "\t\t\t\t\t\tCFqskLive::get().segment(&reads_block_cur->v_reads[i], fqsk_seg_last - i + 1);\n")
    f = replace_once(f, first_read_anchor,
                     "\t\t\t\t\tif (i == fqsk_seg_begin)\n\t\t\t\t\t{\n" + seg + "\t\t\t\t\t}\n"
                     "\t\t\t\t\tCFqskLive::get().set_read(i - fqsk_seg_begin);\n" + first_read_anchor, where)
    # the sync inside a block (application.cpp:645-649 / 1172-1176)
    f = replace_once(f, "\t\t\t\t\t\tdna_comp.InsertKmersToHT();\n\t\t\t\t\t\tbar_synchro.count_down_and_wait();\n\t\t\t\t\t\tdna_comp.ClearKmersToHT();\n",
                     f"\t\t\t\t\t\tCFqskLive::get().sync();\n\t\t\t\t\t\tfqsk_seg_begin = i + {step};\n\t\t\t\t\t\tbar_synchro.count_down_and_wait();\n", where)
    # the sync at the end of a block (application.cpp:657-661 / 1184-1190), over an empty segment when the last read closed one
    f = replace_once(f, "\t\t\t\tdna_comp.InsertKmersToHT();\n",
                     "\t\t\t\tif (fqsk_seg_begin >= my_last)\n\t\t\t\t\tCFqskLive::get().segment(reads_block_cur->v_reads.data() + fqsk_seg_begin, 0);\n"
                     "\t\t\t\tCFqskLive::get().sync();\n", where)
    f = replace_once(f, "\t\t\t\tdna_comp.ClearKmersToHT();\n", "", where)
    assert "InsertKmersToHT" not in f and "ClearKmersToHT" not in f
    # application.cpp:633-641 / 1161-1168 -- the four streams of a worker are independent (own models, own range coder, own output
    # vector; their bytes are concatenated per block, 685-747): the read-length + id streams and the quality stream are coded by two
    # helper threads over the same reads while this thread codes the DNA stream from the engine's records.  Same calls in the same order
    # per stream, so every stream's bytes are unchanged.
    if paired:
        f = replace_once(f, "\t\t\t\t\tmeta_comp.CompressReadLenPE(cur_read_1.read_len(), cur_read_2.read_len());\n", "", where)
        f = replace_once(f, "\t\t\t\t\tid_comp.CompressPE(cur_read_1.id, cur_read_1.id_len(), cur_read_2.id, cur_read_2.id_len());\n", "", where)
        f = replace_once(f, "\t\t\t\t\tquality_comp.Compress(cur_read_1.quality, cur_read_1.read_len());\n\t\t\t\t\tquality_comp.Compress(cur_read_2.quality, cur_read_2.read_len());\n", "", where)
        helpers = ("\t\t\t\tstd::thread fqsk_thr_ids([&] { for (uint64_t q = my_first; q < my_last; q += 2) { auto &r1 = reads_block_cur->v_reads[q]; auto &r2 = reads_block_cur->v_reads[q + 1];\n"
                   "\t\t\t\t\tmeta_comp.CompressReadLenPE(r1.read_len(), r2.read_len()); id_comp.CompressPE(r1.id, r1.id_len(), r2.id, r2.id_len()); } });\n"
                   "\t\t\t\tstd::thread fqsk_thr_qual([&] { for (uint64_t q = my_first; q < my_last; ++q) { auto &r = reads_block_cur->v_reads[q]; quality_comp.Compress(r.quality, r.read_len()); } });\n")
        loop_head = "\t\t\t\tfor (uint64_t i = my_first; i < my_last; i += 2)\n"
    else:
        f = replace_once(f, "\t\t\t\t\tmeta_comp.CompressReadLen(cur_read.read_len());\n", "", where)
        f = replace_once(f, "\t\t\t\t\tid_comp.Compress(cur_read.id, (uint32_t) cur_read.id_len());\n", "", where)
        f = replace_once(f, "\t\t\t\t\tquality_comp.Compress(cur_read.quality, cur_read.read_len());\n", "", where)
        helpers = ("\t\t\t\tstd::thread fqsk_thr_ids([&] { for (uint64_t q = my_first; q < my_last; ++q) { auto &r = reads_block_cur->v_reads[q];\n"
                   "\t\t\t\t\tmeta_comp.CompressReadLen(r.read_len()); id_comp.Compress(r.id, (uint32_t) r.id_len()); } });\n"
                   "\t\t\t\tstd::thread fqsk_thr_qual([&] { for (uint64_t q = my_first; q < my_last; ++q) { auto &r = reads_block_cur->v_reads[q]; quality_comp.Compress(r.quality, r.read_len()); } });\n")
        loop_head = "\t\t\t\tfor (uint64_t i = my_first; i < my_last; ++i)\n"
    f = replace_once(f, loop_head, helpers + loop_head, where)
    # the helpers are done before the coders are flushed (application.cpp:663-664 / 1192-1193)
    f = replace_once(f, "\t\t\t\tfor(auto &x : c_rc_enc)\n\t\t\t\t\tx.End();\n", "\t\t\t\tfqsk_thr_ids.join();\n\t\t\t\tfqsk_thr_qual.join();\n\t\t\t\tfor(auto &x : c_rc_enc)\n\t\t\t\t\tx.End();\n", where)
    # the workers have joined (application.cpp:762 / 1310): engine statistics, release
    f = replace_once(f, "\tv_thr_compress.clear();\n", "\tv_thr_compress.clear();\n\tCFqskLive::finish_all();\n", where)
    # application.cpp:590-592 / 1118-1120: the worker thread takes its engine (one CFqskLive object per worker; -t N = N engines = N GPUs)
    f = replace_once(f, "\t\t\tCDNACompressor &dna_comp = v_dna_comp[thread_id];\n", "\t\t\tCDNACompressor &dna_comp = v_dna_comp[thread_id];\n\t\t\tCFqskLive::bind_worker((uint32_t) thread_id);\n", where)
    return f


def patch_application(path):
    s = open(path, encoding="latin-1").read()
    s = replace_once(s, '#include "application.h"\n', '#include "application.h"\n#include "fqsk_live.h"\n', path)
    # application.cpp:86-91 -- the engine replaces the three global tables; the CPU objects stay as stubs nobody reads
    s = replace_once(s, "\tsiv_pmer = new TSmallIntVector<SIV_FIELD_SIZE>(2 * params.pmer_len);\n"
                        "\tht_smer = new CHT_kmer<uint32_t>(params.smer_len, SMER_COUNTER_BITS, params.ht_max_filling_factor);\n"
                        "\tht_bmer = new CHT_kmer<uint32_t>(params.bmer_len, BMER_COUNTER_BITS, params.ht_max_filling_factor);\n",
                     "\tCFqskLive::create(params.pmer_len, params.smer_len, params.bmer_len, params.prefix_len, params.genome_size,\n"
                     "\t\t(uint32_t) params.dna_mode, (uint32_t) params.no_threads, params.duplicates_check);\n"
                     "\tsiv_pmer = new TSmallIntVector<SIV_FIELD_SIZE>(8);\n"
                     "\tht_smer = new CHT_kmer<uint32_t>(12, SMER_COUNTER_BITS, params.ht_max_filling_factor);\n"
                     "\tht_bmer = new CHT_kmer<uint32_t>(12, BMER_COUNTER_BITS, params.ht_max_filling_factor);\n", path)
    s = in_range(s, "bool CApplication::compress_se_files(", "bool CApplication::decompress_se_file(",
                 lambda f: patch_worker_loop(f, "\t\t\t\t\tauto &cur_read = reads_block_cur->v_reads[i];\n", False, path), path)
    s = in_range(s, "bool CApplication::compress_pe_files(", "bool CApplication::decompress_pe_file(",
                 lambda f: patch_worker_loop(f, "\t\t\t\t\tauto &cur_read_1 = reads_block_cur->v_reads[i];\n", True, path), path)
    open(path, "w", encoding="latin-1").write(s)


def patch_suffix(f, path):
    # dna.cpp:695 -- the count vector of a coded base comes from the segment's records
    # ... or, with fqsk_submit_ctx, the 16-byte context record: whether the base is coded with counts, the 7 context ids, the rank
    f = replace_once(f, "\t\tcounts_level_t counts_level = find_counts(counts);\n",
                     "\t\tconst bool fqsk_cm = CFqskLive::get().ctx_mode();\n"
                     "\t\tfqs_rp_rec rp_rec{};\n\t\tconst fqsk_ctx_rec *fqsk_cx = nullptr;\n\t\tcounts_level_t counts_level;\n"
                     "\t\tif (fqsk_cm)\n\t\t{\n\t\t\tfqsk_cx = &CFqskLive::get().next_ctx();\n"
                     "\t\t\tcounts_level = fqsk_ctx_coded(fqsk_cx) ? counts_level_t::bmer : counts_level_t::none;\t// only none / not none matters from here on\n\t\t}\n"
                     "\t\telse\n\t\t{\n\t\t\trp_rec = fqs_rp_next(i);\n"
                     "\t\t\tfor (int q = 0; q < 4; ++q) counts[q] = rp_rec.c[q];\n"
                     "\t\t\tcounts_level = (counts_level_t) rp_rec.level;\n"
                     "\t\t\tcor_pos = rp_rec.cor_pos;\n\t\t}\n", path)
    # dna.cpp:746-752, 759 -- determine_ctx_codes and rank ran on the device (ctx mode)
    f = replace_once(f, "\t\t\tcontext_levels_t ctx_lev_codes;\n\t\t\tif(!reversed_pe)\n",
                     "\t\t\tcontext_levels_t ctx_lev_codes;\n"
                     "\t\t\tif (fqsk_cm)\n\t\t\t{\n\t\t\t\tuint64_t fqsk_ids[7];\n\t\t\t\tfqsk_ctx_expand(fqsk_cx, fqsk_ids);\n"
                     "\t\t\t\tfor (int q = 0; q < 7; ++q) ctx_lev_codes[q] = fqsk_ids[q];\n\t\t\t}\n\t\t\telse if(!reversed_pe)\n", path)
    f = replace_once(f, "\t\t\tuint8_t r_sym = rank(counts, sym);\n", "\t\t\tuint8_t r_sym = fqsk_cm ? (uint8_t) fqsk_ctx_rsym(fqsk_cx) : rank(counts, sym);\n", path)
    # dna.cpp:709-735 -- no rough searches on the host; the rough flag rides in the record
    a = "\t\tif (counts_level == counts_level_t::none)\n\t\t{\n\t\t\tif (bmer_can.is_full())\n\t\t\t{\n\t\t\t\tif (find_counts_rough_b(counts))"
    f = replace_once(f, a, a.replace("if (counts_level == counts_level_t::none)", "if (false)"), path)
    a = "\t\tif (counts_level != counts_level_t::none && N_run_len < 2)\n\t\t{\n\t\t\tint cor_dist"
    f = replace_once(f, a, "\t\trough_counts = rp_rec.rough != 0;\n" + a, path)
    # dna.cpp:765-773, 785-793 -- no table prefetches: the tables live in HBM
    assert f.count("if(bmer_can.is_almost_full(1))") + f.count("if (bmer_can.is_almost_full(1))") == 2
    f = f.replace("if(bmer_can.is_almost_full(1))", "if (false)").replace("if (bmer_can.is_almost_full(1))", "if (false)")
    # dna.cpp:810-876 -- no register updates, pushes, thread-local inserts or repairs on the host
    a = "\t\tpmer_can.replace_last(sym_to_kmers);\n\t\tsmer_can.replace_last(sym_to_kmers);\n\t\tbmer_can.replace_last(sym_to_kmers);\n"
    f = replace_once(f, a, "\t\tcontinue;\n" + a, path)
    # dna.cpp:684-693 -- the six register shifts at the top of the loop body have no reader left
    a = ("\t\tpmer_can.insert_zero();\n\t\tsmer_can.insert_zero();\n\t\tbmer_can.insert_zero();\n\n"
         "\t\tpmer_can_unc.insert_zero();\n\t\tsmer_can_unc.insert_zero();\n\t\tbmer_can_unc.insert_zero();\n")
    return replace_once(f, a, "", path)


def patch_prefix_sorted(f, path):
    # dna.cpp:589-592 -- flag: 4 when the p-mer equals the previous read's, else siv_pmer->test(p-mer)
    f = replace_once(f, "\tif (pmer_can == pmer_can_prev)\n\t\tflag = max_prefix_sorted_flag_value;\n\telse\n\t\tflag = siv_pmer->test(pmer_can.data_aligned_dir());\n",
                     "\tflag = CFqskLive::get().sorted_flag();\n", path)
    # dna.cpp:600-605 -- dif: the linear scan over the p-mer array between the previous and this p-mer runs on the GPU
    f = replace_once(f, "\t\tuint64_t max_i = pmer_can.data_aligned_dir();\n\n\t\tfor (uint64_t i = pmer_can_prev.data_aligned_dir() + 1; i < max_i; ++i)\n"
                        "\t\t\tif (siv_pmer->test(i) == flag)\n\t\t\t\t++dif;\n",
                     "\t\tdif = CFqskLive::get().sorted_dif();\n", path)
    # dna.cpp:655-660 -- no p-mer pushes on the host
    a = ("\tauto x = pmer_can.data_aligned_dir();\n\t(*my_pmers_to_add)[modulo_divisor(x >> pmer_mod_shift, no_threads)].push_back(x);\n"
         "\tx = pmer_can.data_aligned_rc();\n\t(*my_pmers_to_add)[modulo_divisor(x >> pmer_mod_shift, no_threads)].push_back(x);\n")
    return replace_once(f, a, "", path)


def patch_compress_pe(f, path):
    # dna.cpp:1798-1838 -- find_minim_cand, generate_read_bmers and the candidate search run on the GPU; the decision comes back per pair
    a0 = f.index("\tbool minim_found = find_minim_cand(p1, size1);\n")
    a1 = f.index("\tif (minim2_id < 0)\n\t\tCompressDirect(p2, size2, nullptr, false);")
    f = (f[:a0] + "\tbool minim_found;\n\tuint32_t minim2_pos;\n\tint minim2_id;\n"
         "\tCFqskLive::get().pair_decision(minim_found, minim2_id, minim2_pos);\n\n" + f[a1:])
    # dna.cpp:1877 -- the 14 pair pushes of append_pe_mers3 are made by the engine
    return replace_once(f, "\tappend_pe_mers3(p1, size1, p2, size2);\n", "", path)


def patch_dup(f, path):
    # dna.cpp:1525 / 1724 -- the host keeps its own duplicate test; the engine's flag must agree
    return replace_once(f, "\t\tbool same_read = read_cur == read_prev;\n", "\t\tbool same_read = read_cur == read_prev;\n\t\tCFqskLive::get().check_dup(same_read);\n", path)


def patch_dna(path):
    s = open(path, encoding="latin-1").read()
    s = replace_once(s, '#include "dna.h"\n', '#include "dna.h"\n#include "fqsk_live.h"\n', path)
    sep = "\n//****"
    s = in_range(s, "void CDNACompressor::compress_suffix(", sep, lambda f: patch_suffix(f, path), path)
    s = in_range(s, "void CDNACompressor::compress_prefix_sorted(", sep, lambda f: patch_prefix_sorted(f, path), path)
    s = in_range(s, "bool CDNACompressor::CompressPE(", sep, lambda f: patch_compress_pe(f, path), path)
    s = in_range(s, "bool CDNACompressor::CompressDirect(", sep, lambda f: patch_dup(f, path), path)
    s = in_range(s, "bool CDNACompressor::CompressSorted(", sep, lambda f: patch_dup(f, path), path)
    open(path, "w", encoding="latin-1").write(s)


def patch_io(path):
    """io.h:499-528 -- CSortedFASTQFile::sort_reads keeps its std::sort call; the comparator compares the ranks the GPU computed for the bin
    (fqsk_sort_ranks): the same outcome for every pair of reads, so the same order, ties of the unstable sort included."""
    s = open(path, encoding="latin-1").read()
    s = replace_once(s, "#pragma once\n", '#pragma once\n#include "fqsk_live.h"\n', path) if "#pragma once\n" in s else replace_once(s, '#include "defs.h"\n', '#include "defs.h"\n#include "fqsk_live.h"\n', path)
    a0 = s.index("\tvoid sort_reads()\n\t{\n")
    a1 = s.index("\t//Return ordering of the sorted collection of reads", a0)
    body = ("\tvoid sort_reads()\n\t{\n"
            "\t\tstd::vector<uint32_t> fqsk_rank = CFqskLive::get().sort_ranks(buffer, buffer_filled, v_reads);\n"
            "\t\tsort(v_reads.begin(), v_reads.end(), [&](pair<read_desc_t, uint32_t> &x, pair<read_desc_t, uint32_t> &y) {\n"
            "\t\t\treturn fqsk_rank[x.second] < fqsk_rank[y.second];\n\t\t});\n\t}\n\n")
    assert "dna_convert_NT[x.first.dna[i]]" in s[a0:a1]
    s = s[:a0] + body + s[a1:]
    open(path, "w", encoding="latin-1").write(s)


def compile_dir(src_dir, build_dir, exe):
    os.makedirs(build_dir, exist_ok=True)
    cpps = sorted(f for f in os.listdir(src_dir) if f.endswith(".cpp"))

    def one(f):
        o = os.path.join(build_dir, f[:-4] + ".o")
        subprocess.run(["g++", *CXXFLAGS, "-I", src_dir, "-I", os.path.join(ROOT, "include"), "-I", HERE, "-c", os.path.join(src_dir, f), "-o", o], check=True)
        return o

    with ThreadPoolExecutor(8) as ex:
        objs = list(ex.map(one, cpps))
    subprocess.run(["g++", "-O3", "-pthread", "-o", exe, *objs, "-lm", "-ldl"], check=True)


def main():
    if not os.path.isdir(SRC):
        print(f"[build_host] {SRC} not present -- keeping prebuilt host/_bin (if any)")
        return 0
    scratch = os.path.join(os.environ.get("FQS_REF_SCRATCH", "/tmp/fqs_ref_build"), "host_src")
    if os.path.exists(scratch):
        shutil.rmtree(scratch)
    os.makedirs(scratch)
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(SRC):
        if f.endswith((".h", ".cpp")):
            shutil.copy(os.path.join(SRC, f), scratch)
    patch_application(os.path.join(scratch, "application.cpp"))
    patch_dna(os.path.join(scratch, "dna.cpp"))
    patch_io(os.path.join(scratch, "io.h"))
    compile_dir(scratch, os.path.join(os.path.dirname(scratch), "host_obj"), EXE)
    print("[build_host] ok:", EXE)
    return 0


if __name__ == "__main__":
    sys.exit(main())
