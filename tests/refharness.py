"""ctypes binding of the library-level oracle oracle/_ref/libt4ref.so (TEST INFRASTRUCTURE)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libt4ref.so")


def available():
    return os.path.exists(LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(LIB)
        l.t4ref_create.restype = C.c_void_p
        l.t4ref_create.argtypes = [C.c_int]
        l.t4ref_destroy.argtypes = [C.c_void_p]
        l.t4ref_set_hit_len_required.argtypes = [C.c_void_p, C.c_int]
        l.t4ref_set_novel_seq_similarity.argtypes = [C.c_void_p, C.c_double]
        l.t4ref_set_novel_seq_similarity.restype = C.c_double
        l.t4ref_size.argtypes = [C.c_void_p]
        l.t4ref_merge_sets.restype = C.c_void_p
        l.t4ref_merge_sets.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_int)]
        l.t4ref_input_contig.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        l.t4ref_set_consider_barcode_in_hash.argtypes = [C.c_void_p, C.c_int]
        l.t4ref_reverse_complement_in_place.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        l.t4ref_kmer_length.argtypes = [C.c_void_p]
        l.t4ref_add_read.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_double]
        l.t4ref_repeat_add_read.argtypes = [C.c_void_p, C.c_char_p]
        l.t4ref_input_novel_read.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        l.t4ref_update_all_consensus.argtypes = [C.c_void_p]
        l.t4ref_change_kmer_length.argtypes = [C.c_void_p, C.c_int]
        l.t4ref_has_motif.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        l.t4ref_output_mem.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        l.t4ref_free.argtypes = [C.c_void_p]
        l.t4ref_get_hits.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        l.t4ref_get_chains.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        l.t4ref_get_overlaps.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        l.t4ref_dp_pos_weight.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p]
        l.t4ref_index_lookup.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int]
        l.t4ref_index_checksum.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        l.t4ref_index_checksum.restype = C.c_int64
        l.t4ref_nomatch_gap_limit.argtypes = [C.c_void_p]
        l.t4ref_get_contig.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_char_p, C.c_int,
                                       C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        l.t4ref_run_descs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_char_p), C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        l.t4ref_release_finished_barcode.argtypes = [C.c_void_p, C.c_int, C.c_int]
        l.t4ref_release_shallow_contigs.argtypes = [C.c_void_p, C.c_int]
        l.t4ref_input_novel_fa.argtypes = [C.c_void_p, C.c_char_p]
        l.t4ref_num_read.argtypes = [C.c_void_p, C.c_int]
        l.t4ref_kmer_count_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
        l.t4ref_refset_create.restype = C.c_void_p
        l.t4ref_refset_create.argtypes = [C.c_char_p, C.c_int]
        l.t4ref_seq_name.restype = C.c_char_p
        l.t4ref_seq_name.argtypes = [C.c_void_p, C.c_int]
        l.t4ref_seq_count.argtypes = [C.c_void_p]
        l.t4ref_set_radius.argtypes = [C.c_void_p, C.c_int]
        l.t4ref_has_hit_in_set.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        l.t4ref_is_low_complexity.argtypes = [C.c_char_p]
        l.t4ref_annotate_read.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
        l.t4ref_sort_reads.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.t4ref_is_mate_overlap.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        l.t4ref_lis.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        l.t4ref_input_seqset.restype = C.c_void_p
        l.t4ref_input_seqset.argtypes = [C.c_void_p, C.c_int]
        l.t4ref_assign_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
        l.t4ref_assign_pass.restype = C.c_void_p
        l.t4ref_assign_pass.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


class RefSeqSet:
    """The reference SeqSet (compiled from /root/reference) behind the same call names as trust4_b200.api.SeqSet."""

    def __init__(self, k=9, handle=None):
        self.l = lib()
        self.h = handle if handle is not None else self.l.t4ref_create(k)

    def close(self):
        if self.h:
            self.l.t4ref_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_hit_len_required(self, v):
        return self.l.t4ref_set_hit_len_required(self.h, v)

    def size(self):
        return self.l.t4ref_size(self.h)

    def kmer_length(self):
        return self.l.t4ref_kmer_length(self.h)

    def add_read(self, read, name, strand, barcode, min_kmer_count, repetitive, thr):
        s = C.c_int(strand)
        r = self.l.t4ref_add_read(self.h, read.encode(), name.encode(), C.byref(s), barcode, min_kmer_count, int(repetitive), thr)
        return r, s.value

    def repeat_add_read(self, read):
        return self.l.t4ref_repeat_add_read(self.h, read.encode())

    def input_novel_read(self, name, read, strand, barcode):
        return self.l.t4ref_input_novel_read(self.h, name.encode(), read.encode(), strand, barcode)

    def update_all_consensus(self):
        self.l.t4ref_update_all_consensus(self.h)

    def change_kmer_length(self, k):
        self.l.t4ref_change_kmer_length(self.h, k)

    def release_finished_barcode(self, barcode, contig_min_cov=0):
        self.l.t4ref_release_finished_barcode(self.h, barcode, contig_min_cov)

    def release_shallow_contigs(self, min_cov):
        self.l.t4ref_release_shallow_contigs(self.h, min_cov)

    def input_novel_fa(self, filename):
        self.l.t4ref_input_novel_fa(self.h, filename.encode())

    def num_read(self, slot):
        return self.l.t4ref_num_read(self.h, slot)

    def has_motif(self, read, strand):
        return self.l.t4ref_has_motif(self.h, read.encode(), strand)

    def output(self):
        buf = C.c_void_p()
        n = C.c_size_t()
        self.l.t4ref_output_mem(self.h, C.byref(buf), C.byref(n))
        s = C.string_at(buf, n.value)
        self.l.t4ref_free(buf)
        return s

    def get_hits(self, read, strand=0, barcode=-1, allow_total_skip=False, cap=1 << 20):
        out = np.zeros((cap, 5), dtype=np.int32)
        n = self.l.t4ref_get_hits(self.h, read.encode(), strand, barcode, int(allow_total_skip), out.ctypes.data, cap)
        assert n <= cap
        return out[:n]

    def get_chains(self, read, strand=0, barcode=-1, allow_total_skip=False, filt=1, cap=1 << 14, ccap=1 << 20):
        out = np.zeros((cap, 8), dtype=np.int32)
        co = np.zeros((ccap, 2), dtype=np.int32)
        nc = C.c_int()
        n = self.l.t4ref_get_chains(self.h, read.encode(), strand, barcode, int(allow_total_skip), filt, out.ctypes.data, cap,
                                    co.ctypes.data, ccap, C.byref(nc))
        assert n <= cap and nc.value <= ccap
        return out[:n], co[:nc.value]

    def get_overlaps(self, read, strand=0, barcode=-1, skip_repeats=False, cap=1 << 14):
        out = np.zeros((cap, 8), dtype=np.int32)
        sim = np.zeros(cap, dtype=np.float64)
        n = self.l.t4ref_get_overlaps(self.h, read.encode(), strand, barcode, int(skip_repeats), out.ctypes.data, sim.ctypes.data, cap)
        if n < 0:
            return n, None, None
        return n, out[:n], sim[:n]

    def index_lookup(self, code, barcode=-1, cap=1 << 20):
        out = np.zeros((cap, 2), dtype=np.int32)
        n = self.l.t4ref_index_lookup(self.h, code, barcode, out.ctypes.data, cap)
        return out[:n]

    def index_checksum(self):
        cs = C.c_uint64()
        n = self.l.t4ref_index_checksum(self.h, C.byref(cs))
        return n, cs.value

    def get_contig(self, slot):
        ln = self.l.t4ref_get_contig(self.h, slot, None, 0, None, None, 0, None, None, None, None)
        if ln < 0:
            return None
        cons = C.create_string_buffer(ln + 1)
        pw = np.zeros((ln, 4), dtype=np.int32)
        name = C.create_string_buffer(4096)
        bc, nr, ml, mr = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.l.t4ref_get_contig(self.h, slot, cons, ln + 1, pw.ctypes.data, name, 4096, C.byref(bc), C.byref(nr), C.byref(ml), C.byref(mr))
        return dict(consensus=cons.value.decode(), pos_weight=pw, name=name.value.decode(), barcode=bc.value,
                    num_read=nr.value, min_left=ml.value, min_right=mr.value)

    def run_descs(self, cfg, descs, pool, names):
        n = len(descs)
        ret = np.zeros(n, dtype=np.int32)
        strands = np.zeros(n, dtype=np.int8)
        resc = np.zeros(n, dtype=np.int32)
        arr = (C.c_char_p * max(1, len(names)))(*names)
        descs = np.ascontiguousarray(descs)
        pool = np.ascontiguousarray(pool)
        a = self.l.t4ref_run_descs(self.h, cfg.ctypes.data, descs.ctypes.data, n, pool.ctypes.data, arr, len(names),
                                   ret.ctypes.data, strands.ctypes.data, resc.ctypes.data)
        return a, ret, strands, resc


def dp_pos_weight(tw, p):
    """AlignAlgo::GlobalAlignment_PosWeight via the reference; returns (score, edit list)."""
    tw = np.ascontiguousarray(tw, dtype=np.int32)
    lent, lenp = tw.shape[0], len(p)
    al = np.zeros(lent + lenp + 4, dtype=np.int8)
    sc = lib().t4ref_dp_pos_weight(tw.ctypes.data, lent, p.encode(), lenp, al.ctypes.data)
    e = []
    for v in al:
        if v == -1:
            break
        e.append(int(v))
    return sc, e


class RefGeneSet(RefSeqSet):
    """The reference's `SeqSet refSet(k); refSet.InputRefFa(fasta)` (fastq-extractor / trust4 reference gene set)."""

    def __init__(self, fasta, k=9, hit_len_required=27):
        l = lib()
        super().__init__(k, handle=l.t4ref_refset_create(fasta.encode(), k))
        self.set_hit_len_required(hit_len_required)

    def names(self):
        return [self.l.t4ref_seq_name(self.h, i).decode() for i in range(self.l.t4ref_seq_count(self.h))]

    def set_radius(self, r):
        self.l.t4ref_set_radius(self.h, r)

    def annotate_read(self, read):
        """AnnotateRead(read, 0, ...): (int32[4, 8] for V, D, J, C; similarity[4])."""
        out = np.zeros((4, 8), dtype=np.int32)
        sim = np.zeros(4, dtype=np.float64)
        self.l.t4ref_annotate_read(self.h, read.encode(), out.ctypes.data, sim.ctypes.data)
        return out, sim

    def has_hit_in_set(self, read, mode=0):
        return self.l.t4ref_has_hit_in_set(self.h, read.encode(), mode)


def is_low_complexity(read):
    return lib().t4ref_is_low_complexity(read.encode())


def kmer_count_stats(pool, seq_off, lens, k=21, qual=None):
    """The reference's KmerCount over the reads: (min, median, avg, new length) per read (KmerCount.hpp:64-97, 177-288)."""
    pool = np.ascontiguousarray(pool)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    n = len(lens)
    mn = np.zeros(max(1, n), dtype=np.int32)
    med = np.zeros(max(1, n), dtype=np.int32)
    avg = np.zeros(max(1, n), dtype=np.float32)
    nl = np.zeros(max(1, n), dtype=np.int32)
    if qual is not None:
        qual = np.ascontiguousarray(qual)
    lib().t4ref_kmer_count_stats(pool.ctypes.data, qual.ctypes.data if qual is not None else None, seq_off.ctypes.data, lens.ctypes.data, n, k,
                                 mn.ctypes.data, med.ctypes.data, avg.ctypes.data, nl.ctypes.data)
    return mn[:n], med[:n], avg[:n], nl[:n]


def sort_reads(reads, ids, min_cnt, median_cnt, avg_cnt):
    """std::sort with _sortRead::operator< (main.cpp:103-125): the permutation (order[j] = original index)."""
    n = len(reads)
    ra = (C.c_char_p * max(1, n))(*[r.encode() for r in reads])
    ia = (C.c_char_p * max(1, n))(*[i.encode() for i in ids])
    order = np.zeros(max(1, n), dtype=np.int64)
    lib().t4ref_sort_reads(ra, ia, np.ascontiguousarray(min_cnt, dtype=np.int32).ctypes.data, np.ascontiguousarray(median_cnt, dtype=np.int32).ctypes.data,
                           np.ascontiguousarray(avg_cnt, dtype=np.float32).ctypes.data, n, order.ctypes.data)
    return order[:n]


def assembled_list(ret, resc):
    """The driver's assembledReadIdx (main.cpp:1779, 1933): reads added in the main pass, in order, then the rescued ones."""
    ret = np.asarray(ret)
    main = np.nonzero(ret >= 0)[0]
    if resc is None:
        return main.astype(np.int32)
    resc = np.asarray(resc)
    rescued = np.nonzero((resc != np.iinfo(np.int32).min) & (resc >= 0))[0]
    return np.concatenate([main, rescued]).astype(np.int32)


def assign_pass(src, k, descs, pool, lst, strands, recompute=True):
    """main.cpp:2047-2118 on the reference SeqSet `src`: (extended RefSeqSet, assign int32[n, 8], similarity[n]) per list element."""
    descs = np.ascontiguousarray(descs)
    pool = np.ascontiguousarray(pool)
    lst = np.ascontiguousarray(lst, dtype=np.int32)
    strands = np.ascontiguousarray(strands, dtype=np.int8)
    n = len(lst)
    a = np.zeros((max(1, n), 8), dtype=np.int32)
    s = np.zeros(max(1, n), dtype=np.float64)
    h = lib().t4ref_assign_pass(src.h, k, descs.ctypes.data, pool.ctypes.data, lst.ctypes.data, n, strands.ctypes.data,
                                int(recompute), a.ctypes.data, s.ctypes.data)
    return RefSeqSet(k, handle=h), a[:n], s[:n]


def merge_sets(shards, k=9):
    """SURVEY.md 8e(2) merge oracle over reference SeqSets in (rank, stream) order -> (merged RefSeqSet, n removed)."""
    arr = (C.c_void_p * len(shards))(*[s.h for s in shards])
    removed = C.c_int()
    h = lib().t4ref_merge_sets(arr, len(shards), k, C.byref(removed))
    return RefSeqSet(k, handle=h), removed.value


def set_from_contigs(contigs, k=9):
    """Reference SeqSet holding the given contigs (dicts as produced by trust4_b200.dist.unpack_contigs)."""
    s = RefSeqSet(k)
    for c in contigs:
        pw = np.ascontiguousarray(c["pos_weight"], dtype=np.int32)
        lib().t4ref_input_contig(s.h, c["name"].encode(), c["consensus"].encode(), pw.ctypes.data, c["barcode"], c["num_read"])
    return s
