"""Seeded synthetic immune-repertoire reads and the stage-1 AddRead-loop workload.

Workload generator for tests and bench.py (host side, numpy only).  It follows
the generator spec of SURVEY.md Appendix A / section 8d (transcript = V + random
junction + J + first 250 bp of C, clone abundance ~ rank^-0.8, fragment
N(2L, L/3), 0.5 % substitutions) and then plays the role of the reference's
pre-processing (main.cpp:787-1526: k-mer statistics, sort, rough annotation) so
that the hot path receives what the reference's AddRead loop receives: one
`t4_read_desc` per loop iteration (include/trust4_b200.h).  Gene overlaps come
from the known transcript coordinates instead of refSet.AnnotateRead -- both the
GPU engine and the reference arm consume the identical records, which is the
parity / timing boundary of SURVEY.md section 8e(1).
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GENE_POOL = os.path.join(os.path.dirname(_HERE), "bench", "data", "gene_pool.tsv.gz")

CHAINS = ("IGH", "IGK", "IGL", "TRA", "TRB")
_NUC = np.frombuffer(b"ACGT", dtype=np.uint8)

# Mirrors `struct t4_read_desc` (64 bytes).
READ_DESC = np.dtype(
    [
        ("seq_off", "<u8"),
        ("len", "<i4"),
        ("barcode", "<i4"),
        ("min_cnt", "<i4"),
        ("min_kmer_count", "<i4"),
        ("sim_threshold", "<f8"),
        ("name_id", "<i4"),
        ("mate_idx", "<i4"),
        ("eq_lo", "<i4"),
        ("eq_hi", "<i4"),
        ("flags", "<u4"),
        ("strand_in", "i1"),
        ("novel_strand", "i1"),
        ("gene4", "S4"),
        ("pad_", "i1", (2,)),
    ],
    align=True,
)
assert READ_DESC.itemsize == 64, READ_DESC.itemsize

RD_DUP = 1 << 0
RD_FILTERED = 1 << 1
RD_NOVEL_ON_FAIL = 1 << 2
RD_MOTIF = 1 << 3
RD_GOOD_PLUS = 1 << 4
RD_GOOD_MINUS = 1 << 5
RD_MOTIF_FORCED = 1 << 6

RUN_CFG = np.dtype(
    [
        ("has_barcode", "<i4"),
        ("repetitive", "<i4"),
        ("change_k_threshold", "<i4"),
        ("update_consensus_every", "<i4"),
        ("do_rescue", "<i4"),
        ("first_read_len", "<i4"),
        ("final_update", "<i4"),
        ("release_barcodes", "<i4"),
        ("contig_min_cov", "<i4"),
        ("reserved_", "<i4"),
    ]
)


def run_cfg(has_barcode=0, repetitive=0, change_k_threshold=4096, update_consensus_every=10000, do_rescue=1,
            first_read_len=150, final_update=1, release_barcodes=0, contig_min_cov=0):
    c = np.zeros(1, dtype=RUN_CFG)
    c["has_barcode"] = has_barcode
    c["repetitive"] = repetitive
    c["change_k_threshold"] = change_k_threshold
    c["update_consensus_every"] = update_consensus_every
    c["do_rescue"] = do_rescue
    c["first_read_len"] = first_read_len
    c["final_update"] = final_update
    c["release_barcodes"] = release_barcodes
    c["contig_min_cov"] = contig_min_cov
    return c


def encode(seq: str) -> np.ndarray:
    """ASCII ACGT -> codes 0..3."""
    a = np.frombuffer(seq.encode(), dtype=np.uint8)
    lut = np.zeros(256, dtype=np.uint8)
    lut[ord("C")] = 1
    lut[ord("G")] = 2
    lut[ord("T")] = 3
    return lut[a]


def decode(codes: np.ndarray) -> str:
    return _NUC[codes].tobytes().decode()


def load_gene_pool(path: str = GENE_POOL):
    pool = {c: {"V": [], "J": [], "C": []} for c in CHAINS}
    with gzip.open(path, "rt") as f:
        for line in f:
            chain, seg, name, seq = line.rstrip("\n").split("\t")
            pool[chain][seg].append((name, seq))
    return pool


@dataclass
class Clones:
    seq: np.ndarray        # concatenated transcript codes
    off: np.ndarray        # [nclones+1] offsets
    seg_end: np.ndarray    # [nclones,4] transcript coordinate one past V, junction, J, C
    gene_name: list        # [nclones] (V, J, C) full names
    gene_len: np.ndarray   # [nclones,3] full germline length of V, J, C
    j_trim: np.ndarray     # [nclones] 5' trim of J


def make_clones(nclones: int, seed: int, pool=None, chains=CHAINS) -> Clones:
    pool = pool or load_gene_pool()
    rng = np.random.default_rng([seed, 0xC10E])
    seqs, seg_end, names, glen, jtrim = [], [], [], [], []
    for _ in range(nclones):
        ch = chains[rng.integers(len(chains))]
        vn, vs = pool[ch]["V"][rng.integers(len(pool[ch]["V"]))]
        jn, js = pool[ch]["J"][rng.integers(len(pool[ch]["J"]))]
        cn, cs = pool[ch]["C"][rng.integers(len(pool[ch]["C"]))]
        v = vs[: len(vs) - int(rng.integers(0, 9))]
        ins = decode(rng.integers(0, 4, size=int(rng.integers(2, 19))).astype(np.uint8))
        jt = int(rng.integers(0, 7))
        j = js[jt:]
        c = cs[:250]
        t = v + ins + j + c
        seqs.append(encode(t))
        e1 = len(v)
        e2 = e1 + len(ins)
        e3 = e2 + len(j)
        seg_end.append((e1, e2, e3, e3 + len(c)))
        names.append((vn, jn, cn))
        glen.append((len(vs), len(js), len(cs)))
        jtrim.append(jt)
    off = np.zeros(nclones + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return Clones(np.concatenate(seqs), off, np.array(seg_end, dtype=np.int32), names,
                  np.array(glen, dtype=np.int32), np.array(jtrim, dtype=np.int32))


@dataclass
class Reads:
    codes: np.ndarray     # [n, L] uint8 0..3 (as sequenced, i.e. possibly reverse strand)
    clone: np.ndarray     # [n]
    tstart: np.ndarray    # [n] transcript coordinate of the leftmost base on the transcript strand
    strand: np.ndarray    # [n] +1 read equals transcript strand, -1 reverse complement
    pair: np.ndarray      # [n] pair id (mates share it)
    L: int


def sample_pairs(clones: Clones, npairs: int, L: int, seed: int, alpha: float = 0.8, sub_rate: float = 0.005,
                 paired: bool = True) -> Reads:
    rng = np.random.default_rng([seed, 0x5EAD])
    nclones = len(clones.off) - 1
    w = 1.0 / np.power(np.arange(1, nclones + 1, dtype=np.float64), alpha)
    w /= w.sum()
    cl = rng.choice(nclones, size=npairs, p=w)
    tlen = (clones.off[1:] - clones.off[:-1])[cl]
    ins = np.clip(rng.normal(2 * L, L / 3.0, size=npairs).astype(np.int64), L, tlen)
    start = (rng.random(npairs) * (tlen - ins + 1)).astype(np.int64)
    base = clones.off[cl] + start
    ar = np.arange(L, dtype=np.int64)
    r1 = clones.seq[base[:, None] + ar[None, :]]                       # forward, left end of fragment
    r2 = 3 - clones.seq[(base + ins - 1)[:, None] - ar[None, :]]       # revcomp, right end of fragment
    t1, t2 = start, start + ins - L
    s1 = np.ones(npairs, dtype=np.int8)
    s2 = -np.ones(npairs, dtype=np.int8)
    if paired:
        swap = rng.random(npairs) < 0.5
        a = np.where(swap[:, None], r2, r1)
        b = np.where(swap[:, None], r1, r2)
        ta, tb = np.where(swap, t2, t1), np.where(swap, t1, t2)
        sa, sb = np.where(swap, s2, s1), np.where(swap, s1, s2)
        codes = np.empty((2 * npairs, L), dtype=np.uint8)
        codes[0::2], codes[1::2] = a, b
        tstart = np.empty(2 * npairs, dtype=np.int64)
        tstart[0::2], tstart[1::2] = ta, tb
        strand = np.empty(2 * npairs, dtype=np.int8)
        strand[0::2], strand[1::2] = sa, sb
        clone = np.repeat(cl, 2)
        pair = np.repeat(np.arange(npairs, dtype=np.int64), 2)
    else:
        codes, tstart, strand, clone, pair = r1, t1, s1, cl, np.arange(npairs, dtype=np.int64)
    err = rng.random(codes.shape) < sub_rate
    sub = rng.integers(0, 4, size=codes.shape, dtype=np.uint8)
    codes = np.where(err, sub, codes).astype(np.uint8)
    return Reads(codes, clone.astype(np.int64), tstart, strand, pair, L)


def sample_single_cell(clones: Clones, n_barcodes: int, reads_per_barcode: int, L: int, seed: int, ambient: float = 0.02,
                       sub_rate: float = 0.005):
    """10x-style single-cell reads (SURVEY.md 8d config 4 / BASELINE configs[3]): every cell (barcode) expresses one or
    two clonotypes (e.g. a heavy/beta and a light/alpha chain), `reads_per_barcode` single-end reads each, plus a
    fraction of ambient reads drawn from the clones of other cells.  Returns (Reads, barcode[n])."""
    rng = np.random.default_rng([seed, 0x10C])
    nclones = len(clones.off) - 1
    n = n_barcodes * reads_per_barcode
    bc = np.repeat(np.arange(n_barcodes, dtype=np.int64), reads_per_barcode)
    c1 = rng.integers(0, nclones, size=n_barcodes)
    c2 = rng.integers(0, nclones, size=n_barcodes)
    two = rng.random(n_barcodes) < 0.7
    pick2 = two[bc] & (rng.random(n) < 0.4)
    cl = np.where(pick2, c2[bc], c1[bc])
    amb = rng.random(n) < ambient
    cl = np.where(amb, rng.integers(0, nclones, size=n), cl)
    tlen = (clones.off[1:] - clones.off[:-1])[cl]
    start = (rng.random(n) * (tlen - L + 1)).astype(np.int64)
    base = clones.off[cl] + start
    ar = np.arange(L, dtype=np.int64)
    fwd = clones.seq[base[:, None] + ar[None, :]]
    rev = 3 - clones.seq[(base + L - 1)[:, None] - ar[None, :]]
    minus = rng.random(n) < 0.5
    codes = np.where(minus[:, None], rev, fwd)
    err = rng.random(codes.shape) < sub_rate
    sub = rng.integers(0, 4, size=codes.shape, dtype=np.uint8)
    codes = np.where(err, sub, codes).astype(np.uint8)
    strand = np.where(minus, -1, 1).astype(np.int8)
    return Reads(codes, cl.astype(np.int64), start, strand, np.arange(n, dtype=np.int64), L), bc


def sample_amplicon(clones: Clones, n: int, L: int, seed: int, alpha: float = 1.0, primer_in_c: int = 20, sub_rate: float = 0.005) -> Reads:
    """Bulk TCR-seq amplicon reads (SURVEY.md 8d config 5 / BASELINE configs[4]): every read starts at the C-gene primer
    (`primer_in_c` bases into the constant region) and runs antisense across J and the CDR3 into V; clonotype
    abundance follows rank^-alpha."""
    rng = np.random.default_rng([seed, 0xA3B])
    nclones = len(clones.off) - 1
    w = 1.0 / np.power(np.arange(1, nclones + 1, dtype=np.float64), alpha)
    w /= w.sum()
    cl = rng.choice(nclones, size=n, p=w)
    end = clones.seg_end[cl, 2].astype(np.int64) + primer_in_c          # exclusive transcript coordinate of the primer end
    start = np.maximum(0, end - L)
    ar = np.arange(L, dtype=np.int64)
    codes = 3 - clones.seq[(clones.off[cl] + start + L - 1)[:, None] - ar[None, :]]
    err = rng.random(codes.shape) < sub_rate
    sub = rng.integers(0, 4, size=codes.shape, dtype=np.uint8)
    codes = np.where(err, sub, codes).astype(np.uint8)
    return Reads(codes, cl.astype(np.int64), start, -np.ones(n, dtype=np.int8), np.arange(n, dtype=np.int64), L)


def write_fastq(reads: Reads, prefix: str):
    """Write <prefix>_1.fq / _2.fq (mates interleaved in `reads`) for the reference binary."""
    L = reads.L
    qual = "I" * L
    with open(prefix + "_1.fq", "w") as f1, open(prefix + "_2.fq", "w") as f2:
        for i in range(0, reads.codes.shape[0], 2):
            f1.write("@r%d\n%s\n+\n%s\n" % (i // 2, decode(reads.codes[i]), qual))
            f2.write("@r%d\n%s\n+\n%s\n" % (i // 2, decode(reads.codes[i + 1]), qual))


# ---------------------------------------------------------------------------
# pre-processing stand-in: k-mer statistics, sort, annotation from truth
# ---------------------------------------------------------------------------
def kmer_stats(codes: np.ndarray, k: int = 21, device=None, salt=None):
    """Canonical k-mer counts over all reads -> per-read (min, median, mean) like
    KmerCount::GetCountStatsAndTrim (KmerCount.hpp:177) without trimming.  `device`: a torch device
    to do the counting on (workload preparation only -- 260 M k-mers for 1 M pairs)."""
    n, L = codes.shape
    m = L - k + 1
    if device is not None:
        import torch
        c = torch.from_numpy(np.ascontiguousarray(codes)).to(device).to(torch.int64)
        fw = torch.zeros((n, m), dtype=torch.int64, device=device)
        rc = torch.zeros((n, m), dtype=torch.int64, device=device)
        for j in range(k):
            fw = (fw << 2) | c[:, j:j + m]
            rc = rc | ((3 - c[:, j:j + m]) << (2 * j))
        canon = torch.minimum(fw, rc)
        if salt is not None:      # per-read salt above the 2k code bits: counts per (k-mer, cell) -- the barcode-wise KmerCount of main.cpp:1128-1150
            canon = canon | (torch.from_numpy(np.ascontiguousarray(salt, dtype=np.int64)).to(device)[:, None] << (2 * k))
        canon = canon.reshape(-1)
        del fw, rc, c
        uniq, inv, cnt = torch.unique(canon, return_inverse=True, return_counts=True)
        del canon, uniq
        per = cnt[inv].reshape(n, m)
        del inv, cnt
        per_sorted, _ = torch.sort(per, dim=1)
        mn = per_sorted[:, 0].to(torch.int32).cpu().numpy()
        med = per_sorted[:, m // 2].to(torch.int32).cpu().numpy()
        avg = per.to(torch.float64).mean(dim=1).to(torch.float32).cpu().numpy()
        return mn, med, avg
    c64 = codes.astype(np.uint64)
    fw = np.zeros((n, m), dtype=np.uint64)
    rc = np.zeros((n, m), dtype=np.uint64)
    for j in range(k):
        fw = (fw << np.uint64(2)) | c64[:, j:j + m]
        rc = rc | ((np.uint64(3) - c64[:, j:j + m]) << np.uint64(2 * j))
    canon = np.minimum(fw, rc)
    if salt is not None:
        canon = canon | (np.asarray(salt, dtype=np.uint64)[:, None] << np.uint64(2 * k))
    canon = canon.ravel()
    uniq, inv, cnt = np.unique(canon, return_inverse=True, return_counts=True)
    per = cnt[inv].reshape(n, m)
    per_sorted = np.sort(per, axis=1)
    mn = per_sorted[:, 0].astype(np.int32)
    med = per_sorted[:, m // 2].astype(np.int32)
    avg = per.mean(axis=1).astype(np.float32)
    return mn, med, avg


_AA = None


def _aa_table():
    global _AA
    if _AA is None:
        # standard code on codes A=0,C=1,G=2,T=3 (stop = '_'), as SeqSet::DnaToAa (SeqSet.hpp:638)
        aa = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV_Y_YSSSS_CWCLFLF"
        _AA = np.frombuffer(aa.encode(), dtype=np.uint8)
    return _AA


def has_motif(codes: np.ndarray) -> np.ndarray:
    """SeqSet::HasMotif(read, +-1) != 0 (SeqSet.hpp:5029): any-frame YYC or [FW]G.G."""
    n, L = codes.shape
    idx = codes[:, :-2].astype(np.int32) * 16 + codes[:, 1:-1].astype(np.int32) * 4 + codes[:, 2:].astype(np.int32)
    aa = _aa_table()[idx]                       # [n, L-2], aa[:, i] = codon starting at base i
    Y, C, F, W, G = (aa == ord(ch) for ch in "YCFWG")
    m = aa.shape[1]
    yyc = (Y[:, : m - 6] & Y[:, 3: m - 3] & C[:, 6:]).any(axis=1) if m > 6 else np.zeros(n, bool)
    fg = ((F | W)[:, : m - 9] & G[:, 3: m - 6] & G[:, 9:]).any(axis=1) if m > 9 else np.zeros(n, bool)
    return yyc | fg


@dataclass
class Workload:
    descs: np.ndarray      # READ_DESC[n]
    pool: np.ndarray       # uint8 ASCII read pool
    names: list            # gene names (bytes)
    L: int
    order: np.ndarray      # index into the unsorted reads for each desc
    med_cnt: np.ndarray = None   # medianCnt of the 21-mer statistics per record (cost model of the sharding only)

    def read(self, i: int) -> str:
        d = self.descs[i]
        return self.pool[int(d["seq_off"]): int(d["seq_off"]) + int(d["len"])].tobytes().decode()


def build_workload(clones: Clones, reads: Reads, min_anchor: int = 17, device=None, barcode=None, repseq: bool = False) -> Workload:
    """Truth-annotated stand-in for main.cpp:981-1526 -> AddRead-loop records.

    barcode: per-read cell barcode (10x mode, BASELINE configs[3]): 21-mer statistics are taken per cell (the stand-in
             uses them for both minCnt and barcodeMinCnt), records are ordered by CompReadWithBarcode (main.cpp:128:
             barcode, then barcodeMinCnt desc, then the default order), duplicates need equal barcodes (main.cpp:1596),
             minKmerCount = (minCnt + barcodeMinCnt + 1) / 2 and similarityThreshold = 0.9 (main.cpp:1694, 1701).
    repseq:  --repseq = --trimLevel 2 (BASELINE configs[4]): the V gene assignment becomes a pseudo barcode
             (main.cpp:1224-1235, no re-sort, index not salted) and similarityThreshold = 0.9 (main.cpp:1693)."""
    n, L = reads.codes.shape
    if barcode is not None:
        barcode = np.asarray(barcode, dtype=np.int64)
    mn, med, avg = kmer_stats(reads.codes, device=device, salt=barcode)
    # sort: minCnt desc, medianCnt desc, avgCnt desc, len desc, read asc, id asc  (main.cpp:103-125)
    words = []
    nw = (L + 31) // 32
    padded = np.zeros((n, nw * 32), dtype=np.uint64)
    padded[:, :L] = reads.codes
    for wi in range(nw):
        wv = np.zeros(n, dtype=np.uint64)
        for j in range(32):
            wv = (wv << np.uint64(2)) | padded[:, wi * 32 + j]
        words.append(wv)
    keys = [np.arange(n)] + words[::-1] + [-avg.astype(np.float64), -med.astype(np.int64), -mn.astype(np.int64)]
    if barcode is not None:
        keys.append(barcode)
    order = np.lexsort(keys)
    bc_sorted = barcode[order] if barcode is not None else None
    codes = reads.codes[order]
    cl = reads.clone[order]
    ts = reads.tstart[order]
    st = reads.strand[order]
    pr = reads.pair[order]
    mn, med, avg = mn[order], med[order], avg[order]

    descs = np.zeros(n, dtype=READ_DESC)
    descs["seq_off"] = np.arange(n, dtype=np.uint64) * np.uint64(L)
    descs["len"] = L
    descs["barcode"] = -1
    descs["min_cnt"] = mn
    descs["min_kmer_count"] = mn
    pool = _NUC[codes].reshape(-1)

    same_prev = np.zeros(n, dtype=bool)
    same_prev[1:] = (codes[1:] == codes[:-1]).all(axis=1)
    if bc_sorted is not None:
        same_prev[1:] &= bc_sorted[1:] == bc_sorted[:-1]
        descs["barcode"] = bc_sorted.astype(np.int32)
    flags = np.zeros(n, dtype=np.uint32)
    flags[same_prev] |= RD_DUP
    run_id = np.cumsum(~same_prev) - 1
    run_lo = np.flatnonzero(~same_prev)
    run_hi = np.append(run_lo[1:], n)
    descs["eq_lo"] = run_lo[run_id]
    descs["eq_hi"] = run_hi[run_id]

    # mate index: the other read of the pair
    by_pair = np.argsort(pr, kind="stable")
    mate = np.full(n, -1, dtype=np.int32)
    pp = pr[by_pair]
    is_pair = np.zeros(n, dtype=bool)
    is_pair[:-1] = pp[:-1] == pp[1:]
    a = by_pair[:-1][is_pair[:-1]]
    b = by_pair[1:][is_pair[:-1]]
    mate[a] = b
    mate[b] = a
    descs["mate_idx"] = mate

    # gene overlaps from truth, in transcript-oriented read coordinates
    seg_end = clones.seg_end[cl].astype(np.int64)            # [n,4]
    seg_start = np.zeros_like(seg_end)
    seg_start[:, 1:] = seg_end[:, :-1]
    te = ts + L                                               # exclusive
    lo = np.maximum(ts[:, None], seg_start)
    hi = np.minimum(te[:, None], seg_end)
    ov = (hi - lo)                                            # [n,4] V, junction, J, C
    present = ov >= min_anchor
    present[:, 1] = False                                      # no D annotation in the synthetic stream
    g_rs = lo - ts[:, None]
    g_re = hi - 1 - ts[:, None]
    g_ss = lo - seg_start
    g_ss[:, 2] += clones.j_trim[cl]
    g_se = g_ss + ov - 1
    V, J, C = 0, 2, 3
    pv, pj, pc = present[:, V], present[:, J], present[:, C]

    # main.cpp:1640-1651: reads from the constant gene only
    c_only = pc & ~pv & ~pj
    filt = c_only & (g_ss[:, C] >= 200)
    filt |= c_only & (g_ss[:, C] >= 100) & ((st == 1) | (ov[:, C] < L))
    flags[filt & ~same_prev] |= RD_FILTERED

    # name / strand passed to AddRead (main.cpp:1660-1673): last annotated gene's 4-char prefix
    name_ids = {}
    names = []

    def nid(s):
        if s not in name_ids:
            name_ids[s] = len(names)
            names.append(s.encode())
        return name_ids[s]

    gname = clones.gene_name
    last = np.where(pc, C, np.where(pj, J, np.where(pv, V, -1)))
    first = np.where(pv, V, np.where(pj, J, np.where(pc, C, -1)))
    col = {V: 0, J: 1, C: 2}
    gene4 = np.zeros(n, dtype="S4")
    name_id = np.full(n, -1, dtype=np.int32)
    for seg in (V, J, C):
        idx = np.flatnonzero(last == seg)
        if len(idx):
            gene4[idx] = [gname[c][col[seg]][:4].encode() for c in cl[idx]]
        idx = np.flatnonzero(first == seg)
        if len(idx):
            name_id[idx] = [nid(gname[c][col[seg]]) for c in cl[idx]]
    descs["gene4"] = gene4
    annotated = last >= 0
    descs["strand_in"] = np.where(annotated, st, 0)
    descs["novel_strand"] = np.where(annotated, st, 0)
    descs["name_id"] = name_id

    thr = np.full(n, 0.9)
    thr[mn >= 2] = 0.95
    thr[mn >= 20] = 0.97
    tcr = np.array([g[:1] == b"T" for g in gene4])
    thr[tcr & (thr < 0.95)] = 0.95
    if barcode is not None or repseq:
        thr[:] = 0.9                                           # main.cpp:1692-1694
    descs["sim_threshold"] = thr
    if repseq:
        # pseudo barcode = index of the V gene in the reference set (any injective numbering of the V names does)
        vid = {}
        pb = np.full(n, -1, dtype=np.int32)
        idx = np.flatnonzero(pv)
        pb[idx] = [vid.setdefault(gname[c][0], len(vid)) for c in cl[idx]]
        descs["barcode"] = pb

    # main.cpp:1706-1736: may the read seed a new contig when AddRead fails?
    match_half = (np.where(present, ov, 0)).sum(axis=1)
    vlen = clones.gene_len[cl, 0]
    ok = match_half >= 31
    ok |= pv & pj & (g_re[:, V] < g_rs[:, J])
    ok |= pv & (g_se[:, V] >= vlen - 17)
    ok |= ~pv & pj & (g_ss[:, J] <= 17)
    flags[ok & annotated] |= RD_NOVEL_ON_FAIL

    # main.cpp:1782-1808 (all similarities >= 0.9 for truth annotation)
    span_plus = (pj & (g_rs[:, J] > g_re[:, V])) | (pc & (g_rs[:, C] > g_re[:, V]))
    flags[pv & ~span_plus] |= RD_GOOD_PLUS
    span_minus = pv & ((pj & (g_rs[:, J] > g_re[:, V])) | (pc & (g_rs[:, C] > g_re[:, V])))
    flags[(pj | pc) & ~span_minus] |= RD_GOOD_MINUS
    flags[has_motif(codes)] |= RD_MOTIF

    # duplicates inherit the static fields of the first read of their run (main.cpp:1588 static geneOverlap)
    first_of_run = run_lo[run_id]
    keep = RD_GOOD_PLUS | RD_GOOD_MINUS
    flags = np.where(same_prev, (flags & ~np.uint32(keep)) | (flags[first_of_run] & np.uint32(keep)), flags)
    descs["flags"] = flags
    return Workload(descs, pool, names, L, order, med.astype(np.int32))


# Cost model of one record in the stream kernel, used to size contiguous shards (microseconds at 1965 MHz, only the
# ratios matter).  A duplicate is a RepeatAddRead (posWeight increments only); a distinct read runs the whole AddRead
# path, and what it costs depends on where it sits in the abundance spectrum, which the 21-mer statistics give before
# the launch: minCnt (the rarest k-mer) says whether the read is error free, medianCnt how abundant its clonotype is.
# An erroneous copy of an abundant clonotype (minCnt 1, medianCnt > 65536) is the most expensive kind: it overlaps
# many contigs of its shard and most of its overhangs need the banded DP.  Table fitted by non-negative least squares
# to the per-stream cycle counts of the default bench workload (bench.py --dump-streams, bench/fit_cost_model.py;
# correlation 0.93 with the measured stream times; geometric mean of the fits on equal-count and on cost-balanced shards,
# because the cost of a record also grows with the size of its shard), smoothed where a cell has too few reads.
COST_DUP = 2.5
COST_MIN_EDGES = (1, 2, 4, 8, 16, 64, 256, 1 << 40)
COST_MED_EDGES = (4, 16, 64, 256, 1024, 4096, 16384, 65536, 1 << 40)
COST_TABLE = np.array([[306, 228, 305, 301, 205, 276, 454, 464, 630], [314, 182, 285, 316, 174, 209, 334, 361, 581],
                       [295, 180, 302, 406, 192, 206, 271, 339, 565], [284, 284, 280, 270, 273, 236, 285, 399, 586],
                       [332, 332, 200, 377, 279, 293, 324, 380, 686], [256, 256, 256, 412, 132, 234, 335, 238, 336],
                       [215, 215, 215, 205, 127, 164, 279, 185, 274], [117, 117, 117, 117, 173, 117, 112, 114, 124]], dtype=np.float64)


def read_cost(descs, med_cnt=None) -> np.ndarray:
    mc = descs["min_cnt"].astype(np.int64)
    med = np.asarray(med_cnt, dtype=np.int64) if med_cnt is not None else mc
    c = COST_TABLE[np.searchsorted(np.array(COST_MIN_EDGES), mc, side="left"), np.searchsorted(np.array(COST_MED_EDGES), med, side="left")].copy()
    c[(descs["flags"] & RD_DUP) != 0] = COST_DUP
    return c


def _gene_shards(w: Workload, n_shards: int) -> np.ndarray:
    """group="gene": stream of every record when reads are first grouped by the gene of their rough annotation (the name
    InputNovelRead would give them, i.e. mostly the V gene) and the groups are then packed into n_shards streams of equal
    predicted cost: with t the smallest cap for which everything fits, a group dearer than t is cut into ceil(cost / t)
    contiguous sub-blocks (at run boundaries), the others are bin-packed (largest first into the least loaded stream).  Reads of one clonotype that reach into its V
    gene then meet in the same SeqSet whatever their abundance rank -- the reference itself shards --repseq data by V gene
    (pseudo barcodes, main.cpp:1224-1235)."""
    d = w.descs
    n = len(d)
    head = d["eq_lo"].astype(np.int64)
    key = d["name_id"][head].astype(np.int64)                    # the whole run shares the annotation of its first record
    cost = read_cost(d, w.med_cnt)
    keys, inv = np.unique(key, return_inverse=True)
    gcost = np.bincount(inv, weights=cost, minlength=len(keys))
    shard_of = np.full(n, -1, dtype=np.int64)

    def budget(t):
        """streams needed when no stream may be dearer than t: ceil(cost / t) per big group + the bin-packed rest"""
        bg = np.flatnonzero(gcost > t)
        kb = np.ceil(gcost[bg] / t).astype(np.int64)
        rest = gcost.sum() - gcost[bg].sum()
        return bg, kb, (max(1, int(np.ceil(1.03 * rest / t))) if len(bg) < len(keys) else 0)

    lo = cost.sum() / max(1, n_shards)                            # the smallest cap t whose budget fits n_shards streams
    hi = 2.0 * lo + gcost.max() / max(1, n_shards)
    while True:
        bg, kb, bins = budget(hi)
        if kb.sum() + bins <= n_shards:
            break
        hi *= 2.0
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        bg, kb, bins = budget(mid)
        if kb.sum() + bins <= n_shards:
            hi = mid
        else:
            lo = mid
    big, k_big, bins = budget(hi)
    order = np.argsort(-gcost[big])
    big, k_big = big[order], k_big[order]
    small = np.setdiff1d(np.arange(len(keys)), big)
    small = small[np.argsort(-gcost[small])]
    nxt = 0
    for g, k in zip(big, k_big):
        idx = np.flatnonzero(inv == g)                            # ascending = sorted order
        cum = np.cumsum(cost[idx])
        cuts = np.searchsorted(cum, cum[-1] * np.arange(1, k) / k)
        part = np.zeros(len(idx), dtype=np.int64)
        for c in cuts:
            c = int(np.searchsorted(idx, head[idx[min(c, len(idx) - 1)]]))   # move back to the start of the run
            part[c:] += 1
        _, part = np.unique(part, return_inverse=True)
        shard_of[idx] = nxt + part
        nxt += int(part.max()) + 1
    load = np.zeros(max(1, bins))
    for g in small:                                               # descending cost: largest first into the least loaded stream
        b = int(np.argmin(load))
        load[b] += gcost[g]
        shard_of[inv == g] = nxt + b
    _, shard_of = np.unique(shard_of, return_inverse=True)        # drop empty streams, keep ids dense
    return shard_of.astype(np.int64)


def shard_workload(w: Workload, n_shards: int, deal: bool = False, balance: str = "reads", align: str = "run", group: str = ""):
    """Shard the sorted records into n_shards independent streams.  Returns desc_off[n_shards+1]
    and the records in stream order with mate_idx / eq_* made stream-relative (mates in other streams -> -1).
    A run of identical reads is never split (RepeatAddRead semantics survive), and every stream keeps the global
    sorted order of its reads.

    deal=False: contiguous blocks of the sorted list; balance = "reads" (equal read counts) or "cost" (equal predicted
                cost, read_cost()); align = "run" (never split a run of identical reads) or "barcode" (never split a
                barcode: 10x mode, barcodes are independent assemblies).
    group="gene": runs are grouped by the gene of their rough annotation and the groups packed into streams of equal
                predicted cost (_gene_shards); the number of streams returned may be smaller than n_shards.
    deal=True : runs are dealt round-robin (run r -> stream r mod n_shards).  Every stream then sees a uniform sample
                of the library instead of one abundance class, which equalises the work per stream (contiguous blocks
                of low-abundance reads are ~50x more expensive than blocks of duplicates)."""
    n = len(w.descs)
    d = w.descs.copy()
    if deal or group == "gene":
        same_prev = (d["flags"] & RD_DUP) != 0
        # a run = maximal range of identical read strings (eq_lo..eq_hi), which contains its DUP records
        run_id = np.cumsum(d["eq_lo"] == np.arange(n)) - 1
        if group == "gene":
            shard_of = _gene_shards(w, n_shards)
            n_shards = int(shard_of.max()) + 1 if n else n_shards
        else:
            shard_of = (run_id % n_shards).astype(np.int64)
        order = np.argsort(shard_of, kind="stable")
        counts = np.bincount(shard_of, minlength=n_shards)
        off = np.zeros(n_shards + 1, dtype=np.int64)
        off[1:] = np.cumsum(counts)
        new_pos = np.empty(n, dtype=np.int64)
        new_pos[order] = np.arange(n)
        base_new = off[shard_of]                       # stream start in the new array, per old index
        mate = d["mate_idx"].astype(np.int64)
        has = mate >= 0
        same = np.zeros(n, dtype=bool)
        same[has] = shard_of[mate[has]] == shard_of[has]
        rel_mate = np.full(n, -1, dtype=np.int64)
        rel_mate[same] = new_pos[mate[same]] - base_new[same]
        d["mate_idx"] = rel_mate
        # a run stays contiguous inside its stream: its new range starts at new_pos[eq_lo]
        lo_new = new_pos[d["eq_lo"].astype(np.int64)]
        length = d["eq_hi"].astype(np.int64) - d["eq_lo"].astype(np.int64)
        d["eq_lo"] = lo_new - base_new
        d["eq_hi"] = lo_new - base_new + length
        del same_prev
        return off, d[order]
    bounds = [0]
    if balance == "cost" and n > 0:
        # balance="cost": the same contiguous blocks of the sorted list, cut where the cumulative predicted cost crosses
        # s / n_shards of the total instead of where the read count does (a stream is a serial chain: the launch lasts as
        # long as its most expensive shard)
        cum = np.cumsum(read_cost(w.descs, w.med_cnt))
        cuts = np.searchsorted(cum, cum[-1] * np.arange(1, n_shards) / n_shards, side="left")
    else:
        cuts = (n * np.arange(1, n_shards)) // n_shards
    if align == "barcode":
        # whole barcodes per stream: a cut moves back to the first record of its barcode
        bcs = w.descs["barcode"]
        first_of_bc = np.flatnonzero(np.r_[True, bcs[1:] != bcs[:-1]])
    for b in cuts:
        if b >= n:
            b = n
        elif align == "barcode":
            b = int(first_of_bc[np.searchsorted(first_of_bc, b, side="right") - 1])
        else:
            b = int(w.descs["eq_lo"][b])
        bounds.append(max(b, bounds[-1]))
    bounds.append(n)
    off = np.array(bounds, dtype=np.int64)
    shard_of = np.searchsorted(off, np.arange(n), side="right") - 1
    base = off[shard_of]
    mate = d["mate_idx"].astype(np.int64)
    has = mate >= 0
    same = np.zeros(n, dtype=bool)
    same[has] = shard_of[mate[has]] == shard_of[has]
    d["mate_idx"] = np.where(same, mate - base, -1)
    d["eq_lo"] = d["eq_lo"] - base
    d["eq_hi"] = d["eq_hi"] - base
    return off, d
