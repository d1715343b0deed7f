#!/usr/bin/env python3
"""Assembly quality of a (read-sharded) run on synthetic data: which clonotypes were recovered?

A clonotype counts as *coverable* when at least `min_reads` of its reads span the 24-mer centred on its V(D)J junction
(the CDR3 core: last V bases + random insert + first J bases), and as *recovered* when some contig of the run contains
that 24-mer (either strand).  Read-sharding trades assembly contiguity for parallelism; this number says
how much of the repertoire survives the trade: it is reported for S streams next to S = 1.

Used by bench.py --quality (on the packed contigs of the timed run) and by tests."""
import numpy as np

SIG = 24


def _codes_to_kmers(codes, k=SIG):
    """All k-mers of a 1-D code array as uint64 (k <= 32); returns an empty array when too short."""
    n = len(codes) - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    c = codes.astype(np.uint64)
    out = np.zeros(n, dtype=np.uint64)
    for j in range(k):
        out = (out << np.uint64(2)) | c[j:j + n]
    return out


def junction_signatures(clones):
    """(forward, reverse-complement) 24-mer codes of every clonotype's junction centre, and the centre coordinate."""
    ncl = len(clones.off) - 1
    mid = (clones.seg_end[:, 0].astype(np.int64) + clones.seg_end[:, 1].astype(np.int64)) // 2
    lo = np.maximum(0, mid - SIG // 2)
    fw = np.zeros(ncl, dtype=np.uint64)
    rc = np.zeros(ncl, dtype=np.uint64)
    for j in range(SIG):
        b = clones.seq[clones.off[:-1] + lo + j].astype(np.uint64)
        fw = (fw << np.uint64(2)) | b
        rc = rc | ((np.uint64(3) - b) << np.uint64(2 * j))
    return fw, rc, lo


def coverable(clones, reads, lo, min_reads=2):
    """Clonotypes with >= min_reads reads spanning [lo, lo + 24)."""
    cl = reads.clone
    span = (reads.tstart <= lo[cl]) & (reads.tstart + reads.L >= lo[cl] + SIG)
    cnt = np.bincount(cl[span], minlength=len(clones.off) - 1)
    return cnt >= min_reads


def recovered_fraction(clones, reads, contig_codes_concat, contig_off, min_reads=2):
    """contig_codes_concat: uint8 codes (0..3, 4 = N) of all contigs back to back; contig_off: their offsets."""
    fw, rc, lo = junction_signatures(clones)
    cov = coverable(clones, reads, lo, min_reads)
    sig = np.unique(np.concatenate([fw, rc]))
    c = contig_codes_concat
    n = len(c) - SIG + 1
    found = np.zeros(0, dtype=np.uint64)
    if n > 0:
        km = _codes_to_kmers(np.minimum(c, 3))
        ok = np.ones(n, dtype=bool)
        # windows crossing a contig boundary or holding an N do not count
        bad = np.zeros(len(c) + 1, dtype=np.int64)
        bad[1:] = np.cumsum(c > 3)
        ok &= (bad[SIG:SIG + n] - bad[:n]) == 0
        starts = np.asarray(contig_off[1:-1], dtype=np.int64)
        for s in starts:
            ok[max(0, s - SIG + 1):min(n, s)] = False
        km = km[ok]
        pos = np.searchsorted(sig, km)
        pos[pos >= len(sig)] = len(sig) - 1
        found = np.unique(km[sig[pos] == km])
    rec = np.isin(fw, found) | np.isin(rc, found)
    n_cov = int(cov.sum())
    return {"clonotypes": int(len(fw)), "coverable": n_cov, "recovered": int((rec & cov).sum()),
            "recovered_fraction": float((rec & cov).sum() / max(1, n_cov)), "signature": "%d-mer centred on the V(D)J junction" % SIG,
            "min_spanning_reads": min_reads}


def _kmer_at(clones, pos):
    """24-mer codes (forward, reverse complement) starting at transcript coordinate pos[c] of clone c."""
    fw = np.zeros(len(pos), dtype=np.uint64)
    rc = np.zeros(len(pos), dtype=np.uint64)
    for j in range(SIG):
        b = clones.seq[clones.off[:-1] + pos + j].astype(np.uint64)
        fw = (fw << np.uint64(2)) | b
        rc = rc | ((np.uint64(3) - b) << np.uint64(2 * j))
    return fw, rc


def spanning_fraction(clones, reads, contig_codes_concat, contig_off, v_flank=90, j_flank=20, min_reads=2):
    """Contiguity: a clonotype is *spanned* when ONE contig contains its V(D)J core from `v_flank` bases inside V to
    `j_flank` bases past the end of J (~200 bp, more than a read: it takes real assembly), tested through the 24-mers at
    both ends of that window lying in the same contig at the right distance (either strand).  Denominator: clonotypes
    whose window is covered by reads at all (every base of the window under >= min_reads reads)."""
    ncl = len(clones.off) - 1
    tlen = (clones.off[1:] - clones.off[:-1]).astype(np.int64)
    a = np.maximum(0, clones.seg_end[:, 0].astype(np.int64) - v_flank)
    b = np.minimum(tlen, clones.seg_end[:, 2].astype(np.int64) + j_flank)       # exclusive
    dist = b - SIG - a
    A_fw, A_rc = _kmer_at(clones, a)
    B_fw, B_rc = _kmer_at(clones, b - SIG)
    # per-base read depth over each window (difference arrays on the concatenated transcripts)
    depth = np.zeros(len(clones.seq) + 1, dtype=np.int64)
    st = clones.off[reads.clone] + reads.tstart
    np.add.at(depth, st, 1)
    np.add.at(depth, st + reads.L, -1)
    depth = np.cumsum(depth)[:-1]
    low = np.zeros(len(depth) + 1, dtype=np.int64)
    low[1:] = np.cumsum(depth < min_reads)
    cov = (low[clones.off[:-1] + b] - low[clones.off[:-1] + a]) == 0
    c = contig_codes_concat
    n = len(c) - SIG + 1
    spanned = np.zeros(ncl, dtype=bool)
    if n > 0:
        km = _codes_to_kmers(np.minimum(c, 3))
        bad = np.zeros(len(c) + 1, dtype=np.int64)
        bad[1:] = np.cumsum(c > 3)
        ok = (bad[SIG:SIG + n] - bad[:n]) == 0
        cid = np.searchsorted(np.asarray(contig_off, dtype=np.int64), np.arange(n), side="right") - 1
        ok &= cid == np.searchsorted(np.asarray(contig_off, dtype=np.int64), np.arange(n) + SIG - 1, side="right") - 1
        sig = np.unique(np.concatenate([A_fw, A_rc, B_fw, B_rc]))
        pos = np.searchsorted(sig, km)
        pos[pos >= len(sig)] = len(sig) - 1
        hit = ok & (sig[pos] == km)
        hp = np.flatnonzero(hit)
        hk = km[hp]
        order = np.argsort(hk, kind="stable")
        hk, hp = hk[order], hp[order]

        def places(code):
            lo_, hi_ = np.searchsorted(hk, code, side="left"), np.searchsorted(hk, code, side="right")
            return hp[lo_:hi_]

        for x in np.flatnonzero(cov):
            pa, pb = places(A_fw[x]), places(B_fw[x])
            if len(pa) and len(pb):
                want = pa + dist[x]
                m = np.isin(want, pb)
                if m.any() and (cid[pa[m]] == cid[want[m]]).any():
                    spanned[x] = True
                    continue
            pa, pb = places(A_rc[x]), places(B_rc[x])     # reverse strand: rc(B) comes first, rc(A) `dist` later
            if len(pa) and len(pb):
                want = pb + dist[x]
                m = np.isin(want, pa)
                if m.any() and (cid[pb[m]] == cid[want[m]]).any():
                    spanned[x] = True
    n_cov = int(cov.sum())
    return {"clonotypes": int(ncl), "covered_by_reads": n_cov, "spanned_by_one_contig": int((spanned & cov).sum()),
            "spanned_fraction": float((spanned & cov).sum() / max(1, n_cov)),
            "window": "V end - %d .. J end + %d (about 200 bp), both end 24-mers in one contig at the right distance" % (v_flank, j_flank)}


def contigs_from_packed(buf):
    """Consensus codes of every record of a t4_streams_pack_contigs buffer (numpy uint8) -> (codes concat, offsets, count)."""
    lut = np.full(256, 4, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    parts, offs, o, tot = [], [0], 0, 0
    n = len(buf)
    while o + 32 <= n:
        h = buf[o:o + 32].view(np.uint32)
        ln, rb = int(h[2]), int(h[6])
        if rb == 0:
            break
        parts.append(lut[buf[o + 32:o + 32 + ln]])
        tot += ln
        offs.append(tot)
        o += rb
    return (np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)), np.array(offs, dtype=np.int64), len(parts)


def contigs_from_output(text):
    """The same from SeqSet::Output text (_raw.out)."""
    lut = np.full(256, 4, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    parts, offs, tot = [], [0], 0
    lines = text.split(b"\n")
    for i, l in enumerate(lines):
        if l.startswith(b">") and i + 1 < len(lines):
            s = np.frombuffer(lines[i + 1], dtype=np.uint8)
            parts.append(lut[s])
            tot += len(s)
            offs.append(tot)
    return (np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)), np.array(offs, dtype=np.int64), len(parts)
