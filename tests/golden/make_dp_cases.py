#!/usr/bin/env python3
"""Golden vectors for GlobalAlignment_PosWeight / index rule / gap limit, produced by the compiled reference
(oracle/_ref/libt4ref.so).  Dev container only; output tests/golden/leaf_cases.json.gz."""
import gzip, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import refharness as rh
import parity_cases as pc

out = {"dp": [], "gap_limit": {}, "index": []}
for tw, p in pc.dp_cases(1234, 250):
    sc, ed = rh.dp_pos_weight(tw, p)
    out["dp"].append({"tw": tw.tolist(), "p": p, "score": sc, "edit": ed})
s = rh.RefSeqSet(9)
for k in (7, 9, 11, 13, 15, 17, 21, 31):
    s.change_kmer_length(k)
    out["gap_limit"][str(k)] = rh.lib().t4ref_nomatch_gap_limit(s.h)
# which offsets BuildIndexFromRead inserts (Appendix B of SURVEY.md and random contigs with N / homopolymers)
rng = np.random.default_rng(7)
seqs = ["AAAAAAAAAAAACGTACGTTGCA", "CCCCCCCCCCCCGTACGTTGCAT", "ACGTNACGTACGTTTGACCAGGT", "ACACACACACACACACACACACAC"]
for _ in range(20):
    seqs.append("".join("ACGTN"[c] for c in rng.choice(5, size=int(rng.integers(9, 80)), p=[.3, .23, .23, .23, .01])))
for sq in seqs:
    r = rh.RefSeqSet(9)
    r.input_novel_read("IGHV1", sq, 1, -1)
    posts = []
    # enumerate postings through the k-mers of the sequence itself
    seen = set()
    for i in range(len(sq) - 8):
        km = sq[i:i + 9]
        if "N" in km:
            continue
        code = 0
        for ch in km:
            code = code * 4 + "ACGT".index(ch)
        if code in seen:
            continue
        seen.add(code)
        for idx, off in r.index_lookup(code).tolist():
            posts.append([code, off])
    out["index"].append({"seq": sq, "postings": sorted(posts)})
with gzip.GzipFile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "leaf_cases.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps(out).encode())
print(len(out["dp"]), out["gap_limit"], len(out["index"]))
