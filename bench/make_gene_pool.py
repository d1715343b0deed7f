#!/usr/bin/env python3
"""Build bench/data/gene_pool.tsv.gz from the reference's IMGT gene set.

Run once in the dev container (needs /root/reference/human_IMGT+C.fa, which does
not exist on the GPU box).  Applies the pool filters of SURVEY.md Appendix A:
per chain c in {IGH, IGK, IGL, TRA, TRB}: V = name[3]=='V', len>=250, no '/';
J = name[3]=='J', len>=30; C = name[3] not in VDJ, len>=200 (IGH: only
IGH{M,G,A,E,D} not followed by a digit).  '.' stripped, upper-cased.
Output columns: chain, segment, name, sequence.
"""
import gzip, os, sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/human_IMGT+C.fa"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "gene_pool.tsv.gz")

def read_fa(path):
    name, seq = None, []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if name is not None:
                yield name, "".join(seq)
            name, seq = line[1:].split()[0], []
        else:
            seq.append(line)
    if name is not None:
        yield name, "".join(seq)

rows = []
for name, seq in read_fa(REF):
    seq = seq.replace(".", "").upper()
    chain = name[:3]
    if chain not in ("IGH", "IGK", "IGL", "TRA", "TRB") or len(name) < 4:
        continue
    if any(ch not in "ACGT" for ch in seq):
        continue
    t = name[3]
    if t == "V":
        if len(seq) >= 250 and "/" not in name:
            rows.append((chain, "V", name, seq))
    elif t == "J":
        if len(seq) >= 30:
            rows.append((chain, "J", name, seq))
    else:
        c5 = name[4] if len(name) > 4 else ""
        if t == "D" and c5.isdigit():
            continue                      # a D gene (IGHD1-1, TRBD1 ...)
        if len(seq) < 200:
            continue
        if chain == "IGH" and not (t in "MGAED" and not (t == "D" and c5.isdigit())):
            continue
        rows.append((chain, "C", name, seq))
os.makedirs(os.path.dirname(OUT), exist_ok=True)
with gzip.open(OUT, "wt") as f:
    for r in rows:
        f.write("\t".join(r) + "\n")
from collections import Counter
print(len(rows), Counter((r[0], r[1]) for r in rows))
