"""bench/quality.py: the contiguity figure reported with read-sharded runs (clonotypes whose V(D)J core lies in one contig)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bench"))
import quality  # noqa: E402
from trust4_b200 import synth  # noqa: E402


def _contigs(seqs):
    codes = [synth.encode(s) for s in seqs]
    off = np.zeros(len(codes) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(c) for c in codes])
    return np.concatenate(codes), off


def test_spanning_fraction_on_hand_made_contigs():
    cl = synth.make_clones(12, 3)
    rd = synth.sample_pairs(cl, 1200, 150, 3)
    full = [synth.decode(cl.seq[cl.off[i]:cl.off[i + 1]]) for i in range(12)]
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    # whole transcripts as contigs (half of them reverse-complemented): every covered clonotype is spanned
    c, o = _contigs([s if i % 2 else "".join(comp[x] for x in reversed(s)) for i, s in enumerate(full)])
    q = quality.spanning_fraction(cl, rd, c, o)
    assert q["covered_by_reads"] >= 10 and q["spanned_by_one_contig"] == q["covered_by_reads"]
    # transcripts cut in the middle of the junction: nothing is spanned although every base is still present
    cut = []
    for i, s in enumerate(full):
        m = int(cl.seg_end[i, 0] + cl.seg_end[i, 1]) // 2
        cut += [s[:m], s[m:]]
    c, o = _contigs(cut)
    assert quality.spanning_fraction(cl, rd, c, o)["spanned_by_one_contig"] == 0
    # and the presence figure still finds the junction 24-mers only when they are not cut
    assert quality.recovered_fraction(cl, rd, *_contigs(full))["recovered_fraction"] == 1.0


def test_contigs_from_output_and_packed_agree(tmp_path):
    text = b">assemble0 IGHV1\nACGTNACGT\n1 0 0 0 0 1 0 0 0 \n0 1 0 0 0 0 1 0 0 \n0 0 1 0 0 0 0 1 0 \n0 0 0 1 0 0 0 0 1 \n>assemble3 X\nGGGG\n0 0 0 0 \n0 0 0 0 \n1 1 1 1 \n0 0 0 0 \n"
    c, o, n = quality.contigs_from_output(text)
    assert n == 2 and o.tolist() == [0, 9, 13] and c.tolist() == [0, 1, 2, 3, 4, 0, 1, 2, 3, 2, 2, 2, 2]
