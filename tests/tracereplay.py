"""Replay a recorded SeqSet call trace (tests/golden/*.trace.gz, format tests/trace_format.md)
against any object with the SeqSet method names (reference harness, GPU engine, emulation)."""
import gzip


def load_trace(path):
    ops = []
    with gzip.open(path, "rt") as f:
        for line in f:
            t = line.rstrip("\n").split("\t")
            ops.append(t)
    return ops


def replay(ops, make_set, stop_at_first_mismatch=True, limit=None):
    """Returns (seqset, mismatches).  A mismatch is (op index, op, got)."""
    s = None
    bad = []
    for i, t in enumerate(ops):
        if limit is not None and i >= limit:
            break
        c = t[0]
        if c == "C":
            s = make_set(int(t[1]))
        elif c == "A":
            read, name, sin, bc, mk, rep, thr, ret, sout = t[1], t[2], int(t[3]), int(t[4]), int(t[5]), int(t[6]), float(t[7]), int(t[8]), int(t[9])
            if name == ".":
                name = ""
            r, so = s.add_read(read, name, sin, bc, mk, rep, thr)
            if (r, so) != (ret, sout):
                bad.append((i, t, (r, so)))
        elif c == "R":
            r = s.repeat_add_read(t[1])
            if r != int(t[2]):
                bad.append((i, t, r))
        elif c == "N":
            r = s.input_novel_read(t[1], t[2], int(t[3]), int(t[4]))
            if r != int(t[5]):
                bad.append((i, t, r))
        elif c == "U":
            s.update_all_consensus()
        elif c == "K":
            s.change_kmer_length(int(t[1]))
        elif c == "H":
            s.set_hit_len_required(int(t[1]))
        elif c == "M":
            pass  # HasMotif: pure host utility, checked separately
        elif c == "O":
            break
        elif c in ("L", "B", "F", "S"):
            raise NotImplementedError("trace op " + c)
        if bad and stop_at_first_mismatch:
            break
    return s, bad
