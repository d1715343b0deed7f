"""Host-side workload generator and sharding (trust4_b200/synth.py): record layout, sort order, sharding invariants."""
import ctypes as C

import numpy as np

from trust4_b200 import synth


def _wl(seed=3, nclones=30, npairs=800):
    cl = synth.make_clones(nclones, seed)
    rd = synth.sample_pairs(cl, npairs, 150, seed)
    return cl, rd, synth.build_workload(cl, rd)


def test_read_desc_matches_c_struct():
    class D(C.Structure):                      # include/trust4_b200.h: struct t4_read_desc
        _fields_ = [("seq_off", C.c_uint64), ("len", C.c_int32), ("barcode", C.c_int32), ("min_cnt", C.c_int32),
                    ("min_kmer_count", C.c_int32), ("sim_threshold", C.c_double), ("name_id", C.c_int32), ("mate_idx", C.c_int32),
                    ("eq_lo", C.c_int32), ("eq_hi", C.c_int32), ("flags", C.c_uint32), ("strand_in", C.c_int8),
                    ("novel_strand", C.c_int8), ("gene4", C.c_char * 4), ("pad_", C.c_int8 * 2)]
    assert C.sizeof(D) == synth.READ_DESC.itemsize == 64
    for name, _ in D._fields_:
        assert getattr(D, name).offset == synth.READ_DESC.fields[name][1], name


def test_sorted_order_and_duplicates():
    cl, rd, w = _wl()
    d = w.descs
    n = len(d)
    assert n == 2 * 800
    # minCnt descending (main.cpp:103-125); identical reads adjacent and flagged as duplicates of their predecessor
    assert (np.diff(d["min_cnt"]) <= 0).all()
    reads = w.pool.reshape(n, 150)
    same = (reads[1:] == reads[:-1]).all(axis=1)
    assert (((d["flags"][1:] & synth.RD_DUP) != 0) == same).all()
    assert not (d["flags"][0] & synth.RD_DUP)
    # eq ranges are maximal runs of identical reads
    for i in range(0, n, 37):
        lo, hi = int(d["eq_lo"][i]), int(d["eq_hi"][i])
        assert lo <= i < hi and (reads[lo:hi] == reads[i]).all()
        assert lo == 0 or not (reads[lo - 1] == reads[i]).all()
        assert hi == n or not (reads[hi] == reads[i]).all()
    # mates point at each other
    m = d["mate_idx"]
    ok = m >= 0
    assert ok.all() and (m[m[ok]] == np.arange(n)[ok]).all()
    # thresholds follow main.cpp:1676-1694
    thr = d["sim_threshold"]
    assert set(np.unique(thr)) <= {0.9, 0.95, 0.97}
    assert (thr[d["min_cnt"] >= 20] == 0.97).all()


def test_kmer_stats_torch_equals_numpy():
    cl = synth.make_clones(10, 9)
    rd = synth.sample_pairs(cl, 150, 150, 9)
    a = synth.kmer_stats(rd.codes)
    b = synth.kmer_stats(rd.codes, device="cpu")
    for x, y in zip(a, b):
        assert (x == y).all()


def test_sharding_invariants():
    cl, rd, w = _wl(5)
    n = len(w.descs)
    for deal, balance, group in ((False, "reads", ""), (True, "reads", ""), (False, "cost", ""), (False, "cost", "gene")):
        off, d = synth.shard_workload(w, 7, deal=deal, balance=balance, group=group)
        if balance == "cost" and not group:
            cost = np.add.reduceat(synth.read_cost(w.descs, w.med_cnt), off[:-1][np.diff(off) > 0])
            assert cost.max() < 2.0 * cost.mean()          # cut by predicted cost, not by read count
        assert off[0] == 0 and off[-1] == n and (np.diff(off) >= 0).all()
        seen = np.zeros(n, dtype=int)
        for j in range(len(off) - 1):
            lo, hi = int(off[j]), int(off[j + 1])
            dd = d[lo:hi]
            if hi > lo:
                assert not (dd["flags"][0] & synth.RD_DUP)             # a run of identical reads is never split
                assert (np.diff(dd["seq_off"].astype(np.int64)) > 0).all()   # global sorted order kept inside a stream
            m = dd["mate_idx"]
            assert ((m == -1) | ((m >= 0) & (m < hi - lo))).all()
            ok = m >= 0
            assert (dd["mate_idx"][m[ok]] == np.arange(hi - lo)[ok]).all()
            assert ((dd["eq_lo"] >= 0) & (dd["eq_hi"] <= hi - lo) & (dd["eq_lo"] <= np.arange(hi - lo)) & (np.arange(hi - lo) < dd["eq_hi"])).all()
            seen[(dd["seq_off"] // 150).astype(int)] += 1
        assert (seen == 1).all()                                       # every record in exactly one stream


def test_has_motif_matches_c_abi_host_utility():
    from trust4_b200 import api
    lib = api.Lib()
    cl, rd, w = _wl(7, 20, 300)
    flags = w.descs["flags"]
    for i in range(0, len(w.descs), 11):
        assert bool(flags[i] & synth.RD_MOTIF) == (lib.has_motif(w.read(i).encode(), 1) != 0)


def test_gene_streams_budget():
    """_gene_shards: the cap search leaves no giant leftover stream, never more than n_shards streams, runs stay whole."""
    cl, rd, w = _wl(9, 150, 5000)
    cost = synth.read_cost(w.descs, w.med_cnt)
    for S in (8, 64, 300):
        a = synth._gene_shards(w, S)
        assert a.max() + 1 <= S
        c = np.bincount(a, weights=cost)
        assert c.max() < 2.0 * c.mean() + cost.max() * 40
        head = w.descs["eq_lo"].astype(np.int64)
        assert (a == a[head]).all()
