"""t4_shard_reads (host-side, include/trust4_b200.h): the stream assignment the batch drop-in uses.  Structural properties
on the product library (no device needed) and per-stream parity with the reference through the engine emulation."""
import numpy as np
import pytest

import parity_cases as pc
from trust4_b200 import api, synth


@pytest.fixture(scope="module")
def host_lib():
    return api.default_lib()          # dlopen only: t4_shard_reads does no device work


def _check_structure(w, off, d, order, mode):
    n = len(w.descs)
    S = len(off) - 1
    assert off[0] == 0 and off[-1] == n and (np.diff(off) > 0).all()
    assert sorted(order.tolist()) == list(range(n))                       # a permutation
    old = w.descs[order]
    for f in ("seq_off", "len", "barcode", "min_cnt", "min_kmer_count", "sim_threshold", "name_id", "flags", "strand_in", "novel_strand"):
        assert (old[f] == d[f]).all(), f
    stream_of_new = np.searchsorted(off, np.arange(n), side="right") - 1
    stream_of_old = np.empty(n, dtype=np.int64)
    stream_of_old[order] = stream_of_new
    new_pos = np.empty(n, dtype=np.int64)
    new_pos[order] = np.arange(n)
    for s in range(S):
        o = order[off[s]:off[s + 1]]
        assert (np.diff(o) > 0).all()                                     # a stream keeps the sorted order
    # runs are never split and stay consecutive; eq_* are stream-relative
    lo_old = w.descs["eq_lo"].astype(np.int64)
    assert (stream_of_old == stream_of_old[lo_old]).all()
    base = off[stream_of_new]
    assert (d["eq_lo"] + base == new_pos[lo_old[order]]).all()
    assert (d["eq_hi"] - d["eq_lo"] == (w.descs["eq_hi"] - w.descs["eq_lo"])[order]).all()
    # mates: kept (stream-relative) when in the same stream, else -1
    m_old = w.descs["mate_idx"].astype(np.int64)[order]
    same = (m_old >= 0) & (stream_of_old[np.maximum(m_old, 0)] == stream_of_new)
    assert (d["mate_idx"][~same] == -1).all()
    assert (d["mate_idx"][same] + base[same] == new_pos[m_old[same]]).all()
    if mode == api.SHARD_RANK:
        assert (order == np.arange(n)).all()
    if mode == api.SHARD_GENE:
        # a gene's runs share a stream unless the gene was dear enough to be cut into several
        key = w.descs["name_id"][lo_old]
        cost = synth.read_cost(w.descs, None)
        share = cost.sum() / S
        for g in np.unique(key):
            m = key == g
            if cost[m].sum() <= share:
                assert len(np.unique(stream_of_old[m])) == 1, g


@pytest.mark.parametrize("mode", [api.SHARD_RANK, api.SHARD_GENE])
@pytest.mark.parametrize("S", [1, 3, 16, 5000])
def test_shard_reads_structure(host_lib, mode, S):
    w = pc.small_workload(11, nclones=40, npairs=700)
    off, d, order = api.shard_reads(w.descs, S, mode, host_lib)
    assert 1 <= len(off) - 1 <= min(S, len(w.descs))
    _check_structure(w, off, d, order, mode)
    if S == 1:
        assert (order == np.arange(len(order))).all() and (d["mate_idx"] == w.descs["mate_idx"]).all()


def test_shard_reads_balance(host_lib):
    w = pc.small_workload(12, nclones=120, npairs=6000)
    for mode in (api.SHARD_RANK, api.SHARD_GENE):
        off, d, _ = api.shard_reads(w.descs, 24, mode, host_lib)
        c = np.add.reduceat(synth.read_cost(d, None), off[:-1])
        assert c.max() < 1.6 * c.mean(), (mode, c.max() / c.mean())


def test_shard_reads_barcodes_whole(host_lib):
    cl = synth.make_clones(30, 5)
    rd, bc = synth.sample_single_cell(cl, 12, 60, 150, 5)
    w = synth.build_workload(cl, rd, barcode=bc)
    off, d, order = api.shard_reads(w.descs, 5, api.SHARD_BARCODE, host_lib)
    for b in np.unique(d["barcode"]):
        s = np.unique(np.searchsorted(off, np.flatnonzero(d["barcode"] == b), side="right"))
        assert len(s) == 1 or b == -1


def test_shard_reads_rejects_bad_input(host_lib):
    w = pc.small_workload(3, nclones=10, npairs=50)
    bad = w.descs.copy()
    bad["eq_hi"][0] = 0
    with pytest.raises(api.T4Error):
        api.shard_reads(bad, 4, api.SHARD_GENE, host_lib)
    with pytest.raises(api.T4Error):
        api.shard_reads(w.descs, 4, 7, host_lib)
    off, d, order = api.shard_reads(w.descs[:0], 4, api.SHARD_GENE, host_lib)
    assert list(off) == [0, 0] and len(d) == 0


@pytest.mark.parametrize("mode", [api.SHARD_RANK, api.SHARD_GENE])
def test_emu_batch_c_sharder_vs_ref(emu_lib, ref, mode):
    assert pc.check_batch_vs_ref(emu_lib, ref, 8, 9, nclones=40, npairs=700, c_mode=mode) > 100
