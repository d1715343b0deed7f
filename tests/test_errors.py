"""Error conventions of the C ABI (include/trust4_b200.h): codes < T4_E_BASE, never an abort.  Run against the
test emulation of the engine (same host code paths as the product library)."""
import os
import subprocess
import sys

import pytest

from trust4_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NOMEM_SCRIPT = r"""
import sys
sys.path.insert(0, %r)
from trust4_b200 import api, synth
lib = api.Lib(%r, 't4emu_')
lib.check(lib.init(0, int(sys.argv[1]) << 20))
cl = synth.make_clones(30, 3); rd = synth.sample_pairs(cl, 1500, 150, 3); w = synth.build_workload(cl, rd)
try:
    sets = api.SeqSet.create_many(4, 9, lib)
    off, d = synth.shard_workload(w, 4)
    api.streams_run(sets, synth.run_cfg(), d, off, w.pool, w.names, lib)
    print("OK")
except api.T4Error as e:
    print("T4Error", e.code)
"""


@pytest.mark.parametrize("mb", [1, 3, 6])
def test_arena_exhaustion_is_an_error_code(emu_lib, mb):
    """A device arena that is too small yields T4_E_NOMEM from the batch entry -- in a fresh process, no crash."""
    p = subprocess.run([sys.executable, "-c", NOMEM_SCRIPT % (ROOT, emu_lib.path), str(mb)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    assert p.stdout.strip().splitlines()[-1] == "T4Error %d" % api.T4_E_NOMEM


def test_unsupported_inputs(emu_lib):
    emu_lib.check(emu_lib.reset())
    s = api.SeqSet(9, emu_lib)
    with pytest.raises(api.T4Error) as e:
        s.add_read("ACGT" * 200, "", 0, -1, 1, 0, 0.9)          # 800 bp > device read limit (512)
    assert e.value.code == api.T4_E_UNSUPPORTED
    with pytest.raises(api.T4Error) as e:
        s.set_is_long(1)                                          # isLongSeqSet (first read > 200 bp)
    assert e.value.code == api.T4_E_UNSUPPORTED
    assert s.set_is_long(0) == 0
    with pytest.raises(api.T4Error):
        s.change_kmer_length(40)
    # the set is still usable
    assert s.input_novel_read("IGHV1-2*01", "ACGTTGCATGCAGTCAGTCAGGGTTACCACGATCGATCGATTTGACGGATCGGAT", 1, -1) == 0
    assert s.size() == 1


def test_stale_handle_after_reset(emu_lib):
    emu_lib.check(emu_lib.reset())
    s = api.SeqSet(9, emu_lib)
    emu_lib.check(emu_lib.reset())
    with pytest.raises(api.T4Error) as e:
        s.size()
    assert e.value.code == api.T4_E_INVAL


def test_assign_pass_errors(emu_lib):
    """t4_streams_assign_reads: bad arguments give NULL + a message; a t4_assign does not survive t4_reset."""
    import ctypes as C
    import numpy as np
    from trust4_b200 import synth
    emu_lib.check(emu_lib.reset())
    cl = synth.make_clones(10, 4)
    w = synth.build_workload(cl, synth.sample_pairs(cl, 100, 150, 4))
    sets = api.SeqSet.create_many(1, 9, emu_lib)
    wl = api.Workload(w.descs, w.pool, w.names, emu_lib)
    hs = (C.c_void_p * 1)(sets[0].h)
    bad = np.array([1, len(w.descs)], dtype=np.int64)
    assert not emu_lib.streams_assign_reads(hs, 1, wl.h, bad.ctypes.data, 17, 0, None)
    assert b"desc_off" in emu_lib.last_error()
    assert not emu_lib.streams_assign_reads(hs, 1, None, bad.ctypes.data, 17, 0, None)
    with pytest.raises(api.T4Error):
        api.Assign(sets, wl, np.array([0, len(w.descs)]), kmer_length=40)      # k out of range
    with pytest.raises(api.T4Error):
        api.Assign(sets, wl, np.array([0, len(w.descs)]), 17)                  # no assembly run on this workload yet
    assert b"no assembly results" in emu_lib.last_error()
    off = np.array([0, len(w.descs)], dtype=np.int64)
    emu_lib.check(emu_lib.streams_run_resident(hs, 1, synth.run_cfg().ctypes.data, wl.h, off.ctypes.data, None))
    a = api.Assign(sets, wl, off, 17)
    assert a.stats()["reads"] > 100
    emu_lib.check(emu_lib.reset())
    with pytest.raises(api.T4Error) as e:
        a.results()
    assert e.value.code == api.T4_E_INVAL
    a.close()
    wl.close()
