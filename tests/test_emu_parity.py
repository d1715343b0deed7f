"""CPU checks of the engine's bit-exact logic through the TEST-ONLY emulation build (tests/emu): the same
t4_engine.h compiled by g++ with one emulated thread per stream.  The product path is exercised by
test_gpu_parity.py on the B200; this module only guards the logic while developing without a GPU."""
import pytest

import parity_cases as pc


@pytest.mark.parametrize("name", ["example", "synth2k"])
def test_emu_trace_replay(emu_lib, name):
    pc.check_trace_replay(emu_lib, name)


@pytest.mark.parametrize("seed,shards", [(1, 1), (2, 3), (3, 8)])
def test_emu_batch_vs_reference(emu_lib, ref, seed, shards):
    assert pc.check_batch_vs_ref(emu_lib, ref, seed, shards) > 100


def test_emu_batch_dealt_shards(emu_lib, ref):
    assert pc.check_batch_vs_ref(emu_lib, ref, 6, 5, nclones=30, npairs=600, deal=True) > 100


def test_emu_stage_parity(emu_lib, ref):
    pc.check_stage_parity(emu_lib, ref, "synth2k", every=131, max_checks=25)


def test_emu_dp(emu_lib, ref):
    pc.check_dp(emu_lib, ref)


def test_emu_big_repeats(emu_lib, ref):
    pc.check_big_repeats(emu_lib, ref)


def test_emu_barcode_mode(emu_lib, ref):
    pc.check_barcode_mode(emu_lib, ref)


def test_emu_barcode_release_is_unobservable(emu_lib, ref):
    pc.check_barcode_release_unobservable(emu_lib, ref)
