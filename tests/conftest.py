import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def emu_lib():
    """TEST-ONLY single-thread emulation of the device code (g++ -DT4_EMU): lets the bit-exact logic of the
    engine be checked against the oracle without a GPU.  Exports t4emu_* symbols; never shipped."""
    from trust4_b200 import api
    src = [os.path.join(ROOT, "trust4_b200", "csrc", f) for f in ("t4_api.cu", "t4_engine.h", "t4_common.h", "t4_shard.h", "t4_probe.cuh", "t4_assign.h", "t4_refscan.h", "t4_kcount.h", "t4_annot.h", "t4_readsort.h")]
    src.append(os.path.join(ROOT, "include", "trust4_b200.h"))
    out = os.path.join(ROOT, "tests", "emu", "libt4emu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in src):
        subprocess.run(["g++", "-x", "c++", "-std=c++17", "-DT4_EMU", "-O2", "-g", "-ffp-contract=off", "-fPIC", "-shared",
                        "-o", out, src[0]], check=True)
    lib = api.Lib(out, "t4emu_")
    lib.check(lib.init(0, 2 << 30))
    return lib


@pytest.fixture(scope="session")
def gpu_lib():
    from trust4_b200 import api
    lib = api.default_lib()
    lib.check(lib.init(0, 8 << 30))
    return lib


@pytest.fixture(scope="session")
def ref():
    import refharness
    if not refharness.available():
        pytest.skip("oracle/_ref/libt4ref.so not built (make -C oracle)")
    return refharness
