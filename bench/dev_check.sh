#!/bin/bash
# developer loop: rebuild the test emulation, run the CPU parity tests, rebuild the CUDA library
set -e
cd "$(dirname "$0")/.."
g++ -x c++ -std=c++17 -DT4_EMU -O2 -g -ffp-contract=off -fPIC -shared -Wall -Wno-unused-function -o tests/emu/libt4emu.so trust4_b200/csrc/t4_api.cu 2>&1 | grep -E "error|warning: unused" | head -20 || true
python -m pytest tests/test_emu_parity.py -x -q 2>&1 | tail -3
if [ "$1" != "nocuda" ]; then python -c "import __graft_entry__ as g; g.build_lib()" 2>&1 | grep -E "error" | head; ls -la trust4_b200/*.so; fi
