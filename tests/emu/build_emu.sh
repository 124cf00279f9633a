#!/bin/bash
# TEST INFRASTRUCTURE: compile the kernel sources for the CPU emulation (see emu_runtime.cpp).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
SAN=${MMG_EMU_SANITIZE:-}
g++ -std=c++20 -O1 -g -fPIC -shared -DMMG_CPU_EMU=1 $SAN -Wall -Wno-unknown-pragmas -Wno-unused-function -Wno-unused-variable \
    -x c++ "$ROOT/multimodalgame_b200/csrc/mmg_api.cu" "$HERE/emu_runtime.cpp" \
    -o "$HERE/libmmg_emu.so" -lpthread
echo "built $HERE/libmmg_emu.so"
