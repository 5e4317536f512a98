#!/bin/sh
# Build the host-thread emulation of the kernel sources (test infrastructure, see cuda_emu.h).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
g++ -O2 -std=c++20 -shared -fPIC -include "$HERE/cuda_emu.h" -x c++ \
    "$ROOT/self-supervised-depth-estimation_b200/csrc/pml_api.cu" -o "$HERE/libpml_emu.so" -lpthread
