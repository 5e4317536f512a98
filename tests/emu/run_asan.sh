#!/bin/sh
# AddressSanitizer run of the kernel sources on the host emulator (see asan_cases.py).  Prints "layers ok" last.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT=${TMPDIR:-/tmp}/libpml_emu_asan.so
g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++20 -shared -fPIC -include "$HERE/cuda_emu.h" -x c++ \
    "$ROOT/self-supervised-depth-estimation_b200/csrc/pml_api.cu" -o "$OUT" -lpthread
PML_EMU_ASAN_LIB="$OUT" LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
    ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1 python "$HERE/asan_cases.py"
