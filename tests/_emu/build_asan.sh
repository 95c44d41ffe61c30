#!/bin/sh
# AddressSanitizer / UBSan build of the serial CPU emulation (development aid, see build.sh). Run a check with it like this:
#   LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so.6)" ASAN_OPTIONS=detect_leaks=0 \
#     python -c "import sys; sys.path[:0]=['tests','.']; import parity_checks as P, oracle; from automerge_classic_b200.engine import doc_class_for; \
#                P.check_history_after_load(doc_class_for('tests/_emu/libamgpu_emu_asan.so'), 'C6', 300, 3)"
# (libstdc++ has to be preloaded next to libasan, or the first C++ exception aborts inside the interceptor)
set -e
cd "$(dirname "$0")/../.."
g++ -x c++ -DAMG_EMU -std=c++17 -O1 -g -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -Wno-unknown-pragmas \
  automerge_classic_b200/csrc/capi.cu automerge_classic_b200/csrc/hostsha.cc -o tests/_emu/libamgpu_emu_asan.so -lz -lpthread
