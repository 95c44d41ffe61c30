#!/bin/sh
# Builds the serial CPU emulation of the engine's kernels (development / host-logic tests only; see csrc/common.cuh).
set -e
cd "$(dirname "$0")/../.."
g++ -x c++ -DAMG_EMU -std=c++17 -O1 -g -fPIC -shared -Wall -Wno-unused-variable -Wno-unused-function -Wno-sign-compare \
  automerge_classic_b200/csrc/capi.cu automerge_classic_b200/csrc/hostsha.cc -o tests/_emu/libamgpu_emu.so -lz -lpthread
