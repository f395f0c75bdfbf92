#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY: compile the product sources against the CPU execution emulator.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
mkdir -p "$HERE/_build"
g++ -O1 -g -std=c++17 -fPIC -shared -DJDSP_EMUL -I"$HERE" -I"$ROOT/jeicyboodsp_b200/csrc" \
    -Wall -Wno-unused-function -Wno-unused-variable -Wno-unknown-pragmas \
    -x c++ "$ROOT/jeicyboodsp_b200/csrc/jdsp_api.cu" -x c++ "$HERE/cuda_emul.cpp" \
    -o "$HERE/_build/libjdsp_emul.so"
echo "$HERE/_build/libjdsp_emul.so"
