#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY: compile the product sources against the CPU execution emulator.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
CSRC="$ROOT/jeicyboodsp_b200/csrc"
mkdir -p "$HERE/_build"
FLAGS="-O1 -g -std=c++17 -fPIC -DJDSP_EMUL -I$HERE -I$CSRC -Wall -Wno-unused-function -Wno-unused-variable -Wno-unknown-pragmas"
pids=()
for tu in jdsp_api jdsp_stft jdsp_conv_mfcc jdsp_pitch jdsp_mvdr; do
    g++ $FLAGS -c -x c++ "$CSRC/$tu.cu" -o "$HERE/_build/$tu.o" &
    pids+=($!)
done
g++ $FLAGS -c "$HERE/cuda_emul.cpp" -o "$HERE/_build/cuda_emul.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -o "$HERE/_build/libjdsp_emul.so" "$HERE/_build/jdsp_api.o" "$HERE/_build/jdsp_stft.o" \
    "$HERE/_build/jdsp_conv_mfcc.o" "$HERE/_build/jdsp_pitch.o" "$HERE/_build/jdsp_mvdr.o" "$HERE/_build/cuda_emul.o"
echo "$HERE/_build/libjdsp_emul.so"
