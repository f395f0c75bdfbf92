#!/usr/bin/env bash
# round 2, one-kernel cycle: parity tests, timing at the profiler size and at the bench size, one full ncu capture
# usage: bash tools/gpu_r2_kernel.sh <tag> <pytest -k expression> <prof_small --which> <bench_extras --only> <ncu kernel regex>
TAG=$1; KEXPR=$2; WHICH=$3; ONLY=$4; KREGEX=$5
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "$KEXPR" > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -3 gpurun_out/pytest_$TAG.txt
timeout 300 python tools/prof_small.py --which $WHICH --iters 5 > gpurun_out/prof_$TAG.log 2>&1; RC=$?
tail -2 gpurun_out/prof_$TAG.log
timeout 600 python tools/bench_extras.py --only $ONLY --out gpurun_out/extras_$TAG.json > gpurun_out/extras_$TAG.log 2>&1
tail -2 gpurun_out/extras_$TAG.log
if [ $RC -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 1 -c 1 -o gpurun_out/ncu_$TAG python tools/prof_small.py --which $WHICH --iters 2 > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log
fi
