#!/usr/bin/env bash
# round 2, MFCC kernel cycle: parity, timing at two sizes, one full ncu capture of the kernel
TAG=${1:-mfcc}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "mfcc" > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -3 gpurun_out/pytest_$TAG.txt
timeout 300 python tools/prof_small.py --which mfcc --iters 5 > gpurun_out/prof_$TAG.log 2>&1; RC=$?
tail -2 gpurun_out/prof_$TAG.log
timeout 600 python tools/bench_extras.py --only mfcc --out gpurun_out/extras_$TAG.json > gpurun_out/extras_$TAG.log 2>&1
tail -2 gpurun_out/extras_$TAG.log
if [ $RC -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mfcc_ -s 1 -c 1 -o gpurun_out/ncu_$TAG python tools/prof_small.py --which mfcc --iters 2 > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log
fi
