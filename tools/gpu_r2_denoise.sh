#!/usr/bin/env bash
# round 2, denoise kernel cycle: parity, the headline timing, one full ncu capture of the stream-group kernel
TAG=${1:-dn}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "denoise or host_forms or api" > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -3 gpurun_out/pytest_$TAG.txt
timeout 300 python tools/prof_denoise.py > gpurun_out/prof_$TAG.log 2>&1; RC=$?
tail -4 gpurun_out/prof_$TAG.log
if [ $RC -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:denoise_stream -s 2 -c 1 -o gpurun_out/ncu_$TAG python tools/prof_denoise.py --iters 2 > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log
fi
