#!/usr/bin/env bash
# round 2, FFT cycle: parity of every plan, the sweep from 2^13 up with and without the cluster kernel, one ncu capture at 2^15
TAG=${1:-fft}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "fft" > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -3 gpurun_out/pytest_$TAG.txt
timeout 300 python tools/sweep_quick.py 28 13 > gpurun_out/sweep_$TAG.log 2>&1; RC=$?
cat gpurun_out/sweep_$TAG.log
JDSP_FFT_CLUSTER=1 timeout 300 python tools/sweep_quick.py 28 15 > gpurun_out/sweep_${TAG}_cluster.log 2>&1
cat gpurun_out/sweep_${TAG}_cluster.log
if [ $RC -eq 0 ]; then
  JDSP_FFT_CLUSTER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cluster2 -s 1 -c 1 -o gpurun_out/ncu_$TAG python tools/sweep_quick.py 26 15 > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log
fi
