#!/usr/bin/env bash
# round 2, cluster FFT cycle: parity of the alternate plans, sweep 2^14..2^16 with the default plans and with JDSP_FFT_CLUSTER16, one ncu capture
TAG=${1:-fftc}; NCUN=${2:-15}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "fft_alternate" > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -3 gpurun_out/pytest_$TAG.txt
timeout 300 python tools/sweep_quick.py 28 14 > gpurun_out/sweep_$TAG.log 2>&1
cat gpurun_out/sweep_$TAG.log
JDSP_FFT_CLUSTER16=1 timeout 300 python tools/sweep_quick.py 28 14 > gpurun_out/sweep_${TAG}_c16.log 2>&1; RC=$?
cat gpurun_out/sweep_${TAG}_c16.log
if [ $RC -eq 0 ]; then
  JDSP_FFT_CLUSTER16=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_c2c_cluster_kernel -s 1 -c 1 -o gpurun_out/ncu_$TAG python tools/sweep_quick.py 26 $NCUN > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log
fi
