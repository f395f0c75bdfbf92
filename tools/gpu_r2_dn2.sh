#!/usr/bin/env bash
# round 2, denoise / pitch / MVDR cycle after a change to the shared per-bin or exchange code: parity, timings at one full wave, ncu
TAG=${1:-dn}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -k "denoise or host_forms or api or pitch or mvdr" > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -3 gpurun_out/pytest_$TAG.txt
timeout 300 python tools/prof_denoise.py --streams 4096 --seconds 8 > gpurun_out/prof_$TAG.log 2>&1; RC=$?
tail -4 gpurun_out/prof_$TAG.log
timeout 300 python tools/prof_pitch.py > gpurun_out/prof_pitch_$TAG.log 2>&1
tail -3 gpurun_out/prof_pitch_$TAG.log
if [ $RC -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:denoise_stream -s 2 -c 1 -o gpurun_out/ncu_$TAG python tools/prof_denoise.py --streams 4096 --seconds 4 --iters 2 > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log
fi
