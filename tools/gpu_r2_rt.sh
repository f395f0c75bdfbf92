#!/usr/bin/env bash
# round 2, round-trip cycle: parity (all GPU tests that touch the round trip + denoise after the Wiener change), timing both kernels, ncu
TAG=${1:-rt}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -k "roundtrip or round_trip or denoise or programs" > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -3 gpurun_out/pytest_$TAG.txt
timeout 600 python tools/bench_extras.py --only roundtrip --out gpurun_out/extras_$TAG.json > gpurun_out/extras_$TAG.log 2>&1; RC=$?
tail -3 gpurun_out/extras_$TAG.log | cut -c1-400
JDSP_ROUNDTRIP_CTA=1 timeout 600 python tools/bench_extras.py --only roundtrip --out gpurun_out/extras_${TAG}_cta.json > gpurun_out/extras_${TAG}_cta.log 2>&1
tail -3 gpurun_out/extras_${TAG}_cta.log | cut -c1-400
timeout 300 python tools/prof_denoise.py --streams 4096 --seconds 8 > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log
if [ $RC -eq 0 ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:roundtrip_warp -s 1 -c 1 -o gpurun_out/ncu_$TAG python tools/bench_extras.py --only roundtrip --out gpurun_out/extras_ncu_$TAG.json > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log
fi
