#!/usr/bin/env bash
# Build -> measure cycle on the B200 box: GPU parity tests, small timing run, optional ncu capture, bench line.
# usage: bash tools/gpu_cycle.sh <tag> [ncu]
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -3 gpurun_out/pytest_$TAG.txt
timeout 300 python tools/prof_denoise.py > gpurun_out/prof_$TAG.log 2>&1; RC=$?
tail -4 gpurun_out/prof_$TAG.log
if [ "$2" = "ncu" ] && [ $RC -eq 0 ]; then
  python tools/prof_denoise.py --iters 2 > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:denoise_ -s 2 -c 2 -o gpurun_out/denoise_$TAG python tools/prof_denoise.py --iters 2 > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log
fi
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['per_mode_ms'], d['e2e']['value'], d['parity'], d['clocks'])"
tail -3 gpurun_out/bench_$TAG.err
