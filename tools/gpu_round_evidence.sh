#!/usr/bin/env bash
# Evidence for profiles/: bench line, ncu launch list of the same command, one full capture of the top kernel.
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/bench_plain_$TAG.json 2>/dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:denoise_ -c 40 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/bench_under_ncu_$TAG.log 2>&1
python tools/prof_denoise.py --iters 2 > gpurun_out/prof_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:denoise_ -s 2 -c 2 -o gpurun_out/denoise_$TAG python tools/prof_denoise.py --iters 2 > gpurun_out/ncu_$TAG.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null
cat gpurun_out/bench_$TAG.json | cut -c1-1500; cat gpurun_out/bench_ref_$TAG.json | cut -c1-600; tail -3 gpurun_out/prof_$TAG.log
