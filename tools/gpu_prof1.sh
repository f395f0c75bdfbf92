#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 120 profiles/microbench/mb2 > gpurun_out/mb2.txt 2>&1
python tools/prof_denoise.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:denoise_ -s 2 -c 2 -o gpurun_out/denoise_r1a python tools/prof_denoise.py --iters 2 > gpurun_out/ncu_run.log 2>&1
cat gpurun_out/mb2.txt; cat gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu_run.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu2.txt 2>&1; tail -4 gpurun_out/pytest_gpu2.txt
