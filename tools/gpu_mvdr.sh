#!/usr/bin/env bash
# MVDR on the B200 box: GPU parity tests, timing of both paths, launch list and one full capture per path.
TAG=${1:-mv1}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu ${2:+-k "$2"} > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -4 gpurun_out/pytest_$TAG.txt
timeout 300 python tools/prof_mvdr.py > gpurun_out/prof_mvdr_$TAG.log 2>&1; RC=$?
JDSP_MVDR_PATH=fft timeout 300 python tools/prof_mvdr.py >> gpurun_out/prof_mvdr_$TAG.log 2>&1
timeout 300 python tools/prof_mvdr.py --dtime 2.5e-4 >> gpurun_out/prof_mvdr_$TAG.log 2>&1
cat gpurun_out/prof_mvdr_$TAG.log
timeout 600 python tools/bench_extras.py --only mvdr --out gpurun_out/extras_mvdr_$TAG.json > gpurun_out/extras_mvdr_$TAG.log 2>&1; tail -2 gpurun_out/extras_mvdr_$TAG.log
if [ $RC -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mvdr_ -c 12 --csv --log-file gpurun_out/launches_mvdr_$TAG.csv \
      python tools/prof_mvdr.py --iters 2 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:mvdr_ -s 1 -c 1 -o gpurun_out/mvdr_td_$TAG python tools/prof_mvdr.py --iters 2 > gpurun_out/ncu_mvdr_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_mvdr_$TAG.log
fi
JDSP_MVDR_PATH=fft ncu --set full --clock-control none --import-source on -k regex:mvdr_apply -s 1 -c 1 -o gpurun_out/mvdr_fft_$TAG python tools/prof_mvdr.py --iters 2 > /dev/null 2>&1
