#!/usr/bin/env bash
# Round-end evidence: GPU parity suite, smoke, bench line (both arms), extras for every config, MVDR launch list and captures.
TAG=${1:-fin}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -3 gpurun_out/pytest_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.txt 2>&1; tail -1 gpurun_out/smoke_$TAG.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; echo "ref rc=$?"
timeout 1200 python tools/bench_extras.py --out gpurun_out/extras_$TAG.json > gpurun_out/extras_$TAG.log 2>&1; echo "extras rc=$?"
python tools/prof_mvdr.py > gpurun_out/prof_mvdr_$TAG.log 2>&1
python tools/prof_mvdr.py --dtime 2.5e-4 >> gpurun_out/prof_mvdr_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mvdr_ -c 8 --csv --log-file gpurun_out/launches_mvdr_td_$TAG.csv python tools/prof_mvdr.py --iters 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mvdr_ -c 12 --csv --log-file gpurun_out/launches_mvdr_steered_$TAG.csv python tools/prof_mvdr.py --iters 2 --dtime 2.5e-4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mvdr_ -s 1 -c 1 -o gpurun_out/mvdr_td_$TAG python tools/prof_mvdr.py --iters 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mvdr_ -s 3 -c 3 -o gpurun_out/mvdr_steered_$TAG python tools/prof_mvdr.py --iters 2 --dtime 2.5e-4 > /dev/null 2>&1
cut -c1-900 gpurun_out/bench_$TAG.json; echo; cut -c1-400 gpurun_out/bench_ref_$TAG.json; echo; grep -o '"config": "[a-z_]*".\{0,60\}\|"frac_hbm": [0-9.]*\|"fwd_frac_hbm": [0-9.]*' gpurun_out/extras_$TAG.log | paste -sd' ' | fold -w 200
