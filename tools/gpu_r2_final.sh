#!/usr/bin/env bash
# round 2 evidence: GPU suite, smoke, bench line (5 configs), reference arm, bench launch list (ncu, time only), full ncu captures of the
# five config kernels at profiler-friendly sizes
TAG=${1:-fin}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -3 gpurun_out/pytest_$TAG.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.txt; tail -2 gpurun_out/smoke_$TAG.txt
START=$(date +%s)
timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$? wall=$(( $(date +%s) - START )) s"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("headline", round(d["value"]), d["roofline"]["frac"], d["roofline"]["fp32_frac"], d["roofline"]["per_mode_ms"], "e2e", d["e2e"]["value"], d["parity"])
for c in d["configs"]:
    print(c["name"], c.get("error"), c.get("ms"), (c.get("roofline") or {}).get("frac"), (c.get("roofline") or {}).get("frac_mean"), c.get("parity"), (c.get("e2e") or {}).get("value"), round(c.get("wall_s", 0)))
    if c["name"] == "fft_sweep":
        for r in c["sizes"]: print("   ", r["n"], round(r["fwd_frac_hbm"], 3), round(r["inv_frac_hbm"], 3), r.get("parity_max_rel"))
PY
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:denoise_stream -s 2 -c 2 -o gpurun_out/ncu_denoise_$TAG python tools/prof_denoise.py --streams 4096 --seconds 4 --iters 2 > gpurun_out/ncu_denoise_$TAG.log 2>&1; echo "ncu denoise rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mfcc_kernel|fastconv_stream" -s 1 -c 2 -o gpurun_out/ncu_small_$TAG python tools/prof_small.py --which mfcc,fastconv --iters 2 > gpurun_out/ncu_small_$TAG.log 2>&1; echo "ncu mfcc/fastconv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roundtrip_warp -s 1 -c 1 -o gpurun_out/ncu_rt_$TAG python tools/prof_roundtrip.py > gpurun_out/ncu_rt_$TAG.log 2>&1; echo "ncu roundtrip rc=$?"
