#!/usr/bin/env bash
# round 2: whole GPU test suite, then the full bench line (all five BASELINE configs) with its wall time
TAG=${1:-b}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -4 gpurun_out/pytest_$TAG.txt
START=$(date +%s)
timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$? wall=$(( $(date +%s) - START )) s"
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("headline", round(d["value"]), d["roofline"]["frac"], d["roofline"]["fp32_frac"], d["roofline"]["per_mode_ms"], "e2e", d["e2e"]["value"], d["e2e"]["gbs_per_rank"], d["e2e"]["copy_only_gbs_per_rank"], d["parity"])
for c in d["configs"]:
    print(c["name"], c.get("error"), c.get("ms"), (c.get("roofline") or {}).get("frac"), (c.get("roofline") or {}).get("frac_mean"), c.get("parity"), (c.get("e2e") or {}).get("value"), (c.get("e2e") or {}).get("equals_resident_path"), (c.get("cpu_baseline") or {}).get("value"), round(c.get("wall_s", 0)))
    if c["name"] == "fft_sweep":
        for r in c["sizes"]: print("   ", r["n"], round(r["fwd_frac_hbm"], 3), round(r["inv_frac_hbm"], 3), r.get("parity_max_rel"))
PY
