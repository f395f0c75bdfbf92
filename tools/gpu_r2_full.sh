#!/usr/bin/env bash
# round 2: whole GPU suite, smoke, the full bench line, the reference arm, and a compute-sanitizer attempt (its output or refusal is kept)
TAG=${1:-full}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.txt
tail -3 gpurun_out/pytest_$TAG.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.txt; tail -2 gpurun_out/smoke_$TAG.txt
START=$(date +%s)
timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$? wall=$(( $(date +%s) - START )) s"
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("headline", round(d["value"]), d["roofline"]["frac"], d["roofline"]["fp32_frac"], d["roofline"]["per_mode_ms"], "e2e", d["e2e"]["value"], d["parity"])
for c in d["configs"]:
    print(c["name"], c.get("error"), c.get("ms"), (c.get("roofline") or {}).get("frac"), (c.get("roofline") or {}).get("frac_mean"), c.get("parity"), (c.get("e2e") or {}).get("value"), round(c.get("wall_s", 0)))
PY
START=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$? wall=$(( $(date +%s) - START )) s"
( timeout 240 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py; echo "memcheck rc=$?" ) > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1
tail -5 gpurun_out/sanitizer_memcheck_$TAG.log
