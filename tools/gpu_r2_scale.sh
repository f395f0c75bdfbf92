#!/usr/bin/env bash
# round 2: bench.py on N GPUs of one box the way the driver launches it; prints the headline and every config
N=${1:-2}; TAG=${2:-n$N}
mkdir -p gpurun_out
START=$(date +%s)
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
fi
echo "rc=$? wall=$(( $(date +%s) - START )) s"
tail -4 gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("headline", round(d["value"]), d["roofline"]["frac"], "e2e", d["e2e"]["value"], "copy-only", d["e2e"]["copy_only_msamples_s"])
for c in d["configs"]:
    print(c["name"], c.get("error"), c.get("ms"), c.get("value"), (c.get("roofline") or {}).get("frac"), (c.get("roofline") or {}).get("frac_mean"), c.get("gather"))
PY
