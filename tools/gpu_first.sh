#!/usr/bin/env bash
# First contact with the B200 box: environment probe, pipe micro-benchmarks, GPU parity tests, smoke, bench.
mkdir -p gpurun_out
{
  echo "== env"; ls -d /root/reference oracle/_ref 2>&1 | head; ls oracle/_ref | head -30
  nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv
  python -c "import torch;print(torch.__version__, torch.cuda.is_available(), torch.cuda.device_count())"
} > gpurun_out/env.txt 2>&1
timeout 120 profiles/microbench/mb > gpurun_out/mb.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?" >> gpurun_out/bench1.err
tail -5 gpurun_out/pytest_gpu.txt; cat gpurun_out/smoke.txt | tail -3; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
