#!/bin/bash
# One GPU pass: full suite, bench with the weight-multicast variants, small-batch attention harness.
tag=${1:-r08c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.txt 2>&1; tail -4 gpurun_out/${tag}_pytest_gpu.txt
for mc in 0 2 4; do
  python bench.py --no-train --no-cpu-baseline --opt gemm_mc=$mc > gpurun_out/${tag}_bench512_mc$mc.json 2> gpurun_out/${tag}_bench512_mc$mc.err
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['value'], d['e2e']['value'], d['kernel_ms_per_step'])" gpurun_out/${tag}_bench512_mc$mc.json
  tail -c 300 gpurun_out/${tag}_bench512_mc$mc.err
done
cd scripts/attn_bench
for cfg in "8 3 30 1" "12 3 30 1" "16 3 30 1" "20 3 30 1" "25 3 30 1" "32 3 30 1"; do echo "== $cfg"; timeout 60 ./attn_bench_base $cfg 2>&1 | grep -E "old fused|new attn2|FAIL" ; done > ../../gpurun_out/${tag}_attn_small_batch.txt 2>&1
cat ../../gpurun_out/${tag}_attn_small_batch.txt
