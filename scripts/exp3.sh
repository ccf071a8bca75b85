#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then exit 1; fi
run() { tag=$1; shift; timeout 120 python bench.py --no-cpu-baseline --breakdown "$@" > gpurun_out/exp3_$tag.json 2> gpurun_out/exp3_$tag.err || { echo "== $tag FAILED"; tail -3 gpurun_out/exp3_$tag.err; return; }; echo "== $tag"; cat gpurun_out/exp3_$tag.err | tr '\n' ';' ; python -c "import json,sys; d=json.load(open('gpurun_out/exp3_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['decoder_step_us'])"; }
run b512
run b8 --batch 8 --steps 20 --warmup 5
