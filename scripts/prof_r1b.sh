#!/bin/bash
# Round-1 (session b) GPU pass: gpu tests, bench lines, launch list, ncu --set full of the top kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --breakdown > gpurun_out/bench_b512.json 2> gpurun_out/bench_b512.err
tail -2 gpurun_out/bench_b512.json
timeout 300 python bench.py --batch 8 --steps 20 --warmup 5 --breakdown --no-cpu-baseline > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err
tail -1 gpurun_out/bench_b8.json
timeout 300 python bench.py --workload word --breakdown --no-cpu-baseline > gpurun_out/bench_word.json 2> gpurun_out/bench_word.err
tail -1 gpurun_out/bench_word.json
timeout 300 python scripts/bench_train.py --mode decoder --batch 32 --steps 20 > gpurun_out/bench_train_decoder.json 2> gpurun_out/bench_train_decoder.err
tail -1 gpurun_out/bench_train_decoder.json
B="python bench.py --batch 128 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1b.csv $B > gpurun_out/ncu_launch.log 2>&1
B="python bench.py --batch 512 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fused -s 70 -c 2 -o gpurun_out/prof_attn $B > gpurun_out/ncu_attn.log 2>&1
B="python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 2 -c 8 -o gpurun_out/prof_conv $B > gpurun_out/ncu_conv.log 2>&1
ls -la gpurun_out
