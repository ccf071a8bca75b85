#!/bin/bash
# Final GPU pass of a round: gpu tests, bench lines, launch list and ncu --set full of the two top kernels.
#   gpurun --timeout 1500 -- 'bash scripts/prof_final.sh r01z'
T=${1:-rXX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --breakdown > gpurun_out/${T}_bench512.json 2> gpurun_out/${T}_bench512.err; cut -c1-200 gpurun_out/${T}_bench512.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_reference_arm.json 2>/dev/null; cut -c1-160 gpurun_out/${T}_reference_arm.json
timeout 200 python bench.py --batch 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench8.json 2>/dev/null; cut -c1-160 gpurun_out/${T}_bench8.json
timeout 300 python bench.py --workload word --no-cpu-baseline > gpurun_out/${T}_bench_word512.json 2>/dev/null; cut -c1-160 gpurun_out/${T}_bench_word512.json
for m in decoder cnn_finetune scst; do b=32; if [ $m = scst ]; then b=10; fi
  timeout 200 python scripts/bench_train.py --mode $m --batch $b --steps 10 --warmup 3 > gpurun_out/${T}_train_${m}.json 2>/dev/null; cut -c1-200 gpurun_out/${T}_train_${m}.json; done
B="python bench.py --batch 512 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -s 1400 -c 800 --csv --log-file gpurun_out/${T}_launches_batch512.csv $B > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fused -s 70 -c 1 -o gpurun_out/${T}_attn $B > gpurun_out/${T}_ncu_attn.log 2>&1
B="python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 0 -c 8 -o gpurun_out/${T}_conv $B > gpurun_out/${T}_ncu_conv.log 2>&1
ls -la gpurun_out | grep ${T}
