#!/bin/bash
# Final GPU pass of a round: gpu tests, bench lines, launch list and ncu --set full of the streaming attention kernel.
#   gpurun --timeout 1700 -- 'bash scripts/prof_final.sh r09z'
T=${1:-rXX}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.txt
timeout 500 python bench.py --breakdown > gpurun_out/${T}_bench512.json 2> gpurun_out/${T}_bench512.err; cut -c1-200 gpurun_out/${T}_bench512.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_reference_arm.json 2>/dev/null; cut -c1-160 gpurun_out/${T}_reference_arm.json
timeout 200 python bench.py --batch 8 --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/${T}_bench8.json 2>/dev/null; cut -c1-160 gpurun_out/${T}_bench8.json
timeout 200 python bench.py --batch 25 --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/${T}_bench25.json 2>/dev/null; cut -c1-160 gpurun_out/${T}_bench25.json
timeout 300 python bench.py --workload word --no-cpu-baseline --no-train > gpurun_out/${T}_bench_word512.json 2>/dev/null; cut -c1-160 gpurun_out/${T}_bench_word512.json
# launch list of the bench command (graph replay off so that every launch is a stream launch; durations are cold-cache and
# serialised: shares, not absolutes, are comparable with the bench line)
B="python bench.py --batch 512 --steps 1 --warmup 1 --no-cpu-baseline --no-train"
COMIC_B200_INFER_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -s 1500 -c 800 --csv --log-file gpurun_out/${T}_launches_batch512.csv $B > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn2_kernel -s 20 -c 1 -o gpurun_out/${T}_attn2 python scripts/decode_only.py 512 30 > gpurun_out/${T}_ncu_attn2.log 2>&1
ls -la gpurun_out | grep ${T}
