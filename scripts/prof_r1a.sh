mkdir -p gpurun_out
B="python bench.py --batch 128 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 900 --csv --log-file gpurun_out/launches_r1a.csv $B > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 40 -c 4 -o gpurun_out/prof_tc $B > gpurun_out/ncu_tc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_scores -s 5 -c 2 -o gpurun_out/prof_scores $B > gpurun_out/ncu_scores.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ctx -s 5 -c 2 -o gpurun_out/prof_ctx $B > gpurun_out/ncu_ctx.log 2>&1
timeout 120 python bench.py --batch 8 --steps 20 --warmup 5 --breakdown > gpurun_out/bench_b8.log 2>&1
tail -3 gpurun_out/bench_b8.log
ls -la gpurun_out
