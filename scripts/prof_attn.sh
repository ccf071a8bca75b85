mkdir -p gpurun_out
B="python bench.py --batch 512 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fused -s 70 -c 2 -o gpurun_out/prof_attn $B > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
