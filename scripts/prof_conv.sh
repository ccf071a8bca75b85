mkdir -p gpurun_out
B="python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline"
# launch list: last full step (skip bind + 4 earlier steps)
timeout 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_bf16x3|maxpool|pad_c3|avgpool" --csv --log-file gpurun_out/launches_conv.csv $B > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
