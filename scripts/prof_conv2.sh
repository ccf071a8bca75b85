mkdir -p gpurun_out
B="python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16x3_kernel<256, 2, 1>" -s 4 -c 1 -o gpurun_out/prof_conv256 $B > gpurun_out/ncu_conv256.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16x3_kernel<256, 2, 0>" -s 12 -c 1 -o gpurun_out/prof_gemm256 $B > gpurun_out/ncu_gemm256.log 2>&1
tail -2 gpurun_out/ncu_conv256.log gpurun_out/ncu_gemm256.log
