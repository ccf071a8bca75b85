#!/bin/bash
# Build the development harness variants (binaries are git-ignored): base, trace (clock64 stamps), stream (score math compiled out)
cd "$(dirname "$0")"
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
nvcc $F -o attn_bench_base attn_bench.cu &
nvcc $F -DCOMIC_A2_TRACE=1 -o attn_bench_trace attn_bench.cu &
wait
