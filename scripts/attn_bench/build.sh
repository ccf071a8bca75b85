#!/bin/bash
# Build the development harness variants (binaries are git-ignored):
#   base        shipped configuration
#   trace       clock64 stamps per warp and phase (COMIC_A2_TRACE)
#   k1 k2 k4 .. knock-out timings (COMIC_A2_KNOCK bit 0: no MUFU, bit 2: pass 2 skipped; results are wrong by design)
#   s2          11 score + 2 statistics warps
cd "$(dirname "$0")"
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
nvcc $F -o attn_bench_base attn_bench.cu &
nvcc $F -DCOMIC_A2_TRACE=1 -o attn_bench_trace attn_bench.cu &
nvcc $F -DCOMIC_A2_NSW=11 -DCOMIC_A2_STATW=2 -o attn_bench_s2 attn_bench.cu &
for k in 1 4; do nvcc $F -DCOMIC_A2_KNOCK=$k -o attn_bench_k$k attn_bench.cu & done
wait
