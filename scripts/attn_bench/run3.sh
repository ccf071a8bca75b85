#!/bin/bash
cd "$(dirname "$0")"
out=../../gpurun_out/${1:-r04e}_attn_flaky.txt
{
echo "== streampf96 k=3 (full output)"; timeout 60 ./attn_bench_streampf96 512 3 5 2>&1 | tail -8
for i in 1 2 3 4 5 6; do echo "== pf32 k=1 run $i"; timeout 60 ./attn_bench_pf32 512 1 20 2>&1 | grep -E "ctx:|hist:|PARITY|error|watchdog"; done
for i in 1 2 3 4 5 6; do echo "== base k=1 run $i"; timeout 60 ./attn_bench_base 512 1 20 2>&1 | grep -E "ctx:|hist:|PARITY|error|watchdog"; done
for i in 1 2 3; do echo "== base k=3 B=37 run $i"; timeout 60 ./attn_bench_base 37 3 20 2>&1 | grep -E "ctx:|hist:|PARITY|error|watchdog"; done
} > $out 2>&1
cat $out
