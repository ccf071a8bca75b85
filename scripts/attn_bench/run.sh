#!/bin/bash
cd "$(dirname "$0")"
out=../../gpurun_out/${1:-r06b}_attn_bench.txt
{
for v in base pf16 pf32 n11 n11pf; do for cfg in "512 3" "512 1" "150 3"; do echo "== $v B,k = $cfg"; timeout 60 ./attn_bench_$v $cfg 30 2>&1 | grep -E "new attn2|PARITY|rror|watchdog" | sort | uniq -c | sort -rn | head -4; done; done
} > $out 2>&1
cat $out
