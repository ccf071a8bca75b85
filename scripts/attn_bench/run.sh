#!/bin/bash
cd "$(dirname "$0")"
out=../../gpurun_out/${1:-r05c}_attn_bench.txt
{
echo "== sanitizer B=5 k=3"; timeout 300 compute-sanitizer --tool memcheck ./attn_bench_base 5 3 1 2>&1 | grep -E "PARITY|ERROR SUMMARY|rror" | head -5
for cfg in "1 3" "3 1" "25 3" "37 3" "150 3" "512 3" "512 1" "512 2" "444 3" "511 3" "512 3"; do echo "== base B,k = $cfg"; timeout 60 ./attn_bench_base $cfg 30 2>&1 | grep -E "old fused|new attn2|ctx:|hist:|PARITY|rror|watchdog" | sort | uniq -c | sort -rn | head -8; done
echo "== trace k=3"; timeout 60 ./attn_bench_trace 512 3 20 2>&1 | grep -E "new attn2|trace|score warp|ctx warp|finaliser|PARITY|rror|watchdog"| head -8
} > $out 2>&1
cat $out
