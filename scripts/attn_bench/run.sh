#!/bin/bash
cd "$(dirname "$0")"
out=../../gpurun_out/${1:-r04i}_attn_bench.txt
{
for cfg in "150 3" "512 3" "512 1" "512 2" "444 3" "37 3" "1 3" "512 3" "512 3"; do echo "== base B,k = $cfg"; timeout 60 ./attn_bench_base $cfg 30 2>&1 | grep -E "old fused|new attn2|ctx:|hist:|PARITY|rror|watchdog" | sort | uniq -c | sort -rn | head -8; done
echo "== stream k=3"; timeout 60 ./attn_bench_stream 512 3 30 2>&1 | grep -E "new attn2|rror|watchdog" | head -3
echo "== stream k=1"; timeout 60 ./attn_bench_stream 512 1 30 2>&1 | grep -E "new attn2|rror|watchdog"| head -3
echo "== trace k=3"; timeout 60 ./attn_bench_trace 512 3 20 2>&1 | grep -E "new attn2|trace|score warp|ctx warp|PARITY|rror|watchdog"| head -8
echo "== trace k=1"; timeout 60 ./attn_bench_trace 512 1 20 2>&1 | grep -E "new attn2|trace|score warp|ctx warp|PARITY|rror|watchdog"| head -8
} > $out 2>&1
cat $out
