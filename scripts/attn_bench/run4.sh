#!/bin/bash
cd "$(dirname "$0")"
out=../../gpurun_out/${1:-r04h}_attn_san.txt
{
echo "== plain B=37 k=3"; timeout 60 ./attn_bench_base 37 3 1 2>&1 | grep -v "watchdog: block" | tail -6
echo "== memcheck B=37 k=3"; timeout 300 compute-sanitizer --tool memcheck ./attn_bench_base 37 3 1 2>&1 | grep -v "watchdog: block" | tail -12
echo "== racecheck B=37 k=3"; timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis ./attn_bench_base 37 3 1 2>&1 | grep -v "watchdog: block" | tail -40
echo "== initcheck B=37 k=3"; timeout 300 compute-sanitizer --tool initcheck ./attn_bench_base 37 3 1 2>&1 | grep -v "watchdog: block" | tail -12
} > $out 2>&1
cat $out
