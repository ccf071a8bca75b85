#!/bin/bash
cd "$(dirname "$0")"
out=../../gpurun_out/${1:-r04d}_attn_variants.txt
{
for v in base stream pf32 pf96 streampf96; do
  for k in 3 1; do
    echo "== $v k=$k"; timeout 60 ./attn_bench_$v 512 $k 50 | grep -E "new attn2|PARITY"
  done
done
} > $out 2>&1
cat $out
