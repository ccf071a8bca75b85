#!/bin/bash
# experiment: encoder chunking / pool kernel / attention v3 (768 vs 384 threads)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then exit 1; fi
run() { tag=$1; shift; timeout 120 python bench.py --no-cpu-baseline --breakdown "$@" > gpurun_out/exp1_$tag.json 2> gpurun_out/exp1_$tag.err || { echo "== $tag FAILED"; tail -3 gpurun_out/exp1_$tag.err; return; }; echo "== $tag"; cat gpurun_out/exp1_$tag.err | tr '\n' ';' ; python -c "import json,sys; d=json.load(open('gpurun_out/exp1_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; }
run default
COMIC_B200_LIB=$PWD/comic-compact-image-captioning-with-attention_b200/libcomic_b200_a384.so run a384
run c64 --opt enc_chunk_stem=64 --opt enc_chunk_28=64 --opt enc_chunk_14=64
run c16_128_512 --opt enc_chunk_stem=16 --opt enc_chunk_28=128 --opt enc_chunk_14=512
run c64_256_512 --opt enc_chunk_stem=64 --opt enc_chunk_28=256 --opt enc_chunk_14=512
