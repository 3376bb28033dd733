#!/bin/bash
# usage: tools/gpu_cycle.sh <tag>   (runs on the GPU box): tests, bench, ncu launch list
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 96 --warmup 24 --no-cpu-baseline --no-sweep > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench rc=$?; tail -3 gpurun_out/bench_$tag.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sweep > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sweep > gpurun_out/ncu.log 2>&1; echo ncu rc=$?
python bench.py --steps 48 --warmup 12 --no-cpu-baseline --no-sweep --envs 65536 > gpurun_out/bench_${tag}_64k.json 2> gpurun_out/bench_${tag}_64k.err; echo bench64k rc=$?; tail -2 gpurun_out/bench_${tag}_64k.err
