#!/bin/bash
# usage (on the GPU box): bash tools/gpu_cycle_r02.sh <tag>   — the round's evidence into gpurun_out/<tag>_*
tag=${1:-r02}
o=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $o/${tag}_pytest_gpu.log
python bench.py > $o/${tag}_bench_8192.log 2>&1 && tail -1 $o/${tag}_bench_8192.log > $o/${tag}_bench_8192.json
python bench.py --config 3 --no-rollout > $o/${tag}_bench_config3.log 2>&1 && tail -1 $o/${tag}_bench_config3.log > $o/${tag}_bench_config3.json
python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > $o/${tag}_bench_reference_arm.json
python bench.py --impl reference --config 3 --steps 20 --warmup 5 2>/dev/null | tail -1 > $o/${tag}_bench_reference_arm_config3.json
python tools/probe.py 8192 > $o/${tag}_probes_8192.txt 2>&1
NOFLUSH=1 python tools/probe.py 8192 > $o/${tag}_probes_8192_noflush.txt 2>&1
python tools/probe.py 65536 > $o/${tag}_probes_65536.txt 2>&1
python tools/e2e_probe.py 8192 2000 2>&1 | tail -3 > $o/${tag}_e2e_breakdown.txt
python tools/sustained.py 8192 6 2>&1 | tail -4 > $o/${tag}_sustained_8192.txt
# ncu: launch list of the bench command, then full captures of one step (never a bench number under ncu)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_ncu_launches_8192.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-rollout --no-sweep > $o/ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"post_physics|reset_observe|gae" --launch-skip 8 --launch-count 6 -f \
    -o $o/${tag}_ncu_8192 python tools/ncu_step.py 8192 > $o/ncu_8192.log 2>&1
STEPS=6 ncu --set full --import-source on --clock-control none -k regex:"post_physics|reset_observe|gae" --launch-skip 6 --launch-count 6 -f \
    -o $o/${tag}_ncu_65536 python tools/ncu_step.py 65536 > $o/ncu_65536.log 2>&1
STEPS=6 ncu --set full --import-source on --clock-control none -k regex:"heights|post_physics|reset_observe" --launch-skip 9 --launch-count 3 -f \
    -o $o/${tag}_ncu_config3_8192 python tools/ncu_step.py 8192 config3 > $o/ncu_config3.log 2>&1
ls -la $o | tail -20
