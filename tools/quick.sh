#!/bin/bash
# usage (GPU box): tools/quick.sh [sizes...]  — GPU tests, then working-tree library vs tools/exp/libti5_base.so at each size
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for n in "${@:-8192}"; do
for lib in tools/exp/libti5_base.so ti5_isaacgym_b200/libti5step.so; do
  TI5_LIB=$PWD/$lib python bench.py --steps 96 --warmup 24 --no-cpu-baseline --no-rollout --no-sweep --envs $n 2>&1 | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$lib envs=$n', round(d['value']/1e6,1),'M', round(d['ms_per_step']*1e3,2),'us e2e',round(d['e2e']['value']/1e6,1), {k:round(v['phase_ms']*1e3,1) for k,v in d['roofline']['kernels'].items()})"
done; done
