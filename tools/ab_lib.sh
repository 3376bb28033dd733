#!/bin/bash
# usage (GPU box): tools/ab_lib.sh [envs]  — working-tree library vs tools/exp/libti5_base.so, interleaved twice
n=${1:-8192}
for rep in 1 2; do for lib in tools/exp/libti5_base.so ti5_isaacgym_b200/libti5step.so; do
  TI5_LIB=$PWD/$lib python bench.py --steps 96 --warmup 24 --no-cpu-baseline --no-rollout --no-sweep --envs $n | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$lib envs=$n', round(d['value']/1e6,1),'M', round(d['ms_per_step']*1e3,2),'us e2e',round((d['e2e'] or {}).get('value',0)/1e6,1), {k:round(v['phase_ms']*1e3,1) for k,v in d['roofline']['kernels'].items()})"
done; done
