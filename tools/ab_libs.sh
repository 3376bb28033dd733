#!/bin/bash
# usage (GPU box): tools/ab_libs.sh "<lib1> <lib2> ..." [envs] [extra bench args]  — bench.py with each library (TI5_LIB), interleaved twice
libs=$1; n=${2:-8192}; shift 2
for rep in 1 2; do for lib in $libs; do
  TI5_LIB=$PWD/$lib python bench.py --steps 96 --warmup 24 --no-cpu-baseline --no-rollout --no-sweep --envs $n "$@" | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$lib envs=$n', round(d['value']/1e6,1),'M', round(d['ms_per_step']*1e3,2),'us e2e',round((d['e2e'] or {}).get('value',0)/1e6,1), {k:round(v['phase_ms']*1e3,1) for k,v in d['roofline']['kernels'].items()})"
done; done
