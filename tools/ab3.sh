#!/bin/bash
# usage (GPU box): tools/ab3.sh VAR "v1 v2 v3" [envs] [extra bench args]  — bench.py under VAR=v for every v, interleaved twice
var=$1; vals=$2; n=${3:-8192}; shift 3
for rep in 1 2; do for v in $vals; do
  env $var=$v python bench.py --steps 96 --warmup 24 --no-cpu-baseline --no-rollout --no-sweep --envs $n "$@" | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$var=$v envs=$n', round(d['value']/1e6,1),'M', round(d['ms_per_step']*1e3,2),'us e2e',round((d['e2e'] or {}).get('value',0)/1e6,1), {k:round(v['phase_ms']*1e3,1) for k,v in d['roofline']['kernels'].items()}, d['clocks']['sm_mhz'])"
done; done
