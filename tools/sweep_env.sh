#!/bin/bash
# A-B timing of scheduling variants on one config: sweep_env.sh cfg "VAR=val ..." "VAR=val ..." ...
cfg=$1; shift
for envs in "$@"; do
  steps=300; [ $cfg = c2 ] && steps=2000; [ $cfg = c5 ] && steps=30
  env $envs python bench.py --config $cfg --steps $steps --warmup 20 --no-e2e --no-cpu-baseline > /tmp/sweep.json 2>/tmp/sweep.err
  python - "$cfg $envs" <<'PY'
import json,sys
try:
    d=json.loads(open('/tmp/sweep.json').read().strip().splitlines()[-1])
    r=d['roofline']; i=d.get('issue_roofline') or {}
    print('%-40s ms/step %.5f kernel_ms %.5f issue %.4f'%(sys.argv[1],d['ms_per_step'],r['kernel_ms'],i.get('frac') or 0))
except Exception as e: print(sys.argv[1],'ERR',e, open('/tmp/sweep.err').read()[-1500:])
PY
done
