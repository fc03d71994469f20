#!/bin/bash
# A-B timing of library variants: ab.sh "cfg1 cfg2" lib1.so lib2.so ...   ("default" = the in-tree build)
cfgs=$1; shift
for lib in "$@"; do
  for cfg in $cfgs; do
    steps=200; [ $cfg = c5 ] && steps=30; [ $cfg = c2 ] && steps=2000
    if [ "$lib" = default ]; then unset LTR_SM100_LIB; else export LTR_SM100_LIB=$PWD/$lib; fi
    python bench.py --config $cfg --steps $steps --warmup 10 --windows 2 --no-e2e --no-cpu-baseline --no-sub > /tmp/ab.json 2>/tmp/ab.err
    python - "$cfg $lib" <<'PY'
import json,sys
try:
    d=json.loads(open('/tmp/ab.json').read().strip().splitlines()[-1])
    r=d['roofline']; i=d.get('issue_roofline') or {}
    print('%-40s ms/step %.5f kernel_ms %.5f issue %.4f parity %s'%(sys.argv[1],d['ms_per_step'],r['kernel_ms'],i.get('frac') or 0,(d.get('parity_spot_check') or {}).get('ok')))
except Exception as e: print(sys.argv[1],'ERR',e, open('/tmp/ab.err').read()[-1500:])
PY
  done
done
