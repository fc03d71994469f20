#!/bin/bash
# Quick GPU pass: parity tests + a few bench lines (no cpu baseline / e2e).  Usage: quick_pass.sh tag cfg...
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${tag}_pytest.log
for c in "$@"; do
  steps=300; [ $c = c5 ] && steps=30; [ $c = c2 ] && steps=2000
  python bench.py --config $c --steps $steps --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
  python - gpurun_out/${tag}_bench_$c.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get('roofline') or {}; i=d.get('issue_roofline') or {}
    print(sys.argv[1],' value %.4g ms/step %.4g kernel_ms %s hbm %.4f issue %s parity %s'%(d['value'],d['ms_per_step'],r.get('kernel_ms'),r.get('frac'),i.get('frac'),d.get('parity_spot_check')))
except Exception as e: print(sys.argv[1],' ERR',e); print(open(sys.argv[1][:-4]+'err').read()[-2000:])
PY
done
