#!/bin/bash
# One GPU-box pass: parity tests, bench lines for the BASELINE configs, ncu launch list and
# full captures of the two dominant kernels.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r01c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 2000 --warmup 200 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${tag}_bench_c2_reference.json 2>&1
for c in ns c3 c4 c4x c4m c4f c5; do
  steps=300; [ $c = c5 ] && steps=30; [ $c = c4f ] && steps=30
  python bench.py --config $c --steps $steps --warmup 20 --no-e2e > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c2.csv \
    python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_warp -s 5 -c 1 -o gpurun_out/${tag}_prof_c2 \
    python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_ring -s 3 -c 1 -o gpurun_out/${tag}_prof_ns \
    python bench.py --config ns --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_full_ns.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:topk_metrics -s 3 -c 1 -o gpurun_out/${tag}_prof_c4m \
    python bench.py --config c4m --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_full_c4m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:listnet -s 3 -c 1 -o gpurun_out/${tag}_prof_c4 \
    python bench.py --config c4 --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_full_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches_ns.csv \
    python bench.py --config ns --steps 10 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_launch_ns.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:linear_listnet_kernel -s 2 -c 1 -o gpurun_out/${tag}_prof_c4f \
    python bench.py --config c4f --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_full_c4f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hinge_sorted -s 3 -c 1 -o gpurun_out/${tag}_prof_c3 \
    python bench.py --config c3 --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_full_c3.log 2>&1
python tools/collate_probe.py > gpurun_out/${tag}_collate_probe.txt 2>&1
python tools/e2e_probe.py > gpurun_out/${tag}_e2e_probe.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.txt 2>&1; tail -1 gpurun_out/${tag}_smoke.txt
# summarise the captures on the box and drop the reports (gpurun copies back at most 64 MiB)
for c in c2 ns c4m c4 c4f c3; do
  rep=gpurun_out/${tag}_prof_$c.ncu-rep
  [ -f $rep ] || continue
  python profiles/summarize_ncu.py $rep > gpurun_out/${tag}_ncu_full_$c.txt 2>/dev/null
  python tools/ncu_lines.py $rep 40 > gpurun_out/${tag}_ncu_source_hotlines_$c.txt 2>/dev/null
  rm -f $rep
done
for f in gpurun_out/${tag}_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get('roofline') or {}; i=d.get('issue_roofline') or {}
    print(' value %.4g %s ms/step %.4g kernel_ms %s hbm %s issue %s e2e %s parity %s'%(d['value'],d['unit'],d['ms_per_step'],r.get('kernel_ms'),r.get('frac'),i.get('frac'),(d.get('e2e') or {}).get('value'),d.get('parity_spot_check')))
except Exception as e: print(' ERR',e)
PY
done
