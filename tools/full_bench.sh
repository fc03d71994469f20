#!/bin/bash
# Full measurement pass on the GPU box: default bench line, reference arm, other configs,
# ncu launch list and full capture of the dominant kernel.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 2000 --warmup 200 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${tag}_bench_c2_reference.json 2>&1
python bench.py --steps 2000 --warmup 200 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench_c2_eager.json 2>&1
for c in ns c3 c4 c4m c4a c4d c5; do
  steps=300; [ $c = c5 ] && steps=30
  python bench.py --config $c --steps $steps --warmup 20 --no-e2e > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c2.csv \
    python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_warp -s 5 -c 1 -o gpurun_out/${tag}_prof_c2 \
    python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_cta -s 3 -c 1 -o gpurun_out/${tag}_prof_ns \
    python bench.py --config ns --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_full_ns.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:listnet -s 3 -c 1 -o gpurun_out/${tag}_prof_c4 \
    python bench.py --config c4 --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_full_c4.log 2>&1
ls -la gpurun_out | grep ${tag}_ | head -40
