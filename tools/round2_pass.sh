#!/bin/bash
# Round-2 GPU pass: ncu --set full of the dominant kernel of every bench config (summaries only),
# launch list of the default bench command, the issue-rate micro-benchmark.  round2_pass.sh tag
tag=${1:-r02}
mkdir -p gpurun_out
tools/issue_peak > gpurun_out/${tag}_issue_peaks.json 2>&1
for spec in "c5 pair_ring 3" "ns pair_ring 3" "c2 pair_warp 5" "c3 hinge_sorted 3" "c4 listnet 3" "c4m topk_metrics 3" "c4a rank_metrics_warp 3" "c4d rank_metrics_warp 3"; do
  set -- $spec
  tools/prof.sh $tag $1 $2 $3 > /dev/null 2>&1
  head -2 gpurun_out/${tag}_ncu_full_$1.txt | tail -1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches_c5.csv \
    python bench.py --steps 4 --warmup 3 --windows 0 --no-graph --no-cpu-baseline --no-e2e --no-sub > gpurun_out/${tag}_ncu_launch_c5.log 2>&1
python profiles/summarize_ncu.py --launches gpurun_out/${tag}_launches_c5.csv > gpurun_out/${tag}_launches_c5.txt 2>/dev/null
rm -f gpurun_out/${tag}_launches_c5.csv
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.txt 2>&1; tail -1 gpurun_out/${tag}_smoke.txt
