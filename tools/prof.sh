#!/bin/bash
# ncu --set full of one kernel of one bench config, summarised on the box: prof.sh tag cfg kernel-regex [skip]
tag=$1; cfg=$2; kre=$3; skip=${4:-3}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -o gpurun_out/${tag}_prof_$cfg \
    python bench.py --config $cfg --steps 3 --warmup 3 --windows 0 --no-graph --no-cpu-baseline --no-e2e --no-sub > gpurun_out/${tag}_ncu_$cfg.log 2>&1
rep=gpurun_out/${tag}_prof_$cfg.ncu-rep
python profiles/summarize_ncu.py $rep > gpurun_out/${tag}_ncu_full_$cfg.txt 2>/dev/null
python tools/ncu_lines.py $rep 60 > gpurun_out/${tag}_ncu_source_hotlines_$cfg.txt 2>/dev/null
python tools/ncu_lines.py $rep 90 inst > gpurun_out/${tag}_ncu_source_instlines_$cfg.txt 2>/dev/null
rm -f $rep
head -3 gpurun_out/${tag}_ncu_full_$cfg.txt
