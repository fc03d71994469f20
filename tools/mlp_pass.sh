#!/bin/bash
# MLP ranker measurement pass: ncu --set full of both scorer kernels at the c4mlp shape (summaries only),
# the event trace of the backward kernel, the probe timings.  mlp_pass.sh tag
tag=${1:-r02w}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mlp_ -c 6 -o gpurun_out/${tag}_prof_c4mlp \
    python tools/mlp_probe.py prof > gpurun_out/${tag}_ncu_c4mlp.log 2>&1
rep=gpurun_out/${tag}_prof_c4mlp.ncu-rep
python profiles/summarize_ncu.py $rep > gpurun_out/${tag}_ncu_full_c4mlp.txt 2>/dev/null
python tools/ncu_lines.py $rep 50 > gpurun_out/${tag}_ncu_source_hotlines_c4mlp.txt 2>/dev/null
rm -f $rep
python tools/mlp_trace.py > gpurun_out/${tag}_mlp_trace.txt 2>&1
python tools/mlp_probe.py > gpurun_out/${tag}_mlp_probe_fwd.txt 2>&1
python tools/mlp_probe.py bwd 2>&1 | grep -v "bad\|Warn\|return Var" > gpurun_out/${tag}_mlp_probe_bwd.txt
python tools/umma_probe.py > gpurun_out/${tag}_umma_layout_probe.txt 2>&1
grep -E "kernel:|gpu__time_duration|dram__bytes_read|dram__bytes_write|pipe_tensor" gpurun_out/${tag}_ncu_full_c4mlp.txt
