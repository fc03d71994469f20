#!/bin/bash
# ncu --set full of the MLP scorer kernels: the c4mlp shape (136 features) and Yahoo's width (700, streamed W1 /
# column slabs); summaries only.  mlp_pass2.sh tag
tag=${1:-r02zz}
mkdir -p gpurun_out
for mode in prof prof_wide; do
  rep=gpurun_out/${tag}_${mode}.ncu-rep
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_ -c 6 -o gpurun_out/${tag}_${mode} \
      python tools/mlp_probe.py $mode > gpurun_out/${tag}_ncu_${mode}.log 2>&1
  python profiles/summarize_ncu.py $rep > gpurun_out/${tag}_ncu_full_mlp_${mode}.txt 2>/dev/null
  python tools/ncu_lines.py $rep 40 > gpurun_out/${tag}_ncu_source_hotlines_mlp_${mode}.txt 2>/dev/null
  rm -f $rep
  grep -E "kernel:|gpu__time_duration|dram__bytes_read.sum |dram__bytes_write.sum |pipe_tensor|dram__throughput" gpurun_out/${tag}_ncu_full_mlp_${mode}.txt
done
