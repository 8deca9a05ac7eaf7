#!/bin/bash
# Developer tool (GPU box): one `ncu --set full` capture of a steady-state frame; keeps only what fits gpurun's 64 MiB return
# channel: the text summary (tools/ncu_summary.py) and, per hot kernel, the per-source-line table (CUDA view, gzip'd).
#   tools/ncu_capture.sh <tag> [profile_frame.py arguments...]     ->  gpurun_out/<tag>_summary.txt, gpurun_out/<tag>_<kernel>_lines.csv.gz
tag=$1; shift
rep=/tmp/${tag}.ncu-rep
ncu --set full --import-source on --clock-control none --profile-from-start off -o /tmp/${tag} -f python tools/profile_frame.py "$@" > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
python tools/ncu_summary.py $rep --top 40 --buckets tools/ncu_buckets.json --dump-lines gpurun_out/${tag}_lines > gpurun_out/${tag}_summary.txt 2>&1
ncu -i $rep --page raw --csv 2>/dev/null | gzip > gpurun_out/${tag}_raw.csv.gz
ls -la gpurun_out/${tag}_*
