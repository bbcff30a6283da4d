#!/bin/sh
# Copies the judged summaries of a capture set gpurun_out/<tag>_*.ncu-rep into profiles/ (tracked):
#   <tag>_<kernel>_details.csv  = ncu --page details, <tag>_<kernel>_raw.txt = the headline raw counters + hottest source
#   lines (tools/ncu_summary.py), <tag>_launches.csv = the launch list of the same bench command
# usage: sh tools/export_profiles.sh r2 gmm_tc3 stats_tc stats_pre beta_l2r_warp alpha_l2r
tag=$1; shift
for k in "$@"; do
  [ -f gpurun_out/${tag}_$k.ncu-rep ] || { echo "missing gpurun_out/${tag}_$k.ncu-rep"; continue; }
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page details --csv > profiles/${tag}_${k}_details.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${tag}_$k.ncu-rep 8 > profiles/${tag}_${k}_raw.txt 2>/dev/null
  python tools/ncu_lines.py gpurun_out/${tag}_$k.ncu-rep 12 >> profiles/${tag}_${k}_raw.txt 2>/dev/null
done
[ -f gpurun_out/${tag}_launches.csv ] && cp gpurun_out/${tag}_launches.csv profiles/${tag}_launches.csv
