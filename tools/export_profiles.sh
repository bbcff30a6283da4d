#!/bin/sh
# Copies the judged summaries of a capture set gpurun_out/<tag>_*.ncu-rep into profiles/ (tracked):
#   <tag>_<kernel>_details.csv  = ncu --page details, <tag>_<kernel>_raw.txt = the headline raw counters,
#   <tag>_launches.csv          = the launch list of the same bench command
# usage: sh tools/export_profiles.sh r1h gmm_tc2 beta_l2r alpha_l2r stats5
tag=$1; shift
for k in "$@"; do
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page details --csv > profiles/${tag}_${k}_details.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${tag}_$k.ncu-rep 8 > profiles/${tag}_${k}_raw.txt 2>/dev/null
done
cp gpurun_out/${tag}_launches.csv profiles/${tag}_launches.csv
