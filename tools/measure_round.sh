#!/bin/sh
# One gpurun call that produces everything profiles/README.md and DESIGN.md section 5 quote for a kernel set:
#   sh tools/measure_round.sh <tag>       (on the GPU box, from the repo root)
# The .ncu-rep files (20 MB each) are turned into the text summaries profiles/ keeps ON THE BOX and deleted: gpurun only
# brings 64 MiB back.  Afterwards, here:  cp gpurun_out/<tag>_*_details.csv gpurun_out/<tag>_*_raw.txt gpurun_out/<tag>_launches.csv profiles/
tag=${1:-rX}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --also "" --tool-files 0 > gpurun_out/${tag}_launch_bench.log 2>&1
cap() {   # cap <kernel regex> <name> <bench args...>
   k=$1; n=$2; shift 2
   timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -o gpurun_out/${tag}_$n -f python bench.py --steps 2 --warmup 3 --no-cpu --also "" --tool-files 0 "$@" > gpurun_out/${tag}_ncu_$n.log 2>&1
   if [ -f gpurun_out/${tag}_$n.ncu-rep ]; then
      ncu -i gpurun_out/${tag}_$n.ncu-rep --page details --csv > gpurun_out/${tag}_${n}_details.csv 2>/dev/null
      python tools/ncu_summary.py gpurun_out/${tag}_$n.ncu-rep > gpurun_out/${tag}_${n}_raw.txt 2>/dev/null
      python tools/ncu_lines.py gpurun_out/${tag}_$n.ncu-rep 14 >> gpurun_out/${tag}_${n}_raw.txt 2>/dev/null
      rm -f gpurun_out/${tag}_$n.ncu-rep
   fi
}
for k in gmm_tc4 stats_tc stats_pre beta_l2r_warp alpha_l2r; do cap ${k}_kernel $k; done
cap gmm_tc4_kernel gmm_tc4_cfg2 --workload cfg2
ls -la gpurun_out/
