#!/bin/sh
# One gpurun call that produces everything profiles/README.md and DESIGN.md section 5 quote for a kernel set:
#   sh tools/measure_round.sh <tag>       (on the GPU box, from the repo root; then tools/export_profiles.sh <tag> ... here)
tag=${1:-rX}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
for w in cfg2 cfg4 cfg5; do timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${tag}_$w.json 2>/dev/null; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_launch_bench.log 2>&1
for k in gmm_tc2 stats5 stats_pre beta_l2r_warp alpha_l2r; do
   timeout 600 ncu --set full --import-source on --clock-control none -k regex:${k}_kernel -s 3 -c 1 -o gpurun_out/${tag}_$k -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_$k.log 2>&1
done
ls -la gpurun_out/${tag}_*
