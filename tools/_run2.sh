timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1z_launch_bench.log 2>&1
for k in gmm_tc2 stats5 stats_pre beta_l2r_warp alpha_l2r; do
timeout 200 ncu --set full --import-source on --clock-control none -k regex:${k}_kernel -s 3 -c 1 -o gpurun_out/r1z_$k -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1z_ncu_$k.log 2>&1
done
ls gpurun_out/r1z_* | wc -l
