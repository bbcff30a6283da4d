mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_n.json 2> gpurun_out/bench_r1_n.err
for w in cfg2 cfg4 cfg5; do timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1_n_$w.json 2>/dev/null; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1n_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1n_launch_bench.log 2>&1
for k in beta_l2r_warp gmm_tc2; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:${k}_kernel -s 3 -c 1 -o gpurun_out/r1n_$k -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1n_ncu_$k.log 2>&1
done
ls -la gpurun_out/r1n_*
