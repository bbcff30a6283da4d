set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_m.json 2> gpurun_out/bench_r1_m.err
tail -c 600 gpurun_out/bench_r1_m.err
for w in cfg2 cfg4 cfg5; do timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1_m_$w.json 2>/dev/null; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1m_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1m_launch_bench.log 2>&1
for k in gmm_tc2 stats5 stats_pre beta_l2r alpha_l2r; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:${k}_kernel -s 3 -c 1 -o gpurun_out/r1m_$k -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1m_ncu_$k.log 2>&1
done
ls -la gpurun_out/r1m_*
