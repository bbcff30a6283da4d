timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r1_final.json').read().strip().splitlines()[-1]); print(d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline']['frac_of_peak_over_3'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'], d['kernels_ms_per_step'], d['rooflines']['recursions']['frac'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1q_launch_bench.log 2>&1
for k in alpha_l2r stats5; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:${k}_kernel -s 3 -c 1 -o gpurun_out/r1q_$k -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1q_ncu_$k.log 2>&1
done
