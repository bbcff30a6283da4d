timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -c 300 gpurun_out/bench_r1_final.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r1_final.json').read().strip().splitlines()[-1]); print(d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'], d['kernels_ms_per_step'])"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
