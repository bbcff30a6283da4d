timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pre  ', d['value']/1e6, d['e2e']['value']/1e6, d['kernels_ms_per_step'])"
HFBGPU_NO_STATS_PRE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nopre', d['value']/1e6, d['e2e']['value']/1e6, d['kernels_ms_per_step'])"
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:stats_pre_kernel -s 3 -c 1 -o gpurun_out/r1m_stats_pre -f python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
