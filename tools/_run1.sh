timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dev', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, d['kernels_ms_per_step'])"
done
