timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nostore', d['kernels_ms_per_step'])"
