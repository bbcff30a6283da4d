for d in 0 128 224; do
HFBGPU_TC_DEBUG=$d timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dbg $d', d['kernels_ms_per_step']['gmm']-d['kernels_ms_per_step']['gmm_expand'])"
done
