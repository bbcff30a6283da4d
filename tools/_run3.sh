HFBGPU_TRACE_KERNELS=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu 2> gpurun_out/trace_async.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['ms_per_step'])"
grep "hfbgpu trace" gpurun_out/trace_async.log | tail -30
