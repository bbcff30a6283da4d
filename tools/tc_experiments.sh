#!/bin/sh
# Timing experiments for gmm_tc_kernel (results are WRONG in the debug modes; only the time matters)
# bits: 1 no A_lo x B_hi, 2 no epilogue (no TMEM loads / math / stores), 4 K = 96, 8 no MMAs, 16 no B loads
export HFBGPU_NO_PAIR=1
for d in 0 8 16 24 26; do
  HFBGPU_TC_DEBUG=$d timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dbg=$d', 'gmm ms %.3f' % d['kernels_ms_per_step']['gmm'], 'step ms %.3f' % d['ms_per_step'])"
done
