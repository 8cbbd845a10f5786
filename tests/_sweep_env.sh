#!/bin/bash
# usage: tests/_sweep_env.sh "A=1 B=2" "A=3" ... : bench.py per setting of experiment knobs, prints the per-kernel times
for v in "$@"; do
  env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernels']
print('[$v]', 'step %.3f' % d['ms_per_step'], ' '.join('%s %.3f' % (n, k[n]['ms_per_launch']) for n in ('sdf_bwd_data', 'sdf_fwd_grad', 'dw_gemm', 'sdf_fwd', 'albedo_fwd', 'albedo_bwd', 'colsum') if n in k), 'clk', d['clocks']['sm_mhz'])
"
done
