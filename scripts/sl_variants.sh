#!/bin/bash
# A/B of the symmetric-lattice SpMV build knobs on the GPU box (rebuilds psfem_solver.o per variant, runs bench.py).
# usage: scripts/sl_variants.sh "name:-DFLAG=.. -DFLAG2=.." ...     (results: gpurun_out/sl_variants.txt)
cd "$(dirname "$0")/.."
CS=plane-stress-fem-with-python_b200/csrc
mkdir -p gpurun_out
: > gpurun_out/sl_variants.txt
for v in "$@" "default:"; do
  name="${v%%:*}"; flags="${v#*:}"
  rm -f $CS/psfem_solver.o
  make -C $CS EXTRA="$flags" > /dev/null 2>&1 || { echo "$name build failed" >> gpurun_out/sl_variants.txt; continue; }
  python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', 'value %.4e' % d['value'], 'iter ms %.4f' % d['pcg_iteration']['ms'], 'spmv ms %.4f' % d['roofline']['ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], d['clocks']['sm_mhz'])
" >> gpurun_out/sl_variants.txt
done
cat gpurun_out/sl_variants.txt
