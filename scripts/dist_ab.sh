#!/bin/bash
# A/B of the strip solver's peer protocol on N GPUs of the box: rebuilds psfem_solver.o per variant.
# usage: scripts/dist_ab.sh NGPU "name:-DFLAGS" ...
cd "$(dirname "$0")/.."
N=$1; shift
CS=plane-stress-fem-with-python_b200/csrc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29600
for v in "$@"; do
  name="${v%%:*}"; flags="${v#*:}"
  rm -f $CS/psfem_solver.o
  make -C $CS EXTRA="$flags" > /dev/null 2>&1 || { echo "$name build failed"; continue; }
  for rep in 1 2; do
    port=$((port+1))
    $TR --master-port $port bench.py --gpus $N --steps 3 --warmup 3 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name weak', '%.4e' % d['value'], 'iter ms %.4f' % d['pcg_iteration']['ms'], 'spmv ms %.4f' % d['roofline']['ms_per_launch'])"
  done
  port=$((port+1))
  $TR --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --cells 1024 --scaling strong --cg-iters 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name strong1024', '%.4e' % d['value'], 'iter ms %.4f' % d['pcg_iteration']['ms'], 'spmv ms %.4f' % d['roofline']['ms_per_launch'])"
done
rm -f $CS/psfem_solver.o; make -C $CS > /dev/null 2>&1
