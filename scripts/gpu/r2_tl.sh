#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
N=${1:-2}
P=29600
for solver in cg cg_device; do
  if [ "$N" = "1" ]; then python scripts/gpu/timeline.py poisson7_256 $solver 40
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P scripts/gpu/timeline.py poisson7_256 $solver 40; fi
  P=$((P+1))
done 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" > $O/r2_timeline_n$N.txt
cat $O/r2_timeline_n$N.txt
