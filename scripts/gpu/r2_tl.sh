#!/bin/bash
# device timelines of the CG flavours at N ranks: r2_tl.sh N [solvers...]
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
N=${1:-2}; shift
SOLVERS=${@:-cg cg_device cg_sr}
TAG=${TL_TAG:-}
P=29600
for solver in $SOLVERS; do
  if [ "$N" = "1" ]; then python scripts/gpu/timeline.py poisson7_256 $solver 40
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P scripts/gpu/timeline.py poisson7_256 $solver 40; fi
  P=$((P+1))
done 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$\|NCCL version" > $O/r2_timeline_n$N$TAG.txt
cat $O/r2_timeline_n$N$TAG.txt
