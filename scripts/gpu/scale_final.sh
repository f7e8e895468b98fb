#!/bin/bash
# usage: scale_final.sh N [N...]: bench both workloads at the given rank counts on this box
cd "$(dirname "$0")/../.."
O=gpurun_out
run() { # N workload
  if [ "$1" = 1 ]; then
    python bench.py --gpus 1 --steps 200 --warmup 10 --workload $2 --no-cpu-baseline > $O/scale_$2_$1.json 2> $O/scale_$2_$1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 200 --warmup 10 --workload $2 2> $O/scale_$2_$1.err | grep '^{' > $O/scale_$2_$1.json
  fi
}
for n in "$@"; do
  run $n poisson7_256
  run $n poisson27_512
done
