#!/bin/bash
cd "$(dirname "$0")/../.."
bash scripts/gpu/scale_final.sh 8
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 200 --warmup 10 --workload poisson27_512 2> $O/scale_poisson27_512_4.err | grep '^{' > $O/scale_poisson27_512_4.json
