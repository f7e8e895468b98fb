#!/bin/bash
cd "$(dirname "$0")/../.."
python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_multi.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --workload $1 2>gpurun_out/bench_n2_$1.err | grep '^{' > gpurun_out/bench_n2_$1.json; }
run poisson7_256
