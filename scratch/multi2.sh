#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_multi_gpu.py -q -m gpu -x 2>&1 | tail -30 > gpurun_out/pytest_multi.log
for wl in poisson7_256; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --workload $wl > gpurun_out/bench_n2_$wl.json 2> gpurun_out/bench_n2_$wl.err
done
