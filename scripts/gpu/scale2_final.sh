#!/bin/bash
cd "$(dirname "$0")/../.."
bash scripts/gpu/scale_final.sh 2
python bench.py --gpus 1 --steps 200 --warmup 10 --workload poisson27_512 --no-cpu-baseline > gpurun_out/scale_poisson27_512_1.json 2> gpurun_out/scale_poisson27_512_1.err
