#!/bin/bash
# round 2, first call: whole GPU suite without -x, the run-time kernel tests, a baseline bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_first_gpus.txt
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -80 > gpurun_out/r2_pytest_gpu.log
timeout 300 python -m pytest tests/test_zz_jit_hw.py -q 2>&1 | tail -40 > gpurun_out/r2_pytest_jit.log
timeout 600 python bench.py > gpurun_out/r2_bench_n1_start.json 2> gpurun_out/r2_bench_n1_start.err
tail -5 gpurun_out/r2_pytest_gpu.log; tail -5 gpurun_out/r2_pytest_jit.log; cat gpurun_out/r2_bench_n1_start.json
