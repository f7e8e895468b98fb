#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
