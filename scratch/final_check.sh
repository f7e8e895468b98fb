#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py --steps 200 --warmup 10 > $O/bench_n1.json 2> $O/bench_n1.err
for c in 1 4; do echo -n "static chunk=$c : " >> $O/chunk512.log; FSB_SPMV_CHUNK=$c python scratch/spmv_sweep.py 27 512 2>&1 | tail -1 >> $O/chunk512.log; done
python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > $O/pytest_gpu.log
