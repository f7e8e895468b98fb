#!/bin/bash
cd "$(dirname "$0")/../.."
bash scripts/gpu/scale_final.sh 8 4
python -m pytest tests/test_multi_gpu.py -q -m gpu -k "8" 2>&1 | tail -6 > gpurun_out/pytest_multi8.log
