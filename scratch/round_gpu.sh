#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
python scripts/run_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
