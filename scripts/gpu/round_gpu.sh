#!/bin/bash
cd "$(dirname "$0")/../.."
python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
