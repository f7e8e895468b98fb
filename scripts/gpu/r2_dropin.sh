#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin_gpu.py -q 2>&1 | tail -40 > gpurun_out/r2_dropin.log
tail -15 gpurun_out/r2_dropin.log
