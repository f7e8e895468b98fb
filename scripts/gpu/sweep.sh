#!/bin/bash
cd "$(dirname "$0")/../.."
export FSB_SPMV_DEBUG=1
timeout 300 python scripts/gpu/first_gpu.py 2>&1 | grep -E "bitexact|match|fused dot" | head -12
for rows in 128 256 512; do for st in 2 3 4; do
  FSB_SPMV_ROWS=$rows FSB_SPMV_STAGES=$st timeout 120 python scripts/gpu/spmv_sweep.py 7 256 2>&1
done; done
for rows in 64 128 256; do for st in 1 2 3; do
  FSB_SPMV_ROWS=$rows FSB_SPMV_STAGES=$st timeout 120 python scripts/gpu/spmv_sweep.py 27 192 2>&1
done; done
