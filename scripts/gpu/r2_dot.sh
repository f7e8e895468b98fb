#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_solvers_gpu.py -q -x 2>&1 | tail -4
{
for m in plain dotx doty dotu jacobi; do FSB_SPMV_DEBUG=1 python scripts/gpu/spmv_sweep.py 7 256 256 $m 2>&1 | grep -v "window format"; done
for m in plain dotx dotu jacobi; do FSB_SPMV_DEBUG=1 python scripts/gpu/spmv_sweep.py 27 256 256 $m 2>&1 | grep -v "window format"; done
for m in plain dotx; do FSB_SPMV_DEBUG=1 python scripts/gpu/spmv_sweep.py 27 512 512 $m 2>&1 | grep -v "window format"; done
for r in 384 416 480; do FSB_SPMV_ROWS=$r FSB_SPMV_DEBUG=1 python scripts/gpu/spmv_sweep.py 7 256 256 dotx 2>&1 | grep -v "window format"; done
} > $O/r2_dot.txt 2>&1
cat $O/r2_dot.txt
