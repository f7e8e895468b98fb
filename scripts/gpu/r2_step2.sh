#!/bin/bash
# round 2: new SpMV kernels (window format, one-pass Jacobi) on one GPU: tests, then a sweep of block size / stages
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -60 > $O/r2_step2_pytest.log
tail -12 $O/r2_step2_pytest.log
{
for cfg in "32 2" "32 3" "32 4" "64 2" "64 3" "64 4" "128 1" "96 2"; do set -- $cfg
FSB_SPMV_ROWS=$1 FSB_SPMV_STAGES=$2 FSB_SPMV_DEBUG=1 python scripts/gpu/spmv_sweep.py 27 256 2>&1 | grep -v "window format"
done
for cfg in "128 2" "128 3" "128 4" "256 2" "256 3" "256 4" "512 1" "512 2" "64 4"; do set -- $cfg
FSB_SPMV_WINDOW=2 FSB_SPMV_ROWS=$1 FSB_SPMV_STAGES=$2 FSB_SPMV_DEBUG=1 python scripts/gpu/spmv_sweep.py 7 256 2>&1 | grep -v "window format"
done
} > $O/r2_step2_spmv.txt 2>&1
cat $O/r2_step2_spmv.txt
