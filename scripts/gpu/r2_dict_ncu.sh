#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
run() { "$@" 2>&1 | grep -v "window format\|value dict" | tr '\n' ' ' | sed 's/\[fsb\] spmv window+dictionary config://'; echo; }
{
for t in 192 256 320 384; do for st in 2 3; do FSB_SPMV_THREADS=$t FSB_SPMV_STAGES=$st FSB_SPMV_ROWS=512 FSB_SPMV_DEBUG=1 run python scripts/gpu/spmv_sweep.py 7 256 256 dotx; done; done
for t in 192 128; do FSB_SPMV_THREADS=$t FSB_SPMV_ROWS=384 FSB_SPMV_DEBUG=1 run python scripts/gpu/spmv_sweep.py 7 256 256 dotx; done
FSB_SPMV_THREADS=256 FSB_SPMV_ROWS=512 FSB_SPMV_DEBUG=1 run python scripts/gpu/spmv_sweep.py 7 256 256 plain
FSB_SPMV_THREADS=256 FSB_SPMV_ROWS=512 FSB_SPMV_DEBUG=1 run python scripts/gpu/spmv_sweep.py 7 256 256 jacobi
for r in 64 128 192 256; do for t in 32 64 128; do FSB_SPMV_THREADS=$t FSB_SPMV_ROWS=$r FSB_SPMV_DEBUG=1 run python scripts/gpu/spmv_sweep.py 27 256 256 dotx; done; done
} > $O/r2_dict_sweep2.txt 2>&1
cat $O/r2_dict_sweep2.txt
