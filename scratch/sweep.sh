#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python scratch/first_gpu.py 2>&1 | grep -E "bitexact|match|fused dot" | head -12
for rows in 128 256 512; do for st in 2 3 4; do
  FSB_SPMV_ROWS=$rows FSB_SPMV_STAGES=$st timeout 120 python scratch/spmv_sweep.py 7 256
done; done
for c in 1 2 3; do FSB_SPMV_ROWS=256 FSB_SPMV_STAGES=3 FSB_SPMV_CTAS_PER_SM=$c timeout 120 python scratch/spmv_sweep.py 7 256; done
for rows in 64 128 256; do for st in 1 2 3; do
  FSB_SPMV_ROWS=$rows FSB_SPMV_STAGES=$st timeout 120 python scratch/spmv_sweep.py 27 192
done; done
FSB_SPMV_ROWS=256 FSB_SPMV_STAGES=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_stream -s 3 -c 1 -o gpurun_out/spmv7_r2 python scratch/spmv_sweep.py 7 256 > gpurun_out/ncu_spmv.log 2>&1
