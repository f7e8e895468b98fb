#!/bin/bash
# round 2: full ncu captures of the SpMV kernel on the 27-point operators (512^3: int64 offsets; 256^3: int32) + plain timings
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
{
python scripts/gpu/spmv_sweep.py 27 512
python scripts/gpu/spmv_sweep.py 27 256
python scripts/gpu/spmv_sweep.py 7 256
python scripts/gpu/spmv_sweep.py 7 512
} > $O/r2_spmv_plain.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_stream -s 5 -c 1 -f -o $O/r2_spmv27_512 \
    python scripts/gpu/spmv_sweep.py 27 512 > $O/r2_ncu27_512.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_stream -s 5 -c 1 -f -o $O/r2_spmv27_256 \
    python scripts/gpu/spmv_sweep.py 27 256 > $O/r2_ncu27_256.log 2>&1
cat $O/r2_spmv_plain.txt; ls -la $O/*.ncu-rep
