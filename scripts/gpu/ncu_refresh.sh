#!/bin/bash
# refresh of the ncu evidence for the final SpMV schedule (launch list + one full capture)
cd "$(dirname "$0")/../.."
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 260 --csv --log-file $O/launches_r1.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --repeats 1 > $O/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_stream -s 30 -c 2 -f -o $O/spmv7_r1_final \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --repeats 1 > $O/ncu_spmv.log 2>&1
