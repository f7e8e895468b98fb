#!/bin/bash
# one GPU call: tests, smoke, bench, ncu launch list + full capture of the SpMV kernel
cd "$(dirname "$0")/.."
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_stream -s 20 -c 2 -o gpurun_out/spmv7_r1_final python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ew_program -s 60 -c 6 -o gpurun_out/ew_r1_final python bench.py --steps 20 --warmup 5 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
