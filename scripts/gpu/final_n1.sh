#!/bin/bash
# round-1 final single-GPU evidence: tests, bench (ours + reference arm), launch list, full ncu captures, configs
cd "$(dirname "$0")/../.."
O=gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -15 > $O/pytest_gpu.log
python bench.py --steps 200 --warmup 10 > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python scripts/run_configs.py > $O/configs.jsonl 2> $O/configs.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 260 --csv --log-file $O/launches_r1.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --repeats 1 > $O/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_stream -s 30 -c 2 -f -o $O/spmv7_r1_final \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --repeats 1 > $O/ncu_spmv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ew_program -s 60 -c 4 -f -o $O/ew_r1_final \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --repeats 1 > $O/ncu_ew.log 2>&1
