#!/bin/bash
cd "$(dirname "$0")/../.."
for prob in "7 256" "27 256"; do
  for mode in "0 0" "1 1" "1 2" "1 4"; do
    set -- $mode
    echo -n "repro=$1 chunk=$2 : " >> gpurun_out/chunk_ab.log
    FSB_REPRODUCIBLE=$1 FSB_SPMV_CHUNK=$2 python scripts/gpu/spmv_sweep.py $prob 2>&1 | tail -1 >> gpurun_out/chunk_ab.log
  done
done
