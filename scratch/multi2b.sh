#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_multi_gpu.py tests/test_spmv_gpu.py -q -m gpu 2>&1 | tail -4
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --workload poisson7_256 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', 'it/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'spmv ms', round(r['avg_launch_ms'],4), 'diag', round(r['diag_block_avg_ms'],4), 'offd', round(r['offd_block_avg_ms'],4), 'e2e', round(d['e2e']['value']))"; }
run p2p_halo
FSB_P2P_HALO=0 run nccl_halo
for g in 8 32; do for cfg in "64 1" "128 1" "256 1"; do set -- $cfg; FSB_SPMV_GATHER=$g FSB_SPMV_ROWS=$1 FSB_SPMV_STAGES=$2 FSB_SPMV_DEBUG=1 python scratch/spmv_sweep.py 27 512 64 2>&1 | tail -2; done; done
python scratch/spmv_sweep.py 27 192
python scratch/spmv_sweep.py 7 256
