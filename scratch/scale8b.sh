#!/bin/bash
cd "$(dirname "$0")/.."
summ() { grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', 'N', d['n_gpus'], 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'spmv', round(r['avg_launch_ms'],4), 'diag', round(r['diag_block_avg_ms'] or 0,4), 'offd', round(r['offd_block_avg_ms'] or 0,4), 'iterfrac', round(r['iteration']['frac'],3), 'e2e', round(d['e2e']['value'],1))"; }
python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -3
for n in 8 4; do
  for mode in p2p nccl; do
    if [ $mode = nccl ]; then export FSB_P2P_HALO=0; else unset FSB_P2P_HALO; fi
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 10 --repeats 3 --workload poisson7_256 2>/dev/null | summ poisson7_256_$mode
  done
done
unset FSB_P2P_HALO
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 50 --warmup 10 --repeats 1 --workload poisson27_512 2>/dev/null | summ poisson27_512_p2p
