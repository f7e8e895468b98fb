#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --repeats 2 --workload $1 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$2 $1', 'it/s', round(d['value'],1), 'dev', round(d['other_solver']['value'],1), 'diag us', round(r['diag_block_avg_ms']*1000,1), 'offd', round(r['offd_block_avg_ms']*1000,1))" >> $O/sched_ab.log; }
for w in poisson27_512 poisson7_256; do
  FSB_REPRODUCIBLE=0 run $w dynamic
  FSB_REPRODUCIBLE=1 run $w static
done
