#!/bin/bash
# round 2, two GPUs: sharded-path tests (fused halo / two-launch / NCCL), then the 256^3 bench at N=2
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x -k "2-" 2>&1 | tail -40 > $O/r2_n2_pytest.log
tail -8 $O/r2_n2_pytest.log
P=29517
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 200 --warmup 10 > $O/r2_n2_bench.json 2> $O/r2_n2_bench.err
cat $O/r2_n2_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=2 value', d['value'], 'ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'other', d['other_solver'], 'spmv', d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
FSB_FUSED_HALO=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --steps 200 --warmup 10 > $O/r2_n2_bench_twolaunch.json 2>> $O/r2_n2_bench.err
cat $O/r2_n2_bench_twolaunch.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=2 two-launch value', d['value'], 'ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'other', d['other_solver'])"
tail -3 $O/r2_n2_bench.err
