#!/bin/bash
# round 2, N GPUs: sharded-path tests at N ranks (both transports), bench line at N
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
N=${1:-8}
if [ -z "$SKIP_PYTEST" ]; then timeout 900 python -m pytest tests/test_multi_gpu.py -q -x -k "$N-peer_memory or $N-nccl" > $O/r2_n${N}_pytest.log 2>&1; tail -5 $O/r2_n${N}_pytest.log; fi
P=$((29700 + N))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 200 --warmup 10 > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err
tail -3 $O/r2_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
r=d['roofline']
print('N=$N value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'launches',d['gpu_launches'],'fused',d['fused_halo'])
print('spmv ms',round(r['avg_launch_ms'],4),'frac',round(r['frac'],3),'iter frac',round(r['iteration']['frac'],3))
print('other',d['other_solver']['value']); print('sr',d['single_reduction_solver']['value']); print('general', (d.get('general_values') or {}).get('value'))
print('e2e',d['e2e']['value'],d['e2e']['seconds'])
for k,v in d['workloads'].items(): print(k, round(v['value'],2), round(v['ms_per_step'],3), 'spmv', round(v['roofline']['avg_launch_ms'],3), round(v['roofline']['frac'],3), 'other', round(v['other_solver']['value'],2), 'general', (v.get('general_values') or {}).get('value'))
PY
