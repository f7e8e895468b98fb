#!/bin/bash
# configs 1/3/4 at N ranks (full size) + the sharded-path worker
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
N=${1:-2}
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x -k "$N-peer_memory or $N-nccl or single_rank" 2>&1 | tail -15
if [ "$N" = "1" ]; then timeout 900 python scripts/run_configs.py > $O/r2_configs_n$N.txt 2> $O/r2_configs_n$N.err
else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 scripts/run_configs.py > $O/r2_configs_n$N.txt 2> $O/r2_configs_n$N.err; fi
tail -3 $O/r2_configs_n$N.err | cut -c1-300
python - <<PY
import json
for l in open('gpurun_out/r2_configs_n$N.txt'):
    if l.startswith('{'):
        d=json.loads(l); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k not in ('roofline_note',)})
PY
