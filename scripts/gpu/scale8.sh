#!/bin/bash
# strong-scaling sweep on one 8-GPU box
cd "$(dirname "$0")/../.."
summ() { grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', 'N', d['n_gpus'], 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'spmv', round(r['avg_launch_ms'],4), 'diag', round(r['diag_block_avg_ms'] or 0,4), 'offd', round(r['offd_block_avg_ms'] or 0,4), 'iterfrac', round(r['iteration']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'e2e_iters', d['e2e']['iterations'])"; }
python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -3
for wl in poisson7_256 poisson27_512; do
  steps=200; [ $wl = poisson27_512 ] && steps=50
  for n in 8 4 2; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps $steps --warmup 10 --repeats 1 --workload $wl 2>gpurun_out/scale_${wl}_$n.err | tee gpurun_out/scale_${wl}_$n.json | summ $wl
  done
done
python bench.py --steps 200 --warmup 10 --no-cpu-baseline --workload poisson7_256 2>/dev/null | tee gpurun_out/scale_poisson7_256_1.json | summ poisson7_256
python bench.py --steps 50 --warmup 10 --repeats 1 --no-cpu-baseline --workload poisson27_512 2>gpurun_out/scale_poisson27_512_1.err | tee gpurun_out/scale_poisson27_512_1.json | summ poisson27_512
