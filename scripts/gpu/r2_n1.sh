#!/bin/bash
# round 2, one GPU: whole -m gpu suite, then the bench line (both arms)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > $O/r2_n1_pytest.log
tail -6 $O/r2_n1_pytest.log
timeout 900 python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
tail -3 $O/r2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'launches',d['gpu_launches'],'fmt',d['config']['device_format'][:30])
print('spmv ms',round(r['avg_launch_ms'],4),'frac',round(r['frac'],3),'fmt GB/s',round(r['achieved_format_gbs']),'iter frac',round(r['iteration']['frac'],3))
print('other',d['other_solver']); print('sr',d['single_reduction_solver'])
print('e2e',d['e2e']['value'],d['e2e']['seconds'],d['e2e']['breakdown_ms'])
print('ref templates',d.get('reference_templates'))
print('parity',d.get('parity'))
print('cpu',d.get('cpu_baseline',{}).get('value'),d.get('cpu_baseline',{}).get('cores'))
for k,v in d['workloads'].items(): print(k, round(v['value'],2), round(v['ms_per_step'],3), 'spmv', round(v['roofline']['avg_launch_ms'],3), round(v['roofline']['frac'],3), 'other', round(v['other_solver']['value'],2), 'sr', round(v['single_reduction_solver']['value'],2))
PY
