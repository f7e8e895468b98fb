#!/bin/bash
# value dictionary: tests, SpMV alone with and without it, the bench line
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
run() { "$@" 2>&1 | grep -v "window format\|value dict" | tr '\n' ' ' | sed 's/\[fsb\] spmv window[+a-z]* config://'; echo; }
{
for d in 1 0; do for m in plain dotx jacobi; do FSB_SPMV_DICT=$d FSB_SPMV_DEBUG=1 run python scripts/gpu/spmv_sweep.py 7 256 256 $m; done; done
for d in 1 0; do for m in dotx; do FSB_SPMV_DICT=$d FSB_SPMV_DEBUG=1 run python scripts/gpu/spmv_sweep.py 27 512 512 $m; done; done
for d in 1 0; do for m in dotx; do FSB_SPMV_DICT=$d FSB_SPMV_DEBUG=1 run python scripts/gpu/spmv_sweep.py 27 256 256 $m; done; done
} > $O/r2_dict.txt 2>&1
cat $O/r2_dict.txt
if [ "$1" = "bench" ]; then
timeout 600 python bench.py --no-cpu-baseline > $O/r2_dict_bench.json 2> $O/r2_dict_bench.err; tail -2 $O/r2_dict_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_dict_bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'spmv ms',round(r['avg_launch_ms'],4),'e2e',d['e2e']['value'], 'other', d['other_solver']['value'], 'sr', d['single_reduction_solver']['value'])
print('general', d.get('general_values'))
for k,v in d['workloads'].items(): print(k, round(v['value'],2), round(v['ms_per_step'],3), 'spmv', round(v['roofline']['avg_launch_ms'],3), 'general', v.get('general_values'))
PY
fi
