#!/bin/bash
# round-2 single-GPU evidence: whole -m gpu suite, smoke, bench (ours + reference arm), launch list of the bench command,
# full ncu captures of the SpMV (7-pt 256^3 inside the bench; 27-pt 512^3 alone) and of CG's element-wise kernels
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > $O/r2_final_pytest.log; tail -3 $O/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_final_smoke.log 2>&1; tail -2 $O/r2_final_smoke.log
timeout 900 python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -2 $O/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err; tail -2 $O/r2_bench_ref.err
if [ "$1" != "quick" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-workloads --repeats 1 > $O/r2_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:spmv_window -s 30 -c 1 -f -o $O/r2_final_spmv7 \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-workloads --repeats 1 > $O/r2_ncu_spmv7.log 2>&1
FSB_SPMV_DICT=0 timeout 600 ncu --set full --clock-control none -k regex:spmv_window -s 30 -c 1 -f -o $O/r2_final_spmv7_general \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-workloads --repeats 1 > $O/r2_ncu_spmv7_general.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:ew_program -s 60 -c 3 -f -o $O/r2_final_ew \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-workloads --repeats 1 > $O/r2_ncu_ew.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:spmv_window -s 5 -c 1 -f -o $O/r2_final_spmv27_512 \
    python scripts/gpu/spmv_sweep.py 27 512 512 dotx > $O/r2_ncu_spmv27.log 2>&1
FSB_SPMV_DICT=0 timeout 900 ncu --set full --clock-control none -k regex:spmv_window -s 5 -c 1 -f -o $O/r2_final_spmv27_512_general \
    python scripts/gpu/spmv_sweep.py 27 512 512 dotx > $O/r2_ncu_spmv27_general.log 2>&1
fi
# keep what comes back small (gpurun merges at most 64 MiB): raw metric tables instead of the reports
for r in $O/r2_final_*.ncu-rep; do
  [ -f "$r" ] || continue
  ncu -i "$r" --page raw --csv > "${r%.ncu-rep}.raw.csv" 2>/dev/null
  case "$r" in *r2_final_spmv7.ncu-rep) ;; *) rm -f "$r";; esac
done
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'launches',d['gpu_launches'])
print('spmv ms',round(r['avg_launch_ms'],4),'frac',round(r['frac'],3),'fmt GB/s',round(r['achieved_format_gbs']),'iter frac',round(r['iteration']['frac'],3))
print('e2e',d['e2e']['value'],'parity',d.get('parity'),'cpu',d.get('cpu_baseline',{}).get('value'),d.get('cpu_baseline',{}).get('cores'))
print('clocks',d.get('clocks'))
for k,v in d['workloads'].items(): print(k, round(v['value'],2), round(v['ms_per_step'],3), 'spmv', round(v['roofline']['avg_launch_ms'],3), round(v['roofline']['frac'],3))
print(open('gpurun_out/r2_bench_ref.json').read()[:600])
PY
ls -la $O/*.ncu-rep 2>/dev/null | tail -5
