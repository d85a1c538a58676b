#!/usr/bin/env bash
# Last GPU call of the round (one GPU): smoke(), the whole GPU parity suite, the default bench line
# (headline + aux + cpu baseline), the reference arm, and the ncu capture bench.py reads its
# `roofline.traffic` from (exported to CSV on the box).
set -u
mkdir -p gpurun_out
T=gpurun_out/r02z
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > ${T}_smoke.log 2>&1
echo "smoke: rc=$?" | tee -a ${T}_summary.log
python -m pytest tests -q -m gpu > ${T}_pytest.log 2>&1
echo "pytest: rc=$?" | tee -a ${T}_summary.log
tail -n 6 ${T}_pytest.log
python bench.py > ${T}_bench_default.log 2>&1
echo "bench default: rc=$?" | tee -a ${T}_summary.log
tail -n 1 ${T}_bench_default.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c5', d['ms_per_step'], d['value'], d['hbm_frac_step'], {k:(round(v['ms_per_launch'],3), [round(x,3) for x in v.get('ms_by_stage')]) for k,v in d['roofline']['kernels'].items()}, {k: round(v['ms_per_launch'],3) for k,v in d['roofline'].get('other_kernels',{}).items()})
print('e2e', d['e2e']['value'], d['e2e'].get('pinned_buffers_numa_node'), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('traffic', d['roofline'].get('traffic'), d['roofline'].get('traffic_source'))
print('cpu_baseline', d.get('cpu_baseline'))
for k,v in d.get('aux',{}).items(): print(k, json.dumps(v)[:400])"
python bench.py --impl reference --steps 5 --warmup 0 > ${T}_bench_reference.log 2>&1
echo "bench reference: rc=$?" | tee -a ${T}_summary.log
tail -n 1 ${T}_bench_reference.log | cut -c1-700
python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > ${T}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stage_(a|b|mv2)_kernel|diag_column|velocity_xy" -s 33 -c 11 \
    -f -o /tmp/c5_full python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > ${T}_ncu_full.log 2>&1
ncu -i /tmp/c5_full.ncu-rep --page raw --csv > ${T}_c5_ncu_full_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file ${T}_launches_c5.csv \
    python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > ${T}_ncu_launches.log 2>&1
rm -f /tmp/c5_full.ncu-rep
du -sh gpurun_out
