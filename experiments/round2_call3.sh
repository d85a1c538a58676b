#!/usr/bin/env bash
# Third GPU call of round 2 (one GPU): gpurun --timeout 2400 -- 'bash experiments/round2_call3.sh'
# 1. the whole GPU parity suite (new: division self-test, p2p transport in process, pipeline fix)
# 2. headline bench, lazy and reference velocity flow, after the division / row-ahead-load fixes
# 3. launch list + full ncu capture of the stage kernels
set -u
mkdir -p gpurun_out
T=gpurun_out/r02c
python -m pytest tests -q -m gpu > ${T}_pytest.log 2>&1
echo "pytest: rc=$?" | tee -a ${T}_summary.log
tail -n 15 ${T}_pytest.log
for lazy in 1 0; do
  TB200_LAZY_UV=$lazy python bench.py --steps 10 --warmup 3 --no-aux --no-cpu-baseline > ${T}_bench_c5_lazy$lazy.log 2>&1
  echo "c5 lazy=$lazy rc=$?" | tee -a ${T}_summary.log
  tail -n 1 ${T}_bench_c5_lazy$lazy.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], {k:(round(v['ms_per_launch'],3), [round(x,3) for x in v.get('ms_by_stage')]) for k,v in d['roofline']['kernels'].items()}, d['e2e']['value'])"
done
python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > ${T}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stage_(a|b|mv2)_kernel|diag_column" -s 30 -c 10 \
    -o ${T}_c5_full python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > ${T}_ncu_full.log 2>&1
ls -la gpurun_out | tail -n 12
