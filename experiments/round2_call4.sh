#!/usr/bin/env bash
# GPU call (one GPU): gpurun --timeout 2400 -- 'bash experiments/round2_call4.sh'
set -u
mkdir -p gpurun_out
T=gpurun_out/r02d
python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_config_sizes.py > ${T}_pytest.log 2>&1
echo "pytest: rc=$?" | tee -a ${T}_summary.log
tail -n 12 ${T}_pytest.log
python -m pytest tests/test_gpu_config_sizes.py -q -m gpu -s > ${T}_pytest_cfg.log 2>&1
echo "pytest config sizes: rc=$?" | tee -a ${T}_summary.log
grep -E "relative errors|passed|failed|Error" ${T}_pytest_cfg.log | cut -c1-300
for lazy in 1 0; do
  TB200_LAZY_UV=$lazy python bench.py --steps 10 --warmup 3 --no-aux --no-cpu-baseline > ${T}_bench_c5_lazy$lazy.log 2>&1
  echo "c5 lazy=$lazy rc=$?" | tee -a ${T}_summary.log
  tail -n 1 ${T}_bench_c5_lazy$lazy.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], {k:(round(v['ms_per_launch'],3), [round(x,3) for x in v.get('ms_by_stage')]) for k,v in d['roofline']['kernels'].items()}, {k: round(v['ms_per_launch'],3) for k,v in d['roofline'].get('other_kernels',{}).items()}, d['e2e']['value'])"
done
python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > ${T}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"stage_(a|b|mv2)_kernel|diag_column|velocity_xy" -s 33 -c 11 \
    -o ${T}_c5_full python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > ${T}_ncu_full.log 2>&1
ls -la gpurun_out | tail -n 8
