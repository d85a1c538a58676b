#!/usr/bin/env bash
# Second GPU call of round 2 (one GPU):
#   gpurun --timeout 2400 -- 'bash experiments/round2_call2.sh'
# 1. the GPU parity suite (new: config-size tests, lazy velocities, two-column diffusion, pipeline)
# 2. headline bench with the three block shapes of the momentum kernel + the old data flow
# 3. C4 bench (march2 kernel)
# 4. launch list + full ncu capture of the stage kernels and of the diffusion kernel
set -u
mkdir -p gpurun_out
T=gpurun_out/r02b
python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_config_sizes.py > ${T}_pytest.log 2>&1
echo "pytest (without config sizes): rc=$?" | tee -a ${T}_summary.log
tail -n 3 ${T}_pytest.log
python -m pytest tests/test_gpu_config_sizes.py -x -q -m gpu -s > ${T}_pytest_cfg.log 2>&1
echo "pytest config sizes: rc=$?" | tee -a ${T}_summary.log
grep -E "relative errors|passed|failed|Error" ${T}_pytest_cfg.log | cut -c1-400
for blk in 3x1 6x1 2x2; do
  TB200_MV_BLOCK=$blk python bench.py --steps 10 --warmup 3 --no-aux --no-cpu-baseline > ${T}_bench_c5_$blk.log 2>&1
  echo "c5 $blk rc=$?" | tee -a ${T}_summary.log
  tail -n 1 ${T}_bench_c5_$blk.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], {k:(round(v['ms_per_launch'],3), v.get('ms_by_stage')) for k,v in d['roofline']['kernels'].items()}, d['e2e']['value'])"
done
python bench.py --steps 20 --warmup 3 > ${T}_bench_c5_full.log 2>&1
echo "c5 full (aux) rc=$?" | tee -a ${T}_summary.log
tail -n 1 ${T}_bench_c5_full.log | cut -c1-3000
python bench.py --workload c4 --steps 10 --warmup 3 > ${T}_bench_c4.log 2>&1
echo "c4 rc=$?" | tee -a ${T}_summary.log
tail -n 1 ${T}_bench_c4.log | cut -c1-1200
# ---- profiles (only after the plain runs)
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file ${T}_launches_c5.csv python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > ${T}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"stage_(a|b|mv2)_kernel" -s 27 -c 9 \
    -o ${T}_c5_full python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > ${T}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"march2_kernel|fma_fields_kernel" -s 6 -c 4 \
    -o ${T}_c4_full python bench.py --workload c4 --steps 3 --warmup 3 > ${T}_ncu_full_c4.log 2>&1
ls -la gpurun_out | tail -n 15
