#!/usr/bin/env bash
# GPU call (one GPU): tiled tracer kernel, overlap fix (in-process path), HBM-sized ncu evidence of
# every kernel family
set -u
mkdir -p gpurun_out
T=gpurun_out/r02f
python -m pytest tests/test_gpu_isentropic.py tests/test_gpu_moist_model.py tests/test_gpu_distributed.py tests/test_gpu_pipeline.py tests/test_gpu_stage_variants.py -q -m gpu > ${T}_pytest.log 2>&1
echo "pytest: rc=$?" | tee -a ${T}_summary.log
tail -n 6 ${T}_pytest.log
python -m pytest tests/test_gpu_config_sizes.py -q -m gpu -s -k "c3 or moist" > ${T}_pytest_cfg.log 2>&1
echo "pytest config sizes (moist): rc=$?" | tee -a ${T}_summary.log
grep -E "passed|failed|Error" ${T}_pytest_cfg.log | cut -c1-300
python experiments/small_grids.py --steps 40 > ${T}_small_grids.log 2>&1
echo "small grids rc=$?" | tee -a ${T}_summary.log
grep -E "device_ms|launches_per" ${T}_small_grids.log
python bench.py --steps 20 --warmup 3 --no-aux --no-cpu-baseline > ${T}_bench_c5.log 2>&1
tail -n 1 ${T}_bench_c5.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c5', d['ms_per_step'], d['hbm_frac_step'], {k:(round(v['ms_per_launch'],3), [round(x,3) for x in v.get('ms_by_stage')]) for k,v in d['roofline']['kernels'].items()}, {k: round(v['ms_per_launch'],3) for k,v in d['roofline'].get('other_kernels',{}).items()}, d['e2e']['value'])"
for what in moist burgers halo; do
  python experiments/all_kernels.py $what > ${T}_all_$what.log 2>&1
  echo "all_kernels $what rc=$?" | tee -a ${T}_summary.log
  tail -n 1 ${T}_all_$what.log | cut -c1-300
done
# one full ncu capture per family (after the plain runs above exited 0)
ncu --set full --clock-control none -s 70 -c 75 -o ${T}_moist_full python experiments/all_kernels.py moist --steps 1 > ${T}_ncu_moist.log 2>&1
ncu --set full --clock-control none -k regex:"burgers" -s 4 -c 3 -o ${T}_burgers_full python experiments/all_kernels.py burgers --steps 1 > ${T}_ncu_burgers.log 2>&1
ncu --set full --clock-control none -k regex:"p2p_kernel" -s 6 -c 8 -o ${T}_halo_full python experiments/all_kernels.py halo --steps 1 > ${T}_ncu_halo.log 2>&1
ls -la gpurun_out | tail -n 8
