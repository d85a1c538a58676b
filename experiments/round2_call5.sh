#!/usr/bin/env bash
# GPU call (one GPU): moist fused stage tests, small-grid timings, c3 launch list
set -u
mkdir -p gpurun_out
T=gpurun_out/r02e
python -m pytest tests/test_gpu_isentropic.py tests/test_gpu_moist_model.py tests/test_gpu_velocity_components.py tests/test_gpu_graphs.py tests/test_gpu_plugin_reference.py -q -m gpu > ${T}_pytest.log 2>&1
echo "pytest: rc=$?" | tee -a ${T}_summary.log
tail -n 12 ${T}_pytest.log
python -m pytest tests/test_gpu_config_sizes.py -q -m gpu -s -k "c3 or moist" > ${T}_pytest_cfg.log 2>&1
echo "pytest config sizes (moist): rc=$?" | tee -a ${T}_summary.log
grep -E "relative errors|passed|failed|Error" ${T}_pytest_cfg.log | cut -c1-300
python experiments/small_grids.py --steps 40 > ${T}_small_grids.log 2>&1
echo "small grids rc=$?" | tee -a ${T}_summary.log
tail -n 30 ${T}_small_grids.log | cut -c1-600
TB200_MOIST_FUSED=0 python experiments/small_grids.py --steps 40 > ${T}_small_grids_unfused.log 2>&1
tail -n 12 ${T}_small_grids_unfused.log | cut -c1-600
for blk in 2x2 6x1; do
  TB200_MV_BLOCK=$blk python bench.py --steps 10 --warmup 3 --no-aux --no-cpu-baseline > ${T}_bench_c5_$blk.log 2>&1
  tail -n 1 ${T}_bench_c5_$blk.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$blk', d['ms_per_step'], {k:(round(v['ms_per_launch'],3), [round(x,3) for x in v.get('ms_by_stage')]) for k,v in d['roofline']['kernels'].items()})"
done
python experiments/small_grids.py --only c3 --steps 4 > ${T}_plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_launches_c3.csv \
    python experiments/small_grids.py --only c3 --steps 4 > ${T}_ncu_c3.log 2>&1
ls -la gpurun_out | tail -n 8
