#!/usr/bin/env bash
# GPU call (one GPU): stepped Coriolis / Smagorinsky epilogues and the marching vertical advection
set -u
mkdir -p gpurun_out
T=gpurun_out/r02l
python -m pytest tests/test_gpu_moist_model.py tests/test_gpu_isentropic_physics.py tests/test_gpu_graphs.py -q -m gpu > ${T}_pytest.log 2>&1
echo "pytest (default): rc=$?" | tee -a ${T}_summary.log
tail -n 3 ${T}_pytest.log
python -m pytest tests/test_gpu_config_sizes.py -q -m gpu -k "c3 or moist" > ${T}_pytest_cfg.log 2>&1
echo "pytest config sizes (moist): rc=$?" | tee -a ${T}_summary.log
tail -n 2 ${T}_pytest_cfg.log
TB200_VADV_IMPL=march python -m pytest tests/test_gpu_moist_model.py tests/test_gpu_isentropic_physics.py -q -m gpu > ${T}_pytest_march.log 2>&1
echo "pytest (vadv march): rc=$?" | tee -a ${T}_summary.log
tail -n 3 ${T}_pytest_march.log
TB200_VADV_IMPL=march python -m pytest tests/test_gpu_config_sizes.py -q -m gpu -k "c3 or moist" > ${T}_pytest_cfg_march.log 2>&1
echo "pytest config sizes (moist, vadv march): rc=$?" | tee -a ${T}_summary.log
for impl in point march; do
  TB200_VADV_IMPL=$impl python experiments/small_grids.py --only c3 --steps 40 > ${T}_c3_$impl.log 2>&1
  echo "vadv=$impl $(tail -n 1 ${T}_c3_$impl.log | cut -c1-200)"
  TB200_VADV_IMPL=$impl python experiments/all_kernels.py moist --size 640 640 64 --steps 2 > ${T}_m640_$impl.log 2>&1
  echo "vadv=$impl $(tail -n 1 ${T}_m640_$impl.log | cut -c1-200)"
done
TB200_VADV_IMPL=march python experiments/small_grids.py --only c3 --steps 4 > ${T}_plain_c3.log 2>&1 &&
TB200_VADV_IMPL=march ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"vadv|coriolis|smagorinsky|frame_fma0|fma_fields" --csv --log-file ${T}_launches_c3_march.csv \
    python experiments/small_grids.py --only c3 --steps 4 > ${T}_ncu_c3.log 2>&1
du -sh gpurun_out
