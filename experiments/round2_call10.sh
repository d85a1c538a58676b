#!/usr/bin/env bash
# GPU call (one GPU): validate the marching smoothing kernel, the cooperative diagnostics kernel and
# the moist plugin hook on the device; small-grid timings; smoothing capture on HBM-sized fields
set -u
mkdir -p gpurun_out
T=gpurun_out/r02j
python -m pytest tests/test_gpu_stencils.py tests/test_gpu_zz_late_additions.py tests/test_gpu_plugin_reference.py tests/test_gpu_moist_model.py tests/test_gpu_pipeline.py tests/test_gpu_isentropic.py tests/test_gpu_isentropic_physics.py tests/test_gpu_fullsize.py -q -m gpu > ${T}_pytest.log 2>&1
echo "pytest: rc=$?" | tee -a ${T}_summary.log
tail -n 6 ${T}_pytest.log
python -m pytest tests/test_gpu_stage_variants.py -q -m gpu -k small_grid > ${T}_pytest_variants.log 2>&1
echo "pytest variants: rc=$?" | tee -a ${T}_summary.log
tail -n 3 ${T}_pytest_variants.log
python experiments/small_grids.py --steps 40 > ${T}_small_grids.log 2>&1
grep -E "device_ms|launches_per" ${T}_small_grids.log
TB200_DIAG_IMPL=column TB200_SMOOTH_IMPL=tile python experiments/small_grids.py --steps 40 > ${T}_small_grids_old.log 2>&1
grep -E "device_ms" ${T}_small_grids_old.log
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum"
python experiments/all_kernels.py moist --size 640 640 64 --steps 1 > ${T}_all_moist.log 2>&1 &&
ncu $SECT --clock-control none -k regex:"smooth2_kernel|cross_kernel|box_kernel" -s 25 -c 30 -f -o /tmp/smooth python experiments/all_kernels.py moist --size 640 640 64 --steps 1 > ${T}_ncu_smooth.log 2>&1
ncu -i /tmp/smooth.ncu-rep --page raw --csv > ${T}_smooth_raw.csv 2>/dev/null; rm -f /tmp/smooth.ncu-rep
tail -n 1 ${T}_all_moist.log
du -sh gpurun_out
