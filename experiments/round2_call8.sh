#!/usr/bin/env bash
# GPU call (one GPU): tile height of the tracer kernel, ncu evidence of the moist physics kernels
set -u
mkdir -p gpurun_out
T=gpurun_out/r02h
python -m pytest tests/test_gpu_isentropic.py tests/test_gpu_moist_model.py -q -m gpu > ${T}_pytest.log 2>&1
echo "pytest: rc=$?" | tee -a ${T}_summary.log
tail -n 3 ${T}_pytest.log
for rows in 4 8 16; do
  TB200_T_ROWS=$rows python experiments/small_grids.py --only c3 --steps 40 > ${T}_c3_rows$rows.log 2>&1
  echo "rows=$rows $(tail -n 1 ${T}_c3_rows$rows.log | cut -c1-200)"
  TB200_T_ROWS=$rows python experiments/all_kernels.py moist --size 640 640 64 --steps 2 > ${T}_m640_rows$rows.log 2>&1
  echo "rows=$rows $(tail -n 1 ${T}_m640_rows$rows.log | cut -c1-200)"
done
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum"
python experiments/all_kernels.py moist --size 640 640 64 --steps 1 > ${T}_all_moist.log 2>&1 &&
ncu $SECT --clock-control none -k regex:"box_kernel|vadv|smagorinsky|cross_kernel|fma_fields|tracers|column_kernel|diag_column|stage_|velocity_xy" -s 63 -c 66 -f -o /tmp/moist python experiments/all_kernels.py moist --size 640 640 64 --steps 1 > ${T}_ncu_moist.log 2>&1
ncu -i /tmp/moist.ncu-rep --page raw --csv > ${T}_moist_raw.csv 2>/dev/null; rm -f /tmp/moist.ncu-rep
python experiments/all_kernels.py burgers --steps 1 > ${T}_all_burgers.log 2>&1 &&
ncu $SECT --clock-control none -k regex:"box_kernel" -s 6 -c 3 -f -o /tmp/burgers python experiments/all_kernels.py burgers --steps 1 > ${T}_ncu_burgers.log 2>&1
ncu -i /tmp/burgers.ncu-rep --page raw --csv > ${T}_burgers_raw.csv 2>/dev/null; rm -f /tmp/burgers.ncu-rep
du -sh gpurun_out
