#!/usr/bin/env bash
# GPU call (one GPU): the whole parity suite after the division replacements in the physics kernels,
# small-grid timings, launch list of configs[2]
set -u
mkdir -p gpurun_out
T=gpurun_out/r02k
python -m pytest tests -q -m gpu > ${T}_pytest.log 2>&1
echo "pytest: rc=$?" | tee -a ${T}_summary.log
tail -n 6 ${T}_pytest.log
python experiments/small_grids.py --steps 40 > ${T}_small_grids.log 2>&1
grep -E "device_ms|launches_per" ${T}_small_grids.log
python experiments/all_kernels.py moist --size 640 640 64 --steps 2 > ${T}_all_moist.log 2>&1
tail -n 1 ${T}_all_moist.log
python experiments/small_grids.py --only c3 --steps 4 > ${T}_plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${T}_launches_c3.csv \
    python experiments/small_grids.py --only c3 --steps 4 > ${T}_ncu_c3.log 2>&1
du -sh gpurun_out
