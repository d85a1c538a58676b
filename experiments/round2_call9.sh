#!/usr/bin/env bash
# GPU call (one GPU): full parity suite on the current build, a few one-off timings, Burgers capture
set -u
mkdir -p gpurun_out
T=gpurun_out/r02i
python -m pytest tests -q -m gpu > ${T}_pytest.log 2>&1
echo "pytest: rc=$?" | tee -a ${T}_summary.log
tail -n 8 ${T}_pytest.log
for env in "TB200_B_IMPL=column" "TB200_B_IMPL=coop"; do
  env $env python bench.py --steps 10 --warmup 3 --no-aux --no-cpu-baseline > ${T}_bench_tmp.log 2>&1
  tail -n 1 ${T}_bench_tmp.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$env', d['ms_per_step'], {k:(round(v['ms_per_launch'],3)) for k,v in d['roofline']['kernels'].items()}, {k: round(v['ms_per_launch'],3) for k,v in d['roofline'].get('other_kernels',{}).items()}, d['e2e']['value'], d['e2e'].get('pinned_buffers_numa_node'))"
done
python experiments/small_grids.py --steps 40 > ${T}_small_grids.log 2>&1
grep -E "device_ms|launches_per" ${T}_small_grids.log
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum"
python experiments/all_kernels.py burgers --steps 1 > ${T}_all_burgers.log 2>&1 &&
ncu $SECT --clock-control none -k regex:"box_kernel" -s 3 -c 3 -f -o /tmp/burgers python experiments/all_kernels.py burgers --steps 1 > ${T}_ncu_burgers.log 2>&1
ncu -i /tmp/burgers.ncu-rep --page raw --csv > ${T}_burgers_raw.csv 2>/dev/null; rm -f /tmp/burgers.ncu-rep
tail -n 3 ${T}_ncu_burgers.log
du -sh gpurun_out
