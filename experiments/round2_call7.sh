#!/usr/bin/env bash
# GPU call (one GPU): two-column s-step kernel (variants test + headline), then the HBM-sized ncu
# evidence of every kernel family, exported to CSV on the box (the .ncu-rep files stay there:
# gpurun_out/ is limited to 64 MiB)
set -u
mkdir -p gpurun_out
T=gpurun_out/r02g
python -m pytest tests/test_gpu_stage_variants.py tests/test_gpu_isentropic.py tests/test_gpu_distributed.py -q -m gpu > ${T}_pytest.log 2>&1
echo "pytest: rc=$?" | tee -a ${T}_summary.log
tail -n 6 ${T}_pytest.log
for impl in two one; do
  TB200_A_IMPL=$impl python bench.py --steps 20 --warmup 3 --no-aux --no-cpu-baseline > ${T}_bench_c5_a$impl.log 2>&1
  tail -n 1 ${T}_bench_c5_a$impl.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('A=$impl', d['ms_per_step'], d['hbm_frac_step'], {k:(round(v['ms_per_launch'],3), [round(x,3) for x in v.get('ms_by_stage')]) for k,v in d['roofline']['kernels'].items()}, {k: round(v['ms_per_launch'],3) for k,v in d['roofline'].get('other_kernels',{}).items()}, d['e2e']['value'])"
done
python experiments/small_grids.py --steps 40 > ${T}_small_grids.log 2>&1
grep -E "device_ms|launches_per" ${T}_small_grids.log
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum"
export_csv () {  # name
  ncu -i /tmp/$1.ncu-rep --page raw --csv > ${T}_$1_raw.csv 2>/dev/null
  rm -f /tmp/$1.ncu-rep
}
python experiments/all_kernels.py moist --size 640 640 64 --steps 1 > ${T}_all_moist.log 2>&1 &&
ncu $SECT --clock-control none -s 70 -c 70 -f -o /tmp/moist python experiments/all_kernels.py moist --size 640 640 64 --steps 1 > ${T}_ncu_moist.log 2>&1
export_csv moist
python experiments/all_kernels.py burgers --steps 1 > ${T}_all_burgers.log 2>&1 &&
ncu $SECT --clock-control none -k regex:"burgers" -s 4 -c 3 -f -o /tmp/burgers python experiments/all_kernels.py burgers --steps 1 > ${T}_ncu_burgers.log 2>&1
export_csv burgers
python experiments/all_kernels.py halo --steps 1 > ${T}_all_halo.log 2>&1 &&
ncu $SECT --clock-control none -k regex:"p2p_kernel" -s 8 -c 8 -f -o /tmp/halo python experiments/all_kernels.py halo --steps 1 > ${T}_ncu_halo.log 2>&1
export_csv halo
python bench.py --workload c4 --steps 3 --warmup 3 > ${T}_plain_c4.log 2>&1 &&
ncu $SECT --clock-control none -k regex:"march2_kernel|fma_fields_kernel|periodic" -s 8 -c 6 -f -o /tmp/c4 python bench.py --workload c4 --steps 3 --warmup 3 > ${T}_ncu_c4.log 2>&1
export_csv c4
ls -la gpurun_out | tail -n 12; du -sh gpurun_out
