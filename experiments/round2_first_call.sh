#!/usr/bin/env bash
# First GPU call of round 2 (run under gpurun from the repo root, one GPU):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash experiments/round2_first_call.sh'
# 1. the late round-1 additions that have never run on a B200 (tests/test_gpu_zz_late_additions.py)
# 2. configs[3] (4096x4096x64 diffusion dwarf) with the tiled and the marching kernel
# 3. launch list + one full ncu capture of the diffusion kernels (only after the plain runs exit 0)
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_zz_late_additions.py -q -m gpu > gpurun_out/r02_late_additions.log 2>&1
echo "late additions: rc=$?" | tee -a gpurun_out/r02_summary.log
python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r02_bench_c4_tile.log 2>&1
rc_tile=$?
TB200_DIFF_IMPL=march python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r02_bench_c4_march.log 2>&1
rc_march=$?
echo "c4 tile rc=$rc_tile march rc=$rc_march" | tee -a gpurun_out/r02_summary.log
tail -n 1 gpurun_out/r02_bench_c4_tile.log gpurun_out/r02_bench_c4_march.log | cut -c1-400
if [ $rc_tile -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file gpurun_out/r02_launches_c4_tile.csv \
      python bench.py --workload c4small --steps 3 --warmup 3 > gpurun_out/r02_ncu_c4_tile.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"cross_kernel" -s 6 -c 2 \
      -o gpurun_out/r02_c4_tile python bench.py --workload c4 --steps 3 --warmup 3 \
      > gpurun_out/r02_ncu_full_c4_tile.log 2>&1
fi
if [ $rc_march -eq 0 ]; then
  TB200_DIFF_IMPL=march ncu --set full --clock-control none --import-source on -k regex:"march_kernel" \
      -s 6 -c 2 -o gpurun_out/r02_c4_march python bench.py --workload c4 --steps 3 --warmup 3 \
      > gpurun_out/r02_ncu_full_c4_march.log 2>&1
fi
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c5.log 2>&1
echo "c5 rc=$?" | tee -a gpurun_out/r02_summary.log
tail -n 1 gpurun_out/r02_bench_c5.log | cut -c1-300
