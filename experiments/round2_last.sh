#!/usr/bin/env bash
# Last GPU call: the whole parity suite on the final build (fused stage with slow tendencies added),
# headline sanity
set -u
mkdir -p gpurun_out
T=gpurun_out/r02zz
python -m pytest tests -q -m gpu > ${T}_pytest.log 2>&1
echo "pytest: rc=$?" | tee -a ${T}_summary.log
tail -n 8 ${T}_pytest.log
python bench.py --steps 20 --warmup 3 --no-aux --no-cpu-baseline > ${T}_bench_c5.log 2>&1
echo "bench: rc=$?" | tee -a ${T}_summary.log
tail -n 1 ${T}_bench_c5.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c5', d['ms_per_step'], d['value'], d['hbm_frac_step'], {k:(round(v['ms_per_launch'],3), [round(x,3) for x in v.get('ms_by_stage')]) for k,v in d['roofline']['kernels'].items()}, {k: round(v['ms_per_launch'],3) for k,v in d['roofline'].get('other_kernels',{}).items()})"
