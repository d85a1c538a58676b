#!/usr/bin/env bash
# Multi-GPU call: gpurun --gpus N --timeout 1500 -- 'bash experiments/round2_mgpu.sh N'
# 1. tests/mgpu_check.py (decomposed run == single-domain run, bitwise) for both transports, with and
#    without communication / computation overlap
# 2. bench.py at N GPUs: peer-store transport (default), with overlap, NCCL transport
set -u
N=${1:-2}
mkdir -p gpurun_out
T=gpurun_out/r02_mgpu${N}
port=29540
for extra in "--transport p2p" "--transport p2p --overlap" "--transport p2p --phases 2" ${MGPU_SKIP_NCCL:-"--transport nccl"} ${MGPU_MORE:+"--transport nccl --overlap"}; do
  [ "$extra" = "1" ] && continue
  port=$((port + 1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      tests/mgpu_check.py $extra >> ${T}_check.log 2>&1
  echo "mgpu_check $extra: rc=$?" | tee -a ${T}_summary.log
done
grep -E "MGPU|differs|Error" ${T}_check.log | cut -c1-200
run_bench () {  # name, env, extra args
  port=$((port + 1))
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 20 --warmup 3 $3 > ${T}_bench_$1.log 2>&1
  echo "bench $1: rc=$?" | tee -a ${T}_summary.log
  tail -n 1 ${T}_bench_$1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'exchange ms/stage', d.get('halo_exchange_ms_per_stage'), 'e2e', round(d['e2e']['value'],1))" 2>/dev/null || tail -n 5 ${T}_bench_$1.log
}
run_bench p2p "TB200_HALO=p2p" ""
run_bench p2p_2phase "TB200_HALO=p2p TB200_HALO_PHASES=2" ""
if [ -n "${MGPU_MORE:-}" ]; then run_bench p2p_overlap "TB200_HALO=p2p" "--overlap"; fi
if [ -z "${MGPU_SKIP_NCCL:-}" ]; then run_bench nccl "TB200_HALO=nccl" ""; fi
python bench.py --gpus 1 --steps 20 --warmup 3 --no-aux --no-cpu-baseline > ${T}_bench_1gpu.log 2>&1
tail -n 1 ${T}_bench_1gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('1gpu ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1))"
