#!/usr/bin/env bash
# 4-GPU call (2 x 2 decomposition): parity of the peer-store exchange with corner blocks on real GPUs,
# and the scaling point
set -u
mkdir -p gpurun_out
T=gpurun_out/r02_mgpu4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 \
    tests/mgpu_check.py --transport p2p >> ${T}_check.log 2>&1
echo "mgpu_check p2p: rc=$?" | tee -a ${T}_summary.log
grep -E "MGPU|differs" ${T}_check.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 \
    bench.py --gpus 4 --steps 20 --warmup 3 > ${T}_bench_p2p.log 2>&1
echo "bench p2p: rc=$?" | tee -a ${T}_summary.log
tail -n 1 ${T}_bench_p2p.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('4 GPUs ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],1), 'exchange ms/stage', d.get('halo_exchange_ms_per_stage'), 'e2e', round(d['e2e']['value'],1))"
