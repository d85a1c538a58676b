# -*- coding: utf-8 -*-
"""Multi-GPU parity check, launched by torchrun on a box with >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29520 tests/mgpu_check.py [--overlap] [--transport p2p|nccl]

Every rank runs its share of the decomposed dry core (halo exchange by NVLink peer stores or by
NCCL messages; with --overlap on a side stream under the interior blocks of the momentum kernel)
and, on the same GPU, the single-domain run of
the whole global grid; after a few steps its owned block must equal the corresponding block of
the single-domain result BITWISE."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tasmania_b200.distributed import DecomposedDryRun, InProcessDecomposedRun  # noqa: E402


def main():
    overlap = "--overlap" in sys.argv
    transport = sys.argv[sys.argv.index("--transport") + 1] if "--transport" in sys.argv else None
    if "--phases" in sys.argv:  # peer-store transport: 1 (faces + corners, default) or 2 phases
        os.environ["TB200_HALO_PHASES"] = sys.argv[sys.argv.index("--phases") + 1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    nx, ny, nz, steps = 150, 140, 16, 4
    # a short mountain growth so that the flow develops within the test
    run = DecomposedDryRun(nx, ny, nz, rank, world, overlap=overlap, transport=transport, topo_seconds=20.0)
    d = run.decomp
    single = InProcessDecomposedRun(d.NX, d.NY, nz, 1, 1, domain_x=run.domain_x, domain_y=run.domain_y,
                                    topo_seconds=20.0)
    for _ in range(steps):
        run.step()
        single.step()
    torch.cuda.synchronize()
    if run.transport == "p2p":
        run.sub.halo.check()
    i0, i1, j0, j1 = d.owned(rank)
    ok = True
    if not float(np.abs(single.gather(run.sub.SV)).max()) > 0.0:
        ok = False
        print(f"rank {rank}: the flow has not developed", flush=True)
    for name in run.sub.names:
        mine = run.sub.owned_numpy(name)
        ref = single.gather(name)[i0:i1, j0:j1, :]
        if not np.array_equal(mine, ref):
            ok = False
            print(f"rank {rank}: {name} differs, max abs {np.abs(mine - ref).max():.3e}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU-OK" if int(flag.item()) == 1 else "MGPU-FAIL",
              f"world={world} decomposition={run.decomposition} overlap={overlap} transport={run.transport} "
              f"phases={getattr(run.sub.halo, 'phases', 2)}",
              flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
