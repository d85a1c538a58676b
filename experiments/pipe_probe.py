"""Launch probe kernels (experiments/pipe_probe.cu) between the stages of the running dry model and
time them: instruction-class throughput in the sustained state of the GPU vs after idle."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench

lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpipe_probe.so"))
lib.probe_launch.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
run = bench.DryRun(1024, 1024, 64)
blocks, threads, iters = 148 * 2, 512, 400
io = torch.rand(blocks * threads + 16, dtype=torch.float64, device="cuda") + 0.5
names = ["DFMA", "DADD", "DMUL", "FFMA", "IMAD"]
n_inst = blocks * threads * iters * 64.0


def probe(op):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream().cuda_stream
    e0.record()
    rc = lib.probe_launch(op, io.data_ptr(), blocks, threads, iters, C.c_void_p(st))
    assert rc == 0, rc
    e1.record()
    return e0, e1


def report(tag, res):
    torch.cuda.synchronize()
    out = []
    for op, (e0, e1) in res:
        ms = e0.elapsed_time(e1)
        out.append("%s %.3f ms = %.1f T/s" % (names[op], ms, n_inst / (ms * 1e-3) / 1e12))
    print(tag, " | ".join(out))


for _ in range(2):
    run.step()
torch.cuda.synchronize()
time.sleep(1.0)
report("idle GPU     :", [(op, probe(op)) for op in range(5)])
time.sleep(1.0)
for rep in range(3):
    for _ in range(25):  # ~0.3 s of continuous stepping
        run.step()
    res = [(op, probe(op)) for op in range(5)]
    run.step()
    report("after 25 steps:", res)
clk = (C.c_ulonglong * 2)()
lib.probe_clock(clk)
print("SM clock during the last probe: %.0f MHz" % (1e3 * clk[0] / clk[1]))
