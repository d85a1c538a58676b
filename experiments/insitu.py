"""Per-kernel stage timings inside a running time loop (CUDA events recorded by the library on
the launching stream), sample by sample, next to nvidia-smi clocks / power sampled every 20 ms:
tells apart what a kernel costs alone (ncu) from what it costs inside the step."""
import ctypes as C
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from tasmania_b200 import lib as tblib

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
run = bench.DryRun(1024, 1024, 64)
handle = tblib.load()
samples = []
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,"
                        "clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown",
                        "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"],
                       stdout=subprocess.PIPE, text=True)
lines = []
threading.Thread(target=lambda: [lines.append((time.time(), l.strip())) for l in smi.stdout], daemon=True).start()
for _ in range(3):
    run.step()
torch.cuda.synchronize()
# 1. plain timed loop
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t_a = time.time()
e0.record()
for _ in range(nsteps):
    run.step()
e1.record()
torch.cuda.synchronize()
t_b = time.time()
print("plain loop: %.3f ms/step" % (e0.elapsed_time(e1) / nsteps))
# 2. profiled loop (host sync after every stage)
orig = run.dyc._stage_fused
gap = float(os.environ.get("INSITU_GAP_MS", "0")) * 1e-3
def timed(stage, state, timestep, out_state):
    if gap > 0:
        time.sleep(gap)  # idle GPU before the stage
    orig(stage, state, timestep, out_state)
    ms = (C.c_double * 3)()
    tblib.check(handle.tb200_stage_profile_read(ms), "read")
    samples.append((stage,) + tuple(ms))
tblib.check(handle.tb200_stage_profile(1), "on")
run.dyc._stage_fused = timed
t_c = time.time()
for _ in range(4):
    run.step()
torch.cuda.synchronize()
t_d = time.time()
for s in samples:
    print("stage %d: A %.3f  B %.3f  MV %.3f" % s)
smi.terminate()
for t, l in lines:
    tag = "plain" if t_a <= t <= t_b else "prof" if t_c <= t <= t_d else ""
    if tag:
        print(tag, l)
