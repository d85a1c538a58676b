"""ms per step over a long plain loop (one CUDA event per step) next to nvidia-smi samples:
shows when the sustained-load power management kicks in and what it costs."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 80
run = bench.DryRun(1024, 1024, 64)
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,power.draw.instant,temperature.gpu,"
                        "clocks_event_reasons.sw_power_cap,clocks_event_reasons.active",
                        "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"],
                       stdout=subprocess.PIPE, text=True)
lines = []
threading.Thread(target=lambda: [lines.append((time.time(), l.strip())) for l in smi.stdout], daemon=True).start()
for _ in range(2):
    run.step()
torch.cuda.synchronize()
time.sleep(1.0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps + 1)]
t0 = time.time()
ev[0].record()
for n in range(nsteps):
    run.step()
    ev[n + 1].record()
torch.cuda.synchronize()
t1 = time.time()
ms = [ev[n].elapsed_time(ev[n + 1]) for n in range(nsteps)]
print("finite:", {n: bool(torch.isfinite(run.state[n].t).all()) for n in run.out_names})
print("ms per step:", " ".join("%.2f" % m for m in ms))
smi.terminate()
for t, l in lines:
    if t0 - 0.2 <= t <= t1 + 0.1:
        print("%.2f s: %s" % (t - t0, l))
