#!/usr/bin/env python
"""Host<->device bandwidth per GPU alone and with all GPUs copying at once (pinned host memory), plus
NVLink peer copy.  Tells what srb_multi_eval's PCIe-parallel design can gain on this box."""
import time, torch
G = torch.cuda.device_count()
MB = 96
n = MB * 1024 * 1024 // 8
host = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(G)]
one = torch.empty(n * G, dtype=torch.float64).pin_memory()
dev = [torch.empty(n, dtype=torch.float64, device="cuda:%d" % i) for i in range(G)]
streams = [torch.cuda.Stream(device=i) for i in range(G)]
def sync():
    for i in range(G):
        torch.cuda.synchronize(i)
def run(devs, direction, src_shared=False, reps=5):
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in devs:
            with torch.cuda.stream(streams[i]):
                h = one[i * n:(i + 1) * n] if src_shared else host[i]
                if direction == "h2d":
                    dev[i].copy_(h, non_blocking=True)
                elif direction == "d2h":
                    h.copy_(dev[i], non_blocking=True)
                else:
                    dev[i].copy_(h, non_blocking=True)
        if direction == "both":
            pass
    sync()
    dt = (time.perf_counter() - t0) / reps
    return len(devs) * MB / 1024 / dt
for i in range(G):
    print("gpu %d alone: h2d %.1f GB/s, d2h %.1f GB/s" % (i, run([i], "h2d"), run([i], "d2h")))
for k in (2, 4, 8):
    if k <= G:
        print("%d gpus at once: h2d %.1f GB/s aggregate (separate buffers), %.1f (slices of one buffer), d2h %.1f" %
              (k, run(list(range(k)), "h2d"), run(list(range(k)), "h2d", True), run(list(range(k)), "d2h")))
# duplex on one GPU
s2 = torch.cuda.Stream(device=0)
d2 = torch.empty(n, dtype=torch.float64, device="cuda:0")
h2 = torch.empty(n, dtype=torch.float64).pin_memory()
sync(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(streams[0]):
        dev[0].copy_(host[0], non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
sync(); dt = (time.perf_counter() - t0) / 5
print("gpu 0 duplex: %.1f GB/s each way" % (MB / 1024 / dt))
if G > 1:
    sync(); t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(streams[0]):
            dev[1].copy_(dev[0], non_blocking=True)
    sync(); dt = (time.perf_counter() - t0) / 5
    print("peer copy 0 -> 1: %.1f GB/s" % (MB / 1024 / dt))
import subprocess
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
