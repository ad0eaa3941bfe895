"""Experiment (r02): does a fleet of `nvidia-smi -lms 20` pollers (what bench.py r01 started on every rank)
stretch the BlurImage interval?  Times N steps of blur + resize with per-step events and host clocks,
(a) quiet, (b) with K pollers running.  usage: python tools/exp_pollers.py [K] [steps]"""
import statistics
import subprocess
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import imagemagick_b200 as im

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 100
x = im.Image(torch.rand(8192, 8192, 4, device="cuda") * 65535)
k = im.AcquireKernelInfo("blur:0x4;blur:0x4+90")


def run(tag):
    for _ in range(5):
        im.ResizeImage(im.ConvolveImage(x, k), 4096, 4096, im.LanczosFilter)
    torch.cuda.synchronize()
    marks, host = [], []
    t0 = time.perf_counter()
    for _ in range(STEPS):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        h0 = time.perf_counter()
        e[0].record()
        b = im.ConvolveImage(x, k)
        e[1].record()
        im.ResizeImage(b, 4096, 4096, im.LanczosFilter)
        e[2].record()
        host.append((time.perf_counter() - h0) * 1e3)
        marks.append(e)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / STEPS
    blur = sorted(a.elapsed_time(b) for a, b, _ in marks)
    rs = sorted(b.elapsed_time(c) for _, b, c in marks)
    q = lambda v, p: v[min(len(v) - 1, int(p * len(v)))]
    print(f"{tag:28s} wall/step {wall:7.3f} ms | host enqueue/step median {statistics.median(host):6.3f} max {max(host):7.3f} | "
          f"blur min {blur[0]:6.3f} med {q(blur, .5):6.3f} p95 {q(blur, .95):6.3f} max {blur[-1]:7.3f} | "
          f"resize med {q(rs, .5):6.3f} max {rs[-1]:6.3f}", flush=True)


run("quiet")
Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap"
procs = [subprocess.Popen(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-i", "0", "-lms", "20"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for _ in range(K)]
time.sleep(1.0)
run(f"{K} nvidia-smi -lms 20 pollers")
for p in procs:
    p.terminate()
for p in procs:
    p.wait()
time.sleep(0.5)
run("quiet again")
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    import threading
    stop = False

    def poll():
        while not stop:
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetPowerUsage(h)
            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            time.sleep(0.1)
    th = threading.Thread(target=poll, daemon=True)
    th.start()
    run("in-process NVML @100 ms")
    stop = True
    th.join()
except Exception as exc:
    print("pynvml:", exc)
