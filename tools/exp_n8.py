"""Experiment (r02): what stretched the BlurImage interval 4x at 8 GPUs in r01 (SCALE_r01.json: 1.55 -> 6.07 ms, resize
unchanged)?  Re-creates r01's measurement -- 20 eager steps of ConvolveImage + ResizeImage on ONE 8192^2 image per rank,
events around every operator, a 40-ms timed region -- under torchrun, in three settings:

    quiet                     nothing else running
    pollers                   every rank runs its own `nvidia-smi -lms 20 -i <local>` (what r01's bench.py did)
    quiet again

and prints, per setting, every rank's blur / resize interval medians and the host time per step.
usage: python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/exp_n8.py"""
import statistics
import subprocess
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import imagemagick_b200 as im
from imagemagick_b200 import dist as mdist

rank, world, local = mdist.init_process_group()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
x = im.Image(torch.rand(8192, 8192, 4, device=dev) * 65535)
k = im.AcquireKernelInfo("blur:0x4;blur:0x4+90")
STEPS = 20


def run(tag):
    for _ in range(5):
        im.ResizeImage(im.ConvolveImage(x, k), 4096, 4096, im.LanczosFilter)
    torch.cuda.synchronize()
    mdist.barrier()
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
    blur = [a.elapsed_time(b) for a, b, _ in marks]
    rs = [b.elapsed_time(c) for _, b, c in marks]
    rows = mdist.gather_over_ranks([wall, statistics.median(blur), max(blur), statistics.median(rs), statistics.median(host), max(host)],
                                   device=dev)
    if rank == 0:
        print(f"== {tag}")
        for r, v in enumerate(rows):
            print(f"  rank {r}: wall/step {v[0]:7.3f} ms  blur median {v[1]:6.3f} max {v[2]:7.3f}  resize median {v[3]:6.3f}  "
                  f"host enqueue median {v[4]:6.3f} max {v[5]:7.3f}", flush=True)
    mdist.barrier()


run("quiet")
Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap"
p = subprocess.Popen(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-i", str(local), "-lms", "20"],
                     stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
time.sleep(1.0)
run("one `nvidia-smi -lms 20 -i <local>` poller per rank (r01's bench.py)")
p.terminate()
p.wait()
time.sleep(0.5)
run("quiet again")
try:
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()
except Exception:
    pass
