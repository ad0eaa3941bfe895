"""One-process-per-GPU sharding of independent images (BASELINE.json configs[4]).

The hot path has no cross-image data dependence, so a batch is partitioned
image -> rank with NO pixel traffic between GPUs.  The only collective is ONE
broadcast of the filter parameters (blur taps, resize filter / geometry) from rank 0,
so that every rank convolves with bit-identical weights, followed by whatever
barrier / timing reduction the caller wants.  Backend: NCCL over NVLink on the GPUs,
gloo in the CPU tests (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

MAX_TAPS = 129


@dataclass
class FilterJob:
    """Parameters of the blur + resize pipeline every rank applies to its images."""
    radius: float
    sigma: float
    out_columns: int
    out_rows: int
    resize_filter: int
    taps: np.ndarray          # the 1-D blur taps (float64), as built on rank 0

    def pack(self) -> np.ndarray:
        buf = np.zeros(8 + MAX_TAPS, dtype=np.float64)
        n = int(self.taps.size)
        if n > MAX_TAPS:
            raise ValueError("too many taps")
        buf[:6] = [self.radius, self.sigma, self.out_columns, self.out_rows, self.resize_filter, n]
        buf[8:8 + n] = self.taps
        return buf

    @staticmethod
    def unpack(buf: np.ndarray) -> "FilterJob":
        n = int(buf[5])
        return FilterJob(float(buf[0]), float(buf[1]), int(buf[2]), int(buf[3]), int(buf[4]),
                         np.array(buf[8:8 + n], dtype=np.float64))


def shard_indices(num_images: int, rank: int, world_size: int) -> List[int]:
    """Image i belongs to rank i mod world_size (round-robin keeps ragged batches balanced)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return list(range(rank, num_images, world_size))


def init_process_group(backend: str | None = None):
    """Initialises torch.distributed from the torchrun environment; returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def broadcast_job(job: FilterJob | None, device=None) -> FilterJob:
    """The single collective of the data path: rank 0's filter parameters to everyone."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert job is not None
        return job
    if dist.get_rank() == 0:
        assert job is not None
        t = torch.from_numpy(job.pack())
    else:
        t = torch.zeros(8 + MAX_TAPS, dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return FilterJob.unpack(t.cpu().numpy())


def blur_kernel_from_taps(taps: Sequence[float]):
    """Rebuilds BlurImage's two-kernel list (effect.c:788) from broadcast taps as user kernels:
    an N x 1 row kernel followed by its 1 x N transpose."""
    from . import api
    n = len(taps)
    vals = ",".join(repr(float(v)) for v in taps)
    return api.AcquireKernelInfo(f"{n}x1:{vals};1x{n}:{vals}")


def run_pipeline(images, blur_kernel, job: FilterJob):
    """The configs[4] pipeline on this rank's shard: {index: Image} -> {index: Image}.  BlurImage (as ConvolveImage with
    the broadcast taps) then ResizeImage; device-resident Images stay in HBM, no pixel leaves the GPU."""
    from . import api
    out = {}
    for idx, image in images.items():
        blurred = api.ConvolveImage(image, blur_kernel)
        out[idx] = api.ResizeImage(blurred, job.out_columns, job.out_rows, job.resize_filter)
    return out


def gather_over_ranks(values, device=None):
    """All ranks' float lists (equal length) as a [world][n] nested list on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [list(values)]
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else "cpu")
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [o.cpu().tolist() for o in outs]


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
