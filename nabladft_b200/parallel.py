"""Multi-GPU plumbing for the hot path: one process per GPU, `torch.distributed` (NCCL on GPUs,
gloo in the CPU tests).  Conformations are independent units (every graph op is molecule-local:
`batch`-restricted radius graph nablaDFT/painn_pyg/painn.py:411-416, per-molecule energy scatter
:128), so the path SHARDS with no data-path collective: each rank evaluates a contiguous range
of molecules; the only communication is gathering the results (and, in bench.py, a scalar MAX
of the device time).  Mirrors what Lightning's DDPStrategy + DistributedSampler give the
reference at inference (nablaDFT/utils/pipelines.py:65-68).
"""
from typing import Callable, List, Tuple

import torch
import torch.distributed as dist


def balanced_ranges(weights: torch.Tensor, world: int) -> List[Tuple[int, int]]:
    """Split items 0..n-1 into `world` contiguous ranges with near-equal total weight
    (weight = atoms per molecule ~ edges ~ work).  Deterministic; every item in exactly one range."""
    n = int(weights.numel())
    csum = torch.cumsum(weights.to(torch.float64), 0)
    total = float(csum[-1]) if n else 0.0
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(torch.searchsorted(csum, torch.tensor(target, dtype=torch.float64)).item())
        # choose the cut (k or k+1 items) closest to the target
        if k < n and abs(float(csum[k]) - target) < abs((float(csum[k - 1]) if k > 0 else 0.0) - target):
            k += 1
        bounds.append(max(bounds[-1], min(k, n)))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_batch(z: torch.Tensor, pos: torch.Tensor, mol_ptr: torch.Tensor, rank: int, world: int):
    """Returns (z_r, pos_r, mol_ptr_r, (m0, m1)) -- the molecules [m0, m1) owned by `rank`."""
    n_atoms = (mol_ptr[1:] - mol_ptr[:-1]).cpu()
    m0, m1 = balanced_ranges(n_atoms, world)[rank]
    a0, a1 = int(mol_ptr[m0]), int(mol_ptr[m1])
    return z[a0:a1], pos[a0:a1], (mol_ptr[m0:m1 + 1] - mol_ptr[m0]), (m0, m1)


def gather_variable(t: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather of tensors whose first dimension differs per rank (pad to the max, then trim)."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def energy_forces_sharded(fn: Callable, z, pos, mol_ptr, group=None):
    """Evaluate `fn(z, pos, mol_ptr) -> (energy [b], forces [n,3])` on this rank's molecules and
    gather the whole batch's results on every rank, in the original molecule order."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    z_r, pos_r, ptr_r, _ = shard_batch(z, pos, mol_ptr, rank, world)
    if ptr_r.numel() > 1:
        e_r, f_r = fn(z_r, pos_r, ptr_r)
    else:
        e_r = torch.zeros(0, dtype=pos.dtype, device=pos.device)
        f_r = torch.zeros(0, 3, dtype=pos.dtype, device=pos.device)
    return gather_variable(e_r, group), gather_variable(f_r, group)


def max_over_ranks(value: float, device, group=None) -> float:
    """Device-time aggregation used by bench.py: the step time of the job is the slowest rank's."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def allreduce_gradients(params, group=None, average: bool = True) -> int:
    """Data-parallel training (the reference gets it from Lightning's DDPStrategy, utils/pipelines.py:65-68):
    every rank trains on its own molecules, then exactly ONE all-reduce of the flattened gradient (PaiNN: 1.34 M parameters = 5.4 MB)
    per optimiser step -- NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests.  Parameters without a gradient contribute zeros so
    that every rank sends the same layout.  Returns the number of elements reduced."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    dt = params[0].dtype if all(p.dtype == params[0].dtype for p in params) else torch.float32
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(dt) for p in params])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return int(flat.numel())


class GradBucket:
    """Pre-flattened gradient bucket for the ONE all-reduce per optimiser step.

    Every parameter's `.grad` is a VIEW into one contiguous buffer (the parameters' dtype: fp32 for the CUDA models), so the backward pass accumulates straight into the bucket (no
    `torch.cat` / copy-back per step) and the all-reduce is launched on a side stream the moment the backward has been enqueued; the
    optimiser's stream waits for it.  `zero()` instead of `optimizer.zero_grad(set_to_none=True)` keeps the views alive.
    `last_allreduce_ms()` = device time of the last all-reduce (CUDA events on the side stream; None on CPU / world 1)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=self.params[0].dtype, device=dev)  # fp32 for the CUDA models
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        self._cuda = dev.type == "cuda"
        self._side = torch.cuda.Stream(dev) if self._cuda else None
        self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if self._cuda else None
        self._timed = False

    def zero(self) -> None:
        self.flat.zero_()

    def allreduce(self, average: bool = True) -> int:
        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        self._timed = False
        if world > 1:
            if self._cuda:
                cur = torch.cuda.current_stream(self.flat.device)
                self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    self._ev[0].record()
                    dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                    if average:
                        self.flat /= world
                    self._ev[1].record()
                cur.wait_stream(self._side)
                self._timed = True
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                if average:
                    self.flat /= world
        return int(self.flat.numel())

    def last_allreduce_ms(self):
        if not self._timed:
            return None
        self._ev[1].synchronize()
        return self._ev[0].elapsed_time(self._ev[1])
