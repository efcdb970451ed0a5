"""Multi-GPU sharding of independent image pairs.

`match()` has no cross-pair dependency (BatchNorm is in eval mode, `matcher.py:790`), so a batch of pairs shards
contiguously over ranks with replicated weights and no collective on the data path.  The only communication a caller
may want is distribution of inputs that live on one rank and collection of the results — plain NCCL
scatter / gather over NVLink (25.4 MB in, 29.9 MB out per pair: ~0.07 ms at 770 GB/s against >= 25 ms of compute per
pair, so nothing is fused with compute).  One process per GPU; works with any `torch.distributed` backend (the CPU tests
use gloo with world size 2).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_pairs: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) slice of the pair batch for every rank; the first `n % world` ranks get one more."""
    base, extra = divmod(n_pairs, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def scatter_pairs(tensors: Optional[List[torch.Tensor]], n_pairs: int, shape_tail: List[Tuple[int, ...]], device, src: int = 0,
                  group=None) -> List[torch.Tensor]:
    """Rank `src` holds `tensors` (each [n_pairs, ...]); every rank receives its contiguous shard of each of them.
    `shape_tail` gives the per-pair shape of each tensor so that receivers can allocate without a metadata exchange."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(n_pairs, world)
    lo, hi = bounds[rank]
    out = []
    for i, tail in enumerate(shape_tail):
        recv = torch.empty((hi - lo,) + tuple(tail), dtype=torch.float32, device=device)
        if rank == src:
            full = tensors[i].to(device=device, dtype=torch.float32)
            for r, (a, b) in enumerate(bounds):
                if r == src:
                    recv.copy_(full[a:b])
                elif b > a:
                    dist.send(full[a:b].contiguous(), dst=r, group=group)
        elif hi > lo:
            dist.recv(recv, src=src, group=group)
        out.append(recv)
    return out


def gather_results(local: List[torch.Tensor], n_pairs: int, dst: int = 0, group=None) -> Optional[List[torch.Tensor]]:
    """Inverse of `scatter_pairs` for the per-pair results (warp, certainty): returns the full tensors on `dst`."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(n_pairs, world)
    if rank != dst:
        for t in local:
            if t.shape[0] > 0:
                dist.send(t.contiguous(), dst=dst, group=group)
        return None
    out = []
    for t in local:
        full = torch.empty((n_pairs,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        for r, (a, b) in enumerate(bounds):
            if r == dst:
                full[a:b].copy_(t)
            elif b > a:
                dist.recv(full[a:b], src=r, group=group)
        out.append(full)
    return out


def match_sharded(model, im_A, im_B, im_A_high_res=None, im_B_high_res=None, n_pairs: Optional[int] = None, src: int = 0,
                  group=None):
    """`model.match` over a pair batch sharded across the process group.  Rank `src` passes the full tensors (other ranks
    pass None and `n_pairs` + shapes via the model's configured resolutions); returns (warp, certainty) on `src`, None
    elsewhere."""
    rank = dist.get_rank(group)
    device = model._get_device()
    h, w = model.h_resized, model.w_resized
    tails = [(3, h, w), (3, h, w)]
    tensors = [im_A, im_B]
    if model.upsample_preds:
        hu, wu = model.upsample_res
        tails += [(3, hu, wu), (3, hu, wu)]
        tensors += [im_A_high_res, im_B_high_res]
    if rank == src:
        n_pairs = im_A.shape[0]
    assert n_pairs is not None, "non-source ranks must pass n_pairs"
    shards = scatter_pairs(tensors if rank == src else None, n_pairs, tails, device, src, group)
    if shards[0].shape[0] > 0:
        kw = dict(im_A_high_res=shards[2], im_B_high_res=shards[3]) if model.upsample_preds else {}
        warp, cert = model.match(shards[0], shards[1], **kw)
    else:
        ho, wo = model.get_output_resolution()
        wout = 2 * wo if model.symmetric else wo
        warp = torch.empty(0, ho, wout, 4, device=device)
        cert = torch.empty(0, ho, wout, device=device)
    return gather_results([warp, cert], n_pairs, src, group)
