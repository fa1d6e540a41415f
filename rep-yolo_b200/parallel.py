"""Multi-GPU data path: images are independent, so batches are sharded contiguously across ranks (one process per GPU)
and the ONLY exchange step is a gather of the fixed-shape padded detections.  The reference has no
multi-GPU inference (select_device returns cuda:0, utils/torch_utils.py:86), so the contract is: gathered output ==
single-GPU output on the concatenated batch."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_images: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of rank `rank`; earlier ranks take the remainder."""
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_detections(out: torch.Tensor, counts: torch.Tensor, group=None, n_images: int | None = None):
    """all-gather padded detections: out [B_local, max_det, 6] fp32, counts [B_local] int32
    -> ([n, max_det, 6], [n]) in global image order.  NCCL on CUDA tensors, gloo on CPU.
    Every rank must contribute the same B_local; when the global batch does not divide (``shard_bounds`` gives the first
    ranks one image more) pass ``n_images``: shards are padded to ceil(n_images / world) rows with zero counts for the
    collective and the padding is dropped again afterwards."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return out, counts
    world = dist.get_world_size(group)
    if n_images is not None:
        per = -(-n_images // world)
        lo, hi = shard_bounds(n_images, dist.get_rank(group), world)
        if out.shape[0] != hi - lo:
            raise ValueError(f'gather_detections: rank holds {out.shape[0]} images, shard_bounds says {hi - lo}')
        if hi - lo < per:
            out = torch.cat([out, out.new_zeros((per - (hi - lo),) + tuple(out.shape[1:]))])
            counts = torch.cat([counts, counts.new_zeros(per - (hi - lo))])
        g_out, g_cnt = gather_detections(out, counts, group)
        keep = torch.cat([torch.arange(r * per, r * per + (b[1] - b[0])) for r in range(world)
                          for b in [shard_bounds(n_images, r, world)]]).to(g_out.device)
        return (g_out, g_cnt) if keep.numel() == g_out.shape[0] else (g_out[keep], g_cnt[keep])
    g_out = torch.empty((world * out.shape[0],) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device)
    g_cnt = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(g_out, out.contiguous(), group=group)
    dist.all_gather_into_tensor(g_cnt, counts.contiguous(), group=group)
    return g_out, g_cnt


def to_list(out: torch.Tensor, counts: torch.Tensor):
    """padded -> the reference's list of (n_i, 6) tensors (one host read of the counts)."""
    return [out[i, :c] for i, c in enumerate(counts.cpu().tolist())]
