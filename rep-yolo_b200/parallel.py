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


class DetectionGatherer:
    """Asynchronous detection gather ("overlap with the next batch on a side stream").

    One fused payload per rank and step -- the [B_local, max_det, 6] fp32 rows with the [B_local] int32 counts packed
    behind them -- goes out as ONE all_gather on a side stream, double buffered, so the compute stream never waits for the
    slowest rank: step i+1's forward runs while step i's detections travel.  ``slot(i)`` hands out the (out, counts) views
    ry_nms writes into (no staging copy), ``launch(i)`` starts the collective once the compute stream has produced them,
    ``result(i)`` makes the CURRENT stream wait for it and returns ([world, B_local, max_det, 6], [world, B_local]) views
    of the gathered payload (global image order = rank-major)."""

    def __init__(self, b_local: int, max_det: int, device, group=None, n_buffers: int = 2):
        self.group, self.device = group, torch.device(device)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.b, self.max_det, self.nb = b_local, max_det, n_buffers
        self.rows = b_local * max_det * 6
        self.plen = self.rows + ((b_local + 3) // 4) * 4            # floats per rank payload (counts as int32 bit patterns)
        self.send = [torch.zeros(self.plen, dtype=torch.float32, device=self.device) for _ in range(n_buffers)]
        self.recv = [torch.zeros(self.world * self.plen, dtype=torch.float32, device=self.device) for _ in range(n_buffers)]
        on_gpu = self.device.type == 'cuda'
        self.side = torch.cuda.Stream(self.device) if on_gpu else None
        self.produced = [torch.cuda.Event() for _ in range(n_buffers)] if on_gpu else None
        self.done = [torch.cuda.Event() for _ in range(n_buffers)] if on_gpu else None
        self.pending = [False] * n_buffers

    def slot(self, i: int):
        """(out [B_local, max_det, 6] fp32, counts [B_local] int32) views of send buffer i % n_buffers.  The compute stream
        first waits until the collective that last read this buffer has finished."""
        k = i % self.nb
        if self.side is not None and self.pending[k]:
            torch.cuda.current_stream(self.device).wait_event(self.done[k])
        buf = self.send[k]
        return buf[:self.rows].view(self.b, self.max_det, 6), buf[self.rows:self.rows + self.b].view(torch.int32)

    def launch(self, i: int):
        k = i % self.nb
        if self.world == 1:
            self.recv[k].copy_(self.send[k], non_blocking=True)
            return
        if self.side is None:                                       # CPU tensors (gloo): synchronous
            dist.all_gather_into_tensor(self.recv[k], self.send[k], group=self.group)
            return
        self.produced[k].record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.produced[k])
            dist.all_gather_into_tensor(self.recv[k], self.send[k], group=self.group)
            self.done[k].record(self.side)
        self.pending[k] = True

    def views(self, i: int):
        g = self.recv[i % self.nb].view(self.world, self.plen)
        return (g[:, :self.rows].view(self.world, self.b, self.max_det, 6), g[:, self.rows:self.rows + self.b].view(torch.int32))

    def result(self, i: int, stream=None):
        k = i % self.nb
        if self.side is not None and self.pending[k]:
            (stream or torch.cuda.current_stream(self.device)).wait_event(self.done[k])
        return self.views(i)


def to_list(out: torch.Tensor, counts: torch.Tensor):
    """padded -> the reference's list of (n_i, 6) tensors (one host read of the counts)."""
    return [out[i, :c] for i, c in enumerate(counts.cpu().tolist())]
