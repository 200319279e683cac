"""Multi-GPU inference: whole images are sharded across ranks; ONE collective at the end.

The reference has no multi-GPU inference (its evaler pins cuda:0, yolov6/core/evaler.py:454-466);
images are independent through forward, decode and NMS, so rank r simply takes the contiguous
slice [r*B/W, (r+1)*B/W) of the batch and the fixed-size NMS outputs (`[b, max_det, 6]` fp32 +
`[b]` int32 counts; 7.2 KB per image) are exchanged with a single all-gather on the compute
stream.  Result order = rank order = image order.  Payloads are KBs, i.e. latency-bound on
NVLink 5: there is nothing to overlap, the point is to have exactly one fixed-size collective.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first `batch % world` ranks get one extra image."""
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pack_detections(det: torch.Tensor, count: torch.Tensor, rows: int) -> torch.Tensor:
    """[b, max_det, 6] + [b] -> one fp32 buffer [rows, max_det*6 + 1] (count in the last column, rows
    beyond b zero) so a single collective moves both."""
    b, max_det, _ = det.shape
    buf = torch.zeros((rows, max_det * 6 + 1), dtype=torch.float32, device=det.device)
    buf[:b, :-1] = det.reshape(b, -1)
    buf[:b, -1] = count.to(torch.float32)
    return buf


def all_gather_detections(det: torch.Tensor, count: torch.Tensor, global_batch: int,
                          group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Every rank passes the (det, count) of its own shard; returns the whole batch's
    (det [B, max_det, 6], count [B]) in image order on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    max_det = det.shape[1]
    rows = -(-global_batch // world)  # every rank contributes the same fixed number of rows
    s, e = shard_range(global_batch, rank, world)
    assert det.shape[0] == e - s, f"rank {rank} owns images [{s},{e}) but passed {det.shape[0]} rows"
    mine = pack_detections(det, count, rows)
    out = torch.empty((world * rows, mine.shape[1]), dtype=torch.float32, device=det.device)
    if det.is_cuda:
        dist.all_gather_into_tensor(out, mine, group=group)
    else:  # gloo (CPU tests of the sharding logic)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        out = torch.cat(parts, 0)
    keep: List[torch.Tensor] = []
    for r in range(world):
        rs, re = shard_range(global_batch, r, world)
        keep.append(out[r * rows: r * rows + (re - rs)])
    full = torch.cat(keep, 0)
    return full[:, :-1].reshape(global_batch, max_det, 6), full[:, -1].to(torch.int32)


class DetectionGather:
    """The same single fixed-size collective WITHOUT staging copies: `copies` persistent buffers of
    shape [world * batch, max_det * 6 + 2] fp32; mafb200_nms_select_packed writes this rank's rows in place (detections
    + the count as int32 bits right behind them), `gather(i)` is one in-place `all_gather_into_tensor`, and `views(i)`
    exposes the whole batch as (det [W*B, max_det, 6], count [W*B] int32) without moving anything.  Every rank owns the
    same number of images (weak scaling / evenly divisible batches); uneven shards use all_gather_detections above."""

    def __init__(self, batch: int, max_det: int, device, group: Optional[dist.ProcessGroup] = None, copies: int = 4):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.batch, self.max_det = batch, max_det
        self.row = max_det * 6 + 2  # even: every row starts 8-byte aligned
        self.bufs = [torch.zeros((self.world * batch, self.row), dtype=torch.float32, device=device) for _ in range(copies)]

    def mine(self, i: int) -> torch.Tensor:
        """This rank's rows of buffer i: the `packed` argument of ops.nms_select_packed."""
        return self.bufs[i][self.rank * self.batch:(self.rank + 1) * self.batch]

    def gather(self, i: int) -> None:
        """Enqueues the collective on the current stream (CUDA) / runs it (gloo)."""
        if self.world == 1:
            return
        buf = self.bufs[i]
        if buf.is_cuda:
            dist.all_gather_into_tensor(buf, self.mine(i), group=self.group)
        else:  # gloo (CPU tests of the same logic)
            dist.all_gather(list(buf.chunk(self.world, 0)), self.mine(i).clone(), group=self.group)

    def views(self, i: int) -> Tuple[torch.Tensor, torch.Tensor]:
        buf = self.bufs[i]
        det = buf[:, :self.max_det * 6].unflatten(1, (self.max_det, 6))
        count = buf.view(torch.int32)[:, self.max_det * 6]
        return det, count
