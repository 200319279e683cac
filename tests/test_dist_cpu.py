"""World-size-2 `gloo` test of the sharding + single fixed-size detection all-gather (host logic of
maf_yolo_b200.dist; the NCCL path is the same code with CUDA tensors)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maf_yolo_b200 import dist as mdist


def test_shard_range_is_a_partition():
    for batch in (1, 2, 7, 32, 33, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [mdist.shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, max_det, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        det_all = torch.rand(batch, max_det, 6, generator=g)
        cnt_all = torch.randint(0, max_det + 1, (batch,), generator=g, dtype=torch.int32)
        s, e = mdist.shard_range(batch, rank, world)
        det, cnt = mdist.all_gather_detections(det_all[s:e].clone(), cnt_all[s:e].clone(), batch)
        ok = torch.equal(det, det_all) and torch.equal(cnt, cnt_all)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 5])
def test_all_gather_detections_gloo_world2(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def _gather_worker(rank, world, port, batch, max_det, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        det_all = torch.rand(world * batch, max_det, 6, generator=g)
        cnt_all = torch.randint(0, max_det + 1, (world * batch,), generator=g, dtype=torch.int32)
        gat = mdist.DetectionGather(batch, max_det, "cpu", copies=2)
        ok = True
        for i in range(2):  # both buffers; what mafb200_nms_select_packed writes on the GPU is written here by hand
            mine = gat.mine(i)
            mine.zero_()
            mine[:, :max_det * 6] = det_all[rank * batch:(rank + 1) * batch].reshape(batch, -1) + i
            mine.view(torch.int32)[:, max_det * 6] = cnt_all[rank * batch:(rank + 1) * batch]
            gat.gather(i)
            det, cnt = gat.views(i)
            ok = ok and torch.equal(det, det_all + i) and torch.equal(cnt, cnt_all) and det.shape == (world * batch, max_det, 6)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_detection_gather_in_place_gloo_world2():
    """DetectionGather: every rank fills its rows of the persistent buffer in place (detections + count bits), ONE
    collective, views of the whole batch — the host logic of the multi-GPU path with the gloo backend."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, 3, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
