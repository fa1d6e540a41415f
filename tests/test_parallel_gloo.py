"""-m "not gpu": the N>1 host path (contiguous batch shards + detection gather) with world_size 2 on the gloo backend."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import repyolo_b200 as R
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n_images, max_det = 6, 5
    g = torch.Generator().manual_seed(0)
    full_out = torch.rand(n_images, max_det, 6, generator=g)
    full_cnt = torch.randint(0, max_det + 1, (n_images,), generator=g, dtype=torch.int32)
    lo, hi = R.shard_bounds(n_images, rank, world)
    lo0, hi0 = lo, hi
    out, cnt = R.gather_detections(full_out[lo:hi].clone(), full_cnt[lo:hi].clone())
    ok = torch.equal(out, full_out) and torch.equal(cnt, full_cnt)
    lst = R.to_list(out, cnt)
    ok = ok and all(l.shape[0] == int(c) for l, c in zip(lst, full_cnt))
    # uneven shards (7 images over 2 ranks: 4 + 3): padded for the collective, padding dropped afterwards
    n7 = 7
    f_out = torch.rand(n7, max_det, 6, generator=g)
    f_cnt = torch.randint(0, max_det + 1, (n7,), generator=g, dtype=torch.int32)
    lo, hi = R.shard_bounds(n7, rank, world)
    out7, cnt7 = R.gather_detections(f_out[lo:hi].clone(), f_cnt[lo:hi].clone(), n_images=n7)
    ok = ok and torch.equal(out7, f_out) and torch.equal(cnt7, f_cnt)
    # fused-payload gatherer (rows + counts in one collective, double buffered); on CPU tensors it runs synchronously
    gat = R.DetectionGatherer(hi0 - lo0, max_det, 'cpu')
    for step in range(3):
        o, c = gat.slot(step)
        o.copy_(full_out[lo0:hi0] + step)
        c.copy_(full_cnt[lo0:hi0])
        gat.launch(step)
        go, gc = gat.result(step)
        ok = ok and torch.equal(go.reshape(n_images, max_det, 6), full_out + step) and torch.equal(gc.reshape(-1), full_cnt)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_gather_detections_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_shard_bounds_cover():
    import repyolo_b200 as R
    for n in (1, 7, 64, 512):
        for w in (1, 2, 4, 8):
            spans = [R.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
