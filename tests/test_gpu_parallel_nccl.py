"""-m gpu, needs >= 2 GPUs (skipped otherwise): the N > 1 path on NCCL -- every rank runs forward + NMS on its contiguous batch
shard, the detections are gathered (synchronous gather_detections and the asynchronous fused-payload DetectionGatherer), and the
gathered result must equal, byte for byte, what ONE GPU computes on the whole batch (image order = rank-major shard order)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import repyolo_b200 as R
    from oracle import repyolo_oracle as O
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    layers, save, sd, fz = O.make_model(seed=0, mode='calibrated')
    m = R.Model()
    m.load_state_dict(sd, strict=True)
    m.fuse()
    n_images = 4 * world
    x = torch.rand(n_images, 3, 128, 160, generator=torch.Generator().manual_seed(5))
    lo, hi = R.shard_bounds(n_images, rank, world)
    pred, _ = m(x[lo:hi].to(dev))
    out, cnt = R.nms_padded(pred, 0.25, 0.45)
    g_out, g_cnt = R.gather_detections(out, cnt)
    gat = R.DetectionGatherer(hi - lo, 300, dev)
    ok = True
    for step in range(3):                                   # the double-buffered side-stream gather, three steps
        o, c = gat.slot(step)
        R.nms_padded(pred, 0.25, 0.45, out=o, counts=c)
        gat.launch(step)
        a_out, a_cnt = gat.result(step)
        torch.cuda.synchronize(dev)
        ok = ok and torch.equal(a_cnt.reshape(-1), g_cnt)
        ok = ok and all(torch.equal(a_out.reshape(n_images, 300, 6)[i, :k], g_out[i, :k]) for i, k in enumerate(g_cnt.tolist()))
    if rank == 0:                                           # one GPU, whole batch
        w_pred, _ = m(x.to(dev))
        w_out, w_cnt = R.nms_padded(w_pred, 0.25, 0.45)
        ok = ok and torch.equal(w_cnt, g_cnt) and int(w_cnt.sum()) > 0
        ok = ok and all(torch.equal(w_out[i, :k], g_out[i, :k]) for i, k in enumerate(w_cnt.tolist()))
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_gathered_detections_equal_single_gpu_on_nccl():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
