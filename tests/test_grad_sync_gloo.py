"""Host-side logic of the training configuration's gradient all-reduce on CPU (gloo, world size 2):
bucket layout (GradBucket) and BucketAllReduce semantics (average, copy-out, counters)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    from torch_detection_b200 import training
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(64, 32, 3, 3)), torch.nn.Parameter(torch.zeros(100)),
                  torch.nn.Parameter(torch.zeros(7, 5, 1, 1))]
        bucket = training.GradBucket(params, torch.device("cpu"))
        assert all(o % 64 == 0 for o in bucket.offsets)                 # 256-byte aligned views
        assert bucket.view(1).shape == (100,) and bucket.view(2).shape == (7, 5, 1, 1)
        assert bucket.index_of(params[2]) == 2
        for i in range(3):
            bucket.view(i).fill_(float(rank + 1) * (i + 1))
        sync = training.BucketAllReduce(average=True, defer=True)
        assert sync.world == world
        red = sync.reduce(bucket.flat, bucket.params)
        sync.finish()
        # the plan's accumulator is left untouched (it is re-zeroed by the next backward); the copy is reduced
        assert float(bucket.view(0).flatten()[0]) == float(rank + 1)
        want = sum(r + 1 for r in range(world)) / world
        for i in range(3):
            v = red[bucket.offsets[i]:bucket.offsets[i] + params[i].numel()]
            assert torch.allclose(v, torch.full_like(v, want * (i + 1)))
        summed = training.BucketAllReduce(average=False).reduce(bucket.flat)
        assert float(summed[0]) == sum(r + 1 for r in range(world))
        assert sync.buckets_reduced == 1 and sync.bytes_reduced == bucket.flat.numel() * 4
        out.put((rank, "ok"))
    except Exception as e:  # surface the failure to the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_bucket_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_bucket_allreduce_single_process():
    from torch_detection_b200 import training
    b = training.GradBucket([torch.nn.Parameter(torch.zeros(8, 8))], torch.device("cpu"))
    b.flat.fill_(3.0)
    sync = training.BucketAllReduce()
    out = sync.reduce(b.flat)
    sync.module_done()
    assert out.data_ptr() != b.flat.data_ptr() and float(out[0]) == 3.0


def test_world_size_is_queried_lazily():
    """The object may be constructed before init_process_group (it used to sample the world size in __init__ and
    silently skip the collective)."""
    from torch_detection_b200 import training
    assert not dist.is_initialized()
    sync = training.BucketAllReduce(defer=True)
    assert sync.world == 1
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        assert sync.world == dist.get_world_size() == 1
    finally:
        dist.destroy_process_group()
