"""Worker of tests/test_multi_gpu.py: one process per GPU (torch.distributed.run, NCCL).  Exits non-zero on a
parity failure; rank 0 writes a JSON summary.

    infer: the global batch sharded over the ranks reproduces, bit for bit, what ONE GPU computes for the whole batch
           (images are independent units in eval / frozen-BN mode; no data-path collective).
    train: BASELINE.json config 4's data-parallel step -- per-rank forward + backward of its shard, per-stage flat
           gradient buckets all-reduced (averaged) by NCCL on a side stream under the backward kernels, grids sized
           to leave 8 SMs to NCCL -- gives the gradients of the single-process step over the concatenated batch,
           divided by the world size; then a second backward WITHOUT zero_grad accumulates to exactly twice that.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tests import helpers  # noqa: E402


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    mode, out_path = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_MAX_CTAS", "8")
    dist.init_process_group("nccl", device_id=dev)
    per_rank = 2
    n = per_rank * world
    g = torch.Generator().manual_seed(77)
    x_all = torch.randn(n, 3, 192, 256, generator=g).to(torch.bfloat16)
    summary = {"mode": mode, "world": world}
    ok = True
    if mode == "infer":
        bb, neck = helpers.build_product_pair(50, seed=5, bnstats=True)
        bb, neck = bb.to(dev).eval(), neck.to(dev).eval()
        with torch.no_grad():
            mine = neck(bb(x_all[rank * per_rank:(rank + 1) * per_rank].to(dev)))
        gathered = []
        for o in mine:
            parts = [torch.empty_like(o) for _ in range(world)]
            dist.all_gather(parts, o.contiguous(memory_format=torch.channels_last))
            gathered.append(torch.cat(parts, 0))
        if rank == 0:
            with torch.no_grad():
                whole = neck(bb(x_all.to(dev)))
            same = [bool(torch.equal(a, b)) for a, b in zip(gathered, whole)]
            summary["levels_bit_identical"] = same
            ok = all(same)
    else:
        from torch_detection_b200 import training
        from oracle import resnet_fpn_oracle as orc  # noqa: F401  (checker side only)
        bb, neck = helpers.build_product_pair(50, seed=5, bnstats=True, frozen_stages=1, bn_eval=True, bn_frozen=True)
        bb, neck = bb.to(dev).train(), neck.to(dev).train()
        sync = training.BucketAllReduce(defer=True, sm_reserve=8)
        bb.set_grad_sync(sync)
        neck.set_grad_sync(sync)
        shapes = [(n, 256, 48 >> i, 64 >> i) for i in range(4)] + [(n, 256, 3, 4)]
        g_all = [(torch.randn(s, generator=g)).to(torch.bfloat16) for s in shapes]
        sl = slice(rank * per_rank, (rank + 1) * per_rank)
        params = [p for p in list(bb.parameters()) + list(neck.parameters()) if p.requires_grad]

        def backward_shard():
            outs = neck(bb(x_all[sl].to(dev)))
            torch.autograd.backward(list(outs), [t[sl].to(dev).contiguous(memory_format=torch.channels_last)
                                                 for t in g_all])
            sync.finish()
            torch.cuda.synchronize()

        backward_shard()
        once = [p.grad.detach().clone() for p in params]
        summary["allreduce_mb_per_step"] = sync.bytes_reduced / 1e6
        summary["buckets"] = sync.buckets_reduced
        backward_shard()   # accumulates into the existing gradients
        twice = [p.grad.detach().clone() for p in params]
        # every rank holds the same averaged gradients
        for t in once:
            ref = t.clone()
            dist.broadcast(ref, 0)
            if not torch.equal(ref, t):
                ok = False
                summary["ranks_disagree"] = True
        if rank == 0:
            from torch_detection_b200 import engine
            engine.set_sm_reserve(dev, 0)
            bb1, neck1 = helpers.build_product_pair(50, seed=5, bnstats=True, frozen_stages=1, bn_eval=True,
                                                    bn_frozen=True)
            bb1, neck1 = bb1.to(dev).train(), neck1.to(dev).train()
            outs = neck1(bb1(x_all.to(dev)))
            torch.autograd.backward(list(outs), [t.to(dev).contiguous(memory_format=torch.channels_last) for t in g_all])
            torch.cuda.synchronize()
            single = [p.grad for p in list(bb1.parameters()) + list(neck1.parameters()) if p.requires_grad]
            e1 = max(rel_l2(a, b / world) for a, b in zip(once, single))
            e2 = max(rel_l2(a, 2 * b / world) for a, b in zip(twice, single))
            summary.update(params=len(single), max_rel_l2_vs_single_process=e1, max_rel_l2_accumulated=e2,
                           buckets_not_overlapped=sync.buckets_not_overlapped)
            ok = ok and e1 <= 1e-3 and e2 <= 1e-3 and sync.buckets_not_overlapped > 0
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        summary["ok"] = bool(flag.item())
        with open(out_path, "w") as f:
            json.dump(summary, f)
        print(json.dumps(summary))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
