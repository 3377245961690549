#!/usr/bin/env python
"""Checks and times the peer-memory gradient allreduce (csrc/dm2_peer.cu, multiview.PeerGradBucket)
against torch.distributed.all_reduce (NCCL) on the same data.  Run under torchrun on N GPUs of one box:

    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29521 tools/check_peer_allreduce.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, iters, dev):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from dmesh2_renderer_b200 import load_library
    from dmesh2_renderer_b200.multiview import PeerGradBucket
    load_library()
    out = {"world": world}
    cases = {}
    for fused in (True, False):
        tag = "fused" if fused else "unfused"
        cases.update({f"small_p2p_{tag}": (1001, 333, False, fused), f"small_mc_{tag}": (1001, 333, True, fused),
                      f"tiny_p2p_{tag}": (1, 2, False, fused),          # fewer 16-byte vectors than ranks: empty slices
                      f"cfg2_p2p_{tag}": (3_000_000, 1_000_000, False, fused),
                      f"cfg2_mc_{tag}": (3_000_000, 1_000_000, True, fused)})
    for label, (P, F, mc, fused) in cases.items():
        b = PeerGradBucket(P, F, dev, use_multicast=mc, fused_barriers=fused)
        if mc and not b.uses_multicast:
            out[label] = "no multicast support"
            continue
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        data = torch.randn(b.flat.numel(), device=dev, generator=g) * (1.0 + rank)
        ref = data.clone()
        dist.all_reduce(ref)
        worst = 0.0
        for _ in range(5):                       # repeated use of the same bucket / barriers (epochs)
            b.flat.copy_(data)
            b.allreduce()
            torch.cuda.synchronize()
            err = float(((b.flat - ref).abs().max() / ref.abs().max()).item())
            worst = max(worst, err)
        # identical bits on every rank?
        chk = b.flat.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        res = {"max_rel_err_vs_nccl": worst, "same_on_all_ranks": bool((lo == hi).item()), "bytes": b.nbytes}
        if P > 100_000:
            scratch = data.clone()
            res["peer_ms"] = timed(lambda: b.allreduce(), 20, dev)
            res["nccl_ms"] = timed(lambda: dist.all_reduce(scratch), 20, dev)
        out[label] = res
        assert worst < 1e-5, (label, worst)
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
