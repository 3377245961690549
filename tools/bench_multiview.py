#!/usr/bin/env python
"""cfg4 of BASELINE.json: one DMesh++ multi-view optimisation step - 64 views at 512x512 of the 1 M
triangle scene - sharded over the GPUs of one box.

    python tools/bench_multiview.py --steps 5 --warmup 2                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/bench_multiview.py --steps 5 --warmup 2                 # N GPUs

A step = every rank renders its views (view i belongs to rank i mod N) forward + backward in
micro-batches of 8 views through the public Renderer API, accumulates the replicated-parameter
gradients in a flat bucket and the ranks sum the bucket ONCE per step with the library's peer-memory
allreduce kernel over NVLink (dmesh2_renderer_b200.multiview.PeerGradBucket, csrc/dm2_peer.cu).  STRONG scaling: the 64 views are fixed, N grows.  Timed with CUDA
events between barriers, max over ranks; rank 0 prints one JSON line (views/s over the whole job).
(`bench.py --workload cfg4` is the driver-facing version of this workload, with the e2e feed and the allreduce check.)
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--micro-batch", type=int, default=8)
    ap.add_argument("--faces", type=int, default=1_000_000)
    ap.add_argument("--size", type=int, default=512)
    a = ap.parse_args()
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)                      # library banners go to stderr: exactly one line on stdout

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import dmesh2_renderer_b200 as pkg
    from dmesh2_renderer_b200 import multiview, scenes
    pkg.load_library()
    sc = scenes.cfg4(n_views=a.views, size=a.size, n_faces=a.faces)
    t = torch.from_numpy
    mine = multiview.shard_views(a.views, world, rank)
    renderer = pkg.Renderer(t(sc.mv).to(dev), t(sc.proj).to(dev), sc.width, sc.height, dev)
    faces, bg = t(sc.faces).to(dev), t(sc.background).to(dev)
    verts, colour, opacity = (t(getattr(sc, k)).to(dev) for k in ("verts", "verts_color", "faces_opacity"))
    # per-view tensors: only this rank's views live on this GPU
    intense = {v: t(sc.faces_intense[v]).to(dev) for v in mine}
    wc = {v: t(sc.w_color[v]).to(dev) for v in mine}
    wd = {v: t(sc.w_depth[v]).to(dev) for v in mine}
    bucket = multiview.make_grad_bucket(verts.shape[0], faces.shape[0], dev)
    pm = {}

    def render_view_batch(views, v, c, o):
        n = len(views)
        if n not in pm:
            pm[n] = torch.zeros((n, 2), dtype=torch.int32, device=dev)
        fi = torch.stack([intense[i] for i in views]).requires_grad_(True)
        color, depth = renderer(views, pm[n], sc.width, sc.height, v, faces, c, o, fi, bg)
        return (color * torch.stack([wc[i] for i in views])).sum() + (depth * torch.stack([wd[i] for i in views])).sum()

    def step():
        return multiview.multiview_step(render_view_batch, list(range(a.views)), verts, colour, opacity, bucket,
                                        micro_batch=a.micro_batch, world_size=world, rank=rank, renderer=renderer)

    def timed(n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(a.warmup, 1)):
        step()
    ms = timed(a.steps) / a.steps
    # the collective alone (same bucket), for the record
    ar_ms = None
    if world > 1:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        e0.record()
        for _ in range(10):
            bucket.allreduce()
        e1.record()
        torch.cuda.synchronize()
        ar_ms = e0.elapsed_time(e1) / 10
    exchange = ("NCCL allreduce" if not isinstance(bucket, multiview.PeerGradBucket) else
                "peer-memory allreduce kernel (" + ("NVLS multimem" if bucket.uses_multicast else "P2P ld/st") + ")")
    loss, _ = step()
    grad_norm = float(bucket.dverts.norm().item())
    if rank == 0:
        real_stdout.write(json.dumps({
            "metric": "multi-view optimisation step, views/s (fwd+bwd, 64 views @512^2, 1M tris)",
            "value": a.views * 1000.0 / ms, "unit": "views/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{sc.name}: {a.views} views at {a.size}x{a.size}, {a.faces} triangles, view i on "
                                   f"rank i mod {world}, micro-batches of {a.micro_batch} views, one {exchange} "
                                   f"of [dverts|dverts_color|dfaces_opacity] ({bucket.nbytes / 1e6:.0f} MB) per step"},
            "allreduce_ms": ar_ms, "views_per_rank": len(mine), "dverts_norm_after_allreduce": grad_norm}) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
