"""World-size-2 gloo tests (CPU) of the multi-view sharding and the single gradient allreduce.

The render function here is the CPU oracle wrapped in an autograd Function, standing in for the
CUDA kernels: what is under test is the host logic of dmesh2_renderer_b200/multiview.py (view
ownership, micro-batching, flat bucket layout, one all_reduce per step, equality with the
single-rank gradient).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_balanced_shards():
    from dmesh2_renderer_b200.multiview import balanced_shards
    import random
    rng = random.Random(3)
    for n, g in [(64, 8), (64, 4), (10, 4), (3, 8), (0, 2)]:
        costs = [1.0 + rng.random() * (3.0 if i % 5 == 0 else 0.3) for i in range(n)]
        sh = balanced_shards(costs, g)
        assert sorted(v for s in sh for v in s) == list(range(n))            # a partition
        assert max(len(s) for s in sh) - min(len(s) for s in sh) <= 1        # same number of views per rank
        assert sh == balanced_shards(list(costs), g)                         # deterministic
        if n >= 2 * g:
            load = [sum(costs[v] for v in s) for s in sh]
            rr = [sum(costs[v] for v in range(r, n, g)) for r in range(g)]
            assert max(load) <= max(rr) + 1e-9                               # never worse than round-robin here
    assert balanced_shards([5, 1, 1, 1], 2) == [[0], [1, 2, 3]] or balanced_shards([5, 1, 1, 1], 2) == [[0, 3], [1, 2]]


def test_multiview_step_rejects_a_bad_ownership_table():
    from dmesh2_renderer_b200.multiview import multiview_step
    v, c, o = torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(2)
    fn = lambda views, a, b, d: (a.sum() + b.sum() + d.sum()) * float(len(views))
    for bad in ([[0, 1], [1, 2]], [[0, 1, 2]], [[0], [1], [2]]):
        with pytest.raises(ValueError):
            multiview_step(fn, [0, 1, 2], v, c, o, world_size=2, rank=0, shards=bad)
    loss, b = multiview_step(fn, [0, 1, 2], v, c, o, world_size=1, rank=0, shards=[[0, 1, 2]])
    assert float(b.dverts.sum()) == 3.0 * 12


def test_shard_views_partition():
    from dmesh2_renderer_b200.multiview import micro_batches, shard_views
    for n, g in ((64, 8), (64, 4), (7, 2), (3, 4), (0, 2)):
        shards = [shard_views(n, g, r) for r in range(g)]
        assert sorted(sum(shards, [])) == list(range(n))
        assert all(all(v % g == r for v in s) for r, s in enumerate(shards))
    assert micro_batches(list(range(10)), 4) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)


def test_bucket_views_alias_flat_buffer():
    from dmesh2_renderer_b200.multiview import GradBucket
    b = GradBucket(5, 3, "cpu")
    assert b.flat.numel() == 33 and b.nbytes == 132
    b.accumulate(torch.ones(5, 3), 2 * torch.ones(5, 3), 3 * torch.ones(3))
    assert torch.equal(b.flat, torch.cat((torch.ones(15), 2 * torch.ones(15), 3 * torch.ones(3))))
    b.zero_()
    assert b.flat.abs().sum() == 0


class _OracleRender(torch.autograd.Function):
    """RenderFunction-shaped autograd wrapper around the CPU oracle (tests only)."""

    @staticmethod
    def forward(ctx, verts, verts_color, faces_opacity, base_args, wh, w_color, w_depth):
        from helpers import args_to_numpy
        from oracle import oracle as orc
        args = list(base_args)
        args[4], args[6], args[7] = verts, verts_color, faces_opacity
        a = args_to_numpy(args)
        out = orc.render(a, wh[0], wh[1], 1.0, K=0, nthreads=2, backward=True, dL_dcolor=w_color, dL_ddepth=w_depth)
        ctx.g = tuple(torch.from_numpy(out[k].astype(np.float32)) for k in ("dL_dverts", "dL_dverts_color", "dL_dfaces_opacity"))
        loss = float((out["color"].astype(np.float64) * w_color).sum() + (out["depth"].astype(np.float64) * w_depth).sum())
        return torch.tensor(loss, dtype=torch.float32)

    @staticmethod
    def backward(ctx, g):
        return ctx.g[0] * g, ctx.g[1] * g, ctx.g[2] * g, None, None, None, None


def _make_render_fn():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from dmesh2_renderer_b200 import scenes
    from helpers import render_args
    sc = scenes.small(n_faces=150, size=32, sigma=0.2, seed=17, n_views=4)
    cache = {}

    def render_view_batch(views, verts, verts_color, faces_opacity):
        total = None
        for v in views:          # the oracle renders one view at a time here
            if v not in cache:
                cache[v] = render_args(sc, "cpu", views=[v])[2]
            loss = _OracleRender.apply(verts, verts_color, faces_opacity, cache[v], (sc.width, sc.height),
                                       sc.w_color[v:v + 1], sc.w_depth[v:v + 1])
            total = loss if total is None else total + loss
        return total
    t = torch.from_numpy
    return sc, render_view_batch, (t(sc.verts), t(sc.verts_color), t(sc.faces_opacity))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dmesh2_renderer_b200.multiview import GradBucket, make_grad_bucket, multiview_step
        sc, fn, (v, c, o) = _make_render_fn()
        # under gloo / on the CPU the factory must hand out the portable bucket (torch.distributed.all_reduce)
        bucket = make_grad_bucket(v.shape[0], o.shape[0], "cpu")
        assert type(bucket) is GradBucket
        loss, bucket = multiview_step(fn, list(range(4)), v, c, o, bucket, micro_batch=1)
        total = loss.clone()
        dist.all_reduce(total)
        first = bucket.flat.numpy().copy()
        # cost-aware ownership (balanced_shards): an uneven table - rank 0 renders view 3 only - must give the same sum
        from dmesh2_renderer_b200.multiview import balanced_shards
        shards = balanced_shards([1.0, 1.0, 1.0, 3.0], world)
        assert sorted(sum(shards, [])) == [0, 1, 2, 3]
        loss2, bucket = multiview_step(fn, list(range(4)), v, c, o, bucket, micro_batch=1, shards=shards)
        total2 = loss2.clone()
        dist.all_reduce(total2)
        assert np.allclose(bucket.flat.numpy(), first, rtol=1e-5, atol=1e-6 * np.abs(first).max())
        assert abs(float(total2) - float(total)) <= 1e-5 * abs(float(total))
        q.put((rank, first, float(total)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_allreduce_equals_single_rank(built):
    from dmesh2_renderer_b200.multiview import multiview_step
    sc, fn, (v, c, o) = _make_render_fn()
    loss1, b1 = multiview_step(fn, list(range(4)), v, c, o, micro_batch=2, world_size=1, rank=0)
    single = b1.flat.numpy().copy()
    assert np.abs(single).sum() > 0

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, flat, total in res:
        assert np.allclose(flat, single, rtol=1e-5, atol=1e-6), f"rank {rank}: allreduced bucket != single-rank gradient"
        assert abs(total - float(loss1)) <= 1e-4 * abs(float(loss1)) + 1e-4


def test_make_grad_bucket_is_the_portable_bucket_without_a_multi_gpu_nccl_job():
    from dmesh2_renderer_b200.multiview import GradBucket, PeerGradBucket, make_grad_bucket
    b = make_grad_bucket(4, 2, "cpu")
    assert type(b) is GradBucket and not isinstance(b, PeerGradBucket) and b.flat.numel() == 26
    b.allreduce()          # no process group: a no-op
