#!/usr/bin/env python
"""bench.py - headline benchmark of the dmesh2_renderer hot path on B200.

Metric (BASELINE.json): fwd+bwd frames/s, 1M triangles @ 1024x1024 (cfg2), reported at 1/2/4/8 GPUs,
plus the fraction of the measured HBM roofline of the dominant kernel.

    python bench.py --gpus 1 --steps 20 --warmup 5                 # this implementation
    python bench.py --impl reference --gpus 1 --steps 20 --warmup 5   # the unmodified reference
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...  # one rank per GPU (weak scaling)
    ... bench.py --workload cfg4 ...    # BASELINE.json configs[3]: 64 views @512^2 sharded over the ranks (strong scaling)

A "step" is one optimisation step of one view per GPU: Renderer.forward + loss + backward through
the public API (and, for N > 1, ONE sum of [dverts | dverts_color | dfaces_opacity] over the ranks with the
library's peer-memory allreduce kernel over NVLink, csrc/dm2_peer.cu - the gradients are accumulated by the
backward kernels straight into the bucket, and the summed bucket is checked against torch.distributed.all_reduce
after the timed region; the reference arm uses NCCL).  --workload cfg4: a step renders all 64 views at 512^2,
8 views per micro-batch, views assigned to the ranks by measured per-view cost (multiview.balanced_shards;
--no-balance: view i on rank i % N).
`value` is measured with the scene parameters resident in HBM; `e2e` additionally copies every step's
inputs (verts, colours, opacity, intensity) from pinned host memory inside the timed region (copy stream,
double buffered: the copy for step i+1 overlaps the compute of step i) and copies the loss back to pinned
host memory each step (asynchronously; verified finite after the timed region).
Prints exactly one JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--faces", type=int, default=1_000_000)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"])
    ap.add_argument("--views", type=int, default=64, help="cfg4: views per optimisation step")
    ap.add_argument("--micro-batch", type=int, default=8, help="cfg4: views per Renderer call")
    ap.add_argument("--exchange", default="peer", choices=["peer", "peer_unfused", "nccl"],
                    help="N > 1: peer-memory kernel with in-kernel barriers (default), with two separate barrier "
                         "launches, or torch.distributed.all_reduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-balance", action="store_true",
                    help="cfg4, N > 1: keep view i on rank i %% N instead of balancing the shards by measured per-view cost")
    ap.add_argument("--clock-window-s", type=float, default=1.2,
                    help="untimed repeat of the timed loop during which nvidia-smi samples the clocks")
    return ap.parse_args()


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi (100 ms period) from before the warm-up until after
    an untimed repeat of the timed loop; only samples taken between mark() and stop() (timed region + repeat,
    i.e. under the benchmark's load) are reported."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.t_mark = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark(self):
        self.t_mark = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if self.t_mark is not None and ts < self.t_mark:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons),
                "window": "timed region + an untimed repeat of the same loop (nvidia-smi -lms 100)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(stage, B, P, F, H, W, R):
    """SURVEY.md section 8d; see DESIGN.md 'Algorithmic bytes'."""
    G = 16 * F + 24 * P + B * (20 * P + 28 * F) + 24 * B * H * W
    grads = 24 * P + 4 * F + B * (12 * P + 28 * F)
    table = {
        "raster_forward": G + 28 * B * H * W + 4 * R,
        "raster_backward": G + 28 * B * H * W + 4 * R + grads,
        "tile_sort": 24 * R,
        "preprocess_pack": G - 24 * B * H * W + 16 * B * F,
        "scatter_instances": 12 * R + 16 * B * F,
        "tile_count": 12 * B * F,
        "classify": 28 * R,
    }
    return table.get(stage)


def make_scene(args, rank):
    from dmesh2_renderer_b200 import scenes
    if args.workload == "cfg4":
        # BASELINE.json configs[3]: cfg2's scene seen from 64 Fibonacci-sphere cameras at 512^2 (same on every rank)
        return scenes.cfg4(n_views=args.views, size=512 if args.size == 1024 else args.size, n_faces=args.faces)
    if args.faces == 1_000_000 and args.size == 1024:
        sc = scenes.cfg2()
    else:
        sc = scenes.cfg2(n_faces=args.faces, size=args.size)
    if rank > 0:      # weak scaling: rank r renders the same scene from Fibonacci-sphere camera r
        import numpy as np
        eye = scenes.fibonacci_eyes(64)[rank % 64]
        sc.mv = np.stack([scenes.look_at(tuple(eye))]).astype("float32")
    return sc


def _oracle_fwd_bwd(sc, view, cores):
    """One whole frame of `sc` (view `view`) forward + backward through the C++ oracle; returns seconds."""
    import torch
    from dmesh2_renderer_b200 import Renderer
    from oracle import oracle as orc
    t = torch.from_numpy
    r = Renderer(t(sc.mv), t(sc.proj), sc.width, sc.height, "cpu")
    a = r.prepare_inputs([view], torch.zeros((1, 2), dtype=torch.int32), sc.width, sc.height, t(sc.verts), t(sc.faces),
                         t(sc.verts_color), t(sc.faces_opacity), t(sc.faces_intense[[view]]), t(sc.background), 1.0)
    names = ["background", "patch_min", "pw", "ph", "verts", "faces", "verts_color", "faces_opacity", "verts_ndc",
             "verts_image", "faces_intense", "tau", "aa_face_verts", "aa_face_edges", "aa_face_edges_iszero",
             "aa_face_edges_recip", "aa_face_edges_normal", "aa_face_edges_normal_c", "K", "image_ray_o", "image_ray_d"]
    d = {n: (v.numpy() if isinstance(v, torch.Tensor) else v) for n, v in zip(names, a)}
    t0 = time.perf_counter()
    orc.render(d, sc.width, sc.height, 1.0, K=0, nthreads=cores, backward=True, dL_dcolor=sc.w_color[[view]],
               dL_ddepth=sc.w_depth[[view]])
    return time.perf_counter() - t0


def cpu_baseline(args, sc):
    """The CPU oracle (kind 'port': the reference has no CPU path), timed for real on all host cores:
    one whole frame (forward + backward) of the bench workload, and - SURVEY 8d - one whole frame of cfg1."""
    from dmesh2_renderer_b200 import scenes
    from oracle import oracle as orc
    cores = orc.hardware_threads()
    dt = _oracle_fwd_bwd(sc, 0, cores)
    dt1 = _oracle_fwd_bwd(scenes.cfg1(), 0, cores)
    unit = "views/s" if args.workload == "cfg4" else "frames/s"
    what = (f"one of the {sc.mv.shape[0]} views of the step ({sc.width}x{sc.height})" if args.workload == "cfg4"
            else f"the whole {sc.width}x{sc.height} frame")
    return {"value": 1.0 / dt, "unit": unit, "cores": cores, "kind": "port",
            "sample": f"C++ oracle forward + backward of {what}, {sc.faces.shape[0]} triangles, measured once ({dt:.2f} s); "
                      "host-side torch prep not included",
            "cfg1": {"value": 1.0 / dt1, "unit": "frames/s",
                     "sample": f"cfg1 (10 k triangles, 256x256) forward + backward, measured ({dt1 * 1e3:.1f} ms)"}}


def main():
    args = parse()
    # exactly ONE line on stdout: libraries (NCCL's version banner, ...) write to fd 1 too, so everything
    # but the final JSON line is routed to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg4 = args.workload == "cfg4"
    unit = "views/s" if cfg4 else "frames/s"
    metric = (f"fwd+bwd views/s, {args.views}-view step @512^2 of 1M tris" if cfg4 else "fwd+bwd frames/s, 1M tris @1024^2")

    if args.impl == "reference" and not torch.cuda.is_available():
        # no GPU: the only runnable stand-in is the CPU oracle port
        if rank == 0:
            sc = make_scene(args, 0)
            cb = cpu_baseline(args, sc)
            emit({"impl": "reference", "metric": metric, "value": cb["value"],
                  "unit": unit, "n_gpus": 0, "steps": 1, "warmup": 0, "higher_is_better": True,
                  "cpu_baseline": cb, "config": {"workload": sc.name},
                  "e2e": {"value": cb["value"], "unit": unit, "h2d_bytes_per_step": 0,
                          "d2h_bytes_per_step": 0}})
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sc = make_scene(args, rank)
    W, H = sc.width, sc.height
    t = lambda a: torch.from_numpy(a)

    if args.impl == "reference":
        ref_dir = os.path.join(ROOT, "baseline", "_ref")
        if not os.path.isdir(os.path.join(ref_dir, "dmesh2_renderer")):
            if rank == 0:
                emit({"impl": "reference", "unavailable": "baseline/_ref not built (baseline/build_ref.sh)"})
            return
        sys.path.insert(0, ref_dir)
        import dmesh2_renderer as pkg          # the UNMODIFIED reference package
        prof = None
    else:
        import dmesh2_renderer_b200 as pkg
        from dmesh2_renderer_b200 import _lib as prof
        pkg.load_library()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                    # nvidia-smi needs a few hundred ms to deliver its first sample

    renderer = pkg.Renderer(t(sc.mv).to(dev), t(sc.proj).to(dev), W, H, dev)
    faces = t(sc.faces).to(dev)
    background = t(sc.background).to(dev)
    wc, wd = t(sc.w_color).to(dev), t(sc.w_depth).to(dev)
    host = {k: t(getattr(sc, k)).pin_memory() for k in ("verts", "verts_color", "faces_opacity", "faces_intense")}
    params = {k: v.to(dev).requires_grad_(True) for k, v in host.items()}
    P, F = sc.verts.shape[0], sc.faces.shape[0]
    n_views = sc.mv.shape[0]
    # views of one step on this rank: cfg2 - the rank's own single view (weak scaling); cfg4 - view i on rank i % N
    my_views = list(range(rank, n_views, world)) if cfg4 else [0]
    batches = [my_views[i:i + args.micro_batch] for i in range(0, len(my_views), args.micro_batch)]
    patch_min = {len(b): torch.zeros((len(b), 2), dtype=torch.int32, device=dev) for b in batches}

    # The flat gradient bucket [dverts | dverts_color | dfaces_opacity].  N > 1: it lives in NVLink peer-mapped
    # memory and is summed by the library's own allreduce kernel (multiview.PeerGradBucket, csrc/dm2_peer.cu);
    # the backward kernels ADD their gradients straight into its views (fused.GradSink): no gradient tensors, no
    # pack copy.  The reference arm packs autograd's .grad tensors and uses NCCL - plain torch, nothing of this repo.
    bucket = sink = None
    use_bucket = world > 1 or cfg4
    if use_bucket:
        if args.impl == "reference":
            class _NcclBucket:
                def __init__(self):
                    self.flat = torch.zeros(3 * P + 3 * P + F, dtype=torch.float32, device=dev)

                def allreduce(self):
                    if world > 1:
                        dist.all_reduce(self.flat)
            bucket = _NcclBucket()
        else:
            from dmesh2_renderer_b200.fused import GradSink
            from dmesh2_renderer_b200.multiview import GradBucket, PeerGradBucket, make_grad_bucket
            if world > 1 and args.exchange != "nccl":
                bucket = make_grad_bucket(P, F, dev)
                if isinstance(bucket, PeerGradBucket):
                    bucket.fused_barriers = args.exchange == "peer"
            else:
                bucket = GradBucket(P, F, dev)
            sink = GradSink(bucket.dverts, bucket.dverts_color, bucket.dfaces_opacity)
            renderer.grad_sink = sink

    # cfg4: the per-micro-batch constants (device index tensor of the views, their loss weights) are built once:
    # indexing with a Python list inside the step costs a pageable host->device copy per use, which drains the
    # queue three times per micro-batch (tools/step_trace.py --cfg4: 6 % GPU idle) - harness work, both arms
    _vbc = {}

    def view_batch_constants(vb):
        k = tuple(vb)
        if k not in _vbc:
            idx = torch.tensor(list(vb), dtype=torch.long, device=dev)
            _vbc[k] = (idx, wc.index_select(0, idx).contiguous(), wd.index_select(0, idx).contiguous())
        return _vbc[k]

    def compute(prm, exchange=True):
        for p in prm.values():
            p.grad = None
        if sink is not None:
            bucket.flat.zero_()
        total = None
        for vb in batches:
            if cfg4:
                idx, wcb, wdb = view_batch_constants(vb)
                fi = prm["faces_intense"].index_select(0, idx)
            else:
                wcb, wdb, fi = wc, wd, prm["faces_intense"]
            color, depth = renderer(vb, patch_min[len(vb)], W, H, prm["verts"], faces, prm["verts_color"],
                                    prm["faces_opacity"], fi, background)
            loss = (color * wcb).sum() + (depth * wdb).sum()
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        if use_bucket:
            if sink is None:
                # reference arm: one flat bucket out of autograd's accumulated .grad tensors
                torch.cat((prm["verts"].grad.reshape(-1), prm["verts_color"].grad.reshape(-1),
                           prm["faces_opacity"].grad.reshape(-1)), out=bucket.flat)
            if exchange:
                bucket.allreduce()        # ONE exchange per optimisation step
        return total

    # cfg4, N > 1: the 64 cameras see the scene from different directions and do not cost the same; view i on rank
    # i % N leaves the exchange waiting ~1 ms of a 13 ms step for the slowest rank at 8 GPUs.  The cameras are fixed
    # for the whole run, so every view is timed ONCE here, before the warm-up (a single-view fwd+bwd frame on the
    # rank that owns it round-robin, CUDA events), the costs are shared, and every rank derives the same ownership
    # table (multiview.balanced_shards: longest first, same number of views per rank).  Both arms.
    view_balance = None
    if cfg4 and world > 1 and not args.no_balance:
        from dmesh2_renderer_b200.multiview import balanced_shards
        costs = torch.zeros(n_views, dtype=torch.float32, device=dev)
        patch_min.setdefault(1, torch.zeros((1, 2), dtype=torch.int32, device=dev))
        for v in my_views:
            batches[:] = [[v]]
            compute(params, exchange=False)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            compute(params, exchange=False)
            compute(params, exchange=False)
            e1.record()
            torch.cuda.synchronize()
            costs[v] = e0.elapsed_time(e1) / 2
        dist.all_reduce(costs)
        c = costs.tolist()
        shards = balanced_shards(c, world)
        rr = [sum(c[v] for v in range(r, n_views, world)) for r in range(world)]
        bal = [sum(c[v] for v in sh) for sh in shards]
        view_balance = {"per_view_ms_min": round(min(c), 4), "per_view_ms_max": round(max(c), 4),
                        "round_robin_rank_ms_max_over_mean": round(max(rr) / (sum(rr) / world), 4),
                        "balanced_rank_ms_max_over_mean": round(max(bal) / (sum(bal) / world), 4)}
        my_views = shards[rank]
        batches[:] = [my_views[i:i + args.micro_batch] for i in range(0, len(my_views), args.micro_batch)]
        for b in batches:
            patch_min.setdefault(len(b), torch.zeros((len(b), 2), dtype=torch.int32, device=dev))

    # e2e input feed (both arms): every step's parameters are copied from pinned host memory inside the
    # timed region, on a copy stream and double buffered - a prefetching loader: the copy for step i+1
    # runs while step i computes; the compute stream waits on the copy's event, the copy waits on the
    # event of the step that last used its buffers.  The loss is read back to the host every step.
    copy_stream = torch.cuda.Stream(device=dev)
    loss_host = torch.zeros(max(args.steps, 8) + 8, dtype=torch.float32).pin_memory()   # one read-back slot per step
    sets = [params, {k: torch.empty_like(v).requires_grad_(True) for k, v in params.items()}]
    ready, free = [None, None], [None, None]

    def prefetch(j):
        with torch.cuda.stream(copy_stream), torch.no_grad():
            if free[j] is not None:
                copy_stream.wait_event(free[j])
            for k in sets[j]:
                sets[j][k].copy_(host[k], non_blocking=True)
            ready[j] = torch.cuda.Event()
            ready[j].record(copy_stream)

    def step_e2e(i, last):
        j = i & 1
        if ready[j] is None:
            prefetch(j)
        torch.cuda.current_stream(dev).wait_event(ready[j])
        ready[j] = None
        if not last:
            prefetch(j ^ 1)
        loss = compute(sets[j])
        free[j] = torch.cuda.Event()
        free[j].record(torch.cuda.current_stream(dev))
        # device -> host read of the step's result: asynchronous copy into pinned memory (checked after the
        # timed region), so that logging the loss does not drain the launch queue every step
        if os.environ.get("BENCH_SYNC_LOSS"):
            return float(loss.item())
        loss_host[i % loss_host.numel()].copy_(loss, non_blocking=True)
        return None

    def timed(n, e2e):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            if e2e:
                step_e2e(i, i == n - 1)
            else:
                compute(params)
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        compute(params)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats(dev)
    if prof:
        prof.profile_reset()
        prof.profile_enable(True)
        prof.launch_count(reset=True)
    sampler.mark()
    ms_total = timed(args.steps, False)
    launches = prof.launch_count() if prof else None
    stages = prof.profile_report() if prof else {}
    if prof:
        prof.profile_enable(False)
    peak_mem = torch.cuda.max_memory_allocated(dev)
    for _ in range(2):
        step_e2e(0, True)
    ms_e2e = timed(args.steps, True)
    assert os.environ.get("BENCH_SYNC_LOSS") or (bool(torch.isfinite(loss_host[:args.steps]).all()) and
                                                 float(loss_host[:args.steps].abs().max()) > 0.0), "e2e: loss read-back missing"
    # untimed repeat of the timed loop: gives nvidia-smi (100 ms period) enough samples under the benchmark's load
    t_end = time.time() + args.clock_window_s
    while time.time() < t_end:
        for _ in range(4):
            compute(params)
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # N > 1: the exchange is verified where the driver can see it - the bucket summed by this library's kernel
    # against torch.distributed.all_reduce (NCCL) of the SAME per-rank buckets, and bit-identical on all ranks
    allreduce_check = None
    if world > 1 and args.impl != "reference":
        compute(params, exchange=False)
        ref = bucket.flat.clone()
        dist.all_reduce(ref)
        bucket.allreduce()
        torch.cuda.synchronize()
        err = ((bucket.flat - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).reshape(1)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        chk = bucket.flat.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        allreduce_check = {"max_rel_err_vs_nccl": float(err.item()), "same_on_all_ranks": bool((lo == hi).item()),
                           "bucket_bytes": int(bucket.flat.numel() * 4), "nonzero": bool(ref.abs().max().item() > 0)}
        ok = allreduce_check["max_rel_err_vs_nccl"] <= 1e-6 and allreduce_check["same_on_all_ranks"] and allreduce_check["nonzero"]
        if not ok:
            if rank == 0:
                print(f"bench.py: the peer-memory allreduce disagrees with NCCL: {allreduce_check}", file=sys.stderr)
            dist.destroy_process_group()
            raise SystemExit(1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    views_per_step = n_views if cfg4 else world
    value = views_per_step * 1000.0 / ms_step
    exchange = "NCCL allreduce"
    if bucket is not None and hasattr(bucket, "uses_multicast"):
        exchange = ("peer-memory allreduce kernel (" + ("NVLS multimem" if bucket.uses_multicast else "P2P ld/st")
                    + (", in-kernel barriers)" if bucket.fused_barriers else ", 2 barrier launches)"))

    h2d = sum(v.numel() * v.element_size() for v in host.values())
    if cfg4:
        workload = (f"{sc.name}: DMesh++ multi-view step, {n_views} views at {W}x{H} of {F} random semi-transparent triangles "
                    f"(SURVEY 8d cfg4: cfg2's scene, Fibonacci-sphere cameras), {args.micro_batch} views per Renderer call, "
                    + (f"views assigned to the ranks by measured per-view cost (multiview.balanced_shards)"
                       if view_balance is not None else f"view i on rank i % {world}"))
    else:
        workload = (f"{sc.name}: Renderer fwd+bwd, {F} random semi-transparent triangles, {W}x{H}, "
                    f"1 view per GPU (SURVEY 8d cfg2, seed 2, sigma 0.0115, aa_temperature 1.0)")
    out = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if cfg4 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload,
                   "views_per_step": views_per_step,
                   "parallelism": f"view-sharded dp{world}, 1 {exchange}/step" if world > 1 else "single GPU",
                   "l2": "working set (packed face records 256 MB + inputs ~190 MB per view) exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": views_per_step * 1000.0 * args.steps / ms_e2e, "unit": unit, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "clocks": clocks,
        "peak_mem_gb": peak_mem / 2 ** 30,
    }
    if allreduce_check is not None:
        out["allreduce_check"] = allreduce_check
    if view_balance is not None:
        out["view_balance"] = view_balance
    if args.impl == "reference":
        out["impl"] = "reference"
        out["cpu_baseline"] = {"value": value, "unit": unit, "cores": 0, "kind": "reference",
                               "sample": "the reference has no CPU path: this arm runs its own CUDA extension "
                                         "(baseline/_ref, unmodified, built for sm_100a) through its public "
                                         "Renderer API on the same GPU, same scene, same harness"}
        out["gpu_launches"] = None
    else:
        out["gpu_launches"] = launches
        R = None
        try:
            R = int(renderer_num_rendered(pkg, renderer, W, H, params, faces, background, batches[0], cfg4))
        except Exception:
            pass
        peak, peak_src = peaks()
        if stages and R is not None:
            B = len(batches[0])
            calls = len(batches)                    # Renderer calls per step on this rank
            per_call = {k: v[0] / max(v[1], 1) for k, v in stages.items()}
            dom = max(stages, key=lambda k: stages[k][0])
            dom_ms = per_call[dom]
            ab = algorithmic_bytes(dom, B, P, F, H, W, R)
            achieved = ab / (dom_ms * 1e-3) / 1e9 if ab else None
            # DRAM traffic + warp instructions of that kernel per launch: from the committed ncu --set full capture of
            # THIS command (profiles/), not measured in this run - labelled as such; only for the default cfg2 workload
            traffic = issue_frac = traffic_src = None
            tj = next((q for q in (os.path.join(ROOT, "profiles", n) for n in ("r02_dram_traffic.json", "r01_dram_traffic.json"))
                       if os.path.exists(q)), None)
            if tj and not cfg4 and args.faces == 1_000_000 and args.size == 1024:
                with open(tj) as fh:
                    prof_json = json.load(fh)
                key = next((k for k in prof_json["bytes_per_launch"] if k.startswith("k_" + dom)), None)
                traffic = prof_json["bytes_per_launch"].get(key)
                traffic_src = f"profiles/{os.path.basename(tj)} (committed ncu --set full capture of this command; not measured in this run)"
                # the limiter that actually binds (not a roofline the contract names, reported beside it): warp
                # instructions of that kernel (same ncu capture) / live kernel time, against the SMs' issue rate
                # (4 schedulers per SM, one warp instruction per clock each, at the sampled SM clock)
                wi = prof_json.get("warp_instructions_per_launch", {}).get(key)
                props = torch.cuda.get_device_properties(dev)
                mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz")
                if wi and mhz:
                    issue_frac = wi / (dom_ms * 1e-3) / (props.multi_processor_count * 4 * mhz * 1e6)
            # the whole step against the same roofline: A_fwd + A_bwd of SURVEY 8d over the measured step time
            a_step = calls * (algorithmic_bytes("raster_forward", B, P, F, H, W, R) + 24 * R +
                              algorithmic_bytes("raster_backward", B, P, F, H, W, R))
            out["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                               "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                               "traffic_source": traffic_src,
                               "peak_source": peak_src, "algorithmic_bytes": ab, "kernel_ms": dom_ms,
                               "step_algorithmic_bytes": a_step,
                               "step_frac": a_step / (ms_step * 1e-3) / 1e9 / peak,
                               "issue_slot_frac": issue_frac,
                               "note": "instruction/latency bound, not HBM bound: hundreds of fp32 instructions per "
                                       "covered (pixel, face) pair (profiles/r02_ncu_kernels.md, DESIGN.md section 3); "
                                       "step_frac = (A_fwd + A_bwd of SURVEY 8d) / ms_per_step / peak"}
            out["stages_ms_per_step"] = {k: v[0] / args.steps for k, v in stages.items()}
            out["num_rendered"] = R
            # the binning sort = scatter into the tile segments + per-segment sort (dm2_tilesort.cu)
            sort_ms = (per_call.get("scatter_instances", 0) + per_call.get("tile_sort", 0))
            if sort_ms > 0:
                out["sort_keys_per_s"] = R / (sort_ms * 1e-3)
        if not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline(args, make_scene(args, 0))
            except Exception as ex:
                out["cpu_baseline"] = {"value": None, "unit": unit, "cores": None, "kind": "port",
                                       "sample": f"failed: {ex}"}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def renderer_num_rendered(pkg, renderer, W, H, params, faces, background, views, cfg4):
    """num_rendered of one Renderer call of the step (reference-shaped route: its 10-tuple carries it)."""
    import torch
    with torch.no_grad():
        pm = torch.zeros((len(views), 2), dtype=torch.int32, device=faces.device)
        fi = params["faces_intense"][views] if cfg4 else params["faces_intense"]
        a = renderer.prepare_inputs(views, pm, W, H, params["verts"], faces, params["verts_color"],
                                    params["faces_opacity"], fi, background, 1.0)
        return pkg._C.render_forward_cuda(*a)[0]


if __name__ == "__main__":
    main()
