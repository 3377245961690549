import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from dmesh2_renderer_b200 import scenes
dev = torch.device("cuda:0")
f32, f64 = torch.float32, torch.float64
def fma(a, b, c): return (a.to(f64) * b.to(f64) + c.to(f64)).to(f32)
for size, nv in ((64, 3), (96, 3), (100, 1), (200, 2), (256, 2), (512, 2), (512, 8), (1000, 1), (1024, 1), (1024, 4), (2048, 1)):
    sc = scenes.cfg4(n_views=max(nv, 2), size=size, n_faces=10)
    inv_mv = torch.inverse(torch.from_numpy(sc.mv).to(dev))[-nv:]
    g = torch.Generator(device=dev).manual_seed(1)
    view = torch.randn((nv, size, size, 1, 4), device=dev, generator=g)
    world = torch.matmul(view, inv_mv.transpose(1, 2).unsqueeze(1).unsqueeze(1)).squeeze(-2).reshape(-1, 4)
    xv = view.reshape(nv, -1, 4)
    ch, pr = [], []
    for j in range(4):
        mj = inv_mv[:, j, :].reshape(nv, 1, 4)
        acc = xv[..., 0] * mj[..., 0]
        for k in (1, 2, 3):
            acc = fma(xv[..., k], mj[..., k], acc)
        ch.append(acc)
        P = [xv[..., k] * mj[..., k] for k in range(4)]
        pr.append((P[0] + P[2]) + (P[1] + P[3]))
    ch = torch.stack(ch, -1).reshape(-1, 4); pr = torch.stack(pr, -1).reshape(-1, 4)
    is_ch = (ch == world).all(-1); is_pr = (pr == world).all(-1)
    only_ch = torch.nonzero(is_ch & ~is_pr).flatten(); only_pr = torch.nonzero(is_pr & ~is_ch).flatten()
    neither = int((~is_ch & ~is_pr).sum())
    M = world.shape[0]
    print(f"size {size} views {nv} M {M}: rows only-chain: {only_ch.numel()} (max {int(only_ch.max()) if only_ch.numel() else -1}), "
          f"only-pair: {only_pr.numel()} (min {int(only_pr.min()) if only_pr.numel() else -1} = M - {M - int(only_pr.min()) if only_pr.numel() else 0}), neither {neither}")
