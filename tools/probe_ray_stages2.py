#!/usr/bin/env python
"""World-stage (K = 4 batched matmul) accumulation order of torch on this box, over frame sizes / view counts."""
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmesh2_renderer_b200 import scenes  # noqa: E402

f32, f64 = torch.float32, torch.float64


def fma(a, b, c):
    return (a.to(f64) * b.to(f64) + c.to(f64)).to(f32)


def main():
    dev = torch.device("cuda:0")
    for size, nv in ((64, 3), (96, 3), (256, 2), (512, 2), (512, 8), (512, 64), (1024, 4), (2048, 2)):
        sc = scenes.cfg4(n_views=nv, size=size, n_faces=10)
        mv = torch.from_numpy(sc.mv).to(dev)
        inv_mv = torch.inverse(mv)
        g = torch.Generator(device=dev).manual_seed(1)
        view = torch.randn((nv, size, size, 1, 4), device=dev, generator=g)
        world = torch.matmul(view, inv_mv.transpose(1, 2).unsqueeze(1).unsqueeze(1)).squeeze(-2)
        xv = view.squeeze(-2)
        prods = lambda j, k: (xv[..., k], inv_mv[:, j, :].reshape(nv, 1, 1, 4)[..., k])
        found = []
        for order in itertools.permutations(range(4)):
            outs_chain, outs_nofma, outs_fma0 = [], [], []
            for j in range(4):
                a, b = prods(j, order[0])
                acc = a * b
                acc2 = a * b
                for k in order[1:]:
                    a, b = prods(j, k)
                    acc = fma(a, b, acc)
                    acc2 = acc2 + a * b
                outs_chain.append(acc); outs_nofma.append(acc2)
            if torch.equal(torch.stack(outs_chain, -1), world): found.append("chain" + "".join(map(str, order)))
            if torch.equal(torch.stack(outs_nofma, -1), world): found.append("nofma" + "".join(map(str, order)))
        for (i, j2), (k, l) in (((0, 1), (2, 3)), ((0, 2), (1, 3)), ((0, 3), (1, 2))):
            for o1 in ((i, j2), (j2, i)):
                for o2 in ((k, l), (l, k)):
                    for swap in (False, True):
                        outs = []
                        for j in range(4):
                            a, b = prods(j, o1[0]); a1, b1 = prods(j, o1[1])
                            c, d = prods(j, o2[0]); c1, d1 = prods(j, o2[1])
                            p = fma(a1, b1, a * b); q = fma(c1, d1, c * d)
                            outs.append(p + q)
                        if not swap and torch.equal(torch.stack(outs, -1), world):
                            found.append(f"pair({o1[0]}{o1[1]})({o2[0]}{o2[1]})")
        P = []
        for j in range(4):
            P.append([xv[..., k] * inv_mv[:, j, :].reshape(nv, 1, 1, 4)[..., k] for k in range(4)])
        for (i, j2), (k, l) in (((0, 1), (2, 3)), ((0, 2), (1, 3)), ((0, 3), (1, 2))):
            outs = [(P[j][i] + P[j][j2]) + (P[j][k] + P[j][l]) for j in range(4)]
            if torch.equal(torch.stack(outs, -1), world): found.append(f"nofma_pair({i}{j2})({k}{l})")
        # fp64 accumulate, single rounding
        outs = [sum(xv[..., k].to(f64) * inv_mv[:, j, :].reshape(nv, 1, 1, 4)[..., k].to(f64) for k in range(4)).to(f32) for j in range(4)]
        if torch.equal(torch.stack(outs, -1), world): found.append("f64_exact")
        d = (torch.stack(outs, -1) - world).abs().max().item()
        print(f"size {size} views {nv}: max |f64 - torch| {d:.3g}; {found[:6]}{' ...' if len(found) > 6 else ''} ({len(found)} match)")


if __name__ == "__main__":
    main()
