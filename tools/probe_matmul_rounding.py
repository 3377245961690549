#!/usr/bin/env python
"""Which accumulation order do torch's two clip-space matmuls use on this GPU?  (experiment behind the
decision whether the projection can move into the preprocess kernel bit-exactly).

For K = 4 every candidate formula is emulated in float64 (products of two fp32 numbers are exact in
fp64; each fused multiply-add is rounded to fp32 once) and compared bit-for-bit with torch.matmul."""
import itertools
import sys

import torch

sys.path.insert(0, ".")
from dmesh2_renderer_b200 import scenes        # noqa: E402


def fma32(a, b, c):
    return (a.double() * b.double() + c.double()).float()


def chain(A, Bm, order):
    """A [N,4], Bm [4,4] -> [N,4] with fma accumulation over k in `order`, first term a plain product."""
    out = []
    for j in range(4):
        acc = (A[:, order[0]].double() * Bm[order[0], j].double()).float()
        for k in order[1:]:
            acc = fma32(A[:, k], Bm[k, j].expand_as(acc), acc)
        out.append(acc)
    return torch.stack(out, dim=1)


def main():
    dev = torch.device("cuda:0")
    sc = scenes.cfg2(n_faces=200_000)
    v = torch.from_numpy(sc.verts).to(dev)
    mv = torch.from_numpy(sc.mv).to(dev)
    proj = torch.from_numpy(sc.proj).to(dev)
    hom = torch.cat((v, torch.ones_like(v[:, [0]])), dim=-1)
    view = torch.matmul(hom, mv.transpose(1, 2))            # [1, P, 4]
    clip = torch.matmul(view, proj.transpose(1, 2))          # [1, P, 4]
    for name, A, M, ref in (("hom @ mv^T", hom, mv[0].t(), view[0]), ("view @ proj^T", view[0], proj[0].t(), clip[0])):
        best = None
        for order in itertools.permutations(range(4)):
            e = chain(A, M.contiguous(), order)
            mism = float((e.view(torch.int32) != ref.view(torch.int32)).float().mean())
            if best is None or mism < best[0]:
                best = (mism, order)
            if mism == 0.0:
                break
        print(f"{name}: best order {best[1]} mismatch fraction {best[0]:.3e}")
    # 8-view batch (bmm path)
    sc8 = scenes.cfg4(n_views=8, n_faces=100_000)
    v = torch.from_numpy(sc8.verts).to(dev)
    mv = torch.from_numpy(sc8.mv).to(dev)
    proj = torch.from_numpy(sc8.proj).to(dev)
    hom = torch.cat((v, torch.ones_like(v[:, [0]])), dim=-1)
    view = torch.matmul(hom, mv.transpose(1, 2))
    clip = torch.matmul(view, proj.transpose(1, 2))
    for b in (0, 5):
        for name, A, M, ref in (("B=8 hom @ mv^T", hom, mv[b].t(), view[b]), ("B=8 view @ proj^T", view[b], proj[b].t(), clip[b])):
            best = None
            for order in itertools.permutations(range(4)):
                e = chain(A, M.contiguous(), order)
                mism = float((e.view(torch.int32) != ref.view(torch.int32)).float().mean())
                if best is None or mism < best[0]:
                    best = (mism, order)
                if mism == 0.0:
                    break
            print(f"{name} [view {b}]: best order {best[1]} mismatch fraction {best[0]:.3e}")


if __name__ == "__main__":
    main()
