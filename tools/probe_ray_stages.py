#!/usr/bin/env python
"""Stage-by-stage probe of the rounding of torch's ops in Renderer._init_rays on this box (dev tool for the ray
recipes of dm2_generate_rays): every stage's candidates are evaluated on torch's OWN previous-stage tensors."""
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


def mul(a, b):
    return (a * b)


def main():
    dev = torch.device("cuda:0")
    for size, nv in ((96, 3), (512, 8), (1024, 1)):
        sc = scenes.small(n_faces=10, size=size, n_views=nv)
        mv, proj = torch.from_numpy(sc.mv).to(dev), torch.from_numpy(sc.proj).to(dev)
        inv_mv, inv_proj = torch.inverse(mv), torch.inverse(proj)
        W = H = size
        xs = torch.arange(W, device=dev).float(); ys = torch.arange(H, device=dev).float()
        centre = torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1).float() + 0.5
        centre = centre.transpose(0, 1).unsqueeze(0).expand(nv, -1, -1, -1)
        ndc = torch.zeros_like(centre)
        ndc[..., 0] = (centre[..., 0] / W * 2) - 1
        ndc[..., 1] = (centre[..., 1] / H * 2) - 1
        c0 = centre[..., 0].contiguous()
        cand = {"div": (c0 / torch.tensor(float(W), device=dev)) * 2 - 1,
                "mulrecip": (c0 * torch.tensor(1.0 / W, dtype=f32, device=dev)) * 2 - 1,
                "mulrecip64": (c0.to(f64) * (1.0 / W)).to(f32) * 2 - 1}
        print(f"size {size} views {nv}: ndc.x ==", {k: bool(torch.equal(v, ndc[..., 0])) for k, v in cand.items()})
        near = -torch.ones_like(ndc[..., 0:1])
        clip = torch.cat((ndc, near), dim=-1)
        clip = torch.cat((clip, torch.ones_like(clip[..., 0:1])), dim=-1).unsqueeze(-2)
        view = torch.matmul(clip, inv_proj.transpose(1, 2).unsqueeze(1).unsqueeze(1))
        world = torch.matmul(view, inv_mv.transpose(1, 2).unsqueeze(1).unsqueeze(1))
        for name, x, m, y in (("view", clip, inv_proj, view), ("world", view, inv_mv, world)):
            xv = x.squeeze(-2)                                   # [V,H,W,4]
            res = {}
            for order in itertools.permutations(range(4)):
                # fma chain in this order, first term a plain product
                outs = []
                for j in range(4):
                    mj = m[:, j, :].reshape(nv, 1, 1, 4)
                    acc = xv[..., order[0]] * mj[..., order[0]]
                    for k in order[1:]:
                        acc = fma(xv[..., k], mj[..., k], acc)
                    outs.append(acc)
                res["chain" + "".join(map(str, order))] = bool(torch.equal(torch.stack(outs, -1), y.squeeze(-2)))
            # pairwise
            outs = []
            for j in range(4):
                mj = m[:, j, :].reshape(nv, 1, 1, 4)
                a = fma(xv[..., 1], mj[..., 1], xv[..., 0] * mj[..., 0]); b = fma(xv[..., 3], mj[..., 3], xv[..., 2] * mj[..., 2])
                outs.append(a + b)
            res["pair"] = bool(torch.equal(torch.stack(outs, -1), y.squeeze(-2)))
            # no fma at all, left to right
            outs = []
            for j in range(4):
                mj = m[:, j, :].reshape(nv, 1, 1, 4)
                acc = xv[..., 0] * mj[..., 0]
                for k in (1, 2, 3):
                    acc = acc + xv[..., k] * mj[..., k]
                outs.append(acc)
            res["nofma0123"] = bool(torch.equal(torch.stack(outs, -1), y.squeeze(-2)))
            print("  ", name, "matches:", [k for k, v in res.items() if v])
        origin = torch.zeros((nv, H, W, 3), device=dev)
        for k in range(3):
            origin[..., k] = inv_mv[:, k, 3].reshape(-1, 1, 1)
        target = world[..., :3].squeeze(-2)
        direction = target - origin
        nrm = torch.norm(direction, dim=-1, keepdim=True)
        x, y, z = direction[..., 0:1], direction[..., 1:2], direction[..., 2:3]
        ncand = {
            "fma_xyz": torch.sqrt(fma(z, z, fma(y, y, x * x))),
            "(x2+y2)+z2": torch.sqrt((x * x + y * y) + z * z),
            "x2+(y2+z2)": torch.sqrt(x * x + (y * y + z * z)),
            "(x2+z2)+y2": torch.sqrt((x * x + z * z) + y * y),
            "fma_zyx": torch.sqrt(fma(x, x, fma(y, y, z * z))),
            "f64": torch.sqrt((x.to(f64) ** 2 + y.to(f64) ** 2 + z.to(f64) ** 2)).to(f32),
            "f64sum_f32sqrt": torch.sqrt((x.to(f64) ** 2 + y.to(f64) ** 2 + z.to(f64) ** 2).to(f32)),
            "scaled": None,
        }
        mx = direction.abs().amax(-1, keepdim=True)
        ncand["scaled"] = mx * torch.sqrt(((x / mx) ** 2 + (y / mx) ** 2) + (z / mx) ** 2)
        print("   norm matches:", [k for k, v in ncand.items() if torch.equal(v, nrm)],
              {k: float((v - nrm).abs().max()) for k, v in ncand.items()})
        length = nrm + 1e-6
        rd = direction / length
        print("   div matches:", bool(torch.equal(rd, direction / length)), "mulrecip:", bool(torch.equal(rd, direction * (1.0 / length))))


if __name__ == "__main__":
    main()
