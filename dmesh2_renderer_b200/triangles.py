"""Screen-space triangle set-up for the anti-aliasing path.

Same values as ``Triangles`` in the reference (dmesh2_renderer/pyrenderer.py:6-30, with
``order_ccw`` :521-529 and ``tri_area`` :531-535): counter-clockwise re-ordering, the three edge
vectors, their |e| < 1e-3 flags, reciprocals (inf allowed) and inward edge normals n = (-e.y, e.x)
with offsets c = n . p.  Every intermediate is a separate fp32 torch op in the reference's order,
so the results are bit-identical; the only change is that the CCW swap is a ``torch.where`` instead
of two boolean-mask index assignments (which force a device->host sync for ``nonzero``).
"""
import torch

EDGE_EPS = 1e-3      # reference pyrenderer.py:4


def signed_area(p0, p1, p2):
    ax = p1[:, 0] - p0[:, 0]
    by = p2[:, 1] - p0[:, 1]
    cx = p2[:, 0] - p0[:, 0]
    dy = p1[:, 1] - p0[:, 1]
    return 0.5 * (ax * by - cx * dy)


def _inward_normal(edge, anchor):
    # reference: e[:, [1, 0]] with column 0 negated, then (n * p).sum(dim=1)
    n = torch.stack((-edge[:, 1], edge[:, 0]), dim=1)
    prod = n * anchor
    return n, prod[:, 0] + prod[:, 1]


class Triangles:
    def __init__(self, p0, p1, p2):
        flip = (signed_area(p0, p1, p2) < 0).unsqueeze(1)
        q1 = torch.where(flip, p2, p1)
        q2 = torch.where(flip, p1, p2)
        p1, p2 = q1, q2
        e0, e1, e2 = p1 - p0, p2 - p1, p0 - p2
        self.verts = torch.stack((p0, p1, p2), dim=1)                  # [T, 3, 2]
        self.edges = torch.stack((e0, e1, e2), dim=1)                  # [T, 3, 2]
        self.edges_iszero = self.edges.abs() < EDGE_EPS                # [T, 3, 2] bool
        self.edges_recip = 1.0 / self.edges                            # [T, 3, 2]
        n0, c0 = _inward_normal(e0, p0)
        n1, c1 = _inward_normal(e1, p1)
        n2, c2 = _inward_normal(e2, p2)
        self.edges_normal = torch.stack((n0, n1, n2), dim=1)           # [T, 3, 2]
        self.edges_normal_c = torch.stack((c0, c1, c2), dim=1)         # [T, 3]
