"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the interpenetration term's third-party
package ``mesh_intersection``.

**Parity unpinned.**  The reference imports ``mesh_intersection`` (fork xiyichen/torch-mesh-isect
of vchoutas/torch-mesh-isect; installed from git HEAD, no version pin: reference README.md:38)
at fit_single_frame.py:301-303; the package is not vendored, no reference test or fixture holds
its outputs, and it is a CUDA-only extension that cannot be built here.  This file restates the
*published* algorithm (Pavlakos et al., "Expressive Body Capture", CVPR 2019, sec. 3.4 + supp.;
Tzionas et al., IJCV 2016, eq. 12-15; Karras 2012 for the BVH whose *result set* -- not its
traversal order -- is reproduced) behind the three classes the reference constructs and calls:

* ``BVH(max_collisions)``                       fit_single_frame.py:305, called fitting.py:446
* ``FilterFaces(faces_segm, faces_parents, ign_part_pairs)``   fit_single_frame.py:325, fitting.py:450
* ``DistanceFieldPenetrationLoss(sigma, point2plane, vectorized, penalize_outside)``
                                                 fit_single_frame.py:308, fitting.py:454

so the reference's own, unmodified ``SMPLifyLoss`` (fitting.py:375-461) can run with them.

What is restated:

search tree   every unordered pair of triangles whose axis-aligned boxes overlap, that share no
              vertex (compared by coordinates, as the package does) and that intersect by the
              17-axis separating-axis test (2 normals, 9 edge x edge, 6 in-plane edge normals).
              Output ``[B, F * max_collisions, 2]`` int64, rows (lower face id, higher face id)
              in lexicographic order, padded with -1.  The package's own order comes from an
              atomic counter and is not deterministic; the penalty is a sum, so only the set
              matters.  A triangle with more than ``max_collisions`` box-overlapping partners
              is truncated by the package; the shipped configurations use 128 so that this
              does not happen, and neither this port nor the engine truncates per triangle.
filter        drops pairs of the same part, of parent / child parts and of ``ign_part_pairs``.
penalty       for a pair (f_s, f_t):  sum_{v in f_t} Psi_{f_s}(v)^2 + sum_{v in f_s} Psi_{f_t}(v)^2
              Psi_f(v)   = ((1 - Phi(v)) * Upsilon(n_f.(v - o_f)))^2      if Phi(v) < 1 else 0
              Phi(v)     = ||(v - o_f) - (n_f.(v - o_f)) n_f|| / (-(r_f / sigma) n_f.(v - o_f) + r_f)
              Upsilon(x) = -x + 1 - sigma                                   x <= -sigma
                           -(1 - 2 sigma)/(4 sigma^2) x^2 - x/(2 sigma) + (3 - 2 sigma)/4   |x| < sigma
                           0                                                x >= sigma
              o_f, r_f the circumcentre / circumradius of f, n_f its unit normal.
"""
import numpy as np
import torch


# ------------------------------------------------------------------------------ search tree
def _sat_axes(t1, t2):
    """t1, t2: [N,3,3] -> [N,17,3] separating-axis candidates."""
    e1 = np.stack([t1[:, 1] - t1[:, 0], t1[:, 2] - t1[:, 0], t1[:, 2] - t1[:, 1]], 1)
    e2 = np.stack([t2[:, 1] - t2[:, 0], t2[:, 2] - t2[:, 0], t2[:, 2] - t2[:, 1]], 1)
    n1 = np.cross(e1[:, 0], e1[:, 1])
    n2 = np.cross(e2[:, 0], e2[:, 1])
    axes = [n1, n2]
    for i in range(3):
        for j in range(3):
            axes.append(np.cross(e1[:, i], e2[:, j]))
    for i in range(3):
        axes.append(np.cross(n1, e1[:, i]))
    for j in range(3):
        axes.append(np.cross(n2, e2[:, j]))
    return np.stack(axes, 1)


def triangles_intersect(t1, t2):
    """Separating-axis test of triangle pairs ([N,3,3] each) -> bool [N]."""
    ax = _sat_axes(t1, t2)                                   # [N,17,3]
    p1 = np.einsum('nad,nvd->nav', ax, t1)                   # [N,17,3]
    p2 = np.einsum('nad,nvd->nav', ax, t2)
    sep = (p1.max(-1) < p2.min(-1)) | (p2.max(-1) < p1.min(-1))
    return ~sep.any(-1)


def share_vertex(t1, t2):
    eq = (t1[:, :, None, :] == t2[:, None, :, :]).all(-1)    # [N,3,3]
    return eq.any((-1, -2))


def box_overlap_pairs(lo, hi, chunk=2048):
    """All i < j with overlapping boxes ([F,3] lo / hi): sweep along x, filter y / z."""
    F = lo.shape[0]
    order = np.argsort(lo[:, 0], kind='stable')
    slo, shi = lo[order], hi[order]
    # for sorted position a, partners b > a with slo[b].x <= shi[a].x
    end = np.searchsorted(slo[:, 0], shi[:, 0], side='right')
    out = []
    for a0 in range(0, F, chunk):
        a1 = min(F, a0 + chunk)
        cnt = np.maximum(end[a0:a1] - (np.arange(a0, a1) + 1), 0)
        if cnt.sum() == 0:
            continue
        ai = np.repeat(np.arange(a0, a1), cnt)
        start = np.repeat(np.arange(a0, a1) + 1, cnt)
        off = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        bi = start + off
        ok = ((slo[ai, 1] <= shi[bi, 1]) & (slo[bi, 1] <= shi[ai, 1]) &
              (slo[ai, 2] <= shi[bi, 2]) & (slo[bi, 2] <= shi[ai, 2]))
        out.append(np.stack([order[ai[ok]], order[bi[ok]]], 1))
    if not out:
        return np.zeros((0, 2), np.int64)
    pairs = np.concatenate(out, 0)
    pairs = np.stack([pairs.min(1), pairs.max(1)], 1)
    return pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]


def find_collisions(tri):
    """tri [F,3,3] (numpy) -> int64 [n,2] intersecting pairs, lower id first, lexicographic."""
    lo, hi = tri.min(1), tri.max(1)
    pairs = box_overlap_pairs(lo, hi)
    if len(pairs) == 0:
        return pairs
    t1, t2 = tri[pairs[:, 0]], tri[pairs[:, 1]]
    keep = ~share_vertex(t1, t2)
    pairs, t1, t2 = pairs[keep], t1[keep], t2[keep]
    return pairs[triangles_intersect(t1, t2)]


class BVH(object):
    """mesh_intersection.bvh_search_tree.BVH: ``triangles [B,F,3,3] -> collision_idxs
    [B, F * max_collisions, 2]`` (int64, -1 padded, no gradient)."""

    def __init__(self, max_collisions=8):
        self.max_collisions = max_collisions

    def __call__(self, triangles):
        B, F = triangles.shape[:2]
        out = torch.full((B, F * self.max_collisions, 2), -1, dtype=torch.int64)
        t = triangles.detach().cpu().numpy()
        for b in range(B):
            p = find_collisions(t[b])[:F * self.max_collisions]
            out[b, :len(p)] = torch.from_numpy(p)
        return out.to(triangles.device)


# ------------------------------------------------------------------------------ filter
def allowed_part_matrix(faces_segm, faces_parents, ign_part_pairs):
    """[P,P] bool: may a pair of faces of parts (p, q) collide?  (Same part, parent / child
    and ignored pairs may not.)"""
    segm = np.asarray(faces_segm).astype(np.int64)
    par = np.asarray(faces_parents).astype(np.int64)
    P = int(segm.max()) + 1
    parent_of = np.full(P, -1, np.int64)
    parent_of[segm] = par
    ok = ~np.eye(P, dtype=bool)
    for p in range(P):
        if parent_of[p] >= 0:
            ok[p, parent_of[p]] = ok[parent_of[p], p] = False
    for pair in (ign_part_pairs or []):
        a, b = (int(x) for x in (pair.split(',') if isinstance(pair, str) else pair))
        if a < P and b < P:
            ok[a, b] = ok[b, a] = False
    return ok


class FilterFaces(torch.nn.Module):
    """mesh_intersection.filter_faces.FilterFaces."""

    def __init__(self, faces_segm=None, faces_parents=None, ign_part_pairs=None):
        super(FilterFaces, self).__init__()
        self.register_buffer('faces_segm', torch.as_tensor(np.asarray(faces_segm), dtype=torch.long))
        self.register_buffer('faces_parents',
                             torch.as_tensor(np.asarray(faces_parents), dtype=torch.long))
        ign = [[int(x) for x in (p.split(',') if isinstance(p, str) else p)]
               for p in (ign_part_pairs or [])]
        self.register_buffer('ign_pairs', torch.as_tensor(ign, dtype=torch.long).reshape(-1, 2))

    def forward(self, collision_idxs):
        recv, intr = collision_idxs[..., 0], collision_idxs[..., 1]
        valid = recv >= 0
        rs, is_ = self.faces_segm[recv.clamp(min=0)], self.faces_segm[intr.clamp(min=0)]
        rp, ip = self.faces_parents[recv.clamp(min=0)], self.faces_parents[intr.clamp(min=0)]
        keep = valid & (rs != is_) & (rp != is_) & (ip != rs)
        for a, b in self.ign_pairs.tolist():
            keep &= ~(((rs == a) & (is_ == b)) | ((rs == b) & (is_ == a)))
        return torch.where(keep[..., None], collision_idxs, torch.full_like(collision_idxs, -1))


# ------------------------------------------------------------------------------ penalty
def circumcircle(tri):
    """tri [...,3,3] -> centre [...,3], radius [...], unit normal [...,3]."""
    a = tri[..., 0, :] - tri[..., 2, :]
    b = tri[..., 1, :] - tri[..., 2, :]
    cr = torch.cross(a, b, dim=-1)
    cc = (cr * cr).sum(-1)
    aa, bb = (a * a).sum(-1), (b * b).sum(-1)
    ab = a - b
    radius = torch.sqrt((ab * ab).sum(-1) * aa * bb) / (2 * torch.sqrt(cc))
    w = aa[..., None] * b - bb[..., None] * a
    centre = tri[..., 2, :] + torch.cross(w, cr, dim=-1) / (2 * cc[..., None])
    normal = cr / torch.sqrt(cc)[..., None]
    return centre, radius, normal


def upsilon(x, sigma):
    lin = -x + 1 - sigma
    quad = -(1 - 2 * sigma) / (4 * sigma * sigma) * x * x - x / (2 * sigma) + (3 - 2 * sigma) / 4
    return torch.where(x <= -sigma, lin, torch.where(x < sigma, quad, torch.zeros_like(x)))


def cone_field(points, centre, radius, normal, sigma):
    """psi = (1 - Phi) * Upsilon for points [...,3,3] in the cones [...]: -> [...,3]."""
    d = points - centre[..., None, :]
    x = (d * normal[..., None, :]).sum(-1)
    p = d - x[..., None] * normal[..., None, :]
    rad = torch.sqrt((p * p).sum(-1))
    den = radius[..., None] * (1 - x / sigma)
    phi = rad / den
    psi = (1 - phi) * upsilon(x, sigma)
    return torch.where((phi < 1) & (den > 0), psi, torch.zeros_like(psi))


class DistanceFieldPenetrationLoss(torch.nn.Module):
    """mesh_intersection.loss.DistanceFieldPenetrationLoss: ``(triangles [B,F,3,3],
    collision_idxs [B,C,2]) -> loss [B]``."""

    def __init__(self, sigma=0.5, point2plane=False, vectorized=True, penalize_outside=True,
                 linear_max=1000):
        super(DistanceFieldPenetrationLoss, self).__init__()
        if point2plane or not penalize_outside:
            raise NotImplementedError('only point2plane=False, penalize_outside=True (the values '
                                      'of every shipped configuration) are restated')
        self.sigma = sigma

    def forward(self, triangles, collision_idxs):
        B = triangles.shape[0]
        out = []
        for b in range(B):
            c = collision_idxs[b]
            c = c[c[:, 0] >= 0]
            if c.shape[0] == 0:
                out.append(triangles.new_zeros(()))
                continue
            recv, intr = triangles[b, c[:, 0]], triangles[b, c[:, 1]]
            ro, rr, rn = circumcircle(recv)
            io, ir, in_ = circumcircle(intr)
            psi_r = cone_field(intr, ro, rr, rn, self.sigma)      # intruder vertices, receiver cones
            psi_i = cone_field(recv, io, ir, in_, self.sigma)
            out.append((psi_r ** 4).sum() + (psi_i ** 4).sum())
        return torch.stack(out)


def load_part_segmentation(path):
    import pickle
    with open(path, 'rb') as f:
        d = pickle.load(f, encoding='latin1')
    return np.asarray(d['segm']).astype(np.int64), np.asarray(d['parents']).astype(np.int64)
