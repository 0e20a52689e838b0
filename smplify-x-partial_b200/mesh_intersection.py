"""Drop-in for the three ``mesh_intersection`` objects the reference builds for its
interpenetration term (smplifyx/fit_single_frame.py:300-328) and hands to ``SMPLifyLoss``
(fitting.py:289-296):

    search_tree   = BVH(max_collisions=max_collisions)
    pen_distance  = DistanceFieldPenetrationLoss(sigma=df_cone_height, point2plane=point2plane,
                                                 vectorized=True, penalize_outside=penalize_outside)
    filter_faces  = FilterFaces(faces_segm=..., faces_parents=..., ign_part_pairs=...)

Here they only *describe* the term: search, filter, penalty and gradient all run inside the
per-frame evaluation kernel (``csrc/sfx_collide.cuh``), which ``SMPLifyLoss`` configures from
these objects (``sfx_model_set_collision`` + ``SfxStage.coll_loss_weight / coll_sigma``).  They
are not callable on tensors; there is no CPU path.
"""
import os
import pickle

import numpy as np


class BVH(object):
    """``max_collisions`` bounds, in the third-party package, the box-overlap candidates kept
    per triangle; the shipped configurations set 128 so that it never binds.  The device search
    keeps every pair, so the value is recorded but does not limit anything."""

    def __init__(self, max_collisions=8):
        self.max_collisions = int(max_collisions)

    def __call__(self, triangles):
        raise RuntimeError('the collision search runs inside the evaluation kernel; pass this '
                           'object to fitting.create_loss(search_tree=...)')


class DistanceFieldPenetrationLoss(object):
    def __init__(self, sigma=0.5, point2plane=False, vectorized=True, penalize_outside=True,
                 linear_max=1000):
        if point2plane:
            raise NotImplementedError('point2plane=True is not built (every shipped '
                                      'configuration uses False)')
        if not penalize_outside:
            raise NotImplementedError('penalize_outside=False is not built (every shipped '
                                      'configuration uses True)')
        if not sigma > 0:
            raise ValueError('sigma (df_cone_height) must be positive')
        self.sigma = float(sigma)
        self.point2plane, self.vectorized, self.penalize_outside = False, vectorized, True

    def __call__(self, triangles, collision_idxs):
        raise RuntimeError('the penetration penalty runs inside the evaluation kernel; pass this '
                           'object to fitting.create_loss(pen_distance=...)')

    def to(self, *a, **k):
        return self


class FilterFaces(object):
    """faces_segm [F] (body part of every face), faces_parents [F] (kinematic parent of that
    part), ign_part_pairs (["a,b", ...]): pairs of the same part, of parent / child parts and
    of the listed parts never collide."""

    def __init__(self, faces_segm=None, faces_parents=None, ign_part_pairs=None):
        if faces_segm is None or faces_parents is None:
            raise ValueError('FilterFaces needs faces_segm and faces_parents')
        self.faces_segm = np.ascontiguousarray(np.asarray(faces_segm), dtype=np.int32).reshape(-1)
        self.faces_parents = np.ascontiguousarray(np.asarray(faces_parents),
                                                  dtype=np.int32).reshape(-1)
        self.ign_part_pairs = list(ign_part_pairs) if ign_part_pairs else []

    def to(self, *a, **k):
        return self

    def __call__(self, collision_idxs):
        raise RuntimeError('the face filter runs inside the evaluation kernel; pass this object '
                           'to fitting.create_loss(tri_filtering_module=...)')


def load_part_segmentation(part_segm_fn):
    """The pickle the reference reads at fit_single_frame.py:317-323 -> (segm, parents)."""
    with open(os.path.expandvars(part_segm_fn), 'rb') as f:
        d = pickle.load(f, encoding='latin1')
    return np.asarray(d['segm']), np.asarray(d['parents'])


def create_term(interpenetration, max_collisions=8, df_cone_height=0.5, point2plane=False,
                penalize_outside=True, part_segm_fn='', ign_part_pairs=None, part_segm=None):
    """fit_single_frame.py:296-328: (search_tree, pen_distance, filter_faces), all None when the
    term is off.  ``part_segm`` = {'segm', 'parents'} can stand in for the pickle file."""
    if not interpenetration:
        return None, None, None
    search_tree = BVH(max_collisions=max_collisions)
    pen_distance = DistanceFieldPenetrationLoss(sigma=df_cone_height, point2plane=point2plane,
                                                vectorized=True, penalize_outside=penalize_outside)
    filter_faces = None
    if part_segm is not None:
        filter_faces = FilterFaces(part_segm['segm'], part_segm['parents'], ign_part_pairs)
    elif part_segm_fn:
        segm, parents = load_part_segmentation(part_segm_fn)
        filter_faces = FilterFaces(segm, parents, ign_part_pairs)
    return search_tree, pen_distance, filter_faces
