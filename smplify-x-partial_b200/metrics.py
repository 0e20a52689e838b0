"""Evaluation metrics of the reference on the device: the alignment classes of
``smplifyx/utils.py:540-771`` and ``compute_v2v`` of ``smplifyx/eval.py:14-44`` with the same
names, call signatures and return values, batched -- the reference aligns one frame at a time
with numpy on the host; here one ``sfx_aligned_errors`` launch handles every frame of a batch.

    alignments = {'procrustes': ProcrustesAlignmentMPJPE(), 'pelvis': PelvisAlignmentMPJPE()}
    out = compute_v2v(vertices_fitted [B,V,3], vertices_target [B,V,3], alignments, vids=None)
    out['point']['procrustes']   # [B, n] per-vertex error after similarity alignment (PA-V2V)

F-scores (``point_fscore``, open3d nearest-neighbour distances) are not built: the reference's
``eval.py`` constructs its alignments without thresholds, so they are never computed there.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N


def _device_tensor(x, dtype=None):
    if not torch.cuda.is_available():
        raise RuntimeError('smplifyx_b200 needs a CUDA device (sm_100a); there is no CPU path')
    t = torch.as_tensor(x)
    if dtype is None:
        dtype = torch.float64 if t.dtype == torch.float64 else torch.float32
    return t.to(device='cuda', dtype=dtype).contiguous()


def aligned_errors(est, gt, mode, vids=None, hips_idxs=(2, 3), return_transform=False):
    """est, gt: [B,N,3] (or [N,3]) numpy / torch, any device -> per-point errors [B,n] as a
    tensor on the device (float64 inputs stay float64)."""
    est = _device_tensor(est)
    gt = _device_tensor(gt, est.dtype)
    single = est.dim() == 2
    if single:
        est, gt = est[None], gt[None]
    if est.shape != gt.shape or est.dim() != 3 or est.shape[-1] != 3:
        raise ValueError('point sets must have the same shape [B, N, 3]')
    B, Np, _ = est.shape
    idx = None
    n = Np
    if vids is not None:
        idx = torch.as_tensor(np.asarray(vids), dtype=torch.int32, device=est.device).contiguous()
        n = int(idx.numel())
    err = torch.empty((B, n), dtype=est.dtype, device=est.device)
    tr = torch.empty((B, 13), dtype=est.dtype, device=est.device) if return_transform else None
    lib = N.load_library()
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    with torch.cuda.device(est.device):
        N.check(lib, lib.sfx_aligned_errors(
            p(est), p(gt), p(idx), B, Np, n, int(mode), int(hips_idxs[0]), int(hips_idxs[1]),
            int(est.dtype == torch.float64), p(err), p(tr),
            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    if single:
        err = err[0]
        tr = None if tr is None else tr[0]
    return (err, tr) if return_transform else err


def mpjpe(input_joints, target_joints):
    """utils.py:596-612."""
    return aligned_errors(input_joints, target_joints, N.ALIGN_NONE)


vertex_to_vertex_error = mpjpe        # utils.py:614-615: the same formula


class ProcrustesAlignment(object):
    """utils.py:540-594: returns the aligned copy of S1 ([N,3] or [B,N,3])."""

    def __repr__(self):
        return 'ProcrustesAlignment'

    _mode = N.ALIGN_PROCRUSTES

    def __call__(self, S1, S2):
        S1d = _device_tensor(S1)
        _, tr = aligned_errors(S1d, S2, self._mode, return_transform=True)
        tr = tr.reshape(-1, 13)
        pts = S1d.reshape(tr.shape[0], -1, 3)
        R = tr[:, 1:10].reshape(-1, 3, 3)
        out = tr[:, :1, None] * torch.einsum('brc,bnc->bnr', R, pts) + tr[:, None, 10:13]
        return out.reshape(S1d.shape)


class ScaleAlignment(ProcrustesAlignment):
    """utils.py:729-771."""
    _mode = N.ALIGN_SCALE

    def __repr__(self):
        return 'ScaleAlignment'


class PelvisAlignment(object):
    """utils.py:650-671."""

    def __init__(self, hips_idxs=None):
        self.hips_idxs = [2, 3] if hips_idxs is None else list(hips_idxs)

    def align_by_pelvis(self, joints):
        j = _device_tensor(joints)
        pelvis = j[..., self.hips_idxs, :].mean(dim=-2, keepdim=True)
        return {'joints': j - pelvis, 'pelvis': pelvis}

    def __call__(self, gt, est):
        return self.align_by_pelvis(gt)['joints'], self.align_by_pelvis(est)['joints']


class _MetricMixin(object):
    def _fscore(self):
        if getattr(self, 'fscore_thresholds', None) is not None:
            raise NotImplementedError('point_fscore (open3d) is not built; eval.py never asks for it')
        return {}


class PelvisAlignmentMPJPE(PelvisAlignment, _MetricMixin):
    """utils.py:673-698."""

    def __init__(self, fscore_thresholds=None):
        super(PelvisAlignmentMPJPE, self).__init__()
        self.fscore_thresholds = fscore_thresholds

    def __call__(self, est_points, gt_points, vids=None):
        return {'point': aligned_errors(est_points, gt_points, N.ALIGN_PELVIS, vids=vids,
                                        hips_idxs=self.hips_idxs), 'fscore': self._fscore()}


class ProcrustesAlignmentMPJPE(ProcrustesAlignment, _MetricMixin):
    """utils.py:773-801 (the later definition, the one eval.py imports)."""

    def __init__(self, fscore_thresholds=None):
        super(ProcrustesAlignmentMPJPE, self).__init__()
        self.fscore_thresholds = fscore_thresholds

    def __call__(self, est_points, gt_points, vids=None):
        return {'point': aligned_errors(est_points, gt_points, N.ALIGN_PROCRUSTES, vids=vids),
                'fscore': self._fscore()}


def compute_v2v(vertices_fitted, vertices_target, alignments, vids=None):
    """eval.py:14-44 for a whole batch at once: {'point': {name: [B, n] numpy}, 'fscore': {...}}."""
    out = {'point': {}, 'fscore': {}}
    for name, alignment in alignments.items():
        r = alignment(vertices_fitted, vertices_target, vids=vids)
        out['point'][name] = r['point'].cpu().numpy()
        out['fscore'][name] = {}
    return out
