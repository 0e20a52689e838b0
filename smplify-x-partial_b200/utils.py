"""Host-side helpers mirroring ``smplifyx/utils.py`` of the reference: the model-joint ->
keypoint index tables (``smpl_to_annotation``, utils.py:98-250), ``JointMapper``
(utils.py:68-81), ``GMoF`` (utils.py:84-95), ``rel_change`` (utils.py:60-61) and the
rotation-matrix -> Euler helper used for the regression priors (utils.py:306-436).
"""
import math

import numpy as np
import torch
import torch.nn as nn

# --- index tables: which model joint feeds each keypoint, per keypoint format -------------
# hands: wrist, then thumb/index/middle/ring/pinky with the vertex-picked finger tip last
_HAND = {
    # (left, right) for models whose tips start at `tip0`
    'smplx': lambda tip0: (
        [20, 37, 38, 39, tip0, 25, 26, 27, tip0 + 1, 28, 29, 30, tip0 + 2,
         34, 35, 36, tip0 + 3, 31, 32, 33, tip0 + 4],
        [21, 52, 53, 54, tip0 + 5, 40, 41, 42, tip0 + 6, 43, 44, 45, tip0 + 7,
         49, 50, 51, tip0 + 8, 46, 47, 48, tip0 + 9]),
    'smplh': lambda tip0: (
        [20, 34, 35, 36, tip0, 22, 23, 24, tip0 + 1, 25, 26, 27, tip0 + 2,
         31, 32, 33, tip0 + 3, 28, 29, 30, tip0 + 4],
        [21, 49, 50, 51, tip0 + 5, 37, 38, 39, tip0 + 6, 40, 41, 42, tip0 + 7,
         46, 47, 48, tip0 + 8, 43, 44, 45, tip0 + 9]),
}
_TORSO = [12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7]     # OpenPose rows 1..14
_BODY = {
    ('coco25', 'smpl'): [24] + _TORSO + list(range(25, 35)),
    ('coco25', 'smplh'): [52] + _TORSO + list(range(53, 63)),
    ('coco25', 'smplx'): [55] + _TORSO + list(range(56, 66)),
    ('coco19', 'smpl'): [24] + _TORSO + list(range(25, 29)),
    ('coco19', 'smplh'): [52] + _TORSO + list(range(53, 57)),
    ('coco19', 'smplx'): [55] + _TORSO + list(range(56, 60)),
    ('halpe', 'smplx'): [55, 57, 56, 59, 58, 16, 17, 18, 19, 20, 21, 1, 2, 4, 5, 7, 8,
                         15, 12, 0, 60, 63, 61, 64, 62, 65],
    ('coco_wholebody', 'smplx'): [55, 57, 56, 59, 58, 16, 17, 18, 19, 20, 21, 1, 2, 4, 5, 7, 8,
                                  60, 61, 62, 63, 64, 65],
}
_TIP0 = {('coco25', 'smplh'): 63, ('coco25', 'smplx'): 66, ('coco19', 'smplh'): 57,
         ('coco19', 'smplx'): 60, ('halpe', 'smplx'): 66, ('coco_wholebody', 'smplx'): 66}
_FACE0 = {('coco25', 'smplx'): 76, ('coco19', 'smplx'): 70, ('halpe', 'smplx'): 76,
          ('coco_wholebody', 'smplx'): 76}


def smpl_to_annotation(model_type='smplx', use_hands=True, use_face=True,
                       use_face_contour=False, format='coco25'):
    """Indices of the model joints in keypoint order (same contract as the reference's
    ``smpl_to_annotation``: body, left hand, right hand, 51 face landmarks, 17 contour)."""
    fmt = format.lower() if format != 'coco19' else format
    key = (fmt, model_type)
    if fmt not in ('coco25', 'coco19', 'halpe', 'coco_wholebody'):
        raise ValueError('Unknown joint format: {}'.format(format))
    if key not in _BODY:
        raise ValueError('Unknown model type: {}'.format(model_type))
    parts = [np.array(_BODY[key], dtype=np.int32)]
    if model_type == 'smpl':
        return parts[0]
    if use_hands:
        left, right = _HAND[model_type](_TIP0[key])
        parts += [np.array(left, dtype=np.int32), np.array(right, dtype=np.int32)]
    if use_face and model_type == 'smplx':
        f0 = _FACE0[key]
        parts.append(np.arange(f0, f0 + 51 + 17 * use_face_contour, dtype=np.int32))
    return np.concatenate(parts)


NUM_BODY_KEYPOINTS = {'coco25': 25, 'halpe': 26, 'coco_wholebody': 23}


class JointMapper(nn.Module):
    def __init__(self, joint_maps=None):
        super(JointMapper, self).__init__()
        if joint_maps is None:
            self.joint_maps = joint_maps
        else:
            self.register_buffer('joint_maps', torch.tensor(joint_maps, dtype=torch.long))

    def forward(self, joints, **kwargs):
        if self.joint_maps is None:
            return joints
        return torch.index_select(joints, 1, self.joint_maps)


def rel_change(prev_val, curr_val):
    return (prev_val - curr_val) / max([abs(prev_val), abs(curr_val), 1])


def to_tensor(tensor, dtype=torch.float32):
    if torch.is_tensor(tensor):
        return tensor.clone().detach()
    return torch.tensor(tensor, dtype=dtype)


def euler_xyz_from_matrix(R):
    """Intrinsic x-y-z Euler angles (radians) of rotation matrices ``[..., 3, 3]`` ->
    ``[..., 3]``, i.e. R = Rx(a0) Ry(a1) Rz(a2), computed in the dtype of the input.

    Same algorithm as the reference's ``_compute_euler_from_matrix(dcm, 'xyz', False)``
    (utils.py:306-436, itself scipy's ``Rotation.as_euler``): rotate into the frame where the
    middle axis is z (O = C R C^T Rx(-pi/2)), read the middle angle from O[2,2] and the outer
    ones from the last row / column; at gimbal lock the third angle is set to zero."""
    R = np.asarray(R)
    dt = R.dtype if R.dtype in (np.float32, np.float64) else np.float64
    m = R.reshape(-1, 3, 3).astype(dt)
    C = np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=dt)
    lam = dt.type(math.atan2(1.0, 0.0)) if hasattr(dt, 'type') else np.dtype(dt).type(math.pi / 2)
    rot = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], dtype=dt)
    O = np.einsum('ij,njk->nik', C, m)
    O = np.einsum('nij,jk->nik', O, (C.T @ rot).astype(dt))
    o22 = np.clip(O[:, 2, 2], -1, 1)
    mid = np.arccos(o22).astype(dt)
    safe1 = np.abs(mid) >= 1e-7
    safe2 = np.abs(mid - dt.type(np.pi) if hasattr(dt, 'type') else mid - np.pi) >= 1e-7
    safe = safe1 & safe2
    ang = np.zeros((m.shape[0], 3), dtype=dt)
    ang[:, 1] = mid + lam
    ang[safe, 0] = np.arctan2(O[safe, 0, 2], -O[safe, 1, 2])
    ang[safe, 2] = np.arctan2(O[safe, 2, 0], O[safe, 2, 1])
    ns1, ns2 = ~safe1, ~safe2
    ang[ns1, 0] = np.arctan2(O[ns1, 1, 0] - O[ns1, 0, 1], O[ns1, 0, 0] + O[ns1, 1, 1])
    ang[ns2, 0] = np.arctan2(O[ns2, 1, 0] + O[ns2, 0, 1], O[ns2, 0, 0] - O[ns2, 1, 1])
    adj = ((ang[:, 1] < -np.pi / 2) | (ang[:, 1] > np.pi / 2)) & safe
    ang[adj, 0] += np.pi
    ang[adj, 1] = 2 * lam - ang[adj, 1]
    ang[adj, 2] -= np.pi
    ang[ang < -np.pi] += 2 * np.pi
    ang[ang > np.pi] -= 2 * np.pi
    return ang.reshape(R.shape[:-2] + (3,))


def rodrigues(r):
    """Axis-angle -> rotation matrix (numpy, float64)."""
    r = np.asarray(r, dtype=np.float64).reshape(3)
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)


def inv_rodrigues(R):
    """Rotation matrix -> axis-angle (numpy, float64); the inverse used for the 180-degree
    orientation flip (fit_single_frame.py:527-538 uses cv2.Rodrigues for it)."""
    R = np.asarray(R, dtype=np.float64)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = 0.5 * np.linalg.norm(v)
    c = min(1.0, max(-1.0, 0.5 * (np.trace(R) - 1.0)))
    th = math.acos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        d = np.sqrt(np.maximum((np.diag(R) + 1.0) * 0.5, 0.0))
        if R[0, 1] < 0:
            d[1] = -d[1]
        if R[0, 2] < 0:
            d[2] = -d[2]
        if abs(d[0]) < abs(d[1]) and abs(d[0]) < abs(d[2]) and (R[1, 2] > 0) != (d[1] * d[2] > 0):
            d[2] = -d[2]
        return d * (th / np.linalg.norm(d))
    return v * (th / (2.0 * s))
