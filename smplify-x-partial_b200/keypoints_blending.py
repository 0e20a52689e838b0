"""Keypoint blending of the reference (``smplifyx/keypoints_blending.py:276-381``), vectorised
over frames: OpenPose (BODY_25 + hands + 68 face points) and MMPose (Halpe-26 body + hands + face)
detections are merged per keypoint -- MMPose confidences are moved onto OpenPose's scale through
per-keypoint z-scores (the SHHQ statistics files ``openpose_{means,stds}.json`` /
``mmpose_{means,stds}.json``, which the reference expects under a ``heuristics`` folder and does
not ship), the more confident detection wins; face points always come from OpenPose.

Rows follow the raw OpenPose order the reference writes to ``<name>_blended.json``: body 0-24,
left hand 25-45, right hand 46-66, face 67-134.  Host-side data preparation (numpy), like the
reference's; it feeds ``data_parser`` and nothing on the device path depends on it.

One deviation: the reference writes the JSON after its per-image loop (an indentation slip,
keypoints_blending.py:371-381), i.e. only for the last image; ``blending`` here writes one file
per image.
"""
import glob
import json
import os

import numpy as np

# name -> (MMPose index, OpenPose index)   keypoints_blending.py:288-312
BODY_PAIRS = [
    ('Nose', 0, 0), ('LEye', 1, 16), ('REye', 2, 15), ('LEar', 3, 18), ('REar', 4, 17),
    ('LShoulder', 5, 5), ('RShoulder', 6, 2), ('LElbow', 7, 6), ('RElbow', 8, 3),
    ('LWrist', 9, 7), ('RWrist', 10, 4), ('LHip', 11, 12), ('RHip', 12, 9), ('LKnee', 13, 13),
    ('RKnee', 14, 10), ('LAnkle', 15, 14), ('RAnkle', 16, 11), ('Neck', 18, 1), ('Hip', 19, 8),
    ('LBigToe', 20, 19), ('RBigToe', 21, 22), ('LSmallToe', 22, 20), ('RSmallToe', 23, 23),
    ('LHeel', 24, 21), ('RHeel', 25, 24)]
OPENPOSE_BODY, MMPOSE_BODY = 25, 26


def blended_pairs():
    """[(name, MMPose row, OpenPose row)] of every keypoint that can come from either detector
    (body + both hands; keypoints_blending.py:314-324)."""
    pairs = list(BODY_PAIRS)
    for i in range(21):
        pairs.append(('left_hand_' + str(i + 1), MMPOSE_BODY + i, OPENPOSE_BODY + i))
    for i in range(21):
        pairs.append(('right_hand_' + str(i + 1), MMPOSE_BODY + 21 + i, OPENPOSE_BODY + 21 + i))
    return pairs


def load_statistics(heuristics_dir):
    out = {}
    for name in ('openpose_means', 'openpose_stds', 'mmpose_means', 'mmpose_stds'):
        with open(os.path.join(heuristics_dir, name + '.json')) as f:
            out[name] = json.load(f)
    return out


def blend_keypoints(openpose, mmpose, stats):
    """openpose [..., 135, 3], mmpose [..., 136, 3] (x, y, confidence; rows as read by the
    reference's ``read_keypoints(..., use_face_contour=True)``) -> blended [..., 135, 3] float64
    (keypoints_blending.py:337-369)."""
    op = np.asarray(openpose, dtype=np.float32)
    mm = np.asarray(mmpose, dtype=np.float32)
    if op.shape[-2:] != (135, 3) or mm.shape[-2:] != (136, 3) or op.shape[:-2] != mm.shape[:-2]:
        raise ValueError('expected openpose [..., 135, 3] and mmpose [..., 136, 3]')
    pairs = blended_pairs()
    mi = np.array([p[1] for p in pairs])
    oi = np.array([p[2] for p in pairs])
    st = lambda key: np.array([stats[key][p[0]] for p in pairs], dtype=np.float32)
    out = np.zeros(op.shape, dtype=np.float64)
    # face: OpenPose, confidence clipped to [0, 1]
    out[..., 67:, :2] = op[..., 67:, :2]
    out[..., 67:, 2] = np.clip(op[..., 67:, 2], 0, 1)
    op_conf = np.clip(op[..., oi, 2], 0, 1)
    mm_conf = (mm[..., mi, 2] - st('mmpose_means')) / st('mmpose_stds')
    mm_conf = np.clip(mm_conf * st('openpose_stds') + st('openpose_means'), 0, 1)
    take = mm_conf > op_conf
    out[..., oi, 0] = np.where(take, mm[..., mi, 0], op[..., oi, 0])
    out[..., oi, 1] = np.where(take, mm[..., mi, 1], op[..., oi, 1])
    out[..., oi, 2] = np.where(take, mm_conf, op_conf)
    return out


def blend_keypoints_device(openpose, mmpose, stats, device=None):
    """``blend_keypoints`` on the device (libsfx ``sfx_blend_keypoints``; float32 arithmetic
    without fused multiply-adds, so the result equals the numpy version bit for bit):
    openpose [B,135,3], mmpose [B,136,3] (numpy or CUDA tensors) -> CUDA tensor [B,135,3]."""
    import ctypes as C
    import torch
    from . import _native as N
    lib = N.load_library()
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32) if not torch.is_tensor(a) else a,
                                  dtype=torch.float32, device=dev).contiguous()
    op, mm = t(openpose), t(mmpose)
    if op.shape[-2:] != (135, 3) or mm.shape[-2:] != (136, 3) or op.shape[0] != mm.shape[0]:
        raise ValueError('expected openpose [B, 135, 3] and mmpose [B, 136, 3]')
    pairs = blended_pairs()
    st = np.stack([[stats[key][p[0]] for p in pairs] for key in
                   ('mmpose_means', 'mmpose_stds', 'openpose_means', 'openpose_stds')]).astype(np.float32)
    st_d = torch.as_tensor(st, device=dev).contiguous()
    mi = torch.as_tensor(np.array([p[1] for p in pairs], dtype=np.int32), device=dev)
    oi = torch.as_tensor(np.array([p[2] for p in pairs], dtype=np.int32), device=dev)
    out = torch.zeros((op.shape[0], 135, 3), dtype=torch.float32, device=dev)
    ptr = lambda x: C.c_void_p(x.data_ptr())
    with torch.cuda.device(dev):
        N.check(lib, lib.sfx_blend_keypoints(ptr(op), ptr(mm), ptr(st_d), ptr(mi), ptr(oi), len(pairs),
                                             int(op.shape[0]), ptr(out),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def read_raw_keypoints(keypoint_fn):
    """Person 0 of an OpenPose-style JSON in raw order: body | left hand | right hand | first 68
    face points (keypoints_blending.py:225-274 with its default ``orders``)."""
    with open(keypoint_fn) as f:
        person = json.load(f)['people'][0]
    blk = lambda key: np.array(person[key], dtype=np.float32).reshape(-1, 3)
    return np.concatenate([blk('pose_keypoints_2d'), blk('hand_left_keypoints_2d'),
                           blk('hand_right_keypoints_2d'), blk('face_keypoints_2d')[:68]], axis=0)


def blending(IMAGES_PATH, OPENPOSE_RES_DIR, MMPOSE_RES_DIR, BLENDING_RES_DIR, heuristics_dir):
    """File-level driver with the reference's signature plus the folder of the statistics files:
    ``<name>_keypoints.json`` + ``<name>_mmpose.json`` -> ``<name>_blended.json``."""
    stats = load_statistics(heuristics_dir)
    names = sorted(os.path.basename(fn).split('.')[0]
                   for fn in glob.glob(os.path.join(IMAGES_PATH, '*')))
    if not names:
        return []
    op = np.stack([read_raw_keypoints(os.path.join(OPENPOSE_RES_DIR, n + '_keypoints.json'))
                   for n in names])
    mm = np.stack([read_raw_keypoints(os.path.join(MMPOSE_RES_DIR, n + '_mmpose.json'))
                   for n in names])
    blended = blend_keypoints(op, mm, stats)
    os.makedirs(BLENDING_RES_DIR, exist_ok=True)
    written = []
    for n, b in zip(names, blended):
        flat = b.flatten().tolist()
        person = {'person_id': [-1], 'pose_keypoints_2d': flat[:25 * 3],
                  'hand_left_keypoints_2d': flat[25 * 3:46 * 3],
                  'hand_right_keypoints_2d': flat[46 * 3:67 * 3],
                  'face_keypoints_2d': flat[67 * 3:]}
        fn = os.path.join(BLENDING_RES_DIR, n + '_blended.json')
        with open(fn, 'w') as f:
            json.dump({'people': [person]}, f, indent=2)
        written.append(fn)
    return written
