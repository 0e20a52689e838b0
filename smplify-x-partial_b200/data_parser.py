"""Dataset reader with the reference's contract (smplifyx/data_parser.py:46-282):
``create_dataset(format, data_folder, **kw)`` -> iterable of
``{'fn', 'img_path', 'keypoints' [P,K,3], 'img' [H,W,3] float32 in 0..1}`` over
``<data_folder>/images/*.{png,jpg}`` with OpenPose-style JSON in ``<data_folder>/keypoints/``
(``<name>_*.json``), plus ``get_model2data()``, ``get_joint_weights()`` and the shoulder
indices.  Keypoint row order: body, left hand, right hand, face[17:68], contour face[0:17].
"""
import glob
import json
import os

import numpy as np
import torch

from . import utils as U


def read_keypoints(keypoint_fn, use_hands=True, use_face=True, use_face_contour=False):
    """-> (keypoints list of [K,3] float32 per person, gender_pd, gender_gt)."""
    with open(keypoint_fn) as f:
        data = json.load(f)
    people, gender_pd, gender_gt = [], [], []

    def block(person, key):
        return np.array(person[key], dtype=np.float32).reshape(-1, 3)
    for person in data['people']:
        rows = [block(person, 'pose_keypoints_2d')]
        if use_hands:
            rows += [block(person, 'hand_left_keypoints_2d'),
                     block(person, 'hand_right_keypoints_2d')]
        if use_face:
            face = block(person, 'face_keypoints_2d')
            rows.append(face[17:17 + 51])
            if use_face_contour:
                rows.append(face[:17])
        people.append(np.concatenate(rows, axis=0))
        if 'gender_pd' in person:
            gender_pd.append(person['gender_pd'])
        if 'gender_gt' in person:
            gender_gt.append(person['gender_gt'])
    return people, gender_pd, gender_gt


class KeypointDataset(object):
    def __init__(self, data_folder, fmt='coco25', img_folder='images', keyp_folder='keypoints',
                 use_hands=False, use_face=False, dtype=torch.float32, model_type='smplx',
                 joints_to_ign=None, use_face_contour=False, **kwargs):
        fmt = fmt.lower()
        if fmt not in U.NUM_BODY_KEYPOINTS:
            raise ValueError('Unknown dataset: {}'.format(fmt))
        self.format = fmt
        self.num_joints = U.NUM_BODY_KEYPOINTS[fmt] + 2 * 20 * bool(use_hands)
        self.use_hands, self.use_face = use_hands, use_face
        self.use_face_contour = use_face_contour
        self.model_type, self.dtype = model_type, dtype
        self.joints_to_ign = joints_to_ign
        self.img_folder = os.path.join(data_folder, img_folder)
        self.keyp_folder = os.path.join(data_folder, keyp_folder)
        self.img_paths = sorted(
            os.path.join(self.img_folder, fn) for fn in os.listdir(self.img_folder)
            if fn.endswith('.png') or (fn.endswith('.jpg') and not fn.startswith('.')))
        self.cnt = 0

    def get_model2data(self):
        return U.smpl_to_annotation(self.model_type, use_hands=self.use_hands,
                                    use_face=self.use_face,
                                    use_face_contour=self.use_face_contour, format=self.format)

    def get_left_shoulder(self):
        return 2

    def get_right_shoulder(self):
        return 5

    def get_joint_weights(self):
        w = np.ones(self.num_joints + 2 * self.use_hands + 51 * self.use_face +
                    17 * self.use_face_contour, dtype=np.float32)
        if self.joints_to_ign is not None and -1 not in self.joints_to_ign:
            w[self.joints_to_ign] = 0.
        return torch.tensor(w, dtype=self.dtype)

    def __len__(self):
        return len(self.img_paths)

    def read_item(self, img_path):
        import cv2
        img = cv2.imread(img_path).astype(np.float32)[:, :, ::-1] / 255.0
        fn = os.path.splitext(os.path.basename(img_path))[0]
        matches = glob.glob(os.path.join(self.keyp_folder, fn + '_*.json'))
        if not matches:
            return {}
        people, gpd, ggt = read_keypoints(matches[0], use_hands=self.use_hands,
                                          use_face=self.use_face,
                                          use_face_contour=self.use_face_contour)
        if len(people) < 1:
            return {}
        out = {'fn': fn, 'img_path': img_path, 'keypoints': np.stack(people), 'img': img}
        if gpd:
            out['gender_pd'] = gpd
        if ggt:
            out['gender_gt'] = ggt
        return out

    def read_meta(self, img_path):
        """``read_item`` without the pixels: the fit only ever uses the image's height and width
        (main.py:208-214), so the batched driver reads the header instead of decoding (and
        keeping, 25 MB per 1080p frame) every image of the folder.  -> {} or dict(fn, img_path,
        keypoints [P,K,3], H, W)."""
        fn = os.path.splitext(os.path.basename(img_path))[0]
        matches = glob.glob(os.path.join(self.keyp_folder, fn + '_*.json'))
        if not matches:
            return {}
        people, gpd, ggt = read_keypoints(matches[0], use_hands=self.use_hands,
                                          use_face=self.use_face,
                                          use_face_contour=self.use_face_contour)
        if len(people) < 1:
            return {}
        try:
            from PIL import Image
            with Image.open(img_path) as im:
                W, H = im.size
        except Exception:
            import cv2
            H, W = cv2.imread(img_path).shape[:2]
        return {'fn': fn, 'img_path': img_path, 'keypoints': np.stack(people), 'H': int(H), 'W': int(W)}

    def __getitem__(self, idx):
        return self.read_item(self.img_paths[idx])

    def __iter__(self):
        self.cnt = 0
        return self

    def __next__(self):
        if self.cnt >= len(self.img_paths):
            raise StopIteration
        path = self.img_paths[self.cnt]
        self.cnt += 1
        return self.read_item(path)

    next = __next__


def create_dataset(format='coco25', data_folder='data', **kwargs):
    return KeypointDataset(data_folder, fmt=format, **kwargs)


# ------------------------------------------------------------------------------ device path
def pack_keypoints_device(body, left_hand, right_hand, face, use_face_contour=True, device=None):
    """``read_keypoints`` without the JSON parsing, on the device (libsfx ``sfx_pack_keypoints``):
    raw OpenPose blocks of B frames -- body [B,n_body,3], hands [B,21,3], face [B,>=68,3]
    (numpy or CUDA tensors, float32) -> CUDA tensor [B,K,3] in the reference's row order."""
    import ctypes as C
    from . import _native as N
    lib = N.load_library()
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32) if not torch.is_tensor(a) else a,
                                  dtype=torch.float32, device=dev).contiguous()
    body, lh, rh, face = t(body), t(left_hand), t(right_hand), t(face)
    B, nb, nf = body.shape[0], body.shape[1], face.shape[1]
    K = nb + 42 + 51 + (17 if use_face_contour else 0)
    out = torch.empty((B, K, 3), dtype=torch.float32, device=dev)
    p = lambda x: C.c_void_p(x.data_ptr())
    with torch.cuda.device(dev):
        N.check(lib, lib.sfx_pack_keypoints(p(body), p(lh), p(rh), p(face), B, nb, nf,
                                            int(bool(use_face_contour)), p(out),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out
