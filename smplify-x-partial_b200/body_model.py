"""SMPL-X body model with the surface the reference uses from ``smplx`` (call sites:
smplifyx/main.py:109-127, fitting.py:82,229,248, fit_single_frame.py:274,551,611):

``create(model_path, model_type='smplx', gender=, joint_mapper=, create_*=, num_betas=,
num_pca_comps=, use_face_contour=, dtype=, batch_size=)`` -> module with parameters ``betas``,
``global_orient``, ``body_pose``, ``left_hand_pose``, ``right_hand_pose``, ``jaw_pose``,
``leye_pose``, ``reye_pose``, ``expression``; ``forward(body_pose=, return_verts=,
return_full_pose=)`` -> output with ``.vertices .joints .full_pose .betas ...``;
``reset_params(**)``; ``faces`` / ``faces_tensor``.

The arithmetic runs in libsfx (CUDA): ``forward`` is one ``sfx_eval`` / ``sfx_forward_mesh``
call on the module's frame batch.  Outputs carry no autograd graph: gradients come from the
kernel's analytic adjoint (``fitting.FittingMonitor.create_fitting_closure`` writes them into
``.grad``).
"""
import os
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

from . import _native as N
from . import engine

ModelOutput = namedtuple('ModelOutput', ['vertices', 'joints', 'full_pose', 'betas',
                                         'global_orient', 'body_pose', 'expression',
                                         'left_hand_pose', 'right_hand_pose', 'jaw_pose'])
ModelOutput.__new__.__defaults__ = (None,) * len(ModelOutput._fields)



class _Output(ModelOutput):
    """ModelOutput that remembers the module it came from (the loss needs the engine handle)."""
    owner = None


PARAM_NAMES = ['betas', 'global_orient', 'body_pose', 'left_hand_pose', 'right_hand_pose',
               'jaw_pose', 'leye_pose', 'reye_pose', 'expression']


def load_model(model_path, gender='neutral'):
    """``<model_path>/smplx/SMPLX_<GENDER>.npz`` (or a direct .npz / .pkl path) -> dict."""
    path = model_path
    if os.path.isdir(path):
        path = os.path.join(path, 'smplx', 'SMPLX_{}.npz'.format(gender.upper()))
    if path.endswith('.npz'):
        return dict(np.load(path, allow_pickle=True))
    import pickle
    with open(path, 'rb') as f:
        return dict(pickle.load(f, encoding='latin1'))


def create(model_path=None, model_type='smplx', model_data=None, gender='neutral', **kwargs):
    if model_type != 'smplx':
        raise ValueError('only model_type="smplx" is built (got {})'.format(model_type))
    if model_data is None:
        model_data = load_model(model_path, gender)
    kwargs.pop('model_folder', None)
    return SMPLX(model_data, **kwargs)


class SMPLX(nn.Module):
    NUM_JOINTS = 55

    def __init__(self, model_data, joint_mapper=None, create_global_orient=True,
                 create_body_pose=True, create_betas=True, create_left_hand_pose=True,
                 create_right_hand_pose=True, create_expression=True, create_jaw_pose=True,
                 create_leye_pose=True, create_reye_pose=True, create_transl=False, use_pca=True,
                 num_pca_comps=6, flat_hand_mean=False, num_betas=10, num_expression_coeffs=10,
                 use_face_contour=False, batch_size=1, dtype=torch.float32, device=None,
                 **kwargs):
        super(SMPLX, self).__init__()
        if create_transl:
            raise ValueError('create_transl is not supported (the reference passes False)')
        self.dtype = dtype
        self.batch_size = batch_size
        self.use_pca = use_pca
        self.num_pca_comps = num_pca_comps
        self.use_face_contour = use_face_contour
        self.joint_mapper = joint_mapper
        n_model_joints = 55 + 21 + 51 + (17 if use_face_contour else 0)
        maps = getattr(joint_mapper, 'joint_maps', None) if joint_mapper is not None else None
        joint_map = np.arange(n_model_joints) if maps is None else maps.cpu().numpy()
        self.engine_model = engine.Model(
            model_data, joint_map.astype(np.int32), dtype=dtype, device=device,
            num_betas=num_betas, num_expression_coeffs=num_expression_coeffs, use_pca=use_pca,
            num_pca_comps=num_pca_comps, flat_hand_mean=flat_hand_mean,
            use_face_contour=use_face_contour)
        dev = self.engine_model.device
        self.faces = np.asarray(model_data['f']).astype(np.int64)
        self.register_buffer('faces_tensor', torch.tensor(self.faces, dtype=torch.long, device=dev))
        hand_dim = num_pca_comps if use_pca else 45
        shapes = dict(betas=num_betas, global_orient=3, body_pose=63, left_hand_pose=hand_dim,
                      right_hand_pose=hand_dim, jaw_pose=3, leye_pose=3, reye_pose=3,
                      expression=num_expression_coeffs)
        flags = dict(betas=create_betas, global_orient=create_global_orient,
                     body_pose=create_body_pose, left_hand_pose=create_left_hand_pose,
                     right_hand_pose=create_right_hand_pose, jaw_pose=create_jaw_pose,
                     leye_pose=create_leye_pose, reye_pose=create_reye_pose,
                     expression=create_expression)
        for name in PARAM_NAMES:                     # smplx registration order
            if flags[name]:
                self.register_parameter(name, nn.Parameter(
                    torch.zeros([batch_size, shapes[name]], dtype=dtype, device=dev),
                    requires_grad=True))
            else:
                setattr(self, name, None)
        if use_pca:
            self.register_buffer('left_hand_components', torch.tensor(
                np.asarray(model_data['hands_componentsl'])[:num_pca_comps], dtype=dtype, device=dev))
            self.register_buffer('right_hand_components', torch.tensor(
                np.asarray(model_data['hands_componentsr'])[:num_pca_comps], dtype=dtype, device=dev))
        zeros = np.zeros(45)
        self.register_buffer('left_hand_mean', torch.tensor(
            zeros if flat_hand_mean else np.asarray(model_data['hands_meanl']), dtype=dtype, device=dev))
        self.register_buffer('right_hand_mean', torch.tensor(
            zeros if flat_hand_mean else np.asarray(model_data['hands_meanr']), dtype=dtype, device=dev))
        self._batches = {}
        self._joint_stage = {}

    # ------------------------------------------------------------------ engine plumbing
    def frame_batch(self, use_vposer=False):
        key = bool(use_vposer)
        if key not in self._batches:
            self._batches[key] = engine.FrameBatch(self.engine_model, self.batch_size, use_vposer)
        return self._batches[key]

    def write_params(self, batch, pose_embedding=None, camera_translation=None, overrides=None):
        """Module parameters (+ pose embedding, camera translation) -> the batch's parameter
        matrix on the device."""
        x = batch.params_tensor()
        overrides = overrides or {}
        with torch.no_grad():
            for name, (off, n) in batch.blocks.items():
                if name == 'pose_embedding':
                    src = pose_embedding
                elif name == 'camera_translation':
                    src = camera_translation
                else:
                    src = overrides.get(name, getattr(self, name, None))
                if src is None:
                    x[:, off:off + n].zero_()
                else:
                    x[:, off:off + n] = src.detach().reshape(self.batch_size, n).to(x.dtype)
        return x

    def read_params(self, batch, pose_embedding=None, camera_translation=None, names=None):
        x = batch.params_tensor()
        with torch.no_grad():
            for name, (off, n) in batch.blocks.items():
                if names is not None and name not in names:
                    continue
                if name == 'pose_embedding':
                    dst = pose_embedding
                elif name == 'camera_translation':
                    dst = camera_translation
                else:
                    dst = getattr(self, name, None)
                if dst is not None:
                    dst.copy_(x[:, off:off + n].reshape(dst.shape))

    # ------------------------------------------------------------------ smplx surface
    @torch.no_grad()
    def reset_params(self, **params_dict):
        for name, param in self.named_parameters():
            if name in params_dict:
                val = params_dict[name]
                if torch.is_tensor(val):
                    val = val.detach().clone()
                param[:] = torch.as_tensor(val, dtype=param.dtype, device=param.device).reshape(
                    param.shape)
            else:
                param.fill_(0)

    def forward(self, betas=None, global_orient=None, body_pose=None, left_hand_pose=None,
                right_hand_pose=None, expression=None, jaw_pose=None, leye_pose=None,
                reye_pose=None, return_verts=True, return_full_pose=False, **kwargs):
        given = dict(betas=betas, global_orient=global_orient, left_hand_pose=left_hand_pose,
                     right_hand_pose=right_hand_pose, expression=expression, jaw_pose=jaw_pose,
                     leye_pose=leye_pose, reye_pose=reye_pose)
        overrides = {k: v for k, v in given.items() if v is not None}
        body_pose = body_pose if body_pose is not None else self.body_pose
        batch = self.frame_batch(False)
        self.write_params(batch, pose_embedding=body_pose, overrides=overrides)
        vertices = None
        if return_verts:
            vertices, joints = batch.forward_mesh(want_joints=True)
        else:
            if id(batch) not in self._joint_stage:
                self._joint_stage[id(batch)] = N.make_stage(
                    batch.L, N.CAMERA_STAGE_BLOCKS, loss_kind=N.LOSS_CAMERA_INIT)
            _, _, joints = batch.eval(self._joint_stage[id(batch)], want_joints=True)
        cur = lambda n: overrides.get(n, getattr(self, n))
        lh, rh = cur('left_hand_pose'), cur('right_hand_pose')
        if self.use_pca:
            lh = lh @ self.left_hand_components
            rh = rh @ self.right_hand_components
        full_pose = None
        if return_full_pose:
            full_pose = torch.cat([cur('global_orient'), body_pose.reshape(self.batch_size, -1),
                                   cur('jaw_pose'), cur('leye_pose'), cur('reye_pose'),
                                   lh + self.left_hand_mean, rh + self.right_hand_mean], dim=1)
        out = _Output(vertices=vertices, joints=joints, betas=cur('betas'),
                          expression=cur('expression'), global_orient=cur('global_orient'),
                          body_pose=body_pose, left_hand_pose=lh, right_hand_pose=rh,
                          jaw_pose=cur('jaw_pose'), full_pose=full_pose)
        out.owner = self
        return out
