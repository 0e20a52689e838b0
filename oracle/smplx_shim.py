"""ORACLE (test infrastructure, not product code) -- CPU restatement of the un-vendored
``smplx`` package surface that the reference calls.

PARITY UNPINNED at this boundary: ``smplx`` (fork xiyichen/smplx of vchoutas/smplx 0.1.x,
installed from git HEAD, no version pin -- reference README.md:31) is absent from
/root/reference and no reference test pins its outputs.  This file
restates the published algorithm of upstream smplx 0.1.x:

* ``lbs.batch_rodrigues`` (angle = ||r + 1e-8||, R = I + sin K + (1-cos) K^2)
* ``lbs.blend_shapes``, ``lbs.vertices2joints``, ``lbs.batch_rigid_transform``, ``lbs.lbs``
* ``lbs.vertices2landmarks``, ``lbs.find_dynamic_lmk_idx_and_bcoords``, ``lbs.rot_mat_to_euler``
* ``lbs.transform_mat`` (used by reference smplifyx/camera.py:27,102)
* ``body_models.SMPLX.forward`` / ``reset_params`` / parameter registration order
* ``vertex_joint_selector.VertexJointSelector`` with ``vertex_ids['smplx']``

anchored on the reference's own call sites: smplifyx/main.py:109-127 (constructor kwargs),
smplifyx/fitting.py:82,248 (forward kwargs and the output attributes the loss reads,
fitting.py:378-435), smplifyx/fit_single_frame.py:274,551 (reset_params).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.
"""
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ModelOutput = namedtuple('ModelOutput',
                         ['vertices', 'joints', 'full_pose', 'betas',
                          'global_orient', 'body_pose', 'expression',
                          'left_hand_pose', 'right_hand_pose', 'jaw_pose'])
ModelOutput.__new__.__defaults__ = (None,) * len(ModelOutput._fields)

# smplx.vertex_ids.vertex_ids['smplx'] in VertexJointSelector order
EXTRA_VERTEX_IDS = [9120, 9929, 9448, 616, 6,
                    5770, 5780, 8846, 8463, 8474, 8635,
                    5361, 4933, 5058, 5169, 5286,
                    8079, 7669, 7794, 7905, 8022]

NECK_IDX = 12
NUM_BODY_JOINTS = 21
NUM_HAND_JOINTS = 15
EXPR_OFFSET = 300


# ----------------------------------------------------------------------------- lbs
def transform_mat(R, t):
    """[B,3,3], [B,3,1] -> [B,4,4] homogeneous transform."""
    return torch.cat([F.pad(R, [0, 0, 0, 1]),
                      F.pad(t, [0, 0, 0, 1], value=1)], dim=2)


def batch_rodrigues(rot_vecs, dtype=torch.float32):
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos = torch.unsqueeze(torch.cos(angle), dim=1)
    sin = torch.unsqueeze(torch.sin(angle), dim=1)
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros((n, 1), dtype=dtype, device=rot_vecs.device)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view((n, 3, 3))
    ident = torch.eye(3, dtype=dtype, device=rot_vecs.device).unsqueeze(dim=0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def blend_shapes(betas, shape_disps):
    return torch.einsum('bl,mkl->bmk', [betas, shape_disps])


def vertices2joints(J_regressor, vertices):
    return torch.einsum('bik,ji->bjk', [vertices, J_regressor])


def batch_rigid_transform(rot_mats, joints, parents, dtype=torch.float32):
    joints = torch.unsqueeze(joints, dim=-1)
    rel_joints = joints.clone()
    rel_joints[:, 1:] = rel_joints[:, 1:] - joints[:, parents[1:]]
    transforms_mat = transform_mat(
        rot_mats.reshape(-1, 3, 3),
        rel_joints.reshape(-1, 3, 1)).reshape(-1, joints.shape[1], 4, 4)
    chain = [transforms_mat[:, 0]]
    for i in range(1, parents.shape[0]):
        chain.append(torch.matmul(chain[int(parents[i])], transforms_mat[:, i]))
    transforms = torch.stack(chain, dim=1)
    posed_joints = transforms[:, :, :3, 3]
    joints_homogen = F.pad(joints, [0, 0, 0, 1])
    rel_transforms = transforms - F.pad(
        torch.matmul(transforms, joints_homogen), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed_joints, rel_transforms


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents,
        lbs_weights, dtype=torch.float32):
    batch_size = max(betas.shape[0], pose.shape[0])
    device = betas.device
    v_shaped = v_template + blend_shapes(betas, shapedirs)
    J = vertices2joints(J_regressor, v_shaped)
    ident = torch.eye(3, dtype=dtype, device=device)
    rot_mats = batch_rodrigues(pose.view(-1, 3), dtype=dtype).view([batch_size, -1, 3, 3])
    pose_feature = (rot_mats[:, 1:, :, :] - ident).view([batch_size, -1])
    pose_offsets = torch.matmul(pose_feature, posedirs).view(batch_size, -1, 3)
    v_posed = pose_offsets + v_shaped
    J_transformed, A = batch_rigid_transform(rot_mats, J, parents, dtype=dtype)
    W = lbs_weights.unsqueeze(dim=0).expand([batch_size, -1, -1])
    num_joints = J_regressor.shape[0]
    T = torch.matmul(W, A.view(batch_size, num_joints, 16)).view(batch_size, -1, 4, 4)
    homogen_coord = torch.ones([batch_size, v_posed.shape[1], 1], dtype=dtype, device=device)
    v_posed_homo = torch.cat([v_posed, homogen_coord], dim=2)
    v_homo = torch.matmul(T, torch.unsqueeze(v_posed_homo, dim=-1))
    return v_homo[:, :, :3, 0], J_transformed


def rot_mat_to_euler(rot_mats):
    sy = torch.sqrt(rot_mats[:, 0, 0] * rot_mats[:, 0, 0] +
                    rot_mats[:, 1, 0] * rot_mats[:, 1, 0])
    return torch.atan2(-rot_mats[:, 2, 0], sy)


def find_dynamic_lmk_idx_and_bcoords(vertices, pose, dynamic_lmk_faces_idx,
                                     dynamic_lmk_b_coords, neck_kin_chain,
                                     dtype=torch.float32):
    batch_size = vertices.shape[0]
    aa_pose = torch.index_select(pose.view(batch_size, -1, 3), 1, neck_kin_chain)
    rot_mats = batch_rodrigues(aa_pose.view(-1, 3), dtype=dtype).view(batch_size, -1, 3, 3)
    rel_rot_mat = torch.eye(3, device=vertices.device, dtype=dtype).unsqueeze_(dim=0).repeat(
        batch_size, 1, 1)
    for idx in range(len(neck_kin_chain)):
        rel_rot_mat = torch.bmm(rot_mats[:, idx], rel_rot_mat)
    y_rot_angle = torch.round(
        torch.clamp(-rot_mat_to_euler(rel_rot_mat) * 180.0 / np.pi, max=39)).to(dtype=torch.long)
    neg_mask = y_rot_angle.lt(0).to(dtype=torch.long)
    mask = y_rot_angle.lt(-39).to(dtype=torch.long)
    neg_vals = mask * 78 + (1 - mask) * (39 - y_rot_angle)
    y_rot_angle = (neg_mask * neg_vals + (1 - neg_mask) * y_rot_angle)
    dyn_lmk_faces_idx = torch.index_select(dynamic_lmk_faces_idx, 0, y_rot_angle)
    dyn_lmk_b_coords = torch.index_select(dynamic_lmk_b_coords, 0, y_rot_angle)
    return dyn_lmk_faces_idx, dyn_lmk_b_coords


def vertices2landmarks(vertices, faces, lmk_faces_idx, lmk_bary_coords):
    batch_size, num_verts = vertices.shape[:2]
    device = vertices.device
    lmk_faces = torch.index_select(faces, 0, lmk_faces_idx.view(-1)).view(batch_size, -1, 3)
    lmk_faces = lmk_faces + torch.arange(
        batch_size, dtype=torch.long, device=device).view(-1, 1, 1) * num_verts
    lmk_vertices = vertices.reshape(-1, 3)[lmk_faces].view(batch_size, -1, 3, 3)
    return torch.einsum('blfi,blf->bli', [lmk_vertices, lmk_bary_coords])


# ----------------------------------------------------------------------- body model
class SMPLX(nn.Module):
    NUM_JOINTS = 55

    def __init__(self, model_data, joint_mapper=None, create_global_orient=True,
                 create_body_pose=True, create_betas=True, create_left_hand_pose=True,
                 create_right_hand_pose=True, create_expression=True, create_jaw_pose=True,
                 create_leye_pose=True, create_reye_pose=True, create_transl=False,
                 use_pca=True, num_pca_comps=6, flat_hand_mean=False, num_betas=10,
                 num_expression_coeffs=10, use_face_contour=False, batch_size=1,
                 dtype=torch.float32, **kwargs):
        super().__init__()
        self.dtype = dtype
        self.batch_size = batch_size
        self.use_pca = use_pca
        self.num_pca_comps = num_pca_comps
        self.use_face_contour = use_face_contour
        self.joint_mapper = joint_mapper
        d = model_data

        def buf(name, arr, dt=dtype):
            self.register_buffer(name, torch.tensor(np.asarray(arr), dtype=dt))

        shapedirs = np.asarray(d['shapedirs'])
        buf('shapedirs', shapedirs[:, :, :num_betas])
        buf('expr_dirs', shapedirs[:, :, EXPR_OFFSET:EXPR_OFFSET + num_expression_coeffs])
        self.faces = np.asarray(d['f']).astype(np.int64)
        buf('faces_tensor', self.faces, torch.long)
        buf('v_template', d['v_template'])
        buf('J_regressor', d['J_regressor'])
        posedirs = np.asarray(d['posedirs'])
        buf('posedirs', np.reshape(posedirs, [-1, posedirs.shape[-1]]).T)
        parents = np.asarray(d['kintree_table'])[0].astype(np.int64)
        parents[0] = -1
        buf('parents', parents, torch.long)
        buf('lbs_weights', d['weights'])
        buf('extra_joints_idxs', EXTRA_VERTEX_IDS, torch.long)

        # --- learnable parameters, registered in smplx order (SMPL -> SMPLH -> SMPLX) ---
        def par(name, shape, create):
            if create:
                self.register_parameter(
                    name, nn.Parameter(torch.zeros(shape, dtype=dtype), requires_grad=True))
        par('betas', [batch_size, num_betas], create_betas)
        par('global_orient', [batch_size, 3], create_global_orient)
        par('body_pose', [batch_size, NUM_BODY_JOINTS * 3], create_body_pose)
        par('transl', [batch_size, 3], create_transl)
        hand_dim = num_pca_comps if use_pca else 3 * NUM_HAND_JOINTS
        if use_pca:
            buf('left_hand_components', np.asarray(d['hands_componentsl'])[:num_pca_comps])
            buf('right_hand_components', np.asarray(d['hands_componentsr'])[:num_pca_comps])
        if flat_hand_mean:
            lmean = np.zeros(45)
            rmean = np.zeros(45)
        else:
            lmean = np.asarray(d['hands_meanl'])
            rmean = np.asarray(d['hands_meanr'])
        buf('left_hand_mean', lmean)
        buf('right_hand_mean', rmean)
        par('left_hand_pose', [batch_size, hand_dim], create_left_hand_pose)
        par('right_hand_pose', [batch_size, hand_dim], create_right_hand_pose)
        par('jaw_pose', [batch_size, 3], create_jaw_pose)
        par('leye_pose', [batch_size, 3], create_leye_pose)
        par('reye_pose', [batch_size, 3], create_reye_pose)
        par('expression', [batch_size, num_expression_coeffs], create_expression)

        pose_mean = torch.cat([torch.zeros(3 + 63 + 9, dtype=dtype),
                               self.left_hand_mean, self.right_hand_mean])
        self.register_buffer('pose_mean', pose_mean)

        buf('lmk_faces_idx', d['lmk_faces_idx'], torch.long)
        buf('lmk_bary_coords', d['lmk_bary_coords'])
        if use_face_contour:
            buf('dynamic_lmk_faces_idx', d['dynamic_lmk_faces_idx'], torch.long)
            buf('dynamic_lmk_bary_coords', d['dynamic_lmk_bary_coords'])
            chain = []
            cur = NECK_IDX
            while cur != -1:
                chain.append(cur)
                cur = int(parents[cur])
            buf('neck_kin_chain', chain, torch.long)

    @torch.no_grad()
    def reset_params(self, **params_dict):
        for name, param in self.named_parameters():
            if name in params_dict:
                val = params_dict[name]
                if torch.is_tensor(val):
                    val = val.detach().clone()
                param[:] = torch.as_tensor(val, dtype=param.dtype).reshape(param.shape)
            else:
                param.fill_(0)

    def forward(self, betas=None, global_orient=None, body_pose=None,
                left_hand_pose=None, right_hand_pose=None, expression=None,
                jaw_pose=None, leye_pose=None, reye_pose=None,
                return_verts=True, return_full_pose=False, **kwargs):
        global_orient = global_orient if global_orient is not None else self.global_orient
        body_pose = body_pose if body_pose is not None else self.body_pose
        betas = betas if betas is not None else self.betas
        left_hand_pose = left_hand_pose if left_hand_pose is not None else self.left_hand_pose
        right_hand_pose = right_hand_pose if right_hand_pose is not None else self.right_hand_pose
        jaw_pose = jaw_pose if jaw_pose is not None else self.jaw_pose
        leye_pose = leye_pose if leye_pose is not None else self.leye_pose
        reye_pose = reye_pose if reye_pose is not None else self.reye_pose
        expression = expression if expression is not None else self.expression

        if self.use_pca:
            left_hand_pose = torch.einsum('bi,ij->bj', [left_hand_pose, self.left_hand_components])
            right_hand_pose = torch.einsum('bi,ij->bj', [right_hand_pose, self.right_hand_components])

        full_pose = torch.cat([global_orient, body_pose, jaw_pose, leye_pose, reye_pose,
                               left_hand_pose, right_hand_pose], dim=1)
        full_pose = full_pose + self.pose_mean
        batch_size = max(betas.shape[0], global_orient.shape[0], body_pose.shape[0])

        shape_components = torch.cat([betas, expression], dim=-1)
        shapedirs = torch.cat([self.shapedirs, self.expr_dirs], dim=-1)
        vertices, joints = lbs(shape_components, full_pose, self.v_template, shapedirs,
                               self.posedirs, self.J_regressor, self.parents,
                               self.lbs_weights, dtype=self.dtype)

        lmk_faces_idx = self.lmk_faces_idx.unsqueeze(dim=0).expand(batch_size, -1).contiguous()
        lmk_bary_coords = self.lmk_bary_coords.unsqueeze(dim=0).repeat(batch_size, 1, 1)
        if self.use_face_contour:
            dyn_faces, dyn_bary = find_dynamic_lmk_idx_and_bcoords(
                vertices, full_pose, self.dynamic_lmk_faces_idx,
                self.dynamic_lmk_bary_coords, self.neck_kin_chain, dtype=self.dtype)
            lmk_faces_idx = torch.cat([lmk_faces_idx, dyn_faces], 1)
            lmk_bary_coords = torch.cat([lmk_bary_coords, dyn_bary], 1)
        landmarks = vertices2landmarks(vertices, self.faces_tensor, lmk_faces_idx, lmk_bary_coords)

        extra = torch.index_select(vertices, 1, self.extra_joints_idxs)
        joints = torch.cat([joints, extra, landmarks], dim=1)
        if self.joint_mapper is not None:
            joints = self.joint_mapper(joints=joints, vertices=vertices)

        return ModelOutput(vertices=vertices if return_verts else None,
                           joints=joints, betas=betas, expression=expression,
                           global_orient=global_orient, body_pose=body_pose,
                           left_hand_pose=left_hand_pose, right_hand_pose=right_hand_pose,
                           jaw_pose=jaw_pose,
                           full_pose=full_pose if return_full_pose else None)


def create(model_path=None, model_type='smplx', model_data=None, gender='neutral', **kwargs):
    """``smplx.create`` stand-in.  ``model_data`` (dict) wins over ``model_path`` (npz file
    ``<model_path>/smplx/SMPLX_<GENDER>.npz`` or a direct path)."""
    import os
    if model_type != 'smplx':
        raise ValueError('oracle shim only restates model_type="smplx"')
    if model_data is None:
        path = model_path
        if os.path.isdir(path):
            path = os.path.join(path, 'smplx', 'SMPLX_{}.npz'.format(gender.upper()))
        model_data = dict(np.load(path, allow_pickle=True))
    kwargs.pop('model_folder', None)
    return SMPLX(model_data, **kwargs)
