"""Perspective camera with the reference's surface (smplifyx/camera.py:35-117):
``create_camera('persp', rotation, translation, focal_length_x, focal_length_y, batch_size,
center, dtype)`` -> module with ``rotation`` [B,3,3] / ``translation`` [B,3] parameters,
``focal_length_x/y`` [B] and ``center`` [B,2] buffers, ``forward(points [B,N,3]) -> [B,N,2]``.

Inside a fit the projection runs fused in the CUDA evaluation kernel (csrc/sfx_core.cuh, step
6); ``forward`` here is the same formula in torch for callers that project points themselves
(visualisation, data generation).
"""
import torch
import torch.nn as nn


def create_camera(camera_type='persp', **kwargs):
    if camera_type.lower() == 'persp':
        return PerspectiveCamera(**kwargs)
    raise ValueError('Uknown camera type: {}'.format(camera_type))


class PerspectiveCamera(nn.Module):
    FOCAL_LENGTH = 5000

    def __init__(self, rotation=None, translation=None, focal_length_x=None,
                 focal_length_y=None, batch_size=1, center=None, dtype=torch.float32, **kwargs):
        super(PerspectiveCamera, self).__init__()
        self.batch_size = batch_size
        self.dtype = dtype
        self.register_buffer('zero', torch.zeros([batch_size], dtype=dtype))

        def focal(v):
            if v is None or type(v) in (float, int):
                return torch.full([batch_size], self.FOCAL_LENGTH if v is None else v, dtype=dtype)
            return torch.as_tensor(v, dtype=dtype).reshape(-1).expand(batch_size).clone()
        self.register_buffer('focal_length_x', focal(focal_length_x))
        self.register_buffer('focal_length_y', focal(focal_length_y))
        if center is None:
            center = torch.zeros([batch_size, 2], dtype=dtype)
        self.register_buffer('center', torch.as_tensor(center, dtype=dtype).reshape(-1, 2)
                             .expand(batch_size, 2).clone())
        if rotation is None:
            rotation = torch.eye(3, dtype=dtype).unsqueeze(dim=0).repeat(batch_size, 1, 1)
        self.register_parameter('rotation', nn.Parameter(
            torch.as_tensor(rotation, dtype=dtype).clone(), requires_grad=True))
        if translation is None:
            translation = torch.zeros([batch_size, 3], dtype=dtype)
        self.register_parameter('translation', nn.Parameter(
            torch.as_tensor(translation, dtype=dtype).clone(), requires_grad=True))

    def forward(self, points):
        p = torch.einsum('bij,bnj->bni', self.rotation, points) + self.translation.unsqueeze(1)
        img = p[:, :, :2] / p[:, :, 2:3]
        f = torch.stack([self.focal_length_x, self.focal_length_y], dim=-1).unsqueeze(1)
        return img * f + self.center.unsqueeze(dim=1)
