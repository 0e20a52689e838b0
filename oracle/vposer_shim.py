"""ORACLE (test infrastructure, not product code) -- CPU restatement of VPoser v1.0
(``human_body_prior`` cvpr19 branch, ``human_body_prior/train/vposer_smpl.py``) and of the
``torchgeometry`` 0.1.2 rotation-matrix -> angle-axis conversion it calls.

PARITY UNPINNED at this boundary: ``human_body_prior`` and ``torchgeometry`` are absent from
/root/reference (README.md:32 installs a branch zip without a hash; requirements.txt:3 only says
``torchgeometry>=0.1.2``) and no reference test pins their outputs.  The
published algorithm restated here:

* encoder  BN(63) -> FC(63,512) -> leaky-ReLU(0.2) -> BN(512) -> dropout -> FC(512,512) ->
           leaky-ReLU(0.2) -> (mu = FC(512,32), sigma = softplus(FC(512,32))) -> Normal
* decoder  FC(32,512) -> leaky-ReLU(0.2) -> dropout -> FC(512,512) -> leaky-ReLU(0.2) ->
           FC(512,126) -> continuous 6-D rotation decoder (Gram-Schmidt, columns b1 b2 b3) ->
           rotation matrices [21,3,3] -> ``matrot2aa`` = torchgeometry
           rotation_matrix_to_angle_axis (via quaternions)

anchored on the reference's call sites: ``vposer.decode(pose_embedding, output_type='aa')
.view(1, -1)`` (smplifyx/fitting.py:72,236, fit_single_frame.py:265,515,607,620,654),
``vposer.encode(prior).sample()`` (fit_single_frame.py:245), ``vposer.eval()`` (:243).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def rotation_matrix_to_quaternion(rotation_matrix, eps=1e-6):
    """torchgeometry 0.1.2 core/conversions.py (note the transpose and the branch masks)."""
    rmat_t = torch.transpose(rotation_matrix, 1, 2)
    mask_d2 = rmat_t[:, 2, 2] < eps
    mask_d0_d1 = rmat_t[:, 0, 0] > rmat_t[:, 1, 1]
    mask_d0_nd1 = rmat_t[:, 0, 0] < -rmat_t[:, 1, 1]
    t0 = 1 + rmat_t[:, 0, 0] - rmat_t[:, 1, 1] - rmat_t[:, 2, 2]
    q0 = torch.stack([rmat_t[:, 1, 2] - rmat_t[:, 2, 1], t0,
                      rmat_t[:, 0, 1] + rmat_t[:, 1, 0], rmat_t[:, 2, 0] + rmat_t[:, 0, 2]], -1)
    t0_rep = t0.repeat(4, 1).t()
    t1 = 1 - rmat_t[:, 0, 0] + rmat_t[:, 1, 1] - rmat_t[:, 2, 2]
    q1 = torch.stack([rmat_t[:, 2, 0] - rmat_t[:, 0, 2], rmat_t[:, 0, 1] + rmat_t[:, 1, 0], t1,
                      rmat_t[:, 1, 2] + rmat_t[:, 2, 1]], -1)
    t1_rep = t1.repeat(4, 1).t()
    t2 = 1 - rmat_t[:, 0, 0] - rmat_t[:, 1, 1] + rmat_t[:, 2, 2]
    q2 = torch.stack([rmat_t[:, 0, 1] - rmat_t[:, 1, 0], rmat_t[:, 2, 0] + rmat_t[:, 0, 2],
                      rmat_t[:, 1, 2] + rmat_t[:, 2, 1], t2], -1)
    t2_rep = t2.repeat(4, 1).t()
    t3 = 1 + rmat_t[:, 0, 0] + rmat_t[:, 1, 1] + rmat_t[:, 2, 2]
    q3 = torch.stack([t3, rmat_t[:, 1, 2] - rmat_t[:, 2, 1], rmat_t[:, 2, 0] - rmat_t[:, 0, 2],
                      rmat_t[:, 0, 1] - rmat_t[:, 1, 0]], -1)
    t3_rep = t3.repeat(4, 1).t()
    mask_c0 = (mask_d2 & mask_d0_d1).view(-1, 1).type_as(q0)
    mask_c1 = (mask_d2 & ~mask_d0_d1).view(-1, 1).type_as(q1)
    mask_c2 = (~mask_d2 & mask_d0_nd1).view(-1, 1).type_as(q2)
    mask_c3 = (~mask_d2 & ~mask_d0_nd1).view(-1, 1).type_as(q3)
    q = q0 * mask_c0 + q1 * mask_c1 + q2 * mask_c2 + q3 * mask_c3
    q = q / torch.sqrt(t0_rep * mask_c0 + t1_rep * mask_c1 + t2_rep * mask_c2 + t3_rep * mask_c3)
    q = q * 0.5
    return q


def quaternion_to_angle_axis(quaternion):
    q1, q2, q3 = quaternion[..., 1], quaternion[..., 2], quaternion[..., 3]
    sin_squared_theta = q1 * q1 + q2 * q2 + q3 * q3
    sin_theta = torch.sqrt(sin_squared_theta)
    cos_theta = quaternion[..., 0]
    two_theta = 2.0 * torch.where(cos_theta < 0.0, torch.atan2(-sin_theta, -cos_theta),
                                  torch.atan2(sin_theta, cos_theta))
    k_pos = two_theta / sin_theta
    k_neg = 2.0 * torch.ones_like(sin_theta)
    k = torch.where(sin_squared_theta > 0.0, k_pos, k_neg)
    return torch.stack([q1 * k, q2 * k, q3 * k], dim=-1)


def rotation_matrix_to_angle_axis(rotation_matrix):
    """[N,3,4] (or [N,3,3]) -> [N,3]."""
    return quaternion_to_angle_axis(rotation_matrix_to_quaternion(rotation_matrix[:, :, :3]))


class ContinousRotReprDecoder(nn.Module):
    def forward(self, module_input):
        reshaped_input = module_input.view(-1, 3, 2)
        b1 = F.normalize(reshaped_input[:, :, 0], dim=1)
        dot_prod = torch.sum(b1 * reshaped_input[:, :, 1], dim=1, keepdim=True)
        b2 = F.normalize(reshaped_input[:, :, 1] - dot_prod * b1, dim=-1)
        b3 = torch.cross(b1, b2, dim=1)
        return torch.stack([b1, b2, b3], dim=-1)


class VPoser(nn.Module):
    def __init__(self, num_neurons=512, latentD=32, data_shape=(1, 21, 3), use_cont_repr=True):
        super().__init__()
        self.latentD = latentD
        self.use_cont_repr = use_cont_repr
        n_features = int(np.prod(data_shape))
        self.num_joints = data_shape[1]
        self.bodyprior_enc_bn1 = nn.BatchNorm1d(n_features)
        self.bodyprior_enc_fc1 = nn.Linear(n_features, num_neurons)
        self.bodyprior_enc_bn2 = nn.BatchNorm1d(num_neurons)
        self.bodyprior_enc_fc2 = nn.Linear(num_neurons, num_neurons)
        self.bodyprior_enc_mu = nn.Linear(num_neurons, latentD)
        self.bodyprior_enc_logvar = nn.Linear(num_neurons, latentD)
        self.dropout = nn.Dropout(p=.1, inplace=False)
        self.bodyprior_dec_fc1 = nn.Linear(latentD, num_neurons)
        self.bodyprior_dec_fc2 = nn.Linear(num_neurons, num_neurons)
        self.rot_decoder = ContinousRotReprDecoder()
        self.bodyprior_dec_out = nn.Linear(num_neurons, self.num_joints * 6)

    def encode(self, Pin):
        Xout = Pin.view(Pin.size(0), -1)
        Xout = self.bodyprior_enc_bn1(Xout)
        Xout = F.leaky_relu(self.bodyprior_enc_fc1(Xout), negative_slope=.2)
        Xout = self.bodyprior_enc_bn2(Xout)
        Xout = self.dropout(Xout)
        Xout = F.leaky_relu(self.bodyprior_enc_fc2(Xout), negative_slope=.2)
        return torch.distributions.normal.Normal(self.bodyprior_enc_mu(Xout),
                                                 F.softplus(self.bodyprior_enc_logvar(Xout)))

    def decode(self, Zin, output_type='matrot'):
        assert output_type in ['matrot', 'aa']
        Xout = F.leaky_relu(self.bodyprior_dec_fc1(Zin), negative_slope=.2)
        Xout = self.dropout(Xout)
        Xout = F.leaky_relu(self.bodyprior_dec_fc2(Xout), negative_slope=.2)
        Xout = self.bodyprior_dec_out(Xout)
        Xout = self.rot_decoder(Xout)
        Xout = Xout.view([-1, 1, self.num_joints, 9])
        if output_type == 'aa':
            return VPoser.matrot2aa(Xout)
        return Xout

    @staticmethod
    def matrot2aa(pose_matrot):
        batch_size = pose_matrot.size(0)
        homogen_matrot = F.pad(pose_matrot.view(-1, 3, 3), [0, 1])
        return rotation_matrix_to_angle_axis(homogen_matrot).view(batch_size, 1, -1, 3).contiguous()


def from_weights(w, dtype=torch.float32):
    """VPoser in eval mode from a flat dict of numpy arrays (``synthetic.make_vposer_like`` keys,
    also what ``smplifyx_b200.vposer.load_vposer_weights`` returns for a real checkpoint)."""
    m = VPoser()

    def put(lin, wk, bk):
        lin.weight.data = torch.tensor(w[wk], dtype=dtype)
        lin.bias.data = torch.tensor(w[bk], dtype=dtype)
    put(m.bodyprior_dec_fc1, 'dec_fc1_w', 'dec_fc1_b')
    put(m.bodyprior_dec_fc2, 'dec_fc2_w', 'dec_fc2_b')
    put(m.bodyprior_dec_out, 'dec_out_w', 'dec_out_b')
    put(m.bodyprior_enc_fc1, 'enc_fc1_w', 'enc_fc1_b')
    put(m.bodyprior_enc_fc2, 'enc_fc2_w', 'enc_fc2_b')
    put(m.bodyprior_enc_mu, 'enc_mu_w', 'enc_mu_b')
    put(m.bodyprior_enc_logvar, 'enc_logvar_w', 'enc_logvar_b')
    for bn, p in ((m.bodyprior_enc_bn1, 'enc_bn1'), (m.bodyprior_enc_bn2, 'enc_bn2')):
        bn.weight.data = torch.tensor(w[p + '_w'], dtype=dtype)
        bn.bias.data = torch.tensor(w[p + '_b'], dtype=dtype)
        bn.running_mean = torch.tensor(w[p + '_mean'], dtype=dtype)
        bn.running_var = torch.tensor(w[p + '_var'], dtype=dtype)
    m = m.to(dtype)
    m.eval()
    return m
