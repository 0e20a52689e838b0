"""VPoser v1.0 on the host side: checkpoint loading and the (off-hot-path) encoder.

The reference loads VPoser with ``load_vposer(vposer_ckpt, vp_model='snapshot')`` and calls
``vposer.encode(prior).sample()`` once per frame to initialise the latent pose and
``vposer.decode(z, output_type='aa')`` inside every closure (fit_single_frame.py:241-249,
fitting.py:236).  In this engine the decoder and its adjoint run inside the CUDA evaluation
kernel (csrc/sfx_core.cuh: vposer_decode / vposer_adjoint, weights installed with
``engine.Model.set_vposer``); this module only

* reads a VPoser v1 checkpoint (``<expr_dir>/snapshots/*.pt`` state dict) into the flat weight
  dict the engine takes (``load_vposer_weights``),
* provides the encoder (BN -> FC -> leaky-ReLU -> BN -> FC -> leaky-ReLU -> mu, softplus sigma)
  and a torch decoder for result post-processing (``VPoser``).
"""
import glob
import os

import numpy as np
import torch
import torch.nn.functional as F

_KEYS = {
    'dec_fc1': 'bodyprior_dec_fc1', 'dec_fc2': 'bodyprior_dec_fc2', 'dec_out': 'bodyprior_dec_out',
    'enc_fc1': 'bodyprior_enc_fc1', 'enc_fc2': 'bodyprior_enc_fc2', 'enc_mu': 'bodyprior_enc_mu',
    'enc_logvar': 'bodyprior_enc_logvar',
}


def load_vposer_weights(expr_dir):
    """-> dict of numpy arrays (dec_fc1_w, dec_fc1_b, ..., enc_bn1_mean, ...)."""
    snaps = sorted(glob.glob(os.path.join(expr_dir, 'snapshots', '*.pt')), key=os.path.getmtime)
    if not snaps:
        raise ValueError('No VPoser snapshot (*.pt) under {}/snapshots'.format(expr_dir))
    sd = torch.load(snaps[-1], map_location='cpu')
    sd = {k.replace('module.', ''): v for k, v in sd.items()}
    out = {}
    for short, long in _KEYS.items():
        out[short + '_w'] = sd[long + '.weight'].numpy()
        out[short + '_b'] = sd[long + '.bias'].numpy()
    for short, long in (('enc_bn1', 'bodyprior_enc_bn1'), ('enc_bn2', 'bodyprior_enc_bn2')):
        out[short + '_w'] = sd[long + '.weight'].numpy()
        out[short + '_b'] = sd[long + '.bias'].numpy()
        out[short + '_mean'] = sd[long + '.running_mean'].numpy()
        out[short + '_var'] = sd[long + '.running_var'].numpy()
    return out


def _rot6d_to_aa(x):
    """[N,6] continuous representation -> axis-angle [N,3] (Gram-Schmidt, then the quaternion
    route of torchgeometry 0.1.2 that VPoser.matrot2aa takes)."""
    m = x.reshape(-1, 3, 2)
    b1 = F.normalize(m[:, :, 0], dim=1)
    b2 = F.normalize(m[:, :, 1] - (b1 * m[:, :, 1]).sum(1, keepdim=True) * b1, dim=1)
    b3 = torch.cross(b1, b2, dim=1)
    r = torch.stack([b1, b2, b3], dim=1)           # rows b1, b2, b3 == transpose of the matrix
    r00, r01, r02, r10, r11, r12, r20, r21, r22 = [r[:, i, j] for i in range(3) for j in range(3)]
    d2, d01, d0n1 = r22 < 1e-6, r00 > r11, r00 < -r11
    t = torch.stack([1 + r00 - r11 - r22, 1 - r00 + r11 - r22, 1 - r00 - r11 + r22,
                     1 + r00 + r11 + r22], dim=1)
    q = torch.stack([
        torch.stack([r12 - r21, t[:, 0], r01 + r10, r20 + r02], -1),
        torch.stack([r20 - r02, r01 + r10, t[:, 1], r12 + r21], -1),
        torch.stack([r01 - r10, r20 + r02, r12 + r21, t[:, 2]], -1),
        torch.stack([t[:, 3], r12 - r21, r20 - r02, r01 - r10], -1)], dim=1)
    case = torch.where(d2, torch.where(d01, 0, 1), torch.where(d0n1, 2, 3))
    idx = torch.arange(x.shape[0])
    quat = 0.5 * q[idx, case] / torch.sqrt(t[idx, case]).unsqueeze(-1)
    s2 = (quat[:, 1:] ** 2).sum(-1)
    s = torch.sqrt(s2)
    c = quat[:, 0]
    two_theta = 2.0 * torch.where(c < 0, torch.atan2(-s, -c), torch.atan2(s, c))
    k = torch.where(s2 > 0, two_theta / s, torch.full_like(s, 2.0))
    return quat[:, 1:] * k.unsqueeze(-1)


class VPoser(torch.nn.Module):
    """Eval-mode VPoser v1 from a flat weight dict; ``encode`` / ``decode`` like the original."""

    def __init__(self, weights, dtype=torch.float32):
        super().__init__()
        self.weights = weights
        self.latentD = int(np.asarray(weights['dec_fc1_w']).shape[1])
        for k, v in weights.items():
            self.register_buffer(k, torch.tensor(np.asarray(v), dtype=dtype))

    def _bn(self, x, p):
        g = lambda n: getattr(self, p + n)
        return (x - g('_mean')) / torch.sqrt(g('_var') + 1e-5) * g('_w') + g('_b')

    def encode(self, pose):
        x = pose.reshape(pose.shape[0], -1)
        x = self._bn(x, 'enc_bn1')
        x = F.leaky_relu(F.linear(x, self.enc_fc1_w, self.enc_fc1_b), 0.2)
        x = self._bn(x, 'enc_bn2')
        x = F.leaky_relu(F.linear(x, self.enc_fc2_w, self.enc_fc2_b), 0.2)
        return torch.distributions.normal.Normal(
            F.linear(x, self.enc_mu_w, self.enc_mu_b),
            F.softplus(F.linear(x, self.enc_logvar_w, self.enc_logvar_b)))

    def decode(self, z, output_type='aa'):
        x = F.leaky_relu(F.linear(z, self.dec_fc1_w, self.dec_fc1_b), 0.2)
        x = F.leaky_relu(F.linear(x, self.dec_fc2_w, self.dec_fc2_b), 0.2)
        x = F.linear(x, self.dec_out_w, self.dec_out_b)
        if output_type != 'aa':
            raise ValueError("only output_type='aa' is provided (what the fitting path uses)")
        return _rot6d_to_aa(x.reshape(-1, 6)).reshape(z.shape[0], 1, -1, 3)


def load_vposer(expr_dir, vp_model='snapshot', dtype=torch.float32):
    """Same call as human_body_prior.tools.model_loader.load_vposer -> (module, settings)."""
    return VPoser(load_vposer_weights(expr_dir), dtype=dtype).eval(), None
