"""Priors with the reference's factory surface (smplifyx/prior.py:36-231):
``create_prior('l2' | 'angle' | 'gmm' | 'none', **kw)``.

The objects carry the constants; inside a fit the prior terms are evaluated by the CUDA
kernel (csrc/sfx_core.cuh, steps 10-11).  ``forward`` gives the same value in torch for
callers that evaluate a prior on their own.
"""
import os
import pickle
import sys

import numpy as np
import torch
import torch.nn as nn

DEFAULT_DTYPE = torch.float32


def create_prior(prior_type, **kwargs):
    if prior_type == 'gmm':
        return MaxMixturePrior(**kwargs)
    if prior_type == 'l2':
        return L2Prior(**kwargs)
    if prior_type == 'angle':
        return SMPLifyAnglePrior(**kwargs)
    if prior_type == 'none' or prior_type is None:
        return lambda *a, **k: 0.0
    raise ValueError('Prior {}'.format(prior_type) + ' is not implemented')


class L2Prior(nn.Module):
    kind = 'l2'

    def __init__(self, dtype=DEFAULT_DTYPE, reduction='sum', **kwargs):
        super(L2Prior, self).__init__()

    def forward(self, module_input, *args):
        return torch.sum(module_input.pow(2))


class SMPLifyAnglePrior(nn.Module):
    """exp(theta * sign)^2 on the 4 elbow / knee bending angles (full-pose entries 55, 58, 12,
    15; prior.py:53-89)."""
    kind = 'angle'

    def __init__(self, dtype=DEFAULT_DTYPE, **kwargs):
        super(SMPLifyAnglePrior, self).__init__()
        self.register_buffer('angle_prior_idxs',
                             torch.tensor([55, 58, 12, 15], dtype=torch.long))
        self.register_buffer('angle_prior_signs', torch.tensor([1, -1, -1, -1], dtype=dtype))

    def forward(self, pose, with_global_pose=False):
        idx = self.angle_prior_idxs - (not with_global_pose) * 3
        return torch.exp(pose[:, idx] * self.angle_prior_signs).pow(2)


class MaxMixturePrior(nn.Module):
    """min over M Gaussians of 0.5 (x-mu)^T P (x-mu) - log(w / (c sqrt|S| / min sqrt|S|))
    (prior.py:100-231).  Loads ``gmm_{M:02d}.pkl`` from ``prior_folder`` or takes the mixture
    as a dict (``gmm=`` keyword: means, covars, weights)."""
    kind = 'gmm'

    def __init__(self, prior_folder='prior', num_gaussians=6, dtype=DEFAULT_DTYPE, epsilon=1e-16,
                 use_merged=True, gmm=None, **kwargs):
        super(MaxMixturePrior, self).__init__()
        np_dtype = np.float64 if dtype == torch.float64 else np.float32
        self.num_gaussians = num_gaussians
        self.epsilon = epsilon
        self.use_merged = use_merged
        if gmm is None:
            fn = os.path.join(prior_folder, 'gmm_{:02d}.pkl'.format(num_gaussians))
            if not os.path.exists(fn):
                print('The path to the mixture prior "{}"'.format(fn) + ' does not exist, exiting!')
                sys.exit(-1)
            with open(fn, 'rb') as f:
                gmm = pickle.load(f, encoding='latin1')
        if isinstance(gmm, dict):
            means = np.asarray(gmm['means']).astype(np_dtype)
            covs = np.asarray(gmm['covars']).astype(np_dtype)
            weights = np.asarray(gmm['weights']).astype(np_dtype)
        else:                                  # sklearn GaussianMixture
            means = gmm.means_.astype(np_dtype)
            covs = gmm.covars_.astype(np_dtype)
            weights = gmm.weights_.astype(np_dtype)
        self.register_buffer('means', torch.tensor(means, dtype=dtype))
        self.register_buffer('covs', torch.tensor(covs, dtype=dtype))
        precisions = np.stack([np.linalg.inv(c) for c in covs]).astype(np_dtype)
        self.register_buffer('precisions', torch.tensor(precisions, dtype=dtype))
        sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in np.asarray(gmm['covars'])
                            ]) if isinstance(gmm, dict) else \
            np.array([np.sqrt(np.linalg.det(c)) for c in gmm.covars_])
        const = (2 * np.pi) ** (69 / 2.)
        nll_weights = np.asarray(np.asarray(gmm['weights'] if isinstance(gmm, dict)
                                            else gmm.weights_) / (const * (sqrdets / sqrdets.min())))
        self.register_buffer('nll_weights', torch.tensor(nll_weights, dtype=dtype).unsqueeze(dim=0))
        self.register_buffer('weights', torch.tensor(weights, dtype=dtype).unsqueeze(dim=0))
        self.random_var_dim = self.means.shape[1]

    def get_mean(self):
        return torch.matmul(self.weights, self.means)

    def forward(self, pose, betas=None):
        diff = pose.unsqueeze(dim=1) - self.means
        pd = torch.einsum('mij,bmj->bmi', [self.precisions, diff])
        quad = (pd * diff).sum(dim=-1)
        ll = 0.5 * quad - torch.log(self.nll_weights)
        return torch.min(ll, dim=1)[0]
