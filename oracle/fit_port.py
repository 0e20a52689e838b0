"""ORACLE (test infrastructure, not product code) -- CPU restatement ("port") of the
reference's per-frame fitting path, in torch with autograd exactly as the reference computes
it (op-by-op ATen + ``backward``), one frame at a time (reference asserts batch_size == 1,
smplifyx/fit_single_frame.py:119).

Restated, with the reference location each piece follows:

* ``gmof``                     smplifyx/utils.py:84-95
* ``project``                  smplifyx/camera.py:93-117
* ``camera_init_loss``         smplifyx/fitting.py:499-520 (incl. the double-unsqueeze broadcast)
* ``smplify_loss``             smplifyx/fitting.py:375-461 (no interpenetration term)
* ``angle_prior`` / ``gmm_prior``  smplifyx/prior.py:53-89, :100-196
* ``StrongWolfeLBFGS``         smplifyx/optimizers/lbfgs_ls.py:11-167 (line search), :256-445 (step)
* ``run_fitting``              smplifyx/fitting.py:147-217
* ``guess_init``               smplifyx/fitting.py:36-110
* ``fit_frame``                smplifyx/fit_single_frame.py:209-668 (camera stage, orientation
                               flip, stage annealing, result dict)
* ``euler_xyz_from_matrix``    smplifyx/utils.py:306-436 (intrinsic xyz only, the case the
                               reference uses)

The SMPL-X forward pass comes from ``oracle/smplx_shim.py`` (third-party ``smplx`` restated,
parity unpinned, see that file's header).

PINNING: ``tests/test_oracle_vs_reference.py`` runs this port against the unmodified
reference modules (imported through ``oracle/ref_bridge.py``) on the same model and inputs and
requires equal evaluation counts and matching losses / parameters; ``tests/golden/*.npz`` are
outputs of the reference itself (``tests/golden/make_golden.py``) that this port must
reproduce.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module; the product (``smplify-x-partial_b200/``) never does.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from . import smplx_shim


# ------------------------------------------------------------------------------ pieces
def gmof(residual, rho):
    sq = residual ** 2
    return rho ** 2 * torch.div(sq, sq + rho ** 2)


def project(points, rotation, translation, fx, fy, center):
    """points [B,N,3] -> [B,N,2]; rotation [B,3,3], translation [B,3], fx, fy [B], center [B,2]."""
    cam_tf = smplx_shim.transform_mat(rotation, translation.unsqueeze(dim=-1))
    ones = torch.ones(list(points.shape)[:-1] + [1], dtype=points.dtype)
    pts_h = torch.cat([points, ones], dim=-1)
    proj = torch.einsum('bki,bji->bjk', [cam_tf, pts_h])
    img = torch.div(proj[:, :, :2], proj[:, :, 2].unsqueeze(dim=-1))
    cam_mat = torch.zeros([points.shape[0], 2, 2], dtype=points.dtype)
    cam_mat[:, 0, 0] = fx
    cam_mat[:, 1, 1] = fy
    return torch.einsum('bki,bji->bjk', [cam_mat, img]) + center.unsqueeze(dim=1)


def angle_prior(body_pose):
    idx = torch.tensor([55, 58, 12, 15], dtype=torch.long) - 3
    signs = torch.tensor([1, -1, -1, -1], dtype=body_pose.dtype)
    return torch.exp(body_pose[:, idx] * signs).pow(2)


class GMMPrior(object):
    """MaxMixturePrior.merged_log_likelihood (prior.py:181-196) with the constants of
    prior.py:147-164 (note the hard-coded 69 in the 2*pi power)."""

    def __init__(self, gmm, dtype=torch.float32):
        means = np.asarray(gmm['means'])
        covs = np.asarray(gmm['covars'])
        np_dtype = np.float32 if dtype == torch.float32 else np.float64
        self.means = torch.tensor(means.astype(np_dtype), dtype=dtype)
        prec = np.stack([np.linalg.inv(c) for c in covs.astype(np_dtype)]).astype(np_dtype)
        self.precisions = torch.tensor(prec, dtype=dtype)
        sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in covs])
        const = (2 * np.pi) ** (69 / 2.)
        nll = np.asarray(gmm['weights'] / (const * (sqrdets / sqrdets.min())))
        self.nll_weights = torch.tensor(nll, dtype=dtype).unsqueeze(dim=0)
        self.weights = torch.tensor(np.asarray(gmm['weights']), dtype=dtype).unsqueeze(dim=0)

    def get_mean(self):
        return torch.matmul(self.weights, self.means)

    def __call__(self, pose, betas=None):
        diff = pose.unsqueeze(dim=1) - self.means
        pd = torch.einsum('mij,bmj->bmi', [self.precisions, diff])
        quad = (pd * diff).sum(dim=-1)
        ll = 0.5 * quad - torch.log(self.nll_weights)
        return torch.min(ll, dim=1)[0]


def euler_xyz_from_matrix(R):
    """Intrinsic x-y-z Euler angles of rotation matrices [N,3,3] (or [3,3]) -> [N,3], in the
    dtype of the input, restating the (seq='xyz', extrinsic=False) case of
    utils._compute_euler_from_matrix (utils.py:306-436) with the same torch operations."""
    R = torch.as_tensor(R)
    if R.dim() == 2:
        R = R[None]
    dt = R.dtype
    n1 = torch.tensor([1., 0., 0.])
    n2 = torch.tensor([0., 1., 0.])
    n3 = torch.tensor([0., 0., 1.])
    n12 = torch.linalg.cross(n1, n2)
    sl = torch.dot(n12, n3)
    cl = torch.dot(n1, n3)
    offset = torch.atan2(sl, cl)
    c = torch.stack((n2, n12, n1)).type(dt)
    rot = torch.tensor([[1, 0, 0], [0, cl, sl], [0, -sl, cl]]).type(dt)
    O = torch.einsum('...ij,jk->...ik', torch.einsum('ij,...jk->...ik', c, R),
                     torch.transpose(c, 0, 1) @ rot)
    ang = torch.zeros((R.shape[0], 3), dtype=dt)
    O[O[:, 2, 2] > 1, 2, 2] = 1
    O[O[:, 2, 2] < -1, 2, 2] = -1
    ang[:, 1] = torch.acos(O[:, 2, 2])
    eps = 1e-7
    safe1 = torch.abs(ang[:, 1]) >= eps
    safe2 = torch.abs(ang[:, 1] - np.pi) >= eps
    ang[:, 1] += offset
    safe = safe1 & safe2
    ang[safe, 0] = torch.atan2(O[safe, 0, 2], -O[safe, 1, 2])
    ang[safe, 2] = torch.atan2(O[safe, 2, 0], O[safe, 2, 1])
    ang[~safe, 2] = 0
    ang[~safe1, 0] = torch.atan2(O[~safe1, 1, 0] - O[~safe1, 0, 1],
                                 O[~safe1, 0, 0] + O[~safe1, 1, 1])
    ang[~safe2, 0] = torch.atan2(O[~safe2, 1, 0] + O[~safe2, 0, 1],
                                 O[~safe2, 0, 0] - O[~safe2, 1, 1])
    adjust = ((ang[:, 1] < -np.pi / 2) | (ang[:, 1] > np.pi / 2)) & safe
    ang[adjust, 0] += np.pi
    ang[adjust, 1] = 2 * offset - ang[adjust, 1]
    ang[adjust, 2] -= np.pi
    ang[ang < -np.pi] += 2 * np.pi
    ang[ang > np.pi] -= 2 * np.pi
    return ang


# --------------------------------------------------------------------- L-BFGS (strong Wolfe)
class TorchOps(object):
    """Vector primitives of the optimiser exactly as the reference spells them (ATen dot,
    add_ with alpha).  Tests may substitute an implementation with another summation order
    to replay the engine's arithmetic bit for bit (tests/test_control_flow.py)."""

    @staticmethod
    def dot(a, b):
        return a.dot(b)

    @staticmethod
    def absmax(a):
        return a.abs().max()

    @staticmethod
    def abssum(a):
        return a.abs().sum()

    @staticmethod
    def axpy_(x, alpha, y):          # x += alpha * y
        return x.add_(y, alpha=float(alpha))


def _cubic_min(x1, f1, g1, x2, f2, g2, bounds=None):
    lo, hi = bounds if bounds is not None else ((x1, x2) if x1 <= x2 else (x2, x1))
    d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)
    disc = d1 ** 2 - g1 * g2
    if disc >= 0:
        d2 = disc.sqrt()
        if x1 <= x2:
            pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))
        else:
            pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2 * d2))
        return min(max(pos, lo), hi)
    return (lo + hi) / 2.


def _strong_wolfe(phi, t, d, f, g, gtd, c1=1e-4, c2=0.9, tol_change=1e-9, max_iter=20, max_ls=25,
                  ops=TorchOps):
    """phi(t) -> (loss float, flat grad).  Returns (f, g, t, n_evals)."""
    d_norm = ops.absmax(d)
    g = g.clone()
    f_new, g_new = phi(t)
    n_evals = 1
    gtd_new = ops.dot(g_new, d)
    t_prev, f_prev, g_prev, gtd_prev = 0, f, g, gtd
    done = False
    it = 0
    br = None
    while it < max_ls:
        if f_new > (f + c1 * t * gtd) or (it > 1 and f_new >= f_prev):
            br = ([t_prev, t], [f_prev, f_new], [g_prev, g_new.clone()], [gtd_prev, gtd_new])
            break
        if abs(gtd_new) <= -c2 * gtd:
            br = ([t], [f_new], [g_new], [gtd_new])
            done = True
            break
        if gtd_new >= 0:
            br = ([t_prev, t], [f_prev, f_new], [g_prev, g_new.clone()], [gtd_prev, gtd_new])
            break
        lo = t + 0.01 * (t - t_prev)
        hi = t * 10
        t_old = t
        t = _cubic_min(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, bounds=(lo, hi))
        t_prev, f_prev, g_prev, gtd_prev = t_old, f_new, g_new.clone(), gtd_new
        f_new, g_new = phi(t)
        n_evals += 1
        gtd_new = ops.dot(g_new, d)
        it += 1
    if it == max_ls:
        br = ([0, t], [f, f_new], [g, g_new], [gtd, gtd_new])
    bt, bf, bg, bgtd = br
    stalled = False
    lo_i, hi_i = (0, 1) if bf[0] <= bf[-1] else (1, 0)
    while not done and it < max_iter:
        t = _cubic_min(bt[0], bf[0], bgtd[0], bt[1], bf[1], bgtd[1])
        width = max(bt) - min(bt)
        eps = 0.1 * width
        if min(max(bt) - t, t - min(bt)) < eps:
            if stalled or t >= max(bt) or t <= min(bt):
                if abs(t - max(bt)) < abs(t - min(bt)):
                    t = max(bt) - eps
                else:
                    t = min(bt) + eps
                stalled = False
            else:
                stalled = True
        else:
            stalled = False
        f_new, g_new = phi(t)
        n_evals += 1
        gtd_new = ops.dot(g_new, d)
        it += 1
        if f_new > (f + c1 * t * gtd) or f_new >= bf[lo_i]:
            bt[hi_i], bf[hi_i], bg[hi_i], bgtd[hi_i] = t, f_new, g_new.clone(), gtd_new
            lo_i, hi_i = (0, 1) if bf[0] <= bf[1] else (1, 0)
        else:
            if abs(gtd_new) <= -c2 * gtd:
                done = True
            elif gtd_new * (bt[hi_i] - bt[lo_i]) >= 0:
                bt[hi_i], bf[hi_i], bg[hi_i], bgtd[hi_i] = bt[lo_i], bf[lo_i], bg[lo_i], bgtd[lo_i]
            bt[lo_i], bf[lo_i], bg[lo_i], bgtd[lo_i] = t, f_new, g_new.clone(), gtd_new
        if abs(bt[1] - bt[0]) * d_norm < tol_change:
            break
    return bf[lo_i], bg[lo_i], bt[lo_i], n_evals


class StrongWolfeLBFGS(object):
    """L-BFGS over one flat vector.  ``closure()`` evaluates the objective at ``self.x`` and
    returns ``(loss_tensor, flat_grad)``; state persists across ``step`` calls like the
    reference optimiser's."""

    def __init__(self, x, lr=1.0, max_iter=20, max_eval=None, tol_grad=1e-5, tol_change=1e-9,
                 history=100, ops=TorchOps):
        self.x = x
        self.ops = ops
        self.lr = lr
        self.max_iter = max_iter
        self.max_eval = max_eval if max_eval is not None else max_iter * 5 // 4
        self.tol_grad = tol_grad
        self.tol_change = tol_change
        self.history = history
        self.n_iter = 0
        self.func_evals = 0
        self.d = self.t = self.H_diag = self.prev_g = self.prev_loss = None
        self.Y, self.S, self.rho = [], [], []
        self.al = [None] * history

    def _probe(self, closure, x0, t, d):
        with torch.no_grad():
            self.ops.axpy_(self.x, t, d)
        loss, g = closure()
        loss = float(loss)
        with torch.no_grad():
            self.x.copy_(x0)
        return loss, g

    def step(self, closure):
        ops = self.ops
        orig_loss, g = closure()
        loss = float(orig_loss)
        evals = 1
        self.func_evals += 1
        if ops.absmax(g) <= self.tol_grad:
            return orig_loss
        d, t, H_diag, prev_g, prev_loss = self.d, self.t, self.H_diag, self.prev_g, self.prev_loss
        it = 0
        while it < self.max_iter:
            it += 1
            self.n_iter += 1
            if self.n_iter == 1:
                d = g.neg()
                self.Y, self.S, self.rho = [], [], []
                H_diag = 1
            else:
                y = g.sub(prev_g)
                s = d.mul(t)
                ys = ops.dot(y, s)
                if ys > 1e-10:
                    if len(self.Y) == self.history:
                        self.Y.pop(0)
                        self.S.pop(0)
                        self.rho.pop(0)
                    self.Y.append(y)
                    self.S.append(s)
                    self.rho.append(1. / ys)
                    H_diag = ys / ops.dot(y, y)
                k = len(self.Y)
                q = g.neg()
                for i in range(k - 1, -1, -1):
                    self.al[i] = ops.dot(self.S[i], q) * self.rho[i]
                    ops.axpy_(q, -self.al[i], self.Y[i])
                d = r = torch.mul(q, H_diag)
                for i in range(k):
                    be = ops.dot(self.Y[i], r) * self.rho[i]
                    ops.axpy_(r, self.al[i] - be, self.S[i])
            prev_g = g.clone() if prev_g is None else prev_g.copy_(g)
            prev_loss = loss
            if self.n_iter == 1:
                t = min(1., 1. / ops.abssum(g)) * self.lr
            else:
                t = self.lr
            gtd = ops.dot(g, d)
            if gtd > -self.tol_change:
                break
            x0 = self.x.detach().clone()
            loss, g, t, ls_evals = _strong_wolfe(
                lambda tt: self._probe(closure, x0, tt, d), t, d, loss, g, gtd,
                max_iter=self.max_iter, ops=ops)
            with torch.no_grad():
                ops.axpy_(self.x, t, d)
            opt = ops.absmax(g) <= self.tol_grad
            evals += ls_evals
            self.func_evals += ls_evals
            if it == self.max_iter:
                break
            if evals >= self.max_eval:
                break
            if opt:
                break
            if ops.absmax(d.mul(t)) <= self.tol_change:
                break
            if abs(loss - prev_loss) < self.tol_change:
                break
        self.d, self.t, self.H_diag, self.prev_g, self.prev_loss = d, t, H_diag, prev_g, prev_loss
        return orig_loss


class AdamPort(object):
    """torch.optim.Adam semantics (reference optim_factory.py:45-48) over the flat vector."""

    def __init__(self, x, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        self.x = x
        self.opt = torch.optim.Adam([x], lr=lr, betas=(beta1, beta2), eps=eps)

    def step(self, closure):
        loss, g = closure()
        self.x.grad = g.clone()
        self.opt.step()
        return loss


def rel_change(prev, cur):
    return (prev - cur) / max([abs(prev), abs(cur), 1])


def run_fitting(opt, closure, blocks, last_grad, maxiters=30, ftol=1e-9, gtol=1e-9):
    """fitting.py:147-217.  ``blocks``: list of (start, stop) slices of the flat vector, one per
    live torch parameter; ``last_grad()`` returns the gradient of the latest closure call."""
    prev = None
    for n in range(maxiters):
        loss = opt.step(closure)
        lv = float(loss)
        if math.isnan(lv) or math.isinf(lv):
            break
        if n > 0 and prev is not None and ftol > 0:
            if rel_change(prev, lv) <= ftol:
                break
        g = last_grad()
        if all(abs(float(g[a:b].max())) < gtol for a, b in blocks):
            break
        prev = lv
    return prev


# ------------------------------------------------------------------------------ frame problem
BODY_PARAM_ORDER = ['betas', 'global_orient', 'body_pose', 'left_hand_pose', 'right_hand_pose',
                    'jaw_pose', 'leye_pose', 'reye_pose', 'expression']


class FrameProblem(object):
    """All state of one frame: model parameters (named like the smplx module's), camera,
    targets; builds flat-vector closures for the camera stage and the body stages."""

    def __init__(self, body_model, keypoints, H, W, focal_length, joint_weights,
                 dtype=torch.float32, use_vposer=False, vposer=None, rho=100.,
                 body_prior=None, confidence_threshold=0.0, num_body_joints=25,
                 use_joints_conf=True, camera_rotation=None):
        self.bm = body_model
        self.dtype = dtype
        self.H, self.W, self.focal = H, W, focal_length
        kp = torch.tensor(np.asarray(keypoints), dtype=dtype).reshape(1, -1, 3)
        self.gt = kp[:, :, :2].contiguous()
        self.conf = kp[:, :, 2].reshape(1, -1).contiguous()
        self.jw = torch.as_tensor(joint_weights, dtype=dtype).reshape(1, -1).clone()
        thr = np.array([confidence_threshold] * num_body_joints + [0] * 42 + [0] * 68)
        K = self.conf.shape[1]
        self.low_conf = [i for i in range(K) if float(self.conf[0, i]) < thr[i]]
        self.jw[:, self.low_conf] = 0
        self.nb = num_body_joints
        self.rho = rho
        self.use_vposer = use_vposer
        self.vposer = vposer
        self.body_prior = body_prior
        self.use_joints_conf = use_joints_conf
        self.P = OrderedDict()
        for name in BODY_PARAM_ORDER:
            if hasattr(body_model, name) and getattr(body_model, name) is not None:
                self.P[name] = torch.zeros_like(getattr(body_model, name).detach())
        self.pose_embedding = None
        self.cam_R = (torch.eye(3, dtype=dtype).unsqueeze(0) if camera_rotation is None
                      else torch.as_tensor(camera_rotation, dtype=dtype).reshape(1, 3, 3))
        self.cam_t = torch.zeros([1, 3], dtype=dtype)
        self.center = torch.zeros([1, 2], dtype=dtype)
        self.fx = torch.full([1], focal_length, dtype=dtype)
        self.fy = torch.full([1], focal_length, dtype=dtype)
        self.n_evals = 0
        self._last_grad = None

    # parameter bookkeeping ------------------------------------------------------------
    def reset_params(self, **kw):
        for name in self.P:
            if name in kw:
                self.P[name] = torch.as_tensor(kw[name], dtype=self.dtype).detach().reshape(
                    self.P[name].shape).clone()
            else:
                self.P[name] = torch.zeros_like(self.P[name])

    def body_pose_from(self, emb):
        if self.use_vposer:
            return self.vposer.decode(emb, output_type='aa').view(1, -1)
        return emb.reshape(1, -1)

    def forward(self, params, emb, return_verts=False):
        kw = {k: v for k, v in params.items() if k != 'body_pose'}
        return_verts = return_verts or getattr(self, 'coll', None) is not None
        return self.bm(body_pose=self.body_pose_from(emb), return_verts=return_verts,
                       return_full_pose=True, **kw)

    # flat vector <-> named parameters --------------------------------------------------
    def _layout(self, names):
        off = 0
        lay = []
        for n in names:
            numel = (self.pose_embedding if n == 'pose_embedding' else
                     self.cam_t if n == 'camera_translation' else self.P[n]).numel()
            lay.append((n, off, off + numel))
            off += numel
        return lay, off

    def _gather(self, names):
        vals = []
        for n in names:
            v = (self.pose_embedding if n == 'pose_embedding' else
                 self.cam_t if n == 'camera_translation' else self.P[n])
            vals.append(v.reshape(-1))
        return torch.cat(vals).detach().clone()

    def _scatter(self, names, x):
        lay, _ = self._layout(names)
        for n, a, b in lay:
            v = x[a:b].detach().clone()
            if n == 'pose_embedding':
                self.pose_embedding = v.reshape(self.pose_embedding.shape)
            elif n == 'camera_translation':
                self.cam_t = v.reshape(1, 3)
            else:
                self.P[n] = v.reshape(self.P[n].shape)

    def make_closure(self, names, loss_fn):
        """Returns (x, closure, blocks).  Gradients of names that do not reach the loss
        (the dead ``body_pose`` block, fit_single_frame.py:554-559) stay zero and the block is
        left out of ``blocks`` (``var.grad is None`` in fitting.py:191-192)."""
        lay, n = self._layout(names)
        x = self._gather(names).requires_grad_(True)

        def closure():
            if x.grad is not None:
                x.grad = None
            params = dict(self.P)
            emb = self.pose_embedding
            cam_t = self.cam_t
            for nme, a, b in lay:
                if nme == 'pose_embedding':
                    emb = x[a:b].reshape(self.pose_embedding.shape)
                elif nme == 'camera_translation':
                    cam_t = x[a:b].reshape(1, 3)
                elif nme != 'body_pose':
                    params[nme] = x[a:b].reshape(self.P[nme].shape)
            loss = loss_fn(params, emb, cam_t)
            loss.backward()
            self.n_evals += 1
            g = x.grad.detach().clone()
            self._last_grad = g
            return loss.detach(), g
        blocks = [(a, b) for nme, a, b in lay if nme != 'body_pose']
        return x, closure, blocks

    # losses ---------------------------------------------------------------------------
    def camera_init_loss(self, params, emb, cam_t, init_idxs, trans_est, data_weight,
                         depth_loss_weight, use_conf):
        """Weights are 0-dim tensors of the working dtype, squared in that dtype, as the
        reference's registered buffers are (fitting.py:480-486, :490-497)."""
        dw = torch.tensor(data_weight, dtype=self.dtype)
        dlw = torch.tensor(depth_loss_weight, dtype=self.dtype)
        out = self.forward(params, emb)
        proj = project(out.joints, self.cam_R, cam_t, self.fx, self.fy, self.center)
        idx = torch.as_tensor(init_idxs, dtype=torch.long)
        err = torch.pow(torch.index_select(self.gt, 1, idx) -
                        torch.index_select(proj, 1, idx), 2)
        if use_conf:
            c = torch.index_select(self.conf, 1, idx).unsqueeze(2)
            jl = torch.sum(err * (c.unsqueeze(2) ** 2)) * dw ** 2
        else:
            jl = torch.sum(err) * dw ** 2
        dl = 0.0
        if dlw.item() > 0 and trans_est is not None:
            dl = dlw ** 2 * torch.sum((cam_t[:, 2] - trans_est[:, 2]).pow(2))
        return jl + dl

    def smplify_loss(self, params, emb, cam_t, w, jw, stage, num_stages, regression_pose):
        """``w``: dict of 0-dim tensors (jaw: 3-vector) of the working dtype; terms are summed
        in the order of fitting.py:457-460."""
        out = self.forward(params, emb)
        proj = project(out.joints, self.cam_R, cam_t, self.fx, self.fy, self.center)
        weights = (jw * self.conf if self.use_joints_conf else jw).unsqueeze(dim=-1)
        jd = gmof(self.gt - proj, self.rho)
        joint_loss = torch.sum(weights ** 2 * jd) * w['data_weight'] ** 2
        bpw = w['body_pose_weight']
        if self.use_vposer:
            if stage + 1 == num_stages and regression_pose is not None:
                pprior = (emb - regression_pose).pow(2).sum() * bpw ** 2
            else:
                pprior = emb.pow(2).sum() * bpw ** 2
        elif regression_pose is not None:
            pprior = ((emb - regression_pose).pow(2).sum()) * bpw ** 2
        else:
            pprior = torch.sum(self.body_prior(out.body_pose, out.betas)) * bpw ** 2
        shape_loss = torch.sum(torch.sum(out.betas.pow(2))) * w['shape_weight'] ** 2
        angle_loss = torch.sum(angle_prior(out.full_pose[:, 3:66])) * w['bending_prior_weight']
        lh = torch.sum(torch.sum(out.left_hand_pose.pow(2))) * w['hand_prior_weight'] ** 2
        rh = torch.sum(torch.sum(out.right_hand_pose.pow(2))) * w['hand_prior_weight'] ** 2
        expr = torch.sum(torch.sum(out.expression.pow(2))) * w['expr_prior_weight'] ** 2
        jaw = torch.sum(torch.sum(out.jaw_pose.mul(w['jaw_prior_weight']).pow(2)))
        pen = 0.0
        coll = getattr(self, 'coll', None)
        if coll is not None and float(w.get('coll_loss_weight', 0.0)) > 0:
            # fitting.py:437-455: search (no gradient), part filter, conic penalty
            search_tree, pen_distance, filter_faces, faces = coll
            tri = torch.index_select(out.vertices, 1, faces).view(1, -1, 3, 3)
            with torch.no_grad():
                idxs = search_tree(tri)
            if filter_faces is not None:
                idxs = filter_faces(idxs)
            if idxs.ge(0).sum().item() > 0:
                pen = torch.sum(w['coll_loss_weight'] * pen_distance(tri, idxs))
        return (joint_loss + pprior + shape_loss + angle_loss + pen + jaw + expr + lh + rh)


def l2_body_prior(pose, betas):
    return torch.sum(pose.pow(2))


def guess_init(prob, edge_idxs, focal_length):
    with torch.no_grad():
        out = prob.forward(prob.P, prob.pose_embedding)
        j3 = out.joints
        d3 = torch.stack([j3[:, e[0]] - j3[:, e[1]] for e in edge_idxs], dim=1)
        d2 = torch.stack([prob.gt[:, e[0]] - prob.gt[:, e[1]] for e in edge_idxs], dim=1)
        l2 = d2.pow(2).sum(dim=-1).sqrt()
        l3 = d3.pow(2).sum(dim=-1).sqrt()
        est = focal_length * (l3.mean(dim=1) / l2.mean(dim=1))
        z = torch.zeros([1], dtype=prob.dtype)
        return torch.stack([z, z.clone(), est], dim=1)


def rodrigues_np(r):
    r = np.asarray(r, dtype=np.float64).reshape(3)
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)


def inv_rodrigues_np(R):
    """Rotation matrix -> axis-angle (what cv2.Rodrigues does for a 3x3 input)."""
    R = np.asarray(R, dtype=np.float64)
    rx, ry, rz = R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]
    s = math.sqrt((rx * rx + ry * ry + rz * rz) * 0.25)
    c = (np.trace(R) - 1) * 0.5
    c = min(1.0, max(-1.0, c))
    th = math.acos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = (R[0, 0] + 1) * 0.5
        x = math.sqrt(max(t, 0.0))
        t = (R[1, 1] + 1) * 0.5
        y = math.sqrt(max(t, 0.0)) * (-1.0 if R[0, 1] < 0 else 1.0)
        t = (R[2, 2] + 1) * 0.5
        z = math.sqrt(max(t, 0.0)) * (-1.0 if R[0, 2] < 0 else 1.0)
        if abs(x) < abs(y) and abs(x) < abs(z) and (R[1, 2] > 0) != (y * z > 0):
            z = -z
        v = np.array([x, y, z])
        return v * (th / np.linalg.norm(v))
    vth = 1.0 / (2 * s) * th
    return np.array([rx, ry, rz]) * vth


def fit_frame(body_model, keypoints, H, W, cfg, joint_weights, expose=None, pixie=None,
              vposer=None, body_prior=None, dtype=torch.float32, return_verts=True):
    """One frame through the whole reference flow.  ``cfg``: the flat config dict
    (cmd_parser keys).  Returns dict(result=<pkl dict>, loss, vertices, joints, n_evals,
    stage_losses, cam_loss)."""
    fmt = cfg.get('format', 'coco25')
    nb = {'coco25': 25, 'halpe': 26, 'coco_wholebody': 23}[fmt]
    focal = cfg.get('focal_length') or (W ** 2 + H ** 2) ** 0.5
    use_vposer = bool(cfg.get('use_vposer', False))
    regression_prior = cfg.get('regression_prior')
    prob = FrameProblem(body_model, keypoints, H, W, focal, joint_weights, dtype=dtype,
                        use_vposer=use_vposer, vposer=vposer, rho=cfg.get('rho', 100),
                        body_prior=body_prior or l2_body_prior,
                        confidence_threshold=cfg.get('confidence_threshold', 0),
                        num_body_joints=nb, use_joints_conf=cfg.get('use_joints_conf', True))
    bpw = list(cfg['body_pose_prior_weights'])
    S = len(bpw)
    jaw_w = cfg.get('jaw_pose_prior_weights')
    if jaw_w is None:
        jaw_w = [[x] * 3 for x in cfg['shape_weights']]
    else:
        jaw_w = [[float(v) for v in s.split(',')] if isinstance(s, str) else list(s) for s in jaw_w]
    stages = []
    for i in range(S):
        stages.append(dict(body_pose_weight=bpw[i], shape_weight=cfg['shape_weights'][i],
                           expr_prior_weight=cfg['expr_weights'][i],
                           hand_prior_weight=cfg['hand_pose_prior_weights'][i],
                           jaw_prior_weight=jaw_w[i], hand_weight=cfg['hand_joints_weights'][i],
                           face_weight=cfg['face_joints_weights'][i]))

    # --- regression prior -> initial pose (fit_single_frame.py:209-235) ---
    global_pose = None
    full_prior = None
    if regression_prior:
        def eul(mats):
            return [euler_xyz_from_matrix(torch.tensor(np.asarray(m))) for m in mats]
        if regression_prior in ('PIXIE', 'combined'):
            pix = eul(pixie['body_pose'])
            global_pose = eul([pixie['global_pose']])[0]
        if regression_prior in ('ExPose', 'combined'):
            exp = eul(expose['body_pose'])
            global_pose = eul([expose['global_orient']])[0]
        full_prior = {'PIXIE': lambda: pix, 'ExPose': lambda: exp,
                      'combined': lambda: exp[:19] + pix[19:]}[regression_prior]()
        full_prior = torch.cat(full_prior).reshape(1, -1).to(dtype)
    if use_vposer:
        if regression_prior:
            prob.pose_embedding = vposer.encode(full_prior.reshape(1, -1)).sample().detach().clone()
        else:
            prob.pose_embedding = torch.zeros([1, 32], dtype=dtype)
    else:
        # body_pose_prior.get_mean() (fit_single_frame.py:250-252).  Only MaxMixturePrior has one
        # (prior.py:176-179): with body_prior_type 'l2' the reference raises AttributeError at this
        # line, so an un-initialised L2 fit cannot run there at all.  The port (and the engine,
        # fit_frames.FitPlan) start such a fit from the L2 prior's own mean, the zero pose.
        if regression_prior:
            prob.pose_embedding = full_prior.clone()
        elif hasattr(body_prior, 'get_mean'):
            prob.pose_embedding = body_prior.get_mean().detach().clone().to(dtype)
        else:
            prob.pose_embedding = torch.zeros([1, 63], dtype=dtype)
    if regression_prior:
        prob.reset_params(global_orient=global_pose, body_pose=prob.pose_embedding)
    else:
        prob.reset_params(body_pose=prob.pose_embedding)
    regression_pose = prob.pose_embedding.clone() if regression_prior else None

    init_idxs = [i for i in cfg['init_joints_idxs']
                 if float(prob.gt[0, i, 0]) != 0 and float(prob.gt[0, i, 1]) != 0
                 and i not in prob.low_conf]

    # --- camera initialisation (fit_single_frame.py:359-411) ---
    if cfg.get('use_camera_prior') and regression_prior in ('ExPose', 'combined'):
        cx, cy = [float(v) for v in expose['center']]
        tr = np.array(expose['transl'], dtype=np.float64).copy()
        tr[-1] /= (5000 / focal)
        init_t = torch.tensor(tr, dtype=dtype).reshape(1, -1)
        prob.center = torch.tensor([[cx, cy]], dtype=dtype)
    elif cfg.get('use_camera_prior') and regression_prior == 'PIXIE':
        left, top, right, bottom = [float(v) for v in pixie['bbox']]
        old = max(right - left, bottom - top)
        cen = np.array([right - (right - left) / 2.0, bottom - (bottom - top) / 2.0])
        size = int(old * 1.1)
        cam = pixie['body_cam']
        init_t = torch.tensor([cam[1], cam[2], 2 * focal / (cam[0] * size + 1e-9)],
                              dtype=dtype).reshape(1, -1)
        prob.center = torch.tensor([[cen[0], cen[1]]], dtype=dtype)
    else:
        init_t = guess_init(prob, cfg['body_tri_idxs'], focal).reshape(1, -1)
        prob.center = torch.tensor([[W, H]], dtype=dtype) * 0.5
    prob.cam_t = init_t.clone()
    data_weight = 1000 / H

    opt_kw = dict(lr=cfg.get('lr', 1.0), max_iter=cfg.get('maxiters', 30))

    def make_opt(x):
        kind = cfg.get('optim_type', 'lbfgsls')
        if kind == 'lbfgsls':
            return StrongWolfeLBFGS(x, **opt_kw)
        if kind == 'adam':
            return AdamPort(x, lr=cfg.get('lr', 1e-3))
        raise ValueError(kind)

    fit_kw = dict(maxiters=cfg.get('maxiters', 30), ftol=cfg.get('ftol', 1e-9),
                  gtol=cfg.get('gtol', 1e-9))

    # --- stage C: camera translation + global orientation ---
    names_c = ['camera_translation', 'global_orient']
    x, closure, blocks = prob.make_closure(
        names_c, lambda p, e, t: prob.camera_init_loss(
            p, e, t, init_idxs, init_t, data_weight, cfg.get('depth_loss_weight', 1e2),
            cfg.get('use_conf_for_camera_init', False)))
    opt = make_opt(x)
    cam_loss = run_fitting(opt, closure, blocks, lambda: prob._last_grad, **fit_kw)
    prob._scatter(names_c, x)
    evals_cam = prob.n_evals

    shoulder = torch.dist(prob.gt[:, cfg.get('left_shoulder_idx', 2)],
                          prob.gt[:, cfg.get('right_shoulder_idx', 5)])
    go = prob.P['global_orient'].detach().cpu().numpy()
    orients = [go]
    if float(shoulder) < cfg.get('side_view_thsh', 25.):
        flipped = rodrigues_np(go.reshape(3)).dot(rodrigues_np([0., np.pi, 0.]))
        orients.append(inv_rodrigues_np(flipped).reshape(1, 3).astype(np.float32))

    names_b = [n for n in prob.P] + ['pose_embedding']
    results = []
    out = None
    for orient in orients:
        prob.reset_params(global_orient=orient, body_pose=prob.pose_embedding)
        stage_losses = []
        for si, sw in enumerate(stages):
            w = {k: torch.tensor(v, dtype=dtype) for k, v in sw.items()}
            w['data_weight'] = torch.tensor(data_weight, dtype=dtype)
            w['bending_prior_weight'] = 3.17 * w['body_pose_weight']
            prob.jw[:, nb:nb + 42] = w['hand_weight']
            prob.jw[:, nb + 42:] = w['face_weight']
            prob.jw[:, prob.low_conf] = 0
            jw = prob.jw.clone()
            x, closure, blocks = prob.make_closure(
                names_b, lambda p, e, t, w=w, jw=jw, si=si: prob.smplify_loss(
                    p, e, t, w, jw, si, S, regression_pose))
            opt = make_opt(x)
            final_loss = run_fitting(opt, closure, blocks, lambda: prob._last_grad, **fit_kw)
            prob._scatter(names_b, x)
            stage_losses.append(final_loss)
        with torch.no_grad():
            out = prob.forward(prob.P, prob.pose_embedding, return_verts=return_verts)
        res = {'camera_rotation': prob.cam_R.numpy().copy(),
               'camera_translation': prob.cam_t.numpy().copy(),
               'camera_center': prob.center.numpy().copy(),
               'H': H, 'W': W, 'focal_length': focal}
        for n in prob.P:
            res[n] = prob.P[n].numpy().copy()
        res['body_pose'] = prob.body_pose_from(prob.pose_embedding).detach().numpy().copy()
        results.append(dict(loss=final_loss, result=res, stage_losses=stage_losses))
    best = 0 if len(results) == 1 or results[0]['loss'] < results[1]['loss'] else 1
    return dict(result=results[best]['result'], loss=results[best]['loss'],
                stage_losses=results[best]['stage_losses'], all_results=results,
                vertices=None if out.vertices is None else out.vertices.numpy().copy(),
                joints=out.joints.numpy().copy(), n_evals=prob.n_evals, evals_cam=evals_cam,
                cam_loss=cam_loss, n_orient=len(orients))


def time_coll_closure(body_model, keypoints, H, W, cfg, joint_weights, expose, pixie, part_segm,
                      n_evals=3):
    """Bounded CPU sample for the interpenetration workload (bench.py): seconds per closure
    evaluation (forward + SMPLifyLoss + backward, fitting.py:232-273) of one frame at its
    regression-prior start, second annealing stage, with the interpenetration term
    (fitting.py:437-455 driving oracle/isect_port.py) on and off."""
    import time
    from oracle import isect_port as IP
    dtype = torch.float32
    focal = cfg.get('focal_length') or (W ** 2 + H ** 2) ** 0.5
    prob = FrameProblem(body_model, keypoints, H, W, focal, joint_weights, dtype=dtype,
                        rho=cfg.get('rho', 100), body_prior=l2_body_prior,
                        confidence_threshold=cfg.get('confidence_threshold', 0))

    def eul(mats):
        return [euler_xyz_from_matrix(torch.tensor(np.asarray(m))) for m in mats]
    full = torch.cat(eul(expose['body_pose'])[:19] + eul(pixie['body_pose'])[19:]).reshape(1, -1)
    prob.pose_embedding = full.to(dtype)
    prob.reset_params(global_orient=eul([expose['global_orient']])[0], body_pose=prob.pose_embedding)
    tr = np.array(expose['transl'], dtype=np.float64).copy()
    tr[-1] /= (5000 / focal)
    prob.cam_t = torch.tensor(tr, dtype=dtype).reshape(1, 3)
    prob.center = torch.tensor([[float(v) for v in expose['center']]], dtype=dtype)
    si = 1
    w = dict(body_pose_weight=cfg['body_pose_prior_weights'][si], shape_weight=cfg['shape_weights'][si],
             expr_prior_weight=cfg['expr_weights'][si],
             hand_prior_weight=cfg['hand_pose_prior_weights'][si],
             jaw_prior_weight=[float(v) for v in cfg['jaw_pose_prior_weights'][si].split(',')],
             data_weight=1000.0 / H)
    w = {k: torch.tensor(v, dtype=dtype) for k, v in w.items()}
    w['bending_prior_weight'] = 3.17 * w['body_pose_weight']
    prob.jw[:, 25:67] = cfg['hand_joints_weights'][si]
    prob.jw[:, 67:] = cfg['face_joints_weights'][si]
    prob.jw[:, prob.low_conf] = 0
    jw = prob.jw.clone()
    names = [n for n in prob.P] + ['pose_embedding']
    reg = prob.pose_embedding.clone()
    out = []
    for on in (True, False):
        if on:
            prob.coll = (IP.BVH(max_collisions=cfg.get('max_collisions', 128)),
                         IP.DistanceFieldPenetrationLoss(
                             sigma=cfg.get('df_cone_height', 1e-4), point2plane=False,
                             vectorized=True, penalize_outside=True),
                         IP.FilterFaces(faces_segm=part_segm['segm'],
                                        faces_parents=part_segm['parents'],
                                        ign_part_pairs=cfg.get('ign_part_pairs')),
                         body_model.faces_tensor.view(-1))
            w['coll_loss_weight'] = torch.tensor(cfg['coll_loss_weights'][si], dtype=dtype)
        else:
            prob.coll = None
            w['coll_loss_weight'] = torch.tensor(0.0, dtype=dtype)
        x, closure, _ = prob.make_closure(
            names, lambda p, e, t: prob.smplify_loss(p, e, t, w, jw, si, 3, reg))
        closure()                                    # warm-up
        t0 = time.perf_counter()
        for _ in range(n_evals if on else 10 * n_evals):
            closure()
        out.append((time.perf_counter() - t0) / (n_evals if on else 10 * n_evals))
    return out[0], out[1]
