"""The reference's fitting API (smplifyx/fitting.py) over the CUDA engine:

* ``create_loss('smplify' | 'camera_init', **kw)`` -> ``SMPLifyLoss`` / ``SMPLifyCameraInitLoss``
  with ``forward(body_model_output, camera, gt_joints, joints_conf, body_model_faces,
  joint_weights, use_vposer, pose_embedding, stage)`` and ``reset_loss_weights(dict)``
  (fitting.py:278-520);
* ``FittingMonitor(...)`` context manager with ``create_fitting_closure(...)`` ->
  ``fitting_func(stage=0, backward=True)`` and ``run_fitting(optimizer, closure, params,
  body_model, stage, ...) -> final loss`` (fitting.py:113-275);
* ``guess_init`` (fitting.py:36-110).

Same names, arguments and return values; batch sizes > 1 are accepted (every frame is an
independent problem).  Where the reference builds an autograd graph, this module calls libsfx:
a closure call is one ``sfx_eval`` launch (loss + analytic gradient written to ``.grad``) and
``run_fitting`` with a device optimiser (``lbfgsls``, ``adam``) is ONE ``sfx_fit_stage`` launch
for the whole stage -- no host synchronisation until the final loss is read.
"""
import sys

import numpy as np
import torch
import torch.nn as nn

from . import _native as N
from . import utils
from .optimizers.optim_factory import DeviceOptimizer


@torch.no_grad()
def guess_init(model, joints_2d, edge_idxs, focal_length=5000, pose_embedding=None, vposer=None,
               use_vposer=True, dtype=torch.float32, model_type='smpl', **kwargs):
    """Initial camera translation by similar triangles (fitting.py:36-110)."""
    if use_vposer:
        body_pose = vposer.decode(pose_embedding, output_type='aa').view(
            pose_embedding.shape[0], -1)
    else:
        body_pose = pose_embedding
    output = model(body_pose=body_pose, return_verts=False, return_full_pose=False)
    joints_3d = output.joints
    d3, d2 = [], []
    for e in edge_idxs:
        d3.append(joints_3d[:, e[0]] - joints_3d[:, e[1]])
        d2.append(joints_2d[:, e[0]] - joints_2d[:, e[1]])
    d3, d2 = torch.stack(d3, dim=1), torch.stack(d2, dim=1)
    l2 = d2.pow(2).sum(dim=-1).sqrt()
    l3 = d3.pow(2).sum(dim=-1).sqrt()
    est_d = focal_length * (l3.mean(dim=1) / l2.mean(dim=1))
    zeros = torch.zeros_like(est_d)
    return torch.stack([zeros, zeros.clone(), est_d], dim=1)


def _as_float(v, default=0.0):
    if v is None:
        return default
    if torch.is_tensor(v):
        return float(v.detach().cpu().reshape(-1)[0]) if v.numel() == 1 else \
            [float(x) for x in v.detach().cpu().reshape(-1)]
    return v


class _WeightedLoss(nn.Module):
    """reset_loss_weights of the reference only touches attributes that already exist
    (fitting.py:363-373, :490-497)."""

    def reset_loss_weights(self, loss_weight_dict):
        for key in loss_weight_dict:
            if hasattr(self, key):
                old = getattr(self, key)
                val = loss_weight_dict[key]
                if torch.is_tensor(val):
                    val = val.detach().clone().to(dtype=old.dtype, device=old.device)
                else:
                    val = torch.tensor(val, dtype=old.dtype, device=old.device)
                setattr(self, key, val)


class SMPLifyLoss(_WeightedLoss):
    def __init__(self, search_tree=None, pen_distance=None, tri_filtering_module=None, rho=100,
                 body_pose_prior=None, shape_prior=None, expr_prior=None, angle_prior=None,
                 jaw_prior=None, use_joints_conf=True, use_face=True, use_hands=True,
                 left_hand_prior=None, right_hand_prior=None, interpenetration=True,
                 dtype=torch.float32, data_weight=1.0, body_pose_weight=0.0, shape_weight=0.0,
                 bending_prior_weight=0.0, hand_prior_weight=0.0, expr_prior_weight=0.0,
                 jaw_prior_weight=0.0, coll_loss_weight=0.0, reduction='sum', vposer=None,
                 regression_pose=None, num_stages=3, **kwargs):
        super(SMPLifyLoss, self).__init__()
        self.use_joints_conf = use_joints_conf
        self.angle_prior = angle_prior
        self.rho = rho
        self.body_pose_prior = body_pose_prior
        self.shape_prior = shape_prior
        self.use_hands, self.use_face = use_hands, use_face
        self.left_hand_prior, self.right_hand_prior = left_hand_prior, right_hand_prior
        self.expr_prior, self.jaw_prior = expr_prior, jaw_prior
        self.interpenetration = interpenetration
        self.search_tree, self.pen_distance = search_tree, pen_distance
        self.tri_filtering_module = tri_filtering_module
        self.vposer = vposer
        self.regression_pose = regression_pose
        self.num_stages = num_stages
        t = lambda v: torch.tensor(v, dtype=dtype)
        self.register_buffer('data_weight', t(data_weight))
        self.register_buffer('body_pose_weight', t(body_pose_weight))
        self.register_buffer('shape_weight', t(shape_weight))
        self.register_buffer('bending_prior_weight', t(bending_prior_weight))
        if use_hands:
            self.register_buffer('hand_prior_weight', t(hand_prior_weight))
        if use_face:
            self.register_buffer('expr_prior_weight', t(expr_prior_weight))
            self.register_buffer('jaw_prior_weight', t(jaw_prior_weight))
        if interpenetration:
            self.register_buffer('coll_loss_weight', t(coll_loss_weight))

    def stage_kwargs(self, use_vposer, stage):
        """Weights + pose-prior branch of forward() (fitting.py:389-401) as SfxStage fields."""
        coll_w, coll_sigma = 0.0, 0.5
        if self.interpenetration and _as_float(getattr(self, 'coll_loss_weight', 0.0)) > 0:
            # fitting.py:439-455; tri_filtering_module = None is the reference's path without
            # --part_segm_fn: every pair the search reports is penalised
            if self.pen_distance is None:
                raise NotImplementedError(
                    'interpenetration on the device needs pen_distance '
                    '(mesh_intersection.create_term, reference fit_single_frame.py:300-328)')
            coll_w = _as_float(self.coll_loss_weight)
            coll_sigma = float(self.pen_distance.sigma)
        if use_vposer:
            pk = N.PPRIOR_LATENT
        elif self.regression_pose is not None:
            pk = N.PPRIOR_REGRESSION
        else:
            kind = getattr(self.body_pose_prior, 'kind', None)
            if kind == 'l2':
                pk = N.PPRIOR_L2
            elif kind == 'gmm':
                pk = N.PPRIOR_GMM
            else:
                raise ValueError('body_pose_prior must be an L2Prior or MaxMixturePrior')
        jaw = _as_float(getattr(self, 'jaw_prior_weight', None), 0.0)
        if not isinstance(jaw, list):
            jaw = [jaw] * 3
        return dict(
            loss_kind=N.LOSS_SMPLIFY, pprior_kind=pk, stage_index=stage, num_stages=self.num_stages,
            use_joints_conf=self.use_joints_conf, use_vposer=use_vposer, rho=float(self.rho),
            body_pose_weight=_as_float(self.body_pose_weight),
            shape_weight=_as_float(self.shape_weight),
            bending_prior_weight=_as_float(self.bending_prior_weight),
            hand_prior_weight=_as_float(getattr(self, 'hand_prior_weight', None), 0.0),
            expr_prior_weight=_as_float(getattr(self, 'expr_prior_weight', None), 0.0),
            jaw_prior_weight=jaw, coll_loss_weight=coll_w, coll_sigma=coll_sigma)

    def forward(self, body_model_output, camera, gt_joints, joints_conf, body_model_faces=None,
                joint_weights=None, use_vposer=False, pose_embedding=None, stage=0, **kwargs):
        bundle = _FitBundle.current(body_model_output, camera, gt_joints, joints_conf,
                                    joint_weights, self, use_vposer, pose_embedding)
        loss, _ = bundle.evaluate(stage, None)
        return loss.sum()


class SMPLifyCameraInitLoss(_WeightedLoss):
    def __init__(self, init_joints_idxs, trans_estimation=None, reduction='sum', data_weight=1.0,
                 depth_loss_weight=1e2, dtype=torch.float32, joints_conf=None, use_conf=False,
                 **kwargs):
        super(SMPLifyCameraInitLoss, self).__init__()
        self.dtype = dtype
        if trans_estimation is not None:
            self.register_buffer('trans_estimation', utils.to_tensor(trans_estimation, dtype=dtype))
        else:
            self.trans_estimation = trans_estimation
        self.register_buffer('data_weight', torch.tensor(data_weight, dtype=dtype))
        self.register_buffer('init_joints_idxs', utils.to_tensor(init_joints_idxs, dtype=torch.long))
        self.register_buffer('depth_loss_weight', torch.tensor(depth_loss_weight, dtype=dtype))
        self.joints_conf = joints_conf
        self.use_conf = use_conf

    def stage_kwargs(self, use_vposer, stage):
        dlw = _as_float(self.depth_loss_weight)
        return dict(loss_kind=N.LOSS_CAMERA_INIT, use_conf_camera=bool(self.use_conf),
                    depth_loss_weight=dlw if self.trans_estimation is not None else 0.0,
                    use_vposer=use_vposer)

    def forward(self, body_model_output, camera, gt_joints, **kwargs):
        bundle = _FitBundle.current(body_model_output, camera, gt_joints, self.joints_conf, None,
                                    self, kwargs.get('use_vposer', False),
                                    kwargs.get('pose_embedding'))
        loss, _ = bundle.evaluate(0, None)
        return loss.sum()


def create_loss(loss_type='smplify', **kwargs):
    if loss_type == 'smplify':
        return SMPLifyLoss(**kwargs)
    elif loss_type == 'camera_init':
        return SMPLifyCameraInitLoss(**kwargs)
    else:
        raise ValueError('Unknown loss type: {}'.format(loss_type))


class _FitBundle(object):
    """Everything one closure binds (fitting.py:219-231): model, camera, targets, loss.  Packs
    it into the frame batch on the device and builds the SfxStage for a launch."""

    def __init__(self, body_model, camera, gt_joints, joints_conf, joint_weights, loss,
                 use_vposer, pose_embedding):
        self.body_model, self.camera, self.loss = body_model, camera, loss
        self.gt_joints, self.joints_conf, self.joint_weights = gt_joints, joints_conf, joint_weights
        self.use_vposer, self.pose_embedding = bool(use_vposer), pose_embedding
        if self.use_vposer:
            vp = getattr(loss, 'vposer', None) or _FitBundle._last_vposer
            if vp is None:
                raise ValueError('use_vposer needs the vposer module (create_loss(vposer=...))')
            body_model.engine_model.set_vposer(vp.weights)
        self.batch = body_model.frame_batch(self.use_vposer)

    @staticmethod
    def current(body_model_output, camera, gt_joints, joints_conf, joint_weights, loss,
                use_vposer, pose_embedding):
        owner = getattr(body_model_output, 'owner', None) or _FitBundle._last_model
        if owner is None:
            raise RuntimeError('loss.forward needs an output of smplifyx_b200.body_model')
        return _FitBundle(owner, camera, gt_joints, joints_conf, joint_weights, loss, use_vposer,
                          pose_embedding)

    _last_model = None
    _last_vposer = None

    def push(self):
        bm, cam, loss, batch = self.body_model, self.camera, self.loss, self.batch
        B, K = bm.batch_size, bm.engine_model.K
        dev, dt = bm.engine_model.device, bm.dtype
        emb = self.pose_embedding if self.pose_embedding is not None else bm.body_pose
        bm.write_params(batch, pose_embedding=emb, camera_translation=cam.translation)
        gt = self.gt_joints.to(device=dev, dtype=dt).reshape(B, K, 2).contiguous()
        if self.joints_conf is not None:
            conf = self.joints_conf.to(device=dev, dtype=dt).reshape(B, K).contiguous()
        else:
            conf = torch.ones([B, K], dtype=dt, device=dev)
        if self.joint_weights is not None:
            jw = self.joint_weights.to(device=dev, dtype=dt).expand(B, K).contiguous()
        else:
            jw = torch.ones([B, K], dtype=dt, device=dev)
        lowconf = torch.zeros([B, K], dtype=torch.uint8, device=dev)
        init_mask = torch.zeros([B, K], dtype=torch.uint8, device=dev)
        camrow = torch.zeros([B, N.SFX_CAM_STRIDE], dtype=dt, device=dev)
        camrow[:, N.SFX_CAM_FX] = cam.focal_length_x
        camrow[:, N.SFX_CAM_FY] = cam.focal_length_y
        camrow[:, 2:4] = cam.center
        camrow[:, N.SFX_CAM_R:N.SFX_CAM_R + 9] = cam.rotation.detach().reshape(B, 9)
        camrow[:, N.SFX_CAM_DW] = loss.data_weight.to(device=dev, dtype=dt)
        reg = None
        if isinstance(loss, SMPLifyCameraInitLoss):
            init_mask[:, loss.init_joints_idxs.to(dev)] = 1
            if loss.trans_estimation is not None:
                camrow[:, N.SFX_CAM_TZ] = loss.trans_estimation.to(device=dev, dtype=dt)[:, 2]
        elif loss.regression_pose is not None:
            reg = loss.regression_pose.to(device=dev, dtype=dt).reshape(B, -1).contiguous()
        if getattr(getattr(loss, 'body_pose_prior', None), 'kind', '') == 'gmm':
            bm.engine_model.set_gmm(loss.body_pose_prior)
        ff = getattr(loss, 'tri_filtering_module', None)
        if getattr(loss, 'interpenetration', False):
            if ff is not None:
                bm.engine_model.set_collision(ff.faces_segm, ff.faces_parents, ff.ign_part_pairs)
            else:
                bm.engine_model.set_collision_unfiltered()     # fit_single_frame.py:317-328: no filter
            batch.enable_collisions()
        self._keep = (gt, conf, jw, lowconf, init_mask, camrow, reg)
        batch.set_targets_dev(gt, conf, jw, lowconf, init_mask, camrow, reg)

    def make_stage(self, stage, optimizer, params=None, maxiters=30, ftol=1e-9, gtol=1e-9):
        bm, batch = self.body_model, self.batch
        kw = self.loss.stage_kwargs(self.use_vposer, stage)
        # explicit joint weights come from the caller (fit_single_frame.py:569-574): no
        # per-stage hand / face overwrite inside the kernel
        kw.update(n_body_kpts=bm.engine_model.K, hand_joint_weight=0.0, face_joint_weight=0.0)
        if params is None:
            blocks = N.BODY_STAGE_BLOCKS if kw['loss_kind'] == N.LOSS_SMPLIFY \
                else N.CAMERA_STAGE_BLOCKS
        else:
            blocks = self.block_names(params)
        if optimizer is not None:
            kw.update(opt_kind=N.OPT_ADAM if optimizer.kind == 'adam' else N.OPT_LBFGSLS,
                      lr=optimizer.lr, max_iter=optimizer.max_iter,
                      adam_beta1=optimizer.beta1, adam_beta2=optimizer.beta2,
                      adam_eps=optimizer.eps, tol_grad=optimizer.tolerance_grad,
                      tol_change=optimizer.tolerance_change, history=optimizer.history_size)
        return N.make_stage(batch.L, blocks, maxiters=maxiters, ftol=ftol, gtol=gtol, **kw), blocks

    def block_names(self, params):
        """Optimiser parameter list -> engine block names, by identity; the reference's dead
        ``body_pose`` parameter (fit_single_frame.py:554-559) owns no block."""
        bm = self.body_model
        by_id = {id(getattr(bm, n)): n for n in
                 ('betas', 'global_orient', 'left_hand_pose', 'right_hand_pose', 'jaw_pose',
                  'leye_pose', 'reye_pose', 'expression') if getattr(bm, n, None) is not None}
        if self.pose_embedding is not None:
            by_id[id(self.pose_embedding)] = 'pose_embedding'
        by_id[id(self.camera.translation)] = 'camera_translation'
        names = []
        for p in params:
            if bm.body_pose is not None and p is bm.body_pose and p is not self.pose_embedding:
                continue
            if id(p) not in by_id:
                raise ValueError('optimiser parameter is not a body-model / camera parameter')
            names.append(by_id[id(p)])
        return names

    def evaluate(self, stage, params):
        self.push()
        st, blocks = self.make_stage(stage, None, params)
        loss, grad, _ = self.batch.eval(st)
        return loss, (grad, blocks)

    def pull(self, names=None):
        emb = self.pose_embedding if self.pose_embedding is not None else None
        self.body_model.read_params(self.batch, pose_embedding=emb,
                                    camera_translation=self.camera.translation, names=names)


class FittingMonitor(object):
    def __init__(self, summary_steps=1, visualize=False, maxiters=100, ftol=2e-09, gtol=1e-05,
                 body_color=(1.0, 1.0, 0.9, 1.0), model_type='smpl', **kwargs):
        super(FittingMonitor, self).__init__()
        self.maxiters, self.ftol, self.gtol = maxiters, ftol, gtol
        self.visualize = visualize
        self.summary_steps = summary_steps
        self.body_color = body_color
        self.model_type = model_type
        self.steps = 0

    def __enter__(self):
        self.steps = 0
        if self.visualize:
            raise NotImplementedError('interactive visualisation is out of scope (SURVEY.md #16)')
        return self

    def __exit__(self, exception_type, exception_value, traceback):
        pass

    def set_colors(self, vertex_color):
        pass

    def create_fitting_closure(self, optimizer, body_model, camera=None, gt_joints=None,
                               loss=None, joints_conf=None, joint_weights=None,
                               return_verts=True, return_full_pose=False, use_vposer=False,
                               vposer=None, pose_embedding=None, create_graph=False, **kwargs):
        """-> ``fitting_func(backward=True)``: one evaluation of the objective for every frame
        on the device; with ``backward`` the analytic gradient is written into ``.grad`` of the
        body-model parameters, the pose embedding and the camera translation
        (fitting.py:232-273).  Returns the summed loss."""
        if vposer is not None:
            _FitBundle._last_vposer = vposer
        bundle = _FitBundle(body_model, camera, gt_joints, joints_conf, joint_weights, loss,
                            use_vposer, pose_embedding)
        _FitBundle._last_model = body_model
        def fitting_func(stage=0, backward=True):
            if backward:
                optimizer.zero_grad()
            loss_b, (grad, _) = bundle.evaluate(stage, None)
            if backward:
                bm = bundle.body_model
                for name, (off, n) in bundle.batch.blocks.items():
                    if name == 'pose_embedding':
                        p = bundle.pose_embedding
                    elif name == 'camera_translation':
                        p = bundle.camera.translation
                    else:
                        p = getattr(bm, name, None)
                    if p is not None and p.requires_grad:
                        p.grad = grad[:, off:off + n].reshape(p.shape).clone()
            self.steps += 1
            return loss_b.sum()
        fitting_func.bundle = bundle
        return fitting_func

    def run_fitting(self, optimizer, closure, params, body_model, stage, use_vposer=True,
                    pose_embedding=None, vposer=None, **kwargs):
        """Runs the stage; returns the reference's value (loss at the start of the last completed
        iteration): a float for batch size 1, a list of floats otherwise (fitting.py:147-217)."""
        bundle = closure.bundle
        if isinstance(optimizer, DeviceOptimizer):
            bundle.push()
            st, blocks = bundle.make_stage(stage, optimizer, params, maxiters=self.maxiters,
                                           ftol=self.ftol, gtol=self.gtol)
            bundle.batch.reset_counters()       # the status flags are sticky: report this stage only
            final = bundle.batch.fit_stage(st)
            bundle.pull(names=set(blocks))
            vals = final.detach().cpu().numpy().astype(np.float64)
            flags = bundle.batch.flags().cpu().numpy()
            if (flags & 1).any():
                print('NaN loss value, stopping!')
            if (flags & 2).any():
                print('Inf loss value, stopping!')
            out = [None if np.isnan(v) else float(v) for v in vals]
            return out[0] if len(out) == 1 else out
        # host-driven torch optimisers (lbfgs, rmsprop, sgd): the reference's loop verbatim in
        # behaviour, every closure call is a device evaluation
        prev_loss = None
        for n in range(self.maxiters):
            loss = optimizer.step(lambda: closure(stage=stage))
            if torch.isnan(loss).sum() > 0:
                print('NaN loss value, stopping!')
                break
            if torch.isinf(loss).sum() > 0:
                print('Inf loss value, stopping!')
                break
            if n > 0 and prev_loss is not None and self.ftol > 0:
                if utils.rel_change(prev_loss, loss.item()) <= self.ftol:
                    break
            if all([torch.abs(var.grad.view(-1).max()).item() < self.gtol
                    for var in params if var.grad is not None]):
                break
            prev_loss = loss.item()
        return prev_loss
