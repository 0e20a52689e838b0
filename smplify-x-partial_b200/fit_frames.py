"""Batched per-frame driver: the flow of the reference's ``fit_single_frame``
(smplifyx/fit_single_frame.py:59-676) for B independent frames fitted in lock-step on the
device.  The reference asserts ``batch_size == 1`` (fit_single_frame.py:119); here every frame
keeps its own optimiser state, line search and termination tests, so a batch gives the same
per-frame flow as B separate calls.

Split in two so the host logic is testable without a GPU:

* planning (pure numpy): weight schedules -> ``SfxStage`` list, regression-prior pose,
  keypoint masks, camera initialisation, orientation flip;
* execution: a handful of ``libsfx`` launches -- camera stage, then every annealing stage, one
  launch each, with no host synchronisation inside a stage.
"""
import numpy as np

from . import _native as N
from . import utils as U


# ------------------------------------------------------------------------------ planning
def stage_weights(cfg):
    """Per-stage weight dicts with the reference's defaults and validation
    (fit_single_frame.py:136-207, :331-348)."""
    bpw = cfg.get('body_pose_prior_weights')
    if bpw is None:
        bpw = [4.04 * 1e2, 4.04 * 1e2, 57.4, 4.78]
    S = len(bpw)
    dflt4 = [1e2, 5 * 1e1, 1e1, .5 * 1e1]

    def get(key, default, what):
        v = cfg.get(key)
        v = list(default) if v is None else list(v)
        if len(v) != S:
            raise AssertionError('Number of Body pose prior weights does not match the number '
                                 'of {}'.format(what))
        return v
    data_w = get('data_weights', [1] * S, 'data term weights')
    shape_w = get('shape_weights', dflt4, 'Shape prior weights')
    use_hands = cfg.get('use_hands', True)
    use_face = cfg.get('use_face', True)
    hand_prior_w = get('hand_pose_prior_weights', dflt4, 'hand pose prior weights') \
        if use_hands else [0.0] * S
    hand_joint_w = get('hand_joints_weights', [0.0, 0.0, 0.0, 1.0],
                       'hand joint distance weights') if use_hands else [0.0] * S
    if use_face:
        jaw = cfg.get('jaw_pose_prior_weights')
        if jaw is None:
            jaw = [[x] * 3 for x in shape_w]
        else:
            jaw = [[float(v) for v in s.split(',')] if isinstance(s, str) else
                   [float(v) for v in s] for s in jaw]
        if len(jaw) != S:
            raise AssertionError('Number of Body pose prior weights does not match the number '
                                 'of jaw pose prior weights')
        expr_w = get('expr_weights', dflt4, 'Expression prior weights')
        face_joint_w = get('face_joints_weights', [0.0, 0.0, 0.0, 1.0],
                           'face joint distance weights')
    else:
        jaw, expr_w, face_joint_w = [[0.0] * 3] * S, [0.0] * S, [0.0] * S
    coll_w = get('coll_loss_weights', [0.0] * S, 'collision loss weights')
    out = []
    for i in range(S):
        out.append(dict(data_weight=data_w[i], body_pose_weight=bpw[i], shape_weight=shape_w[i],
                        expr_prior_weight=expr_w[i], jaw_prior_weight=jaw[i],
                        hand_prior_weight=hand_prior_w[i], hand_weight=hand_joint_w[i],
                        face_weight=face_joint_w[i], coll_loss_weight=coll_w[i]))
    return out


_OPT_KIND = {'lbfgsls': N.OPT_LBFGSLS, 'adam': N.OPT_ADAM}


def _opt_kw(cfg):
    kind = cfg.get('optim_type', 'lbfgsls')
    if kind not in _OPT_KIND:
        raise ValueError('Optimizer {} not supported on the device (lbfgsls, adam)'.format(kind))
    return dict(opt_kind=_OPT_KIND[kind], lr=cfg.get('lr', 1.0), maxiters=cfg.get('maxiters', 30),
                ftol=cfg.get('ftol', 1e-9), gtol=cfg.get('gtol', 1e-9),
                adam_beta1=cfg.get('beta1', 0.9), adam_beta2=cfg.get('beta2', 0.999),
                two_loop=cfg.get('two_loop'))


def body_pose_prior_kind(cfg):
    """Which branch of SMPLifyLoss.forward's pose prior applies (fitting.py:389-401)."""
    if cfg.get('use_vposer', False):
        return N.PPRIOR_LATENT
    if cfg.get('regression_prior'):
        return N.PPRIOR_REGRESSION
    kind = cfg.get('body_prior_type', 'l2')
    if kind == 'l2':
        return N.PPRIOR_L2
    if kind == 'gmm':
        return N.PPRIOR_GMM
    raise ValueError('Prior {} is not implemented'.format(kind))


def make_stages(cfg, L):
    """-> (camera SfxStage, [body SfxStage per annealing stage])."""
    fmt = cfg.get('format', 'coco25')
    nb = U.NUM_BODY_KEYPOINTS[fmt]
    okw = _opt_kw(cfg)
    cam = N.make_stage(L, N.CAMERA_STAGE_BLOCKS, loss_kind=N.LOSS_CAMERA_INIT,
                       use_conf_camera=bool(cfg.get('use_conf_for_camera_init', False)),
                       depth_loss_weight=cfg.get('depth_loss_weight', 1e2), n_body_kpts=nb,
                       rho=cfg.get('rho', 100), use_vposer=cfg.get('use_vposer', False), **okw)
    weights = stage_weights(cfg)
    pk = body_pose_prior_kind(cfg)
    stages = []
    coll_on = bool(cfg.get('interpenetration', False))      # fitting.py:439
    for i, w in enumerate(weights):
        stages.append(N.make_stage(
            L, N.BODY_STAGE_BLOCKS, loss_kind=N.LOSS_SMPLIFY, pprior_kind=pk, stage_index=i,
            num_stages=len(weights), use_joints_conf=cfg.get('use_joints_conf', True),
            use_vposer=cfg.get('use_vposer', False), n_body_kpts=nb, rho=cfg.get('rho', 100),
            body_pose_weight=w['body_pose_weight'], shape_weight=w['shape_weight'],
            bending_prior_weight=3.17 * w['body_pose_weight'],
            hand_prior_weight=w['hand_prior_weight'], expr_prior_weight=w['expr_prior_weight'],
            jaw_prior_weight=w['jaw_prior_weight'], hand_joint_weight=w['hand_weight'],
            face_joint_weight=w['face_weight'],
            coll_loss_weight=w['coll_loss_weight'] if coll_on else 0.0,
            coll_sigma=cfg.get('df_cone_height', 0.5), **okw))
    return cam, stages


def base_joint_weights(cfg, K):
    """dataset.get_joint_weights() (data_parser.py:159-171)."""
    jw = np.ones(K, dtype=np.float64)
    ign = cfg.get('joints_to_ign')
    if ign is not None and -1 not in ign:
        jw[np.asarray(ign, dtype=np.int64)] = 0
    return jw


def keypoint_masks(keypoints, cfg, base_jw):
    """keypoints [B,K,3] -> (joint_weights [B,K], lowconf [B,K] u8, init_mask [B,K] u8)
    (fit_single_frame.py:285-294)."""
    B, K, _ = keypoints.shape
    nb = U.NUM_BODY_KEYPOINTS[cfg.get('format', 'coco25')]
    if K < nb + 42:
        raise ValueError('keypoints need the body, hand and face blocks (use_hands, use_face)')
    thr = np.zeros(K)
    thr[:nb] = cfg.get('confidence_threshold', 0)
    conf = keypoints[:, :, 2]
    lowconf = conf < thr[None]
    jw = np.repeat(np.asarray(base_jw, dtype=np.float64)[None], B, axis=0)
    jw[lowconf] = 0
    init = np.zeros((B, K), dtype=bool)
    idx = np.asarray(cfg.get('init_joints_idxs', (9, 12, 2, 5)), dtype=np.int64)
    init[:, idx] = True
    init &= (keypoints[:, :, 0] != 0) & (keypoints[:, :, 1] != 0) & ~lowconf
    return jw, lowconf.astype(np.uint8), init.astype(np.uint8)


def regression_pose(cfg, expose=None, pixie=None, dtype=np.float32, pare=None):
    """Regression prior -> (pose [63], global_orient [3]) as xyz-Euler angles used *as if*
    they were axis-angle, exactly like the reference (fit_single_frame.py:209-235)."""
    kind = cfg.get('regression_prior')
    if not kind:
        return None, None
    pix = exp = gp = None
    if kind == 'PARE':
        # pred_pose [1, 24, 3, 3]: joint 0 is the global orientation, 1..21 the body
        pp = np.asarray(pare['pred_pose'], dtype=dtype)
        full = U.euler_xyz_from_matrix(pp[0, 1:22])
        gp = U.euler_xyz_from_matrix(pp[0, :1])[0]
        return full.reshape(-1).astype(dtype), gp.reshape(-1).astype(dtype)
    if kind in ('PIXIE', 'combined'):
        pix = U.euler_xyz_from_matrix(np.asarray(pixie['body_pose'], dtype=dtype))
        gp = U.euler_xyz_from_matrix(np.asarray(pixie['global_pose'], dtype=dtype))[0]
    if kind in ('ExPose', 'combined'):
        exp = U.euler_xyz_from_matrix(np.asarray(expose['body_pose'], dtype=dtype))
        gp = U.euler_xyz_from_matrix(np.asarray(expose['global_orient'], dtype=dtype))[0]
    if kind == 'PIXIE':
        full = pix
    elif kind == 'ExPose':
        full = exp
    elif kind == 'combined':
        full = np.concatenate([exp[:19], pix[19:]])
    else:
        raise ValueError('unknown regression prior {}'.format(kind))
    return full.reshape(-1).astype(dtype), gp.reshape(-1).astype(dtype)


def regression_pose_batch(cfg, expose, pixie, B, dtype=np.float32, pare=None):
    """``regression_pose`` for B frames at once (the Euler conversion is elementwise, so the
    result is bit-identical to B separate calls): -> (pose [B,63], global_orient [B,3])."""
    kind = cfg.get('regression_prior')
    pix = exp = gp = None
    stack = lambda rows, key: np.stack([np.asarray(r[key], dtype=dtype) for r in rows])
    if kind == 'PARE':
        pp = stack(pare, 'pred_pose')[:, 0]                                      # [B,24,3,3]
        full = U.euler_xyz_from_matrix(pp[:, 1:22])
        gp = U.euler_xyz_from_matrix(pp[:, :1])[:, 0]
        return full.reshape(B, -1).astype(dtype), gp.reshape(B, -1).astype(dtype)
    if kind in ('PIXIE', 'combined'):
        pix = U.euler_xyz_from_matrix(stack(pixie, 'body_pose'))                 # [B,21,3]
        gp = U.euler_xyz_from_matrix(stack(pixie, 'global_pose'))[:, 0]
    if kind in ('ExPose', 'combined'):
        exp = U.euler_xyz_from_matrix(stack(expose, 'body_pose'))
        gp = U.euler_xyz_from_matrix(stack(expose, 'global_orient'))[:, 0]
    if kind == 'PIXIE':
        full = pix
    elif kind == 'ExPose':
        full = exp
    elif kind == 'combined':
        full = np.concatenate([exp[:, :19], pix[:, 19:]], axis=1)
    else:
        raise ValueError('unknown regression prior {}'.format(kind))
    return full.reshape(B, -1).astype(dtype), gp.reshape(B, -1).astype(dtype)


def camera_prior(cfg, focal, expose=None, pixie=None, pare=None):
    """-> (translation [3], centre [2]) or None when guess_init applies
    (fit_single_frame.py:359-411)."""
    kind = cfg.get('regression_prior')
    if not cfg.get('use_camera_prior') or not kind:
        return None
    if kind == 'PARE':
        # bounding box (cx, cy, b, _) of the 224-pixel crop and its weak-perspective camera
        RES = 224
        cx, cy, b, _ = [float(v) for v in np.asarray(pare['bboxes'])[0]]
        cam = np.asarray(pare['pred_cam'])[0]
        r = b / RES
        return (np.array([cam[1], cam[2], (2 * focal) / (r * RES * cam[0])], dtype=np.float64),
                np.array([cx, cy], dtype=np.float64))
    if kind in ('ExPose', 'combined'):
        t = np.array(expose['transl'], dtype=np.float64).copy()
        t[-1] /= (5000 / focal)
        return t, np.asarray(expose['center'], dtype=np.float64)
    if kind == 'PIXIE':
        left, top, right, bottom = [float(v) for v in pixie['bbox']]
        old = max(right - left, bottom - top)
        cen = np.array([right - (right - left) / 2.0, bottom - (bottom - top) / 2.0])
        size = int(old * 1.1)
        cam = pixie['body_cam']
        return np.array([cam[1], cam[2], 2 * focal / (cam[0] * size + 1e-9)]), cen
    return None


def guess_init_depth(joints3d, gt2d, edge_idxs, focal):
    """fitting.guess_init (fitting.py:36-110): similar-triangles depth from the mean 3-D and
    2-D edge lengths.  joints3d [B,K,3], gt2d [B,K,2] -> translation [B,3]."""
    e = np.asarray(edge_idxs, dtype=np.int64).reshape(-1, 2)
    d3 = joints3d[:, e[:, 0]] - joints3d[:, e[:, 1]]
    d2 = gt2d[:, e[:, 0]] - gt2d[:, e[:, 1]]
    l3 = np.sqrt((d3 ** 2).sum(-1))
    l2 = np.sqrt((d2 ** 2).sum(-1))
    est = focal * (l3.mean(axis=1) / l2.mean(axis=1))
    t = np.zeros((joints3d.shape[0], 3), dtype=joints3d.dtype)
    t[:, 2] = est
    return t


def flipped_orientation(go):
    """Rodrigues(go) . R_y(pi) back to axis-angle (fit_single_frame.py:527-538)."""
    R = U.rodrigues(go).dot(U.rodrigues([0.0, np.pi, 0.0]))
    return U.inv_rodrigues(R)


# ------------------------------------------------------------------------------ execution
class FitResult(object):
    """Per-batch output: ``results[b]`` is the reference's result dict of frame b
    (fit_single_frame.py:644-660), ``vertices`` [B,V,3] the mesh written to vertices.ply."""

    def __init__(self):
        self.results = []
        self.vertices = self.joints = self.loss = self.cam_loss = None
        self.n_evals = self.flags = self.n_orient = self.params = None
        self.h2d_bytes = self.d2h_bytes = 0


class FitPlan(object):
    """Host-side plan of one batch: initial parameters, targets, stage list, which frames get
    the second (flipped) orientation.  Pure numpy; built without touching the device."""

    def __init__(self, L, K, keypoints, H, W, cfg, expose=None, pixie=None, body_mean_pose=None,
                 np_dtype=np.float32, body_pose_prior=None, vposer=None, part_segm=None,
                 pare=None):
        B = keypoints.shape[0]
        if B < 1:
            raise ValueError('a fit needs at least one frame (sfx_batch_create rejects empty batches)')
        self.B, self.K, self.L, self.cfg = B, K, L, cfg
        npd = self.np_dtype = np_dtype
        self.keypoints = keypoints = np.asarray(keypoints, dtype=np.float64).reshape(B, K, 3)
        self.H = H = np.broadcast_to(np.asarray(H, dtype=np.float64), (B,))
        self.W = W = np.broadcast_to(np.asarray(W, dtype=np.float64), (B,))
        fl = cfg.get('focal_length')
        self.focal = focal = (np.sqrt(W ** 2 + H ** 2) if fl is None
                              else np.broadcast_to(float(fl), (B,)))
        # interpenetration term (fit_single_frame.py:296-328)
        self.collision = None
        if cfg.get('interpenetration', False) and any(
                w['coll_loss_weight'] > 0 for w in stage_weights(cfg)):
            from . import mesh_intersection as MI
            _, _, ff = MI.create_term(
                True, max_collisions=cfg.get('max_collisions', 8),
                df_cone_height=cfg.get('df_cone_height', 0.5),
                point2plane=cfg.get('point2plane', False),
                penalize_outside=cfg.get('penalize_outside', True),
                part_segm_fn=cfg.get('part_segm_fn', ''),
                ign_part_pairs=cfg.get('ign_part_pairs'), part_segm=part_segm)
            # no segmentation: the reference runs the term without FilterFaces
            # (fit_single_frame.py:317-328, fitting.py:449-450) -- so does the device search
            self.collision = ff if ff is not None else 'unfiltered'

        self.cam_stage, self.stages = make_stages(cfg, L)
        self.jw, self.lowconf, self.init_mask = keypoint_masks(
            keypoints, cfg, base_joint_weights(cfg, K))
        # --- initial parameters (fit_single_frame.py:209-274) ---
        x = np.zeros((B, L.np), dtype=np.float64)
        self.reg = None
        use_vposer = bool(cfg.get('use_vposer', False))
        self.vposer = vposer
        if use_vposer and (vposer is None or L.n_pose != 32):
            raise ValueError('use_vposer needs vposer= (vposer.VPoser) and a batch created with '
                             'use_vposer=True')
        if not use_vposer and L.n_pose != 63:
            raise ValueError('the batch was created with use_vposer=True but the config says False')
        if cfg.get('regression_prior'):
            self.reg = np.zeros((B, L.n_pose), dtype=np.float64)
            poses, gos = regression_pose_batch(cfg, expose, pixie, B, dtype=npd, pare=pare)
            x[:, L.off_go:L.off_go + 3] = gos
            if use_vposer:
                # pose_embedding = vposer.encode(prior).sample()  (fit_single_frame.py:245):
                # stochastic in the reference too; seed torch for reproducible runs
                import torch
                with torch.no_grad():
                    z = vposer.encode(torch.as_tensor(poses, dtype=vposer.dec_fc1_w.dtype)).sample()
                poses = z.cpu().numpy()
            self.reg[:] = poses
            x[:, L.off_pose:L.off_pose + L.n_pose] = poses
        elif not use_vposer:
            # body_mean_pose = body_pose_prior.get_mean() (fit_single_frame.py:250-252)
            if body_mean_pose is None and hasattr(body_pose_prior, 'get_mean'):
                body_mean_pose = body_pose_prior.get_mean().detach().cpu().numpy()
            if body_mean_pose is not None:
                x[:, L.off_pose:L.off_pose + L.n_pose] = np.asarray(body_mean_pose).reshape(1, -1)
        self.body_pose_prior = body_pose_prior
        if body_pose_prior_kind(cfg) == N.PPRIOR_GMM and getattr(body_pose_prior, 'kind', '') != 'gmm':
            raise ValueError("body_prior_type 'gmm' needs body_pose_prior=prior.MaxMixturePrior(...)")
        # --- camera initialisation (fit_single_frame.py:359-411) ---
        cam = np.zeros((B, N.SFX_CAM_STRIDE), dtype=np.float64)
        cam[:, N.SFX_CAM_FX] = focal
        cam[:, N.SFX_CAM_FY] = focal
        cam[:, N.SFX_CAM_R:N.SFX_CAM_R + 9] = np.eye(3).reshape(-1)
        cam[:, N.SFX_CAM_DW] = 1000.0 / H
        self.need_guess = []
        for b in range(B):
            pr = camera_prior(cfg, focal[b], None if expose is None else expose[b],
                              None if pixie is None else pixie[b],
                              None if pare is None else pare[b])
            if pr is None:
                self.need_guess.append(b)
                cam[b, 2:4] = (W[b] * 0.5, H[b] * 0.5)
            else:
                x[b, L.off_camt:L.off_camt + 3] = np.asarray(pr[0], dtype=npd)
                cam[b, 2:4] = np.asarray(pr[1], dtype=npd)
        cam[:, N.SFX_CAM_TZ] = x[:, L.off_camt + 2]
        self.x0, self.cam = x, cam
        # --- orientations (fit_single_frame.py:461-463): decided by the 2-D shoulders alone ---
        li, ri = cfg.get('left_shoulder_idx', 2), cfg.get('right_shoulder_idx', 5)
        kp32 = keypoints[:, :, :2].astype(npd)
        sh = np.sqrt(((kp32[:, li] - kp32[:, ri]) ** 2).sum(-1))
        self.flip_ids = np.flatnonzero(sh < cfg.get('side_view_thsh', 25.)).astype(np.int32)


def _set_targets(batch, plan):
    """Targets of the batch: the host-built masks (default), or with cfg['device_ingest'] the
    keypoints alone, split / thresholded / masked on the device (sfx_keypoint_masks) -- both give
    the same bits (tests/test_gpu_ingest.py)."""
    cfg = plan.cfg
    if cfg.get('device_ingest', False):
        nb = U.NUM_BODY_KEYPOINTS[cfg.get('format', 'coco25')]
        return batch.set_targets_from_keypoints(
            plan.keypoints, base_joint_weights(cfg, plan.K), cfg.get('init_joints_idxs', (9, 12, 2, 5)),
            nb, cfg.get('confidence_threshold', 0), plan.cam, plan.reg)
    return batch.set_targets(plan.keypoints, plan.jw, plan.lowconf, plan.init_mask, plan.cam, plan.reg)


def upload(batch, plan):
    """Host -> device copies of one batch (async on the current stream); returns the byte count."""
    import torch
    if plan.vposer is not None and plan.cfg.get('use_vposer', False):
        batch.model.set_vposer(plan.vposer.weights)
    if getattr(plan.body_pose_prior, 'kind', '') == 'gmm':
        batch.model.set_gmm(plan.body_pose_prior)
    if plan.collision is not None:
        ff = plan.collision
        if isinstance(ff, str):
            batch.model.set_collision_unfiltered()
        else:
            batch.model.set_collision(ff.faces_segm, ff.faces_parents, ff.ign_part_pairs)
        batch.enable_collisions()
    n = _set_targets(batch, plan)
    n += batch.set_params(plan.x0)
    if plan.need_guess:
        # guess_init needs the model joints at the initial parameters (fitting.py:36-110): evaluated
        # and turned into the camera depth on the device (sfx_batch_guess_init, the arithmetic of
        # guess_init_depth below), so nothing waits for the host here.  cfg['guess_init_host'] keeps
        # the round trip (cross-check); plan.x0 / plan.cam hold the depth only on that path.
        L = plan.L
        _, _, j3 = batch.eval(plan.cam_stage, want_joints=True)
        edges = plan.cfg.get('body_tri_idxs', [(5, 12), (2, 9)])
        if plan.cfg.get('guess_init_host', False):
            t = guess_init_depth(j3.cpu().numpy().astype(np.float64), plan.keypoints[:, :, :2],
                                 edges, plan.focal)
            g = plan.need_guess
            plan.x0[g, L.off_camt:L.off_camt + 3] = t[g].astype(plan.np_dtype)
            plan.cam[:, N.SFX_CAM_TZ] = plan.x0[:, L.off_camt + 2]
            n += _set_targets(batch, plan)
            n += batch.set_params(plan.x0)
        else:
            e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
            d2 = plan.keypoints[:, e[:, 0], :2] - plan.keypoints[:, e[:, 1], :2]
            len2d = np.sqrt((d2 ** 2).sum(-1)).mean(axis=1)
            need = np.zeros(plan.B, dtype=np.uint8)
            need[plan.need_guess] = 1
            n += batch.guess_init(j3, e, plan.focal, len2d, need)
    plan.flip_dev = plan.flip_mask_dev = plan.order_dev = None
    if len(plan.flip_ids):
        dev = batch.model.device
        plan.flip_dev = torch.as_tensor(plan.flip_ids, device=dev)
        mask = np.zeros(plan.B, dtype=np.uint8)
        mask[plan.flip_ids] = 1
        # longest frames first: the ones that fit two orientations
        order = np.concatenate([plan.flip_ids, np.flatnonzero(mask == 0)]).astype(np.int32)
        plan.flip_mask_dev = torch.as_tensor(mask, device=dev)
        plan.order_dev = torch.as_tensor(order, device=dev)
        n += plan.flip_ids.nbytes + mask.nbytes + order.nbytes
    # "wide" frames (csrc/sfx_stream.cuh): the frames that fit two orientations run twice as long as
    # the rest of a batch; while the batch leaves SMs idle each of them gets a cluster of 8 CTAs
    # (the library takes as many as fit next to the one-block frames).  cfg['wide_frames']: 'auto'
    # (default) or 'off' (every frame by one block: the bit-reproducible reference path)
    n_wide = 0
    if (plan.cfg.get('wide_frames') or 'auto') != 'off' and plan.order_dev is not None \
            and plan.collision is None and plan.np_dtype == np.float32:
        n_wide = len(plan.flip_ids)
    plan.pipeline = N.make_pipeline(plan.cam_stage, plan.stages, n_wide)
    batch.reset_counters()
    return n


def run(batch, plan, return_verts=True, out=None):
    """The whole multi-stage fit: ONE launch of the per-frame pipeline kernel (camera stage,
    annealing stages, flipped orientation + selection) and the full-mesh forward; nothing in
    here synchronises with the host.  Returns (cam_loss, verts, joints, n_launches)."""
    batch.fit_pipeline(plan.pipeline, plan.order_dev, plan.flip_mask_dev)
    verts = joints = None
    launches = 1
    if return_verts:
        verts, joints = batch.forward_mesh(last_orientation=True, out=out)
        launches += 4
    return batch.cam_loss(), verts, joints, launches


def run_staged(batch, plan, return_verts=True):
    """Same flow, one launch per stage (the path the FittingMonitor.run_fitting mirror uses);
    kept as the cross-check of the pipeline kernel."""
    launches = 0
    cam_loss = batch.fit_stage(plan.cam_stage)                 # stage C (:473-496)
    batch.begin_orientation(False)                             # reset_params (:546-551)
    launches += 2
    for st in plan.stages:
        batch.fit_stage(st)
        launches += 1
    if plan.flip_dev is not None:
        batch.begin_orientation(True, plan.flip_dev)
        launches += 1
        for st in plan.stages:
            batch.fit_stage(st, frame_ids=plan.flip_dev)
            launches += 1
    verts = joints = None
    if return_verts:
        # body_model(return_verts=True) after the LAST orientation (:611, :671-676)
        verts, joints = batch.forward_mesh()
        launches += 4
    if plan.flip_dev is not None:
        batch.select_orientation(plan.flip_dev)                # results[argmin] (:662-668)
        launches += 1
    return cam_loss, verts, joints, launches


class _LazyResults(object):
    """``FitResult.results``: the reference's per-frame result dicts (fit_single_frame.py:644-660),
    built when a frame is asked for -- a batch of thousands of frames does not pay for thousands of
    small dicts it may never read.  Behaves like a list (len, index, iteration)."""

    def __init__(self, plan, blocks, params, body_pose):
        self._plan, self._blocks, self._params, self._body_pose = plan, blocks, params, body_pose
        self._cache = {}

    def __len__(self):
        return self._plan.B

    def __iter__(self):
        return (self[b] for b in range(len(self)))

    def __getitem__(self, b):
        if isinstance(b, slice):
            return [self[i] for i in range(*b.indices(len(self)))]
        if b < 0:
            b += len(self)
        if not 0 <= b < len(self):
            raise IndexError(b)
        r = self._cache.get(b)
        if r is None:
            plan, L, params, npd = self._plan, self._plan.L, self._params, self._plan.np_dtype
            r = {'camera_rotation': plan.cam[b, 4:13].reshape(1, 3, 3).astype(npd),
                 'camera_translation': params[b, L.off_camt:L.off_camt + 3].reshape(1, 3).copy(),
                 'camera_center': plan.cam[b, 2:4].reshape(1, 2).astype(npd),
                 'H': int(plan.H[b]), 'W': int(plan.W[b]), 'focal_length': float(plan.focal[b])}
            for name in ('betas', 'global_orient', 'left_hand_pose', 'right_hand_pose', 'jaw_pose',
                         'leye_pose', 'reye_pose', 'expression'):
                off, n = self._blocks[name]
                r[name] = params[b, off:off + n].reshape(1, n).copy()
            if self._body_pose is not None:
                r['body_pose'] = self._body_pose[b:b + 1].copy()
            else:
                off, n = self._blocks['pose_embedding']
                r['body_pose'] = params[b, off:off + n].reshape(1, n).copy()
            self._cache[b] = r
        return r


class PendingFit(object):
    """A fit that has been queued on a CUDA stream (``submit``); ``finish`` waits for it and
    returns the ``FitResult``.  Several may be in flight on different ``FrameBatch`` objects: the
    straggler frames of one batch then overlap the next batch's frames."""

    def __init__(self):
        self.batch = self.plan = self.event = self.host = None
        self.h2d_bytes = self.gpu_launches = 0
        self.return_verts = True


def _host_buffers(batch, return_verts):
    """Pinned host buffers of one batch for the asynchronous device -> host copies (made once)."""
    import torch
    hb = getattr(batch, '_fit_host', None)
    if hb is None:
        B, dt = batch.B, batch.model.dtype
        pin = lambda shape, dtype: torch.empty(shape, dtype=dtype, pin_memory=True)
        hb = {'params': pin((B, batch.L.np), dt), 'loss': pin((B,), dt), 'cam_loss': pin((B,), dt),
              'n_evals': pin((B,), torch.int32), 'flags': pin((B,), torch.int32)}
        batch._fit_host = hb
    if return_verts and 'vertices' not in hb:
        hb['vertices'] = torch.empty((batch.B, batch.model.V, 3), dtype=batch.model.dtype, pin_memory=True)
        hb['joints'] = torch.empty((batch.B, batch.model.K, 3), dtype=batch.model.dtype, pin_memory=True)
    return hb


def submit(batch, keypoints, H, W, cfg, expose=None, pixie=None, return_verts=True,
           body_mean_pose=None, body_pose_prior=None, vposer=None, part_segm=None, pare=None):
    """Plans and queues the whole fit of ``batch`` on the current CUDA stream -- host -> device
    copies, the pipeline launch, the full-mesh forward, device -> host copies into pinned buffers
    -- and returns a ``PendingFit`` without waiting for the device (only ``guess_init``, when a
    frame has no usable camera prior, reads the model joints back first, as the reference does)."""
    import torch
    plan = FitPlan(batch.L, batch.model.K, np.asarray(keypoints), H, W, cfg, expose, pixie,
                   body_mean_pose, batch.model.np_dtype, body_pose_prior, vposer, part_segm, pare)
    p = PendingFit()
    p.batch, p.plan, p.return_verts = batch, plan, return_verts
    p.h2d_bytes = upload(batch, plan)
    cam_loss, verts, joints, p.gpu_launches = run(batch, plan, return_verts)
    p.host = _queue_download(batch, cam_loss, verts, joints)
    p.event = torch.cuda.Event()
    p.event.record()
    return p


def _queue_download(batch, cam_loss, verts, joints):
    """Device -> host copies of a fit's outputs into the batch's pinned buffers, queued on the
    current stream."""
    hb = _host_buffers(batch, verts is not None)
    hb['params'].copy_(batch.params_tensor(), non_blocking=True)
    hb['loss'].copy_(batch.final_loss(), non_blocking=True)
    hb['cam_loss'].copy_(cam_loss, non_blocking=True)
    hb['n_evals'].copy_(batch.evals(), non_blocking=True)
    hb['flags'].copy_(batch.flags(), non_blocking=True)
    if verts is not None:
        hb['vertices'].copy_(verts, non_blocking=True)
        if joints is not None:
            hb['joints'].copy_(joints, non_blocking=True)
    return hb


def _build_result(batch, plan, hb, return_verts):
    """FitResult from the pinned buffers (arrays are copies: the buffers belong to the batch and
    are reused by its next fit)."""
    out = FitResult()
    B = plan.B
    out.params = params = hb['params'].numpy().copy()
    out.loss = hb['loss'].numpy().astype(np.float64)
    out.cam_loss = hb['cam_loss'].numpy().copy()
    out.n_evals = hb['n_evals'].numpy().copy()
    out.flags = hb['flags'].numpy().copy()
    out.n_orient = np.ones(B, dtype=np.int64)
    out.n_orient[plan.flip_ids] = 2
    out.d2h_bytes = params.nbytes + 4 * B * 4
    if return_verts:
        out.vertices = hb['vertices'].numpy().copy()
        out.joints = hb['joints'].numpy().copy()
        out.d2h_bytes += out.vertices.nbytes + out.joints.nbytes
    body_pose = None
    if plan.vposer is not None and plan.cfg.get('use_vposer', False):
        # result['body_pose'] = vposer.decode(pose_embedding, 'aa') (fit_single_frame.py:653-657)
        import torch
        off, n = batch.blocks['pose_embedding']
        with torch.no_grad():
            z = torch.as_tensor(params[:, off:off + n], dtype=plan.vposer.dec_fc1_w.dtype)
            body_pose = plan.vposer.decode(z, output_type='aa').reshape(B, -1).cpu().numpy().astype(plan.np_dtype)
    out.results = _LazyResults(plan, batch.blocks, params, body_pose)
    return out


def download(batch, plan, cam_loss, verts, joints):
    """Device -> host: fitted parameters, losses, counters and (optionally) the meshes; waits for
    the current stream."""
    import torch
    hb = _queue_download(batch, cam_loss, verts, joints)
    torch.cuda.current_stream().synchronize()
    return _build_result(batch, plan, hb, verts is not None)


def finish(p):
    """Waits for a submitted fit and returns its ``FitResult``."""
    p.event.synchronize()
    out = _build_result(p.batch, p.plan, p.host, p.return_verts)
    out.h2d_bytes = p.h2d_bytes
    out.gpu_launches = p.gpu_launches
    out.batch_index = getattr(p, 'batch_index', None)
    return out


def fit_frames(batch, keypoints, H, W, cfg, expose=None, pixie=None, return_verts=True,
               body_mean_pose=None, body_pose_prior=None, vposer=None, part_segm=None, pare=None):
    """Fits every frame of ``batch`` (an ``engine.FrameBatch``).

    keypoints [B,K,3] (x, y, confidence) in the reference's row order; ``H``, ``W`` scalars or
    [B]; ``cfg`` the flat config dict of ``cmd_parser.parse_config`` (same keys as the
    reference's YAML files); ``expose`` / ``pixie`` lists of per-frame regression results.
    Synchronous form of ``submit`` + ``finish``.
    """
    return finish(submit(batch, keypoints, H, W, cfg, expose, pixie, return_verts, body_mean_pose,
                         body_pose_prior, vposer, part_segm, pare))
