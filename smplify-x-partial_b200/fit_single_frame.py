"""``fit_single_frame`` with the reference's signature and outputs
(smplifyx/fit_single_frame.py:59-676): camera stage, optional 180-degree flipped orientation,
annealing stages, ``result_fn`` pickle (protocol 2) with the reference's keys and
``<result_folder>/vertices.ply``.

It is written against this package's mirrors of the reference API (``fitting.create_loss``,
``FittingMonitor.create_fitting_closure`` / ``run_fitting``, ``optim_factory.create_optimizer``,
``camera``, ``body_model``), so every stage is one CUDA launch.  The reference asserts
``batch_size == 1``; here ``keypoints`` may hold B frames ([B,K,3]) when ``body_model`` and
``camera`` were created with ``batch_size=B`` -- ``result_fn`` then receives a list of B result
dicts.  For large batches prefer ``fit_frames.fit_frames`` (whole flow in one launch).
"""
import os
import pickle
from collections import defaultdict

import numpy as np
import torch

from . import fit_frames as FF
from . import fitting
from .optimizers import optim_factory


def write_ply_vertices(path, vertices):
    """Binary little-endian PLY with float x, y, z -- what plyfile writes for the reference's
    ``PlyData([PlyElement.describe(..., 'vertices')], text=False, byte_order='<')``."""
    v = np.ascontiguousarray(vertices, dtype='<f4').reshape(-1, 3)
    header = ('ply\nformat binary_little_endian 1.0\nelement vertices {}\n'
              'property float x\nproperty float y\nproperty float z\nend_header\n').format(len(v))
    with open(path, 'wb') as f:
        f.write(header.encode('ascii'))
        f.write(v.tobytes())


def fit_single_frame(img, keypoints, body_model, camera, joint_weights, body_pose_prior,
                     jaw_prior, left_hand_prior, right_hand_prior, shape_prior, expr_prior,
                     angle_prior, result_fn='out.pkl', mesh_fn='out.obj', loss_type='smplify',
                     use_cuda=True, init_joints_idxs=(9, 12, 2, 5), use_face=True, use_hands=True,
                     data_weights=None, body_pose_prior_weights=None,
                     hand_pose_prior_weights=None, jaw_pose_prior_weights=None,
                     shape_weights=None, expr_weights=None, hand_joints_weights=None,
                     face_joints_weights=None, global_orient_weights=None, depth_loss_weight=1e2,
                     interpenetration=True, coll_loss_weights=None, df_cone_height=0.5,
                     penalize_outside=True, max_collisions=8, point2plane=False, part_segm_fn='',
                     focal_length=5000., side_view_thsh=25., rho=100, vposer_latent_dim=32,
                     vposer_ckpt='', use_joints_conf=False, interactive=True, visualize=False,
                     degrees=None, batch_size=1, dtype=torch.float32, ign_part_pairs=None,
                     left_shoulder_idx=2, right_shoulder_idx=5, result_folder='.', img_name='',
                     pixie_results=None, expose_results=None, pare_results=None,
                     regression_prior=None, format='coco25', smplx_path='', curr_img_folder='.',
                     **kwargs):
    B = body_model.batch_size
    dev = body_model.engine_model.device
    H, W = int(np.asarray(img).shape[0]), int(np.asarray(img).shape[1])
    use_vposer = kwargs.get('use_vposer', True)
    vposer = kwargs.pop('vposer', None)
    if use_vposer and vposer is None:
        from .vposer import load_vposer
        vposer, _ = load_vposer(os.path.expandvars(vposer_ckpt), vp_model='snapshot', dtype=dtype)
    if use_vposer:
        vposer = vposer.to(device=dev)
        vposer.eval()
    if visualize:
        raise NotImplementedError('visualisation is out of scope (SURVEY.md #16)')
    cfg = dict(kwargs)
    cfg.update(data_weights=data_weights, body_pose_prior_weights=body_pose_prior_weights,
               hand_pose_prior_weights=hand_pose_prior_weights,
               jaw_pose_prior_weights=jaw_pose_prior_weights, shape_weights=shape_weights,
               expr_weights=expr_weights, hand_joints_weights=hand_joints_weights,
               face_joints_weights=face_joints_weights, coll_loss_weights=coll_loss_weights,
               use_hands=use_hands, use_face=use_face, format=format,
               regression_prior=regression_prior, init_joints_idxs=list(init_joints_idxs),
               confidence_threshold=kwargs.get('confidence_threshold', 0))
    weights = FF.stage_weights(cfg)                      # validation + defaults (:136-207)
    # search tree, penetration penalty and face filter (:296-328)
    from . import mesh_intersection as MI
    search_tree, pen_distance, filter_faces = MI.create_term(
        interpenetration, max_collisions=max_collisions, df_cone_height=df_cone_height,
        point2plane=point2plane, penalize_outside=penalize_outside, part_segm_fn=part_segm_fn,
        ign_part_pairs=ign_part_pairs, part_segm=kwargs.get('part_segm'))
    from . import utils as U
    nb = U.NUM_BODY_KEYPOINTS[format]

    kp = np.asarray(keypoints, dtype=np.float32).reshape(B, -1, 3)
    K = kp.shape[1]
    expose = expose_results if isinstance(expose_results, (list, tuple)) else [expose_results] * B
    pixie = pixie_results if isinstance(pixie_results, (list, tuple)) else [pixie_results] * B
    pare = pare_results if isinstance(pare_results, (list, tuple)) else [pare_results] * B

    # --- regression prior -> initial pose (:209-274) ---
    pose0 = torch.zeros([B, 63], dtype=dtype, device=dev)
    go0 = None
    if regression_prior:
        po, go = zip(*[FF.regression_pose(cfg, expose[b], pixie[b], pare=pare[b]) for b in range(B)])
        pose0 = torch.tensor(np.stack(po), dtype=dtype, device=dev)
        go0 = torch.tensor(np.stack(go), dtype=dtype, device=dev)
    elif not use_vposer:
        pose0 = body_pose_prior.get_mean().detach().to(device=dev, dtype=dtype).expand(B, -1).clone()
    if use_vposer:
        if regression_prior:
            with torch.no_grad():
                pose0 = vposer.encode(pose0).sample()
        else:
            pose0 = torch.zeros([B, 32], dtype=dtype, device=dev)
    pose_embedding = pose0.clone().requires_grad_(True)
    if go0 is not None:
        body_model.reset_params(global_orient=go0, body_pose=pose_embedding)
    else:
        body_model.reset_params(body_pose=pose_embedding)

    gt_joints = torch.tensor(kp[:, :, :2], dtype=dtype, device=dev)
    joints_conf = torch.tensor(kp[:, :, 2], dtype=dtype, device=dev) if use_joints_conf else None
    base_jw = joint_weights.detach().cpu().numpy().reshape(-1, K)[0].astype(np.float64)
    jw_np, lowconf, init_mask = FF.keypoint_masks(kp.astype(np.float64), cfg, base_jw)
    if B > 1 and not np.all(init_mask == init_mask[0]):
        raise ValueError('frames of one fit_single_frame batch must share their visible '
                         'init_joints_idxs (use fit_frames.fit_frames for ragged batches)')
    init_idxs = torch.tensor(np.flatnonzero(init_mask[0]), dtype=torch.long, device=dev)
    jw = torch.tensor(jw_np, dtype=dtype, device=dev)
    low = torch.tensor(lowconf.astype(bool), device=dev)

    # --- camera initialisation (:359-411) ---
    with torch.no_grad():
        for b in range(B):
            pr = FF.camera_prior(cfg, focal_length, expose[b], pixie[b], pare[b])
            if pr is not None:
                camera.translation[b] = torch.tensor(pr[0], dtype=dtype, device=dev)
                camera.center[b] = torch.tensor(pr[1], dtype=dtype, device=dev)
            else:
                camera.center[b] = torch.tensor([W, H], dtype=dtype, device=dev) * 0.5
        if FF.camera_prior(cfg, focal_length, expose[0], pixie[0], pare[0]) is None:
            init_t = fitting.guess_init(body_model, gt_joints, kwargs.get('body_tri_idxs'),
                                        use_vposer=use_vposer, vposer=vposer,
                                        pose_embedding=pose_embedding,
                                        model_type=kwargs.get('model_type', 'smpl'),
                                        focal_length=focal_length, dtype=dtype)
            camera.translation[:] = init_t.view_as(camera.translation)
        init_t = camera.translation.detach().clone()

    camera_loss = fitting.create_loss(
        'camera_init', trans_estimation=init_t, init_joints_idxs=init_idxs,
        depth_loss_weight=depth_loss_weight, dtype=dtype, joints_conf=joints_conf,
        use_conf=kwargs.get('use_conf_for_camera_init')).to(device=dev)
    loss = fitting.create_loss(
        loss_type, joint_weights=jw, rho=rho, use_joints_conf=use_joints_conf, use_face=use_face,
        use_hands=use_hands, vposer=vposer, pose_embedding=pose_embedding,
        body_pose_prior=body_pose_prior, shape_prior=shape_prior, angle_prior=angle_prior,
        expr_prior=expr_prior, left_hand_prior=left_hand_prior, right_hand_prior=right_hand_prior,
        jaw_prior=jaw_prior, interpenetration=interpenetration, search_tree=search_tree,
        pen_distance=pen_distance, tri_filtering_module=filter_faces, dtype=dtype,
        regression_pose=pose_embedding.clone().detach() if regression_prior else None,
        num_stages=len(weights)).to(device=dev)

    monitor_kw = {k: kwargs[k] for k in ('maxiters', 'ftol', 'gtol', 'summary_steps', 'model_type')
                  if k in kwargs}
    with fitting.FittingMonitor(batch_size=B, visualize=False, **monitor_kw) as monitor:
        data_weight = 1000 / H
        camera_loss.reset_loss_weights({'data_weight': data_weight})
        # --- stage C (:473-496) ---
        camera.translation.requires_grad = True
        cam_params = [camera.translation, body_model.global_orient]
        cam_opt, cam_graph = optim_factory.create_optimizer(cam_params, **kwargs)
        closure = monitor.create_fitting_closure(
            cam_opt, body_model, camera, gt_joints, camera_loss, create_graph=cam_graph, use_vposer=False, vposer=None, pose_embedding=pose_embedding,
            return_full_pose=False, return_verts=False)
        monitor.run_fitting(cam_opt, closure, cam_params, body_model, stage=0,
                            use_vposer=use_vposer, pose_embedding=pose_embedding, vposer=vposer)
        camera.translation.requires_grad = False

        # --- orientations (:461-463, :527-538) ---
        sh = torch.sqrt(((gt_joints[:, left_shoulder_idx] - gt_joints[:, right_shoulder_idx]) ** 2)
                        .sum(-1)).cpu().numpy()
        try_both = bool((sh < side_view_thsh).all())
        if B > 1 and bool((sh < side_view_thsh).any()) != try_both:
            raise ValueError('frames of one fit_single_frame batch must agree on the side-view '
                             'test (use fit_frames.fit_frames for mixed batches)')
        orient0 = body_model.global_orient.detach().cpu().numpy().copy()
        orientations = [orient0]
        if try_both:
            orientations.append(np.stack([FF.flipped_orientation(o) for o in orient0]).astype(np.float32))

        results = []
        body_model_output = None
        for orient in orientations:
            new_params = defaultdict(global_orient=orient, body_pose=pose_embedding)
            body_model.reset_params(**new_params)
            final_loss_val = 0
            for opt_idx, w in enumerate(weights):
                final_params = [p for p in body_model.parameters() if p.requires_grad]
                final_params.append(pose_embedding)
                body_opt, body_graph = optim_factory.create_optimizer(final_params, **kwargs)
                body_opt.zero_grad()
                cur = dict(w)
                cur['data_weight'] = data_weight
                cur['bending_prior_weight'] = 3.17 * cur['body_pose_weight']
                if use_hands:
                    jw[:, nb:nb + 42] = cur['hand_weight']
                if use_face:
                    jw[:, nb + 42:] = cur['face_weight']
                jw[low] = 0
                loss.reset_loss_weights(cur)
                closure = monitor.create_fitting_closure(
                    body_opt, body_model, camera=camera, gt_joints=gt_joints,
                    joints_conf=joints_conf, joint_weights=jw, loss=loss, create_graph=body_graph,
                    use_vposer=use_vposer, vposer=vposer, pose_embedding=pose_embedding,
                    return_verts=True, return_full_pose=True)
                final_loss_val = monitor.run_fitting(
                    body_opt, closure, final_params, body_model, opt_idx,
                    pose_embedding=pose_embedding, vposer=vposer, use_vposer=use_vposer)

            def decoded():
                if use_vposer:
                    with torch.no_grad():
                        return vposer.decode(pose_embedding, output_type='aa').view(B, -1)
                return pose_embedding
            body_model_output = body_model(return_verts=True, body_pose=decoded())
            result = {'camera_' + str(k): v.detach().cpu().numpy()
                      for k, v in camera.named_parameters()}
            result['camera_center'] = camera.center.detach().cpu().numpy()
            result['H'], result['W'], result['focal_length'] = H, W, focal_length
            result.update({k: v.detach().cpu().numpy() for k, v in body_model.named_parameters()})
            result['body_pose'] = decoded().detach().cpu().numpy()
            results.append({'loss': final_loss_val, 'result': result})

    def pick(b):
        if len(results) == 1:
            return 0
        l0 = results[0]['loss'] if B == 1 else results[0]['loss'][b]
        l1 = results[1]['loss'] if B == 1 else results[1]['loss'][b]
        return 0 if l0 < l1 else 1
    if B == 1:
        chosen = results[pick(0)]['result']
    else:
        chosen = []
        for b in range(B):
            r = results[pick(b)]['result']
            chosen.append({k: (v[b:b + 1] if isinstance(v, np.ndarray) and v.shape[:1] == (B,) else v)
                           for k, v in r.items()})
    with open(result_fn, 'wb') as f:
        pickle.dump(chosen, f, protocol=2)
    if kwargs.get('save_vertices'):
        verts = body_model_output.vertices.detach().cpu().numpy()
        if B == 1:
            write_ply_vertices(os.path.join(result_folder, 'vertices.ply'), verts[0])
        else:
            for b in range(B):
                write_ply_vertices(os.path.join(result_folder, 'vertices_{:03d}.ply'.format(b)),
                                   verts[b])
    return chosen
