"""Generates the committed fixtures under tests/golden/ by running the UNMODIFIED reference
(``/root/reference/smplifyx``, imported in place through ``oracle/ref_bridge.py``) in the
authoring container.  ``/root/reference`` does not exist on the GPU box, so the tests only
read the ``.npz`` files this script wrote.

    python tests/golden/make_golden.py            # all fixtures (about 2-3 minutes of CPU)

Fixtures
--------
demo_inputs.npz      the reference's demo/ inputs for the two frames: blended keypoints
                     [135,3] in reference row order (data_parser.py:57-104), image size, the
                     ExPose / PIXIE regression results the combined prior reads
                     (fit_single_frame.py:209-235, :370-401).  Inputs only, no code.
ref_eval_<case>.npz  one closure evaluation of the reference (fitting.py:232-273): reference
                     SMPLifyLoss / SMPLifyCameraInitLoss + reference PerspectiveCamera on the
                     restated smplx forward; loss and gradients w.r.t. every parameter.
ref_fit_02.npz       reference fit_single_frame end to end on demo frame 02 (combined
                     regression prior, camera prior, 3 stages, lbfgsls, no interpenetration):
                     result dict, final vertices, closure-evaluation count.
ref_stage_*.npz      reference FittingMonitor.run_fitting on one stage from a fixed start.

The SMPL-X model is the seeded synthetic one (``smplifyx_b200.synthetic``); the licensed model
files are not available.
"""
import os
import pickle
import sys
import tempfile
import json

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_bridge, smplx_shim          # noqa: E402
from smplifyx_b200 import synthetic                # noqa: E402
from smplifyx_b200 import utils as U               # noqa: E402

REF = ref_bridge.REF_ROOT
FRAMES = ['02_cropped', '18_cropped']


def cfg_combined():
    """cfg_files/fit_smplx_combined_coco25.yaml with the switches of BASELINE config 1."""
    from smplifyx_b200.cmd_parser import parse_config
    cfg = parse_config(['-c', os.path.join(REF, 'cfg_files', 'fit_smplx_combined_coco25.yaml')])
    cfg.pop('config')
    cfg.update(interpenetration=False, visualize=False, interactive=False, use_cuda=False,
               use_gender_classifier=False, gender='neutral', use_vposer=False)
    return cfg


def cfg_smplifyx():
    """cfg_files/fit_smplx_smplifyx.yaml (5 stages, VPoser latent pose, no regression prior)
    with the switches of BASELINE config 3."""
    from smplifyx_b200.cmd_parser import parse_config
    cfg = parse_config(['-c', os.path.join(REF, 'cfg_files', 'fit_smplx_smplifyx.yaml')])
    cfg.pop('config')
    cfg.update(interpenetration=False, visualize=False, interactive=False, use_cuda=False,
               use_gender_classifier=False, gender='neutral', use_vposer=True)
    return cfg


def demo_inputs():
    import cv2
    import joblib
    ref = ref_bridge.load()
    out = {}
    for fr in FRAMES:
        kp = ref.data_parser.read_keypoints(
            os.path.join(REF, 'demo', 'keypoints', fr + '_blended.json'),
            use_hands=True, use_face=True, use_face_contour=True).keypoints
        out[fr + '/keypoints'] = np.stack(kp)[0].astype(np.float32)
        img = cv2.imread(os.path.join(REF, 'demo', 'images', fr + '.jpg'))
        out[fr + '/HW'] = np.array(img.shape[:2], dtype=np.int64)
        ex = np.load(os.path.join(REF, 'demo', 'ExPose_results', fr + '.jpg',
                                  fr + '.jpg_params.npz'), allow_pickle=True)
        for k in ['global_orient', 'body_pose', 'betas', 'expression', 'transl', 'center',
                  'jaw_pose']:
            out[fr + '/expose/' + k] = np.asarray(ex[k])
        px = joblib.load(os.path.join(REF, 'demo', 'PIXIE_results', fr, fr + '_param.pkl'))
        for k in ['body_pose', 'global_pose', 'body_cam', 'bbox']:
            out[fr + '/pixie/' + k] = np.asarray(px[k])
    np.savez_compressed(os.path.join(HERE, 'demo_inputs.npz'), **out)
    return out


def build_reference_objects(cfg, dtype=torch.float32, use_vposer=False, md=None):
    """Body model (restated smplx), reference priors, reference joint weights."""
    ref = ref_bridge.load()
    md = synthetic.cached_smplx_like(0) if md is None else md
    jm = U.smpl_to_annotation('smplx', use_hands=True, use_face=True,
                              use_face_contour=cfg.get('use_face_contour', True),
                              format=cfg.get('format', 'coco25'))
    ref_jm = ref.utils.smpl_to_annotation('smplx', use_hands=True, use_face=True,
                                          use_face_contour=cfg.get('use_face_contour', True),
                                          format=cfg.get('format', 'coco25'))
    assert np.array_equal(np.asarray(ref_jm), jm), 'joint map differs from the reference'
    joint_mapper = ref.utils.JointMapper(jm.astype(np.int64))
    body_model = smplx_shim.create(
        model_data=md, joint_mapper=joint_mapper, create_body_pose=not use_vposer,
        use_pca=cfg.get('use_pca', True), num_pca_comps=cfg.get('num_pca_comps', 12),
        flat_hand_mean=cfg.get('flat_hand_mean', False), num_betas=cfg.get('num_betas', 10),
        num_expression_coeffs=cfg.get('num_expression_coeffs', 10),
        use_face_contour=cfg.get('use_face_contour', True), dtype=dtype)
    args = dict(cfg)
    args.pop('dtype', None)
    pri = {}
    mk = ref.prior.create_prior
    pri['body_pose_prior'] = mk(prior_type=args.get('body_prior_type'), dtype=dtype, **args)
    pri['jaw_prior'] = mk(prior_type=args.get('jaw_prior_type'), dtype=dtype, **args)
    pri['expr_prior'] = mk(prior_type='l2', dtype=dtype, **args)
    pri['left_hand_prior'] = mk(prior_type=args.get('left_hand_prior_type'), dtype=dtype, **args)
    pri['right_hand_prior'] = mk(prior_type=args.get('right_hand_prior_type'), dtype=dtype, **args)
    pri['shape_prior'] = mk(prior_type='l2', dtype=dtype, **args)
    pri['angle_prior'] = mk(prior_type='angle', dtype=dtype)
    nb = U.NUM_BODY_KEYPOINTS[cfg.get('format', 'coco25')]
    jw = np.ones(nb + 2 * 20 + 2 + 51 + 17 * int(cfg.get('use_face_contour', True)),
                 dtype=np.float32)
    ign = cfg.get('joints_to_ign')
    if ign is not None and -1 not in ign:
        jw[ign] = 0
    return body_model, pri, torch.tensor(jw, dtype=dtype).unsqueeze(0), jm


class _Counter(object):
    """Counts closure evaluations by wrapping body_model.forward."""

    def __init__(self, bm):
        self.n = 0
        self.bm = bm
        self.orig = bm.forward

        def fwd(*a, **k):
            self.n += 1
            return self.orig(*a, **k)
        bm.forward = fwd


def ref_fit(frame='02_cropped', inputs=None, cfg=None, tag='ref_fit_02', save=True,
            dtype=torch.float32):
    ref = ref_bridge.load()
    cfg = cfg or cfg_combined()
    torch.manual_seed(0)
    torch.set_num_threads(1)          # bit-stable reductions
    use_vposer = bool(cfg.get('use_vposer'))
    body_model, pri, jw, _ = build_reference_objects(cfg, dtype, use_vposer=use_vposer)
    if use_vposer:
        # the reference calls load_vposer(vposer_ckpt) (fit_single_frame.py:241); hand it the
        # restated VPoser v1 with the seeded synthetic weights
        from oracle import vposer_shim
        vp = vposer_shim.from_weights(synthetic.make_vposer_like(seed=2), dtype=dtype)
        ref.fit_single_frame.load_vposer = lambda *a, **k: (vp, None)
    H, W = [int(v) for v in inputs[frame + '/HW']]
    focal = cfg.get('focal_length')                      # main.py:212-214
    if focal is None:
        focal = (W ** 2 + H ** 2) ** 0.5
    args = dict(cfg)
    camera = ref.camera.create_camera(focal_length_x=focal, focal_length_y=focal, dtype=dtype,
                                      **args)
    camera.rotation.requires_grad = False
    args['focal_length'] = focal
    expose = {k.split('/')[-1]: inputs[k] for k in inputs if k.startswith(frame + '/expose/')}
    pixie = {k.split('/')[-1]: inputs[k] for k in inputs if k.startswith(frame + '/pixie/')}
    kp = inputs[frame + '/keypoints'][None]
    counter = _Counter(body_model)
    tmp = tempfile.mkdtemp()
    res_fn = os.path.join(tmp, '000.pkl')
    for k in ('dtype', 'output_folder', 'result_folder'):
        args.pop(k, None)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref.fit_single_frame.fit_single_frame(
            np.zeros((H, W, 3), np.float32), kp, body_model=body_model, camera=camera,
            joint_weights=jw.clone(), dtype=dtype, output_folder=tmp, result_folder=tmp,
            out_img_fn=os.path.join(tmp, 'o.png'), result_fn=res_fn,
            mesh_fn=os.path.join(tmp, 'o.obj'), img_name=frame, pixie_results=pixie,
            expose_results=expose, pare_results=None, smplx_path='', curr_img_folder=tmp,
            **pri, **args)
    with open(res_fn, 'rb') as f:
        result = pickle.load(f)
    verts = np.load(os.path.join(tmp, 'vertices.ply.npy'))
    out = {'result/' + k: np.asarray(v) for k, v in result.items()}
    out['vertices'] = verts.astype(np.float32 if dtype == torch.float32 else np.float64)
    out['n_forward_calls'] = np.array(counter.n)
    out['cfg_json'] = np.array(json.dumps({k: v for k, v in cfg.items()}))
    if save:
        np.savez_compressed(os.path.join(HERE, tag + '.npz'), **out)
    print(tag, 'forward calls', counter.n)
    return out


def _rand_params(rng, n_hand=12, scale=1.0):
    return dict(betas=rng.normal(size=(1, 10)) * scale,
                global_orient=rng.normal(size=(1, 3)) * 0.3 * scale + np.array([[3.0, 0.1, -0.1]]),
                left_hand_pose=rng.normal(size=(1, n_hand)) * 0.5 * scale,
                right_hand_pose=rng.normal(size=(1, n_hand)) * 0.5 * scale,
                jaw_pose=rng.normal(size=(1, 3)) * 0.1 * scale,
                leye_pose=rng.normal(size=(1, 3)) * 0.1 * scale,
                reye_pose=rng.normal(size=(1, 3)) * 0.1 * scale,
                expression=rng.normal(size=(1, 10)) * scale,
                pose_embedding=rng.normal(size=(1, 63)) * 0.2 * scale)


def ref_eval(inputs, dtype=torch.float64, tag='f64'):
    """One reference closure evaluation per loss kind at a random parameter point."""
    ref = ref_bridge.load()
    cfg = cfg_combined()
    frame = '18_cropped'
    body_model, pri, jw, _ = build_reference_objects(cfg, dtype)
    rng = np.random.default_rng(7)
    P = _rand_params(rng)
    H, W = [int(v) for v in inputs[frame + '/HW']]
    focal = (W ** 2 + H ** 2) ** 0.5
    camera = ref.camera.create_camera(focal_length_x=focal, focal_length_y=focal, dtype=dtype)
    with torch.no_grad():
        camera.translation[:] = torch.tensor([[0.05, 0.25, 3.2]], dtype=dtype)
        camera.center[:] = torch.tensor([[W * 0.5 + 3, H * 0.5 - 5]], dtype=dtype)
    camera.rotation.requires_grad = False
    camera.translation.requires_grad = True
    kp = torch.tensor(inputs[frame + '/keypoints'][None], dtype=dtype)
    gt, conf = kp[:, :, :2], kp[:, :, 2]
    reg = torch.tensor(rng.normal(size=(1, 63)) * 0.2, dtype=dtype)
    body_model.reset_params(**{k: v for k, v in P.items() if k != 'pose_embedding'})
    emb = torch.tensor(P['pose_embedding'], dtype=dtype, requires_grad=True)
    jw = jw.clone()
    jw[:, 25:67] = 0.1
    jw[:, 67:] = 2.0
    low = [i for i in range(25) if float(conf[0, i]) < 0.2]
    jw[:, low] = 0
    out = {'jw': jw.numpy(), 'keypoints': kp.numpy()[0], 'reg_pose': reg.numpy(),
           'cam_t': camera.translation.detach().numpy().copy(),
           'center': camera.center.numpy().copy(), 'focal': np.array(focal),
           'HW': np.array([H, W])}
    for k, v in P.items():
        out['param/' + k] = np.asarray(v)
    weights = dict(data_weight=1000.0 / H, body_pose_weight=300.0, shape_weight=50.0,
                   bending_prior_weight=3.17 * 300.0, hand_prior_weight=4.78,
                   expr_prior_weight=5.0, jaw_prior_weight=[100.0, 1000.0, 1000.0],
                   coll_loss_weight=0.0)
    out['weights_json'] = np.array(json.dumps(weights))
    # synthetic 8-component mixture prior written in the gmm_08.pkl format the reference loads
    gdir = tempfile.mkdtemp()
    with open(os.path.join(gdir, 'gmm_08.pkl'), 'wb') as f:
        pickle.dump(synthetic.make_gmm_like(seed=1, num_gaussians=8, dim=63), f)
    gmm_prior = ref.prior.create_prior('gmm', prior_folder=gdir, num_gaussians=8, dtype=dtype)
    for case, regression in (('l2', None), ('reg', reg), ('gmm', None)):
        pri_case = dict(pri)
        if case == 'gmm':
            pri_case['body_pose_prior'] = gmm_prior
        loss = ref.fitting.create_loss(
            loss_type='smplify', rho=100, use_joints_conf=True, use_face=True, use_hands=True,
            vposer=None, interpenetration=False, dtype=dtype, regression_pose=regression,
            num_stages=3, **pri_case)
        loss.reset_loss_weights({k: v for k, v in weights.items()})
        for p in list(body_model.parameters()) + [emb, camera.translation]:
            p.grad = None
        o = body_model(return_verts=True, body_pose=emb, return_full_pose=True)
        val = loss(o, camera=camera, gt_joints=gt, body_model_faces=None, joints_conf=conf,
                   joint_weights=jw, pose_embedding=emb, use_vposer=False, stage=1)
        val.backward()
        out[case + '/loss'] = val.detach().numpy()
        out[case + '/joints'] = o.joints.detach().numpy()[0]
        if case == 'l2':
            out['vertices'] = o.vertices.detach().numpy()[0]
        for name, p in body_model.named_parameters():
            if name != 'body_pose':
                out[case + '/grad/' + name] = p.grad.numpy().copy()
        out[case + '/grad/pose_embedding'] = emb.grad.numpy().copy()
        out[case + '/grad/camera_translation'] = camera.translation.grad.numpy().copy()
    # VPoser latent pose (fitting.py:235-236, :389-395): reference loss + restated VPoser v1
    from oracle import vposer_shim
    vposer = vposer_shim.from_weights(synthetic.make_vposer_like(seed=2), dtype=dtype)
    z = torch.tensor(rng.normal(size=(1, 32)) * 0.7, dtype=dtype, requires_grad=True)
    z_reg = torch.tensor(rng.normal(size=(1, 32)) * 0.7, dtype=dtype)
    out['vposer/latent'] = z.detach().numpy().copy()
    out['vposer/latent_reg'] = z_reg.numpy().copy()
    for case, regression, stage in (('vposer', None, 1), ('vposer_reg', z_reg, 2)):
        loss = ref.fitting.create_loss(
            loss_type='smplify', rho=100, use_joints_conf=True, use_face=True, use_hands=True,
            vposer=vposer, interpenetration=False, dtype=dtype, regression_pose=regression,
            num_stages=3, **pri)
        loss.reset_loss_weights({k: v for k, v in weights.items()})
        for p in list(body_model.parameters()) + [z, camera.translation]:
            p.grad = None
        bp = vposer.decode(z, output_type='aa').view(1, -1)
        o = body_model(return_verts=True, body_pose=bp, return_full_pose=True)
        val = loss(o, camera=camera, gt_joints=gt, body_model_faces=None, joints_conf=conf,
                   joint_weights=jw, pose_embedding=z, use_vposer=True, stage=stage)
        val.backward()
        out[case + '/loss'] = val.detach().numpy()
        out[case + '/body_pose'] = bp.detach().numpy()
        out[case + '/joints'] = o.joints.detach().numpy()[0]
        for name, p in body_model.named_parameters():
            if name != 'body_pose':
                out[case + '/grad/' + name] = p.grad.numpy().copy()
        out[case + '/grad/pose_embedding'] = z.grad.numpy().copy()
        out[case + '/grad/camera_translation'] = camera.translation.grad.numpy().copy()
    # camera-init loss, both confidence modes
    init_idxs = [i for i in cfg['init_joints_idxs'] if i not in low]
    out['init_idxs'] = np.array(init_idxs)
    for case, use_conf in (('cam', False), ('camconf', True)):
        closs = ref.fitting.create_loss(
            loss_type='camera_init', trans_estimation=torch.tensor([[0., 0., 3.5]], dtype=dtype),
            init_joints_idxs=torch.tensor(init_idxs), depth_loss_weight=100.0,
            camera_mode='moving', dtype=dtype, use_conf=use_conf, joints_conf=conf)
        closs.reset_loss_weights({'data_weight': 1000.0 / H})
        for p in list(body_model.parameters()) + [emb, camera.translation]:
            p.grad = None
        o = body_model(return_verts=False, body_pose=emb, return_full_pose=False)
        val = closs(o, camera, gt, body_model=body_model, joints_conf=conf)
        val.backward()
        out[case + '/loss'] = val.detach().numpy()
        out[case + '/grad/global_orient'] = body_model.global_orient.grad.numpy().copy()
        out[case + '/grad/camera_translation'] = camera.translation.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'ref_eval_{}.npz'.format(tag)), **out)
    print('ref_eval', tag, {k: float(out[k]) for k in out if k.endswith('/loss')})
    return out


COLL_IGN = ["9,16", "9,17", "6,16", "6,17", "1,2", "12,22"]      # the shipped yaml files


def coll_modules(md, max_collisions=128, sigma=1e-4):
    """The three objects fit_single_frame.py:305-328 builds, from the restated package."""
    from oracle import isect_port as IP
    seg = synthetic.parts_segm_like(md)
    ign = COLL_IGN + synthetic.sibling_part_pairs(md)
    return (IP.BVH(max_collisions=max_collisions),
            IP.DistanceFieldPenetrationLoss(sigma=sigma, point2plane=False, vectorized=True,
                                            penalize_outside=True),
            IP.FilterFaces(faces_segm=seg['segm'], faces_parents=seg['parents'],
                           ign_part_pairs=ign))


def ref_eval_coll(inputs, dtype=torch.float64, tag='f64', cases=(('coll', 0.1, True), ('nocoll', 0.0, True)),
                  fname='ref_eval_coll', clean_faces=False):
    """One closure evaluation of the reference's SMPLifyLoss WITH the interpenetration term
    (fitting.py:437-455) driving the restated mesh_intersection package (oracle/isect_port.py),
    at a pose whose arms cut into the torso.  ``cases``: (key, coll_loss_weight, with FilterFaces);
    without the filter (fit_single_frame.py:317-328 when part_segm_fn is empty) every intersecting
    pair that shares no vertex is penalised (``fname`` = 'ref_eval_collnf', case 'collnf');
    ``clean_faces``: on synthetic.without_degenerate_faces (a face without area has no cone)."""
    ref = ref_bridge.load()
    cfg = cfg_combined()
    frame = '18_cropped'
    md = synthetic.cached_smplx_like(0)
    if clean_faces:
        md = synthetic.without_degenerate_faces(md)
    body_model, pri, jw, _ = build_reference_objects(cfg, dtype, md=md)
    rng = np.random.default_rng(11)
    P = _rand_params(rng, scale=0.5)
    pe = P['pose_embedding']
    pe[0, 15 * 3 + 2] += -1.3          # shoulders down, elbows bent across the chest
    pe[0, 16 * 3 + 2] += 1.3
    pe[0, 17 * 3 + 1] += -1.6
    pe[0, 18 * 3 + 1] += 1.6
    H, W = [int(v) for v in inputs[frame + '/HW']]
    focal = (W ** 2 + H ** 2) ** 0.5
    camera = ref.camera.create_camera(focal_length_x=focal, focal_length_y=focal, dtype=dtype)
    with torch.no_grad():
        camera.translation[:] = torch.tensor([[0.05, 0.25, 3.2]], dtype=dtype)
        camera.center[:] = torch.tensor([[W * 0.5 + 3, H * 0.5 - 5]], dtype=dtype)
    camera.rotation.requires_grad = False
    camera.translation.requires_grad = True
    kp = torch.tensor(inputs[frame + '/keypoints'][None], dtype=dtype)
    gt, conf = kp[:, :, :2], kp[:, :, 2]
    body_model.reset_params(**{k: v for k, v in P.items() if k != 'pose_embedding'})
    emb = torch.tensor(pe, dtype=dtype, requires_grad=True)
    jw = jw.clone()
    jw[:, 25:67] = 0.1
    jw[:, 67:] = 2.0
    low = [i for i in range(25) if float(conf[0, i]) < 0.2]
    jw[:, low] = 0
    out = {'jw': jw.numpy(), 'keypoints': kp.numpy()[0],
           'cam_t': camera.translation.detach().numpy().copy(),
           'center': camera.center.numpy().copy(), 'focal': np.array(focal),
           'HW': np.array([H, W]), 'reg_pose': np.zeros((1, 63))}
    for k, v in P.items():
        out['param/' + k] = np.asarray(v)
    weights = dict(data_weight=1000.0 / H, body_pose_weight=300.0, shape_weight=50.0,
                   bending_prior_weight=3.17 * 300.0, hand_prior_weight=4.78,
                   expr_prior_weight=5.0, jaw_prior_weight=[100.0, 1000.0, 1000.0],
                   coll_loss_weight=0.1)
    out['weights_json'] = np.array(json.dumps(weights))
    out['sigma'] = np.array(1e-4)
    search_tree, pen_distance, filter_faces = coll_modules(md, sigma=1e-4)
    faces = body_model.faces_tensor.view(-1)
    for case, w_coll, filtered in cases:
        loss = ref.fitting.create_loss(
            loss_type='smplify', rho=100, use_joints_conf=True, use_face=True, use_hands=True,
            vposer=None, interpenetration=True, search_tree=search_tree,
            pen_distance=pen_distance, tri_filtering_module=filter_faces if filtered else None, dtype=dtype,
            regression_pose=None, num_stages=3, **pri)
        ww = dict(weights)
        ww['coll_loss_weight'] = w_coll
        loss.reset_loss_weights(ww)
        for p in list(body_model.parameters()) + [emb, camera.translation]:
            p.grad = None
        o = body_model(return_verts=True, body_pose=emb, return_full_pose=True)
        val = loss(o, camera=camera, gt_joints=gt, body_model_faces=faces, joints_conf=conf,
                   joint_weights=jw, pose_embedding=emb, use_vposer=False, stage=1)
        val.backward()
        out[case + '/loss'] = val.detach().numpy()
        for name, p in body_model.named_parameters():
            if name != 'body_pose':
                out[case + '/grad/' + name] = p.grad.numpy().copy()
        out[case + '/grad/pose_embedding'] = emb.grad.numpy().copy()
        out[case + '/grad/camera_translation'] = camera.translation.grad.numpy().copy()
        if w_coll > 0:
            tri = o.vertices[:, faces].view(1, -1, 3, 3)
            ci = search_tree(tri)
            if filtered:
                ci = filter_faces(ci)
            pairs = ci[0][ci[0, :, 0] >= 0].numpy()
            out[case + '/pairs'] = pairs.astype(np.int32)
            out[case + '/vertices'] = o.vertices.detach().numpy()[0]
    np.savez_compressed(os.path.join(HERE, '{}_{}.npz'.format(fname, tag)), **out)
    print(fname, tag, {c[0]: float(out[c[0] + '/loss']) for c in cases},
          'pairs', {c[0]: len(out[c[0] + '/pairs']) for c in cases if c[1] > 0})
    return out


def ref_blending():
    """The reference's keypoints_blending.blending() (keypoints_blending.py:276-381), unmodified,
    on seeded detections and seeded statistics.  Its non-arithmetic imports (mmcv, mmpose, PIL,
    matplotlib) are stubbed and its hard-coded '/content/heuristics/' paths are redirected to a
    temporary folder by giving the module its own ``open``."""
    import builtins
    import importlib
    import types
    ref_bridge.load()
    for name in ('mmcv', 'mmcv.visualization', 'mmcv.visualization.image', 'mmcv.image', 'mmpose',
                 'mmpose.core', 'PIL', 'matplotlib', 'matplotlib.pyplot'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__getattr__ = lambda attr: (lambda *a, **k: None)
            sys.modules[name] = m
    sys.path.insert(0, os.path.join(REF, 'smplifyx'))
    KB = importlib.import_module('keypoints_blending')
    from smplifyx_b200 import keypoints_blending as PB
    rng = np.random.default_rng(33)
    tmp = tempfile.mkdtemp()
    heur = os.path.join(tmp, 'heuristics')
    for d in ('images', 'op', 'mm', 'out', 'heuristics'):
        os.makedirs(os.path.join(tmp, d))
    keys = [p[0] for p in PB.blended_pairs()]
    stats = {'openpose_means': {k: float(rng.uniform(0.4, 0.8)) for k in keys},
             'openpose_stds': {k: float(rng.uniform(0.1, 0.3)) for k in keys},
             'mmpose_means': {k: float(rng.uniform(0.5, 0.9)) for k in keys},
             'mmpose_stds': {k: float(rng.uniform(0.05, 0.2)) for k in keys}}
    for k, v in stats.items():
        with open(os.path.join(heur, k + '.json'), 'w') as f:
            json.dump(v, f)

    def redirected_open(fn, *a, **kw):
        if isinstance(fn, str) and fn.startswith('/content/heuristics/'):
            fn = os.path.join(heur, os.path.basename(fn))
        return builtins.open(fn, *a, **kw)
    KB.open = redirected_open
    out = {'stats_json': np.array(json.dumps(stats))}
    ops, mms, res = [], [], []
    for n in range(6):
        # the reference writes only the last image's file: one image per call
        for d in ('images', 'op', 'mm'):
            for fn in os.listdir(os.path.join(tmp, d)):
                os.remove(os.path.join(tmp, d, fn))
        name = 'img%02d' % n
        open(os.path.join(tmp, 'images', name + '.jpg'), 'w').close()
        op = np.concatenate([rng.uniform(0, 800, size=(137, 2)),
                             rng.uniform(-0.1, 1.2, size=(137, 1))], axis=1).astype(np.float32)
        mm = np.concatenate([rng.uniform(0, 800, size=(138, 2)),
                             rng.uniform(0.0, 1.1, size=(138, 1))], axis=1).astype(np.float32)
        op[rng.uniform(size=137) < 0.1] = 0            # undetected keypoints
        for arr, body, fn in ((op, 25, os.path.join(tmp, 'op', name + '_keypoints.json')),
                              (mm, 26, os.path.join(tmp, 'mm', name + '_mmpose.json'))):
            person = {'pose_keypoints_2d': arr[:body].reshape(-1).tolist(),
                      'hand_left_keypoints_2d': arr[body:body + 21].reshape(-1).tolist(),
                      'hand_right_keypoints_2d': arr[body + 21:body + 42].reshape(-1).tolist(),
                      'face_keypoints_2d': arr[body + 42:].reshape(-1).tolist()}     # 70 points
            with open(fn, 'w') as f:
                json.dump({'people': [person]}, f)
        KB.blending(os.path.join(tmp, 'images'), os.path.join(tmp, 'op'), os.path.join(tmp, 'mm'),
                    os.path.join(tmp, 'out'))
        with open(os.path.join(tmp, 'out', name + '_blended.json')) as f:
            person = json.load(f)['people'][0]
        res.append(np.concatenate([np.array(person[k]).reshape(-1, 3) for k in
                                   ('pose_keypoints_2d', 'hand_left_keypoints_2d',
                                    'hand_right_keypoints_2d', 'face_keypoints_2d')]))
        ops.append(op[:135])
        mms.append(mm[:136])
    out['openpose'] = np.stack(ops)
    out['mmpose'] = np.stack(mms)
    out['blended'] = np.stack(res)
    np.savez_compressed(os.path.join(HERE, 'ref_blending.npz'), **out)
    print('ref_blending: rows taken from MMPose', int((out['blended'][:, :67, :2] != out['openpose'][:, :67, :2]).any(-1).sum()), 'of', 6 * 67)
    return out


def ref_metrics():
    """The reference's alignment classes (utils.py:540-801) and eval.py's compute_v2v loop on
    seeded point sets: a rotated / scaled / noisy copy of a mesh-sized cloud, a reflected one
    (exercises the det(R) = +1 fix) and a vertex subset."""
    ref = ref_bridge.load()
    U = ref.utils
    rng = np.random.default_rng(21)
    B, Np = 4, 600
    gt = rng.normal(size=(B, Np, 3)) * np.array([0.3, 0.8, 0.2])
    est = np.zeros_like(gt)
    for b in range(B):
        q, r = np.linalg.qr(rng.normal(size=(3, 3)))
        q = q * np.sign(np.diag(r))
        if (np.linalg.det(q) < 0) != (b == 1):   # frame 1 is a reflection: Z must flip the last axis
            q[:, 0] *= -1
        est[b] = (gt[b] @ q.T) * rng.uniform(0.5, 2.0) + rng.normal(size=3) + \
            rng.normal(size=(Np, 3)) * 0.01
    vids = np.sort(rng.choice(Np, size=200, replace=False))
    out = {'est': est, 'gt': gt, 'vids': vids.astype(np.int64)}
    pa, pe, sa = U.ProcrustesAlignmentMPJPE(), U.PelvisAlignmentMPJPE(), U.ScaleAlignment()
    out['procrustes'] = np.stack([pa(est[b], gt[b])['point'] for b in range(B)])
    out['procrustes_aligned'] = np.stack([U.ProcrustesAlignment()(est[b], gt[b]) for b in range(B)])
    out['pelvis'] = np.stack([pe(est[b], gt[b])['point'] for b in range(B)])
    out['scale_aligned'] = np.stack([sa(est[b], gt[b]) for b in range(B)])
    out['none'] = np.stack([U.mpjpe(est[b], gt[b]) for b in range(B)])
    out['procrustes_vids'] = np.stack([pa(est[b][vids], gt[b][vids])['point'] for b in range(B)])
    out['pelvis_vids'] = np.stack([pe(est[b][vids], gt[b][vids])['point'] for b in range(B)])
    np.savez_compressed(os.path.join(HERE, 'ref_metrics.npz'), **out)
    print('ref_metrics: PA error mean', out['procrustes'].mean(), 'pelvis', out['pelvis'].mean())
    return out


def ref_stage(inputs, dtype=torch.float64, tag='f64', perturb=None, save=True):
    """Reference run_fitting + reference LBFGS on one body stage from the ref_eval start.
    ``perturb=(eps, seed)`` multiplies the start by 1 + eps * N(0,1) (envelope runs)."""
    ref = ref_bridge.load()
    ev = dict(np.load(os.path.join(HERE, 'ref_eval_{}.npz'.format(
        'f64' if dtype == torch.float64 else 'f32'))))
    cfg = cfg_combined()
    body_model, pri, _, _ = build_reference_objects(cfg, dtype)
    H, W = [int(v) for v in ev['HW']]
    focal = float(ev['focal'])
    camera = ref.camera.create_camera(focal_length_x=focal, focal_length_y=focal, dtype=dtype)
    with torch.no_grad():
        camera.translation[:] = torch.tensor(ev['cam_t'], dtype=dtype)
        camera.center[:] = torch.tensor(ev['center'], dtype=dtype)
    camera.rotation.requires_grad = False
    camera.translation.requires_grad = False
    kp = torch.tensor(ev['keypoints'][None], dtype=dtype)
    gt, conf = kp[:, :, :2], kp[:, :, 2]
    jw = torch.tensor(ev['jw'], dtype=dtype)
    P = {k[6:]: ev[k] for k in ev if k.startswith('param/')}
    if perturb is not None:
        prng = np.random.default_rng(perturb[1])
        P = {k: v * (1 + perturb[0] * prng.normal(size=v.shape)) for k, v in P.items()}
    body_model.reset_params(**{k: v for k, v in P.items() if k != 'pose_embedding'})
    emb = torch.tensor(P['pose_embedding'], dtype=dtype, requires_grad=True)
    weights = json.loads(str(ev['weights_json']))
    loss = ref.fitting.create_loss(
        loss_type='smplify', rho=100, use_joints_conf=True, use_face=True, use_hands=True,
        vposer=None, interpenetration=False, dtype=dtype,
        regression_pose=torch.tensor(ev['reg_pose'], dtype=dtype), num_stages=3, **pri)
    loss.reset_loss_weights(weights)
    counter = _Counter(body_model)
    params = [p for p in body_model.parameters() if p.requires_grad] + [emb]
    import warnings
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        with ref.fitting.FittingMonitor(batch_size=1, visualize=False, maxiters=30, ftol=1e-9,
                                        gtol=1e-9, model_type='smplx') as monitor:
            opt, cg = ref.optim_factory.create_optimizer(params, optim_type='lbfgsls', lr=1.0,
                                                         maxiters=30)
            closure = monitor.create_fitting_closure(
                opt, body_model, camera=camera, gt_joints=gt, joints_conf=conf,
                joint_weights=jw, loss=loss, create_graph=cg, use_vposer=False, vposer=None,
                pose_embedding=emb, return_verts=True, return_full_pose=True)
            final = monitor.run_fitting(opt, closure, params, body_model, pose_embedding=emb,
                                        vposer=None, use_vposer=False, stage=1)
    out['final_loss'] = np.array(final)
    out['n_forward_calls'] = np.array(counter.n)
    for name, p in body_model.named_parameters():
        out['param/' + name] = p.detach().numpy().copy()
    out['param/pose_embedding'] = emb.detach().numpy().copy()
    with torch.no_grad():
        o = body_model(return_verts=True, body_pose=emb)
    out['vertices'] = o.vertices.numpy()[0].astype(np.float32)
    out['joints'] = o.joints.numpy()[0]
    if save:
        np.savez_compressed(os.path.join(HERE, 'ref_stage_{}.npz'.format(tag)), **out)
    print('ref_stage', tag, perturb, 'final', final, 'forward calls', counter.n)
    return out


def _pairwise(vs):
    mx, mean = 0.0, 0.0
    for i in range(len(vs)):
        for j in range(i + 1, len(vs)):
            d = np.abs(vs[i] - vs[j])
            mx, mean = max(mx, float(d.max())), max(mean, float(d.mean()))
    return mx, mean


def ref_envelope(inputs):
    """Run-to-run envelope of the reference itself: the same float32 problems started from
    points that differ by one float32 ulp (relative 6e-8).  The fitting problem is chaotic
    (line-search branch flips), so these runs end at visibly different points; the spread is
    the yardstick the engine's fitted results are held to."""
    torch.set_num_threads(1)
    runs = [ref_stage(inputs, torch.float32, 'f32', perturb=(6e-8, s), save=False)
            for s in range(1, 7)]
    base = ref_stage(inputs, torch.float32, 'f32', save=False)
    allr = [base] + runs
    out = {'stage/final_loss': np.array([float(r['final_loss']) for r in allr]),
           'stage/n_forward_calls': np.array([int(r['n_forward_calls']) for r in allr]),
           'stage/vertices_base': base['vertices']}
    mx, mean = _pairwise([r['vertices'] for r in allr])
    out['stage/vertex_pairwise_max'] = np.array(mx)
    out['stage/vertex_pairwise_mean'] = np.array(mean)
    for k in base:
        if k.startswith('param/'):
            st = np.stack([r[k] for r in allr])
            out['stage/spread/' + k[6:]] = np.array(float((st.max(0) - st.min(0)).max()))
    fits = []
    for s in range(1, 6):
        prng = np.random.default_rng(100 + s)
        inp = dict(inputs)
        kp = inp['02_cropped/keypoints'].copy()
        kp[:, :2] *= (1 + 6e-8 * prng.normal(size=kp[:, :2].shape)).astype(np.float32)
        inp['02_cropped/keypoints'] = kp
        fits.append(ref_fit('02_cropped', inp, save=False))
    fits.append(dict(np.load(os.path.join(HERE, 'ref_fit_02.npz'))))
    out['fit/n_runs_nan'] = np.array(sum(1 for f in fits if not np.isfinite(f['vertices']).all()))
    fits = [f for f in fits if np.isfinite(f['vertices']).all()]
    mx, mean = _pairwise([f['vertices'] for f in fits])
    out['fit/vertex_pairwise_max'] = np.array(mx)
    out['fit/vertex_pairwise_mean'] = np.array(mean)
    out['fit/n_forward_calls'] = np.array([int(f['n_forward_calls']) for f in fits])
    for k in fits[0]:
        if k.startswith('result/') and np.asarray(fits[0][k]).ndim == 2:
            st = np.stack([f[k] for f in fits])
            out['fit/spread/' + k[7:]] = np.array(float((st.max(0) - st.min(0)).max()))
    np.savez_compressed(os.path.join(HERE, 'ref_envelope.npz'), **out)
    for k, v in out.items():
        if np.asarray(v).size < 10:
            print(k, v)


if __name__ == '__main__':
    if not ref_bridge.available():
        raise SystemExit('reference tree not found at ' + REF)
    inp = demo_inputs()
    if len(sys.argv) > 1 and sys.argv[1] == 'envelope':
        ref_envelope(inp)
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'blending':
        ref_blending()
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'metrics':
        ref_metrics()
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'fit64':
        # float64 full fits of both demo frames: the well-conditioned end-to-end fixtures.  No
        # regression prior: the reference builds that prior with torch.tensor(<float32 array>)
        # (fit_single_frame.py:211-235), so its pose embedding stays a float32 parameter even in a
        # float64 run and is re-quantised by every optimiser update.  The un-initialised flow with
        # the mixture prior (pose_embedding = body_pose_prior.get_mean(), :250-252; guess_init
        # camera) is float64 throughout.
        cfg = cfg_combined()
        gdir = tempfile.mkdtemp()
        with open(os.path.join(gdir, 'gmm_08.pkl'), 'wb') as f:
            pickle.dump(synthetic.make_gmm_like(seed=1, num_gaussians=8, dim=63), f)
        cfg.update(regression_prior=None, use_camera_prior=False, body_prior_type='gmm',
                   prior_folder=gdir, num_gaussians=8, float_dtype='float64')
        for fr, tag in (('02_cropped', 'ref_fit_02_f64'), ('18_cropped', 'ref_fit_18_f64')):
            out = ref_fit(fr, inp, cfg=cfg, tag=tag, dtype=torch.float64, save=False)
            c = json.loads(str(out['cfg_json']))
            c.pop('prior_folder')
            out['cfg_json'] = np.array(json.dumps(c))
            np.savez_compressed(os.path.join(HERE, tag + '.npz'), **out)
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'collnf':
        nf = (('collnf', 0.1, False), ('nocoll', 0.0, True))
        ref_eval_coll(inp, torch.float64, 'f64', nf, 'ref_eval_collnf', clean_faces=True)
        ref_eval_coll(inp, torch.float32, 'f32', nf, 'ref_eval_collnf', clean_faces=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'coll':
        ref_eval_coll(inp, torch.float64, 'f64')
        ref_eval_coll(inp, torch.float32, 'f32')
        raise SystemExit(0)
    ref_eval(inp, torch.float64, 'f64')
    ref_eval(inp, torch.float32, 'f32')
    ref_stage(inp, torch.float64, 'f64')
    ref_eval_coll(inp, torch.float64, 'f64')
    ref_eval_coll(inp, torch.float32, 'f32')
    ref_metrics()
    ref_blending()
    ref_fit('02_cropped', inp)
    ref_fit('18_cropped', inp, cfg=cfg_smplifyx(), tag='ref_fit_18_vposer')
