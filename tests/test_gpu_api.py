"""The reference-facing Python API (fitting.create_loss / FittingMonitor / create_optimizer /
create_camera / body_model.create / fit_single_frame) on the GPU, used the way the reference's
fit_single_frame uses it, checked against the golden vectors of the unmodified reference."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from tests import common as Cm

pytestmark = pytest.mark.gpu


def _objects(dtype, batch_size=1):
    from smplifyx_b200 import body_model as BM, utils as U, prior as P
    jm = Cm.joint_map()
    bm = BM.create(model_data=Cm.model_data(), joint_mapper=U.JointMapper(jm.astype(np.int64)),
                   dtype=dtype, batch_size=batch_size, **Cm.MODEL_KW)
    pri = dict(body_pose_prior=P.create_prior('l2'), jaw_prior=P.create_prior('l2'),
               expr_prior=P.create_prior('l2'), left_hand_prior=P.create_prior('l2'),
               right_hand_prior=P.create_prior('l2'), shape_prior=P.create_prior('l2'),
               angle_prior=P.create_prior('angle', dtype=dtype))
    return bm, pri


def test_closure_writes_reference_gradients():
    """create_fitting_closure()(backward=True): loss and .grad of every parameter equal the
    reference's autograd results (float64)."""
    from smplifyx_b200 import fitting, camera as C
    from smplifyx_b200.optimizers import optim_factory
    dtype = torch.float64
    ev = Cm.golden('ref_eval_f64.npz')
    bm, pri = _objects(dtype)
    dev = bm.engine_model.device
    H, W = [int(v) for v in ev['HW']]
    cam = C.create_camera(focal_length_x=float(ev['focal']), focal_length_y=float(ev['focal']),
                          dtype=dtype).to(dev)
    with torch.no_grad():
        cam.translation[:] = torch.tensor(ev['cam_t'], dtype=dtype)
        cam.center[:] = torch.tensor(ev['center'], dtype=dtype)
    named = {k[6:]: ev[k] for k in ev if k.startswith('param/')}
    emb = torch.tensor(named.pop('pose_embedding'), dtype=dtype, device=dev, requires_grad=True)
    bm.reset_params(**named)
    kp = torch.tensor(ev['keypoints'][None], dtype=dtype, device=dev)
    gt, conf = kp[:, :, :2], kp[:, :, 2]
    jw = torch.tensor(ev['jw'], dtype=dtype, device=dev)
    w = json.loads(str(ev['weights_json']))
    reg = torch.tensor(ev['reg_pose'], dtype=dtype, device=dev)
    loss = fitting.create_loss('smplify', rho=100, use_joints_conf=True, use_face=True,
                               use_hands=True, interpenetration=False, dtype=dtype,
                               regression_pose=reg, num_stages=3, **pri).to(dev)
    loss.reset_loss_weights(w)
    params = [p for p in bm.parameters() if p.requires_grad] + [emb]
    with fitting.FittingMonitor(maxiters=30, ftol=1e-9, gtol=1e-9) as monitor:
        opt, cg = optim_factory.create_optimizer(params, optim_type='lbfgsls', lr=1.0, maxiters=30)
        closure = monitor.create_fitting_closure(
            opt, bm, camera=cam, gt_joints=gt, joints_conf=conf, joint_weights=jw, loss=loss,
            create_graph=cg, use_vposer=False, vposer=None, pose_embedding=emb,
            return_verts=True, return_full_pose=True)
        val = closure(stage=1, backward=True)
        assert float(val) == pytest.approx(float(ev['reg/loss']), rel=1e-10)
        for name, p in bm.named_parameters():
            if name == 'body_pose':
                continue
            g = ev['reg/grad/' + name]
            assert np.abs(p.grad.cpu().numpy() - g).max() <= 1e-9 * max(1.0, np.abs(g).max()), name
        g = ev['reg/grad/pose_embedding']
        assert np.abs(emb.grad.cpu().numpy() - g).max() <= 1e-9 * np.abs(g).max()
        # SMPLifyLoss.forward on a body-model output: same value
        out = bm(return_verts=True, body_pose=emb, return_full_pose=True)
        v2 = loss(out, camera=cam, gt_joints=gt, body_model_faces=None, joints_conf=conf,
                  joint_weights=jw, pose_embedding=emb, use_vposer=False, stage=1)
        assert float(v2) == pytest.approx(float(ev['reg/loss']), rel=1e-10)
        assert np.abs(out.vertices.cpu().numpy()[0] - ev['vertices']).max() < 1e-10
        assert np.abs(out.joints.cpu().numpy()[0] - ev['l2/joints']).max() < 1e-10
        # run_fitting: one launch for the stage, parameters written back
        before = emb.detach().clone()
        final = monitor.run_fitting(opt, closure, params, bm, 1, pose_embedding=emb, vposer=None,
                                    use_vposer=False)
        sg = Cm.golden('ref_stage_f64.npz')
        assert abs(final - float(sg['final_loss'])) < 2e-3 * float(sg['final_loss'])
        assert float((emb.detach() - before).abs().max()) > 1e-3


def test_torch_optimizer_through_device_closure():
    """optim_type 'sgd' / 'lbfgs' / 'rmsprop' stay torch optimisers stepped from the host; every
    closure call is a device evaluation and the loss goes down."""
    from smplifyx_b200 import fitting, camera as C
    from smplifyx_b200.optimizers import optim_factory
    dtype = torch.float32
    ev = Cm.golden('ref_eval_f32.npz')
    bm, pri = _objects(dtype)
    dev = bm.engine_model.device
    cam = C.create_camera(focal_length_x=float(ev['focal']), focal_length_y=float(ev['focal']),
                          dtype=dtype).to(dev)
    with torch.no_grad():
        cam.translation[:] = torch.tensor(ev['cam_t'], dtype=dtype)
        cam.center[:] = torch.tensor(ev['center'], dtype=dtype)
    kp = torch.tensor(ev['keypoints'][None], dtype=dtype, device=dev)
    gt, conf = kp[:, :, :2], kp[:, :, 2]
    emb = torch.zeros([1, 63], dtype=dtype, device=dev, requires_grad=True)
    closs = fitting.create_loss('camera_init', trans_estimation=torch.tensor([[0., 0., 3.5]]),
                                init_joints_idxs=torch.tensor(ev['init_idxs']),
                                depth_loss_weight=100.0, dtype=dtype, joints_conf=conf).to(dev)
    closs.reset_loss_weights({'data_weight': 1000.0 / int(ev['HW'][0])})
    params = [cam.translation, bm.global_orient]
    with fitting.FittingMonitor(maxiters=5, ftol=0, gtol=0) as monitor:
        opt, cg = optim_factory.create_optimizer(params, optim_type='lbfgs', lr=1.0, maxiters=5)
        closure = monitor.create_fitting_closure(opt, bm, cam, gt, closs, create_graph=cg, use_vposer=False,
                                                 pose_embedding=emb, return_verts=False)
        first = float(closure(backward=False))
        last = monitor.run_fitting(opt, closure, params, bm, 0, use_vposer=False, pose_embedding=emb)
        assert last is not None and last < first


def test_fit_single_frame_mirror_against_reference(tmp_path):
    from smplifyx_b200 import camera as C, fit_single_frame as FSF
    inp = Cm.golden('demo_inputs.npz')
    ref = Cm.golden('ref_fit_02.npz')
    env = Cm.golden('ref_envelope.npz')
    cfg = json.loads(str(ref['cfg_json']))
    cfg['body_tri_idxs'] = [tuple(p) for p in cfg['body_tri_idxs']]
    fr = '02_cropped'
    expose = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/expose/')}
    pixie = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/pixie/')}
    H, W = [int(v) for v in inp[fr + '/HW']]
    dtype = torch.float32
    bm, pri = _objects(dtype)
    focal = (W ** 2 + H ** 2) ** 0.5
    cam = C.create_camera(focal_length_x=focal, focal_length_y=focal, dtype=dtype).to(
        bm.engine_model.device)
    cam.rotation.requires_grad = False
    jw = torch.ones([1, 135], dtype=dtype)
    jw[:, cfg['joints_to_ign']] = 0
    args = dict(cfg)
    for k in ('dtype', 'output_folder', 'result_folder', 'focal_length'):
        args.pop(k, None)
    res_fn = str(tmp_path / '000.pkl')
    FSF.fit_single_frame(np.zeros((H, W, 3), np.float32), inp[fr + '/keypoints'][None],
                         body_model=bm, camera=cam, joint_weights=jw, dtype=dtype,
                         result_folder=str(tmp_path), result_fn=res_fn, img_name=fr,
                         pixie_results=pixie, expose_results=expose, focal_length=focal,
                         **pri, **args)
    with open(res_fn, 'rb') as f:
        result = pickle.load(f)
    want_keys = {k[7:] for k in ref if k.startswith('result/')}
    assert set(result.keys()) == want_keys
    for k in want_keys:
        if isinstance(result[k], np.ndarray):
            assert result[k].shape == ref['result/' + k].shape, k
            assert result[k].dtype == ref['result/' + k].dtype, k
    raw = (tmp_path / 'vertices.ply').read_bytes()
    verts = np.frombuffer(raw.split(b'end_header\n', 1)[1], dtype='<f4').reshape(-1, 3)
    err = np.abs(verts - ref['vertices'])
    assert err.max() <= float(env['fit/vertex_pairwise_max'])
    assert err.mean() <= float(env['fit/vertex_pairwise_mean'])


def test_closure_with_interpenetration_objects():
    """fitting.create_loss(search_tree=, pen_distance=, tri_filtering_module=) the way the
    reference builds it (fit_single_frame.py:300-328, :413-428): the closure's loss and .grad
    equal the reference loss driving the restated mesh_intersection package (float64)."""
    from smplifyx_b200 import fitting, camera as C, mesh_intersection as MI
    from smplifyx_b200.optimizers import optim_factory
    dtype = torch.float64
    ev = Cm.golden('ref_eval_coll_f64.npz')
    bm, pri = _objects(dtype)
    dev = bm.engine_model.device
    cam = C.create_camera(focal_length_x=float(ev['focal']), focal_length_y=float(ev['focal']),
                          dtype=dtype).to(dev)
    with torch.no_grad():
        cam.translation[:] = torch.tensor(ev['cam_t'], dtype=dtype)
        cam.center[:] = torch.tensor(ev['center'], dtype=dtype)
    named = {k[6:]: ev[k] for k in ev if k.startswith('param/')}
    emb = torch.tensor(named.pop('pose_embedding'), dtype=dtype, device=dev, requires_grad=True)
    bm.reset_params(**named)
    kp = torch.tensor(ev['keypoints'][None], dtype=dtype, device=dev)
    gt, conf = kp[:, :, :2], kp[:, :, 2]
    jw = torch.tensor(ev['jw'], dtype=dtype, device=dev)
    w = json.loads(str(ev['weights_json']))
    segm, par, ign = Cm.coll_segmentation()
    search_tree = MI.BVH(max_collisions=128)
    pen_distance = MI.DistanceFieldPenetrationLoss(sigma=float(ev['sigma']), point2plane=False,
                                                   vectorized=True, penalize_outside=True)
    filter_faces = MI.FilterFaces(faces_segm=segm, faces_parents=par, ign_part_pairs=ign)
    loss = fitting.create_loss('smplify', rho=100, use_joints_conf=True, use_face=True,
                               use_hands=True, interpenetration=True, search_tree=search_tree,
                               pen_distance=pen_distance, tri_filtering_module=filter_faces,
                               dtype=dtype, regression_pose=None, num_stages=3, **pri).to(dev)
    loss.reset_loss_weights(w)
    params = [p for p in bm.parameters() if p.requires_grad] + [emb]
    with fitting.FittingMonitor(maxiters=30, ftol=1e-9, gtol=1e-9) as monitor:
        opt, cg = optim_factory.create_optimizer(params, optim_type='lbfgsls', lr=1.0, maxiters=30)
        closure = monitor.create_fitting_closure(
            opt, bm, camera=cam, gt_joints=gt, joints_conf=conf, joint_weights=jw, loss=loss,
            create_graph=cg, use_vposer=False, vposer=None, pose_embedding=emb,
            return_verts=True, return_full_pose=True)
        val = closure(stage=1, backward=True)
        assert float(val) == pytest.approx(float(ev['coll/loss']), rel=1e-11)
        g_pen = ev['coll/grad/pose_embedding'] - ev['nocoll/grad/pose_embedding']
        got = emb.grad.cpu().numpy() - ev['nocoll/grad/pose_embedding']
        assert np.abs(got - g_pen).max() <= 1e-8 * np.abs(g_pen).max()
        # weight 0 switches the term off (fitting.py:439)
        loss.reset_loss_weights(dict(w, coll_loss_weight=0.0))
        assert float(closure(stage=1, backward=False)) == pytest.approx(float(ev['nocoll/loss']), rel=1e-11)
    # without the face filter the term keeps every pair (reference fitting.py:449-450); without a
    # penalty object the device path refuses instead of dropping the term
    nofilter = fitting.create_loss('smplify', interpenetration=True, search_tree=search_tree,
                                   pen_distance=pen_distance, tri_filtering_module=None, dtype=dtype,
                                   num_stages=3, **pri).to(dev)
    nofilter.reset_loss_weights(w)
    assert nofilter.stage_kwargs(False, 1)['coll_loss_weight'] == w['coll_loss_weight']
    bad = fitting.create_loss('smplify', interpenetration=True, search_tree=search_tree,
                              pen_distance=None, tri_filtering_module=None, dtype=dtype,
                              num_stages=3, **pri).to(dev)
    bad.reset_loss_weights(w)
    with pytest.raises(NotImplementedError):
        bad.stage_kwargs(False, 1)
    with pytest.raises(NotImplementedError):
        MI.DistanceFieldPenetrationLoss(sigma=1e-4, point2plane=True)


def test_fit_single_frame_mirror_with_part_segm_fn(tmp_path):
    """The reference call with interpenetration=True and --part_segm_fn (README.md:55): the
    pickle is read, the term runs in the last two stages, the result keys are the reference's."""
    from smplifyx_b200 import camera as C, fit_single_frame as FSF, synthetic
    inp = Cm.golden('demo_inputs.npz')
    ref = Cm.golden('ref_fit_02.npz')
    cfg = json.loads(str(ref['cfg_json']))
    cfg['body_tri_idxs'] = [tuple(p) for p in cfg['body_tri_idxs']]
    segm, par, ign = Cm.coll_segmentation()
    seg_fn = str(tmp_path / 'parts_segm.pkl')
    with open(seg_fn, 'wb') as f:
        pickle.dump({'segm': segm, 'parents': par}, f, protocol=2)
    cfg.update(interpenetration=True, coll_loss_weights=[0.0, 0.1, 1.0], df_cone_height=1e-4,
               max_collisions=128, penalize_outside=True, point2plane=False, ign_part_pairs=ign,
               part_segm_fn=seg_fn, maxiters=6)
    fr = '02_cropped'
    expose = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/expose/')}
    pixie = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/pixie/')}
    H, W = [int(v) for v in inp[fr + '/HW']]
    dtype = torch.float32
    bm, pri = _objects(dtype)
    focal = (W ** 2 + H ** 2) ** 0.5
    cam = C.create_camera(focal_length_x=focal, focal_length_y=focal, dtype=dtype).to(
        bm.engine_model.device)
    cam.rotation.requires_grad = False
    jw = torch.ones([1, 135], dtype=dtype)
    jw[:, cfg['joints_to_ign']] = 0
    args = dict(cfg)
    for k in ('dtype', 'output_folder', 'result_folder', 'focal_length'):
        args.pop(k, None)
    res_fn = str(tmp_path / '000.pkl')
    FSF.fit_single_frame(np.zeros((H, W, 3), np.float32), inp[fr + '/keypoints'][None],
                         body_model=bm, camera=cam, joint_weights=jw, dtype=dtype,
                         result_folder=str(tmp_path), result_fn=res_fn, img_name=fr,
                         pixie_results=pixie, expose_results=expose, focal_length=focal,
                         **pri, **args)
    with open(res_fn, 'rb') as f:
        result = pickle.load(f)
    assert set(result.keys()) == {k[7:] for k in ref if k.startswith('result/')}
    assert all(np.all(np.isfinite(v)) for v in result.values() if isinstance(v, np.ndarray))
    batch = bm.frame_batch(False)
    cs = batch.coll_stats().cpu().numpy()
    assert cs[0, 0] > 0                     # the search ran and saw candidates
    assert int(batch.flags().cpu().numpy().max()) == 0


def test_parameter_vector_longer_than_the_kernel_limit_is_refused():
    """use_pca=False (45 + 45 hand components) with 26 shape / expression coefficients needs 194
    parameters > SFX_NP_MAX = 192: refused at batch creation instead of overrunning the
    kernel's shared arrays."""
    from smplifyx_b200 import engine
    kw = dict(Cm.MODEL_KW, use_pca=False, num_betas=16, num_expression_coeffs=10)
    model = engine.Model(Cm.model_data(), Cm.joint_map(), dtype=torch.float32, **kw)
    with pytest.raises(RuntimeError, match='SFX_NP_MAX'):
        engine.FrameBatch(model, 2)
    kw = dict(Cm.MODEL_KW, use_pca=False, num_betas=10, num_expression_coeffs=10)     # 188: fine
    model = engine.Model(Cm.model_data(), Cm.joint_map(), dtype=torch.float32, **kw)
    batch = engine.FrameBatch(model, 2)
    assert batch.L.np == 188
    verts, _ = batch.forward_mesh()
    assert bool(torch.isfinite(verts).all())


@pytest.mark.skipif(not os.path.isfile(os.environ.get('SFX_SMPLX_MODEL', '')),
                    reason='needs a licensed SMPLX_NEUTRAL.npz: set SFX_SMPLX_MODEL=<path>')
def test_real_model_known_answer_from_the_expose_fixture():
    """The one fixture of the reference that pins the un-vendored ``smplx`` package: ExPose's
    (pose, betas, expression) -> (vertices [10475,3], joints [144,3]) of demo frame 02
    (reference demo/ExPose_results/02_cropped.jpg/02_cropped.jpg_params.npz, copied to
    tests/golden/expose_02_known_answer.npz).  Runs only where the licensed model file is
    available; both float32 (tensor-core mesh path) and float64."""
    from smplifyx_b200 import body_model as BM, engine, utils as U
    K = Cm.golden('expose_02_known_answer.npz')
    md = BM.load_model(os.environ['SFX_SMPLX_MODEL'])
    aa = lambda R: np.concatenate([U.inv_rodrigues(np.asarray(r, np.float64)) for r in R])
    for dtype, tol in ((torch.float64, 2e-4), (torch.float32, 5e-4)):
        model = engine.Model(md, Cm.joint_map(), dtype=dtype, num_betas=10, num_expression_coeffs=10,
                             use_pca=False, flat_hand_mean=True, use_face_contour=True)
        batch = engine.FrameBatch(model, 1)
        L = batch.L
        x = Cm.pack_params(L, dict(global_orient=aa(K['global_orient']), betas=K['betas'],
                                   expression=K['expression'], jaw_pose=aa(K['jaw_pose']),
                                   left_hand_pose=aa(K['left_hand_pose']),
                                   right_hand_pose=aa(K['right_hand_pose']),
                                   pose_embedding=aa(K['body_pose'])))
        batch.set_params(x[None])
        verts, _ = batch.forward_mesh(want_joints=False)
        v = verts.cpu().numpy()[0]
        # ExPose stores the mesh before its camera translation; allow a rigid offset of the root
        d = v - K['vertices']
        d = d - d.mean(0, keepdims=True)
        print('ExPose known answer, %s: max vertex deviation %.3g m' % (dtype, np.abs(d).max()))
        assert np.abs(d).max() < tol
