"""Host-side logic of the drop-in surface (no GPU): config parsing, stage schedules, keypoint
masks, regression-prior pose, camera prior, orientation flip, result writers, sharding."""
import glob
import json
import os
import struct

import numpy as np
import pytest
import torch

from tests import common as Cm
from smplifyx_b200 import _native as N
from smplifyx_b200 import fit_frames as FF
from smplifyx_b200 import utils as U
from smplifyx_b200.cmd_parser import parse_config

CFG_DIR = os.path.join(Cm.ROOT, 'cfg_files')


@pytest.mark.parametrize('fn', sorted(os.path.basename(p) for p in glob.glob(os.path.join(CFG_DIR, '*.yaml'))))
def test_every_shipped_profile_parses_and_schedules(fn):
    cfg = parse_config(['-c', os.path.join(CFG_DIR, fn)])
    assert isinstance(cfg['ftol'], float) and isinstance(cfg['maxiters'], int)
    assert all(isinstance(p, tuple) and len(p) == 2 for p in cfg['body_tri_idxs'])
    w = FF.stage_weights(cfg)
    assert len(w) == len(cfg['body_pose_prior_weights'])
    assert all(len(x['jaw_prior_weight']) == 3 for x in w)
    if not cfg['use_vposer'] and cfg['body_prior_type'] == 'l2':
        cam, stages = FF.make_stages(cfg, Cm.layout(n_hand=cfg['num_pca_comps']))
        assert cam.loss_kind == N.LOSS_CAMERA_INIT and cam.n_active == 6
        for i, st in enumerate(stages):
            assert st.stage_index == i and st.num_stages == len(stages)
            assert st.bending_prior_weight == pytest.approx(3.17 * st.body_pose_weight)
            assert st.opt_kind == N.OPT_LBFGSLS and st.max_iter == cfg['maxiters']
            assert st.max_eval == cfg['maxiters'] * 5 // 4


def test_cli_overrides_yaml(tmp_path):
    cfg = parse_config(['-c', os.path.join(CFG_DIR, 'fit_smplx_combined_coco25.yaml'),
                        '--maxiters', '7', '--use_vposer', 'True', '--data_folder', 'D',
                        '--joints_to_ign', '3', '4'])
    assert cfg['maxiters'] == 7 and cfg['use_vposer'] is True and cfg['data_folder'] == 'D'
    assert cfg['joints_to_ign'] == [3, 4]
    assert cfg['rho'] == 100 and cfg['regression_prior'] == 'combined'
    bad = tmp_path / 'bad.yaml'
    bad.write_text('body_tri_idxs: [1, 2, 3]\n')
    with pytest.raises(AssertionError):
        parse_config(['-c', str(bad)])


def test_stage_weight_validation():
    cfg = dict(body_pose_prior_weights=[1, 2, 3], shape_weights=[1, 2])
    with pytest.raises(AssertionError):
        FF.stage_weights(cfg)
    w = FF.stage_weights(dict(body_pose_prior_weights=[1, 2, 3, 4]))
    assert [x['shape_weight'] for x in w] == [1e2, 5e1, 1e1, 5.0]
    assert w[2]['jaw_prior_weight'] == [1e1] * 3         # defaults to the shape weights


def test_keypoint_masks_follow_reference_rules():
    inp = Cm.golden('demo_inputs.npz')
    kp = np.stack([inp['02_cropped/keypoints'], inp['18_cropped/keypoints']]).astype(np.float64)
    cfg = dict(format='coco25', confidence_threshold=0.2, joints_to_ign=[1, 9, 12],
               init_joints_idxs=[0, 1, 2, 3, 5, 6, 8, 9, 12, 15, 16, 17, 18])
    jw, low, init = FF.keypoint_masks(kp, cfg, FF.base_joint_weights(cfg, 135))
    for b in range(2):
        exp_low = np.zeros(135, bool)
        exp_low[:25] = kp[b, :25, 2] < 0.2
        assert np.array_equal(low[b].astype(bool), exp_low)       # hands / face never "low"
        assert np.all(jw[b][exp_low] == 0) and np.all(jw[b][[1, 9, 12]] == 0)
        idx = [i for i in cfg['init_joints_idxs']
               if kp[b, i, 0] != 0 and kp[b, i, 1] != 0 and not exp_low[i]]
        assert list(np.flatnonzero(init[b])) == sorted(idx)


def test_regression_pose_and_camera_prior():
    from oracle import fit_port as FP
    inp = Cm.golden('demo_inputs.npz')
    fr = '18_cropped'
    expose = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/expose/')}
    pixie = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/pixie/')}
    for kind in ('ExPose', 'PIXIE', 'combined'):
        pose, go = FF.regression_pose(dict(regression_prior=kind), expose, pixie)
        e = [FP.euler_xyz_from_matrix(torch.tensor(m)) for m in expose['body_pose']]
        p = [FP.euler_xyz_from_matrix(torch.tensor(m)) for m in pixie['body_pose']]
        want = {'ExPose': e, 'PIXIE': p, 'combined': e[:19] + p[19:]}[kind]
        assert np.allclose(pose, torch.cat(want).reshape(-1).numpy(), atol=2e-6)
        src = pixie['global_pose'] if kind == 'PIXIE' else expose['global_orient']
        assert np.allclose(go, FP.euler_xyz_from_matrix(torch.tensor(src[0])).numpy()[0], atol=2e-6)
    # the batched planner gives exactly what per-frame calls give
    fr2 = '02_cropped'
    expose2 = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr2 + '/expose/')}
    pixie2 = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr2 + '/pixie/')}
    for kind in ('ExPose', 'PIXIE', 'combined'):
        cfg = dict(regression_prior=kind)
        pb, gb = FF.regression_pose_batch(cfg, [expose, expose2, expose], [pixie, pixie2, pixie], 3)
        for b, (e_, p_) in enumerate([(expose, pixie), (expose2, pixie2), (expose, pixie)]):
            pose, go = FF.regression_pose(cfg, e_, p_)
            assert np.array_equal(pb[b], pose) and np.array_equal(gb[b], go)
    t, c = FF.camera_prior(dict(regression_prior='combined', use_camera_prior=True), 1000.0,
                           expose, pixie)
    assert np.allclose(t[:2], expose['transl'][:2]) and t[2] == pytest.approx(expose['transl'][2] / 5.0)
    assert np.allclose(c, expose['center'])
    t, c = FF.camera_prior(dict(regression_prior='PIXIE', use_camera_prior=True), 1000.0,
                           expose, pixie)
    l, tp, r, b = [float(v) for v in pixie['bbox']]
    size = int(max(r - l, b - tp) * 1.1)
    assert t[2] == pytest.approx(2000.0 / (float(pixie['body_cam'][0]) * size + 1e-9))
    assert FF.camera_prior(dict(regression_prior=None, use_camera_prior=True), 1000.0) is None


def test_pare_regression_and_camera_prior():
    """regression_prior 'PARE' (fit_single_frame.py:220-231, :360-369): body pose from
    pred_pose[:, 1:22], orientation from pred_pose[:, :1], camera from the 224-pixel crop."""
    from oracle import fit_port as FP
    rng = np.random.default_rng(5)

    def rotmats(n):
        q, r = np.linalg.qr(rng.normal(size=(n, 3, 3)))
        q = q * np.sign(np.diagonal(r, axis1=1, axis2=2))[:, None, :]
        q[np.linalg.det(q) < 0, :, 0] *= -1
        return q.astype(np.float32)
    pares = [dict(pred_pose=rotmats(24)[None], bboxes=np.array([[410.0, 305.0, 380.0, 380.0]]),
                  pred_cam=np.array([[0.9, 0.02, -0.05]], dtype=np.float32)) for _ in range(3)]
    cfg = dict(regression_prior='PARE', use_camera_prior=True)
    for p in pares:
        pose, go = FF.regression_pose(cfg, None, None, pare=p)
        want = torch.cat([FP.euler_xyz_from_matrix(torch.tensor(m)) for m in p['pred_pose'][0, 1:22]])
        assert pose.shape == (63,) and np.allclose(pose, want.reshape(-1).numpy(), atol=2e-6)
        assert np.allclose(go, FP.euler_xyz_from_matrix(torch.tensor(p['pred_pose'][0, :1])).numpy()[0],
                           atol=2e-6)
        t, c = FF.camera_prior(cfg, 1000.0, pare=p)
        assert np.allclose(c, [410.0, 305.0])
        assert np.allclose(t, [0.02, -0.05, 2 * 1000.0 / (380.0 * 0.9)], rtol=1e-6)
    pb, gb = FF.regression_pose_batch(cfg, None, None, 3, pare=pares)
    for b, p in enumerate(pares):
        pose, go = FF.regression_pose(cfg, None, None, pare=p)
        assert np.array_equal(pb[b], pose) and np.array_equal(gb[b], go)
    # and through the plan: initial pose, orientation, camera row
    L = Cm.layout()
    kp = np.zeros((3, 135, 3))
    kp[:, :, :2] = rng.uniform(100, 500, size=(3, 135, 2))
    kp[:, :, 2] = 0.9
    plan_cfg = dict(cfg, format='coco25', use_hands=True, use_face=True, joints_to_ign=[1, 9, 12],
                    body_pose_prior_weights=[1.0], shape_weights=[1.0], expr_weights=[1.0],
                    hand_pose_prior_weights=[1.0], hand_joints_weights=[1.0],
                    face_joints_weights=[1.0], coll_loss_weights=[0.0],
                    jaw_pose_prior_weights=['1,1,1'])
    plan = FF.FitPlan(L, 135, kp, 600, 800, plan_cfg, None, None, None, np.float32, pare=pares)
    assert np.allclose(plan.x0[:, L.off_pose:L.off_pose + 63], pb)
    assert np.allclose(plan.x0[:, L.off_go:L.off_go + 3], gb)
    assert np.allclose(plan.cam[:, 2:4], [410.0, 305.0]) and not plan.need_guess
    assert np.allclose(plan.x0[:, L.off_camt + 2], 2 * 1000.0 / (380.0 * 0.9), rtol=1e-6)


def test_plan_carries_the_interpenetration_settings():
    """The shipped profile's interpenetration keys reach the stages (coll weights per stage,
    df_cone_height) and the face filter; without a segmentation the plan carries the reference's
    unfiltered term (fit_single_frame.py:317-328: filter_faces = None)."""
    from smplifyx_b200.cmd_parser import parse_config
    cfg = parse_config(['-c', os.path.join(Cm.ROOT, 'cfg_files', 'fit_smplx_combined_coco25.yaml')])
    cfg.pop('config')
    cfg.update(regression_prior=None, use_camera_prior=False)
    L = Cm.layout()
    kp = np.zeros((2, 135, 3))
    kp[:, :, :2] = 300.0
    kp[:, :, 2] = 0.9
    plan = FF.FitPlan(L, 135, kp, 600, 800, cfg, None, None, None, np.float32)
    assert plan.collision == 'unfiltered' and [st.coll_loss_weight for st in plan.stages] == [0.0, 0.1, 1.0]
    segm, par, ign = Cm.coll_segmentation()
    plan = FF.FitPlan(L, 135, kp, 600, 800, cfg, None, None, None, np.float32,
                      part_segm={'segm': segm, 'parents': par})
    assert [st.coll_loss_weight for st in plan.stages] == [0.0, 0.1, 1.0]
    assert all(st.coll_sigma == 1e-4 for st in plan.stages) and plan.cam_stage.coll_loss_weight == 0
    assert plan.collision.ign_part_pairs == cfg['ign_part_pairs']
    assert plan.collision.faces_segm.dtype == np.int32 and len(plan.collision.faces_segm) == 20908
    off = dict(cfg, interpenetration=False)
    plan = FF.FitPlan(L, 135, kp, 600, 800, off, None, None, None, np.float32)
    assert plan.collision is None and all(st.coll_loss_weight == 0 for st in plan.stages)
    with pytest.raises(NotImplementedError):
        FF.FitPlan(L, 135, kp, 600, 800, dict(cfg, point2plane=True), None, None, None, np.float32,
                   part_segm={'segm': segm, 'parents': par})


def test_flipped_orientation_matches_cv2():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(5)
    cases = [rng.normal(size=3) for _ in range(20)] + [np.zeros(3), np.array([0, 1e-9, 0]),
                                                       np.array([0., np.pi, 0.]), np.array([np.pi, 0, 0])]
    for go in cases:
        want = cv2.Rodrigues(cv2.Rodrigues(go.reshape(3, 1))[0].dot(
            cv2.Rodrigues(np.array([0., np.pi, 0]))[0]))[0].ravel()
        got = FF.flipped_orientation(go)
        Rw, Rg = cv2.Rodrigues(want)[0], cv2.Rodrigues(got)[0]
        assert np.allclose(Rw, Rg, atol=1e-9)
        if np.linalg.norm(want) < np.pi - 1e-3:
            assert np.allclose(want, got, atol=1e-7)


def test_guess_init_depth_formula():
    rng = np.random.default_rng(1)
    j3 = rng.normal(size=(3, 20, 3))
    g2 = rng.normal(size=(3, 20, 2)) * 50
    t = FF.guess_init_depth(j3, g2, [(5, 12), (2, 9)], 1000.0)
    for b in range(3):
        l3 = np.mean([np.linalg.norm(j3[b, 5] - j3[b, 12]), np.linalg.norm(j3[b, 2] - j3[b, 9])])
        l2 = np.mean([np.linalg.norm(g2[b, 5] - g2[b, 12]), np.linalg.norm(g2[b, 2] - g2[b, 9])])
        assert t[b, 2] == pytest.approx(1000.0 * l3 / l2) and t[b, 0] == 0 and t[b, 1] == 0


def test_fit_plan_matches_reference_initialisation():
    """Initial parameters / camera of the plan == what the reference set up for demo frame 02
    (golden result keeps camera_center; the start translation is the ExPose prior)."""
    inp = Cm.golden('demo_inputs.npz')
    ref = Cm.golden('ref_fit_02.npz')
    cfg = json.loads(str(ref['cfg_json']))
    fr = '02_cropped'
    expose = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/expose/')}
    pixie = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/pixie/')}
    L = Cm.layout()
    plan = FF.FitPlan(L, 135, inp[fr + '/keypoints'][None], 600, 800, cfg, [expose], [pixie])
    assert np.allclose(plan.cam[0, 2:4], ref['result/camera_center'][0])
    assert plan.cam[0, N.SFX_CAM_DW] == pytest.approx(1000 / 600)
    assert plan.cam[0, 0] == pytest.approx(1000.0)
    assert plan.cam[0, N.SFX_CAM_TZ] == pytest.approx(float(expose['transl'][2]) / 5.0, rel=1e-6)
    assert len(plan.stages) == 3 and len(plan.flip_ids) == 0 and not plan.need_guess
    assert np.all(plan.x0[0, L.off_betas:L.off_betas + 10] == 0)
    assert np.abs(plan.x0[0, L.off_pose:L.off_pose + 63]).max() > 0.1
    assert plan.stages[2].pprior_kind == N.PPRIOR_REGRESSION


def test_ply_writer_roundtrip(tmp_path):
    from smplifyx_b200.fit_single_frame import write_ply_vertices
    v = np.random.default_rng(0).normal(size=(10475, 3)).astype(np.float32)
    p = tmp_path / 'vertices.ply'
    write_ply_vertices(str(p), v)
    raw = p.read_bytes()
    head, body = raw.split(b'end_header\n', 1)
    assert b'format binary_little_endian 1.0' in head and b'element vertices 10475' in head
    assert head.count(b'property float') == 3
    assert np.array_equal(np.frombuffer(body, dtype='<f4').reshape(-1, 3), v)


def test_joint_maps_have_reference_sizes():
    assert len(U.smpl_to_annotation('smplx', True, True, True, 'coco25')) == 135
    assert len(U.smpl_to_annotation('smplx', True, True, False, 'coco25')) == 118
    assert len(U.smpl_to_annotation('smplx', True, True, True, 'halpe')) == 136
    assert len(U.smpl_to_annotation('smplx', True, True, True, 'coco_wholebody')) == 133
    with pytest.raises(ValueError):
        U.smpl_to_annotation('smplx', format='mpii')


def test_shard_ranges_partition_the_frames():
    from smplifyx_b200 import sharding
    for n in (0, 1, 7, 128, 4096, 4099):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                rg = sharding.shard_range(n, r, world)
                seen += list(rg)
                assert abs(len(rg) - n / world) < 1
            assert seen == list(range(n))


def test_two_loop_option_reaches_the_stages(monkeypatch):
    """The engine-only key ``two_loop`` (yaml / CLI / SFX_TWO_LOOP) selects SfxStage's
    L-BFGS-direction mode for the camera stage and every annealing stage; unknown names raise."""
    from smplifyx_b200 import _native as N
    from smplifyx_b200.cmd_parser import parse_config
    cfg_file = os.path.join(Cm.ROOT, 'cfg_files', 'fit_smplx_combined_coco25.yaml')
    cfg = parse_config(['-c', cfg_file])
    assert cfg['two_loop'] is None
    L = Cm.layout()
    monkeypatch.delenv('SFX_TWO_LOOP', raising=False)
    cam, stages = FF.make_stages(dict(cfg, interpenetration=False), L)
    assert cam.generic_two_loop == 0 and all(st.generic_two_loop == 0 for st in stages)
    cfg = parse_config(['-c', cfg_file, '--two_loop', 'gram'])
    cam, stages = FF.make_stages(dict(cfg, interpenetration=False), L)
    assert cam.generic_two_loop == 2 and all(st.generic_two_loop == 2 for st in stages)
    monkeypatch.setenv('SFX_TWO_LOOP', 'gram')
    assert N.make_stage(L, N.BODY_STAGE_BLOCKS).generic_two_loop == 2
    assert N.make_stage(L, N.BODY_STAGE_BLOCKS, two_loop='exact').generic_two_loop == 0
    with pytest.raises(ValueError):
        N.make_stage(L, N.BODY_STAGE_BLOCKS, two_loop='fast')


def test_degenerate_inputs_at_plan_level():
    """Empty batches are refused with a clear message; frames without a single detected
    keypoint still plan (no init joints, both orientations because the shoulders coincide,
    finite start) -- the reference would run them too (fit_single_frame.py:285-294, :477-480)."""
    import bench
    cfg = dict(bench.bench_cfg(), regression_prior=None, use_camera_prior=False)
    L = Cm.layout()
    with pytest.raises(ValueError, match='at least one frame'):
        FF.FitPlan(L, 135, np.zeros((0, 135, 3)), 600, 800, cfg, None, None, None, np.float32)
    plan = FF.FitPlan(L, 135, np.zeros((2, 135, 3)), 600, 800, cfg, None, None, None, np.float32)
    assert list(plan.flip_ids) == [0, 1]
    assert np.isfinite(plan.x0).all()


def test_dataset_read_meta_equals_read_item_without_pixels(tmp_path):
    """The batched driver reads keypoints + image size only (``KeypointDataset.read_meta``): same
    fields as ``read_item`` except the decoded pixels, whose shape it reports."""
    import json
    import cv2
    from smplifyx_b200 import data_parser as DP
    (tmp_path / 'images').mkdir()
    (tmp_path / 'keypoints').mkdir()
    rng = np.random.default_rng(0)
    for name, (h, w) in (('a', (48, 64)), ('b', (30, 20))):
        cv2.imwrite(str(tmp_path / 'images' / (name + '.png')), np.zeros((h, w, 3), np.uint8))
        person = {'pose_keypoints_2d': rng.uniform(size=75).tolist(),
                  'hand_left_keypoints_2d': rng.uniform(size=63).tolist(),
                  'hand_right_keypoints_2d': rng.uniform(size=63).tolist(),
                  'face_keypoints_2d': rng.uniform(size=210).tolist()}
        with open(tmp_path / 'keypoints' / (name + '_keypoints.json'), 'w') as f:
            json.dump({'people': [person]}, f)
    cv2.imwrite(str(tmp_path / 'images' / 'c.png'), np.zeros((8, 8, 3), np.uint8))   # no keypoints
    ds = DP.create_dataset(data_folder=str(tmp_path), use_hands=True, use_face=True,
                           use_face_contour=True)
    metas = [ds.read_meta(p) for p in ds.img_paths]
    items = [ds.read_item(p) for p in ds.img_paths]
    assert [bool(m) for m in metas] == [bool(i) for i in items] == [True, True, False]
    for m, i in zip(metas[:2], items[:2]):
        assert m['fn'] == i['fn'] and np.array_equal(m['keypoints'], i['keypoints'])
        assert (m['H'], m['W']) == i['img'].shape[:2]


def test_lazy_result_dicts():
    """FitResult.results builds the reference's result dict (fit_single_frame.py:644-660) of a
    frame when it is asked for: list-like (len, negative index, slice, iteration), cached, with
    the parameter blocks cut out of the fitted rows."""
    inp = Cm.golden('demo_inputs.npz')
    ref = Cm.golden('ref_fit_02.npz')
    cfg = json.loads(str(ref['cfg_json']))
    fr = '02_cropped'
    expose = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/expose/')}
    pixie = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/pixie/')}
    L = Cm.layout()
    kp = np.stack([inp[fr + '/keypoints']] * 3)
    plan = FF.FitPlan(L, 135, kp, [600, 610, 620], 800, cfg, [expose] * 3, [pixie] * 3)
    params = np.arange(3 * L.np, dtype=np.float32).reshape(3, L.np)
    blocks = {'betas': (L.off_betas, L.n_betas), 'global_orient': (L.off_go, 3),
              'left_hand_pose': (L.off_lh, L.n_hand), 'right_hand_pose': (L.off_rh, L.n_hand),
              'jaw_pose': (L.off_jaw, 3), 'leye_pose': (L.off_leye, 3), 'reye_pose': (L.off_reye, 3),
              'expression': (L.off_expr, L.n_expr), 'pose_embedding': (L.off_pose, L.n_pose)}
    res = FF._LazyResults(plan, blocks, params, None)
    assert len(res) == 3 and not res._cache
    r = res[1]
    assert res[1] is r and res[-2] is r and list(res._cache) == [1]
    assert r['H'] == 610 and r['W'] == 800 and r['betas'].shape == (1, L.n_betas)
    assert np.array_equal(r['body_pose'][0], params[1, L.off_pose:L.off_pose + L.n_pose])
    assert np.array_equal(r['camera_translation'][0], params[1, L.off_camt:L.off_camt + 3])
    assert [x['H'] for x in res] == [600, 610, 620] and [x['H'] for x in res[1:]] == [610, 620]
    with pytest.raises(IndexError):
        res[3]
    decoded = np.ones((3, 63), dtype=np.float32)
    assert np.array_equal(FF._LazyResults(plan, blocks, params, decoded)[2]['body_pose'], decoded[2:3])
