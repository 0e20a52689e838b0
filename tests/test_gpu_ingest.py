"""Device ingestion and blending kernels (csrc/sfx_ingest.cuh) against the
host mirrors, which tests/test_oracle.py and tests/test_keypoints_blending.py pin to the
unmodified reference: bit-identical on seeded inputs and on the reference's demo detections."""
import numpy as np
import pytest
import torch

from tests import common as Cm

pytestmark = pytest.mark.gpu


def test_pack_keypoints_matches_the_reader_row_order():
    from smplifyx_b200 import data_parser as DP
    rng = np.random.default_rng(3)
    B = 5
    body = rng.normal(size=(B, 25, 3)).astype(np.float32)
    lh = rng.normal(size=(B, 21, 3)).astype(np.float32)
    rh = rng.normal(size=(B, 21, 3)).astype(np.float32)
    face = rng.normal(size=(B, 70, 3)).astype(np.float32)
    for contour in (True, False):
        got = DP.pack_keypoints_device(body, lh, rh, face, use_face_contour=contour).cpu().numpy()
        rows = [body, lh, rh, face[:, 17:68]] + ([face[:, :17]] if contour else [])
        assert np.array_equal(got, np.concatenate(rows, axis=1))
    # halpe body block (26 rows)
    body26 = rng.normal(size=(B, 26, 3)).astype(np.float32)
    got = DP.pack_keypoints_device(body26, lh, rh, face, True).cpu().numpy()
    assert got.shape == (B, 136, 3) and np.array_equal(got[:, :26], body26)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_device_masks_equal_host_masks_and_fit_is_unchanged(dtype):
    """sfx_keypoint_masks == fit_frames.keypoint_masks (fit_single_frame.py:276-294) bit for bit,
    on the demo detections and on seeded ones with zero coordinates / low confidences; a fit fed
    through the device path ends at exactly the same parameters."""
    import json
    from smplifyx_b200 import engine, fit_frames as FF
    inp = Cm.golden('demo_inputs.npz')
    ref = Cm.golden('ref_fit_02.npz')
    cfg = json.loads(str(ref['cfg_json']))
    cfg['maxiters'] = 2
    rng = np.random.default_rng(5)
    kp = np.stack([inp['02_cropped/keypoints'], inp['18_cropped/keypoints'],
                   inp['02_cropped/keypoints'] * (rng.uniform(size=(135, 3)) > 0.2)]).astype(np.float32)
    model = engine.Model(Cm.model_data(), Cm.joint_map(), dtype=dtype, **Cm.MODEL_KW)
    frames = ['02_cropped', '18_cropped', '02_cropped']
    ex = [{k.split('/')[-1]: inp[k] for k in inp if k.startswith(f + '/expose/')} for f in frames]
    px = [{k.split('/')[-1]: inp[k] for k in inp if k.startswith(f + '/pixie/')} for f in frames]
    outs = []
    for dev_ingest in (False, True):
        batch = engine.FrameBatch(model, 3)
        c = dict(cfg, device_ingest=dev_ingest, two_loop='exact')
        plan = FF.FitPlan(batch.L, model.K, kp, 600, 800, c, ex, px, None, model.np_dtype)
        FF.upload(batch, plan)
        if dev_ingest:
            kp_d, gt, conf, jw, low, init, _, _ = batch._keep_dev
            assert np.array_equal(low.cpu().numpy(), plan.lowconf)
            assert np.array_equal(init.cpu().numpy(), plan.init_mask)
            assert np.array_equal(jw.cpu().numpy(), plan.jw.astype(model.np_dtype))
            assert np.array_equal(gt.cpu().numpy(), plan.keypoints[:, :, :2].astype(model.np_dtype))
            assert np.array_equal(conf.cpu().numpy(), plan.keypoints[:, :, 2].astype(model.np_dtype))
            assert plan.lowconf.sum() > 0 and plan.init_mask.sum() > 0
        cam_loss, verts, joints, _ = FF.run(batch, plan, False)
        outs.append(FF.download(batch, plan, cam_loss, verts, joints))
    assert np.array_equal(outs[0].params, outs[1].params)
    assert np.array_equal(outs[0].n_evals, outs[1].n_evals)


def test_blend_keypoints_kernel_equals_host_blending():
    """sfx_blend_keypoints == keypoints_blending.blend_keypoints, itself bit-identical to the
    unmodified reference function (tests/golden/ref_blending.npz)."""
    from smplifyx_b200 import keypoints_blending as KB
    import json
    G = Cm.golden('ref_blending.npz')
    stats = json.loads(str(G['stats_json']))
    rng = np.random.default_rng(9)
    # the committed reference vectors first: device == the unmodified reference's output
    dev0 = KB.blend_keypoints_device(G['openpose'], G['mmpose'], stats).cpu().numpy()
    assert np.array_equal(dev0.astype(np.float64), G['blended'])
    B = 7
    op = rng.uniform(0, 800, size=(B, 135, 3)).astype(np.float32)
    mm = rng.uniform(0, 800, size=(B, 136, 3)).astype(np.float32)
    op[..., 2] = rng.uniform(-0.2, 1.3, size=(B, 135))
    mm[..., 2] = rng.uniform(-0.2, 1.3, size=(B, 136))
    host = KB.blend_keypoints(op, mm, stats)
    dev = KB.blend_keypoints_device(op, mm, stats).cpu().numpy()
    assert np.array_equal(dev.astype(np.float64), host)
    assert (host[:, :67, 2] != np.clip(op[:, :67, 2], 0, 1)).any()      # MMPose won somewhere
