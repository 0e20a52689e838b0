"""The reference's command line on the engine (smplifyx/main.py:51-323): a data folder with images
and OpenPose-style keypoint JSON, a model folder, the shipped YAML profile and regression results
go in; ``results/<name>/000.pkl``, ``vertices.ply`` and ``conf.yaml`` come out -- once with the
profile as shipped minus the interpenetration term, once with the term and ``--part_segm_fn``
(reference README.md:55)."""
import json
import os
import pickle

import numpy as np
import pytest

from tests import common as Cm

pytestmark = pytest.mark.gpu
ROOT = Cm.ROOT
FRAMES = ['02_cropped', '18_cropped']


def _write_inputs(tmp):
    """demo_inputs.npz (the reference's demo/ inputs) back into the on-disk layout main reads."""
    import cv2
    import joblib
    inp = Cm.golden('demo_inputs.npz')
    data = tmp / 'data'
    (data / 'images').mkdir(parents=True)
    (data / 'keypoints').mkdir()
    for fr in FRAMES:
        H, W = [int(v) for v in inp[fr + '/HW']]
        cv2.imwrite(str(data / 'images' / (fr + '.png')), np.zeros((H, W, 3), np.uint8))
        kp = inp[fr + '/keypoints']                      # body 25 | lhand 21 | rhand 21 | face 51 | contour 17
        face = np.zeros((70, 3), np.float32)
        face[17:68] = kp[67:118]
        face[:17] = kp[118:135]
        person = {'pose_keypoints_2d': kp[:25].reshape(-1).tolist(),
                  'hand_left_keypoints_2d': kp[25:46].reshape(-1).tolist(),
                  'hand_right_keypoints_2d': kp[46:67].reshape(-1).tolist(),
                  'face_keypoints_2d': face.reshape(-1).tolist()}
        with open(data / 'keypoints' / (fr + '_keypoints.json'), 'w') as f:
            json.dump({'people': [person]}, f)
        ex = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/expose/')}
        d = tmp / 'expose' / (fr + '.jpg')
        d.mkdir(parents=True)
        np.savez(str(d / (fr + '.jpg_params.npz')), **ex)
        px = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/pixie/')}
        d = tmp / 'pixie' / fr
        d.mkdir(parents=True)
        joblib.dump(px, str(d / (fr + '_param.pkl')))
    md = Cm.model_data()
    (tmp / 'models' / 'smplx').mkdir(parents=True)
    np.savez(str(tmp / 'models' / 'smplx' / 'SMPLX_NEUTRAL.npz'), **md)
    return data


def _run(tmp, out_name, extra):
    from smplifyx_b200 import main as M
    from smplifyx_b200.cmd_parser import parse_config
    data = tmp / 'data'
    out = tmp / out_name
    argv = ['-c', os.path.join(ROOT, 'cfg_files', 'fit_smplx_combined_coco25.yaml'),
            '--data_folder', str(data), '--output_folder', str(out),
            '--model_folder', str(tmp / 'models'), '--use_gender_classifier', 'False',
            '--visualize', 'False', '--interactive', 'False', '--batch_size', '2',
            '--expose_results_directory', str(tmp / 'expose'),
            '--pixie_results_directory', str(tmp / 'pixie')] + extra
    cfg = parse_config(argv)
    cfg.pop('config')
    M.main(**cfg)
    return out


def _check_outputs(out, want_keys):
    assert (out / 'conf.yaml').is_file()
    res = {}
    for fr in FRAMES:
        with open(out / 'results' / fr / '000.pkl', 'rb') as f:
            r = pickle.load(f)
        assert set(r.keys()) == want_keys
        assert all(np.all(np.isfinite(v)) for v in r.values() if isinstance(v, np.ndarray))
        raw = (out / 'results' / fr / 'vertices.ply').read_bytes()
        v = np.frombuffer(raw.split(b'end_header\n', 1)[1], dtype='<f4').reshape(-1, 3)
        assert v.shape == (10475, 3) and np.all(np.isfinite(v))
        res[fr] = (r, v)
    return res


def test_cli_end_to_end(tmp_path):
    ref = Cm.golden('ref_fit_02.npz')
    env = Cm.golden('ref_envelope.npz')
    want_keys = {k[7:] for k in ref if k.startswith('result/')}
    _write_inputs(tmp_path)
    out = _run(tmp_path, 'out_plain', ['--interpenetration', 'False'])
    res = _check_outputs(out, want_keys)
    # frame 02 against the reference's own fit of the same inputs (same bar as the driver tests)
    r, v = res['02_cropped']
    err = np.abs(v - ref['vertices'])
    assert err.max() <= float(env['fit/vertex_pairwise_max'])
    assert err.mean() <= float(env['fit/vertex_pairwise_mean'])
    assert r['body_pose'].shape == (1, 63) and r['betas'].dtype == np.float32

    # the profile as shipped (interpenetration True, no part_segm_fn) runs the term without
    # FilterFaces, as the reference does (fit_single_frame.py:317-328) -- on a mesh whose faces
    # all have area (the tube-man's area-less cap faces turn the reference's penalty into NaN too)
    (tmp_path / 'models_clean' / 'smplx').mkdir(parents=True)
    np.savez(str(tmp_path / 'models_clean' / 'smplx' / 'SMPLX_NEUTRAL.npz'),
             **Cm.synthetic.without_degenerate_faces(Cm.model_data()))
    out = _run(tmp_path, 'out_unfiltered', ['--model_folder', str(tmp_path / 'models_clean'),
                                            '--maxiters', '3'])
    res3 = _check_outputs(out, want_keys)
    assert not np.array_equal(res3['02_cropped'][1], res['02_cropped'][1])
    # ... and with the face segmentation
    segm, par, ign = Cm.coll_segmentation()
    seg_fn = tmp_path / 'parts_segm.pkl'
    with open(seg_fn, 'wb') as f:
        pickle.dump({'segm': segm, 'parents': par}, f, protocol=2)
    out = _run(tmp_path, 'out_coll', ['--part_segm_fn', str(seg_fn), '--maxiters', '6',
                                      '--ign_part_pairs'] + ign)
    res2 = _check_outputs(out, want_keys)
    assert not np.array_equal(res2['02_cropped'][1], res['02_cropped'][1])
