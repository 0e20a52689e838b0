"""Shared helpers of the test-suite: seeded model, golden fixtures, packing between the
reference's named parameters and the engine's per-frame parameter vector."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

from smplifyx_b200 import _native as N          # noqa: E402
from smplifyx_b200 import synthetic             # noqa: E402
from smplifyx_b200 import utils as U            # noqa: E402

MODEL_KW = dict(num_betas=10, num_expression_coeffs=10, use_pca=True, num_pca_comps=12,
                flat_hand_mean=False, use_face_contour=True)


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def model_data(seed=0):
    return synthetic.cached_smplx_like(seed)


def joint_map(fmt='coco25', contour=True):
    return U.smpl_to_annotation('smplx', use_hands=True, use_face=True,
                                use_face_contour=contour, format=fmt)


def layout(use_vposer=False, n_hand=12):
    return N.make_layout(10, 10, n_hand, use_vposer)


def pack_params(L, named, cam_t=None, dtype=np.float64):
    """dict of named arrays (reference names) -> [np] vector in the engine layout."""
    x = np.zeros(L.np, dtype=dtype)
    for name, (off, n) in N.param_blocks(L).items():
        if name == 'camera_translation':
            if cam_t is not None:
                x[off:off + n] = np.asarray(cam_t, dtype=dtype).reshape(-1)
        elif name in named:
            x[off:off + n] = np.asarray(named[name], dtype=dtype).reshape(-1)
    return x


def unpack_params(L, x):
    return {name: np.asarray(x[off:off + n]) for name, (off, n) in N.param_blocks(L).items()}


def cam_row(focal, center, data_weight, tz_est=0.0, R=None, dtype=np.float64):
    row = np.zeros(N.SFX_CAM_STRIDE, dtype=dtype)
    row[0] = row[1] = focal
    row[2:4] = np.asarray(center).reshape(-1)
    row[4:13] = (np.eye(3) if R is None else np.asarray(R)).reshape(-1)
    row[13] = data_weight
    row[14] = tz_est
    return row


def eval_case_inputs(ev, case):
    """Engine-side inputs for one case of tests/golden/ref_eval_*.npz."""
    vposer = case.startswith('vposer')
    L = layout(use_vposer=vposer)
    named = {k[6:]: ev[k] for k in ev if k.startswith('param/')}
    if vposer:
        named['pose_embedding'] = ev['vposer/latent']
    x = pack_params(L, named, cam_t=ev['cam_t'])
    kp = ev['keypoints']
    H = int(ev['HW'][0])
    w = json.loads(str(ev['weights_json']))
    K = kp.shape[0]
    lowconf = np.zeros(K, np.uint8)
    lowconf[:25] = kp[:25, 2] < 0.2
    init_mask = np.zeros(K, np.uint8)
    init_mask[ev['init_idxs']] = 1
    cam = cam_row(float(ev['focal']), ev['center'], 1000.0 / H, tz_est=3.5)
    if case in ('l2', 'reg', 'gmm', 'vposer', 'vposer_reg'):
        st = N.make_stage(
            L, N.BODY_STAGE_BLOCKS, loss_kind=N.LOSS_SMPLIFY,
            pprior_kind={'reg': N.PPRIOR_REGRESSION, 'l2': N.PPRIOR_L2, 'gmm': N.PPRIOR_GMM,
                         'vposer': N.PPRIOR_LATENT, 'vposer_reg': N.PPRIOR_LATENT}[case],
            use_vposer=vposer, stage_index=2 if case == 'vposer_reg' else 1, num_stages=3,
            body_pose_weight=w['body_pose_weight'],
            shape_weight=w['shape_weight'], bending_prior_weight=w['bending_prior_weight'],
            hand_prior_weight=w['hand_prior_weight'], expr_prior_weight=w['expr_prior_weight'],
            jaw_prior_weight=w['jaw_prior_weight'], hand_joint_weight=0.1,
            face_joint_weight=2.0)
    else:
        st = N.make_stage(L, N.CAMERA_STAGE_BLOCKS, loss_kind=N.LOSS_CAMERA_INIT,
                          use_conf_camera=(case == 'camconf'), depth_loss_weight=100.0)
    # base joint weights: before the per-stage hand/face overwrite the engine applies itself
    jw_base = np.ones(K)
    jw_base[[1, 9, 12]] = 0
    jw_base[lowconf.astype(bool)] = 0
    reg_pose = ev['reg_pose'][0]
    if vposer:
        reg_pose = ev['vposer/latent_reg'][0] if case == 'vposer_reg' else None
    return dict(L=L, x=x, gt=kp[:, :2].copy(), conf=kp[:, 2].copy(), jw=jw_base,
                lowconf=lowconf, init_mask=init_mask, cam=cam, reg_pose=reg_pose,
                stage=st, named=named)


def vposer_weights():
    return synthetic.make_vposer_like(seed=2)


def golden_grad_vector(L, ev, case):
    """Reference gradients of one golden case packed in the engine layout."""
    named = {k[len(case) + 6:]: ev[k] for k in ev if k.startswith(case + '/grad/')}
    cam = named.pop('camera_translation', None)
    return pack_params(L, named, cam_t=cam)


def gmm_prior(dtype):
    """The synthetic 8 x 63 mixture (same seed as make_golden.py) as a product MaxMixturePrior."""
    from smplifyx_b200 import prior as P
    return P.MaxMixturePrior(num_gaussians=8, dtype=dtype,
                             gmm=synthetic.make_gmm_like(seed=1, num_gaussians=8, dim=63))


def gmm_arrays(dtype):
    import torch
    pr = gmm_prior(dtype)
    return (pr.means.numpy(), pr.precisions.numpy(),
            torch.log(pr.nll_weights).reshape(-1).numpy())


COLL_IGN = ["9,16", "9,17", "6,16", "6,17", "1,2", "12,22"]      # the shipped yaml files


def coll_segmentation():
    """(segm, parents, ign_part_pairs) of the synthetic model -- as tests/golden/make_golden.py
    builds them for the reference-side FilterFaces."""
    md = model_data()
    seg = synthetic.parts_segm_like(md)
    return seg['segm'], seg['parents'], COLL_IGN + synthetic.sibling_part_pairs(md)


def coll_case_inputs(ev, case):
    """Engine-side inputs for tests/golden/ref_eval_coll_*.npz (case 'coll' or 'nocoll')."""
    L = layout()
    named = {k[6:]: ev[k] for k in ev if k.startswith('param/')}
    x = pack_params(L, named, cam_t=ev['cam_t'])
    kp = ev['keypoints']
    H = int(ev['HW'][0])
    w = json.loads(str(ev['weights_json']))
    K = kp.shape[0]
    lowconf = np.zeros(K, np.uint8)
    lowconf[:25] = kp[:25, 2] < 0.2
    cam = cam_row(float(ev['focal']), ev['center'], 1000.0 / H, tz_est=3.5)
    jw_base = np.ones(K)
    jw_base[[1, 9, 12]] = 0
    jw_base[lowconf.astype(bool)] = 0
    st = N.make_stage(
        L, N.BODY_STAGE_BLOCKS, loss_kind=N.LOSS_SMPLIFY, pprior_kind=N.PPRIOR_L2, stage_index=1,
        num_stages=3, body_pose_weight=w['body_pose_weight'], shape_weight=w['shape_weight'],
        bending_prior_weight=w['bending_prior_weight'], hand_prior_weight=w['hand_prior_weight'],
        expr_prior_weight=w['expr_prior_weight'], jaw_prior_weight=w['jaw_prior_weight'],
        hand_joint_weight=0.1, face_joint_weight=2.0,
        coll_loss_weight=w['coll_loss_weight'] if case != 'nocoll' else 0.0,
        coll_sigma=float(ev['sigma']))
    return dict(L=L, x=x, gt=kp[:, :2].copy(), conf=kp[:, 2].copy(), jw=jw_base, lowconf=lowconf,
                init_mask=np.zeros(K, np.uint8), cam=cam, reg_pose=None, stage=st)
