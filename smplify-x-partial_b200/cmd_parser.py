"""CLI / YAML surface of the reference (``smplifyx/cmd_parser.py:27-317``), table driven.

Same flag names, types, defaults and post-processing (``body_tri_idxs`` pairs,
cmd_parser.py:307-316), same precedence (command line > YAML > default).  ``configargparse``
is not installed in this image, so the YAML file given by ``-c/--config`` is read with
``yaml`` and merged by hand; keys of the YAML file are exactly the flag names, as in the
reference's ``cfg_files/*.yaml``.
"""
import argparse
import os

import yaml


def _b(x):            # reference bool flags: ``lambda arg: arg.lower() == 'true'`` / in ['true','1']
    if isinstance(x, bool):
        return x
    return str(x).lower() in ('true', '1')


# (name, type, default, nargs)
_FLAGS = [
    ('data_folder', str, None, None), ('max_persons', int, 3, None),
    ('loss_type', str, 'smplify', None), ('interactive', _b, False, None),
    ('save_meshes', _b, True, None), ('visualize', _b, False, None),
    ('degrees', float, [0, 90, 180, 270], '*'), ('use_cuda', _b, True, None),
    ('format', str, 'coco_wholebody', None), ('joints_to_ign', int, -1, '*'),
    ('output_folder', str, 'output', None), ('img_folder', str, 'images', None),
    ('keyp_folder', str, 'keypoints', None), ('summary_folder', str, 'summaries', None),
    ('result_folder', str, 'results', None), ('mesh_folder', str, 'meshes', None),
    ('gender', str, 'neutral', None), ('float_dtype', str, 'float32', None),
    ('model_type', str, 'smpl', None), ('camera_type', str, 'persp', None),
    ('optim_jaw', _b, True, None), ('optim_hands', _b, True, None),
    ('optim_expression', _b, True, None), ('optim_shape', _b, True, None),
    ('model_folder', str, 'models', None), ('use_joints_conf', _b, True, None),
    ('batch_size', int, 1, None), ('num_gaussians', int, 8, None),
    ('use_pca', _b, True, None), ('num_pca_comps', int, 6, None),
    ('flat_hand_mean', _b, False, None), ('body_prior_type', str, 'mog', None),
    ('left_hand_prior_type', str, 'mog', None), ('right_hand_prior_type', str, 'mog', None),
    ('jaw_prior_type', str, 'l2', None), ('use_vposer', _b, False, None),
    ('vposer_ckpt', str, '', None), ('init_joints_idxs', int, [9, 12, 2, 5], '*'),
    ('body_tri_idxs', int, [5, 12, 2, 9], '*'), ('prior_folder', str, 'prior', None),
    ('focal_length', float, None, None), ('rho', float, 100, None),
    ('interpenetration', _b, False, None), ('penalize_outside', _b, False, None),
    ('data_weights', float, None, '*'),
    ('body_pose_prior_weights', float, [4.04e2, 4.04e2, 57.4, 4.78], '*'),
    ('shape_weights', float, [1e2, 5e1, 1e1, 0.5e1], '*'),
    ('expr_weights', float, [1e2, 5e1, 1e1, 0.5e1], '*'),
    ('face_joints_weights', float, [0.0, 0.0, 0.0, 2.0], '*'),
    ('hand_joints_weights', float, [0.0, 0.0, 0.0, 2.0], '*'),
    ('jaw_pose_prior_weights', str, None, '*'),
    ('hand_pose_prior_weights', float, [1e2, 5e1, 1e1, 0.5e1], '*'),
    ('coll_loss_weights', float, [0.0, 0.0, 0.0, 2.0], '*'),
    ('depth_loss_weight', float, 1e2, None), ('df_cone_height', float, 0.5, None),
    ('max_collisions', int, 8, None), ('point2plane', _b, False, None),
    ('part_segm_fn', str, '', None), ('ign_part_pairs', str, None, '*'),
    ('use_hands', _b, False, None), ('use_face', _b, False, None),
    ('use_face_contour', _b, False, None), ('side_view_thsh', float, 25, None),
    ('optim_type', str, 'adam', None), ('lr', float, 1e-6, None),
    ('gtol', float, 1e-8, None), ('ftol', float, 2e-9, None),
    ('maxiters', int, 100, None), ('num_betas', int, 10, None),
    ('num_expression_coeffs', int, 10, None), ('regression_prior', str, None, None),
    ('pixie_results_directory', str, None, None),
    ('expose_results_directory', str, None, None),
    ('pare_results_directory', str, None, None),
    ('homogeneous_ckpt', str, './homogeneous/trained_models/tf/', None),
    ('use_camera_prior', _b, False, None), ('use_conf_for_camera_init', _b, False, None),
    ('use_gender_classifier', _b, False, None), ('save_vertices', _b, False, None),
    ('confidence_threshold', float, 0, None),
    # engine option (not in the reference): how the L-BFGS direction is computed, _native.TWO_LOOP_MODES
    ('two_loop', str, None, None),
    # engine options (not in the reference): batches kept in flight by main.py (default 2), frames
    # that fit two orientations on clusters of blocks ('auto' / 'off', fit_frames.upload)
    ('batches_in_flight', int, None, None), ('wide_frames', str, None, None),
]


def _coerce(typ, nargs, value):
    if value is None:
        return None
    if nargs == '*':
        if not isinstance(value, (list, tuple)):
            value = [value]
        return [typ(v) for v in value]
    return typ(value)


def parse_config(argv=None):
    """Returns the flat ``dict`` the reference's ``parse_config`` returns."""
    ap = argparse.ArgumentParser(prog='SMPLifyX', description='B200-native SMPLify-X fitting')
    ap.add_argument('-c', '--config', required=True, help='config file path')
    for name, typ, _default, nargs in _FLAGS:
        kw = dict(default=argparse.SUPPRESS, type=typ)
        if nargs:
            kw['nargs'] = nargs
        ap.add_argument('--' + name, **kw)
    cli = vars(ap.parse_args(argv))
    with open(cli['config']) as f:
        cfg = yaml.safe_load(f) or {}
    known = {n for n, _, _, _ in _FLAGS}
    unknown = set(cfg) - known
    if unknown:
        raise SystemExit('unrecognized config keys: ' + ', '.join(sorted(unknown)))
    out = {'config': cli['config']}
    for name, typ, default, nargs in _FLAGS:
        if name in cli:
            out[name] = cli[name]
        elif name in cfg:
            out[name] = _coerce(typ, nargs, cfg[name])
        else:
            out[name] = os.getcwd() if name == 'data_folder' else default
    tri = out['body_tri_idxs']
    if len(tri) % 2 != 0:
        raise AssertionError('Number of body_tri_idxs arguments must be divisble by 2.'
                             ' Got: {}'.format(len(tri)))
    out['body_tri_idxs'] = [(tri[i], tri[i + 1]) for i in range(0, len(tri), 2)]
    return out
