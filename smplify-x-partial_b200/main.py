"""Command line of the reference (smplifyx/main.py:51-323) on the batched CUDA engine:

    python -m smplifyx_b200.main -c cfg_files/fit_smplx_combined_coco25.yaml \\
        --data_folder DATA --output_folder OUT [--model_folder MODELS] [--batch_size 128]

Same YAML keys / flags, same outputs (``<output>/results/<fn>/000.pkl``, ``vertices.ply``,
``conf.yaml``).  Unlike the reference, which fits image after image, the frames of the data
folder are packed ``batch_size`` at a time and each batch is fitted in one persistent CUDA
launch (``fit_frames.submit`` / ``finish``, ``batches_in_flight`` of them overlapping on their own
CUDA streams); with ``torchrun`` the image list is sharded over the ranks
(each rank reads only its own frames), and the fitted parameter rows are all-gathered once at the
end (``sharding.gather_frames``): rank 0 writes ``<output>/fitted_params.npz`` (frame names,
[n, np] parameter table, layout offsets) next to the per-frame pickles every rank writes.
Limits kept from round 1: model_type 'smplx' and one gender per run (the reference builds three
gendered models, main.py:109-127; its gender classifier is out of scope).
"""
import os
import pickle
import sys
import time

import numpy as np
import torch
import yaml

from . import body_model as BM
from . import fit_frames as FF
from . import sharding
from . import _native as N
from . import utils as U
from .cmd_parser import parse_config
from .data_parser import create_dataset
from .fit_single_frame import write_ply_vertices


def _load_regression(args, img_name):
    import joblib
    pixie = expose = pare = None
    if args.get('regression_prior'):
        d = args.get('pixie_results_directory')
        if d:
            pixie = joblib.load(os.path.join(d, img_name, img_name + '_param.pkl'))
        d = args.get('expose_results_directory')
        if d:
            expose = dict(np.load(os.path.join(d, img_name + '.jpg', img_name + '.jpg_params.npz'),
                                  allow_pickle=True))
        d = args.get('pare_results_directory')
        if d:
            pare = joblib.load(os.path.join(d, img_name + '.pkl'))          # main.py:292-293
    return pixie, expose, pare


def main(**args):
    output_folder = os.path.expandvars(args.pop('output_folder'))
    os.makedirs(output_folder, exist_ok=True)
    with open(os.path.join(output_folder, 'conf.yaml'), 'w') as f:
        yaml.dump(args, f)
    result_folder = os.path.join(output_folder, args.pop('result_folder', 'results'))
    os.makedirs(result_folder, exist_ok=True)
    if args.get('use_gender_classifier'):
        raise NotImplementedError('gender classifier (TF1 Homogenus) is out of scope; pass '
                                  '--use_gender_classifier False --gender <g>')
    float_dtype = args.get('float_dtype', 'float32')
    if float_dtype not in ('float32', 'float64'):
        raise ValueError('Unknown float type {}, exiting!'.format(float_dtype))
    dtype = torch.float64 if float_dtype == 'float64' else torch.float32
    rank, world = sharding.init_from_env()
    start = time.time()
    dataset = create_dataset(img_folder=args.pop('img_folder', 'images'),
                             keyp_folder=args.pop('keyp_folder', 'keypoints'),
                             data_folder=args.pop('data_folder'), dtype=dtype, **args)
    joint_map = dataset.get_model2data()
    gender = args.get('gender', 'neutral')
    md = BM.load_model(args.get('model_folder'), gender)
    from . import engine
    model = engine.Model(md, joint_map, dtype=dtype, num_betas=args.get('num_betas', 10),
                         num_expression_coeffs=args.get('num_expression_coeffs', 10),
                         use_pca=args.get('use_pca', True),
                         num_pca_comps=args.get('num_pca_comps', 6),
                         flat_hand_mean=args.get('flat_hand_mean', False),
                         use_face_contour=args.get('use_face_contour', False))
    body_pose_prior = None
    if args.get('body_prior_type') == 'gmm':
        from .prior import create_prior
        body_pose_prior = create_prior(prior_type='gmm', dtype=dtype, **args)
    vposer = None
    if args.get('use_vposer'):
        from .vposer import load_vposer
        vposer, _ = load_vposer(os.path.expandvars(args.get('vposer_ckpt')), dtype=dtype)
        model.set_vposer(vposer.weights)
    # The image paths are sharded first and read batch by batch: keypoints + image size only
    # (the reference streams one decoded image at a time, main.py:207; nothing but H and W of the
    # pixels is used by the fit).
    paths = list(dataset.img_paths)
    mine = sharding.shard_range(len(paths), rank, world)
    paths = paths[mine.start:mine.stop]
    bs = max(1, int(args.get('batch_size', 1)))
    args.setdefault('device_ingest', True)       # masks / thresholds on the device (sfx_keypoint_masks)
    names, rows = [], []
    L = None
    # batches in flight (key batches_in_flight, default 2): batch i + 1 is read, planned and queued
    # on its own FrameBatch / CUDA stream while batch i is being fitted, so the straggler frames
    # of one batch overlap the next batch's frames and the host work hides behind the device
    in_flight = max(1, int(args.get('batches_in_flight') or 2))
    use_vposer = bool(args.get('use_vposer'))
    slots = [None] * in_flight                      # FrameBatch per slot (re-made when B changes)
    streams = [torch.cuda.Stream(device=model.device) for _ in range(in_flight)]
    pending = []

    def write_results(chunk, pf):
        out = FF.finish(pf)
        for b, d in enumerate(chunk):
            print('Processing: {}'.format(d['img_path']))
            fl = int(out.flags[b])
            if fl & N.SFX_FLAG_NAN:
                print('NaN loss value, stopping!')                 # fitting.py:177-179
            if fl & N.SFX_FLAG_INF:
                print('Infinite loss value, stopping!')            # fitting.py:181-183
            if fl & N.SFX_FLAG_COLL_OVERFLOW:
                print('Warning: interpenetration candidate list truncated for {}'.format(d['fn']))
            folder = os.path.join(result_folder, d['fn'])
            os.makedirs(folder, exist_ok=True)
            with open(os.path.join(folder, '000.pkl'), 'wb') as f:
                pickle.dump(out.results[b], f, protocol=2)
            if args.get('save_vertices'):
                write_ply_vertices(os.path.join(folder, 'vertices.ply'), out.vertices[b])
            names.append(d['fn'])
            rows.append(out.params[b])

    step = 0
    for lo in range(0, len(paths), bs):
        chunk = [d for d in (dataset.read_meta(p) for p in paths[lo:lo + bs]) if d]
        if not chunk:
            continue
        B = len(chunk)
        kp = np.stack([d['keypoints'][0] for d in chunk])          # person 0 only (main.py:242-246)
        H = [d['H'] for d in chunk]
        W = [d['W'] for d in chunk]
        if args.get('focal_length') is None:
            # the reference stores the first image's focal length back into its arguments
            # (main.py:212-218): every later image of the folder re-uses it
            args['focal_length'] = float((W[0] ** 2 + H[0] ** 2) ** 0.5)
        reg = [_load_regression(args, d['fn']) for d in chunk]
        k = step % in_flight
        step += 1
        if len(pending) == in_flight:                              # slot k's previous fit
            write_results(*pending.pop(0))
        if slots[k] is None or slots[k].B != B:
            if slots[k] is not None:
                slots[k].close()
            slots[k] = engine.FrameBatch(model, B, use_vposer=use_vposer)
        L = slots[k].L
        with torch.cuda.stream(streams[k]):
            pf = FF.submit(slots[k], kp, H, W, args, expose=[r[1] for r in reg],
                           pixie=[r[0] for r in reg], pare=[r[2] for r in reg],
                           return_verts=bool(args.get('save_vertices')),
                           body_pose_prior=body_pose_prior, vposer=vposer)
        pending.append((chunk, pf))
    while pending:
        write_results(*pending.pop(0))
    for sl in slots:
        if sl is not None:
            sl.close()
    # one all-gather of the fitted parameter rows: rank 0 writes the whole job's table
    if world > 1:
        dev = model.device
        local = torch.as_tensor(np.stack(rows) if rows else np.zeros((0, L.np if L else 0)),
                                dtype=dtype, device=dev)
        allp = sharding.gather_frames(local)
        all_names = [None] * world
        torch.distributed.all_gather_object(all_names, names)
        names = [n for part in all_names for n in part]
        table = allp.cpu().numpy()
    else:
        table = np.stack(rows) if rows else np.zeros((0, 0))
    if rank == 0 and len(names):
        np.savez(os.path.join(output_folder, 'fitted_params.npz'), names=np.array(names),
                 params=table, **({} if L is None else {
                     'layout_' + k: np.array(getattr(L, k)) for k, _ in L._fields_}))
    sharding.finalize()
    if rank == 0:
        print('Processing the data took: {}'.format(
            time.strftime('%H hours, %M minutes, %S seconds', time.gmtime(time.time() - start))))


if __name__ == '__main__':
    main(**parse_config())
