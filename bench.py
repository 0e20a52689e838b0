#!/usr/bin/env python
"""Benchmark of the hot path: full multi-stage SMPLify-X fit of a batch of frames.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames B] [--impl b200|reference]

A "step" is one complete fit (camera stage + every annealing stage + final full-mesh forward)
of one batch of B synthetic frames.  Default workload = BASELINE.json configs[1] as SURVEY.md
section 8d specifies it: 128 frames, neutral SMPL-X-shaped model, GMoF data term + L2 priors,
3-stage weight schedule of cfg_files/fit_smplx_combined_coco25.yaml, lbfgsls, NO regression
prior: the fit starts from the prior's mean pose (fit_single_frame.py:250-252) and the camera from
guess_init (:404-411); no interpenetration.  ``--regression-prior`` adds the synthetic "combined"
regression + camera prior (the per-frame setup of config 5; round 1's default line),
``--vposer`` is config 3, ``--interpenetration`` config 4.

* ``value``   frames/s with the inputs already resident in HBM (device-timed, CUDA events);
* ``e2e``     frames/s through the public call ``fit_frames`` with host buffers: planning,
              host->device copies, every launch, device->host read of the fitted parameters
              and meshes, all inside the timed region;
* ``roofline`` for the dominant kernel ``fit_pipeline_kernel`` (DESIGN.md "Measurement");
* ``cpu_baseline`` the oracle port of the reference (oracle/fit_port.py) on a bounded sample
              of the same frames on this host's cores (rank 0, N = 1 only);
* ``parity``  the engine's fitted frames against the CPU leg's fits of the SAME frames (final
              loss, vertices, reprojected keypoints, evaluation counts);
* ``value_exact`` the same device-timed figure with the reference's own two-loop operation
              order (``value`` runs the Gram formulation of the same recursion).

``--impl reference`` times the reference's CPU path (the oracle port; a step fits one frame per
host core, one single-threaded process per core -- the reference itself is batch-size-1 and
single-process) and prints the same JSON line with ``"impl": "reference"``.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'frames/sec (full multi-stage fit, 127 body kpts)'
UNIT = 'frames/s'
SUPPORT_ROWS = 225 * 3
ROW_BYTES = 512 * 4
# dram__bytes_read.sum + dram__bytes_write.sum of the fit_pipeline_kernel<float> launches of one
# step (128 frames), from `ncu --set full` captures summarised under profiles/ (the per-frame Gram
# blocks of the Gram two-loop add 80 KB per frame; `_reg` = the --regression-prior workload)
NCU_TRAFFIC_BYTES = {'gram_reg': (24584448.0, 'profiles/r01v_ncu_raw_pipeline_kernel.csv'),
                     'exact_reg': (12591616.0, 'profiles/r01p_ncu_raw_pipeline_kernel.csv'),
                     # default workload, round 2: one-block grid 11.02 + 13.34 MB, wide grid 2.24 + 0.06 MB
                     # default workload, round 2, one-block grid (what runs with steps in flight): 11.46 + 12.04 MB
                     'gram': (23501568.0, 'profiles/r02u_ncu_raw_pipeline_kernel.csv')}


# ------------------------------------------------------------------------------ workload
COLL_POSE_CORRECTIVE_SCALE = 0.1      # config 4: see synthetic.cached_smplx_like
# How the engine computes the L-BFGS direction (_native.TWO_LOOP_MODES): 'gram' runs the
# reference's two-loop recursion on the inner products of the history (same mathematics,
# different rounding, ~2.4x shorter); 'exact' reproduces the reference's operation order.
TWO_LOOP = os.environ.get('SFX_TWO_LOOP', 'gram')
# frames on clusters of blocks (fit_frames cfg key wide_frames) only while one step runs at a time:
# with several steps in flight every SM is busy and a cluster's helpers would take SMs from frames
WIDE_IN_FLIGHT = 'off'
SM_CLOCK_KHZ = 1965.0e3      # clock64 ticks per millisecond at the 1965 MHz SM clock of the B200 (see clocks.sm_mhz)


def bench_cfg(interpenetration=False, vposer=False, regression_prior=False, two_loop=None):
    """fit_smplx_combined_coco25.yaml (reference cfg_files/) with BASELINE config-2 switches.

    Default: no regression prior -- pose from the prior's mean
    (fit_single_frame.py:250-252; the zero pose for the L2 prior, see oracle/fit_port.py), camera
    depth from guess_init (:404-411).  ``regression_prior`` restores the yaml's "combined"
    regression + camera prior (config 5's per-frame setup); ``interpenetration`` turns the
    yaml's interpenetration settings back on (config 4, with the regression prior as the yaml
    has it); ``vposer`` selects the 5-stage fit_smplx_smplifyx.yaml schedule with the VPoser
    latent pose, zero-latent start, guess_init camera and focal length 5000 (config 3)."""
    cfg = _bench_cfg()
    cfg['two_loop'] = two_loop or TWO_LOOP
    if not (regression_prior or interpenetration):
        cfg.update(regression_prior=None, use_camera_prior=False)
    if vposer:
        cfg.update(
            use_vposer=True, regression_prior=None, use_camera_prior=False,
            use_conf_for_camera_init=False, init_joints_idxs=[9, 12, 2, 5], focal_length=5000.0,
            body_pose_prior_weights=[404.0, 404.0, 57.4, 4.78, 4.78],
            coll_loss_weights=[0.0, 0.0, 0.0, 0.0, 0.0],
            shape_weights=[100.0, 50.0, 10.0, 5.0, 5.0], expr_weights=[100.0, 50.0, 10.0, 5.0, 5.0],
            hand_pose_prior_weights=[404.0, 404.0, 57.4, 4.78, 4.78],
            jaw_pose_prior_weights=['4.04e03,4.04e04,4.04e04', '4.04e03,4.04e04,4.04e04',
                                    '574,5740,5740', '47.8,478,478', '47.8,478,478'],
            hand_joints_weights=[0.0, 0.0, 0.0, 0.1, 2.0],
            face_joints_weights=[0.0, 0.0, 0.0, 0.0, 2.0])
    if interpenetration:
        from smplifyx_b200 import synthetic
        md = synthetic.cached_smplx_like(0, COLL_POSE_CORRECTIVE_SCALE)
        cfg.update(interpenetration=True, coll_loss_weights=[0.0, 0.1, 1.0], df_cone_height=1e-4,
                   max_collisions=128, penalize_outside=True, point2plane=False,
                   ign_part_pairs=["9,16", "9,17", "6,16", "6,17", "1,2", "12,22"] +
                   synthetic.sibling_part_pairs(md) + synthetic.REST_POSE_TOUCHING_PART_PAIRS)
    return cfg


def _bench_cfg():
    return dict(
        format='coco25', joints_to_ign=[1, 9, 12], gender='neutral', model_type='smplx',
        float_dtype='float32', use_joints_conf=True, use_pca=True, use_hands=True, use_face=True,
        flat_hand_mean=False, body_prior_type='l2', left_hand_prior_type='l2',
        right_hand_prior_type='l2', jaw_prior_type='l2', num_pca_comps=12, rho=100,
        interpenetration=False, optim_type='lbfgsls', ftol=1e-9, gtol=1e-9, lr=1.0, maxiters=30,
        body_pose_prior_weights=[500, 300, 200], coll_loss_weights=[0.0, 0.0, 0.0],
        shape_weights=[75, 50, 35], expr_weights=[10.0, 5.0, 5.0],
        hand_pose_prior_weights=[57.4, 4.78, 4.78],
        jaw_pose_prior_weights=['1000, 10000, 10000', '100, 1000, 1000', '100, 1000, 1000'],
        hand_joints_weights=[0.0, 0.1, 2.0], face_joints_weights=[0.0, 0.0, 2.0],
        use_face_contour=True, init_joints_idxs=[0, 1, 2, 3, 5, 6, 8, 9, 12, 15, 16, 17, 18],
        body_tri_idxs=[(5, 12), (2, 9)], use_vposer=False, num_betas=10,
        num_expression_coeffs=10, regression_prior='combined', use_camera_prior=True,
        use_conf_for_camera_init=True, confidence_threshold=0.2, depth_loss_weight=1e2,
        side_view_thsh=25.0, focal_length=None, two_loop=TWO_LOOP)


MODEL_KW = dict(num_betas=10, num_expression_coeffs=10, use_pca=True, num_pca_comps=12,
                flat_hand_mean=False, use_face_contour=True)
H_IMG, W_IMG = 600, 800


def _rodrigues_batch(r):
    th = np.linalg.norm(r, axis=-1, keepdims=True)
    k = r / np.maximum(th, 1e-12)
    K = np.zeros(r.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s, c = np.sin(th)[..., None], np.cos(th)[..., None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def ground_truth(B, seed, shape_scale=1.0):
    """Seeded ground-truth parameters of B frames.

    ``shape_scale`` < 1 (config 4): the synthetic model's shape directions are per-vertex noise,
    so at unit betas the tube mesh self-intersects in ~10^4 triangle pairs wherever two tubes
    meet -- two orders of magnitude beyond a real body.  Scaling the ground-truth shape keeps the
    interpenetration term's load at what the pose causes (arms against the torso)."""
    from smplifyx_b200 import utils as U
    rng = np.random.default_rng(seed)
    gt = dict(
        body_pose=rng.normal(size=(B, 63)) * 0.2, betas=rng.normal(size=(B, 10)),
        expression=rng.normal(size=(B, 10)), left_hand_pose=rng.normal(size=(B, 12)) * 0.5,
        right_hand_pose=rng.normal(size=(B, 12)) * 0.5, jaw_pose=rng.normal(size=(B, 3)) * 0.05,
        leye_pose=np.zeros((B, 3)), reye_pose=np.zeros((B, 3)))
    # upright in the image: the model is y-up, the camera y-down -> rotate by pi about x
    go = np.zeros((B, 3))
    Rx = U.rodrigues([np.pi, 0, 0])
    for b in range(B):
        go[b] = U.inv_rodrigues(U.rodrigues(rng.normal(size=3) * 0.3).dot(Rx))
    gt['global_orient'] = go
    gt['betas'] *= shape_scale
    gt['expression'] *= shape_scale
    t = np.stack([rng.normal(size=B) * 0.1, rng.normal(size=B) * 0.1,
                  rng.uniform(2.5, 6.0, size=B)], axis=1)
    gt['transl'] = t
    return gt, rng


def observations(gt, joints3d, rng, focal_length=None):
    """Noisy 2-D keypoints + synthetic regression results from GT joints [B,K,3].  With an
    explicit ``focal_length`` (config 3: 5000) the ground-truth depth is scaled along so the
    person keeps its size in the 800 x 600 image."""
    from smplifyx_b200 import utils as U
    B, K, _ = joints3d.shape
    focal = float(np.sqrt(H_IMG ** 2 + W_IMG ** 2))
    if focal_length is not None:
        gt['transl'][:, 2] *= float(focal_length) / focal
        focal = float(focal_length)
    c = np.array([W_IMG * 0.5, H_IMG * 0.5])
    p = joints3d + gt['transl'][:, None]
    uv = focal * p[:, :, :2] / p[:, :, 2:3] + c
    uv = uv + rng.normal(size=uv.shape) * 2.0
    conf = rng.uniform(0.3, 1.0, size=(B, K))
    conf[rng.uniform(size=(B, K)) < 0.15] = 0.0
    kp = np.concatenate([uv, conf[:, :, None]], axis=2).astype(np.float32)
    expose, pixie = [], []
    for b in range(B):
        def noisy(aa):
            aa = np.asarray(aa).reshape(-1, 3)
            return _rodrigues_batch(aa + rng.normal(size=aa.shape) * 0.1).astype(np.float32)
        tr = gt['transl'][b].copy()
        tr[:2] += rng.normal(size=2) * 0.05
        tr[2] = tr[2] * (1 + 0.05 * rng.normal()) * (5000.0 / focal)
        expose.append(dict(body_pose=noisy(gt['body_pose'][b]),
                           global_orient=noisy(gt['global_orient'][b]), transl=tr,
                           center=(c + rng.normal(size=2) * 2.0).astype(np.float32)))
        pixie.append(dict(body_pose=noisy(gt['body_pose'][b]),
                          global_pose=noisy(gt['global_orient'][b])))
    return kp, expose, pixie


def gt_param_matrix(L, gt):
    from smplifyx_b200 import _native as N
    B = gt['betas'].shape[0]
    x = np.zeros((B, L.np))
    for name, (off, n) in N.param_blocks(L).items():
        key = {'pose_embedding': 'body_pose'}.get(name, name)
        if key in gt:
            x[:, off:off + n] = gt[key]
    return x


# ------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(mx)) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------ reference arm
def oracle_objects(cfg):
    """-> (restated smplx body model, base joint weights, restated VPoser or None)."""
    import torch
    from oracle import smplx_shim
    from smplifyx_b200 import synthetic, utils as U
    jm = U.smpl_to_annotation('smplx', use_hands=True, use_face=True, use_face_contour=True,
                              format='coco25')
    scale = COLL_POSE_CORRECTIVE_SCALE if cfg.get('interpenetration') else 1.0
    bm = smplx_shim.create(model_data=synthetic.cached_smplx_like(0, scale),
                           joint_mapper=U.JointMapper(jm.astype(np.int64)), dtype=torch.float32,
                           create_body_pose=not cfg.get('use_vposer', False), **MODEL_KW)
    jw = np.ones(len(jm))
    jw[cfg['joints_to_ign']] = 0
    vp = None
    if cfg.get('use_vposer', False):
        from oracle import vposer_shim
        vp = vposer_shim.from_weights(synthetic.make_vposer_like(seed=2), dtype=torch.float32)
    return bm, jw, vp


def oracle_joints(bm, gt):
    import torch
    with torch.no_grad():
        out = bm(**{k: torch.tensor(v, dtype=torch.float32) for k, v in gt.items()
                    if k != 'transl'}, return_verts=False)
    return out.joints.numpy().astype(np.float64)


def _fit_one(W, b, want_result):
    import torch
    from oracle import fit_port as FP
    ex = None if W['expose'] is None else W['expose'][b]
    px = None if W['pixie'] is None else W['pixie'][b]
    t0 = time.perf_counter()
    r = FP.fit_frame(W['bm'], W['kp'][b], H_IMG, W_IMG, W['cfg'], W['jw'], expose=ex, pixie=px,
                     vposer=W['vp'], dtype=torch.float32, return_verts=True)
    dt = time.perf_counter() - t0
    out = {'secs': dt, 'evals': int(r['n_evals'])}
    if want_result:
        out.update(loss=float(r['loss']), vertices=np.asarray(r['vertices'], np.float32)[0],
                   joints=np.asarray(r['joints'], np.float32)[0],
                   cam_t=np.asarray(r['result']['camera_translation'], np.float32).reshape(3),
                   center=np.asarray(r['result']['camera_center'], np.float32).reshape(2),
                   n_orient=int(r['n_orient']))
    return out


def time_oracle_frames(cfg, kp, expose, pixie, frames, threads, want_result=False):
    """Fits ``frames`` sequentially with the oracle port; returns one dict per frame."""
    import torch
    import warnings
    torch.set_num_threads(threads)
    bm, jw, vp = oracle_objects(cfg)
    W = dict(cfg=cfg, kp=kp, expose=expose, pixie=pixie, bm=bm, jw=jw, vp=vp)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return [_fit_one(W, b, want_result) for b in frames]


# The reference fits one frame at a time (fit_single_frame.py:119 asserts batch_size == 1) and its
# tensors are far too small for intra-op threads to help (1 -> 4 threads: 1.8x), so the way to use
# every host core is the one a user of the reference has: one process per core, each fitting its
# own frames with one thread.  That is what the CPU arm times.
_W = {}


def _oracle_worker_init(cfg, kp, expose, pixie):
    import warnings
    import torch
    torch.set_num_threads(1)
    warnings.simplefilter('ignore')
    bm, jw, vp = oracle_objects(cfg)
    _W.update(cfg=cfg, kp=kp, expose=expose, pixie=pixie, bm=bm, jw=jw, vp=vp)


def _oracle_worker_ready(_):
    time.sleep(0.2)         # keeps the task on this worker long enough for every worker to get one
    return os.getpid()


def _oracle_worker_fit(task):
    b, want_result = task
    return _fit_one(_W, b, want_result)


class OraclePool(object):
    """``procs`` single-threaded worker processes (spawned: the parent may hold a CUDA context),
    each with its own copy of the oracle's model."""

    def __init__(self, cfg, kp, expose, pixie, procs):
        import multiprocessing
        ctx = multiprocessing.get_context('spawn')
        self.procs = procs
        self.pool = ctx.Pool(procs, initializer=_oracle_worker_init,
                             initargs=(cfg, kp, expose, pixie))
        self.pool.map(_oracle_worker_ready, range(4 * procs), chunksize=1)

    def round(self, frames, want_result=False):
        """Fits ``frames`` (one task each) -> (wall seconds, [dict per frame])."""
        t0 = time.perf_counter()
        res = self.pool.map(_oracle_worker_fit, [(b, want_result) for b in frames], chunksize=1)
        return time.perf_counter() - t0, res

    def close(self):
        self.pool.close()
        self.pool.join()


def oracle_coll_eval_seconds(cfg, kp, expose, pixie, n_evals=3):
    """config 4's CPU leg: the interpenetration term's restated search is seconds per evaluation
    in numpy, a full fit hours; so a bounded sample times single closure evaluations (forward,
    loss with the term on, backward) of frame 0 and the caller scales by the evaluation counts.
    -> (seconds per evaluation with the term on, seconds per evaluation with it off)."""
    import torch
    import warnings
    from oracle import fit_port as FP
    from smplifyx_b200 import synthetic
    torch.set_num_threads(1)
    bm, jw, vp = oracle_objects(cfg)
    md = synthetic.cached_smplx_like(0, COLL_POSE_CORRECTIVE_SCALE)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return FP.time_coll_closure(bm, kp[0], H_IMG, W_IMG, cfg, jw, expose[0], pixie[0],
                                    synthetic.parts_segm_like(md), n_evals=n_evals)


def make_inputs_cpu(args, cfg):
    """Reference arm: keypoints from the oracle's own forward pass."""
    B = args.frames
    gt, rng = ground_truth(B, args.seed, 0.2 if cfg.get('interpenetration') else 1.0)
    bm, _, _ = oracle_objects(cfg)
    kp, expose, pixie = observations(gt, oracle_joints(bm, gt), rng, cfg.get('focal_length'))
    if not cfg.get('regression_prior'):
        expose = pixie = None
    return kp, expose, pixie


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = bench_cfg(args.interpenetration, args.vposer, args.regression_prior)
    B = args.frames
    kp, expose, pixie = make_inputs_cpu(args, cfg)
    procs = os.cpu_count() or 1
    walls, evals = [], []
    if args.interpenetration:
        # bounded sample: single evaluations, scaled by the reference's own evaluation count of
        # the same schedule without the term (the engine's count is not available to this arm)
        t_on, t_off = oracle_coll_eval_seconds(cfg, kp, expose, pixie, n_evals=max(1, args.steps))
        cfg0 = bench_cfg(False, False, True)
        r = time_oracle_frames(cfg0, kp, expose, pixie, [0], 1)[0]
        per_frame = r['evals'] * (t_off + 2.0 / 3.0 * (t_on - t_off))     # term on in 2 of 3 stages
        value = procs / per_frame
        total, nsteps = per_frame, 1
        per_step, how = procs, ('extrapolated: {:.2f} s per evaluation with the interpenetration '
                                'term on, {:.3f} s without, x {} evaluations of frame 0 (term on in 2 '
                                'of 3 stages), one single-threaded process per host core assumed'
                                .format(t_on, t_off, r['evals']))
        evals = [r['evals']]
        ms_per_step = 1e3 * per_frame
    else:
        try:
            pool = OraclePool(cfg, kp, expose, pixie, procs)
            for step in range(args.warmup + args.steps):
                frames = [(step * procs + i) % B for i in range(procs)]
                wall, res = pool.round(frames)
                if step >= args.warmup:
                    walls.append(wall)
                    evals += [r['evals'] for r in res]
            pool.close()
            per_step, how = procs, 'one single-threaded process per host core, each fitting one frame'
        except Exception as exc:         # no worker processes on this host: one process, all threads
            frames = [i % B for i in range(args.warmup + args.steps)]
            res = time_oracle_frames(cfg, kp, expose, pixie, frames, procs)
            walls = [r['secs'] for r in res][args.warmup:]
            evals = [r['evals'] for r in res][args.warmup:]
            per_step, how = 1, 'fitted sequentially with {} torch threads (worker processes ' \
                               'unavailable: {})'.format(procs, type(exc).__name__)
        total = float(np.sum(walls))
        value = per_step * len(walls) / total
        ms_per_step = 1e3 * total / len(walls)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': workload_config(B, args.gpus, args.interpenetration, args.vposer,
                                  args.regression_prior, steps_in_flight(args), distinct_batches(args)),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': 'port',
                         'sample': '{} frame(s) of the batch per step, {} (the reference asserts '
                                   'batch_size == 1); mean evals/frame {:.0f}'.format(
                                       per_step, how, np.mean(evals))},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def steps_in_flight(args):
    # (config 4: 0.6 GB of collision workspace per 128-frame batch in flight)
    return max(1, int(getattr(args, 'depth', 6)))


def distinct_batches(args):
    return (1 if args.interpenetration else max(1, int(getattr(args, 'batches', 2)))) * steps_in_flight(args)


def workload_config(B, n_gpus, interpenetration=False, vposer=False, regression_prior=False, depth=6, nb=12):
    common = {'frames_per_gpu': B, 'global_frames': B * n_gpus,
              'batches': '{} distinct synthetic batches per rank (seed + 1000 rank + 7919 j), cycled step '
                         'after step; the reference arm, cpu_baseline and the parity block use batch 0'.format(nb),
              'steps_in_flight': '{} (engine arm: consecutive steps alternate between that many '
                                 'FrameBatch objects on their own CUDA streams, so the straggler '
                                 'frames of one step overlap the next steps\' frames; every step is a '
                                 'complete fit of its own batch, one block per frame; '
                                 'value_one_step_at_a_time is the same measurement with one step in '
                                 'flight, where the frames that fit two orientations run on clusters of '
                                 '8 blocks; the reference arm fits one frame per process)'.format(depth),
              'parallelism': 'frames sharded, dp{}'.format(n_gpus),
              'l2': 'flushed between timed steps (256 MiB write)',
              'two_loop': TWO_LOOP + ' (engine option for the L-BFGS direction, value_exact is '
                          'the other one; the reference arm always runs the reference recursion)'}
    if vposer:
        w = ('batch={} synthetic frames per GPU, 135 keypoints, neutral SMPL-X-shaped synthetic '
             'model, VPoser latent pose prior (synthetic VPoser v1 weights), 5-stage '
             'fit_smplx_smplifyx schedule, lbfgsls, guess_init camera, focal 5000, interpenetration '
             'off (BASELINE config 3)'.format(B))
    else:
        init = ('combined regression + camera prior (config 5 per-frame setup)'
                if (regression_prior or interpenetration) else
                'no regression prior: zero (prior-mean) start pose, guess_init camera (BASELINE '
                'config 2 as SURVEY 8d specifies it)')
        w = ('batch={} synthetic frames per GPU, 135 keypoints (127 model joints + 17 face-contour), '
             'neutral SMPL-X-shaped synthetic model, GMoF + L2 priors, 3-stage '
             'fit_smplx_combined_coco25 schedule, lbfgsls, {}, interpenetration {}'.format(
                 B, init, 'ON (coll_loss_weights 0 / 0.1 / 1.0, df_cone_height 1e-4, part filter; '
                 'BASELINE config 4)' if interpenetration else 'off'))
    common['workload'] = w
    return common


# ------------------------------------------------------------------------------ parity
def project_np(joints, cam_t, center, focal):
    p = joints + cam_t[None]
    return focal * p[:, :2] / p[:, 2:3] + center[None]


def parity_block(out, oracle_res, sample, kp, focal):
    """Engine fit (``out``: fit_frames result) against the CPU leg's fit of the same frames."""
    loss_rel, v_mean, v_max, px_mean, px_max, ev_e, ev_o = [], [], [], [], [], [], []
    for b, r in zip(sample, oracle_res):
        loss_rel.append(abs(float(out.loss[b]) - r['loss']) / max(abs(r['loss']), 1.0))
        d = np.sqrt(((out.vertices[b].astype(np.float64) - r['vertices']) ** 2).sum(-1))
        v_mean.append(float(d.mean()))
        v_max.append(float(d.max()))
        res = out.results[b]
        pe = project_np(out.joints[b].astype(np.float64),
                        res['camera_translation'].reshape(3).astype(np.float64),
                        res['camera_center'].reshape(2).astype(np.float64), focal)
        po = project_np(r['joints'].astype(np.float64), r['cam_t'].astype(np.float64),
                        r['center'].astype(np.float64), focal)
        live = kp[b, :, 2] > 0
        dpx = np.sqrt(((pe - po) ** 2).sum(-1))[live]
        px_mean.append(float(dpx.mean()))
        px_max.append(float(dpx.max()))
        ev_e.append(int(out.n_evals[b]))
        ev_o.append(int(r['evals']))
    return {'frames': len(sample),
            'final_loss_rel_median': float(np.median(loss_rel)),
            'final_loss_rel_max': float(np.max(loss_rel)),
            'vertex_err_mean_m': float(np.mean(v_mean)), 'vertex_err_max_m': float(np.max(v_max)),
            'vertex_err_mean_m_median_frame': float(np.median(v_mean)),
            'reproj_px_mean': float(np.mean(px_mean)), 'reproj_px_max': float(np.max(px_max)),
            'evals_engine': float(np.mean(ev_e)), 'evals_oracle': float(np.mean(ev_o)),
            'note': 'engine (two_loop {}) vs the oracle port of the reference on the same frames '
                    'of this batch; vertices in model space (last orientation, as vertices.ply), '
                    'reprojection over the detected keypoints; float32 fits are chaotic -- the '
                    'reference-vs-reference envelope on these frames is in '
                    'tests/golden/bench_parity.npz (tests/test_gpu_bench_parity.py)'.format(TWO_LOOP)}


# ------------------------------------------------------------------------------ B200 arm
# floating-point operations of one closure evaluation besides the blend rows (pose prologue,
# kinematic chain, skinning of <= 225 support vertices, projection, loss, their adjoints):
# counted from the loops of csrc/sfx_core.cuh eval_frame -- see DESIGN.md section 4
EVAL_FIXED_FLOPS = 2.3e5
FP32_SIMT_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # FMA lanes x 2 x boost clock


def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from smplifyx_b200 import engine, fit_frames as FF, synthetic, utils as U, _native as N

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cfg = bench_cfg(args.interpenetration, args.vposer, args.regression_prior)
    md = synthetic.cached_smplx_like(
        0, COLL_POSE_CORRECTIVE_SCALE if args.interpenetration else 1.0)
    part_segm = synthetic.parts_segm_like(md) if args.interpenetration else None
    B = args.frames
    jm = U.smpl_to_annotation('smplx', use_hands=True, use_face=True, use_face_contour=True,
                              format='coco25')
    model = engine.Model(md, jm, dtype=torch.float32, **MODEL_KW)
    vp = None
    if args.vposer:
        from smplifyx_b200 import vposer as V
        vp = V.VPoser(synthetic.make_vposer_like(seed=2))
        model.set_vposer(vp.weights)
    batch = engine.FrameBatch(model, B, use_vposer=args.vposer)
    L = batch.L

    # ---- synthetic inputs: GT parameters -> model joints (engine forward) -> noisy keypoints.
    # NB distinct batches (seeds seed + 1000 rank + 7919 j), cycled step after step: a stream of
    # different batches, as a service sees it; batch 0 is the one the CPU arm and the parity block use
    depth = steps_in_flight(args)
    NB = distinct_batches(args)
    K = model.K
    gbatch = engine.FrameBatch(model, B) if args.vposer else batch     # axis-angle pose block
    Lg = gbatch.L
    cam_st = N.make_stage(Lg, N.CAMERA_STAGE_BLOCKS, loss_kind=N.LOSS_CAMERA_INIT)
    zero_cam = np.zeros((B, N.SFX_CAM_STRIDE))
    zero_cam[:, 0:2] = 1.0
    zero_cam[:, 4:13] = np.eye(3).reshape(-1)
    inputs = []
    for j in range(NB):
        gt, rng = ground_truth(B, args.seed + 1000 * rank + 7919 * j, 0.2 if args.interpenetration else 1.0)
        xg = gt_param_matrix(Lg, gt)
        xg[:, Lg.off_camt + 2] = 1.0
        gbatch.set_targets(np.zeros((B, K, 3)), np.zeros((B, K)), np.zeros((B, K), np.uint8),
                           np.zeros((B, K), np.uint8), zero_cam, None)
        gbatch.set_params(xg)
        _, _, j3 = gbatch.eval(cam_st, want_joints=True)
        kp_j, expose_j, pixie_j = observations(gt, j3.cpu().numpy().astype(np.float64), rng,
                                               cfg.get('focal_length'))
        if not cfg.get('regression_prior'):
            expose_j = pixie_j = None
        inputs.append((kp_j, expose_j, pixie_j))
    if args.vposer:
        gbatch.close()
    kp, expose, pixie = inputs[0]
    focal = float(cfg.get('focal_length') or np.sqrt(H_IMG ** 2 + W_IMG ** 2))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gathered = torch.empty((world * B, L.np), dtype=torch.float32, device=dev) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- steps in flight -------------------------------------------------------------------------
    # A step is one complete fit of one batch.  Consecutive steps alternate between `depth`
    # independent FrameBatch objects, each on its own CUDA stream: the straggler frames of step i
    # (a few SMs busy) overlap the first frames of step i + 1, the way a service that is fed batch
    # after batch runs.  Every step still fits its own 128 frames from scratch and the timed region
    # ends only when all of them are complete.  depth 1 = one batch at a time (also reported).
    batches = [batch] + [engine.FrameBatch(model, B, use_vposer=args.vposer) for _ in range(NB - 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
    # the single NCCL all-gather of a step's fitted parameters reads a copy of the rows
    # and runs on NCCL's own stream: the next step of that batch slot does not wait for the other
    # ranks.  The copies rotate through NG buffers (a buffer is waited for when it comes round again)
    NG = 32
    gathered_k = [torch.empty((world * B, L.np), dtype=torch.float32, device=dev) if world > 1 else None
                  for _ in range(NG)]
    staged_k = [torch.empty((B, L.np), dtype=torch.float32, device=dev) if world > 1 else None
                for _ in range(NG)]
    gather_work = [None] * NG
    gather_count = [0]

    def gather_params(k):
        if world > 1:
            g = gather_count[0] % NG
            gather_count[0] += 1
            if gather_work[g] is not None:
                gather_work[g].wait()                 # the buffer's previous gather has completed
            staged_k[g].copy_(batches[k].params_tensor())
            gather_work[g] = dist.all_gather_into_tensor(gathered_k[g], staged_k[g], async_op=True)

    def gather_drain():
        for g in range(NG):
            if gather_work[g] is not None:
                gather_work[g].wait()
                gather_work[g] = None

    def timed_steps(step_fn, d, steps):
        """`steps` steps, step i on stream i % d; device time of the whole region (ms)."""
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for st_ in streams[:d]:
            st_.wait_event(start)
        n = 0
        for i in range(steps):
            k = i % NB                              # the batch (and its FrameBatch object) of this step
            with torch.cuda.stream(streams[k % d]):
                flush.fill_(i & 0xff)
                n += step_fn(k)
        for st_ in streams[:d]:
            torch.cuda.current_stream().wait_stream(st_)
        gather_drain()
        end.record()
        barrier()
        return start.elapsed_time(end), n

    # ---- device-timed steps, inputs resident: default two-loop mode first, then the other ----
    other = {'gram': 'exact', 'exact': 'gram'}[TWO_LOOP]
    modes = [TWO_LOOP] if (args.interpenetration or args.single_mode) else [TWO_LOOP, other]
    sampler = ClockSampler(local)
    timed, timed_serial, plans, by_depth = {}, {}, {}, {}
    launches = 0
    for mi, mode in enumerate(modes):
        # steps in flight keep every SM busy, so a frame gets one block there (wide_frames off: a
        # cluster's helper blocks would only take SMs from other frames); one step at a time runs
        # the frames that fit two orientations on clusters (wide_frames auto)
        def make_plans(wide):
            mcfg = dict(cfg, two_loop=mode, wide_frames=wide)
            pls, xs = [], []
            for bk, (kp_j, expose_j, pixie_j) in zip(batches, inputs):
                pl = FF.FitPlan(L, K, kp_j, H_IMG, W_IMG, mcfg, expose_j, pixie_j, None, np.float32,
                                part_segm=part_segm, vposer=vp)
                FF.upload(bk, pl)
                pls.append(pl)
                xs.append(bk.params_tensor().clone())
            return pls, xs

        def make_step(pls, xs):
            def resident_step(k):
                batches[k].params_tensor().copy_(xs[k])
                _, verts, joints, n = FF.run(batches[k], pls[k], return_verts=True)
                gather_params(k)
                return n
            return resident_step

        mplans, x0s = make_plans(WIDE_IN_FLIGHT if depth > 1 else 'auto')
        plans[mode] = mplans[0]                     # the launch configuration `value` is measured with
        step = make_step(mplans, x0s)
        timed_steps(step, depth, args.warmup)
        if rank == 0 and mi == 0:
            sampler.start()
        timed[mode], n = timed_steps(step, depth, args.steps)
        if mi == 0:
            launches += n
            # the same launches with fewer steps in flight (the curve behind `value`)
            for d_ in (2, 4):
                if d_ < depth:
                    by_depth[d_] = reduce_max(timed_steps(step, d_, args.steps)[0])
        if depth > 1:
            mplans, x0s = make_plans('auto')
            step = make_step(mplans, x0s)
            timed_steps(step, 1, 1)
            timed_serial[mode], _ = timed_steps(step, 1, args.steps)
        else:
            timed_serial[mode] = timed[mode]
    total_ms_rank = timed[TWO_LOOP]
    total_ms = reduce_max(total_ms_rank)
    total_ms_serial = reduce_max(timed_serial[TWO_LOOP])
    total_ms_other = reduce_max(timed[other]) if other in timed else None
    plan = plans[TWO_LOOP]
    FF.upload(batch, plan)
    x0_dev = batch.params_tensor().clone()

    # ---- end to end through the public call with host buffers --------------------------------
    # FF.submit plans the batch on the host, queues H2D copies, the fit and the D2H copies of the
    # results (pinned buffers) on the step's stream; FF.finish waits for that step and hands the
    # results over.  With `depth` steps in flight the host reads step i - depth + 1 while step i runs.
    def e2e_steps(d, steps):
        import collections
        ecfg = dict(cfg, wide_frames=WIDE_IN_FLIGHT if d > 1 else 'auto')
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pending = collections.deque()
        res = None
        barrier()
        start.record()
        for st_ in streams[:d]:
            st_.wait_event(start)
        for i in range(steps):
            k = i % NB
            if len(pending) == d:
                r_ = FF.finish(pending.popleft())
                res = r_ if (res is None or r_.batch_index == 0) else res
            with torch.cuda.stream(streams[k % d]):
                flush.fill_(i & 0xff)
                kp_j, expose_j, pixie_j = inputs[k]
                pf = FF.submit(batches[k], kp_j, H_IMG, W_IMG, ecfg, expose_j, pixie_j, return_verts=True,
                               part_segm=part_segm, vposer=vp)
                pf.batch_index = k
                gather_params(k)
            pending.append(pf)
        while pending:
            r_ = FF.finish(pending.popleft())
            res = r_ if (res is None or r_.batch_index == 0) else res
        for st_ in streams[:d]:
            torch.cuda.current_stream().wait_stream(st_)
        gather_drain()
        end.record()
        barrier()
        return start.elapsed_time(end), res

    # untimed: every one of the NB batches once, so that its pinned result buffers exist and the
    # host path is warm before the timed loop (a fresh box measured 2 600 - 3 500 frames/s end to
    # end on its first process with the short warm-up, 4 270 afterwards)
    e2e_steps(depth, max(args.warmup, NB))
    e2e_ms_rank, out = e2e_steps(depth, args.steps)
    e2e_serial_rank = e2e_steps(1, args.steps)[0] if depth > 1 else e2e_ms_rank
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = reduce_max(e2e_ms_rank)
    e2e_serial_ms = reduce_max(e2e_serial_rank)

    # ---- roofline accounting for the dominant kernel (one untimed replay with counters) ------
    batch.params_tensor().copy_(x0_dev)
    batch.reset_counters()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.fill_(1)
    a.record()
    batch.fit_pipeline(plan.pipeline, plan.order_dev, plan.flip_mask_dev)
    b_.record()
    torch.cuda.synchronize()
    kern_ms = a.elapsed_time(b_)
    frame_cycles = batch.frame_cycles().cpu().numpy().astype(np.float64)
    passes = batch.passes().cpu().numpy().astype(np.int64)
    evals = batch.evals().cpu().numpy().astype(np.int64)
    row_bytes = float(passes.sum()) * ROW_BYTES        # rows of the blend matrix streamed x 2 KiB
    flops = float(passes.sum()) * 2 * 512 + float(evals.sum()) * EVAL_FIXED_FLOPS
    evals_per_frame = float(evals.mean())
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    tf_peak = float(peaks.get('bf16_tflops_sustained', 1400.0))
    l2_gbs = C.c_double(0.0)
    N.check(batch.lib, batch.lib.sfx_diag_l2_read_gbs(48 << 20, 20, C.byref(l2_gbs)))
    traffic = args.traffic
    traffic_src = 'command line' if traffic is not None else None
    if traffic is None and B == 128 and not (args.interpenetration or args.vposer):
        key = TWO_LOOP + ('_reg' if args.regression_prior else '')
        if key in NCU_TRAFFIC_BYTES:
            traffic, traffic_src = NCU_TRAFFIC_BYTES[key]
    sec = kern_ms * 1e-3
    achieved_tf = flops / sec / 1e12
    roofline = {
        'bound': 'latency', 'kernel': 'fit_pipeline_kernel<float>',
        'achieved': achieved_tf, 'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': achieved_tf / tf_peak,
        'peak_source': ('MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else
                        'fallback (B200_PROFILING.md)'),
        'traffic': traffic, 'traffic_source': traffic_src,
        'launches': 1, 'kernel_ms_per_step': kern_ms, 'executed_flops_per_step': flops,
        'frac_of_fp32_simt_peak': achieved_tf / FP32_SIMT_PEAK_TFLOPS,
        'hbm': None if traffic is None else {
            'achieved': traffic / sec / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
            'frac': traffic / sec / 1e9 / hbm_peak},
        'l2_to_sm': {'achieved': row_bytes / sec / 1e9, 'peak': float(l2_gbs.value), 'unit': 'GB/s',
                     'frac': row_bytes / sec / 1e9 / max(float(l2_gbs.value), 1e-9),
                     'bytes_per_step': row_bytes,
                     'peak_source': 'sfx_diag_l2_read_gbs, measured in this run (all SMs '
                                    'sweeping a 48 MiB L2-resident buffer)'},
        'evals_max_frame': int(evals.max()), 'evals_min_frame': int(evals.min()),
        'sm_time_busy_frac': float(evals.sum()) / (float(evals.max()) * model_sms(batch)),
        'frame_ms_mean_max': [float(frame_cycles.mean() / SM_CLOCK_KHZ), float(frame_cycles.max() / SM_CLOCK_KHZ)],
        'note': 'per-frame serial latency bound (DESIGN.md section 4): executed flops = blend '
                'rows streamed x 1024 + evaluations x {:.0f}; skipped full-mesh work is not '
                'credited.  The blend rows are shared by all frames and served from L2: '
                'l2_to_sm is that stream against the measured L2 read peak; hbm is the ncu '
                'dram traffic of the launch.  sm_time_busy_frac = sum of evaluations / '
                '(evaluations of the longest frame x SMs): the single-wave tail of ONE launch '
                '(kernel_ms_per_step, this launch alone); sm_time_busy_frac_measured the same from '
                'the per-frame cycle counters; sm_time_busy_frac_in_flight = the frames\' SM-time '
                'over SMs x ms_per_step with the steps overlapping; achieved_in_flight / '
                'frac_in_flight (also under l2_to_sm) = the same flops / rows over ms_per_step, '
                'the launch configuration of value.  While a block streams rows it draws its '
                'SM\'s share of the L2 peak (about 88 of 84 GB/s per SM: DESIGN.md section 4); '
                'the passes are a third of an evaluation'.format(
                    EVAL_FIXED_FLOPS)}
    coll_stats = batch.coll_stats()
    if coll_stats is not None:
        cs = coll_stats.cpu().numpy()
        roofline['collision_candidates_per_frame_median_max'] = [int(np.median(cs[:, 0])), int(cs[:, 0].max())]
        roofline['collision_touched_vertices_median_max'] = [int(np.median(cs[:, 1])), int(cs[:, 1].max())]

    # ---- per-rank evidence of the straggler (N > 1: all ranks) -------------------------------
    mine = torch.tensor([total_ms_rank / args.steps, kern_ms, float(evals.max()),
                         float(evals.mean()), float(len(plan.flip_ids))],
                        dtype=torch.float64, device=dev)
    if world > 1:
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    per_rank = [{'rank': i, 'ms_per_step': float(t[0]), 'kernel_ms': float(t[1]),
                 'evals_max_frame': int(t[2]), 'evals_mean': float(t[3]),
                 'frames_second_orientation': int(t[4])} for i, t in enumerate(allr)]

    value = world * B * args.steps / (total_ms * 1e-3)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    # SM-time the frames of one step occupy (clock64 per frame, one block each) over SMs x step time
    sm_ms = float(frame_cycles.sum() / SM_CLOCK_KHZ)
    roofline['sm_time_busy_frac_measured'] = sm_ms / (model_sms(batch) * kern_ms)
    roofline['sm_time_busy_frac_in_flight'] = sm_ms / (model_sms(batch) * (total_ms / args.steps))
    # the same accounting at the launch configuration of `value` (steps overlapping): this batch's
    # executed flops / streamed rows over the in-flight step time
    step_s = total_ms / args.steps * 1e-3
    roofline['achieved_in_flight'] = flops / step_s / 1e12
    roofline['frac_in_flight'] = flops / step_s / 1e12 / tf_peak
    roofline['l2_to_sm']['achieved_in_flight'] = row_bytes / step_s / 1e9
    roofline['l2_to_sm']['frac_in_flight'] = row_bytes / step_s / 1e9 / max(float(l2_gbs.value), 1e-9)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(B, world, args.interpenetration, args.vposer,
                                  args.regression_prior, depth, NB),
        'value_by_steps_in_flight': dict(
            [('1 (clusters on)', world * B * args.steps / (total_ms_serial * 1e-3))] +
            [(str(d_), world * B * args.steps / (ms_ * 1e-3)) for d_, ms_ in sorted(by_depth.items())] +
            [(str(depth), world * B * args.steps / (total_ms * 1e-3))]),
        'value_one_step_at_a_time': world * B * args.steps / (total_ms_serial * 1e-3),
        'ms_per_step_one_step_at_a_time': total_ms_serial / args.steps,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(out.h2d_bytes),
                'd2h_bytes_per_step': int(out.d2h_bytes), 'ms_per_step': e2e_ms / args.steps,
                'value_one_step_at_a_time': world * B * args.steps / (e2e_serial_ms * 1e-3)},
        'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline,
        'evals_per_frame': evals_per_frame, 'per_rank': per_rank,
        'fit': {'final_loss_median': float(np.median(out.loss)),
                'frames_with_nan_or_inf': int((out.flags != 0).sum()),
                'frames_second_orientation': int(len(plan.flip_ids))},
    }
    if total_ms_other is not None:
        line['value_' + other] = world * B * args.steps / (total_ms_other * 1e-3)
        line['ms_per_step_' + other] = total_ms_other / args.steps
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        procs = os.cpu_count() or 1
        if args.interpenetration:
            t_on, t_off = oracle_coll_eval_seconds(cfg, kp, expose, pixie, n_evals=3)
            per_frame = evals_per_frame * (t_off + 2.0 / 3.0 * (t_on - t_off))
            line['cpu_baseline'] = {
                'value': procs / per_frame, 'unit': UNIT, 'cores': procs, 'kind': 'port',
                'sample': 'extrapolated from 3 timed closure evaluations of frame 0 by the oracle '
                          'port: {:.2f} s per evaluation with the interpenetration term on (restated '
                          'mesh_intersection search in numpy), {:.3f} s without; x the engine\'s {:.0f} '
                          'evaluations per frame (term on in 2 of 3 stages); {} single-threaded '
                          'processes assumed.  A full CPU fit of one frame would take ~{:.0f} min'
                          .format(t_on, t_off, evals_per_frame, procs, per_frame / 60.0)}
        else:
            sample = list(range(min(args.cpu_frames or procs, B)))
            try:
                pool = OraclePool(cfg, kp, expose, pixie, min(procs, len(sample)))
                wall, res = pool.round(sample, want_result=True)
                pool.close()
                line['cpu_baseline'] = {
                    'value': len(sample) / wall, 'unit': UNIT, 'cores': pool.procs, 'kind': 'port',
                    'sample': 'frames 0..{} of the same batch fitted by the oracle port of the '
                              'reference, one single-threaded process per host core ({} processes, '
                              'one frame each); {:.1f} s wall, mean {:.0f} evals/frame'.format(
                                  len(sample) - 1, pool.procs, wall,
                                  float(np.mean([r['evals'] for r in res])))}
            except Exception as exc:   # no worker processes on this host: one process, all threads
                sample = sample[:2]
                res = time_oracle_frames(cfg, kp, expose, pixie, sample, procs, want_result=True)
                secs = [r['secs'] for r in res]
                line['cpu_baseline'] = {
                    'value': len(secs) / float(np.sum(secs)), 'unit': UNIT, 'cores': procs,
                    'kind': 'port',
                    'sample': 'frames 0..{} fitted sequentially by the oracle port ({} torch '
                              'threads; worker processes unavailable: {}); {:.1f} s'.format(
                                  len(secs) - 1, procs, type(exc).__name__, float(np.sum(secs)))}
            line['parity'] = parity_block(out, res, sample, kp, focal)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def model_sms(batch):
    import torch
    return torch.cuda.get_device_properties(batch.model.device).multi_processor_count


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None,
                    help='timed steps (default 20; 5 for --impl reference: 14 s of CPU work each)')
    ap.add_argument('--warmup', type=int, default=None, help='untimed steps (default 5; 3 for --impl reference)')
    ap.add_argument('--frames', type=int, default=128, help='frames per GPU')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-frames', type=int, default=0,
                    help='frames of the CPU baseline sample (default: one per host core)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--regression-prior', action='store_true',
                    help='config 5\'s per-frame setup on the config-2 batch: synthetic "combined" '
                         'regression prior + camera prior (round 1\'s default line)')
    ap.add_argument('--interpenetration', action='store_true',
                    help='BASELINE config 4: the regression-prior workload with the '
                         'interpenetration term on (not the default bench line; the CPU baseline '
                         'is extrapolated from timed single evaluations)')
    ap.add_argument('--vposer', action='store_true',
                    help='BASELINE config 3: 5-stage schedule with the VPoser latent pose prior '
                         '(not the default bench line)')
    ap.add_argument('--two-loop', default=None, choices=['exact', 'gram'],
                    help="L-BFGS direction of `value`: 'gram' (default; coefficient-space "
                         "recursion) or 'exact' (the reference's operation order); the other "
                         "one is reported as value_<mode>")
    ap.add_argument('--single-mode', action='store_true', help='time only the default two-loop mode')
    ap.add_argument('--batches', type=int, default=2,
                    help='distinct synthetic batches per step slot: steps cycle through batches x depth '
                         'different batches')
    ap.add_argument('--depth', type=int, default=6,
                    help='steps in flight (each on its own FrameBatch and CUDA stream); 1 = one batch '
                         'at a time')
    ap.add_argument('--traffic', type=float, default=None,
                    help='dram bytes per launch from an ncu capture (profiles/), recorded as-is')
    args = ap.parse_args()
    # with six steps in flight a short run is mostly pipeline fill and drain (the timed region is
    # bracketed by synchronisations): the engine arm defaults to the driver's 20 / 5
    if args.steps is None:
        args.steps = 5 if args.impl == 'reference' else 20
    if args.warmup is None:
        args.warmup = 3 if args.impl == 'reference' else 5
    global TWO_LOOP
    if args.two_loop:
        TWO_LOOP = args.two_loop
    elif args.interpenetration and 'SFX_TWO_LOOP' not in os.environ:
        # config 4 was measured (profiles/r01o_*) with the exact recursion; its time is the
        # collision search, not the two-loop
        TWO_LOOP = 'exact'
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
