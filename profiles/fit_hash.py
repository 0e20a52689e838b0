"""Bit-level fingerprint and kernel time of the default bench workload (B = 128, config 2), for
the two L-BFGS direction modes and with / without wide frames.  Used while optimising the
pipeline kernel: a change that is meant to keep the arithmetic must leave every hash unchanged.

    python profiles/fit_hash.py [B] [--lib libsfx_x.so]        # on a B200 (gpurun)
"""
import sys, os, hashlib
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from smplifyx_b200 import _native as N
if '--lib' in sys.argv:
    N.LIB_PATH = os.path.join(os.path.dirname(N.LIB_PATH), sys.argv[sys.argv.index('--lib') + 1])
import bench
from smplifyx_b200 import engine, fit_frames as FF, synthetic, utils as U

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 128
jm = U.smpl_to_annotation('smplx', True, True, True, 'coco25')
md = synthetic.cached_smplx_like(0, 1.0)
model = engine.Model(md, jm, dtype=torch.float32, **bench.MODEL_KW)
batch = engine.FrameBatch(model, B); L = batch.L
gt, rng = bench.ground_truth(B, 0, 1.0)
cam_st = N.make_stage(L, N.CAMERA_STAGE_BLOCKS, loss_kind=N.LOSS_CAMERA_INIT)
zc = np.zeros((B, 16)); zc[:, 0:2] = 1; zc[:, 4:13] = np.eye(3).reshape(-1)
xg = bench.gt_param_matrix(L, gt); xg[:, L.off_camt + 2] = 1
batch.set_targets(np.zeros((B, 135, 3)), np.zeros((B, 135)), np.zeros((B, 135), np.uint8),
                  np.zeros((B, 135), np.uint8), zc, None)
batch.set_params(xg)
_, _, j3 = batch.eval(cam_st, want_joints=True)
kp, ex, px = bench.observations(gt, j3.cpu().numpy().astype(np.float64), rng)

for two_loop in ('gram', 'exact'):
    for wide in ('auto', 'off'):
        cfg = bench.bench_cfg(False, False, False, two_loop=two_loop)
        cfg['wide_frames'] = wide
        plan = FF.FitPlan(L, 135, kp, 600, 800, cfg, None, None, None, np.float32)
        FF.upload(batch, plan)
        x0 = batch.params_tensor().clone()
        ms = []
        for it in range(3):
            batch.params_tensor().copy_(x0); batch.reset_counters()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); FF.run(batch, plan, False); b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        x = batch.params_tensor().cpu().numpy()
        ev = batch.evals().cpu().numpy()
        print('two_loop %-5s wide %-4s  params sha1 %s  evals sum %d max %d  kernel ms %s' % (
            two_loop, wide, hashlib.sha1(x.tobytes()).hexdigest()[:16], ev.sum(), ev.max(),
            ' '.join('%.2f' % m for m in ms)), flush=True)
