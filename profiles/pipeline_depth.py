"""Throughput of the default bench workload (B = 128 per step) with 1, 2, 3 steps in flight:
each step is a complete fit of one batch, consecutive steps alternate between independent
FrameBatch objects on their own streams, so the straggler tail of one step overlaps the next
step's frames.  Diagnostics (bench.py reports the same through --depth).

    python profiles/pipeline_depth.py [B] [--lib libsfx_x.so]
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
STEPS = 24
jm = U.smpl_to_annotation('smplx', True, True, True, 'coco25')
md = synthetic.cached_smplx_like(0, 1.0)
model = engine.Model(md, jm, dtype=torch.float32, **bench.MODEL_KW)
MAXD = 4
batches = [engine.FrameBatch(model, B) for _ in range(MAXD)]
L = batches[0].L
gt, rng = bench.ground_truth(B, 0, 1.0)
cam_st = N.make_stage(L, N.CAMERA_STAGE_BLOCKS, loss_kind=N.LOSS_CAMERA_INIT)
zc = np.zeros((B, 16)); zc[:, 0:2] = 1; zc[:, 4:13] = np.eye(3).reshape(-1)
xg = bench.gt_param_matrix(L, gt); xg[:, L.off_camt + 2] = 1
b0 = batches[0]
b0.set_targets(np.zeros((B, 135, 3)), np.zeros((B, 135)), np.zeros((B, 135), np.uint8),
               np.zeros((B, 135), np.uint8), zc, None)
b0.set_params(xg)
_, _, j3 = b0.eval(cam_st, want_joints=True)
kp, ex, px = bench.observations(gt, j3.cpu().numpy().astype(np.float64), rng)
streams = [torch.cuda.Stream() for _ in range(MAXD)]
for wide in ('auto', 'off'):
  cfg = bench.bench_cfg(False, False, False)
  cfg['wide_frames'] = wide
  plans, x0 = [], []
  for b in batches:
    plan = FF.FitPlan(L, 135, kp, 600, 800, cfg, None, None, None, np.float32)
    FF.upload(b, plan)
    plans.append(plan)
    x0.append(b.params_tensor().clone())
  torch.cuda.synchronize()
  print('wide frames', wide)
  for depth in (1, 2, 3, 4):
      for rep in range(2):
          torch.cuda.synchronize()
          a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          a.record()
          for s in streams[:depth]:
              s.wait_event(a)
          for i in range(STEPS):
              k = i % depth
              with torch.cuda.stream(streams[k]):
                  batches[k].params_tensor().copy_(x0[k])
                  FF.run(batches[k], plans[k], True)
          for s in streams[:depth]:
              torch.cuda.current_stream().wait_stream(s)
          e.record()
          torch.cuda.synchronize()
      ms = a.elapsed_time(e) / STEPS
      h = [hashlib.sha1(b.params_tensor().cpu().numpy().tobytes()).hexdigest()[:12] for b in batches[:depth]]
      print('depth %d: %.2f ms per step, %.0f frames/s   params %s' % (depth, ms, B / ms * 1e3, ' '.join(h)), flush=True)
