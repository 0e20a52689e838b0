"""Cycle-counter profile of the per-frame pipeline kernel (clock64 around the phases of an
evaluation, per frame; thread 0's view).  Needs the profiling build of the library next to the
product one:

    cd smplify-x-partial_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 \
        -std=c++17 --extended-lambda -Xcompiler -fPIC -shared -DSFX_CYCLE_PROF \
        -o libsfx_prof.so sfx_lib.cu

    python profiles/prof_cycles.py [--interpenetration]        # on a B200 (gpurun)

Numbers taken with this build are diagnostics, never bench values."""
import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from smplifyx_b200 import _native as N
N.LIB_PATH = os.path.join(os.path.dirname(N.LIB_PATH), sys.argv[sys.argv.index('--lib') + 1] if '--lib' in sys.argv else 'libsfx_prof.so')
import bench
from smplifyx_b200 import engine, fit_frames as FF, synthetic, utils as U
COLL = '--interpenetration' in sys.argv
REG = '--regression-prior' in sys.argv
cfg = bench.bench_cfg(COLL, False, REG); B = 128
jm = U.smpl_to_annotation('smplx', True, True, True, 'coco25')
md = synthetic.cached_smplx_like(0, bench.COLL_POSE_CORRECTIVE_SCALE if COLL else 1.0)
model = engine.Model(md, jm, dtype=torch.float32, **bench.MODEL_KW)
batch = engine.FrameBatch(model, B); L = batch.L
gt, rng = bench.ground_truth(B, 0, 0.2 if COLL else 1.0)
cam_st = N.make_stage(L, N.CAMERA_STAGE_BLOCKS, loss_kind=N.LOSS_CAMERA_INIT)
zc = np.zeros((B, 16)); zc[:,0:2]=1; zc[:,4:13]=np.eye(3).reshape(-1)
xg = bench.gt_param_matrix(L, gt); xg[:, L.off_camt+2]=1
batch.set_targets(np.zeros((B,135,3)), np.zeros((B,135)), np.zeros((B,135),np.uint8), np.zeros((B,135),np.uint8), zc, None)
batch.set_params(xg)
_,_,j3 = batch.eval(cam_st, want_joints=True)
kp, ex, px = bench.observations(gt, j3.cpu().numpy().astype(np.float64), rng)
if not cfg.get('regression_prior'): ex = px = None
part_segm = synthetic.parts_segm_like(md) if COLL else None
plan = FF.FitPlan(L, 135, kp, 600, 800, cfg, ex, px, None, np.float32, part_segm=part_segm)
FF.upload(batch, plan)
x0 = batch.params_tensor().clone()
for it in range(2):
    batch.params_tensor().copy_(x0); batch.reset_counters()
    FF.run(batch, plan, True); torch.cuda.synchronize()
lib = batch.lib; lib.sfx_batch_prof_dev.restype = C.c_void_p; lib.sfx_batch_prof_dev.argtypes=[C.c_void_p]
ptr = lib.sfx_batch_prof_dev(batch.h)
prof = engine._wrap(ptr, (B*128,), torch.int32, model.device, batch).cpu().numpy().view(np.int64).reshape(B,64)
laps = prof[:, 16:]
ev = batch.evals().cpu().numpy()
tot = prof[:,4].astype(float)
i = np.argmax(tot)
print('slowest frame', i, 'evals', ev[i], 'cycles total %.3g (%.1f ms @1.965GHz)' % (tot[i], tot[i]/1.965e6))
for name, k in (('eval',0),('two_loop',1),('blend_fwd',2),('blend_adj',3),('chain_fwd_w0',5),('stream_fwd_w1',6),('chain_adj_w0',7),
                ('coll_skin|prologue',8),('coll_boxes|skin+proj+adj',9),('coll_narrow|eval_tail',10),('coll_vgather|gram_dots',11),('coll_skin_adj|gram_chain',12),('coll_walk_t0|gram_combine',13),('coll_pairs_t0',14)):
    print('%-10s slowest: %5.1f%%   all frames: %5.1f%%   per eval (slowest) %.1f us' % (name, 100*prof[i,k]/tot[i], 100*prof[:,k].sum()/tot.sum(), prof[i,k]/ev[i]/1965.))
print('mean frame total ms', tot.mean()/1.965e6, 'median evals', np.median(ev))
print('evals sorted', np.sort(ev)[::-1][:24], 'flip frames', len(plan.flip_ids), 'their evals', ev[plan.flip_ids])
print('frame ms sorted', np.round(np.sort(tot)[::-1][:24]/1.965e6, 1))

LAPS = ['LS logic -> eval entry (scatter)', 'prologue: pose vector, hand PCA, shape', 'prologue: rodrigues, rest joints, yaw row',
        'prologue: blend coefficients, rel (+contour tables)', 'blend forward || chain', 'skinning of support slots', 'extra joints + landmarks',
        'projection + data term', 'nine sums (multi_sum)', 'dX (keypoints -> joints)', 'dvert', 'dvp + dA', 'blend adjoint || chain adjoint',
        'rodrigues adjoint, rest-joint adjoint', 'shape gradient partials', 'priors, gradient vector, loss', '', '', '', '',
        'line-search logic between probes', 'gather_grad after eval', 'probe: copy + g.d', 'end-of-iteration tests', 'history update (y, s, ys, yy)',
        'two-loop recursion', 'step length, gtd, copies', 'line-search epilogue']
print('lap timers, slowest frame (us per evaluation) | all frames (share of frame time)')
for k, name in enumerate(LAPS):
    if name:
        print('  %2d %-52s %6.2f us   %5.1f%%' % (k, name, laps[i, k] / ev[i] / 1965., 100 * laps[:, k].sum() / tot.sum()))
print('  laps total %.1f us per evaluation (frame total %.1f)' % (laps[i].sum() / ev[i] / 1965., tot[i] / ev[i] / 1965.))

cs = batch.coll_stats()
if cs is not None:
    cs = cs.cpu().numpy()
    print('candidates per frame (max over evals): median %d max %d; touched vertices: median %d max %d' % (np.median(cs[:,0]), cs[:,0].max(), np.median(cs[:,1]), cs[:,1].max()))
    print('warp 0 (1/16 of the candidates): sweep iterations median %d max %d; listed partners median %d max %d; slowest frame: cand %d iters %d hits %d' % (np.median(cs[:,2]), cs[:,2].max(), np.median(cs[:,3]), cs[:,3].max(), cs[i,0], cs[i,2], cs[i,3]))
