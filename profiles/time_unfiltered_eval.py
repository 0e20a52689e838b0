"""Times one closure evaluation with the interpenetration term on the golden self-contact pose:
with FilterFaces (8 343 pairs) and without (109 412 pairs; the reference's path when part_segm_fn
is empty), float32, 8 frames (one block each), CUDA events."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from tests import common as Cm
from smplifyx_b200 import engine
B = 8
for name in ('filtered', 'unfiltered'):
    ev = Cm.golden('ref_eval_coll_f32.npz' if name == 'filtered' else 'ref_eval_collnf_f32.npz')
    md = Cm.model_data() if name == 'filtered' else Cm.synthetic.without_degenerate_faces(Cm.model_data())
    model = engine.Model(md, Cm.joint_map(), dtype=torch.float32, **Cm.MODEL_KW)
    if name == 'filtered':
        segm, par, ign = Cm.coll_segmentation()
        model.set_collision(segm, par, [[int(x) for x in p.split(',')] for p in ign])
    else:
        model.set_collision_unfiltered()
    batch = engine.FrameBatch(model, B)
    batch.enable_collisions()
    I = Cm.coll_case_inputs(ev, 'coll' if name == 'filtered' else 'collnf')
    rep = lambda a: np.repeat(np.asarray(a)[None], B, axis=0)
    kp = np.concatenate([I['gt'], I['conf'][:, None]], axis=1)
    batch.set_targets(rep(kp), rep(I['jw']), rep(I['lowconf']), rep(I['init_mask']), rep(I['cam']), None)
    batch.set_params(rep(I['x']))
    ts = []
    for it in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); loss, grad, _ = batch.eval(I['stage']); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print('%-10s one evaluation of %d frames: median %.2f ms  (loss %.1f, reference %.1f)' % (
        name, B, np.median(ts[1:]), float(loss[0]), float(ev[('coll' if name == 'filtered' else 'collnf') + '/loss'])))
