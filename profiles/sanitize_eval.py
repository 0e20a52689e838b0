"""One closure evaluation with the interpenetration term on two frames -- the smallest run that
touches every kernel path of the term; meant for compute-sanitizer:

    compute-sanitizer --tool memcheck python profiles/sanitize_eval.py
"""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
import torch

from tests import common as Cm
from smplifyx_b200 import engine

ev = Cm.golden('ref_eval_coll_f32.npz')
model = engine.Model(Cm.model_data(), Cm.joint_map(), dtype=torch.float32, **Cm.MODEL_KW)
model.set_collision(*Cm.coll_segmentation())
batch = engine.FrameBatch(model, 2)
batch.enable_collisions()
I = Cm.coll_case_inputs(ev, 'coll')
rep = lambda a: np.repeat(np.asarray(a)[None], 2, axis=0)
kp = np.concatenate([I['gt'], I['conf'][:, None]], axis=1)
batch.set_targets(rep(kp), rep(I['jw']), rep(I['lowconf']), rep(I['init_mask']), rep(I['cam']))
batch.set_params(rep(I['x']))
loss, grad, _ = batch.eval(I['stage'])
torch.cuda.synchronize()
print('loss', loss.cpu().numpy(), 'reference', float(ev['coll/loss']))
I['stage'].maxiters = 1
I['stage'].max_iter = 2
final = batch.fit_stage(I['stage'])
torch.cuda.synchronize()
print('stage', final.cpu().numpy(), 'flags', batch.flags().cpu().numpy())
