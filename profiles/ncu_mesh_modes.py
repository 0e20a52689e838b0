"""Two forward-mesh calls per implementation (fused tensor-core kernel, round-1 pair, SIMT) plus
one closure evaluation, for an ncu capture of every kernel of the mesh / evaluation path."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from smplifyx_b200 import engine, synthetic, utils as U, _native as N
B = 128
jm = U.smpl_to_annotation('smplx', True, True, True, 'coco25')
model = engine.Model(synthetic.cached_smplx_like(0), jm, dtype=torch.float32, **bench.MODEL_KW)
batch = engine.FrameBatch(model, B); L = batch.L
gt, rng = bench.ground_truth(B, 0)
x = bench.gt_param_matrix(L, gt); x[:, L.off_camt + 2] = 3
zc = np.zeros((B, 16)); zc[:, 0:2] = 1; zc[:, 4:13] = np.eye(3).reshape(-1)
batch.set_targets(np.ones((B, 135, 3)), np.ones((B, 135)), np.zeros((B, 135), np.uint8), np.ones((B, 135), np.uint8), zc, None)
batch.set_params(x)
for env in ({}, {'SFX_MESH_UNFUSED': '1'}, {'SFX_MESH_SIMT': '1'}):
    for k in ('SFX_MESH_SIMT', 'SFX_MESH_UNFUSED'):
        os.environ.pop(k, None)
    os.environ.update(env)
    for it in range(2):
        batch.forward_mesh(want_joints=False)
    torch.cuda.synchronize()
st = N.make_stage(L, N.BODY_STAGE_BLOCKS, body_pose_weight=300.0, shape_weight=50.0, hand_joint_weight=0.1, face_joint_weight=2.0)
batch.eval(st); batch.eval(st)
torch.cuda.synchronize()
