"""Times the full-mesh forward (sfx_forward_mesh: pose prologue + blend + skinning) for 128 frames:
the fused tensor-core kernel (default), the round-1 pair (SFX_MESH_UNFUSED=1: tcgen05 blend + SIMT
skinning) and the float32 SIMT kernels (SFX_MESH_SIMT=1).  CUDA events, L2 flushed between runs."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from smplifyx_b200 import engine, synthetic, utils as U, _native as N
if '--lib' in sys.argv:     # another build of the library (csrc/<name>), e.g. -DSFX_FU_TF=48
    N.LIB_PATH = os.path.join(os.path.dirname(N.LIB_PATH), sys.argv[sys.argv.index('--lib') + 1])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
jm = U.smpl_to_annotation('smplx', True, True, True, 'coco25')
model = engine.Model(synthetic.cached_smplx_like(0), jm, dtype=torch.float32, **bench.MODEL_KW)
batch = engine.FrameBatch(model, B); L = batch.L
gt, rng = bench.ground_truth(B, 0)
x = bench.gt_param_matrix(L, gt); x[:, L.off_camt + 2] = 3
zc = np.zeros((B, 16)); zc[:, 0:2] = 1; zc[:, 4:13] = np.eye(3).reshape(-1)
batch.set_targets(np.zeros((B, 135, 3)), np.zeros((B, 135)), np.zeros((B, 135), np.uint8), np.zeros((B, 135), np.uint8), zc, None)
batch.set_params(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
out = (torch.empty((B, model.V, 3), device='cuda'), None)
ref = None
for name, env in (('simt', {'SFX_MESH_SIMT': '1'}), ('unfused', {'SFX_MESH_UNFUSED': '1'}), ('fused', {})):
    for k in ('SFX_MESH_SIMT', 'SFX_MESH_UNFUSED'):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for it in range(8):
        flush.fill_(it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); v, _ = batch.forward_mesh(want_joints=False, out=out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    v = v.clone()
    if ref is None: ref = v
    print('%-8s B=%d  median %.1f us  min %.1f us   max |v - simt| %.3g m' % (name, B, np.median(ts[2:]), min(ts[2:]), (v - ref).abs().max().item()))
