"""Engine vs the oracle port of the reference ON THE BENCH WORKLOADS (bench.py; BASELINE configs
2, 3 and a shard of config 5): the frames of tests/golden/bench_parity.npz, fitted by
oracle/fit_port.py in the authoring container (tests/golden/make_bench_parity.py), against
``fit_frames`` on the device through the C ABI -- both two-loop modes.

float32 fits are chaotic: the fixture also holds the oracle's OWN fit of the same frames started
from keypoints one float32 ulp apart, and the thresholds below are set from that
oracle-vs-oracle envelope (measured per workload), not from a guess:

    workload   loss rel median / max     mean vertex distance median / max
    cfg2       3.3e-2 / 0.27             10 mm / 50 mm
    cfg2_reg   4.2e-3 / 0.20             2.1 mm / 32 mm
    cfg3       4.6e-2 / 1.2              21 mm / 264 mm
    cfg5       3.1e-3 / 0.54             1.8 mm / 12 mm
"""
import numpy as np
import pytest
import torch

from tests import common as Cm

pytestmark = pytest.mark.gpu

VSTRIDE = 5
H_IMG, W_IMG = 600, 800


def _case_inputs(G, name):
    kp = G[name + '/keypoints']
    n = kp.shape[0]
    expose = pixie = None
    if name + '/expose/body_pose' in G:
        expose = [{k: G['{}/expose/{}'.format(name, k)][b] for k in
                   ('body_pose', 'global_orient', 'transl', 'center')} for b in range(n)]
        pixie = [{k: G['{}/pixie/{}'.format(name, k)][b] for k in ('body_pose', 'global_pose')}
                 for b in range(n)]
    return kp, expose, pixie


def _metrics(loss_a, verts_a, loss_b, verts_b):
    rel = np.abs(loss_a - loss_b) / np.maximum(np.abs(loss_b), 1.0)
    d = np.sqrt(((verts_a.astype(np.float64) - verts_b.astype(np.float64)) ** 2).sum(-1)).mean(1)
    return rel, d


def _project(joints, cam_t, center, focal):
    p = joints.astype(np.float64) + cam_t.astype(np.float64)[:, None]
    return focal * p[:, :, :2] / p[:, :, 2:3] + center.astype(np.float64)[:, None]


CASES = [('cfg2', dict()), ('cfg2_reg', dict(regression_prior=True)),
         ('cfg3', dict(vposer=True)), ('cfg5', dict(regression_prior=True))]


@pytest.mark.parametrize('two_loop', ['gram', 'exact'])
@pytest.mark.parametrize('name,kw', CASES)
def test_bench_workload_against_oracle_fit(name, kw, two_loop):
    import bench
    from smplifyx_b200 import engine, fit_frames as FF, synthetic
    G = Cm.golden('bench_parity.npz')
    cfg = bench.bench_cfg(two_loop=two_loop, **kw)
    kp, expose, pixie = _case_inputs(G, name)
    n = kp.shape[0]
    model = engine.Model(synthetic.cached_smplx_like(0), Cm.joint_map(), dtype=torch.float32,
                         **bench.MODEL_KW)
    vp = None
    if kw.get('vposer'):
        from smplifyx_b200 import vposer as V
        vp = V.VPoser(synthetic.make_vposer_like(seed=2))
        model.set_vposer(vp.weights)
    batch = engine.FrameBatch(model, n, use_vposer=bool(kw.get('vposer')))
    out = FF.fit_frames(batch, kp, H_IMG, W_IMG, cfg, expose, pixie, return_verts=True, vposer=vp)
    assert out.flags.max() == 0
    ref, ulp = name + '/ref/', name + '/ref_ulp/'
    assert np.array_equal(out.n_orient, G[ref + 'n_orient'])       # same orientation decisions
    verts = out.vertices[:, ::VSTRIDE]
    rel, dv = _metrics(out.loss, verts, G[ref + 'loss'], G[ref + 'vertices'])
    erel, edv = _metrics(G[ulp + 'loss'], G[ulp + 'vertices'], G[ref + 'loss'], G[ref + 'vertices'])
    print('{} [{}]: loss rel median {:.3g} max {:.3g} (oracle envelope {:.3g} / {:.3g}); mean '
          'vertex distance median {:.4f} max {:.4f} m (envelope {:.4f} / {:.4f}); evals {:.0f} '
          '(oracle {:.0f}, one ulp apart {:.0f})'.format(
              name, two_loop, np.median(rel), rel.max(), np.median(erel), erel.max(),
              np.median(dv), dv.max(), np.median(edv), edv.max(), out.n_evals.mean(),
              G[ref + 'evals'].mean(), G[ulp + 'evals'].mean()))
    # the engine sits inside the spread the reference shows against itself on these frames
    assert np.median(rel) <= 2.0 * np.median(erel) + 2e-3
    assert np.median(dv) <= 2.0 * np.median(edv) + 1e-3
    assert rel.max() <= 2.0 * erel.max() + 0.05
    assert dv.max() <= 2.0 * edv.max() + 0.01
    # evaluation counts: same workload, same amount of work
    lo = min(G[ref + 'evals'].mean(), G[ulp + 'evals'].mean())
    hi = max(G[ref + 'evals'].mean(), G[ulp + 'evals'].mean())
    assert 0.75 * lo <= out.n_evals.mean() <= 1.25 * hi
    # reprojected keypoints are the well-conditioned quantity: medians within a pixel
    focal = float(cfg.get('focal_length') or np.sqrt(H_IMG ** 2 + W_IMG ** 2))
    cam_t = np.stack([r['camera_translation'].reshape(3) for r in out.results])
    cen = np.stack([r['camera_center'].reshape(2) for r in out.results])
    pe = _project(out.joints, cam_t, cen, focal)
    po = _project(G[ref + 'joints'], G[ref + 'cam_t'], G[ref + 'center'], focal)
    live = kp[:, :, 2] > 0
    dpx = np.sqrt(((pe - po) ** 2).sum(-1))
    pu = _project(G[ulp + 'joints'], G[ulp + 'cam_t'], G[ulp + 'center'], focal)
    epx = np.sqrt(((pu - po) ** 2).sum(-1))
    print('   reprojection distance to the oracle fit: median {:.3f} px (envelope {:.3f} px)'.format(
        np.median(dpx[live]), np.median(epx[live])))
    assert np.median(dpx[live]) <= 2.0 * np.median(epx[live]) + 0.5
