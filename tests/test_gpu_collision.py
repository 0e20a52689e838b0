"""Interpenetration term on the GPU: one closure evaluation with the term
against the reference's SMPLifyLoss (fitting.py:437-455) driving the restated
mesh_intersection package (oracle/isect_port.py; parity unpinned at that third-party boundary),
golden vectors tests/golden/ref_eval_coll_*.npz."""
import numpy as np
import pytest
import torch

from tests import common as Cm
from smplifyx_b200 import _native as N

pytestmark = pytest.mark.gpu


def _batch(dtype, B=2):
    from smplifyx_b200 import engine
    model = engine.Model(Cm.model_data(), Cm.joint_map(), dtype=dtype, **Cm.MODEL_KW)
    model.set_collision(*Cm.coll_segmentation())
    batch = engine.FrameBatch(model, B)
    batch.enable_collisions()
    return model, batch


def _load(batch, I, B):
    rep = lambda a: np.repeat(np.asarray(a)[None], B, axis=0)
    kp = np.concatenate([I['gt'], I['conf'][:, None]], axis=1)
    batch.set_targets(rep(kp), rep(I['jw']), rep(I['lowconf']), rep(I['init_mask']), rep(I['cam']))
    batch.set_params(rep(I['x']))


@pytest.mark.parametrize('dt,tol_loss,tol_pen,tol_grad', [('f64', 1e-12, 1e-9, 1e-9),
                                                          ('f32', 1e-5, 2e-3, 2e-3)])
def test_eval_with_interpenetration(dt, tol_loss, tol_pen, tol_grad):
    dtype = torch.float64 if dt == 'f64' else torch.float32
    ev = Cm.golden('ref_eval_coll_{}.npz'.format(dt))
    B = 2
    model, batch = _batch(dtype, B)
    out = {}
    for case in ('nocoll', 'coll'):
        I = Cm.coll_case_inputs(ev, case)
        _load(batch, I, B)
        loss, grad, _ = batch.eval(I['stage'])
        out[case] = (loss.cpu().numpy().astype(np.float64), grad.cpu().numpy().astype(np.float64))
        ref = float(ev[case + '/loss'])
        assert np.all(np.abs(out[case][0] - ref) <= tol_loss * abs(ref)), (case, out[case][0], ref)
    assert int(batch.flags().cpu().numpy().max()) & N.SFX_FLAG_COLL_OVERFLOW == 0
    L = Cm.layout()
    # the term on its own: difference of the two evaluations, value and gradient
    pen_ref = float(ev['coll/loss']) - float(ev['nocoll/loss'])
    g_ref = Cm.golden_grad_vector(L, ev, 'coll') - Cm.golden_grad_vector(L, ev, 'nocoll')
    assert pen_ref > 100 and len(ev['coll/pairs']) > 1000
    for f in range(B):
        pen = out['coll'][0][f] - out['nocoll'][0][f]
        g = out['coll'][1][f] - out['nocoll'][1][f]
        assert abs(pen - pen_ref) <= tol_pen * pen_ref + tol_loss * abs(float(ev['coll/loss'])), (pen, pen_ref)
        assert np.abs(g - g_ref).max() <= tol_grad * np.abs(g_ref).max() + \
            tol_loss * np.abs(out['coll'][1][f]).max(), np.abs(g - g_ref).max()
    # frames are independent and launches deterministic
    assert np.array_equal(out['coll'][1][0], out['coll'][1][1])
    I = Cm.coll_case_inputs(ev, 'coll')
    loss2, grad2, _ = batch.eval(I['stage'])
    assert np.array_equal(grad2.cpu().numpy().astype(np.float64), out['coll'][1])


def test_stage_needs_workspace():
    from smplifyx_b200 import engine
    ev = Cm.golden('ref_eval_coll_f32.npz')
    model = engine.Model(Cm.model_data(), Cm.joint_map(), dtype=torch.float32, **Cm.MODEL_KW)
    batch = engine.FrameBatch(model, 1)
    I = Cm.coll_case_inputs(ev, 'coll')
    _load(batch, I, 1)
    with pytest.raises(RuntimeError, match='sfx_model_set_collision'):
        batch.eval(I['stage'])
    with pytest.raises(RuntimeError, match='sfx_model_set_collision'):
        batch.enable_collisions()


def test_short_fit_with_interpenetration_reduces_penalty():
    """A body stage with the term on: the fit runs on the device, finishes without flags and
    lowers the total loss; the penalty at the fitted pose is below the start's."""
    ev = Cm.golden('ref_eval_coll_f32.npz')
    B = 3
    model, batch = _batch(torch.float32, B)
    I = Cm.coll_case_inputs(ev, 'coll')
    I['stage'].coll_loss_weight = 10.0
    I['stage'].maxiters = 5
    _load(batch, I, B)
    l0, _, _ = batch.eval(I['stage'])
    final = batch.fit_stage(I['stage']).cpu().numpy()
    l1, _, _ = batch.eval(I['stage'])
    assert np.all(np.isfinite(final))
    assert np.all(l1.cpu().numpy() < l0.cpu().numpy())
    assert int(batch.flags().cpu().numpy().max()) == 0
    x = batch.get_params()
    assert np.array_equal(x[0], x[1]) and np.array_equal(x[0], x[2])


def test_fit_frames_with_interpenetration_pipeline_equals_staged():
    """The whole multi-stage flow with the term on (coll_loss_weights 0 / 0.1 / 1.0 as in the
    shipped yaml) through the batched driver: one persistent launch == one launch per stage,
    bit for bit; no flags; the term lowers the penalty left at the end of the fit."""
    import json
    from smplifyx_b200 import engine, fit_frames as FF, synthetic
    inp = Cm.golden('demo_inputs.npz')
    ref = Cm.golden('ref_fit_02.npz')
    cfg = json.loads(str(ref['cfg_json']))
    md = Cm.model_data()
    segm, par, ign = Cm.coll_segmentation()
    cfg.update(interpenetration=True, coll_loss_weights=[0.0, 0.1, 1.0], df_cone_height=1e-4,
               max_collisions=128, penalize_outside=True, point2plane=False, ign_part_pairs=ign,
               maxiters=8)
    part_segm = {'segm': segm, 'parents': par}
    model = engine.Model(md, Cm.joint_map(), dtype=torch.float32, **Cm.MODEL_KW)
    frames = ['02_cropped', '18_cropped']
    data = []
    for fr in frames:
        expose = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/expose/')}
        pixie = {k.split('/')[-1]: inp[k] for k in inp if k.startswith(fr + '/pixie/')}
        H, W = [int(v) for v in inp[fr + '/HW']]
        data.append((inp[fr + '/keypoints'], H, W, expose, pixie))
    kp = np.stack([d[0] for d in data])
    res = []
    for runner in (FF.run, FF.run_staged):
        batch = engine.FrameBatch(model, 2)
        plan = FF.FitPlan(batch.L, model.K, kp, [d[1] for d in data], [d[2] for d in data], cfg,
                          [d[3] for d in data], [d[4] for d in data], None, np.float32,
                          part_segm=part_segm)
        assert plan.stages[2].coll_loss_weight == 1.0 and plan.stages[0].coll_loss_weight == 0.0
        FF.upload(batch, plan)
        cam_loss, verts, joints, _ = runner(batch, plan, True)
        res.append(FF.download(batch, plan, cam_loss, verts, joints))
    a, b = res
    assert a.flags.max() == 0 and np.all(np.isfinite(a.params))
    assert np.array_equal(a.params, b.params) and np.array_equal(a.loss, b.loss)
    assert np.array_equal(a.n_evals, b.n_evals) and np.array_equal(a.vertices, b.vertices)
    # without part_segm the plan carries the reference's unfiltered term (fit_single_frame.py:317-328)
    assert FF.FitPlan(batch.L, model.K, kp, 600, 800, cfg, [d[3] for d in data],
                      [d[4] for d in data], None, np.float32).collision == 'unfiltered'


@pytest.mark.parametrize('dt,tol_loss,tol_pen,tol_grad', [('f64', 1e-12, 1e-9, 1e-9),
                                                          ('f32', 1e-5, 2e-3, 2e-3)])
def test_eval_with_interpenetration_without_the_face_filter(dt, tol_loss, tol_pen, tol_grad):
    """The reference's path when no part_segm_fn is given (fit_single_frame.py:317-328,
    fitting.py:449-450: filter_faces = None): every intersecting pair of triangles that share no
    vertex is penalised -- 109 412 pairs on the golden pose against 8 343 with the filter
    (tests/golden/ref_eval_collnf_*.npz, the reference's SMPLifyLoss on the restated package; on
    the synthetic mesh without its 98 area-less cap faces, whose cones are NaN on both sides).  On
    the device the faces are only grouped (by dominant joint) for the broad phase; every face is a
    candidate, so the sweep arrays live in the block's global workspace."""
    from smplifyx_b200 import engine
    dtype = torch.float64 if dt == 'f64' else torch.float32
    ev = Cm.golden('ref_eval_collnf_{}.npz'.format(dt))
    B = 2
    model = engine.Model(Cm.synthetic.without_degenerate_faces(Cm.model_data()), Cm.joint_map(),
                         dtype=dtype, **Cm.MODEL_KW)
    model.set_collision_unfiltered()
    batch = engine.FrameBatch(model, B)
    batch.enable_collisions()
    out = {}
    for case in ('nocoll', 'collnf'):
        I = Cm.coll_case_inputs(ev, case)
        _load(batch, I, B)
        loss, grad, _ = batch.eval(I['stage'])
        out[case] = (loss.cpu().numpy().astype(np.float64), grad.cpu().numpy().astype(np.float64))
        ref = float(ev[case + '/loss'])
        assert np.all(np.abs(out[case][0] - ref) <= tol_loss * abs(ref)), (case, out[case][0], ref)
    assert int(batch.flags().cpu().numpy().max()) & N.SFX_FLAG_COLL_OVERFLOW == 0
    L = Cm.layout()
    pen_ref = float(ev['collnf/loss']) - float(ev['nocoll/loss'])
    g_ref = Cm.golden_grad_vector(L, ev, 'collnf') - Cm.golden_grad_vector(L, ev, 'nocoll')
    assert pen_ref > 1000 and len(ev['collnf/pairs']) > 100000
    for f in range(B):
        pen = out['collnf'][0][f] - out['nocoll'][0][f]
        g = out['collnf'][1][f] - out['nocoll'][1][f]
        assert abs(pen - pen_ref) <= tol_pen * pen_ref + tol_loss * abs(float(ev['collnf/loss'])), (pen, pen_ref)
        assert np.abs(g - g_ref).max() <= tol_grad * np.abs(g_ref).max() + \
            tol_loss * np.abs(out['collnf'][1][f]).max(), np.abs(g - g_ref).max()
    assert np.array_equal(out['collnf'][1][0], out['collnf'][1][1])
