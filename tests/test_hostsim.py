"""The kernel source (csrc/sfx_core.cuh) compiled for the host, single-threaded, against the
golden vectors of the unmodified reference.  This checks the evaluation maths (forward pass,
loss stack, analytic adjoint) of the exact code the GPU runs, without a GPU."""
import numpy as np
import pytest

from tests import common as Cm
from tests.hostsim.hostsim import HostSim
from smplifyx_b200 import _native as N


@pytest.fixture(scope='module')
def hs64():
    return HostSim(Cm.model_data(), Cm.joint_map(), use_double=True, **Cm.MODEL_KW)


@pytest.fixture(scope='module')
def hs32():
    return HostSim(Cm.model_data(), Cm.joint_map(), use_double=False, **Cm.MODEL_KW)


def _eval(hs, I):
    return hs.eval(I['stage'], I['x'], I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                   I['init_mask'], I['reg_pose'])


@pytest.mark.parametrize('case', ['l2', 'reg', 'cam', 'camconf'])
def test_eval_f64(hs64, case):
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, case)
    r = _eval(hs64, I)
    ref = float(ev[case + '/loss'])
    assert abs(r['loss'] - ref) <= 1e-13 * abs(ref)
    g_ref = Cm.golden_grad_vector(I['L'], ev, case)
    live = g_ref != 0
    assert np.abs(r['grad'][live] - g_ref[live]).max() <= 1e-12 * np.abs(g_ref).max()
    if case == 'l2':
        assert np.abs(r['joints'] - ev['l2/joints']).max() < 1e-14
        # every parameter block the reference differentiates is live in the engine too
        assert live.sum() == I['L'].np


@pytest.mark.parametrize('case', ['l2', 'reg', 'cam', 'camconf'])
def test_eval_f32(hs32, case):
    ev = Cm.golden('ref_eval_f32.npz')
    I = Cm.eval_case_inputs(ev, case)
    r = _eval(hs32, I)
    ref = float(ev[case + '/loss'])
    assert abs(r['loss'] - ref) <= 1e-5 * abs(ref)
    g_ref = Cm.golden_grad_vector(I['L'], ev, case)
    live = g_ref != 0
    assert np.abs(r['grad'][live] - g_ref[live]).max() <= 5e-4 * np.abs(g_ref).max()


def test_mixture_prior_f64(hs64):
    """MaxMixturePrior branch of the pose prior (prior.py:181-196) against the reference."""
    import torch
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, 'gmm')
    hs64.set_gmm(*Cm.gmm_arrays(torch.float64))
    r = _eval(hs64, I)
    ref = float(ev['gmm/loss'])
    assert abs(r['loss'] - ref) <= 1e-12 * abs(ref)
    g_ref = Cm.golden_grad_vector(I['L'], ev, 'gmm')
    assert np.abs(r['grad'] - g_ref).max() <= 1e-11 * np.abs(g_ref).max()
    # the prior is what makes this case differ from plain L2
    assert abs(ref - float(ev['l2/loss'])) > 1e3


@pytest.mark.parametrize('case', ['vposer', 'vposer_reg'])
def test_vposer_latent_pose_f64(case):
    """Latent pose through the VPoser decoder (fitting.py:235-236) + latent prior (:389-395):
    loss and gradient wrt the 32-D embedding and every other parameter against the reference
    loss stack driving the restated VPoser v1 (oracle/vposer_shim.py)."""
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, case)
    hs = HostSim(Cm.model_data(), Cm.joint_map(), use_double=True, use_vposer=True, **Cm.MODEL_KW)
    hs.set_vposer(Cm.vposer_weights())
    r = hs.eval(I['stage'], I['x'], I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                I['init_mask'], I['reg_pose'])
    ref = float(ev[case + '/loss'])
    assert abs(r['loss'] - ref) <= 1e-12 * abs(ref)
    g_ref = Cm.golden_grad_vector(I['L'], ev, case)
    assert I['L'].n_pose == 32
    assert np.abs(r['grad'] - g_ref).max() <= 1e-10 * np.abs(g_ref).max()
    assert np.abs(r['joints'] - ev[case + '/joints']).max() < 1e-13


@pytest.mark.parametrize('case', ['l2', 'reg', 'cam', 'camconf'])
def test_skipping_dead_support_rows_is_exact(hs64, case):
    """A stage streams only the support rows a keypoint with a non-zero weight depends on
    (zero hand / face weights, undetected keypoints, the camera stage's dozen joints): loss and
    gradient equal those of the evaluation that streams all 675 rows."""
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, case)
    I['stage'].face_joint_weight = 0.0          # as in the first annealing stages
    conf = I['conf'].copy()
    conf[30:40] = 0.0                            # a few undetected hand keypoints
    run = lambda: hs64.eval(I['stage'], I['x'], I['gt'], conf, I['jw'], I['cam'], I['lowconf'],
                            I['init_mask'], I['reg_pose'])
    full = run()
    hs64.lib.hs_set_all_rows(0)
    try:
        lean = run()
    finally:
        hs64.lib.hs_set_all_rows(1)
    assert lean['loss'] == pytest.approx(full['loss'], rel=1e-15)
    assert np.abs(lean['grad'] - full['grad']).max() <= 1e-13 * np.abs(full['grad']).max()
    # joints of live keypoints are untouched; the face block is what was skipped
    eff = I['jw'].copy()                          # per-stage weights (fit_single_frame.py:569-574)
    eff[25:67] = I['stage'].hand_joint_weight
    eff[67:] = I['stage'].face_joint_weight
    eff[I['lowconf'].astype(bool)] = 0
    live = (eff * conf != 0) if case in ('l2', 'reg') else I['init_mask'].astype(bool)
    assert np.array_equal(lean['joints'][live], full['joints'][live])
    assert not np.array_equal(lean['joints'], full['joints'])


def test_gradient_against_finite_differences(hs64):
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, 'reg')
    r0 = _eval(hs64, I)
    rng = np.random.default_rng(3)
    for i in rng.choice(I['L'].np, size=12, replace=False):
        h = 1e-6
        xp, xm = I['x'].copy(), I['x'].copy()
        xp[i] += h
        xm[i] -= h
        fp = hs64.eval(I['stage'], xp, I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                       I['init_mask'], I['reg_pose'])['loss']
        fm = hs64.eval(I['stage'], xm, I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                       I['init_mask'], I['reg_pose'])['loss']
        fd = (fp - fm) / (2 * h)
        assert abs(fd - r0['grad'][i]) <= 1e-5 * max(1.0, abs(fd)), (i, fd, r0['grad'][i])


def test_stage_fit_inside_reference_envelope(hs32):
    """float32 stage fit of the host build lands inside the reference's own run-to-run
    envelope (tests/golden/ref_envelope.npz)."""
    ev = Cm.golden('ref_eval_f32.npz')
    env = Cm.golden('ref_envelope.npz')
    I = Cm.eval_case_inputs(ev, 'reg')
    r = hs32.fit(I['stage'], I['x'], I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                 I['init_mask'], I['reg_pose'])
    lo, hi = env['stage/final_loss'].min(), env['stage/final_loss'].max()
    assert lo - 0.1 * (hi - lo) < r['loss'] < hi + 0.1 * (hi - lo)
    assert r['flags'] == 0 and 40 < r['n_evals'] < 600


def test_adam_matches_torch(hs64):
    import torch
    from oracle import fit_port as FP
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, 'cam')
    L = I['L']
    st = N.make_stage(L, N.CAMERA_STAGE_BLOCKS, loss_kind=N.LOSS_CAMERA_INIT, opt_kind=N.OPT_ADAM,
                      lr=1e-2, depth_loss_weight=100.0, maxiters=30)
    r = hs64.fit(st, I['x'], I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'], I['init_mask'],
                 I['reg_pose'])
    pb = N.param_blocks(L)
    act = np.concatenate([np.arange(pb[n][0], pb[n][0] + pb[n][1]) for n in N.CAMERA_STAGE_BLOCKS])
    x = torch.tensor(I['x'][act], dtype=torch.float64)
    full = I['x'].copy()
    last = {}

    def closure():
        full[act] = x.detach().numpy()
        e = hs64.eval(st, full, I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                      I['init_mask'], I['reg_pose'])
        last['g'] = torch.tensor(e['grad'][act], dtype=torch.float64)
        return torch.tensor(e['loss'], dtype=torch.float64), last['g']
    FP.run_fitting(FP.AdamPort(x, lr=1e-2), closure, [(0, 3), (3, 6)], lambda: last['g'],
                   maxiters=30)
    assert np.abs(r['params'][act] - x.numpy()).max() < 1e-12


@pytest.mark.parametrize('case', ['l2', 'camconf'])
def test_gram_two_loop_is_the_same_recursion(hs64, case):
    """The coefficient-space ("Gram") two-loop (sfx_core.cuh gram_two_loop) is the recursion of
    lbfgs_ls.py:336-358 in exact arithmetic: in float64 a stage follows the default recursion
    loss for loss until round-off (chaos) separates them, and ends at the same minimum."""
    import copy
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, case)
    runs = []
    for mode in ('exact', 'gram'):
        st = copy.copy(I['stage'])
        st.generic_two_loop = N.two_loop_mode(mode)
        hs64.trace()
        r = hs64.fit(st, I['x'], I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                     I['init_mask'], I['reg_pose'])
        runs.append((r, hs64.trace().copy()))
    (r0, t0), (r1, t1) = runs
    n = min(len(t0), len(t1), 30)
    assert n >= 30
    assert np.abs(t0[:n] - t1[:n]).max() <= 1e-9 * np.abs(t0[:n]).max()
    assert abs(r0['loss'] - r1['loss']) <= 1e-8 * abs(r0['loss'])
    assert abs(r0['n_evals'] - r1['n_evals']) <= 10


def test_gram_two_loop_short_history_wraps(hs64):
    """history = 6: the ring of (s, y) pairs wraps after a few iterations, so the per-slot Gram
    blocks are overwritten in place (pop oldest / store newest) many times over a stage."""
    import copy
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, 'l2')
    runs = []
    for mode in ('exact', 'gram'):
        st = copy.copy(I['stage'])
        st.history = 6
        st.generic_two_loop = N.two_loop_mode(mode)
        hs64.trace()
        r = hs64.fit(st, I['x'], I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                     I['init_mask'], I['reg_pose'])
        runs.append((r, hs64.trace().copy()))
    (r0, t0), (r1, t1) = runs
    n = min(len(t0), len(t1), 40)
    assert n >= 40                      # well past the first wrap
    assert np.abs(t0[:n] - t1[:n]).max() <= 1e-9 * np.abs(t0[:n]).max()


@pytest.mark.parametrize('k,cond', [(0, 1e2), (1, 1e2), (5, 1e5), (30, 1e2), (100, 1e5)])
def test_gram_direction_accuracy(k, cond):
    """One L-BFGS direction from a given history (quadratic model with condition number
    ``cond``, steps shrinking towards the newest pair): the Gram two-loop equals the reference's
    recursion to round-off in float64, and in float32 its distance from the float64 direction
    is of the order of the exact recursion's own float32 error."""
    from tests.hostsim import hostsim as HS
    rng = np.random.default_rng(k)
    D = 119
    Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    A = (Q * np.geomspace(1, cond, D)) @ Q.T
    S = rng.standard_normal((k, D)) * np.geomspace(1, 1e-3, max(k, 1))[::-1][:k, None]
    Y = S @ A
    Y = Y + 1e-3 * rng.standard_normal((k, D)) * np.linalg.norm(Y, axis=1, keepdims=True) / np.sqrt(D)
    g = rng.standard_normal(D) * 1e-2
    hd = (S[-1] @ Y[-1]) / (Y[-1] @ Y[-1]) if k else 1.0
    ref = HS.lbfgs_direction('exact', S, Y, g, hd, True)
    nrm = np.linalg.norm(ref)
    assert np.linalg.norm(HS.lbfgs_direction('gram', S, Y, g, hd, True) - ref) <= 1e-12 * nrm
    e32 = np.linalg.norm(HS.lbfgs_direction('exact', S, Y, g, hd, False) - ref) / nrm
    g32 = np.linalg.norm(HS.lbfgs_direction('gram', S, Y, g, hd, False) - ref) / nrm
    assert e32 < 2e-6 and g32 < 2e-6
    assert g32 <= 5 * e32 + 2e-7
