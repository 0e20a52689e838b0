"""GPU parity tests proper: the CUDA path, called through the C ABI, against the committed
golden vectors of the unmodified reference (tests/golden/, made by make_golden.py) and against
the host build of the same kernel source.

Tolerances (north_star: 1e-3 relative on float32 outputs):
* one closure evaluation, float64 build: 1e-9 relative on the loss and on every gradient entry
  (scaled by the largest entry) -- pure round-off;
* float32 build: 1e-5 on the loss, 5e-4 of the largest gradient entry on the gradient;
* a fitted stage: the trajectory is chaotic (DESIGN.md "Parity"), so the comparison is against
  the envelope of the reference's own round-off-perturbed runs.
"""
import json

import numpy as np
import pytest
import torch

from tests import common as Cm
from smplifyx_b200 import _native as N

pytestmark = pytest.mark.gpu


def _engine():
    from smplifyx_b200 import engine
    return engine


@pytest.fixture(scope='module')
def model32():
    return _engine().Model(Cm.model_data(), Cm.joint_map(), dtype=torch.float32, **Cm.MODEL_KW)


@pytest.fixture(scope='module')
def model64():
    return _engine().Model(Cm.model_data(), Cm.joint_map(), dtype=torch.float64, **Cm.MODEL_KW)


def _load(batch, I, B):
    rep = lambda a: np.repeat(np.asarray(a)[None], B, axis=0)
    kp = np.concatenate([I['gt'], I['conf'][:, None]], axis=1)
    batch.set_targets(rep(kp), rep(I['jw']), rep(I['lowconf']), rep(I['init_mask']),
                      rep(I['cam']), rep(I['reg_pose']))
    batch.set_params(rep(I['x']))


@pytest.mark.parametrize('case', ['l2', 'reg', 'cam', 'camconf'])
def test_eval_matches_reference_f64(model64, case):
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, case)
    batch = _engine().FrameBatch(model64, 3)
    _load(batch, I, 3)
    loss, grad, joints = batch.eval(I['stage'], want_joints=True)
    ref = float(ev[case + '/loss'])
    assert np.allclose(loss.cpu().numpy(), ref, rtol=1e-9, atol=0)
    g_ref = Cm.golden_grad_vector(I['L'], ev, case)
    g = grad.cpu().numpy()
    live = g_ref != 0
    scale = np.abs(g_ref).max()
    for b in range(3):
        assert np.abs(g[b][live] - g_ref[live]).max() <= 1e-9 * scale
    if case == 'l2':
        assert np.abs(joints.cpu().numpy()[1] - ev['l2/joints']).max() < 1e-12


@pytest.mark.parametrize('case', ['l2', 'reg', 'cam', 'camconf'])
def test_eval_matches_reference_f32(model32, case):
    ev = Cm.golden('ref_eval_f32.npz')
    I = Cm.eval_case_inputs(ev, case)
    batch = _engine().FrameBatch(model32, 5)
    _load(batch, I, 5)
    loss, grad, joints = batch.eval(I['stage'], want_joints=True)
    ref = float(ev[case + '/loss'])
    assert np.allclose(loss.cpu().numpy(), ref, rtol=1e-5, atol=0)
    g_ref = Cm.golden_grad_vector(I['L'], ev, case)
    g = grad.cpu().numpy()
    live = g_ref != 0
    scale = np.abs(g_ref).max()
    for b in range(5):
        assert np.abs(g[b][live] - g_ref[live]).max() <= 5e-4 * scale
    if case == 'l2':
        assert np.abs(joints.cpu().numpy()[0] - ev['l2/joints']).max() < 2e-5


def test_mixture_prior_matches_reference(model32, model64):
    ev64, ev32 = Cm.golden('ref_eval_f64.npz'), Cm.golden('ref_eval_f32.npz')
    for model, ev, dt, tl, tg in ((model64, ev64, torch.float64, 1e-10, 1e-9),
                                  (model32, ev32, torch.float32, 1e-5, 5e-4)):
        model.set_gmm(Cm.gmm_prior(dt))
        I = Cm.eval_case_inputs(ev, 'gmm')
        batch = _engine().FrameBatch(model, 2)
        _load(batch, I, 2)
        loss, grad, _ = batch.eval(I['stage'])
        ref = float(ev['gmm/loss'])
        assert np.allclose(loss.cpu().numpy(), ref, rtol=tl, atol=0)
        g_ref = Cm.golden_grad_vector(I['L'], ev, 'gmm')
        assert np.abs(grad.cpu().numpy()[1] - g_ref).max() <= tg * np.abs(g_ref).max()
        final = batch.fit_stage(I['stage']).cpu().numpy()
        assert np.all(np.isfinite(final)) and np.all(final < 0.8 * ref)


@pytest.mark.parametrize('case', ['vposer', 'vposer_reg'])
def test_vposer_latent_pose_matches_reference(case):
    """VPoser decoder + adjoint inside the evaluation kernel (fitting.py:235-236, :389-395)."""
    for dt, evn, tl, tg in ((torch.float64, 'ref_eval_f64.npz', 1e-10, 1e-8),
                            (torch.float32, 'ref_eval_f32.npz', 1e-5, 1e-3)):
        ev = Cm.golden(evn)
        model = _engine().Model(Cm.model_data(), Cm.joint_map(), dtype=dt, **Cm.MODEL_KW)
        model.set_vposer(Cm.vposer_weights())
        I = Cm.eval_case_inputs(ev, case)
        batch = _engine().FrameBatch(model, 3, use_vposer=True)
        assert batch.L.n_pose == 32
        B = 3
        rep = lambda a: None if a is None else np.repeat(np.asarray(a)[None], B, axis=0)
        kp = np.concatenate([I['gt'], I['conf'][:, None]], axis=1)
        batch.set_targets(rep(kp), rep(I['jw']), rep(I['lowconf']), rep(I['init_mask']),
                          rep(I['cam']), rep(I['reg_pose']))
        batch.set_params(rep(I['x']))
        loss, grad, joints = batch.eval(I['stage'], want_joints=True)
        ref = float(ev[case + '/loss'])
        assert np.allclose(loss.cpu().numpy(), ref, rtol=tl, atol=0)
        g_ref = Cm.golden_grad_vector(I['L'], ev, case)
        assert np.abs(grad.cpu().numpy()[2] - g_ref).max() <= tg * np.abs(g_ref).max()
        assert np.abs(joints.cpu().numpy()[0] - ev[case + '/joints']).max() < (1e-11 if dt == torch.float64 else 3e-5)
        final = batch.fit_stage(I['stage']).cpu().numpy()
        assert np.all(np.isfinite(final)) and np.all(final < 0.8 * ref)
        verts, _ = batch.forward_mesh()
        assert bool(torch.isfinite(verts).all())


def test_ring_and_direct_streams_agree(model32, monkeypatch):
    """The TMA ring and the plain-load path of the blend passes give the same bits."""
    ev = Cm.golden('ref_eval_f32.npz')
    I = Cm.eval_case_inputs(ev, 'reg')
    batch = _engine().FrameBatch(model32, 2)
    _load(batch, I, 2)
    l0, g0, _ = batch.eval(I['stage'])
    monkeypatch.setenv('SFX_STREAM_DIRECT', '1')
    l1, g1, _ = batch.eval(I['stage'])
    assert torch.equal(l0, l1) and torch.equal(g0, g1)
    # the register-pipelined variant (no ring) and the shuffle-sum two-loop are A/B switches of
    # the same arithmetic
    monkeypatch.delenv('SFX_STREAM_DIRECT')
    monkeypatch.setenv('SFX_STREAM_REGS', '1')
    l2, g2, _ = batch.eval(I['stage'])
    assert torch.equal(l0, l2) and torch.equal(g0, g2)
    monkeypatch.delenv('SFX_STREAM_REGS')
    ref = batch.fit_stage(I['stage']).clone()
    x_ref = batch.get_params()
    monkeypatch.setenv('SFX_TWO_LOOP_SHFL', '1')
    _load(batch, I, 2)
    again = batch.fit_stage(I['stage'])
    assert torch.equal(ref, again) and np.array_equal(x_ref, batch.get_params())


def test_two_loop_variants_agree(model32, monkeypatch):
    """Shared-address, generic-pointer (both TMA-staged) and block-wide two-loop recursion: a
    whole stage ends with the same bits and the same evaluation count."""
    ev = Cm.golden('ref_eval_f32.npz')
    I = Cm.eval_case_inputs(ev, 'reg')
    out = []
    for env, generic in ((None, 0), ('SFX_TWO_LOOP_GENERIC', 0), (None, 1)):
        if env:
            monkeypatch.setenv(env, '1')
        batch = _engine().FrameBatch(model32, 2)
        _load(batch, I, 2)
        I['stage'].generic_two_loop = generic
        final = batch.fit_stage(I['stage'])
        out.append((batch.get_params(), final.cpu().numpy(), batch.evals().cpu().numpy()))
        if env:
            monkeypatch.delenv(env)
    I['stage'].generic_two_loop = 0
    for o in out[1:]:
        assert np.array_equal(out[0][0], o[0]) and np.array_equal(out[0][1], o[1])
        assert np.array_equal(out[0][2], o[2])
    assert out[0][2].min() > 60


def test_tensor_core_mesh_against_simt(model32, monkeypatch):
    """Fused tcgen05 / TMA mesh kernel (blend in tf32, skinning in three products of float16
    hi / lo halves, fp32 accumulation in tensor memory; csrc/sfx_mesh_fused.cuh) and the round-1 pair (tcgen05 blend + SIMT skinning)
    against the fp32 SIMT kernel on 130 frames (two frame tiles, ragged) with distinct
    parameters: the tf32 rounding of the blend inputs bounds the vertex error by ~1e-4 m;
    structure errors would be centimetres."""
    ev = Cm.golden('ref_eval_f32.npz')
    I = Cm.eval_case_inputs(ev, 'l2')
    B = 130
    rng = np.random.default_rng(11)
    x = np.repeat(I['x'][None], B, axis=0)
    x[:, :I['L'].off_camt] += rng.normal(size=(B, I['L'].off_camt)) * 0.2
    batch = _engine().FrameBatch(model32, B)
    _load(batch, I, B)
    batch.set_params(x)
    v_tc, j_tc = batch.forward_mesh()              # fused kernel: blend + skinning on tcgen05
    monkeypatch.setenv('SFX_MESH_UNFUSED', '1')
    v_un, _ = batch.forward_mesh()                 # tcgen05 blend kernel + SIMT skinning kernel
    monkeypatch.delenv('SFX_MESH_UNFUSED')
    monkeypatch.setenv('SFX_MESH_SIMT', '1')
    v_simt, j_simt = batch.forward_mesh()
    err = (v_tc - v_simt).abs().max().item()
    print('fused tensor-core mesh vs float32 SIMT: max %.3g m, mean %.3g m; unfused: max %.3g m'
          % (err, (v_tc - v_simt).abs().mean().item(), (v_un - v_simt).abs().max().item()))
    assert torch.isfinite(v_tc).all()
    assert err < 2e-4, err
    assert (v_tc - v_simt).abs().mean().item() < 2e-5
    assert (v_un - v_simt).abs().max().item() < 2e-4
    assert torch.equal(j_tc, j_simt)
    # frames are distinct: a tile / lane mix-up would not pass
    assert (v_simt[0] - v_simt[129]).abs().max().item() > 1e-2


def test_full_mesh_matches_reference(model32, model64):
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, 'l2')
    for model, tol in ((model64, 1e-11), (model32, 1e-4)):
        batch = _engine().FrameBatch(model, 3)
        _load(batch, I, 3)
        verts, joints = batch.forward_mesh()
        v = verts.cpu().numpy()
        assert np.abs(v[0] - ev['vertices']).max() < tol
        assert np.abs(v[2] - ev['vertices']).max() < tol
        assert np.abs(joints.cpu().numpy()[1] - ev['l2/joints']).max() < tol


def test_stage_fit_f64_matches_host_build(model64):
    """Same source, device build vs single-threaded host build, float64.  libm differences
    (sin/cos/exp are not bit-identical between CUDA and glibc) make bit equality impossible;
    the early trajectory must agree to round-off and the run must end inside the reference's
    own perturbation envelope (tests/golden/ref_stage_f64.npz and DESIGN.md)."""
    from tests.hostsim.hostsim import HostSim
    ev = Cm.golden('ref_eval_f64.npz')
    sg = Cm.golden('ref_stage_f64.npz')
    I = Cm.eval_case_inputs(ev, 'reg')
    batch = _engine().FrameBatch(model64, 2)
    _load(batch, I, 2)
    final = batch.fit_stage(I['stage']).cpu().numpy()
    n_ev = batch.evals().cpu().numpy()
    ref_final = float(sg['final_loss'])
    # envelope of 1-ulp-perturbed reference runs measured in the authoring container:
    # final loss 486559 .. 486937, 165 .. 273 evaluations
    assert np.all(np.abs(final - ref_final) < 2e-3 * ref_final), (final, ref_final)
    assert np.all((n_ev > 80) & (n_ev < 600)), n_ev
    assert final[0] == final[1] and n_ev[0] == n_ev[1]        # frames are independent + deterministic
    hs = HostSim(Cm.model_data(), Cm.joint_map(), use_double=True, **Cm.MODEL_KW)
    r = hs.fit(I['stage'], I['x'], I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
               I['init_mask'], I['reg_pose'])
    assert abs(r['loss'] - final[0]) < 2e-3 * ref_final


def test_stage_fit_f32_envelope(model32):
    """A float32 stage from a far-away random start: the reference's own float32 runs started
    one ulp apart end with final losses 403k .. 528k and meshes up to 39 cm apart
    (tests/golden/ref_envelope.npz).  The engine must land inside that envelope."""
    ev = Cm.golden('ref_eval_f32.npz')
    env = Cm.golden('ref_envelope.npz')
    I = Cm.eval_case_inputs(ev, 'reg')
    batch = _engine().FrameBatch(model32, 4)
    _load(batch, I, 4)
    final = batch.fit_stage(I['stage']).cpu().numpy()
    lo, hi = env['stage/final_loss'].min(), env['stage/final_loss'].max()
    assert np.all((final > lo - 0.1 * (hi - lo)) & (final < hi + 0.1 * (hi - lo))), (final, lo, hi)
    assert int(batch.flags().cpu().numpy().max()) == 0
    assert np.all(final == final[0])
    verts, _ = batch.forward_mesh()
    err = np.abs(verts.cpu().numpy()[0] - env['stage/vertices_base'])
    assert err.max() <= float(env['stage/vertex_pairwise_max']), err.max()
    assert err.mean() <= float(env['stage/vertex_pairwise_mean']), err.mean()


def test_two_loop_fast_path_bit_identical(model32):
    """The single-warp, register-resident two-loop recursion gives the same bits as the
    block-wide one (same summation order by construction)."""
    import copy
    ev = Cm.golden('ref_eval_f32.npz')
    I = Cm.eval_case_inputs(ev, 'reg')
    outs = []
    for generic in (0, 1):
        st = copy.copy(I['stage'])
        st.generic_two_loop = generic
        batch = _engine().FrameBatch(model32, 2)
        _load(batch, I, 2)
        final = batch.fit_stage(st).cpu().numpy()
        outs.append((final, batch.get_params(), batch.evals().cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])


def test_frame_subset_launch(model32):
    """frame_ids restricts a stage to a subset; the other frames keep their parameters."""
    ev = Cm.golden('ref_eval_f32.npz')
    I = Cm.eval_case_inputs(ev, 'reg')
    batch = _engine().FrameBatch(model32, 4)
    _load(batch, I, 4)
    before = batch.get_params()
    ids = torch.tensor([1, 3], dtype=torch.int32, device='cuda')
    batch.fit_stage(I['stage'], frame_ids=ids)
    after = batch.get_params()
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[2], after[2])
    assert not np.array_equal(before[1], after[1])
    assert np.array_equal(after[1], after[3])


def test_adam_stage(model64):
    """optim_type adam (optim_factory.py:45-48): 30 steps of torch.optim.Adam semantics."""
    from oracle import fit_port as FP
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, 'cam')
    from smplifyx_b200 import _native as N
    st = N.make_stage(I['L'], N.CAMERA_STAGE_BLOCKS, loss_kind=N.LOSS_CAMERA_INIT,
                      opt_kind=N.OPT_ADAM, lr=1e-2, depth_loss_weight=100.0, maxiters=30)
    batch = _engine().FrameBatch(model64, 1)
    _load(batch, I, 1)
    batch.fit_stage(st)
    got = batch.get_params()[0]
    # oracle: torch.optim.Adam driven by the engine's own gradient at each iterate
    L = I['L']
    pb = N.param_blocks(L)
    act = np.concatenate([np.arange(pb[n][0], pb[n][0] + pb[n][1]) for n in N.CAMERA_STAGE_BLOCKS])
    x = torch.tensor(I['x'][act], dtype=torch.float64)
    b2 = _engine().FrameBatch(model64, 1)
    _load(b2, I, 1)
    full = I['x'].copy()
    last = {}

    def closure():
        full[act] = x.detach().numpy()
        b2.set_params(full[None])
        l, g, _ = b2.eval(st)
        last['g'] = torch.tensor(g.cpu().numpy()[0][act], dtype=torch.float64)
        return torch.tensor(float(l.cpu().numpy()[0]), dtype=torch.float64), last['g']
    opt = FP.AdamPort(x, lr=1e-2)
    FP.run_fitting(opt, closure, [(0, 3), (3, 6)], lambda: last['g'], maxiters=30)
    assert np.abs(got[act] - x.detach().numpy()).max() < 1e-9
    assert np.abs(got[act] - I['x'][act]).max() > 1e-2          # it moved


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['l2', 'camconf'])
def test_gram_two_loop_f64_matches_exact_recursion(model64, case):
    """Coefficient-space two-loop on the device (registers + broadcast chain, Gram blocks staged
    in shared memory or read in place) against the default recursion, float64: same minimum,
    same number of evaluations up to the chaotic tail."""
    import copy
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, case)
    outs = []
    for mode in ('exact', 'gram'):
        st = copy.copy(I['stage'])
        st.generic_two_loop = N.two_loop_mode(mode)
        batch = _engine().FrameBatch(model64, 2)
        _load(batch, I, 2)
        final = batch.fit_stage(st).cpu().numpy()
        outs.append((final, batch.get_params(), batch.evals().cpu().numpy()))
    assert np.abs(outs[0][0] - outs[1][0]).max() <= 1e-8 * np.abs(outs[0][0]).max()
    assert np.abs(outs[0][2] - outs[1][2]).max() <= 10
    assert np.array_equal(outs[1][1][0], outs[1][1][1])      # frames independent of position


@pytest.mark.gpu
def test_gram_two_loop_f32_short_history_follows_exact(model32):
    """float32, staged path: with a handful of iterations the two recursions still agree to
    round-off in the parameters they reach (before chaos amplifies the difference)."""
    import copy
    ev = Cm.golden('ref_eval_f32.npz')
    I = Cm.eval_case_inputs(ev, 'camconf')
    outs = []
    for mode in ('exact', 'gram'):
        st = copy.copy(I['stage'])
        st.generic_two_loop = N.two_loop_mode(mode)
        st.maxiters = 1
        st.max_iter = 6
        st.max_eval = 8
        batch = _engine().FrameBatch(model32, 2)
        _load(batch, I, 2)
        final = batch.fit_stage(st).cpu().numpy()
        outs.append((final, batch.get_params(), batch.evals().cpu().numpy()))
    assert np.abs(outs[0][0] - outs[1][0]).max() <= 1e-4 * np.abs(outs[0][0]).max()
    assert np.abs(outs[0][1] - outs[1][1]).max() <= 1e-3 * max(1.0, np.abs(outs[0][1]).max())


@pytest.mark.gpu
def test_gram_two_loop_f32_staging_variants_bit_identical(model32, monkeypatch):
    """float32: the zero-padded fixed-stride block + shared-address chain (TMA-ring build of the
    launch) and the tight block / in-place reads + generic chain (plain-load build) perform the
    same operations in the same order: a whole stage ends with the same bits."""
    import copy
    ev = Cm.golden('ref_eval_f32.npz')
    I = Cm.eval_case_inputs(ev, 'reg')
    st = copy.copy(I['stage'])
    st.generic_two_loop = N.two_loop_mode('gram')
    outs = []
    for direct in (False, True):
        if direct:
            monkeypatch.setenv('SFX_STREAM_DIRECT', '1')
        batch = _engine().FrameBatch(model32, 2)
        _load(batch, I, 2)
        final = batch.fit_stage(st).cpu().numpy()
        outs.append((final, batch.get_params(), batch.evals().cpu().numpy()))
        if direct:
            monkeypatch.delenv('SFX_STREAM_DIRECT')
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])
    assert outs[0][2].min() > 60


def test_float64_model_keeps_the_arrays_of_the_npz():
    """float_dtype: float64 (reference main.py:99-105): the model arrays reach the device as the
    npz holds them (the licensed files store float64), not through float32.  A template shifted by
    1e-9 m -- far below float32 resolution at body scale -- moves every vertex and every joint of
    the float64 model by 1e-9 m, and the evaluation (whose per-frame float32 copies of the
    skinning weights are bypassed for such a model) follows."""
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, 'l2')
    md = dict(Cm.model_data())
    base = _engine().Model(md, Cm.joint_map(), dtype=torch.float64, **Cm.MODEL_KW)
    md['v_template'] = np.asarray(md['v_template'], dtype=np.float64) + np.array([1e-9, 0.0, 0.0])
    moved = _engine().Model(md, Cm.joint_map(), dtype=torch.float64, **Cm.MODEL_KW)
    out = []
    for model in (base, moved):
        batch = _engine().FrameBatch(model, 2)
        _load(batch, I, 2)
        v, j = batch.forward_mesh()
        loss, grad, joints = batch.eval(I['stage'], want_joints=True)
        out.append((v.cpu().numpy(), joints.cpu().numpy(), loss.cpu().numpy()))
    dv = np.linalg.norm(out[1][0] - out[0][0], axis=-1)
    dj = np.linalg.norm(out[1][1] - out[0][1], axis=-1)
    assert np.all(np.abs(dv - 1e-9) < 1e-12), (dv.min(), dv.max())       # a rigid shift survives skinning
    assert 0.2e-9 < dj[:, 55:].mean() < 5e-9                              # vertex-derived joints move with it
    # the unshifted float64 model still reproduces the reference's float64 loss
    assert abs(out[0][2][0] - float(ev['l2/loss'])) <= 1e-9 * abs(float(ev['l2/loss']))
