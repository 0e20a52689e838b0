"""Interpenetration term without a GPU: known-answer tests of the restated mesh_intersection
package (oracle/isect_port.py -- third party, un-vendored, parity unpinned: SURVEY.md 8c) and
the kernel source compiled for the host against the reference's SMPLifyLoss driving it
(tests/golden/ref_eval_coll_*.npz)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import isect_port as IP
from tests import common as Cm
from tests.hostsim import hostsim
from tests.hostsim.hostsim import HostSim
from smplifyx_b200 import _native as N


def _tri(*pts):
    return np.array(pts, dtype=np.float64)


# ------------------------------------------------------------------ oracle known answers
def test_sat_known_answers():
    a = _tri([0, 0, 0], [1, 0, 0], [0, 1, 0])
    pierce = _tri([0.2, 0.2, -1], [0.2, 0.2, 1], [0.9, 0.9, 1])         # crosses the plane inside a
    above = a + np.array([0, 0, 0.5])                                    # parallel, separated
    far = _tri([0.2, 0.2, 1], [0.2, 0.2, 2], [0.9, 0.9, 2])              # same line, never reaches a
    beside = _tri([2, 2, -1], [2, 2, 1], [3, 3, 1])                      # crosses the plane outside a
    coplanar_in = _tri([0.1, 0.1, 0], [0.3, 0.1, 0], [0.1, 0.3, 0])      # inside a, same plane
    coplanar_out = _tri([2, 2, 0], [3, 2, 0], [2, 3, 0])
    got = IP.triangles_intersect(np.stack([a] * 6),
                                 np.stack([pierce, above, far, beside, coplanar_in, coplanar_out]))
    assert got.tolist() == [True, False, False, False, True, False]


def test_search_finds_exactly_the_crossing_pairs():
    rng = np.random.default_rng(0)
    # a grid of disjoint small triangles + three long needles that pierce some of them
    tris = []
    for i in range(12):
        for j in range(12):
            o = np.array([i, j, 0.0])
            tris.append(o + _tri([0.1, 0.1, 0], [0.9, 0.1, 0], [0.1, 0.9, 0]))
    needles = [(3, 4), (7, 7), (10, 1)]
    for (i, j) in needles:
        tris.append(_tri([i + 0.3, j + 0.3, -1], [i + 0.3, j + 0.3, 1], [i + 0.35, j + 0.3, 1]))
    tris = np.stack(tris)
    pairs = IP.find_collisions(tris)
    want = sorted((i * 12 + j, 144 + n) for n, (i, j) in enumerate(needles))
    assert [tuple(p) for p in pairs.tolist()] == want
    # brute force agrees on a random soup, and shared corners are never reported
    soup = rng.normal(size=(60, 3, 3))
    soup[1, 0] = soup[0, 0]
    got = set(map(tuple, IP.find_collisions(soup).tolist()))
    ii, jj = np.triu_indices(60, 1)
    hit = IP.triangles_intersect(soup[ii], soup[jj]) & ~IP.share_vertex(soup[ii], soup[jj])
    assert got == set(zip(ii[hit].tolist(), jj[hit].tolist())) and (0, 1) not in got
    out = IP.BVH(max_collisions=4)(torch.tensor(soup[None]))
    assert out.shape == (1, 240, 2) and int((out[0, :, 0] >= 0).sum()) == min(len(got), 240)


def test_filter_faces_rules():
    segm = np.array([0, 0, 1, 2, 3, 4])
    parents = np.array([-1, -1, 0, 1, 0, 3])
    ff = IP.FilterFaces(segm, parents, ign_part_pairs=['2,3'])
    c = torch.tensor([[[0, 1], [0, 2], [0, 3], [2, 3], [3, 4], [3, 5], [2, 4], [-1, -1]]])
    keep = (ff(c)[0, :, 0] >= 0).tolist()
    #        same   parent grand  parent ignored free   sibling pad
    assert keep == [False, False, True, False, False, True, True, False]
    ok = IP.allowed_part_matrix(segm, parents, ['2,3'])
    assert ok[0, 2] and not ok[0, 1] and not ok[2, 3] and not ok[3, 2] and ok[2, 4] and not ok[4, 4]


def test_distance_field_known_values():
    sigma = 0.5
    tri = torch.tensor([[[1.0, 0, 0], [-0.5, 3 ** 0.5 / 2, 0], [-0.5, -3 ** 0.5 / 2, 0]]],
                       dtype=torch.float64)                     # unit circumcircle, normal +z
    o, r, n = IP.circumcircle(tri)
    assert torch.allclose(o, torch.zeros(1, 3, dtype=torch.float64), atol=1e-15)
    assert abs(float(r) - 1) < 1e-15 and torch.allclose(n, torch.tensor([[0.0, 0, 1]], dtype=torch.float64))
    pts = torch.tensor([[[0, 0, -1.0], [0.5, 0, 0.0], [0, 0, 0.6]]], dtype=torch.float64)
    psi = IP.cone_field(pts, o, r, n, sigma)[0]
    # on the axis one unit inside: Phi = 0, Upsilon = 1 + 1 - 0.5
    assert abs(float(psi[0]) - 1.5) < 1e-15
    # in the plane half a radius out: Phi = 0.5, Upsilon(0) = (3 - 2 sigma) / 4 = 0.5
    assert abs(float(psi[1]) - 0.25) < 1e-15
    # above the cone's apex: nothing
    assert float(psi[2]) == 0.0
    # Upsilon is continuous at +-sigma
    x = torch.tensor([-sigma - 1e-9, -sigma + 1e-9, sigma - 1e-9, sigma + 1e-9], dtype=torch.float64)
    u = IP.upsilon(x, sigma)
    assert abs(float(u[0] - u[1])) < 1e-8 and abs(float(u[2] - u[3])) < 1e-8 and abs(float(u[0]) - 1) < 1e-8
    # loss of one pair = sum of psi^4 in both directions
    other = tri + torch.tensor([0.0, 0, -0.2], dtype=torch.float64)
    both = torch.stack([tri[0], other[0]])[None]
    L = IP.DistanceFieldPenetrationLoss(sigma=sigma)(both, torch.tensor([[[0, 1], [-1, -1]]]))
    o2, r2, n2 = IP.circumcircle(other)
    want = (IP.cone_field(other, o, r, n, sigma) ** 4).sum() + (IP.cone_field(tri, o2, r2, n2, sigma) ** 4).sum()
    assert abs(float(L[0]) - float(want)) < 1e-15 and float(L[0]) > 0


# ------------------------------------------------------------------ host build of the kernels
def _lib():
    lib = C.CDLL(hostsim.build())
    lib.hs_pair_terms.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    lib.hs_triangles_intersect.argtypes = [C.c_void_p, C.c_void_p]
    return lib


def test_pair_penalty_and_adjoint_against_autograd():
    lib = _lib()
    ev = Cm.golden('ref_eval_coll_f64.npz')
    faces = np.asarray(Cm.model_data()['f']).astype(np.int64)
    tri = ev['coll/vertices'][faces]
    pairs = ev['coll/pairs']
    sigma = float(ev['sigma'])
    sel = pairs[np.linspace(0, len(pairs) - 1, 150).astype(int)]
    for a, b in sel:
        ta = torch.tensor(tri[a], requires_grad=True)
        tb = torch.tensor(tri[b], requires_grad=True)
        o, r, n = IP.circumcircle(ta[None])
        o2, r2, n2 = IP.circumcircle(tb[None])
        l1 = (IP.cone_field(tb[None], o, r, n, sigma) ** 4).sum()
        l2 = (IP.cone_field(ta[None], o2, r2, n2, sigma) ** 4).sum()
        (l1 + l2).backward()
        loss, gi = np.zeros(1), np.zeros(9)
        ti, tj = np.ascontiguousarray(tri[a].reshape(-1)), np.ascontiguousarray(tri[b].reshape(-1))
        assert lib.hs_triangles_intersect(ti.ctypes.data, tj.ctypes.data) == 1
        lib.hs_pair_terms(ti.ctypes.data, tj.ctypes.data, sigma, loss.ctypes.data, gi.ctypes.data)
        gref = ta.grad.numpy().reshape(-1)
        assert abs(loss[0] - float(l1.detach())) <= 1e-12 * max(1.0, abs(float(l1.detach())))
        assert np.abs(gi - gref).max() <= 1e-9 * max(1e-12, np.abs(gref).max())


def test_host_sat_agrees_with_oracle_on_a_soup():
    lib = _lib()
    rng = np.random.default_rng(5)
    soup = rng.normal(size=(80, 3, 3))
    ii, jj = np.triu_indices(80, 1)
    want = IP.triangles_intersect(soup[ii], soup[jj])
    got = [lib.hs_triangles_intersect(np.ascontiguousarray(soup[i]).ctypes.data,
                                      np.ascontiguousarray(soup[j]).ctypes.data)
           for i, j in zip(ii, jj)]
    assert np.array_equal(np.array(got, bool), want) and 50 < want.sum() < len(want)


def _hs(dt):
    hs = HostSim(Cm.model_data(), Cm.joint_map(), use_double=(dt == 'f64'), **Cm.MODEL_KW)
    segm, par, ign = Cm.coll_segmentation()
    hs.set_collision(segm, par, [[int(x) for x in p.split(',')] for p in ign])
    return hs


@pytest.mark.parametrize('dt,tol_loss,tol_g', [('f64', 1e-13, 1e-10), ('f32', 1e-5, 2e-3)])
def test_eval_with_interpenetration_hostsim(dt, tol_loss, tol_g):
    ev = Cm.golden('ref_eval_coll_{}.npz'.format(dt))
    hs = _hs(dt)
    L = Cm.layout()
    res = {}
    for case in ('nocoll', 'coll'):
        I = Cm.coll_case_inputs(ev, case)
        r = hs.eval(I['stage'], I['x'], I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                    I['init_mask'], None)
        ref = float(ev[case + '/loss'])
        assert abs(r['loss'] - ref) <= tol_loss * abs(ref)
        assert r['flags'] == 0
        res[case] = r
    assert hs.last_touched() > 500
    g = res['coll']['grad'] - res['nocoll']['grad']
    g_ref = Cm.golden_grad_vector(L, ev, 'coll') - Cm.golden_grad_vector(L, ev, 'nocoll')
    assert np.abs(g_ref).max() > 1.0
    assert np.abs(g - g_ref).max() <= tol_g * np.abs(g_ref).max() + \
        (0 if dt == 'f64' else 1e-6 * np.abs(res['coll']['grad']).max())


@pytest.mark.parametrize('dt,tol_loss,tol_g', [('f64', 1e-13, 1e-10), ('f32', 1e-5, 2e-3)])
def test_eval_without_the_face_filter_hostsim(dt, tol_loss, tol_g):
    """The reference's path without part_segm_fn (fit_single_frame.py:317-328: filter_faces =
    None): 109 412 pairs instead of 8 343; the faces are only grouped for the broad phase."""
    ev = Cm.golden('ref_eval_collnf_{}.npz'.format(dt))
    md = Cm.synthetic.without_degenerate_faces(Cm.model_data())
    hs = HostSim(md, Cm.joint_map(), use_double=(dt == 'f64'), **Cm.MODEL_KW)
    faces = np.asarray(md['f']).astype(np.int64)
    hs.set_collision(np.asarray(md['weights']).argmax(1)[faces[:, 0]] % 64, None, [])
    L = Cm.layout()
    res = {}
    for case in ('nocoll', 'collnf'):
        I = Cm.coll_case_inputs(ev, case)
        r = hs.eval(I['stage'], I['x'], I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                    I['init_mask'], None)
        ref = float(ev[case + '/loss'])
        assert abs(r['loss'] - ref) <= tol_loss * abs(ref) and r['flags'] == 0
        res[case] = r
    assert hs.last_touched() > 9000 and len(ev['collnf/pairs']) > 100000
    g = res['collnf']['grad'] - res['nocoll']['grad']
    g_ref = Cm.golden_grad_vector(L, ev, 'collnf') - Cm.golden_grad_vector(L, ev, 'nocoll')
    assert np.abs(g_ref).max() > 1000.0
    assert np.abs(g - g_ref).max() <= tol_g * np.abs(g_ref).max() + \
        (0 if dt == 'f64' else 1e-6 * np.abs(res['collnf']['grad']).max())


def test_candidate_overflow_falls_back_to_the_global_area():
    """With a tiny shared work area the candidates spill to the block's global arrays and the
    result does not change."""
    ev = Cm.golden('ref_eval_coll_f64.npz')
    I = Cm.coll_case_inputs(ev, 'coll')
    out = []
    for wb in (1 << 20, 48000):
        hs = HostSim(Cm.model_data(), Cm.joint_map(), use_double=True, **Cm.MODEL_KW)
        segm, par, ign = Cm.coll_segmentation()
        hs.set_collision(segm, par, [[int(x) for x in p.split(',')] for p in ign], work_bytes=wb)
        r = hs.eval(I['stage'], I['x'], I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                    I['init_mask'], None)
        assert r['flags'] == 0
        out.append(r)
    assert out[0]['loss'] == out[1]['loss'] and np.array_equal(out[0]['grad'], out[1]['grad'])


def test_tiny_hit_regions_take_the_inline_path():
    """With a hit region of 16 entries most candidates overflow it and are walked again partner
    by partner in the narrow phase: same result, no overflow flag."""
    ev = Cm.golden('ref_eval_coll_f64.npz')
    I = Cm.coll_case_inputs(ev, 'coll')
    hs = _hs('f64')
    run = lambda: hs.eval(I['stage'], I['x'], I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                          I['init_mask'], None)
    a = run()
    hs.lib.hs_set_hits_cap(16)
    try:
        b = run()
    finally:
        hs.lib.hs_set_hits_cap(0)
    assert a['flags'] == 0 and b['flags'] == 0
    assert a['loss'] == b['loss'] and np.array_equal(a['grad'], b['grad'])


def test_gradient_of_the_term_against_finite_differences():
    ev = Cm.golden('ref_eval_coll_f64.npz')
    hs = _hs('f64')
    I = Cm.coll_case_inputs(ev, 'coll')
    I['stage'].coll_loss_weight = 1000.0            # make the term dominate rounding
    f = lambda x: hs.eval(I['stage'], x, I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                          I['init_mask'], None)
    r0 = f(I['x'])
    L = I['L']
    idx = [L.off_pose + 15 * 3 + 2, L.off_pose + 17 * 3 + 1, L.off_betas, L.off_go + 1]
    for i in idx:
        h = 1e-7
        xp, xm = I['x'].copy(), I['x'].copy()
        xp[i] += h
        xm[i] -= h
        fd = (f(xp)['loss'] - f(xm)['loss']) / (2 * h)
        # the set of colliding pairs may change between the two probes (the term is only
        # piecewise smooth), hence the loose bound
        assert abs(fd - r0['grad'][i]) <= 2e-3 * max(1.0, abs(fd)), (i, fd, r0['grad'][i])
