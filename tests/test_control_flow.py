"""Control-flow parity of the on-device optimiser (csrc/sfx_core.cuh: run_fitting, lbfgs_step,
strong_wolfe) against the oracle port of the reference optimiser (oracle/fit_port.py, itself
pinned against the unmodified reference in test_oracle.py).

The fitting problem is chaotic: a 1-ulp change of one dot product changes the number of closure
evaluations of a stage (see DESIGN.md "parity").  To separate control flow from round-off the
oracle optimiser is replayed here with (a) the engine's own closure (host build of the kernel
source) and (b) the engine's summation order, in float64.  With identical arithmetic the two
must take identical decisions: same evaluation count, same loss at every evaluation, same final
parameters, bit for bit.
"""
import math

import numpy as np
import pytest
import torch

from tests import common as Cm
from tests.hostsim.hostsim import HostSim
from smplifyx_b200 import _native as N
from oracle import fit_port as FP


class LaneOps(object):
    """Reductions in the engine's fixed order (sfx_core.cuh block_reduce): lane l sums the
    elements l, l+32, ... sequentially, then an xor butterfly over the 32 partials; elementwise
    updates without fused multiply-add."""

    @staticmethod
    def _reduce(v):
        p = np.zeros(32)
        for l in range(32):
            acc = 0.0
            for x in v[l::32]:
                acc = acc + x
            p[l] = acc
        idx = np.arange(32)
        for o in (16, 8, 4, 2, 1):
            p = p + p[idx ^ o]
        return p[0]

    @staticmethod
    def dot(a, b):
        return torch.tensor(LaneOps._reduce(a.detach().numpy() * b.detach().numpy()),
                            dtype=torch.float64)

    @staticmethod
    def absmax(a):
        return a.abs().max()

    @staticmethod
    def abssum(a):
        return torch.tensor(LaneOps._reduce(np.abs(a.detach().numpy())), dtype=torch.float64)

    @staticmethod
    def axpy_(x, alpha, y):
        return x.add_(y.mul(float(alpha)))


def _float_cubic(monkeypatch):
    """torch evaluates ``python_float / tensor`` as ``tensor.reciprocal() * float`` (two
    roundings), which the reference's cubic interpolation hits whenever a bracket end is a
    0-dim tensor (lbfgs_ls.py:22).  The engine divides once, in double.  The replay removes that
    1-ulp artefact by evaluating the oracle's interpolation on python floats."""
    def cubic(x1, f1, g1, x2, f2, g2, bounds=None):
        x1, f1, g1, x2, f2, g2 = [float(v) for v in (x1, f1, g1, x2, f2, g2)]
        lo, hi = ([float(b) for b in bounds] if bounds is not None
                  else ((x1, x2) if x1 <= x2 else (x2, x1)))
        d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)
        disc = d1 ** 2 - g1 * g2
        if disc >= 0:
            d2 = math.sqrt(disc)
            if x1 <= x2:
                pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))
            else:
                pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2 * d2))
            return min(max(pos, lo), hi)
        return (lo + hi) / 2.
    monkeypatch.setattr(FP, '_cubic_min', cubic)


def _replay(hs, I, blocks_names, maxiters=30):
    L, st = I['L'], I['stage']
    pb = N.param_blocks(L)
    act = np.concatenate([np.arange(pb[n][0], pb[n][0] + pb[n][1]) for n in blocks_names])
    x_full = I['x'].copy()
    x = torch.tensor(x_full[act], dtype=torch.float64)
    trace = []
    last = {}

    def closure():
        xf = x_full.copy()
        xf[act] = x.numpy()
        r = hs.eval(st, xf, I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
                    I['init_mask'], I['reg_pose'])
        g = torch.tensor(r['grad'][act], dtype=torch.float64)
        trace.append(r['loss'])
        last['g'] = g
        return torch.tensor(r['loss'], dtype=torch.float64), g

    opt = FP.StrongWolfeLBFGS(x, lr=st.lr, max_iter=st.max_iter, max_eval=st.max_eval,
                              tol_grad=st.tol_grad, tol_change=st.tol_change,
                              history=st.history, ops=LaneOps)
    blocks = []
    pos = 0
    for n in blocks_names:
        blocks.append((pos, pos + pb[n][1]))
        pos += pb[n][1]
    final = FP.run_fitting(opt, closure, blocks, lambda: last['g'], maxiters=maxiters,
                           ftol=st.ftol, gtol=st.gtol)
    xf = x_full.copy()
    xf[act] = x.numpy()
    return final, np.array(trace), xf


@pytest.fixture(scope='module')
def hs():
    return HostSim(Cm.model_data(), Cm.joint_map(), use_double=True, **Cm.MODEL_KW)


@pytest.mark.parametrize('case,blocks', [('reg', N.BODY_STAGE_BLOCKS),
                                         ('l2', N.BODY_STAGE_BLOCKS),
                                         ('camconf', N.CAMERA_STAGE_BLOCKS)])
def test_stage_trajectory_bit_identical(hs, case, blocks, monkeypatch):
    _float_cubic(monkeypatch)
    ev = Cm.golden('ref_eval_f64.npz')
    I = Cm.eval_case_inputs(ev, case)
    final_o, trace_o, x_o = _replay(hs, I, blocks)
    hs.trace()
    r = hs.fit(I['stage'], I['x'], I['gt'], I['conf'], I['jw'], I['cam'], I['lowconf'],
               I['init_mask'], I['reg_pose'])
    trace_e = hs.trace()
    assert r['n_evals'] == len(trace_o)
    assert np.array_equal(trace_e, trace_o)
    assert r['loss'] == final_o
    assert np.array_equal(r['params'], x_o)
    assert len(trace_o) > 25        # a real multi-iteration run, not an early exit
