"""TEST INFRASTRUCTURE: ctypes wrapper of the single-threaded host build of csrc/sfx_core.cuh
(tests/hostsim/sfx_hostsim.cpp).  Used to check the evaluation maths and the optimiser control
flow against the oracle without a GPU.  Not part of the product."""
import ctypes as C
import os
import subprocess

import numpy as np

from smplifyx_b200 import _native as N

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libsfx_hostsim.so')
_SRC = os.path.join(_HERE, 'sfx_hostsim.cpp')
_CORE = os.path.join(_HERE, '..', '..', 'smplify-x-partial_b200', 'csrc')


def build(force=False):
    deps = [_SRC] + [os.path.join(_CORE, f) for f in
                     ('sfx_core.cuh', 'sfx_collide.cuh', 'sfx_types.h', 'sfx_model_prep.h')]
    if (not force and os.path.isfile(_SO) and
            all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps)):
        return _SO
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-o', _SO, _SRC])
    return _SO


def lbfgs_direction(mode, s, y, g, h_diag, use_double):
    """L-BFGS direction from k pairs ``s, y [k, D]`` (oldest first) and the gradient ``g [D]``:
    ``mode`` 'exact' = the reference's recursion as lbfgs_step runs it, 'gram' = gram_two_loop
    with its Gram blocks built pair by pair (host build of csrc/sfx_core.cuh)."""
    lib = C.CDLL(build())
    dt = np.float64 if use_double else np.float32
    s = np.ascontiguousarray(s, dtype=dt)
    y = np.ascontiguousarray(y, dtype=dt)
    g = np.ascontiguousarray(g, dtype=dt)
    k, D = s.shape
    out = np.zeros(D, dtype=dt)
    lib.hs_lbfgs_direction.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_double, C.c_void_p]
    lib.hs_lbfgs_direction(int(use_double), N.two_loop_mode(mode), k, D, s.ctypes.data,
                           y.ctypes.data, g.ctypes.data, float(h_diag), out.ctypes.data)
    return out


class HostSim(object):
    def __init__(self, model_data, joint_map, use_double=True, use_vposer=False, **model_kw):
        self.lib = C.CDLL(build())
        self.lib.hs_model_create.restype = C.c_void_p
        self.lib.hs_model_create.argtypes = [C.POINTER(N.SfxModelDesc), C.c_char_p, C.c_int]
        self.lib.hs_model_destroy.argtypes = [C.c_void_p]
        self.lib.hs_layout.argtypes = [C.c_void_p, C.c_int, C.POINTER(N.SfxLayout)]
        self.lib.hs_run.argtypes = [C.c_void_p, C.POINTER(N.SfxStage), C.c_int, C.c_int] + \
            [C.c_void_p] * 11 + [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        desc, keep = N.build_model_desc(model_data, joint_map, use_double=use_double, **model_kw)
        err = C.create_string_buffer(256)
        self.h = self.lib.hs_model_create(C.byref(desc), err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.dt = np.float64 if use_double else np.float32
        self.use_vposer = use_vposer
        self.K = len(joint_map)
        self.L = N.SfxLayout()
        self.lib.hs_layout(self.h, int(use_vposer), C.byref(self.L))

    def set_vposer(self, w):
        a = [np.ascontiguousarray(w[k], dtype=self.dt) for k in
             ('dec_fc1_w', 'dec_fc1_b', 'dec_fc2_w', 'dec_fc2_b', 'dec_out_w', 'dec_out_b')]
        self._vp_keep = a
        self.lib.hs_set_vposer.argtypes = [C.c_void_p] * 7
        self.lib.hs_set_vposer(self.h, *[x.ctypes.data_as(C.c_void_p) for x in a])

    def set_gmm(self, means, precisions, log_nll_weights):
        a = [np.ascontiguousarray(x, dtype=self.dt) for x in (means, precisions, log_nll_weights)]
        self.lib.hs_set_gmm.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3
        self.lib.hs_set_gmm(self.h, a[0].shape[0], a[0].shape[1],
                            *[x.ctypes.data_as(C.c_void_p) for x in a])

    def set_collision(self, faces_segm, faces_parents, ign_part_pairs=(), work_bytes=None):
        segm = np.ascontiguousarray(faces_segm, dtype=np.int32)
        par = np.ascontiguousarray(faces_parents, dtype=np.int32) if faces_parents is not None else None
        ign = np.ascontiguousarray(np.asarray(ign_part_pairs, dtype=np.int32).reshape(-1, 2))
        if work_bytes is None:          # what the device has: the idle blend ring
            work_bytes = 131072 if self.dt == np.float32 else 65536
        err = C.create_string_buffer(256)
        self.lib.hs_set_collision.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_char_p, C.c_int]
        rc = self.lib.hs_set_collision(self.h, segm.ctypes.data_as(C.c_void_p),
                                       par.ctypes.data_as(C.c_void_p) if par is not None else None,
                                       ign.ctypes.data_as(C.c_void_p), ign.shape[0],
                                       int(work_bytes), err, 256)
        if rc:
            raise RuntimeError(err.value.decode())

    def last_touched(self):
        self.lib.hs_last_touch.argtypes = [C.c_void_p]
        return self.lib.hs_last_touch(self.h)

    def __del__(self):
        try:
            self.lib.hs_model_destroy(self.h)
        except Exception:
            pass

    def _run(self, stage, do_fit, params, gt, conf, jw, lowconf, init_mask, cam, reg_pose):
        a = lambda x, dt=None: np.ascontiguousarray(x, dtype=dt or self.dt)
        params = a(params).copy()
        gt, conf, jw, cam = a(gt), a(conf), a(jw), a(cam)
        lowconf = a(lowconf if lowconf is not None else np.zeros(self.K), np.uint8)
        init_mask = a(init_mask if init_mask is not None else np.zeros(self.K), np.uint8)
        reg = a(reg_pose) if reg_pose is not None else None
        loss = np.zeros(1, self.dt)
        grad = np.zeros(self.L.np, self.dt)
        joints = np.zeros((self.K, 3), self.dt)
        ne, fl = C.c_int(0), C.c_int(0)
        p = lambda x: x.ctypes.data_as(C.c_void_p) if x is not None else None
        self.lib.hs_run(self.h, C.byref(stage), int(self.use_vposer), int(do_fit), p(params), p(gt),
                        p(conf), p(jw), p(lowconf), p(init_mask), p(cam), p(reg), p(loss), p(grad),
                        p(joints), C.byref(ne), C.byref(fl))
        return dict(loss=float(loss[0]), grad=grad, joints=joints, params=params,
                    n_evals=ne.value, flags=fl.value)

    def trace(self, cap=100000):
        """Loss of every closure evaluation since the last call."""
        buf = np.zeros(cap)
        n = self.lib.hs_trace(buf.ctypes.data_as(C.c_void_p), cap)
        return buf[:min(n, cap)].copy()

    def steps(self, cap=100000):
        """Step length of every line-search probe since the last call."""
        buf = np.zeros(cap)
        n = self.lib.hs_steps(buf.ctypes.data_as(C.c_void_p), cap)
        return buf[:min(n, cap)].copy()

    def eval(self, stage, params, gt, conf, jw, cam, lowconf=None, init_mask=None, reg_pose=None):
        return self._run(stage, 0, params, gt, conf, jw, lowconf, init_mask, cam, reg_pose)

    def fit(self, stage, params, gt, conf, jw, cam, lowconf=None, init_mask=None, reg_pose=None):
        return self._run(stage, 1, params, gt, conf, jw, lowconf, init_mask, cam, reg_pose)
