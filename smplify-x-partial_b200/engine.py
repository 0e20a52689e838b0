"""Thin host-side owner of the C-ABI handles (``include/sfx.h``): one ``Model`` per SMPL-X
file and device, one ``FrameBatch`` per set of B frames fitted in lock-step.

torch is used for device memory and streams only; every computation is a ``libsfx.so`` call.
There is no CPU path: constructing a ``Model`` without a CUDA device raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Model(object):
    """smplx.create(...) + JointMapper on the device (reference main.py:107-127)."""

    def __init__(self, model_data, joint_map, dtype=torch.float32, device=None, **model_kw):
        if not torch.cuda.is_available():
            raise RuntimeError('smplifyx_b200 needs a CUDA device (sm_100a); there is no CPU path')
        self.lib = N.load_library()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.dtype = dtype
        self.np_dtype = np.float64 if dtype == torch.float64 else np.float32
        self.K = int(len(joint_map))
        self.model_kw = dict(model_kw)
        faces = np.asarray(model_data['f']).astype(np.int64)
        self.faces = faces
        self.V = int(np.asarray(model_data['v_template']).shape[0])
        # joint with the largest skinning weight per vertex (face grouping of the unfiltered collision search)
        self.dominant_joint = np.asarray(model_data['weights']).argmax(axis=1).astype(np.int64)
        desc, keep = N.build_model_desc(model_data, joint_map, use_double=dtype == torch.float64,
                                        **model_kw)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(self.lib, self.lib.sfx_model_create(C.byref(desc), C.byref(h)))
        del keep
        self.h = h
        self.n_hand = int(desc.n_hand)
        self.n_betas = int(desc.num_betas)
        self.n_expr = int(desc.num_expr)

    def set_vposer(self, weights):
        """Installs the VPoser decoder: dict with dec_fc1_w/b, dec_fc2_w/b, dec_out_w/b (numpy,
        torch Linear layout) -- see ``vposer.load_vposer_weights``."""
        if getattr(self, '_vposer_id', None) == id(weights):
            return
        arrs = [np.ascontiguousarray(np.asarray(weights[k]), dtype=self.np_dtype) for k in
                ('dec_fc1_w', 'dec_fc1_b', 'dec_fc2_w', 'dec_fc2_b', 'dec_out_w', 'dec_out_b')]
        shapes = [(512, 32), (512,), (512, 512), (512,), (126, 512), (126,)]
        for a, sh in zip(arrs, shapes):
            if a.shape != sh:
                raise ValueError('VPoser decoder weight has shape {}, expected {}'.format(a.shape, sh))
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            N.check(self.lib, self.lib.sfx_model_set_vposer(
                self.h, *[C.c_void_p(a.ctypes.data) for a in arrs]))
        self._vposer_id = id(weights)

    def set_gmm(self, prior):
        """Installs a ``prior.MaxMixturePrior`` (or anything with ``means`` [M,D], ``precisions``
        [M,D,D], ``nll_weights`` [1,M] tensors) as the body-pose mixture prior of this model."""
        if getattr(self, '_gmm_id', None) == id(prior):
            return
        tt = self.dtype
        means = prior.means.detach().to('cpu', tt).contiguous().numpy()
        prec = prior.precisions.detach().to('cpu', tt).contiguous().numpy()
        logw = torch.log(prior.nll_weights.detach().to('cpu', tt)).reshape(-1).contiguous().numpy()
        M, D = means.shape
        p = lambda a: C.c_void_p(a.ctypes.data)
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            N.check(self.lib, self.lib.sfx_model_set_gmm(self.h, M, D, p(means), p(prec), p(logw)))
        self._gmm_id = id(prior)

    def set_collision(self, faces_segm, faces_parents, ign_part_pairs=None):
        """Installs the face segmentation of the interpenetration term: ``segm`` / ``parents``
        of ``part_segm_fn`` (reference fit_single_frame.py:317-328) and ``ign_part_pairs``
        (strings "a,b" as in the yaml files, or pairs of ints)."""
        segm = np.ascontiguousarray(np.asarray(faces_segm), dtype=np.int32).reshape(-1)
        par = np.ascontiguousarray(np.asarray(faces_parents), dtype=np.int32).reshape(-1)
        # same tables as the ones installed (by content: every FitPlan builds its own arrays): nothing to do
        key = (hash(segm.tobytes()), hash(par.tobytes()), tuple(map(str, ign_part_pairs or ())))
        if getattr(self, '_coll_key', None) == key:
            return
        if segm.shape[0] != self.faces.shape[0] or par.shape[0] != segm.shape[0]:
            raise ValueError('part segmentation has {} faces, the model {}'.format(
                segm.shape[0], self.faces.shape[0]))
        ign = [[int(x) for x in (p.split(',') if isinstance(p, str) else p)]
               for p in (ign_part_pairs or [])]
        ign = np.ascontiguousarray(np.asarray(ign, dtype=np.int32).reshape(-1, 2))
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            N.check(self.lib, self.lib.sfx_model_set_collision(
                self.h, C.c_void_p(segm.ctypes.data), C.c_void_p(par.ctypes.data),
                C.c_void_p(ign.ctypes.data) if ign.shape[0] else None, int(ign.shape[0])))
        self._coll_key = key
        self.has_collision = True

    def set_collision_unfiltered(self, faces_group=None):
        """The interpenetration term without FilterFaces (the reference's path when no
        ``part_segm_fn`` is given, fit_single_frame.py:317-328): every intersecting pair of
        triangles that share no vertex counts.  ``faces_group`` [F] in 0..63 only groups the faces
        for the search; default: the joint with the largest skinning weight of a face's first
        vertex, folded into 64 groups."""
        if faces_group is None:
            faces_group = (self.dominant_joint[self.faces[:, 0]] % 64)
        grp = np.ascontiguousarray(np.asarray(faces_group), dtype=np.int32).reshape(-1)
        if grp.shape[0] != self.faces.shape[0]:
            raise ValueError('face grouping has {} faces, the model {}'.format(grp.shape[0], self.faces.shape[0]))
        key = ('unfiltered', hash(grp.tobytes()))
        if getattr(self, '_coll_key', None) == key:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            N.check(self.lib, self.lib.sfx_model_set_collision_unfiltered(self.h, C.c_void_p(grp.ctypes.data)))
        self._coll_key = key
        self.has_collision = True

    def close(self):
        if getattr(self, 'h', None):
            self.lib.sfx_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FrameBatch(object):
    """B independent frames: parameters, targets and optimiser workspace on the device."""

    def __init__(self, model, num_frames, use_vposer=False):
        self.model = model
        self.lib = model.lib
        self.B = int(num_frames)
        self.use_vposer = bool(use_vposer)
        h = C.c_void_p()
        with torch.cuda.device(model.device):
            N.check(self.lib, self.lib.sfx_batch_create(model.h, self.B, int(use_vposer),
                                                        C.byref(h)))
        self.h = h
        self.L = N.SfxLayout()
        N.check(self.lib, self.lib.sfx_batch_layout(self.h, C.byref(self.L)))
        self.blocks = N.param_blocks(self.L)

    def enable_collisions(self):
        """Allocates the full-mesh workspace of the interpenetration term (needs
        ``Model.set_collision``)."""
        with torch.cuda.device(self.model.device):
            N.check(self.lib, self.lib.sfx_batch_enable_collisions(self.h))

    def close(self):
        if getattr(self, 'h', None):
            self.lib.sfx_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ host <-> device
    def _host(self, a, shape, dtype=None):
        a = np.ascontiguousarray(np.asarray(a), dtype=dtype or self.model.np_dtype)
        if tuple(a.shape) != tuple(shape):
            raise ValueError('expected shape {}, got {}'.format(tuple(shape), a.shape))
        return a

    def set_targets(self, keypoints, joint_weights, lowconf, init_mask, cam, reg_pose=None):
        """Host arrays (numpy, or pinned CPU tensors) -> device, asynchronously on the current
        stream.  Shapes as in ``sfx_batch_set_targets`` (include/sfx.h)."""
        B, K = self.B, self.model.K
        kp = self._host(keypoints, (B, K, 3))
        jw = self._host(joint_weights, (B, K))
        lc = self._host(lowconf, (B, K), np.uint8)
        im = self._host(init_mask, (B, K), np.uint8)
        cm = self._host(cam, (B, N.SFX_CAM_STRIDE))
        rp = None if reg_pose is None else self._host(reg_pose, (B, self.L.n_pose))
        self._keep = (kp, jw, lc, im, cm, rp)
        p = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
        with torch.cuda.device(self.model.device):
            N.check(self.lib, self.lib.sfx_batch_set_targets(self.h, p(kp), p(jw), p(lc), p(im),
                                                             p(cm), p(rp), _stream()))
        return kp.nbytes + jw.nbytes + lc.nbytes + im.nbytes + cm.nbytes + \
            (0 if rp is None else rp.nbytes)

    def set_targets_dev(self, gt, conf, joint_weights, lowconf, init_mask, cam, reg_pose=None):
        """Same as set_targets with CUDA tensors in the batch dtype (gt [B,K,2], conf [B,K])."""
        with torch.cuda.device(self.model.device):
            N.check(self.lib, self.lib.sfx_batch_set_targets_dev(
                self.h, _ptr(gt), _ptr(conf), _ptr(joint_weights), _ptr(lowconf), _ptr(init_mask),
                _ptr(cam), _ptr(reg_pose)))

    def set_targets_from_keypoints(self, keypoints, base_joint_weights, init_joints_idxs, n_body,
                                   confidence_threshold, cam, reg_pose=None):
        """Device ingestion (``sfx_keypoint_masks``, fit_single_frame.py:276-294): keypoints
        [B,K,3] float32 (numpy, copied once, or a CUDA tensor) -> gt / conf / joint weights /
        low-confidence mask / trimmed init joints computed on the device and installed as the
        batch targets.  Returns the host->device byte count."""
        B, K = self.B, self.model.K
        dev, dt = self.model.device, self.model.dtype
        nbytes = 0
        if torch.is_tensor(keypoints) and keypoints.is_cuda:
            kp = keypoints.to(torch.float32).contiguous()
        else:
            kp_h = self._host(keypoints, (B, K, 3), np.float32)
            kp = torch.as_tensor(kp_h, device=dev)
            nbytes += kp_h.nbytes
        if tuple(kp.shape) != (B, K, 3):
            raise ValueError('expected keypoints [{}, {}, 3]'.format(B, K))
        bjw = torch.as_tensor(np.asarray(base_joint_weights, dtype=np.float32).reshape(K), device=dev)
        idx = torch.as_tensor(np.asarray(init_joints_idxs, dtype=np.int32).reshape(-1), device=dev)
        cm_h = self._host(cam, (B, N.SFX_CAM_STRIDE))
        cm = torch.as_tensor(cm_h, device=dev)
        rp = None
        if reg_pose is not None:
            rp_h = self._host(reg_pose, (B, self.L.n_pose))
            rp = torch.as_tensor(rp_h, device=dev)
            nbytes += rp_h.nbytes
        nbytes += bjw.numel() * 4 + idx.numel() * 4 + cm_h.nbytes
        gt, conf, jw = self._new(B, K, 2), self._new(B, K), self._new(B, K)
        low = torch.empty((B, K), dtype=torch.uint8, device=dev)
        init = torch.empty((B, K), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib, self.lib.sfx_keypoint_masks(
                _ptr(kp), _ptr(bjw), _ptr(idx), int(idx.numel()), int(n_body),
                float(confidence_threshold), B, K, int(dt == torch.float64), _ptr(gt), _ptr(conf),
                _ptr(jw), _ptr(low), _ptr(init), _stream()))
        self.set_targets_dev(gt, conf, jw, low, init, cm, rp)
        self._keep_dev = (kp, gt, conf, jw, low, init, cm, rp)
        return nbytes

    def set_params(self, params):
        x = self._host(params, (self.B, self.L.np))
        self._keep_x = x
        with torch.cuda.device(self.model.device):
            N.check(self.lib, self.lib.sfx_batch_set_params(self.h, C.c_void_p(x.ctypes.data),
                                                            _stream()))
        return x.nbytes

    def get_params(self):
        out = np.empty((self.B, self.L.np), dtype=self.model.np_dtype)
        with torch.cuda.device(self.model.device):
            N.check(self.lib, self.lib.sfx_batch_get_params(self.h, C.c_void_p(out.ctypes.data),
                                                            _stream()))
            torch.cuda.current_stream().synchronize()
        return out

    def params_tensor(self):
        """Device view [B, np] of the live parameters (no copy)."""
        ptr = self.lib.sfx_batch_params_dev(self.h)
        return _wrap(ptr, (self.B, self.L.np), self.model.dtype, self.model.device, self)

    def _new(self, *shape, dtype=None):
        return torch.empty(shape, dtype=dtype or self.model.dtype, device=self.model.device)

    # ------------------------------------------------------------------ compute
    def eval(self, stage, want_joints=False):
        """One closure evaluation per frame -> (loss [B], grad [B, np], joints [B, K, 3]|None)."""
        loss = self._new(self.B)
        grad = self._new(self.B, self.L.np)
        joints = self._new(self.B, self.model.K, 3) if want_joints else None
        with torch.cuda.device(self.model.device):
            N.check(self.lib, self.lib.sfx_eval(self.h, C.byref(stage), _ptr(loss), _ptr(grad),
                                                _ptr(joints), _stream()))
        return loss, grad, joints

    def fit_stage(self, stage, frame_ids=None, out=None):
        """FittingMonitor.run_fitting for every frame (or the int32 device tensor
        ``frame_ids``) in one launch; returns the per-frame return values [B] (device)."""
        final = out if out is not None else self._new(self.B)
        n = 0
        if frame_ids is not None:
            if frame_ids.dtype != torch.int32 or not frame_ids.is_cuda:
                raise ValueError('frame_ids must be an int32 CUDA tensor')
            n = int(frame_ids.numel())
        with torch.cuda.device(self.model.device):
            N.check(self.lib, self.lib.sfx_fit_stage(self.h, C.byref(stage), _ptr(frame_ids), n,
                                                     _ptr(final), _stream()))
        return final

    def fit_pipeline(self, pipeline, order=None, flip=None):
        """Camera stage + every annealing stage (+ flipped orientation and argmin selection for
        the frames flagged in the uint8 device tensor ``flip``) in ONE launch."""
        with torch.cuda.device(self.model.device):
            N.check(self.lib, self.lib.sfx_fit_pipeline(self.h, C.byref(pipeline), _ptr(order),
                                                        _ptr(flip), _stream()))

    def cam_loss(self):
        ptr = self.lib.sfx_batch_cam_loss_dev(self.h)
        return _wrap(ptr, (self.B,), self.model.dtype, self.model.device, self)

    def forward_mesh(self, want_joints=True, last_orientation=False, out=None):
        """body_model(return_verts=True): vertices [B, V, 3] (+ mapped joints [B, K, 3]).
        ``last_orientation``: at the parameters of the last fitted orientation of a pipeline
        launch (what the reference writes to vertices.ply)."""
        verts = out[0] if out is not None else self._new(self.B, self.model.V, 3)
        joints = (out[1] if out is not None else self._new(self.B, self.model.K, 3)) \
            if want_joints else None
        fn = self.lib.sfx_forward_mesh_last if last_orientation else self.lib.sfx_forward_mesh
        with torch.cuda.device(self.model.device):
            N.check(self.lib, fn(self.h, _ptr(verts), _ptr(joints), _stream()))
        return verts, joints

    def begin_orientation(self, flip, frame_ids=None):
        """reset_params(global_orient=orient, body_pose=pose_embedding) on the device
        (fit_single_frame.py:546-551); ``flip`` starts from the 180-degree rotated orientation."""
        n = 0 if frame_ids is None else int(frame_ids.numel())
        with torch.cuda.device(self.model.device):
            N.check(self.lib, self.lib.sfx_batch_begin_orientation(
                self.h, int(bool(flip)), _ptr(frame_ids), n, _stream()))

    def select_orientation(self, frame_ids):
        """Keeps, per frame, the orientation with the lower final loss."""
        with torch.cuda.device(self.model.device):
            N.check(self.lib, self.lib.sfx_batch_select_orientation(
                self.h, _ptr(frame_ids), int(frame_ids.numel()), _stream()))

    def final_loss(self):
        ptr = self.lib.sfx_batch_final_loss_dev(self.h)
        return _wrap(ptr, (self.B,), self.model.dtype, self.model.device, self)

    def evals(self):
        ptr = self.lib.sfx_batch_evals_dev(self.h)
        return _wrap(ptr, (self.B,), torch.int32, self.model.device, self)

    def passes(self):
        ptr = self.lib.sfx_batch_passes_dev(self.h)
        return _wrap(ptr, (self.B,), torch.int32, self.model.device, self)

    def flags(self):
        ptr = self.lib.sfx_batch_flags_dev(self.h)
        return _wrap(ptr, (self.B,), torch.int32, self.model.device, self)

    def guess_init(self, joints, edge_idxs, focal, mean_len2d, need):
        """fitting.guess_init on the device (``sfx_batch_guess_init``): writes the estimated camera
        translation of the frames flagged in ``need`` into the parameters and the depth-prior
        target.  ``joints``: device tensor [B,K,3] from ``eval(..., want_joints=True)``."""
        e = np.ascontiguousarray(np.asarray(edge_idxs, dtype=np.int32).reshape(-1, 2))
        f = np.ascontiguousarray(np.broadcast_to(np.asarray(focal, dtype=np.float64), (self.B,)))
        l2 = np.ascontiguousarray(mean_len2d, dtype=np.float64)
        nd = np.ascontiguousarray(need, dtype=np.uint8)
        p = lambda a: C.c_void_p(a.ctypes.data)
        with torch.cuda.device(self.model.device):
            N.check(self.lib, self.lib.sfx_batch_guess_init(self.h, _ptr(joints), p(e), int(e.shape[0]),
                                                            p(f), p(l2), p(nd), _stream()))
        return f.nbytes + l2.nbytes + nd.nbytes

    def frame_cycles(self):
        """[B] int64: SM cycles the block(s) of the last pipeline launch spent on each frame (the
        leader block of a frame that ran on a cluster)."""
        self.lib.sfx_batch_prof_dev.restype = C.c_void_p
        self.lib.sfx_batch_prof_dev.argtypes = [C.c_void_p]
        ptr = self.lib.sfx_batch_prof_dev(self.h)
        raw = _wrap(ptr, (self.B * 128,), torch.int32, self.model.device, self)
        return raw.view(torch.int64).view(self.B, 64)[:, 4]

    def coll_stats(self):
        """[B, 4] int32 per frame, largest over its evaluations of the interpenetration term:
        candidate faces, touched vertices, sweep iterations and listed partners of warp 0 (None
        when the term is not enabled)."""
        ptr = self.lib.sfx_batch_coll_stat_dev(self.h)
        if not ptr:
            return None
        return _wrap(ptr, (self.B, 4), torch.int32, self.model.device, self)

    def reset_counters(self):
        with torch.cuda.device(self.model.device):
            N.check(self.lib, self.lib.sfx_batch_reset_counters(self.h, _stream()))


class _CudaArray(object):
    """__cuda_array_interface__ holder so torch can view library-owned device memory."""

    def __init__(self, ptr, shape, typestr, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {
            'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False), 'version': 2}


def _wrap(ptr, shape, dtype, device, owner):
    typestr = {torch.float32: '<f4', torch.float64: '<f8', torch.int32: '<i4'}[dtype]
    with torch.cuda.device(device):
        return torch.as_tensor(_CudaArray(ptr, shape, typestr, owner), device=device)
