"""ctypes binding of the C ABI declared in ``include/sfx.h`` (``libsfx.so``).

There is no CPU path: if the CUDA library is missing or no GPU is visible the product raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SFX_LIB: another build of the library in csrc/ (development A/B runs only)
LIB_PATH = os.path.join(_HERE, 'csrc', os.environ.get('SFX_LIB', 'libsfx.so'))

SFX_MAX_BLOCKS = 12
SFX_NP_MAX = 192
SFX_KMAX = 160
SFX_CAM_STRIDE = 16
SFX_FLAG_NAN, SFX_FLAG_INF, SFX_FLAG_COLL_OVERFLOW = 1, 2, 4
SFX_CAM_FX, SFX_CAM_FY, SFX_CAM_CX, SFX_CAM_CY, SFX_CAM_R, SFX_CAM_DW, SFX_CAM_TZ = 0, 1, 2, 3, 4, 13, 14

LOSS_SMPLIFY, LOSS_CAMERA_INIT = 0, 1
ALIGN_NONE, ALIGN_PROCRUSTES, ALIGN_PELVIS, ALIGN_SCALE = 0, 1, 2, 3
OPT_LBFGSLS, OPT_ADAM = 0, 1
PPRIOR_L2, PPRIOR_REGRESSION, PPRIOR_GMM, PPRIOR_LATENT = 0, 1, 2, 3

# smplx.vertex_ids.vertex_ids['smplx'] in VertexJointSelector order (third-party smplx 0.1.x)
EXTRA_VERTEX_IDS = np.array(
    [9120, 9929, 9448, 616, 6, 5770, 5780, 8846, 8463, 8474, 8635,
     5361, 4933, 5058, 5169, 5286, 8079, 7669, 7794, 7905, 8022], dtype=np.int32)


class SfxLayout(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ['n_betas', 'n_expr', 'n_hand', 'n_pose', 'off_betas', 'off_go', 'off_lh',
                 'off_rh', 'off_jaw', 'off_leye', 'off_reye', 'off_expr', 'off_pose',
                 'off_camt', 'np']]


class SfxStage(C.Structure):
    _fields_ = [
        ('loss_kind', C.c_int), ('opt_kind', C.c_int), ('pprior_kind', C.c_int),
        ('stage_index', C.c_int), ('num_stages', C.c_int), ('use_joints_conf', C.c_int),
        ('use_conf_camera', C.c_int), ('use_vposer', C.c_int), ('n_body_kpts', C.c_int),
        ('rho', C.c_double),
        ('body_pose_weight', C.c_double), ('shape_weight', C.c_double),
        ('bending_prior_weight', C.c_double), ('hand_prior_weight', C.c_double),
        ('expr_prior_weight', C.c_double), ('jaw_prior_weight', C.c_double * 3),
        ('hand_joint_weight', C.c_double), ('face_joint_weight', C.c_double),
        ('depth_loss_weight', C.c_double),
        ('maxiters', C.c_int), ('ftol', C.c_double), ('gtol', C.c_double),
        ('lr', C.c_double), ('max_iter', C.c_int), ('max_eval', C.c_int), ('history', C.c_int),
        ('tol_grad', C.c_double), ('tol_change', C.c_double),
        ('adam_beta1', C.c_double), ('adam_beta2', C.c_double), ('adam_eps', C.c_double),
        ('n_active', C.c_int), ('n_blocks', C.c_int),
        ('block_start', C.c_int * SFX_MAX_BLOCKS), ('block_len', C.c_int * SFX_MAX_BLOCKS),
        ('block_off', C.c_int * SFX_MAX_BLOCKS), ('need_blend_grad', C.c_int),
        ('generic_two_loop', C.c_int),
        ('coll_loss_weight', C.c_double), ('coll_sigma', C.c_double),
    ]


SFX_MAX_STAGES = 8


class SfxPipeline(C.Structure):
    _fields_ = [('n_stages', C.c_int), ('n_wide', C.c_int), ('cam', SfxStage),
                ('body', SfxStage * SFX_MAX_STAGES)]


WIDE_CLUSTER = 8      # CTAs per wide frame (csrc/sfx_stream.cuh SFX_WIDE_CLUSTER)


def make_pipeline(cam_stage, body_stages, n_wide=0):
    """``n_wide``: the first n_wide frames of the launch order each get a cluster of
    ``WIDE_CLUSTER`` CTAs (SfxPipeline.n_wide)."""
    if not 1 <= len(body_stages) <= SFX_MAX_STAGES:
        raise ValueError('between 1 and {} annealing stages are supported'.format(SFX_MAX_STAGES))
    P = SfxPipeline()
    P.n_stages = len(body_stages)
    P.n_wide = int(n_wide)
    P.cam = cam_stage
    for i, st in enumerate(body_stages):
        P.body[i] = st
    return P


class SfxModelDesc(C.Structure):
    _fields_ = [
        ('num_verts', C.c_int32), ('num_faces', C.c_int32),
        ('v_template', C.c_void_p), ('shapedirs', C.c_void_p), ('shape_stride', C.c_int32),
        ('num_betas', C.c_int32), ('expr_offset', C.c_int32), ('num_expr', C.c_int32),
        ('posedirs', C.c_void_p), ('J_regressor', C.c_void_p), ('lbs_weights', C.c_void_p),
        ('parents', C.c_void_p), ('faces', C.c_void_p), ('n_hand', C.c_int32),
        ('hand_components_l', C.c_void_p), ('hand_components_r', C.c_void_p),
        ('hand_mean_l', C.c_void_p), ('hand_mean_r', C.c_void_p),
        ('extra_vertex_ids', C.c_void_p), ('lmk_faces_idx', C.c_void_p),
        ('lmk_bary_coords', C.c_void_p), ('use_face_contour', C.c_int32),
        ('dyn_lmk_faces_idx', C.c_void_p), ('dyn_lmk_bary_coords', C.c_void_p),
        ('joint_map', C.c_void_p), ('num_keypoints', C.c_int32), ('use_double', C.c_int32),
        ('arrays_float64', C.c_int32),
    ]


def build_model_desc(model_data, joint_map, num_betas=10, num_expression_coeffs=10,
                     use_pca=True, num_pca_comps=6, flat_hand_mean=False,
                     use_face_contour=False, use_double=False):
    """``model_data``: dict with the SMPL-X npz keys.  Returns (desc, keepalive)."""
    keep = []

    def arr(x, dt):
        a = np.ascontiguousarray(np.asarray(x), dtype=dt)
        keep.append(a)
        return a.ctypes.data_as(C.c_void_p)

    # float64 batches take the model arrays as the npz holds them (the licensed files store float64);
    # float32 batches the float32 values the reference's float_dtype: float32 path works with
    fdt = np.float64 if use_double else np.float32
    d = SfxModelDesc()
    vt = np.asarray(model_data['v_template'])
    d.num_verts = vt.shape[0]
    d.num_faces = np.asarray(model_data['f']).shape[0]
    d.v_template = arr(vt, fdt)
    sd = np.asarray(model_data['shapedirs'])
    d.shapedirs = arr(sd, fdt)
    d.shape_stride = sd.shape[2]
    d.num_betas = num_betas
    d.expr_offset = 300
    d.num_expr = num_expression_coeffs
    if sd.shape[2] < 300 + num_expression_coeffs:
        raise ValueError('shapedirs has {} columns; expression block needs [300, {})'.format(
            sd.shape[2], 300 + num_expression_coeffs))
    d.posedirs = arr(model_data['posedirs'], fdt)
    d.J_regressor = arr(model_data['J_regressor'], fdt)
    d.lbs_weights = arr(model_data['weights'], fdt)
    parents = np.asarray(model_data['kintree_table'])[0].astype(np.int64).copy()
    parents[0] = -1
    d.parents = arr(parents, np.int32)
    d.faces = arr(np.asarray(model_data['f']).astype(np.int64), np.int32)
    if use_pca:
        n_hand = int(num_pca_comps)
        cl = np.asarray(model_data['hands_componentsl'])[:n_hand]
        cr = np.asarray(model_data['hands_componentsr'])[:n_hand]
    else:
        n_hand = 45
        cl = cr = np.eye(45)
    d.n_hand = n_hand
    d.hand_components_l = arr(cl, fdt)
    d.hand_components_r = arr(cr, fdt)
    zeros = np.zeros(45)
    d.hand_mean_l = arr(zeros if flat_hand_mean else model_data['hands_meanl'], fdt)
    d.hand_mean_r = arr(zeros if flat_hand_mean else model_data['hands_meanr'], fdt)
    d.extra_vertex_ids = arr(EXTRA_VERTEX_IDS, np.int32)
    d.lmk_faces_idx = arr(model_data['lmk_faces_idx'], np.int32)
    d.lmk_bary_coords = arr(model_data['lmk_bary_coords'], fdt)
    d.use_face_contour = 1 if use_face_contour else 0
    if use_face_contour:
        d.dyn_lmk_faces_idx = arr(model_data['dynamic_lmk_faces_idx'], np.int32)
        d.dyn_lmk_bary_coords = arr(model_data['dynamic_lmk_bary_coords'], fdt)
    jm = np.asarray(joint_map, dtype=np.int32)
    d.joint_map = arr(jm, np.int32)
    d.num_keypoints = jm.shape[0]
    d.use_double = 1 if use_double else 0
    d.arrays_float64 = 1 if use_double else 0
    return d, keep


def make_layout(n_betas, n_expr, n_hand, use_vposer):
    """Mirror of ``sfx::make_layout`` (csrc/sfx_model_prep.h)."""
    L = SfxLayout()
    L.n_betas, L.n_expr, L.n_hand = n_betas, n_expr, n_hand
    L.n_pose = 32 if use_vposer else 63
    o = 0
    for name, n in (('off_betas', n_betas), ('off_go', 3), ('off_lh', n_hand),
                    ('off_rh', n_hand), ('off_jaw', 3), ('off_leye', 3), ('off_reye', 3),
                    ('off_expr', n_expr), ('off_pose', L.n_pose), ('off_camt', 3)):
        setattr(L, name, o)
        o += n
    L.np = o
    return L


# name -> (offset field, length) of every parameter block, in optimiser order
def param_blocks(L):
    return {
        'betas': (L.off_betas, L.n_betas), 'global_orient': (L.off_go, 3),
        'left_hand_pose': (L.off_lh, L.n_hand), 'right_hand_pose': (L.off_rh, L.n_hand),
        'jaw_pose': (L.off_jaw, 3), 'leye_pose': (L.off_leye, 3), 'reye_pose': (L.off_reye, 3),
        'expression': (L.off_expr, L.n_expr), 'pose_embedding': (L.off_pose, L.n_pose),
        'camera_translation': (L.off_camt, 3),
    }


BODY_STAGE_BLOCKS = ['betas', 'global_orient', 'left_hand_pose', 'right_hand_pose', 'jaw_pose',
                     'leye_pose', 'reye_pose', 'expression', 'pose_embedding']
CAMERA_STAGE_BLOCKS = ['camera_translation', 'global_orient']


# L-BFGS direction (lbfgs_ls.py:336-358): 'exact' = the reference's recursion and operation order,
# one warp (default); 'block' = the same by the whole block (debug A/B); 'gram' = the same
# recursion run on the inner products of the history (csrc/sfx_core.cuh gram_two_loop): equal
# in exact arithmetic, different rounding, several times shorter dependent chain.
TWO_LOOP_MODES = {'exact': 0, 'block': 1, 'gram': 2}


def two_loop_mode(name=None):
    if name is None:
        name = os.environ.get('SFX_TWO_LOOP', 'exact')
    if isinstance(name, int):
        return name
    if name not in TWO_LOOP_MODES:
        raise ValueError('two_loop must be one of {}'.format(sorted(TWO_LOOP_MODES)))
    return TWO_LOOP_MODES[name]


def make_stage(L, blocks, loss_kind=LOSS_SMPLIFY, opt_kind=OPT_LBFGSLS, pprior_kind=PPRIOR_L2,
               stage_index=0, num_stages=1, use_joints_conf=True, use_conf_camera=False,
               use_vposer=False, n_body_kpts=25, rho=100.0, body_pose_weight=0.0,
               shape_weight=0.0, bending_prior_weight=0.0, hand_prior_weight=0.0,
               expr_prior_weight=0.0, jaw_prior_weight=(0.0, 0.0, 0.0), hand_joint_weight=0.0,
               face_joint_weight=0.0, depth_loss_weight=0.0, maxiters=30, ftol=1e-9, gtol=1e-9,
               lr=1.0, max_iter=None, max_eval=None, history=100, tol_grad=1e-5,
               tol_change=1e-9, adam_beta1=0.9, adam_beta2=0.999, adam_eps=1e-8,
               coll_loss_weight=0.0, coll_sigma=0.5, two_loop=None):
    """Builds an SfxStage.  ``max_iter`` defaults to ``maxiters`` and ``max_eval`` to
    ``max_iter * 5 // 4`` exactly like create_optimizer('lbfgsls') (optim_factory.py:51-53,
    lbfgs_ls.py:202-203).  ``two_loop`` selects how the L-BFGS direction is computed (see
    ``TWO_LOOP_MODES``; default: the environment variable ``SFX_TWO_LOOP``, else 'exact')."""
    st = SfxStage()
    st.generic_two_loop = two_loop_mode(two_loop)
    st.loss_kind, st.opt_kind, st.pprior_kind = loss_kind, opt_kind, pprior_kind
    st.stage_index, st.num_stages = stage_index, num_stages
    st.use_joints_conf = int(bool(use_joints_conf))
    st.use_conf_camera = int(bool(use_conf_camera))
    st.use_vposer = int(bool(use_vposer))
    st.n_body_kpts = n_body_kpts
    st.rho = rho
    st.body_pose_weight, st.shape_weight = body_pose_weight, shape_weight
    st.bending_prior_weight, st.hand_prior_weight = bending_prior_weight, hand_prior_weight
    st.expr_prior_weight = expr_prior_weight
    for i in range(3):
        st.jaw_prior_weight[i] = float(jaw_prior_weight[i])
    st.hand_joint_weight, st.face_joint_weight = hand_joint_weight, face_joint_weight
    st.depth_loss_weight = depth_loss_weight
    st.maxiters, st.ftol, st.gtol = maxiters, ftol, gtol
    st.lr = lr
    st.max_iter = maxiters if max_iter is None else max_iter
    st.max_eval = st.max_iter * 5 // 4 if max_eval is None else max_eval
    st.history = history
    st.tol_grad, st.tol_change = tol_grad, tol_change
    st.adam_beta1, st.adam_beta2, st.adam_eps = adam_beta1, adam_beta2, adam_eps
    st.coll_loss_weight, st.coll_sigma = float(coll_loss_weight), float(coll_sigma)
    pb = param_blocks(L)
    pos = 0
    if len(blocks) > SFX_MAX_BLOCKS:
        raise ValueError('too many parameter blocks')
    for i, name in enumerate(blocks):
        off, n = pb[name]
        st.block_start[i], st.block_len[i], st.block_off[i] = pos, n, off
        pos += n
    st.n_blocks = len(blocks)
    st.n_active = pos
    st.need_blend_grad = int(any(b not in ('camera_translation', 'global_orient')
                                 for b in blocks))
    return st


_lib = None


def load_library(path=None):
    """Loads ``libsfx.so`` and declares the prototypes of ``include/sfx.h``."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.isfile(p):
        raise RuntimeError(
            'CUDA library {} is missing: build it with `python -c "import __graft_entry__ as g; '
            'g.build()"` (nvcc, sm_100a).  There is no CPU fallback.'.format(p))
    lib = C.CDLL(p)
    vp, i32 = C.c_void_p, C.c_int32
    lib.sfx_model_create.argtypes = [C.POINTER(SfxModelDesc), C.POINTER(vp)]
    lib.sfx_model_set_vposer.argtypes = [vp] * 7
    lib.sfx_model_set_gmm.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.sfx_model_set_collision.argtypes = [vp, vp, vp, vp, i32]
    lib.sfx_model_set_collision_unfiltered.argtypes = [vp, vp]
    lib.sfx_batch_enable_collisions.argtypes = [vp]
    lib.sfx_batch_coll_stat_dev.argtypes = [vp]
    lib.sfx_batch_coll_stat_dev.restype = vp
    lib.sfx_model_destroy.argtypes = [vp]
    lib.sfx_model_destroy.restype = None
    lib.sfx_batch_create.argtypes = [vp, i32, i32, C.POINTER(vp)]
    lib.sfx_batch_destroy.argtypes = [vp]
    lib.sfx_batch_destroy.restype = None
    lib.sfx_batch_layout.argtypes = [vp, C.POINTER(SfxLayout)]
    lib.sfx_batch_set_targets.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    lib.sfx_batch_set_targets_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    lib.sfx_batch_set_params.argtypes = [vp, vp, vp]
    lib.sfx_batch_get_params.argtypes = [vp, vp, vp]
    lib.sfx_batch_params_dev.argtypes = [vp]
    lib.sfx_batch_params_dev.restype = vp
    lib.sfx_eval.argtypes = [vp, C.POINTER(SfxStage), vp, vp, vp, vp]
    lib.sfx_fit_stage.argtypes = [vp, C.POINTER(SfxStage), vp, i32, vp, vp]
    lib.sfx_forward_mesh.argtypes = [vp, vp, vp, vp]
    lib.sfx_fit_pipeline.argtypes = [vp, C.POINTER(SfxPipeline), vp, vp, vp]
    lib.sfx_batch_cam_loss_dev.argtypes = [vp]
    lib.sfx_batch_cam_loss_dev.restype = vp
    lib.sfx_forward_mesh_last.argtypes = [vp, vp, vp, vp]
    lib.sfx_batch_begin_orientation.argtypes = [vp, i32, vp, i32, vp]
    lib.sfx_batch_select_orientation.argtypes = [vp, vp, i32, vp]
    lib.sfx_batch_final_loss_dev.argtypes = [vp]
    lib.sfx_batch_final_loss_dev.restype = vp
    lib.sfx_batch_evals_dev.argtypes = [vp]
    lib.sfx_batch_evals_dev.restype = vp
    lib.sfx_batch_flags_dev.argtypes = [vp]
    lib.sfx_batch_flags_dev.restype = vp
    lib.sfx_batch_passes_dev.argtypes = [vp]
    lib.sfx_batch_passes_dev.restype = vp
    lib.sfx_batch_reset_counters.argtypes = [vp, vp]
    lib.sfx_aligned_errors.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.sfx_batch_guess_init.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp]
    lib.sfx_pack_keypoints.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp]
    lib.sfx_keypoint_masks.argtypes = [vp, vp, vp, i32, i32, C.c_float, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.sfx_blend_keypoints.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp]
    lib.sfx_diag_l2_read_gbs.argtypes = [C.c_int64, i32, C.POINTER(C.c_double)]
    lib.sfx_last_error.restype = C.c_char_p
    lib.sfx_version.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def check(lib, status):
    if status != 0:
        raise RuntimeError('libsfx: ' + (lib.sfx_last_error() or b'?').decode())
