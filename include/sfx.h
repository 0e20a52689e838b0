/* sfx.h -- C ABI of the B200 SMPL-X fitting engine (libsfx.so).
 *
 * The reference (xiyichen/smplify-x-partial) has no FFI: its seam is the Python API of
 * smplifyx/fitting.py, camera.py, prior.py, optimizers/ (SURVEY.md section 8b).  Each entry
 * point below names the reference interface it stands in for; the Python mirror in
 * smplify-x-partial_b200/ binds them with ctypes (see INTEGRATION.md for the stub a
 * maintainer of the reference would add).
 *
 * Conventions: plain pointers and sizes, no torch types.  Every function returns 0 on success
 * or a negative sfx_status; sfx_last_error() gives the message of the last failure on the
 * calling thread.  "dev" pointers are CUDA device pointers owned by the caller, row-major
 * contiguous; "host" pointers are ordinary memory.  `stream` is a cudaStream_t passed as
 * void*.  A model handle is immutable and may be shared by several batches; a batch handle is
 * not re-entrant.  Nothing allocates after sfx_batch_create.
 */
#ifndef SFX_H
#define SFX_H
#include <stdint.h>
#include "../smplify-x-partial_b200/csrc/sfx_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SFX_OK = 0,
    SFX_ERR_ARG = -1,
    SFX_ERR_CUDA = -2,
    SFX_ERR_UNSUPPORTED = -3
} sfx_status;

typedef struct sfx_model sfx_model;
typedef struct sfx_batch sfx_batch;

/* Raw SMPL-X arrays as stored in SMPLX_{NEUTRAL,MALE,FEMALE}.npz (host pointers; the
 * floating-point arrays are float32, or float64 when arrays_float64 is set -- the shipped files
 * hold float64 and the reference's float_dtype: float64 path keeps them so; integers are int32).
 * Stands in for the constructor arguments of smplx.create (reference smplifyx/main.py:109-127)
 * plus JointMapper(dataset.get_model2data()) (main.py:107). */
typedef struct {
    int32_t num_verts, num_faces;
    const void* v_template;          /* [V,3] */
    const void* shapedirs;           /* [V,3,shape_stride] */
    int32_t shape_stride;            /* 400 in the shipped files */
    int32_t num_betas;               /* columns [0, num_betas) */
    int32_t expr_offset, num_expr;   /* columns [expr_offset, expr_offset + num_expr) */
    const void* posedirs;            /* [V,3,486] */
    const void* J_regressor;         /* [55,V] */
    const void* lbs_weights;         /* [V,55] */
    const int32_t* parents;          /* [55], root = -1 */
    const int32_t* faces;            /* [F,3] */
    int32_t n_hand;                  /* num_pca_comps, or 45 when use_pca is False */
    const void* hand_components_l;   /* [n_hand,45] (identity rows when use_pca is False) */
    const void* hand_components_r;   /* [n_hand,45] */
    const void* hand_mean_l;         /* [45] (zeros when flat_hand_mean) */
    const void* hand_mean_r;         /* [45] */
    const int32_t* extra_vertex_ids; /* [21] smplx vertex_ids['smplx'] in selector order */
    const int32_t* lmk_faces_idx;    /* [51] */
    const void* lmk_bary_coords;     /* [51,3] */
    int32_t use_face_contour;
    const int32_t* dyn_lmk_faces_idx;   /* [79,17] or NULL */
    const void* dyn_lmk_bary_coords;    /* [79,17,3] or NULL */
    const int32_t* joint_map;        /* [num_keypoints] model joint index per keypoint */
    int32_t num_keypoints;
    int32_t use_double;              /* float_dtype float64 (reference main.py:99-105) */
    int32_t arrays_float64;          /* the floating-point arrays above are double, not float */
} sfx_model_desc;

/* smplx.create(...).to(device) -- copies and re-lays the constants on the current device. */
int sfx_model_create(const sfx_model_desc* desc, sfx_model** out);
void sfx_model_destroy(sfx_model* m);
/* VPoser v1 decoder weights (human_body_prior cvpr19 `bodyprior_dec_fc1/fc2/out`, loaded by the
 * reference at fit_single_frame.py:241-243): fc1 [512,32] + [512], fc2 [512,512] + [512], out
 * [126,512] + [126], torch nn.Linear layout, model dtype, host pointers.  Needed before
 * sfx_batch_create(..., use_vposer = 1): the pose block of such a batch is the 32-D latent and
 * every evaluation decodes it (fitting.py:236). */
int sfx_model_set_vposer(sfx_model* m, const void* fc1_w, const void* fc1_b, const void* fc2_w,
                         const void* fc2_b, const void* out_w, const void* out_b);
/* MaxMixturePrior on the body pose (reference prior.py:100-231, created at main.py:133-136):
 * means [M,D], precisions [M,D,D], log(nll_weights) [M] as the reference's buffers hold them, in
 * the model dtype (host pointers).  Call before any batch of this model is evaluated. */
int sfx_model_set_gmm(sfx_model* m, int32_t num_gaussians, int32_t dim, const void* means,
                      const void* precisions, const void* log_nll_weights);

/* Interpenetration term (reference fitting.py:437-455): the face segmentation the reference
 * loads from part_segm_fn at fit_single_frame.py:317-328 (`segm` [F] = body part of every face,
 * `parents` [F] = kinematic parent of that part; parts 0..63) and its ign_part_pairs option
 * ([n,2] part ids) -- the inputs of mesh_intersection.FilterFaces.  Stands in for the
 * construction of BVH / FilterFaces / DistanceFieldPenetrationLoss at fit_single_frame.py:300-328.
 * This entry takes the segmentation (sfx_model_set_collision_unfiltered below runs without);
 * point2plane = False and penalize_outside = True, the values of every shipped configuration,
 * are what is built. */
int sfx_model_set_collision(sfx_model* m, const int32_t* faces_segm, const int32_t* faces_parents,
                            const int32_t* ign_part_pairs, int32_t n_ign_pairs);
/* The same term WITHOUT FilterFaces -- the reference's path when part_segm_fn is empty
 * (fit_single_frame.py:317-328: filter_faces = None; fitting.py:449-450 then keeps every pair the
 * search tree reports): every intersecting pair of triangles that share no vertex is penalised.
 * faces_group [F] (0..63) only groups the faces for the broad phase (any spatially coherent
 * grouping, e.g. the joint with the largest skinning weight); it does not change the result. */
int sfx_model_set_collision_unfiltered(sfx_model* m, const int32_t* faces_group);

/* Per-batch workspace: parameters, targets, L-BFGS history (154 KB per frame in float32) and its
 * inner products (80 KB per frame; SfxStage.generic_two_loop = 2) for B independent frames.
 * use_vposer selects a 32-D latent pose block instead of the 63-D axis-angle one. */
int sfx_batch_create(const sfx_model* m, int32_t num_frames, int32_t use_vposer, sfx_batch** out);
void sfx_batch_destroy(sfx_batch* b);
int sfx_batch_layout(const sfx_batch* b, SfxLayout* out);
/* Allocates the full-mesh workspace of the interpenetration term (about 4.9 MB per frame in
 * float32).  Needed before a stage with coll_loss_weight > 0 (SfxStage) is evaluated or fitted;
 * df_cone_height travels as SfxStage.coll_sigma. */
int sfx_batch_enable_collisions(sfx_batch* b);
/* Diagnostics of the term ([B][4] int32, device; NULL before sfx_batch_enable_collisions), per
 * frame the largest value any of its evaluations saw: candidate faces, touched vertices, and for
 * warp 0 (which owns 1/16 of the candidates) sweep iterations and listed partner faces. */
int32_t* sfx_batch_coll_stat_dev(sfx_batch* b);

/* Targets of every frame (host pointers; copied with cudaMemcpyAsync on `stream`):
 *   keypoints [B,K,3] (x, y, conf)        fit_single_frame.py:276-284
 *   joint_weights [B,K]                   main.py:191-195 after data_parser.py:159-171
 *   lowconf [B,K] (uint8)                 fit_single_frame.py:285-287
 *   init_mask [B,K] (uint8)               fit_single_frame.py:289-294 (trimmed init joints)
 *   cam [B,16]: fx fy cx cy R[9] data_weight trans_est_z pad
 *   reg_pose [B,n_pose] or NULL           fit_single_frame.py:442 (regression_pose) */
int sfx_batch_set_targets(sfx_batch* b, const void* keypoints, const void* joint_weights,
                          const uint8_t* lowconf, const uint8_t* init_mask, const void* cam,
                          const void* reg_pose, void* stream);
/* Same, all pointers on the device (elements in the batch dtype). */
int sfx_batch_set_targets_dev(sfx_batch* b, const void* gt, const void* conf,
                              const void* joint_weights, const uint8_t* lowconf,
                              const uint8_t* init_mask, const void* cam, const void* reg_pose);

/* Parameter vectors [B,np] in the batch dtype (layout: sfx_batch_layout). */
int sfx_batch_set_params(sfx_batch* b, const void* host_params, void* stream);
int sfx_batch_get_params(const sfx_batch* b, void* host_params, void* stream);
void* sfx_batch_params_dev(sfx_batch* b);

/* fitting.guess_init (fitting.py:36-110) on the device, for the frames with need[b] != 0: camera
 * translation (0, 0, focal[b] * mean 3-D edge length / mean_len2d[b]) written into the frame's
 * parameter vector and into its depth-prior target (cam row, trans_estimation z,
 * fit_single_frame.py:404-411).  joints_dev: the mapped model joints [B,K,3] at the initial
 * parameters (sfx_eval); edge_idxs: n_edges pairs of keypoint indices (body_tri_idxs); focal,
 * mean_len2d, need: HOST arrays of B entries (copied before the call returns).  Double
 * arithmetic in the order of the host formula, so no host round trip is needed before the fit. */
int sfx_batch_guess_init(sfx_batch* b, const void* joints_dev, const int32_t* edge_idxs, int32_t n_edges,
                         const double* focal, const double* mean_len2d, const uint8_t* need, void* stream);

/* One closure evaluation for every frame: loss [B] and gradient [B,np] w.r.t. the whole
 * parameter vector.  Stands in for FittingMonitor.create_fitting_closure's fitting_func
 * (fitting.py:232-273): SMPL-X forward + SMPLifyLoss / SMPLifyCameraInitLoss + backward.
 * Also writes the mapped joints [B,K,3] when joints_dev != NULL. */
int sfx_eval(sfx_batch* b, const SfxStage* stage, void* loss_dev, void* grad_dev,
             void* joints_dev, void* stream);

/* FittingMonitor.run_fitting (fitting.py:147-217) with the optimiser of
 * optim_factory.create_optimizer (lbfgsls or adam) for every frame, entirely on the device:
 * one launch, no host synchronisation.  frame_ids_dev (int32 [n]) restricts the launch to a
 * subset (second orientation, fit_single_frame.py:527-538) or NULL for all frames.
 * final_loss_dev [B] receives run_fitting's return value per frame. */
int sfx_fit_stage(sfx_batch* b, const SfxStage* stage, const int32_t* frame_ids_dev,
                  int32_t n_ids, void* final_loss_dev, void* stream);

/* Orientation bookkeeping of fit_single_frame.py:527-551 on the device, between the camera
 * stage and the body stages.  Both stand in for body_model.reset_params(global_orient=orient,
 * body_pose=pose_embedding): every block except the orientation, the pose embedding and the
 * camera translation restarts from zero.
 *   flip = 0: remembers each selected frame's camera-stage orientation and starts from it;
 *   flip = 1: snapshots the first orientation's fitted parameters and loss, then starts from
 *             Rodrigues(saved orientation) . R_y(pi) (cv2.Rodrigues both ways, float32 result). */
int sfx_batch_begin_orientation(sfx_batch* b, int32_t flip, const int32_t* frame_ids_dev,
                                int32_t n_ids, void* stream);
/* results[argmin loss] (fit_single_frame.py:662-668): restores the snapshot of a frame when the
 * first orientation's loss is strictly lower than the second's. */
int sfx_batch_select_orientation(sfx_batch* b, const int32_t* frame_ids_dev, int32_t n_ids,
                                 void* stream);
/* run_fitting's return value of the last stage per frame ([B], batch dtype, device). */
void* sfx_batch_final_loss_dev(sfx_batch* b);

/* The whole per-frame flow of fit_single_frame.py:447-668 in ONE launch: camera stage, the
 * annealing stages, and for frames with flip_dev[f] != 0 (2-D shoulder distance below
 * side_view_thsh, :461-463) the second, 180-degree flipped orientation followed by the argmin
 * selection.  Persistent blocks pull frames in the order given by order_dev (int32 [B]; NULL =
 * 0..B-1; list the flipped frames first).  Results: parameters (sfx_batch_get_params), final
 * loss (sfx_batch_final_loss_dev), camera-stage loss (sfx_batch_cam_loss_dev), and the last
 * orientation's parameters for sfx_forward_mesh_last. */
int sfx_fit_pipeline(sfx_batch* b, const SfxPipeline* pipe, const int32_t* order_dev,
                     const uint8_t* flip_dev, void* stream);
void* sfx_batch_cam_loss_dev(sfx_batch* b);
/* [B][64] int64 cycle counters per frame (all zero unless the library was built with
 * -DSFX_CYCLE_PROF): 0 evaluations, 1 two-loop recursion, 2 blend forward, 3 blend adjoint,
 * 4 whole frame, 8..12 phases of the interpenetration term; 16..63 lap timers of the phases of an
 * evaluation and of the line search (profiles/prof_cycles.py names them). */
long long* sfx_batch_prof_dev(sfx_batch* b);
/* body_model(return_verts=True) at the LAST fitted orientation of every frame -- the mesh the
 * reference writes to vertices.ply (fit_single_frame.py:611, :671-676). */
int sfx_forward_mesh_last(sfx_batch* b, void* vertices_dev, void* joints_dev, void* stream);

/* body_model(return_verts=True) for every frame (fit_single_frame.py:611): full-mesh
 * vertices [B,V,3] and mapped joints [B,K,3] from the current parameters. */
int sfx_forward_mesh(sfx_batch* b, void* vertices_dev, void* joints_dev, void* stream);

/* Diagnostics: closure evaluations and status flags per frame ([B] int32, device). */
int32_t* sfx_batch_evals_dev(sfx_batch* b);
int32_t* sfx_batch_flags_dev(sfx_batch* b);
/* rows of the blend matrix (2 KiB each in float32) streamed per frame, forward + adjoint passes. */
int32_t* sfx_batch_passes_dev(sfx_batch* b);
int sfx_batch_reset_counters(sfx_batch* b, void* stream);

/* Evaluation metrics (reference utils.py:540-771, eval.py:14-44 compute_v2v): per-point error
 * sqrt(sum((aligned est - gt)^2)) between B fitted and ground-truth point sets [B,N,3] (device,
 * float32 or float64 per use_double) after an alignment:
 *   mode 0 none; 1 Procrustes / similarity transform (ProcrustesAlignment.__call__);
 *   2 pelvis: both sets minus the mean of their points hip0, hip1 (PelvisAlignment);
 *   3 scale + translation only (ScaleAlignment).
 * idx_dev (int32 [n]) selects a subset of the N points (eval.py `vids`) or is NULL with n == N;
 * hip0 / hip1 index the selected points.  err_dev [B,n]; transform_dev NULL or [B,13]: scale,
 * R[9] row major, t[3] applied to est. */
int sfx_aligned_errors(const void* est_dev, const void* gt_dev, const int32_t* idx_dev, int32_t B,
                       int32_t N, int32_t n, int32_t mode, int32_t hip0, int32_t hip1,
                       int32_t use_double, void* err_dev, void* transform_dev, void* stream);

/* Keypoint ingestion on the device (all pointers device memory, float32 detections).
 * sfx_pack_keypoints: read_keypoints of data_parser.py:57-104 without the JSON parsing -- the raw
 * OpenPose blocks body [B,n_body,3], hands [B,21,3] each and face [B,n_face >= 68,3] become
 * out [B,K,3], rows body | left hand | right hand | face[17:68] | face[0:17] (the last block only
 * with use_face_contour), K = n_body + 42 + 51 (+ 17). */
int sfx_pack_keypoints(const float* body_dev, const float* lhand_dev, const float* rhand_dev,
                       const float* face_dev, int32_t B, int32_t n_body, int32_t n_face,
                       int32_t use_face_contour, float* out_dev, void* stream);
/* fit_single_frame.py:276-294 for B frames: keypoints [B,K,3] -> gt [B,K,2], conf [B,K] and
 * joint weights [B,K] in the batch dtype (use_double), the low-confidence mask (conf <
 * confidence_threshold on the first n_body rows, < 0 elsewhere) and the trimmed camera-init
 * joints (listed, detected, not low-confidence): the arguments of sfx_batch_set_targets_dev.
 * base_joint_weights [K] = dataset.get_joint_weights() (data_parser.py:159-171). */
int sfx_keypoint_masks(const float* keypoints_dev, const float* base_joint_weights_dev,
                       const int32_t* init_joints_idxs_dev, int32_t n_init, int32_t n_body,
                       float confidence_threshold, int32_t B, int32_t K, int32_t use_double, void* gt_dev,
                       void* conf_dev, void* joint_weights_dev, uint8_t* lowconf_dev, uint8_t* init_mask_dev,
                       void* stream);
/* keypoints_blending.py:337-369 for B frames: OpenPose [B,135,3] and MMPose [B,136,3] detections
 * in raw order -> blended [B,135,3].  stats [4][n_pairs]: mmpose means, mmpose stds, openpose
 * means, openpose stds of the n_pairs (<= 67) keypoints either detector can supply, whose rows
 * are pair_mmpose / pair_openpose; face rows 67..134 come from OpenPose. */
int sfx_blend_keypoints(const float* openpose_dev, const float* mmpose_dev, const float* stats_dev,
                        const int32_t* pair_mmpose_dev, const int32_t* pair_openpose_dev, int32_t n_pairs,
                        int32_t B, float* out_dev, void* stream);

/* Diagnostic, no counterpart in the reference: measured L2 -> SM read bandwidth (GB/s) of the
 * current device -- every SM sweeping one L2-resident buffer of buffer_bytes, iters times.  The
 * benchmark reports the fit kernel's blend-row traffic (rows shared by all frames, served from
 * L2) against this figure.  Synchronous. */
int sfx_diag_l2_read_gbs(int64_t buffer_bytes, int32_t iters, double* gbs_out);

const char* sfx_last_error(void);
int sfx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SFX_H */
