// POD types shared by the CUDA kernels, the C-ABI layer and the host-side simulation build.
// Names follow the reference's domain: frames, stages, keypoints, joints, landmarks.
#pragma once
#include <stdint.h>

#define SFX_NJ 55            // SMPL-X skeleton joints
#define SFX_NPOSE 165        // full pose length (55 * 3)
#define SFX_NPF 486          // pose-corrective feature length (54 * 9)
#define SFX_KPAD 512         // padded contraction length: 486 pose features + <= 26 shape coeffs
#define SFX_NSHAPE_MAX 26
#define SFX_NEXTRA 21        // vertex-picked extra joints (nose, eyes, ears, feet, finger tips)
#define SFX_NLMK 51          // static face landmarks
#define SFX_NDYN 17          // dynamic face-contour landmarks
#define SFX_NDYNROWS 79      // rows of the yaw look-up table
#define SFX_NSTATIC (SFX_NEXTRA + 3 * SFX_NLMK)           // 174 static support-vertex slots
#define SFX_NSLOT (SFX_NSTATIC + 3 * SFX_NDYN)            // 225 support-vertex slots
#define SFX_NJOUT_MAX 144    // model joints before the joint mapper (127 without contour)
#define SFX_KMAX 160         // max keypoints after the joint mapper
#define SFX_NP_MAX 192       // max length of the per-frame parameter vector
#define SFX_HIST 100         // L-BFGS history (reference lbfgs_ls.py:200)
#define SFX_WROW 56          // padded row length of the dense skinning-weight table
#define SFX_NW 8             // non-zero skinning weights kept per support vertex (dense fall-back beyond)
#define SFX_MAX_BLOCKS 12    // parameter blocks of the optimised vector (for the gtol test)
#define SFX_NLATENT 32       // VPoser latent size
#define SFX_NSTREAM 15       // warps of a block that stream blend rows (warp 0 runs the kinematic chain)
#define SFX_NPART_MAX 64     // body parts of the face segmentation (smplx_parts_segm.pkl: 55)

// loss kinds (reference fitting.py:278-284)
#define SFX_LOSS_SMPLIFY 0
#define SFX_LOSS_CAMERA_INIT 1
// optimiser kinds (reference optimizers/optim_factory.py:27-65)
#define SFX_OPT_LBFGSLS 0
#define SFX_OPT_ADAM 1
// body-pose prior kinds used by SMPLifyLoss.forward (fitting.py:389-401)
#define SFX_PPRIOR_L2 0          // body_pose_prior = L2Prior on body_pose
#define SFX_PPRIOR_REGRESSION 1  // || pose_embedding - regression_pose ||^2
#define SFX_PPRIOR_GMM 2         // MaxMixturePrior
#define SFX_PPRIOR_LATENT 3      // VPoser: || z ||^2 (or || z - z_reg ||^2 in the last stage)

// Layout of the per-frame parameter vector (all optimisable quantities of one frame).
struct SfxLayout {
    int n_betas, n_expr, n_hand, n_pose;     // n_pose = 63 (axis-angle) or 32 (VPoser latent)
    int off_betas, off_go, off_lh, off_rh, off_jaw, off_leye, off_reye, off_expr, off_pose,
        off_camt;
    int np;                                   // total length
};

// One optimisation stage (camera stage or one annealing stage) -- reference
// fit_single_frame.py:450-496 (camera) and :553-596 (body stages).
struct SfxStage {
    int loss_kind;            // SFX_LOSS_*
    int opt_kind;             // SFX_OPT_*
    int pprior_kind;          // SFX_PPRIOR_*
    int stage_index, num_stages;
    int use_joints_conf;      // fitting.py:380-382
    int use_conf_camera;      // fitting.py:509-511 (the double-unsqueeze broadcast)
    int use_vposer;
    int n_body_kpts;          // NUM_BODY_JOINTS of the keypoint format (25 / 26 / 23)
    // --- weights (fitting.py:341-359; they enter squared except bending and coll) ---
    double rho;
    double body_pose_weight, shape_weight, bending_prior_weight, hand_prior_weight,
        expr_prior_weight, jaw_prior_weight[3], hand_joint_weight, face_joint_weight,
        depth_loss_weight;
    // --- run_fitting (fitting.py:147-217) ---
    int maxiters;
    double ftol, gtol;
    // --- optimiser (lbfgs_ls.py:199-207; adam: optim_factory.py:45-48) ---
    double lr;
    int max_iter, max_eval, history;
    double tol_grad, tol_change;
    double adam_beta1, adam_beta2, adam_eps;
    // --- optimised subset of the parameter vector ---
    int n_active;                              // D
    int n_blocks;
    int block_start[SFX_MAX_BLOCKS];           // into the compact vector
    int block_len[SFX_MAX_BLOCKS];
    int block_off[SFX_MAX_BLOCKS];             // into the full parameter vector
    int need_blend_grad;                       // 0 when only global_orient / camera are optimised
    int generic_two_loop;                      // L-BFGS direction: 0 the reference's recursion, one warp (default);
                                               // 1 the same by the whole block (A/B test); 2 the same recursion on
                                               // the inner products of the history (gram_two_loop: same mathematics,
                                               // different rounding, shorter dependent chain)
    // --- interpenetration term (fitting.py:437-455; third-party mesh_intersection) ---
    double coll_loss_weight;                   // enters unsquared (fitting.py:453-455); 0 = term off
    double coll_sigma;                         // df_cone_height (fit_single_frame.py:310)
};

// The whole per-frame flow of fit_single_frame.py:447-668 as one launch: camera stage, then the
// annealing stages for the first orientation and, for frames flagged for it, for the flipped one.
#define SFX_MAX_STAGES 8
struct SfxPipeline {
    int n_stages;
    int n_wide;               // "wide" frames: the first n_wide frames of the launch order are each run by a
                              // cluster of 8 CTAs whose helpers keep the live blend rows resident in shared
                              // memory (csrc/sfx_stream.cuh).  0 = every frame by one block (default).  Float32
                              // only; the adjoint's summation order differs from the one-block path.
    SfxStage cam;
    SfxStage body[SFX_MAX_STAGES];
};

// Per-frame constants, struct-of-arrays over frames.  "cam" row: fx fy cx cy R[9] data_weight
// trans_est_z pad (16 values).
#define SFX_CAM_STRIDE 16
#define SFX_CAM_FX 0
#define SFX_CAM_FY 1
#define SFX_CAM_CX 2
#define SFX_CAM_CY 3
#define SFX_CAM_R 4
#define SFX_CAM_DW 13
#define SFX_CAM_TZ 14

// per-frame status flags written by the fit kernel
#define SFX_FLAG_NAN 1
#define SFX_FLAG_INF 2
#define SFX_FLAG_COLL_OVERFLOW 4     // candidate or touched-vertex list of the interpenetration term was truncated
