// Core of the per-frame fitting engine: SMPL-X sparse-support forward pass, loss stack and
// analytic adjoint, strong-Wolfe L-BFGS / Adam and the run_fitting loop -- one CUDA block per
// frame, the whole optimisation stage inside one kernel launch (no host synchronisation).
//
// The same source compiles twice:
//   * nvcc, device pass: the product (threads of one block cooperate; SFX_FOR strides by
//     blockDim.x, SFX_SYNC is __syncthreads);
//   * g++, single-threaded "host simulation" used ONLY by tests/hostsim to debug the maths and
//     the control flow without a GPU.  It is never linked into the product library.
//
// Reference behaviour restated here (file:line in /root/reference/smplifyx unless noted):
//   SMPL-X forward ............ smplx 0.1.x lbs.py / body_models.py [third party, un-vendored]
//   PerspectiveCamera.forward . camera.py:93-117
//   GMoF ...................... utils.py:84-95
//   SMPLifyLoss.forward ....... fitting.py:375-461
//   SMPLifyCameraInitLoss ..... fitting.py:499-520
//   SMPLifyAnglePrior, L2Prior  prior.py:53-97
//   LBFGS.step / strong Wolfe . optimizers/lbfgs_ls.py:39-167, :256-445
//   FittingMonitor.run_fitting  fitting.py:147-217
#pragma once
#include <math.h>
#include "sfx_types.h"

#ifdef __CUDACC__
#define SFX_FN __device__ __forceinline__
#define SFX_MFN __device__ __forceinline__
#define SFX_FN_NOINLINE __device__ __noinline__
#define SFX_TID ((int)threadIdx.x)
#define SFX_NT ((int)blockDim.x)
#define SFX_SYNC() __syncthreads()
#else
#define SFX_FN static inline
#define SFX_MFN inline
#define SFX_FN_NOINLINE static
#define SFX_TID 0
#define SFX_NT 1
#define SFX_SYNC() ((void)0)
#endif
#if defined(__CUDACC__) && defined(SFX_CYCLE_PROF)
#define SFX_PROF_BEGIN(name) long long _t_##name = clock64()
#define SFX_PROF_END(S, slot, name) do { if (threadIdx.x == 0) (S).prof[slot] += clock64() - _t_##name; } while (0)
// lap timer: the time since the previous lap goes to slot `i` (thread 0's view)
#define SFX_LAP(S, i) do { if (threadIdx.x == 0) { long long _n = clock64(); (S).lap[i] += _n - (S).lap_t; (S).lap_t = _n; } } while (0)
#define SFX_NLAP 48
#else
#define SFX_PROF_BEGIN(name) ((void)0)
#define SFX_PROF_END(S, slot, name) ((void)0)
#define SFX_LAP(S, i) ((void)0)
#endif
// Address-space hint: the per-frame working set, the staged model tables and the evaluation
// context all live in shared memory, but the out-of-line functions below only see generic
// references; without the hint every access is a generic LD/ST that re-derives the shared window
// (measured on the two-loop recursion: 42 -> 25 us per evaluation from shared addressing alone).
#ifdef __CUDACC__
#define SFX_ASSUME_SHARED(ref) __builtin_assume(__isShared((const void*)&(ref)))
#define SFX_ASSUME_SHARED_PTR(p) __builtin_assume(__isShared((const void*)(p)))
#else
#define SFX_ASSUME_SHARED(ref) ((void)0)
#define SFX_ASSUME_SHARED_PTR(p) ((void)0)
#endif
#define SFX_FOR(i, n) for (int i = SFX_TID; i < (n); i += SFX_NT)
// same, the iteration space starting at thread `first` (spreads independent loops of one phase
// over different warps)
#define SFX_FOR_FROM(i, n, first) \
    for (int i = (SFX_TID + SFX_NT - ((first) % SFX_NT)) % SFX_NT; i < (n); i += SFX_NT)
#ifdef __CUDACC__
#define SFX_IS_WARP0 (threadIdx.x < 32)
#define SFX_ON_THREAD(t) ((int)threadIdx.x == ((t) % (int)blockDim.x))
#define SFX_LANE_FOR(i, n) for (int i = (int)(threadIdx.x & 31); i < (n); i += 32)
#define SFX_SYNCWARP() __syncwarp()
#else
#define SFX_IS_WARP0 true
#define SFX_ON_THREAD(t) true
#define SFX_LANE_FOR(i, n) for (int i = 0; i < (n); ++i)
#define SFX_SYNCWARP() ((void)0)
#endif

namespace sfx {

// ------------------------------------------------------------------------------ views
// Support tables of the 51 dynamic-contour slots for ONE row of the yaw look-up table, exactly as
// support_slots() + support_by_joint() build them: prepared once per model for all 79 rows, so an
// evaluation whose head yaw crossed into another row copies 5 KB instead of rebuilding them
// (that rebuild was ~6 % of the average frame's time).
#define SFX_NDYNSLOT (3 * SFX_NDYN)
template <typename T>
struct DynRowPack {
    int vid[SFX_NDYNSLOT];
    int jtd_ptr[SFX_NJ + 1];                 // offsets into jt_slot / jt_w (first entry SFX_NSTATIC * SFX_NW)
    int overflow, pad;
    T bary[SFX_NDYNSLOT];
    T vt_s[SFX_NDYNSLOT * 3];
    float ww[SFX_NDYNSLOT * SFX_NW];
    float jt_w[SFX_NDYNSLOT * SFX_NW];
    unsigned char wn[SFX_NDYNSLOT + 1];
    unsigned char wj[SFX_NDYNSLOT * SFX_NW];
    unsigned char jt_slot[SFX_NDYNSLOT * SFX_NW];
};

template <typename T>
struct ModelView {
    int V, NS, NB, NE, NH, K, NJOUT, use_contour, n_neck, nlev;
    int wide_model;      // float64 arrays in a float64 model: the per-frame float32 copies (skinning weights
                         // by slot / by joint, hand components, mean pose) would truncate them, so the
                         // evaluation reads the model's own tables (dense-weight path of every frame)
    const T* PK;         // [3V][SFX_KPAD]  row 3v+c: 486 pose-corrective dirs | NS shape dirs | 0
    const T* vt;         // [3V]            template
    const T* J0;         // [165]           J_regressor . v_template
    const T* JS;         // [165][32]       J_regressor . shapedirs
    const T* Wd;         // [V][SFX_WROW]   dense skinning weights
    const int* sk_ptr;            // [V + 1]  the same weights as per-vertex lists (CSR)
    const unsigned char* sk_j;    // [nnz]    joint
    const T* sk_w;                // [nnz]    weight
    const T* hand_l;     // [NH][45]
    const T* hand_r;     // [NH][45]
    const T* pose_mean;  // [165]
    const int* sv_vid;   // [SFX_NSTATIC]   vertex id of every static support slot
    const T* lmk_bary;   // [51*3]
    const int* dyn_vid;  // [79][51]        vertex ids of the contour landmarks per yaw row
    const T* dyn_bary;   // [79][51]
    const DynRowPack<T>* dyn_pack;   // [79] (one entry without the contour) prepared support tables, or nullptr
    const int* joint_map;   // [K]  keypoint -> model joint (JointMapper, utils.py:68-81)
    const int* inv_ptr;     // [NJOUT+1]  model joint -> keypoints (CSR)
    const int* inv_idx;     // [K]
    // MaxMixturePrior on the body pose (prior.py:100-231); gmm_M = 0 when not set
    int gmm_M, gmm_D;
    const T* gmm_means;     // [M][D]
    const T* gmm_prec;      // [M][D][D]
    const T* gmm_logw;      // [M]   log(nll_weights)
    // VPoser v1 decoder (human_body_prior vposer_smpl.py); vp_ready = 0 when not set
    int vp_ready;
    const T* vp_w1;         // [512][32]   bodyprior_dec_fc1.weight
    const T* vp_b1;         // [512]
    const T* vp_w2;         // [512][512]  bodyprior_dec_fc2.weight
    const T* vp_b2;         // [512]
    const T* vp_w3;         // [128][512]  bodyprior_dec_out.weight (126 rows + 2 zero rows)
    const T* vp_b3;         // [128]
    // interpenetration term (sfx_collide.cuh); coll_ready = 0 until sfx_model_set_collision
    int coll_ready, F, n_parts;
    const int* faces;                      // [F][3]
    const int* part_ptr;                   // [n_parts + 1]  faces grouped by body part (CSR)
    const int* part_faces;                 // [F]            face ids, along the part's long axis
    int n_clusters;
    const int* cl_ptr;                     // [n_clusters + 1] runs of <= 64 consecutive part_faces entries
    const int* part_cl_ptr;                // [n_parts + 1]  clusters of every part
    const unsigned char* face_part;        // [F]            part of every face
    const unsigned long long* part_allow;  // [SFX_NPART_MAX] bit q of word p: FilterFaces keeps (p, q)
    const int* vf_ptr;                     // [V + 1]        vertex -> incident faces (CSR)
    const int* vf_idx;                     // [3F]           face * 4 + corner
    int parents[SFX_NJ];
    int order[SFX_NJ];      // joints sorted by depth
    int level_off[16];
    int child_off[SFX_NJ + 1];
    int child_idx[SFX_NJ];
    int neck[8];            // neck kinematic chain (neck -> root)
};

template <typename T>
struct BatchView {
    int B;
    SfxLayout lay;
    T* params;           // [B][np]       in/out
    const T* gt;         // [B][K][2]
    const T* conf;       // [B][K]
    const T* jw_base;    // [B][K]   joint weights after joints_to_ign / low-confidence zeroing
    const unsigned char* lowconf;    // [B][K]
    const unsigned char* init_mask;  // [B][K]  camera-init joints (trimmed)
    const T* cam;        // [B][16]
    const T* reg_pose;   // [B][n_pose] or nullptr
    T* hist_s;           // [B][HIST][SFX_NP_MAX]
    T* hist_y;           // [B][HIST][SFX_NP_MAX]
    T* gram;             // [B][2][HIST][HIST]  inner products of the history (Gram two-loop)
    T* final_loss;       // [B]
    int* n_evals;        // [B]  (accumulated)
    int* n_passes;       // [B]  rows of the blend matrix streamed (forward + adjoint passes)
    int* flags;          // [B]
    const int* frame_ids;   // optional indirection (nullptr: block b -> frame b)
    long long* prof;        // [B][16] cycle counters (builds with -DSFX_CYCLE_PROF only)
    // interpenetration term: global workspace, one slot per block (nullptr until
    // sfx_batch_enable_collisions).  Values: vp[3V] | vert[3V] | dvert[3V] | dtri[9F] | box[6F] | tri[9F];
    // indices: tv[V] | face[F]
    T* coll_vals;
    unsigned short* coll_idx;
    int* coll_stat;         // [B][4] per frame, largest over its evaluations: candidates, touched vertices,
                            // sweep iterations and listed partners of warp 0 (1/16 of the candidates)
    long coll_vals_stride, coll_idx_stride;
};

// per-frame working set (shared memory on the device)
template <typename T>
struct Scratch {
    T x[SFX_NP_MAX];          // full parameter vector
    T gfull[SFX_NP_MAX];      // gradient wrt the full parameter vector
    T fp[SFX_NPOSE], dfp[SFX_NPOSE];
    T hand[90];
    T shape[32], dshape[32];
    T R[SFX_NJ * 9], Rw[SFX_NJ * 9], tw[SFX_NJ * 3], Jr[SFX_NJ * 3], rel[SFX_NJ * 3];
    T dRw[SFX_NJ * 9], dtw[SFX_NJ * 3], dR[SFX_NJ * 9], dJ[SFX_NJ * 3], drel[SFX_NJ * 3];
    T A[SFX_NJ * 12], dA[SFX_NJ * 12];
    T c[SFX_KPAD], dc[SFX_KPAD];
    T vp[SFX_NSLOT * 3], vert[SFX_NSLOT * 3], dvert[SFX_NSLOT * 3], dvp[SFX_NSLOT * 3];
    T Trot[SFX_NSLOT * 9];
    int vid[SFX_NSLOT];
    T bary[SFX_NSLOT];
    T X[SFX_NJOUT_MAX * 3], dX[SFX_NJOUT_MAX * 3];
    T kl[SFX_KMAX];           // per-keypoint data term
    T dp[SFX_KMAX * 3];       // per-keypoint dL/d(camera-space point)
    T jw[SFX_KMAX];           // per-stage effective joint weights
    // optimiser vectors (compact, length D)
    T xa[SFX_NP_MAX], g[SFX_NP_MAX], d[SFX_NP_MAX], prev_g[SFX_NP_MAX], x0[SFX_NP_MAX];
    T g_prev[SFX_NP_MAX], bg0[SFX_NP_MAX], bg1[SFX_NP_MAX], q[SFX_NP_MAX], gl[SFX_NP_MAX];
    T m1[SFX_NP_MAX], m2[SFX_NP_MAX];   // Adam moments
    T al[SFX_HIST], ro[SFX_HIST];
    T sg[SFX_HIST], yg[SFX_HIST], cf[SFX_HIST];   // Gram two-loop: s_i.g, y_i.g, second-loop coefficients
    int act[SFX_NP_MAX];      // compact index -> full parameter index
    T red[12];
    T confsq;                 // camera stage: sum of squared confidences of the init joints
    T vt_s[SFX_NSLOT * 3];    // template position of the support rows
    float ww[SFX_NSLOT * SFX_NW], jt_w[SFX_NSLOT * SFX_NW];    // skinning weights by slot / by joint
    unsigned char wj[SFX_NSLOT * SFX_NW], jt_slot[SFX_NSLOT * SFX_NW], wn[SFX_NSLOT];
    int jt_ptr[SFX_NJ + 1];   // by-joint table of the static slots ...
    int jtd_ptr[SFX_NJ + 1];  // ... and of the dynamic contour slots (entries from NSTATIC * NW on)
    int w_overflow, dynrow_cached;
    T bp[64];                 // VPoser: decoded body pose (63) and its gradient
    T dbp[64];
    T o6[128], do6[128];      // VPoser: 6-D rotation outputs and their gradient
    unsigned char vm1[512], vm2[512]; // VPoser: leaky-ReLU masks of the two hidden layers
    T tl_red[64];             // two-loop recursion: per-lane partial sums (double buffered)
#if defined(__CUDACC__) && defined(SFX_CYCLE_PROF)
    long long lap[SFX_NLAP], lap_t;    // fine-grained lap timers (profiling build only)
#endif
    long long prof[16];       // cycle counters: 0 eval, 1 two-loop, 2 blend fwd, 3 blend adj, 4 total,
                              // 8 mesh skinning, 9 part boxes + candidates, 10 narrow phase + penalty,
                              // 11 per-vertex gather, 12 skinning adjoint of touched vertices
    T gq[16];                 // mixture prior: per-component negative log-likelihood
    // interpenetration term: ordered-compaction state, touched-vertex count
    float hand_c[2 * 12 * 45];    // hand PCA components of both hands (when there are <= 12 of them);
    float pose_mean_c[SFX_NPOSE]; // the model stores them as float32, so a float copy is exact
    int hand_cached;
    unsigned short rows[SFX_NSLOT * 3];   // support rows the stage needs (slots with a live keypoint),
                                          // grouped by the streaming warp that owns them (row % 15)
    unsigned short wptr[SFX_NSTREAM + 1]; // start of every warp's group in rows[]
    unsigned char slot_live[SFX_NSLOT];
    int n_rows;
    int rows_dirty;           // wide frames: the helpers' resident copy of the live rows is stale
    int cscan[2][32];
    int cscan_total, cscan_calls, coll_overflow, n_touch, coll_max_cand, coll_max_touch, coll_max_iters, coll_max_hits;
    T coll_loss;
    T loss;
    int dynrow;
    int n_evals;
    int n_passes;
};

// ------------------------------------------------------------------------------ math
SFX_FN float sfx_sin(float x) { return sinf(x); }
SFX_FN double sfx_sin(double x) { return sin(x); }
SFX_FN float sfx_cos(float x) { return cosf(x); }
SFX_FN double sfx_cos(double x) { return cos(x); }
SFX_FN float sfx_sqrt(float x) { return sqrtf(x); }
SFX_FN double sfx_sqrt(double x) { return sqrt(x); }
SFX_FN float sfx_exp(float x) { return expf(x); }
SFX_FN double sfx_exp(double x) { return exp(x); }
SFX_FN float sfx_atan2(float y, float x) { return atan2f(y, x); }
SFX_FN double sfx_atan2(double y, double x) { return atan2(y, x); }
SFX_FN float sfx_rint(float x) { return rintf(x); }
SFX_FN double sfx_rint(double x) { return rint(x); }
SFX_FN float sfx_abs(float x) { return fabsf(x); }
SFX_FN double sfx_abs(double x) { return fabs(x); }

// C = A * B (3x3, row major)
template <typename T>
SFX_FN void mat3_mul(const T* a, const T* b, T* c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

// Rodrigues as smplx.lbs.batch_rodrigues writes it: angle = ||r + 1e-8||, dir = r / angle.
template <typename T>
SFX_FN void rodrigues(const T* r, T* R) {
    const T e = (T)1e-8;
    T a0 = r[0] + e, a1 = r[1] + e, a2 = r[2] + e;
    T th = sfx_sqrt(a0 * a0 + a1 * a1 + a2 * a2);
    T x = r[0] / th, y = r[1] / th, z = r[2] / th;
    T s = sfx_sin(th), c1 = (T)1 - sfx_cos(th);
    // K = [0 -z y; z 0 -x; -y x 0],  K^2 = dir dir^T - |dir|^2 I
    T xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = (T)1 + c1 * (-(yy + zz));
    R[1] = -s * z + c1 * xy;
    R[2] = s * y + c1 * xz;
    R[3] = s * z + c1 * xy;
    R[4] = (T)1 + c1 * (-(xx + zz));
    R[5] = -s * x + c1 * yz;
    R[6] = -s * y + c1 * xz;
    R[7] = s * x + c1 * yz;
    R[8] = (T)1 + c1 * (-(xx + yy));
}

// Adjoint of rodrigues(): G = dL/dR (row major 3x3) -> dr = dL/dr.
template <typename T>
SFX_FN void rodrigues_bwd(const T* r, const T* G, T* dr) {
    const T e = (T)1e-8;
    T a0 = r[0] + e, a1 = r[1] + e, a2 = r[2] + e;
    T th = sfx_sqrt(a0 * a0 + a1 * a1 + a2 * a2);
    T inv = (T)1 / th;
    T x = r[0] * inv, y = r[1] * inv, z = r[2] * inv;
    T s = sfx_sin(th), c = sfx_cos(th), c1 = (T)1 - c;
    T K[9] = {0, -z, y, z, 0, -x, -y, x, 0};
    T K2[9];
    mat3_mul(K, K, K2);
    // dL/dtheta = <G, cos K + sin K^2>
    T dth = 0;
    for (int i = 0; i < 9; ++i) dth += G[i] * (c * K[i] + s * K2[i]);
    // dL/dK = sin G + (1 - cos)(G K^T + K^T G)
    T GKt[9], KtG[9], Kt[9] = {K[0], K[3], K[6], K[1], K[4], K[7], K[2], K[5], K[8]};
    mat3_mul(G, Kt, GKt);
    mat3_mul(Kt, G, KtG);
    T dK[9];
    for (int i = 0; i < 9; ++i) dK[i] = s * G[i] + c1 * (GKt[i] + KtG[i]);
    T dx = dK[7] - dK[5], dy = dK[2] - dK[6], dz = dK[3] - dK[1];
    // dir = r / th ; th = ||r + e||
    T dot = dx * r[0] + dy * r[1] + dz * r[2];
    T k = dth * inv - dot * inv * inv * inv;
    dr[0] = dx * inv + k * a0;
    dr[1] = dy * inv + k * a1;
    dr[2] = dz * inv + k * a2;
}

// ------------------------------------------------------------------------ reductions
// Fixed-order reductions so a frame's trajectory is reproducible run to run: lane l of warp 0
// accumulates elements l, l+32, ... then an xor butterfly combines the 32 partials.
template <typename T, typename F, typename Op>
SFX_FN T block_reduce(int n, F f, Op op, T init, T* slot) {
    SFX_SYNC();
#ifdef __CUDACC__
    if (threadIdx.x < 32) {
        T p = init;
        for (int i = threadIdx.x; i < n; i += 32) p = op(p, f(i));
        for (int o = 16; o > 0; o >>= 1) p = op(p, __shfl_xor_sync(0xffffffffu, p, o));
        if (threadIdx.x == 0) *slot = p;
    }
#else
    T part[32];
    for (int l = 0; l < 32; ++l) {
        T p = init;
        for (int i = l; i < n; i += 32) p = op(p, f(i));
        part[l] = p;
    }
    for (int o = 16; o > 0; o >>= 1) {
        T nxt[32];
        for (int l = 0; l < 32; ++l) nxt[l] = op(part[l], part[l ^ o]);
        for (int l = 0; l < 32; ++l) part[l] = nxt[l];
    }
    *slot = part[0];
#endif
    SFX_SYNC();
    return *slot;
}

template <typename T>
struct OpAdd { SFX_MFN T operator()(T a, T b) const { return a + b; } };

template <typename T>
struct OpMax { SFX_MFN T operator()(T a, T b) const { return a > b ? a : b; } };

// Two reductions over the same index range in one pair of barriers: warp 0 takes the first,
// warp 1 the second, each exactly as block_reduce() would (same partial sums, same butterfly).
template <typename T, typename F0, typename Op0, typename F1, typename Op1>
SFX_FN void block_reduce2(int n, F0 f0, Op0 op0, T init0, F1 f1, Op1 op1, T init1, T* slots) {
#ifdef __CUDACC__
    SFX_SYNC();
    if (threadIdx.x < 64) {
        const int lane = threadIdx.x & 31;
        if (threadIdx.x < 32) {
            T p = init0;
            for (int i = lane; i < n; i += 32) p = op0(p, f0(i));
            for (int o = 16; o > 0; o >>= 1) p = op0(p, __shfl_xor_sync(0xffffffffu, p, o));
            if (lane == 0) slots[0] = p;
        } else {
            T p = init1;
            for (int i = lane; i < n; i += 32) p = op1(p, f1(i));
            for (int o = 16; o > 0; o >>= 1) p = op1(p, __shfl_xor_sync(0xffffffffu, p, o));
            if (lane == 0) slots[1] = p;
        }
    }
    SFX_SYNC();
#else
    block_reduce<T>(n, f0, op0, init0, slots);
    block_reduce<T>(n, f1, op1, init1, slots + 1);
#endif
}

template <typename T>
SFX_FN T block_dot(const T* a, const T* b, int n, T* slot) {
    return block_reduce<T>(n, [=](int i) { return a[i] * b[i]; }, OpAdd<T>(), (T)0, slot);
}
template <typename T>
SFX_FN T block_absmax(const T* a, int n, T* slot) {
    return block_reduce<T>(n, [=](int i) { return sfx_abs(a[i]); }, OpMax<T>(), (T)0, slot);
}
template <typename T>
SFX_FN T block_abssum(const T* a, int n, T* slot) {
    return block_reduce<T>(n, [=](int i) { return sfx_abs(a[i]); }, OpAdd<T>(), (T)0, slot);
}

// ------------------------------------------------------------------- blend streaming
// vp[r] = vt[row] + PK[row] . c   for the 3 rows of every support slot     (forward)
// dc[k] = sum_r PK[row][k] * dvp[r]                                         (adjoint)
// These two passes move 225*3*512 values each and dominate the evaluation; the device
// version lives in sfx_stream.cuh (TMA bulk copies into per-warp shared-memory rings).
#ifdef __CUDACC__
template <typename T>
__device__ __forceinline__ void blend_forward(const ModelView<T>& M, Scratch<T>& S, void* wsp);
template <typename T>
__device__ __forceinline__ void blend_adjoint(const ModelView<T>& M, Scratch<T>& S, void* wsp);
template <typename T>
__device__ __forceinline__ void rows_dot(const T* W, const T* bias, int nrows, const T* x, T* y, void* wsp);
template <typename T>
__device__ __forceinline__ void rows_accum(const T* W, int nrows, const T* g, T* out, void* wsp);
template <typename T>
__device__ __forceinline__ void blend_forward_full(const ModelView<T>& M, Scratch<T>& S, void* wsp, T* vp_g);
template <typename T>
__device__ __forceinline__ void blend_adjoint_ext(const ModelView<T>& M, Scratch<T>& S, void* wsp,
                                                  const unsigned short* tv, const T* dvp_g);
template <typename T>
struct Scratch;
template <typename T>
__device__ __forceinline__ bool two_loop_staged(Scratch<T>& S, int k, int head, int H, T hd,
                                                const T* hist_s, const T* hist_y, int D, void* wsp);
// "wide" frames (a cluster of CTAs per frame, sfx_stream.cuh): the helpers keep the stage's live
// rows resident in their shared memory and run both blend passes on them
__device__ __forceinline__ bool wide_active(void* wsp);
template <typename T>
__device__ __forceinline__ void wide_begin(Scratch<T>& S, void* wsp, int cmd);
template <typename T>
__device__ __forceinline__ void wide_end_forward(Scratch<T>& S, void* wsp);
template <typename T>
__device__ __forceinline__ void wide_end_adjoint(Scratch<T>& S, void* wsp);
#define SFX_WIDE_EXIT 0
#define SFX_WIDE_LOAD 1
#define SFX_WIDE_FWD 2
#define SFX_WIDE_ADJ 3
// shared work area that is idle between evaluations (the blend ring), and its release
template <typename T>
__device__ __forceinline__ unsigned char* idle_area(void* wsp, size_t* bytes);
__device__ __forceinline__ void idle_area_release(void* wsp);
// float32 Gram chain over a [SFX_GRAM_ROWS][SFX_GRAM_LDF] zero-padded block in shared memory
#define SFX_GRAM_ROWS 128
#define SFX_GRAM_LDF 129
__device__ __forceinline__ void gram_chain_f32(Scratch<float>& S, int k, float hd, const float* G, int lane);
#else
template <typename T>
static void rows_dot(const T* W, const T* bias, int nrows, const T* x, T* y, void*) {
    for (int r = 0; r < nrows; ++r) {
        T acc = 0;
        for (int k = 0; k < SFX_KPAD; ++k) acc += W[(long)r * SFX_KPAD + k] * x[k];
        y[r] = bias[r] + acc;
    }
}
template <typename T>
static void rows_accum(const T* W, int nrows, const T* g, T* out, void*) {
    for (int k = 0; k < SFX_KPAD; ++k) out[k] = 0;
    for (int r = 0; r < nrows; ++r)
        for (int k = 0; k < SFX_KPAD; ++k) out[k] += W[(long)r * SFX_KPAD + k] * g[r];
}
template <typename T>
static void blend_forward(const ModelView<T>& M, Scratch<T>& S, void*) {
    for (int i = 0; i < S.n_rows; ++i) {
        const int r = S.rows[i];
        long row = (long)S.vid[r / 3] * 3 + (r % 3);
        const T* p = M.PK + row * SFX_KPAD;
        T acc = 0;
        for (int k = 0; k < SFX_KPAD; ++k) acc += p[k] * S.c[k];
        S.vp[r] = S.vt_s[r] + acc;
    }
}
template <typename T>
static void blend_adjoint(const ModelView<T>& M, Scratch<T>& S, void*) {
    for (int k = 0; k < SFX_KPAD; ++k) S.dc[k] = 0;
    for (int i = 0; i < S.n_rows; ++i) {
        const int r = S.rows[i];
        long row = (long)S.vid[r / 3] * 3 + (r % 3);
        const T* p = M.PK + row * SFX_KPAD;
        T w = S.dvp[r];
        for (int k = 0; k < SFX_KPAD; ++k) S.dc[k] += p[k] * w;
    }
}
// every row of the blend matrix: vp_g[r] = vt[r] + PK[r] . c   (interpenetration term)
template <typename T>
static void blend_forward_full(const ModelView<T>& M, Scratch<T>& S, void*, T* vp_g) {
    for (long r = 0; r < 3L * M.V; ++r) {
        const T* p = M.PK + r * SFX_KPAD;
        T acc = 0;
        for (int k = 0; k < SFX_KPAD; ++k) acc += p[k] * S.c[k];
        vp_g[r] = M.vt[r] + acc;
    }
}
// support rows followed by the rows of the touched vertices tv (weights dvpc[3 t + k])
template <typename T>
static void blend_adjoint_ext(const ModelView<T>& M, Scratch<T>& S, void* w, const unsigned short* tv,
                              const T* dvpc) {
    blend_adjoint(M, S, w);
    for (int t = 0; t < S.n_touch; ++t)
        for (int c = 0; c < 3; ++c) {
            const long row = 3L * tv[t] + c;
            const T* p = M.PK + row * SFX_KPAD;
            const T wgt = dvpc[3 * t + c];
            for (int k = 0; k < SFX_KPAD; ++k) S.dc[k] += p[k] * wgt;
        }
}
#endif

// --------------------------------------------------------------------- evaluation
struct FrameConst {          // per-frame constants decoded from the BatchView "cam" row
    double fx, fy, cx, cy, Rc[9], dw, tz_est;
};

// ------------------------------------------------------------------ support-vertex tables
// The loss reads the mesh through <= 225 "support" vertex slots: 21 extra-joint vertices, 51 x 3
// static landmark triangle corners and 17 x 3 dynamic-contour corners (these change with the yaw
// row of the contour look-up table).  Per frame the kernel keeps, in shared memory, each slot's
// vertex id, barycentric weight, template position and its non-zero skinning weights both by slot
// (forward) and by joint (adjoint), so an evaluation touches global memory only for the blend rows.
template <typename T>
SFX_FN void support_slots(const ModelView<T>& M, Scratch<T>& S, int s_begin, int s_end) {
    SFX_FOR(i, s_end - s_begin) {
        const int s = s_begin + i;
        int vid;
        T b;
        if (s < SFX_NSTATIC) {
            vid = M.sv_vid[s];
            b = s < SFX_NEXTRA ? (T)1 : M.lmk_bary[s - SFX_NEXTRA];
        } else if (M.use_contour) {
            vid = M.dyn_vid[S.dynrow * 51 + s - SFX_NSTATIC];
            b = M.dyn_bary[S.dynrow * 51 + s - SFX_NSTATIC];
        } else {
            vid = M.sv_vid[0];
            b = 0;
        }
        S.vid[s] = vid;
        S.bary[s] = b;
        for (int k = 0; k < 3; ++k) S.vt_s[3 * s + k] = M.vt[3L * vid + k];
        // the vertex's non-zero skinning weights, joints ascending (per-vertex lists of the model)
        const int e0 = M.sk_ptr[vid], n = M.sk_ptr[vid + 1] - e0;
        for (int e = 0; e < n && e < SFX_NW; ++e) {
            S.wj[s * SFX_NW + e] = M.sk_j[e0 + e];
            S.ww[s * SFX_NW + e] = (float)M.sk_w[e0 + e];  // model weights are float32: exact
        }
        S.wn[s] = (unsigned char)(n <= SFX_NW ? n : SFX_NW);
        if (n > SFX_NW) S.w_overflow = 1;                  // dense fall-back for this frame
    }
}

// By-joint transpose of the per-slot lists of slots [s_begin, s_end) (entries of a joint in
// ascending slot order), written from `base` on in jt_slot / jt_w with offsets ptr[0 .. NJ].
// Two such tables exist: the static slots (built once per frame) and the dynamic contour slots
// (rebuilt whenever the head's yaw crosses into another row of the look-up table -- every few
// evaluations during a line search, which is why it is kept small).
template <typename T>
SFX_FN void support_by_joint(Scratch<T>& S, int s_begin, int s_end, int* ptr, int base) {
    SFX_SYNC();
    SFX_FOR(j, SFX_NJ) {
        int cnt = 0;
        for (int s = s_begin; s < s_end; ++s)
            for (int e = 0; e < S.wn[s]; ++e) cnt += S.wj[s * SFX_NW + e] == j;
        ptr[j + 1] = cnt;
    }
    SFX_SYNC();
    if (SFX_TID == 0) {
        ptr[0] = base;
        for (int j = 0; j < SFX_NJ; ++j) ptr[j + 1] += ptr[j];
    }
    SFX_SYNC();
    SFX_FOR(j, SFX_NJ) {
        int pos = ptr[j];
        for (int s = s_begin; s < s_end; ++s)
            for (int e = 0; e < S.wn[s]; ++e)
                if (S.wj[s * SFX_NW + e] == j) {
                    S.jt_slot[pos] = (unsigned char)s;
                    S.jt_w[pos] = S.ww[s * SFX_NW + e];
                    ++pos;
                }
    }
    SFX_SYNC();
}

// The dynamic slots' tables of one yaw row from the model's prepared copy (same contents as
// support_slots + support_by_joint over [NSTATIC, NSLOT)).  Ends synchronised.
template <typename T>
SFX_FN void support_load_dyn_row(const DynRowPack<T>& P, Scratch<T>& S) {
    SFX_FOR(i, SFX_NDYNSLOT) {
        S.vid[SFX_NSTATIC + i] = P.vid[i];
        S.bary[SFX_NSTATIC + i] = P.bary[i];
        S.wn[SFX_NSTATIC + i] = P.wn[i];
    }
    SFX_FOR_FROM(i, SFX_NDYNSLOT * 3, 64) S.vt_s[3 * SFX_NSTATIC + i] = P.vt_s[i];
    SFX_FOR_FROM(i, SFX_NJ + 1, 256) S.jtd_ptr[i] = P.jtd_ptr[i];
    SFX_FOR(i, SFX_NDYNSLOT * SFX_NW) {
        S.wj[SFX_NSTATIC * SFX_NW + i] = P.wj[i];
        S.ww[SFX_NSTATIC * SFX_NW + i] = P.ww[i];
        S.jt_slot[SFX_NSTATIC * SFX_NW + i] = P.jt_slot[i];
        S.jt_w[SFX_NSTATIC * SFX_NW + i] = P.jt_w[i];
    }
    if (SFX_TID == 0 && P.overflow) S.w_overflow = 1;
    SFX_SYNC();
}

// small model constants every evaluation reads (hand PCA components, mean pose): kept next to the data
template <typename T>
SFX_FN void frame_constants(const ModelView<T>& M, Scratch<T>& S) {
    if (SFX_TID == 0) S.hand_cached = M.NH <= 12 && !M.wide_model;
    if (M.NH <= 12 && !M.wide_model)
        SFX_FOR(i, 2 * M.NH * 45)
            S.hand_c[i] = (float)(i < M.NH * 45 ? M.hand_l[i] : M.hand_r[i - M.NH * 45]);
    SFX_FOR(i, SFX_NPOSE) S.pose_mean_c[i] = (float)M.pose_mean[i];
}

// once per frame (and per kernel launch): static slots; the dynamic ones follow the yaw row
template <typename T>
SFX_FN void support_begin_frame(const ModelView<T>& M, Scratch<T>& S) {
    SFX_SYNC();
    if (SFX_TID == 0) {
        S.w_overflow = M.wide_model ? 1 : 0;
        S.dynrow_cached = -1;
    }
    SFX_SYNC();
    frame_constants(M, S);
    support_slots(M, S, 0, SFX_NSTATIC);
    support_by_joint(S, 0, SFX_NSTATIC, S.jt_ptr, 0);
}


// ------------------------------------------------------------------ VPoser v1 decoder
// z [32] -> FC 512 -> leaky-ReLU(0.2) -> FC 512 -> leaky-ReLU(0.2) -> FC 126 -> per joint: 6-D
// continuous rotation -> Gram-Schmidt -> rotation matrix -> quaternion (torchgeometry's branchy
// conversion) -> axis-angle.  Third-party algorithm (human_body_prior, torchgeometry), restated
// from the published code; reference call site: fitting.py:236.
SFX_FN float sfx_max(float a, float b) { return a > b ? a : b; }
SFX_FN double sfx_max(double a, double b) { return a > b ? a : b; }

// one joint: o[6] -> aa[3]; keeps nothing (the adjoint recomputes the forward values)
template <typename T>
struct Rot6 {
    T b1[3], b2[3], b3[3], n1, n2, dt, u[3];
    T rt[9];            // transposed rotation matrix (torchgeometry works on R^T)
    int kase;
    T t, qraw[4], q[4], ss, sn, k, tt;
};

template <typename T>
SFX_FN void rot6_forward(const T* o, Rot6<T>& F, T* aa) {
    const T eps = (T)1e-12;
    T a1[3] = {o[0], o[2], o[4]}, a2[3] = {o[1], o[3], o[5]};
    F.n1 = sfx_max(sfx_sqrt(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), eps);
    for (int k = 0; k < 3; ++k) F.b1[k] = a1[k] / F.n1;
    F.dt = F.b1[0] * a2[0] + F.b1[1] * a2[1] + F.b1[2] * a2[2];
    for (int k = 0; k < 3; ++k) F.u[k] = a2[k] - F.dt * F.b1[k];
    F.n2 = sfx_max(sfx_sqrt(F.u[0] * F.u[0] + F.u[1] * F.u[1] + F.u[2] * F.u[2]), eps);
    for (int k = 0; k < 3; ++k) F.b2[k] = F.u[k] / F.n2;
    F.b3[0] = F.b1[1] * F.b2[2] - F.b1[2] * F.b2[1];
    F.b3[1] = F.b1[2] * F.b2[0] - F.b1[0] * F.b2[2];
    F.b3[2] = F.b1[0] * F.b2[1] - F.b1[1] * F.b2[0];
    // R[r][c] = (b1 b2 b3)[c][r];  rt = R^T: rt[i][j] = R[j][i] -> rt row i is b_i
    for (int k = 0; k < 3; ++k) { F.rt[k] = F.b1[k]; F.rt[3 + k] = F.b2[k]; F.rt[6 + k] = F.b3[k]; }
    const T* m = F.rt;
    const bool d2 = m[8] < (T)1e-6, d0d1 = m[0] > m[4], d0nd1 = m[0] < -m[4];
    if (d2 && d0d1) {
        F.kase = 0; F.t = (T)1 + m[0] - m[4] - m[8];
        F.qraw[0] = m[5] - m[7]; F.qraw[1] = F.t; F.qraw[2] = m[1] + m[3]; F.qraw[3] = m[6] + m[2];
    } else if (d2) {
        F.kase = 1; F.t = (T)1 - m[0] + m[4] - m[8];
        F.qraw[0] = m[6] - m[2]; F.qraw[1] = m[1] + m[3]; F.qraw[2] = F.t; F.qraw[3] = m[5] + m[7];
    } else if (d0nd1) {
        F.kase = 2; F.t = (T)1 - m[0] - m[4] + m[8];
        F.qraw[0] = m[1] - m[3]; F.qraw[1] = m[6] + m[2]; F.qraw[2] = m[5] + m[7]; F.qraw[3] = F.t;
    } else {
        F.kase = 3; F.t = (T)1 + m[0] + m[4] + m[8];
        F.qraw[0] = F.t; F.qraw[1] = m[5] - m[7]; F.qraw[2] = m[6] - m[2]; F.qraw[3] = m[1] - m[3];
    }
    const T rs = sfx_sqrt(F.t);
    for (int i = 0; i < 4; ++i) F.q[i] = F.qraw[i] / rs * (T)0.5;
    F.ss = F.q[1] * F.q[1] + F.q[2] * F.q[2] + F.q[3] * F.q[3];
    F.sn = sfx_sqrt(F.ss);
    const T c = F.q[0];
    F.tt = (T)2 * (c < (T)0 ? sfx_atan2(-F.sn, -c) : sfx_atan2(F.sn, c));
    F.k = F.ss > (T)0 ? F.tt / F.sn : (T)2;
    for (int i = 0; i < 3; ++i) aa[i] = F.q[1 + i] * F.k;
}

template <typename T>
SFX_FN void rot6_backward(const Rot6<T>& F, const T* g, T* dO) {
    // angle-axis <- quaternion
    T dq[4] = {0, g[0] * F.k, g[1] * F.k, g[2] * F.k};
    if (F.ss > (T)0) {
        const T dk = g[0] * F.q[1] + g[1] * F.q[2] + g[2] * F.q[3];
        const T dtt = dk / F.sn;
        T ds = -dk * F.tt / (F.sn * F.sn);
        const T c = F.q[0], den = F.sn * F.sn + c * c;
        ds += dtt * (T)2 * c / den;
        dq[0] = dtt * (T)2 * (-F.sn) / den;
        const T dss = ds / ((T)2 * F.sn);
        for (int i = 1; i < 4; ++i) dq[i] += (T)2 * F.q[i] * dss;
    }
    // q = 0.5 qraw / sqrt(t)
    const T rs = sfx_sqrt(F.t);
    T dqr[4], dt = 0;
    for (int i = 0; i < 4; ++i) {
        dqr[i] = dq[i] * (T)0.5 / rs;
        dt += -dq[i] * (T)0.25 * F.qraw[i] / (F.t * rs);
    }
    T dm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (F.kase == 0) {
        dt += dqr[1];
        dm[5] += dqr[0]; dm[7] -= dqr[0]; dm[1] += dqr[2]; dm[3] += dqr[2]; dm[6] += dqr[3]; dm[2] += dqr[3];
        dm[0] += dt; dm[4] -= dt; dm[8] -= dt;
    } else if (F.kase == 1) {
        dt += dqr[2];
        dm[6] += dqr[0]; dm[2] -= dqr[0]; dm[1] += dqr[1]; dm[3] += dqr[1]; dm[5] += dqr[3]; dm[7] += dqr[3];
        dm[0] -= dt; dm[4] += dt; dm[8] -= dt;
    } else if (F.kase == 2) {
        dt += dqr[3];
        dm[1] += dqr[0]; dm[3] -= dqr[0]; dm[6] += dqr[1]; dm[2] += dqr[1]; dm[5] += dqr[2]; dm[7] += dqr[2];
        dm[0] -= dt; dm[4] -= dt; dm[8] += dt;
    } else {
        dt += dqr[0];
        dm[5] += dqr[1]; dm[7] -= dqr[1]; dm[6] += dqr[2]; dm[2] -= dqr[2]; dm[1] += dqr[3]; dm[3] -= dqr[3];
        dm[0] += dt; dm[4] += dt; dm[8] += dt;
    }
    // rt rows are b1, b2, b3
    T db1[3] = {dm[0], dm[1], dm[2]}, db2[3] = {dm[3], dm[4], dm[5]}, db3[3] = {dm[6], dm[7], dm[8]};
    // b3 = b1 x b2
    db1[0] += F.b2[1] * db3[2] - F.b2[2] * db3[1];
    db1[1] += F.b2[2] * db3[0] - F.b2[0] * db3[2];
    db1[2] += F.b2[0] * db3[1] - F.b2[1] * db3[0];
    db2[0] += db3[1] * F.b1[2] - db3[2] * F.b1[1];
    db2[1] += db3[2] * F.b1[0] - db3[0] * F.b1[2];
    db2[2] += db3[0] * F.b1[1] - db3[1] * F.b1[0];
    // b2 = u / |u|
    const T pb2 = F.b2[0] * db2[0] + F.b2[1] * db2[1] + F.b2[2] * db2[2];
    T du[3], da2[3], da1[3];
    for (int k = 0; k < 3; ++k) du[k] = (db2[k] - F.b2[k] * pb2) / F.n2;
    // u = a2 - dt b1 ; dt = b1 . a2
    const T ddt = -(du[0] * F.b1[0] + du[1] * F.b1[1] + du[2] * F.b1[2]);
    T a2[3];
    for (int k = 0; k < 3; ++k) a2[k] = F.u[k] + F.dt * F.b1[k];
    for (int k = 0; k < 3; ++k) {
        da2[k] = du[k] + ddt * F.b1[k];
        db1[k] += -F.dt * du[k] + ddt * a2[k];
    }
    // b1 = a1 / |a1|
    const T pb1 = F.b1[0] * db1[0] + F.b1[1] * db1[1] + F.b1[2] * db1[2];
    for (int k = 0; k < 3; ++k) da1[k] = (db1[k] - F.b1[k] * pb1) / F.n1;
    dO[0] = da1[0]; dO[2] = da1[1]; dO[4] = da1[2];
    dO[1] = da2[0]; dO[3] = da2[1]; dO[5] = da2[2];
}

// z = S.x[off_pose ..] -> S.bp[63].  h1 lives in S.dc, h2 in S.dvp (both free before the blend
// passes); only the leaky-ReLU masks and the 6-D outputs are kept for the adjoint.
template <typename T>
SFX_FN void vposer_decode(const ModelView<T>& M, const T* z, Scratch<T>& S, void* wsp) {
    SFX_SYNC();
    SFX_FOR(i, 512) {
        const T* w = M.vp_w1 + i * SFX_NLATENT;
        T acc = 0;
        for (int k = 0; k < SFX_NLATENT; ++k) acc += w[k] * z[k];
        acc += M.vp_b1[i];
        S.vm1[i] = acc > (T)0;
        S.dc[i] = acc > (T)0 ? acc : (T)0.2 * acc;
    }
    SFX_SYNC();
    rows_dot(M.vp_w2, M.vp_b2, 512, S.dc, S.dvp, wsp);
    SFX_SYNC();
    SFX_FOR(i, 512) {
        const T v = S.dvp[i];
        S.vm2[i] = v > (T)0;
        S.dvp[i] = v > (T)0 ? v : (T)0.2 * v;
    }
    SFX_SYNC();
    rows_dot(M.vp_w3, M.vp_b3, 128, S.dvp, S.o6, wsp);
    SFX_SYNC();
    SFX_FOR(j, 21) {
        Rot6<T> F;
        rot6_forward(S.o6 + 6 * j, F, S.bp + 3 * j);
    }
    SFX_SYNC();
}

// S.dbp[63] (gradient wrt the decoded pose) -> dz[32] (written to out).  Uses S.c / S.dc as
// 512-wide temporaries (free at this point of the evaluation).
template <typename T>
SFX_FN void vposer_adjoint(const ModelView<T>& M, Scratch<T>& S, T* out, void* wsp) {
    SFX_SYNC();
    SFX_FOR(j, 128 / 6 + 1) {
        if (j < 21) {
            Rot6<T> F;
            T aa[3];
            rot6_forward(S.o6 + 6 * j, F, aa);
            rot6_backward(F, S.dbp + 3 * j, S.do6 + 6 * j);
        } else {
            S.do6[126] = 0;
            S.do6[127] = 0;
        }
    }
    SFX_SYNC();
    rows_accum(M.vp_w3, 128, S.do6, S.c, wsp);                 // dL/dh2
    SFX_SYNC();
    SFX_FOR(i, 512) S.c[i] = S.vm2[i] ? S.c[i] : (T)0.2 * S.c[i];
    SFX_SYNC();
    rows_accum(M.vp_w2, 512, S.c, S.dc, wsp);                  // dL/dh1
    SFX_SYNC();
    SFX_FOR(i, 512) S.dc[i] = S.vm1[i] ? S.dc[i] : (T)0.2 * S.dc[i];
    SFX_SYNC();
    // dz[k] = sum_i W1[i][k] g1[i]: 16 partial sums per k in a fixed order
    SFX_FOR(t, 512) {
        const int k = t & 31, part = t >> 5;
        T acc = 0;
        for (int i = part; i < 512; i += 16) acc += M.vp_w1[i * SFX_NLATENT + k] * S.dc[i];
        S.c[t] = acc;
    }
    SFX_SYNC();
    SFX_FOR(k, SFX_NLATENT) {
        T acc = 0;
        for (int part = 0; part < 16; ++part) acc += S.c[part * 32 + k];
        out[k] = acc;
    }
    SFX_SYNC();
}

// ------------------------------------------------------------------ pose prologue
// Full pose (hand PCA + mean), Rodrigues, rest joints from the shape, blend coefficients, yaw row
// of the contour table.  Reads S.x.  Ends synchronised.
template <typename T>
SFX_FN void pose_prologue(const ModelView<T>& M, const SfxLayout& L, Scratch<T>& S,
                          bool use_vposer = false, void* wsp = nullptr, bool support_tables = true) {
    const int NS = M.NS;
    if (use_vposer) vposer_decode(M, S.x + L.off_pose, S, wsp);     // body pose = decode(z)
    SFX_FOR(i, SFX_NPOSE) {
        T v;
        if (i < 3) v = S.x[L.off_go + i];
        else if (i < 66) v = use_vposer ? S.bp[i - 3] : S.x[L.off_pose + i - 3];
        else if (i < 69) v = S.x[L.off_jaw + i - 66];
        else if (i < 72) v = S.x[L.off_leye + i - 69];
        else if (i < 75) v = S.x[L.off_reye + i - 72];
        else {
            int h = (i - 75) / 45, e = (i - 75) % 45;
            const T* pc = S.x + (h ? L.off_rh : L.off_lh);
            T acc = 0;
            if (S.hand_cached) {
                const float* C = S.hand_c + h * L.n_hand * 45;
                for (int k = 0; k < L.n_hand; ++k) acc += pc[k] * (T)C[k * 45 + e];
            } else {
                const T* C = h ? M.hand_r : M.hand_l;
                for (int k = 0; k < L.n_hand; ++k) acc += pc[k] * C[k * 45 + e];
            }
            S.hand[h * 45 + e] = acc;
            v = acc;
        }
        S.fp[i] = v + (M.wide_model ? M.pose_mean[i] : (T)S.pose_mean_c[i]);
    }
    SFX_FOR_FROM(i, 32, 192) {
        T v = 0;
        if (i < L.n_betas) v = S.x[L.off_betas + i];
        else if (i < L.n_betas + L.n_expr) v = S.x[L.off_expr + i - L.n_betas];
        S.shape[i] = v;
    }
    SFX_SYNC();
    SFX_LAP(S, 1);
    // joint rotations (threads 0..54), rest joints (threads 64..228), yaw row (thread 256)
    SFX_FOR(j, SFX_NJ) rodrigues(S.fp + 3 * j, S.R + 9 * j);
    SFX_FOR_FROM(i, SFX_NJ * 3, 64) {
        T acc = M.J0[i];
        const T* js = M.JS + i * 32;
        for (int s = 0; s < NS; ++s) acc += js[s] * S.shape[s];
        S.Jr[i] = acc;
    }
    if (SFX_ON_THREAD(256)) {
        // dynamic-contour look-up row (smplx lbs.find_dynamic_lmk_idx_and_bcoords)
        int row = 0;
        if (M.use_contour) {
            T rel[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tmp[9], Rn[9];
            for (int i = 0; i < M.n_neck; ++i) {
                rodrigues(S.fp + 3 * M.neck[i], Rn);
                mat3_mul(Rn, rel, tmp);
                for (int k = 0; k < 9; ++k) rel[k] = tmp[k];
            }
            T sy = sfx_sqrt(rel[0] * rel[0] + rel[3] * rel[3]);
            T ang = sfx_atan2(-rel[6], sy);
            T deg = (-ang * (T)180.0) / (T)3.14159265358979323846;
            if (deg > (T)39) deg = (T)39;
            long y = (long)sfx_rint(deg);
            if (y < 0) y = (y < -39) ? 78 : (39 - y);
            row = (int)y;
        }
        S.dynrow = row;
    }
    SFX_SYNC();
    SFX_LAP(S, 2);
    // blend coefficients: pose feature (R_j - I for j >= 1) | shape | 0
    SFX_FOR(i, SFX_KPAD) {
        T v = 0;
        if (i < SFX_NPF) {
            int e = i % 9;
            v = S.R[9 + i] - ((e == 0 || e == 4 || e == 8) ? (T)1 : (T)0);
        } else if (i < SFX_NPF + NS) {
            v = S.shape[i - SFX_NPF];
        }
        S.c[i] = v;
    }
    SFX_FOR(i, SFX_NJ * 3) {
        int j = i / 3, p = M.parents[j];
        S.rel[i] = p < 0 ? S.Jr[i] : S.Jr[i] - S.Jr[3 * p + (i % 3)];
    }
    if (support_tables && S.dynrow != S.dynrow_cached) {   // uniform: dynrow was written before the barrier
        SFX_SYNC();
        if (M.dyn_pack) {
            support_load_dyn_row(M.dyn_pack[M.use_contour ? S.dynrow : 0], S);
        } else {
            support_slots(M, S, SFX_NSTATIC, SFX_NSLOT);
            support_by_joint(S, SFX_NSTATIC, SFX_NSLOT, S.jtd_ptr, SFX_NSTATIC * SFX_NW);
        }
        if (SFX_TID == 0) {
            S.dynrow_cached = S.dynrow;
            S.rows_dirty = 1;
        }
    }
    SFX_SYNC();
    SFX_LAP(S, 3);
}

// Kinematic chain, skinning transforms A and the 55 posed skeleton joints -- one warp, level by
// level with warp-level synchronisation only (the rest of the block streams blend rows meanwhile).
// Operands are staged in registers: the shared-memory arrays may alias as far as the compiler
// knows, which would serialise every load behind the previous store.
template <typename T>
SFX_FN void chain_forward(const ModelView<T>& M, Scratch<T>& S) {
    for (int lv = 0; lv < M.nlev; ++lv) {
        int a = M.level_off[lv], b = M.level_off[lv + 1];
        SFX_LANE_FOR(i, b - a) {
            const int j = M.order[a + i], p = M.parents[j];
            T Rj[9], out[9], t[3];
            for (int k = 0; k < 9; ++k) Rj[k] = S.R[9 * j + k];
            if (p < 0) {
                for (int k = 0; k < 9; ++k) out[k] = Rj[k];
                for (int k = 0; k < 3; ++k) t[k] = S.Jr[3 * j + k];
            } else {
                T Rp[9], rl[3], tp[3];
                for (int k = 0; k < 9; ++k) Rp[k] = S.Rw[9 * p + k];
                for (int k = 0; k < 3; ++k) { rl[k] = S.rel[3 * j + k]; tp[k] = S.tw[3 * p + k]; }
                mat3_mul(Rp, Rj, out);
                for (int k = 0; k < 3; ++k)
                    t[k] = Rp[3 * k] * rl[0] + Rp[3 * k + 1] * rl[1] + Rp[3 * k + 2] * rl[2] + tp[k];
            }
            for (int k = 0; k < 9; ++k) S.Rw[9 * j + k] = out[k];
            for (int k = 0; k < 3; ++k) S.tw[3 * j + k] = t[k];
            // A_j = [Rw | tw - Rw J_rest], posed joint = tw
            T J[3], Aj[12];
            for (int k = 0; k < 3; ++k) J[k] = S.Jr[3 * j + k];
            for (int r = 0; r < 3; ++r) {
                Aj[4 * r] = out[3 * r];
                Aj[4 * r + 1] = out[3 * r + 1];
                Aj[4 * r + 2] = out[3 * r + 2];
                Aj[4 * r + 3] = t[r] - (out[3 * r] * J[0] + out[3 * r + 1] * J[1] + out[3 * r + 2] * J[2]);
            }
            for (int k = 0; k < 12; ++k) S.A[12 * j + k] = Aj[k];
            for (int k = 0; k < 3; ++k) S.X[3 * j + k] = t[k];
        }
        SFX_SYNCWARP();
    }
}

// Adjoint of chain_forward (one warp): dA, dX[0..55) -> dR (chain part), drel, dJ (A part).
// Three passes, so that only what a parent needs from its children sits on the level-by-level
// chain: (1) every joint's own terms, all joints side by side; (2) deepest level first, the joints
// WITH children add their children's terms (own terms first, children in table order -- the order
// of the sums is that of the one-pass formulation); (3) dR / drel of every joint from its
// finished dRw / dtw, all joints side by side.
template <typename T>
SFX_FN void chain_adjoint(const ModelView<T>& M, Scratch<T>& S) {
    // ---- (1) own terms from A_j and the posed joint ----
    SFX_LANE_FOR(j, SFX_NJ) {
        T dA[12], J[3], Rwj[9];
        for (int k = 0; k < 12; ++k) dA[k] = S.dA[12 * j + k];
        for (int k = 0; k < 3; ++k) J[k] = S.Jr[3 * j + k];
        for (int k = 0; k < 9; ++k) Rwj[k] = S.Rw[9 * j + k];
        for (int r = 0; r < 3; ++r) {
            const T dat = dA[4 * r + 3];
            for (int k = 0; k < 3; ++k) S.dRw[9 * j + 3 * r + k] = dA[4 * r + k] - dat * J[k];
            S.dtw[3 * j + r] = dat + S.dX[3 * j + r];
        }
        for (int k = 0; k < 3; ++k)
            S.dJ[3 * j + k] = -(Rwj[k] * dA[3] + Rwj[3 + k] * dA[7] + Rwj[6 + k] * dA[11]);
    }
    SFX_SYNCWARP();
    // ---- (2) children (complete: they sit one level deeper) ----
    for (int lv = M.nlev - 2; lv >= 0; --lv) {
        int a = M.level_off[lv], b = M.level_off[lv + 1];
        SFX_LANE_FOR(i, b - a) {
            const int j = M.order[a + i];
            const int e0 = M.child_off[j], e1 = M.child_off[j + 1];
            if (e0 < e1) {
                T dRw[9], dtw[3];
                for (int k = 0; k < 9; ++k) dRw[k] = S.dRw[9 * j + k];
                for (int k = 0; k < 3; ++k) dtw[k] = S.dtw[3 * j + k];
                for (int e = e0; e < e1; ++e) {
                    const int ch = M.child_idx[e];
                    T dRc[9], Rch[9], dtc[3], rl[3];
                    for (int k = 0; k < 9; ++k) { dRc[k] = S.dRw[9 * ch + k]; Rch[k] = S.R[9 * ch + k]; }
                    for (int k = 0; k < 3; ++k) { dtc[k] = S.dtw[3 * ch + k]; rl[k] = S.rel[3 * ch + k]; }
                    for (int r = 0; r < 3; ++r)
                        for (int k = 0; k < 3; ++k)
                            dRw[3 * r + k] += dRc[3 * r] * Rch[3 * k] + dRc[3 * r + 1] * Rch[3 * k + 1] +
                                              dRc[3 * r + 2] * Rch[3 * k + 2] + dtc[r] * rl[k];
                    for (int r = 0; r < 3; ++r) dtw[r] += dtc[r];
                }
                for (int k = 0; k < 9; ++k) S.dRw[9 * j + k] = dRw[k];
                for (int k = 0; k < 3; ++k) S.dtw[3 * j + k] = dtw[k];
            }
        }
        SFX_SYNCWARP();
    }
    // ---- (3) dR (chain part) and drel ----
    SFX_LANE_FOR(j, SFX_NJ) {
        const int p = M.parents[j];
        T dRw[9], dtw[3];
        for (int k = 0; k < 9; ++k) dRw[k] = S.dRw[9 * j + k];
        for (int k = 0; k < 3; ++k) dtw[k] = S.dtw[3 * j + k];
        if (p < 0) {
            for (int k = 0; k < 9; ++k) S.dR[9 * j + k] = dRw[k];
            for (int k = 0; k < 3; ++k) S.drel[3 * j + k] = dtw[k];
        } else {
            T Rp[9];
            for (int k = 0; k < 9; ++k) Rp[k] = S.Rw[9 * p + k];
            for (int r = 0; r < 3; ++r)
                for (int k = 0; k < 3; ++k)
                    S.dR[9 * j + 3 * r + k] = Rp[r] * dRw[k] + Rp[3 + r] * dRw[3 + k] + Rp[6 + r] * dRw[6 + k];
            for (int r = 0; r < 3; ++r)
                S.drel[3 * j + r] = Rp[r] * dtw[0] + Rp[3 + r] * dtw[1] + Rp[6 + r] * dtw[2];
        }
    }
    SFX_SYNCWARP();
}

// several sums at once: warp q (q, q + nwarps, ...) reduces quantity q over i < n in the fixed
// lane order of block_reduce(); results in out[0..nq)
template <typename T, typename F>
SFX_FN void multi_sum(int nq, int n, F f, T* out) {
    SFX_SYNC();
#ifdef __CUDACC__
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int q = warp; q < nq; q += nw) {
        T p = 0;
        for (int i = lane; i < n; i += 32) p = p + f(q, i);
        for (int o = 16; o > 0; o >>= 1) p = p + __shfl_xor_sync(0xffffffffu, p, o);
        if (lane == 0) out[q] = p;
    }
#else
    for (int q = 0; q < nq; ++q) {
        T part[32];
        for (int l = 0; l < 32; ++l) {
            T p = 0;
            for (int i = l; i < n; i += 32) p = p + f(q, i);
            part[l] = p;
        }
        for (int o = 16; o > 0; o >>= 1) {
            T nxt[32];
            for (int l = 0; l < 32; ++l) nxt[l] = part[l] + part[l ^ o];
            for (int l = 0; l < 32; ++l) part[l] = nxt[l];
        }
        out[q] = part[0];
    }
#endif
    SFX_SYNC();
}

// MaxMixturePrior.merged_log_likelihood (prior.py:181-196) and its gradient for one frame:
// value = min_m [ 0.5 (x - mu_m)^T P_m (x - mu_m) - log w_m ],  grad = 0.5 (P_m* + P_m*^T)(x - mu_m*).
// Uses S.c (free after the blend passes) for the M x D products; returns the value, writes grad[D].
template <typename T>
SFX_FN T gmm_prior(const ModelView<T>& M, const T* pose, Scratch<T>& S, T* grad) {
    const int Mg = M.gmm_M, D = M.gmm_D;
    SFX_SYNC();
    SFX_FOR(t, Mg * D) {
        const int m = t / D;
        const T* row = M.gmm_prec + (long)t * D;
        const T* mu = M.gmm_means + m * D;
        T acc = 0;
        for (int j = 0; j < D; ++j) acc += row[j] * (pose[j] - mu[j]);
        S.c[t] = acc;
    }
    SFX_SYNC();
    SFX_FOR(m, Mg) {
        const T* mu = M.gmm_means + m * D;
        T q = 0;
        for (int i = 0; i < D; ++i) q += S.c[m * D + i] * (pose[i] - mu[i]);
        S.gq[m] = (T)0.5 * q - M.gmm_logw[m];
    }
    SFX_SYNC();
    int best = 0;
    for (int m = 1; m < Mg; ++m)
        if (S.gq[m] < S.gq[best]) best = m;           // torch.min keeps the first minimum
    const T val = S.gq[best];
    SFX_FOR(i, D) {
        const T* mu = M.gmm_means + best * D;
        const T* P = M.gmm_prec + (long)best * D * D;
        T acc = 0;
        for (int j = 0; j < D; ++j) acc += P[j * D + i] * (pose[j] - mu[j]);
        grad[i] = (T)0.5 * (S.c[best * D + i] + acc);
    }
    SFX_SYNC();
    return val;
}

}  // namespace sfx
#include "sfx_collide.cuh"
namespace sfx {

// Out-of-line entry points of the interpenetration term, so that the evaluation without the
// term keeps its register allocation.
template <typename T>
SFX_FN_NOINLINE void coll_blend_forward(const ModelView<T>& M, Scratch<T>& S, const CollWS<T>& W, void* ws) {
    SFX_ASSUME_SHARED(M);
    SFX_ASSUME_SHARED(S);
    SFX_ASSUME_SHARED(W);
    SFX_ASSUME_SHARED_PTR(ws);
    // the whole mesh is needed: stream every row once, the support rows are a subset
    blend_forward_full(M, S, ws, W.vp_g);
    SFX_SYNC();
    SFX_FOR(r, SFX_NSLOT * 3) S.vp[r] = W.vp_g[3L * S.vid[r / 3] + (r % 3)];
}
template <typename T>
SFX_FN_NOINLINE void coll_mesh_and_penalty(const ModelView<T>& M, Scratch<T>& S, const CollWS<T>& W,
                                           T sigma, T weight) {
    SFX_ASSUME_SHARED(M);
    SFX_ASSUME_SHARED(S);
    SFX_ASSUME_SHARED(W);
    SFX_PROF_BEGIN(cs);
    coll_skin_mesh(M, S, W);
    SFX_PROF_END(S, 8, cs);
    coll_search_and_penalty(M, S, W, sigma, weight);
}
template <typename T>
SFX_FN_NOINLINE void coll_blend_adjoint(const ModelView<T>& M, Scratch<T>& S, const CollWS<T>& W, void* ws) {
    SFX_ASSUME_SHARED(M);
    SFX_ASSUME_SHARED(S);
    SFX_ASSUME_SHARED(W);
    SFX_ASSUME_SHARED_PTR(ws);
    blend_adjoint_ext(M, S, ws, W.tv_g, W.vert_g);
}

// One evaluation of the stage objective and its gradient for one frame (the reference's
// closure, fitting.py:232-273).  Reads S.x, writes S.loss and S.gfull.  Needs
// support_begin_frame() once per frame and stage_joint_weights() once per stage beforehand.
template <typename T>
SFX_FN_NOINLINE void eval_frame(const ModelView<T>& M, const SfxLayout& L, const SfxStage& st,
                                const T* gt, const T* conf, const unsigned char* init_mask,
                                const T* cam, const T* reg_pose, Scratch<T>& S, void* stream_ws,
                                const CollWS<T>* CW = nullptr) {
    SFX_ASSUME_SHARED(M);
    SFX_ASSUME_SHARED(L);
    SFX_ASSUME_SHARED(st);
    SFX_ASSUME_SHARED(S);
    SFX_ASSUME_SHARED_PTR(stream_ws);
#ifdef __CUDACC__
    __builtin_assume(CW == nullptr || __isShared((const void*)CW));
#endif
    const int NS = M.NS;
    const int nj = M.NJOUT;
    const int K = M.K;
    SFX_PROF_BEGIN(eval);
    const bool vposer = st.use_vposer != 0;
    // interpenetration term: active in body stages with a positive weight (fitting.py:439)
    const bool coll = CW != nullptr && M.coll_ready && st.loss_kind == SFX_LOSS_SMPLIFY &&
                      st.coll_loss_weight > 0;
    SFX_PROF_BEGIN(pp);
    SFX_LAP(S, 0);
    pose_prologue(M, L, S, vposer, stream_ws);
    if (!coll) SFX_PROF_END(S, 8, pp);
    // ---- 3. kinematic chain (warp 0) || blendshapes on the support vertices (every warp) ----
    SFX_PROF_BEGIN(bf);
#ifdef __CUDACC__
    const bool wide = !coll && wide_active(stream_ws);
    if (wide) wide_begin(S, stream_ws, SFX_WIDE_FWD);       // helpers start on their resident rows
#else
    const bool wide = false;
#endif
    if (SFX_IS_WARP0) chain_forward(M, S);
    SFX_PROF_END(S, 5, bf);                     // chain alone (warp 0)
    SFX_PROF_BEGIN(bs);
    // rows no live keypoint depends on are not streamed: their slots stay at the template
    // position (finite; every term they enter carries a zero weight)
    SFX_FOR(r, SFX_NSLOT * 3)
        if (!S.slot_live[r / 3]) S.vp[r] = S.vt_s[r];
    if (coll) coll_blend_forward(M, S, *CW, stream_ws);
#ifdef __CUDACC__
    else if (wide) wide_end_forward(S, stream_ws);
#endif
    else blend_forward(M, S, stream_ws);
#if defined(__CUDACC__) && defined(SFX_CYCLE_PROF)
    if (threadIdx.x == 32) S.prof[6] += clock64() - _t_bs;      // warp 1's own streaming time
#endif
    SFX_SYNC();
    SFX_PROF_END(S, 2, bf);
    SFX_LAP(S, 4);
    if (coll) coll_mesh_and_penalty(M, S, *CW, (T)st.coll_sigma, (T)st.coll_loss_weight);
    SFX_PROF_BEGIN(mid);
    // ---- 4. skinning of the support vertices ------------------------------------------
    SFX_FOR(s, SFX_NSLOT) {
        T Tm[12];
        for (int k = 0; k < 12; ++k) Tm[k] = 0;
        if (!S.w_overflow) {
            for (int e = 0; e < S.wn[s]; ++e) {
                const T wj = (T)S.ww[s * SFX_NW + e];
                const T* A = S.A + 12 * S.wj[s * SFX_NW + e];
                for (int k = 0; k < 12; ++k) Tm[k] += wj * A[k];
            }
        } else {
            const T* w = M.Wd + (long)S.vid[s] * SFX_WROW;
            for (int j = 0; j < SFX_NJ; ++j) {
                T wj = w[j];
                if (wj != (T)0) {
                    const T* A = S.A + 12 * j;
                    for (int k = 0; k < 12; ++k) Tm[k] += wj * A[k];
                }
            }
        }
        const T* v = S.vp + 3 * s;
        for (int r = 0; r < 3; ++r) {
            S.vert[3 * s + r] = Tm[4 * r] * v[0] + Tm[4 * r + 1] * v[1] + Tm[4 * r + 2] * v[2] +
                                Tm[4 * r + 3];
            S.Trot[9 * s + 3 * r] = Tm[4 * r];
            S.Trot[9 * s + 3 * r + 1] = Tm[4 * r + 1];
            S.Trot[9 * s + 3 * r + 2] = Tm[4 * r + 2];
        }
    }
    SFX_SYNC();
    SFX_LAP(S, 5);
    // ---- 5. extra joints and landmarks ------------------------------------------------
    SFX_FOR(i, (nj - SFX_NJ) * 3) {
        int j = SFX_NJ + i / 3, k = i % 3;
        T v;
        if (j < SFX_NJ + SFX_NEXTRA) {
            v = S.vert[3 * (j - SFX_NJ) + k];
        } else {
            int s0 = SFX_NEXTRA + 3 * (j - SFX_NJ - SFX_NEXTRA);
            v = S.bary[s0] * S.vert[3 * s0 + k] + S.bary[s0 + 1] * S.vert[3 * (s0 + 1) + k] +
                S.bary[s0 + 2] * S.vert[3 * (s0 + 2) + k];
        }
        S.X[3 * j + k] = v;
    }
    SFX_SYNC();
    SFX_LAP(S, 6);
    // ---- 6. projection + data term per keypoint ---------------------------------------
    const T fx = cam[SFX_CAM_FX], fy = cam[SFX_CAM_FY], cx = cam[SFX_CAM_CX], cy = cam[SFX_CAM_CY];
    const T* Rc = cam + SFX_CAM_R;
    const T dw = cam[SFX_CAM_DW];
    const T dw2 = dw * dw;
    const T* ct = S.x + L.off_camt;
    const T rho2 = (T)(st.rho * st.rho);
    // fitting.py:510-511: conf [1,n,1,1]^2 broadcast against err [1,n,2] -> (sum conf^2)(sum err)
    const T confsq = (st.loss_kind == SFX_LOSS_CAMERA_INIT && st.use_conf_camera) ? S.confsq : (T)1;
    SFX_FOR(k, K) {
        const T* Xj = S.X + 3 * M.joint_map[k];
        T px = Rc[0] * Xj[0] + Rc[1] * Xj[1] + Rc[2] * Xj[2] + ct[0];
        T py = Rc[3] * Xj[0] + Rc[4] * Xj[1] + Rc[5] * Xj[2] + ct[1];
        T pz = Rc[6] * Xj[0] + Rc[7] * Xj[1] + Rc[8] * Xj[2] + ct[2];
        T ix = px / pz, iy = py / pz;
        T u = fx * ix + cx, v = fy * iy + cy;
        T ru = gt[2 * k] - u, rv = gt[2 * k + 1] - v;
        T lk, du, dv;       // loss of this keypoint, dL/du, dL/dv
        if (st.loss_kind == SFX_LOSS_CAMERA_INIT) {
            T w = init_mask[k] ? confsq * dw2 : (T)0;
            lk = w * (ru * ru + rv * rv);
            du = -(T)2 * w * ru;
            dv = -(T)2 * w * rv;
        } else {
            T wk = st.use_joints_conf ? S.jw[k] * conf[k] : S.jw[k];
            T w = wk * wk * dw2;
            T su = ru * ru, sv = rv * rv;
            T qu = su + rho2, qv = sv + rho2;
            lk = w * (rho2 * (su / qu) + rho2 * (sv / qv));
            // d/dr [rho2 r^2/(r^2+rho2)] = 2 r rho2^2 / (r^2+rho2)^2
            du = -w * ((T)2 * ru * rho2 * rho2 / (qu * qu));
            dv = -w * ((T)2 * rv * rho2 * rho2 / (qv * qv));
        }
        S.kl[k] = lk;
        // back through the projection
        T dix = du * fx, diy = dv * fy;
        T dpx = dix / pz, dpy = diy / pz;
        T dpz = -(dix * ix + diy * iy) / pz;
        S.dp[3 * k] = dpx;
        S.dp[3 * k + 1] = dpy;
        S.dp[3 * k + 2] = dpz;
    }
    SFX_LAP(S, 7);
    // data loss and camera-translation gradient (= sum_k dL/dp_k) and, in body stages, the five
    // sums of squares of the priors (pose, betas, expression, left hand, right hand; they only
    // need the parameters): nine sums, one pass, one pair of barriers
    const bool body = st.loss_kind == SFX_LOSS_SMPLIFY;
    const bool latent_reg = vposer && reg_pose != nullptr && st.stage_index + 1 == st.num_stages;
    {
        const T* kl = S.kl;
        const T* dp = S.dp;
        const T* pe = S.x + L.off_pose;
        const T* be = S.x + L.off_betas;
        const T* ex = S.x + L.off_expr;
        const T* hv = S.hand;
        const int n_pose = L.n_pose, n_betas = L.n_betas, n_expr = L.n_expr;
        const bool reg = vposer ? latent_reg : st.pprior_kind == SFX_PPRIOR_REGRESSION;
        multi_sum<T>(body ? 9 : 4, K > 64 ? K : 64, [=](int q, int i) -> T {
            if (q < 4) return i < K ? (q == 0 ? kl[i] : dp[3 * i + q - 1]) : (T)0;
            if (q == 4) {
                if (i >= n_pose) return (T)0;
                T d = reg ? pe[i] - reg_pose[i] : pe[i];
                return d * d;
            }
            if (q == 5) return i < n_betas ? be[i] * be[i] : (T)0;
            if (q == 6) return i < n_expr ? ex[i] * ex[i] : (T)0;
            if (q == 7) return i < 45 ? hv[i] * hv[i] : (T)0;
            return i < 45 ? hv[45 + i] * hv[45 + i] : (T)0;
        }, S.red);
    }
    const T data_loss = S.red[0];
    const T gct[3] = {S.red[1], S.red[2], S.red[3]};
    SFX_LAP(S, 8);
    // ---- 7. adjoint: keypoints -> model joints -> support vertices ---------------------
    SFX_FOR(i, nj * 3) {
        int j = i / 3, a = i % 3;
        T acc = 0;
        for (int e = M.inv_ptr[j]; e < M.inv_ptr[j + 1]; ++e) {
            const T* d = S.dp + 3 * M.inv_idx[e];
            acc += Rc[a] * d[0] + Rc[3 + a] * d[1] + Rc[6 + a] * d[2];     // Rc^T dp
        }
        S.dX[i] = acc;
    }
    SFX_SYNC();
    SFX_LAP(S, 9);
    SFX_FOR(i, SFX_NSLOT * 3) {
        int s = i / 3, k = i % 3;
        int j = s < SFX_NEXTRA ? SFX_NJ + s : SFX_NJ + SFX_NEXTRA + (s - SFX_NEXTRA) / 3;
        S.dvert[i] = j < nj ? S.bary[s] * S.dX[3 * j + k] : (T)0;
    }
    SFX_SYNC();
    SFX_LAP(S, 10);
    SFX_FOR(i, SFX_NSLOT * 3) {
        int s = i / 3, k = i % 3;
        const T* Tr = S.Trot + 9 * s;
        const T* dv = S.dvert + 3 * s;
        S.dvp[i] = Tr[k] * dv[0] + Tr[3 + k] * dv[1] + Tr[6 + k] * dv[2];
    }
    // dA[j][r][cc] = sum_s W[vid_s][j] * dvert[s][r] * (cc < 3 ? vp[s][cc] : 1)
    SFX_FOR_FROM(i, SFX_NJ * 12, 160) {
        int j = i / 12, r = (i % 12) / 4, cc = i % 4;
        T acc = 0;
        if (!S.w_overflow) {
            // static slots, then the dynamic ones: ascending slot order
            for (int e = S.jt_ptr[j]; e < S.jt_ptr[j + 1]; ++e) {
                const int s = S.jt_slot[e];
                acc += (T)S.jt_w[e] * S.dvert[3 * s + r] * (cc < 3 ? S.vp[3 * s + cc] : (T)1);
            }
            for (int e = S.jtd_ptr[j]; e < S.jtd_ptr[j + 1]; ++e) {
                const int s = S.jt_slot[e];
                acc += (T)S.jt_w[e] * S.dvert[3 * s + r] * (cc < 3 ? S.vp[3 * s + cc] : (T)1);
            }
        } else {
            for (int s = 0; s < SFX_NSLOT; ++s) {
                T w = M.Wd[(long)S.vid[s] * SFX_WROW + j];
                if (w != (T)0) acc += w * S.dvert[3 * s + r] * (cc < 3 ? S.vp[3 * s + cc] : (T)1);
            }
        }
        S.dA[i] = acc;
    }
    SFX_SYNC();
    SFX_LAP(S, 11);
    const bool coll_adj = coll && S.n_touch > 0;
    if (coll_adj) coll_skin_adjoint(M, S, *CW);
    if (!coll) SFX_PROF_END(S, 9, mid);
    // ---- 8. adjoint of the chain (warp 0) || adjoint of the blendshapes (every warp) -----
    SFX_PROF_BEGIN(ba);
#ifdef __CUDACC__
    const bool wide_adj = wide && !coll_adj && st.need_blend_grad;
    if (wide_adj) wide_begin(S, stream_ws, SFX_WIDE_ADJ);
#endif
    if (SFX_IS_WARP0) chain_adjoint(M, S);
    SFX_PROF_END(S, 7, ba);                     // chain adjoint alone (warp 0)
    if (st.need_blend_grad) {
        if (coll_adj) coll_blend_adjoint(M, S, *CW, stream_ws);
#ifdef __CUDACC__
        else if (wide_adj) wide_end_adjoint(S, stream_ws);
#endif
        else blend_adjoint(M, S, stream_ws);
    } else {
        SFX_FOR(i, SFX_KPAD) S.dc[i] = 0;
    }
    SFX_SYNC();
    SFX_PROF_END(S, 3, ba);
    SFX_LAP(S, 12);
    SFX_PROF_BEGIN(tail);
    // ---- 9. pose-feature gradient joins the chain's; Rodrigues adjoint; rest-joint adjoint --
    SFX_FOR(j, SFX_NJ) {
        if (j >= 1)
            for (int k = 0; k < 9; ++k) S.dR[9 * j + k] += S.dc[9 * (j - 1) + k];
        rodrigues_bwd(S.fp + 3 * j, S.dR + 9 * j, S.dfp + 3 * j);
    }
    // rel_j = J_j - J_parent(j)
    SFX_FOR_FROM(i, SFX_NJ * 3, 64) {
        int j = i / 3, k = i % 3;
        T acc = S.dJ[i] + S.drel[i];
        for (int e = M.child_off[j]; e < M.child_off[j + 1]; ++e) acc -= S.drel[3 * M.child_idx[e] + k];
        S.gl[i] = acc;                       // dL/dJ_rest (gl is rewritten after every evaluation)
    }
    SFX_SYNC();
    SFX_LAP(S, 13);
    // dshape[s] = dc[486 + s] + sum_i JS[i][s] dJ[i]: 16 partial sums per s, fixed order
    SFX_FOR(t, 512) {
        const int sidx = t & 31, part = t >> 5;
        T acc = 0;
        if (sidx < NS) {
            constexpr int NT = (SFX_NJ * 3 + 15) / 16;     // terms per thread (global loads: all in flight)
            T jv[NT];
#pragma unroll
            for (int u = 0; u < NT; ++u) {
                const int i = part + 16 * u;
                jv[u] = i < SFX_NJ * 3 ? M.JS[i * 32 + sidx] : (T)0;
            }
#pragma unroll
            for (int u = 0; u < NT; ++u) {
                const int i = part + 16 * u;
                if (i < SFX_NJ * 3) acc += jv[u] * S.gl[i];
            }
        }
        S.c[t] = acc;
    }
    SFX_SYNC();
    SFX_LAP(S, 14);
    SFX_FOR(sidx, 32) {
        T acc = 0;
        if (sidx < NS) {
            acc = S.dc[SFX_NPF + sidx];
            for (int part = 0; part < 16; ++part) acc += S.c[part * 32 + sidx];
        }
        S.dshape[sidx] = acc;
    }
    // ---- 10. priors + gradient wrt the parameter vector --------------------------------
    const T bpw2 = (T)st.body_pose_weight * (T)st.body_pose_weight;
    const T sw2 = (T)st.shape_weight * (T)st.shape_weight;
    const T hw2 = (T)st.hand_prior_weight * (T)st.hand_prior_weight;
    const T ew2 = (T)st.expr_prior_weight * (T)st.expr_prior_weight;
    const T bendw = (T)st.bending_prior_weight;
    T gmm_val = 0;
    const bool use_gmm = body && st.pprior_kind == SFX_PPRIOR_GMM;
    if (use_gmm) gmm_val = gmm_prior(M, S.x + L.off_pose, S, S.dvp);    // c, dvp are free by now
    if (vposer) {
        // gradient wrt the decoded body pose: data term (dfp) + bending prior, then the decoder
        SFX_SYNC();
        SFX_FOR(e, 63) {
            T gv = S.dfp[3 + e];
            if (body && (e == 52 || e == 55 || e == 9 || e == 12)) {
                T sg = e == 52 ? (T)1 : (T)-1;
                T ex = sfx_exp(S.fp[3 + e] * sg);
                gv += bendw * (T)2 * ex * ex * sg;
            }
            S.dbp[e] = gv;
        }
        vposer_adjoint(M, S, S.gl, stream_ws);
    }
    SFX_SYNC();
    SFX_FOR(i, L.np) {
        T gv = 0;
        if (i >= L.off_camt && i < L.off_camt + 3) {
            gv = gct[i - L.off_camt];
            if (!body && st.depth_loss_weight > 0 && i == L.off_camt + 2) {
                T dlw = (T)st.depth_loss_weight;
                gv += (T)2 * dlw * dlw * (S.x[i] - cam[SFX_CAM_TZ]);
            }
        } else if (i >= L.off_go && i < L.off_go + 3) {
            gv = S.dfp[i - L.off_go];
        } else if (vposer && i >= L.off_pose && i < L.off_pose + L.n_pose) {
            // latent pose: decoder adjoint (S.gl) + latent prior (fitting.py:389-395)
            int e = i - L.off_pose;
            gv = S.gl[e];
            if (body) gv += (T)2 * bpw2 * (latent_reg ? S.x[i] - reg_pose[e] : S.x[i]);
        } else if (i >= L.off_pose && i < L.off_pose + L.n_pose) {
            int e = i - L.off_pose;
            gv = S.dfp[3 + e];
            if (body) {
                if (st.pprior_kind == SFX_PPRIOR_REGRESSION)
                    gv += (T)2 * bpw2 * (S.x[i] - reg_pose[e]);
                else if (st.pprior_kind == SFX_PPRIOR_L2)
                    gv += (T)2 * bpw2 * S.x[i];
                else if (st.pprior_kind == SFX_PPRIOR_GMM)
                    gv += bpw2 * S.dvp[e];
                // bending prior on full_pose[3:66][52, 55, 9, 12] with signs (+ - - -)
                if (e == 52 || e == 55 || e == 9 || e == 12) {
                    T sg = e == 52 ? (T)1 : (T)-1;
                    T ex = sfx_exp(S.fp[3 + e] * sg);
                    gv += bendw * (T)2 * ex * ex * sg;
                }
            }
        } else if (i >= L.off_jaw && i < L.off_jaw + 3) {
            int e = i - L.off_jaw;
            gv = S.dfp[66 + e];
            if (body) {
                T jw = (T)st.jaw_prior_weight[e];
                gv += (T)2 * jw * jw * S.x[i];
            }
        } else if (i >= L.off_leye && i < L.off_leye + 3) {
            gv = S.dfp[69 + i - L.off_leye];
        } else if (i >= L.off_reye && i < L.off_reye + 3) {
            gv = S.dfp[72 + i - L.off_reye];
        } else if (i >= L.off_betas && i < L.off_betas + L.n_betas) {
            gv = S.dshape[i - L.off_betas];
            if (body) gv += (T)2 * sw2 * S.x[i];
        } else if (i >= L.off_expr && i < L.off_expr + L.n_expr) {
            gv = S.dshape[L.n_betas + i - L.off_expr];
            if (body) gv += (T)2 * ew2 * S.x[i];
        } else if (i >= L.off_lh && i < L.off_lh + 2 * L.n_hand) {
            int h = (i - L.off_lh) / L.n_hand, k = (i - L.off_lh) % L.n_hand;
            const T* dh = S.dfp + 75 + 45 * h;
            const T* hv = S.hand + 45 * h;
            T acc = 0;
            if (S.hand_cached) {
                const float* C = S.hand_c + (h * L.n_hand + k) * 45;
                for (int e0 = 0; e0 < 45; e0 += 5) {       // operands of five terms first, sums in order
                    T cv[5], tv[5];
#pragma unroll
                    for (int u = 0; u < 5; ++u) {
                        cv[u] = (T)C[e0 + u];
                        tv[u] = dh[e0 + u] + (body ? (T)2 * hw2 * hv[e0 + u] : (T)0);
                    }
#pragma unroll
                    for (int u = 0; u < 5; ++u) acc += cv[u] * tv[u];
                }
            } else {
                const T* C = (h ? M.hand_r : M.hand_l) + k * 45;
                for (int e = 0; e < 45; ++e)
                    acc += C[e] * (dh[e] + (body ? (T)2 * hw2 * hv[e] : (T)0));
            }
            gv = acc;
        }
        S.gfull[i] = gv;
    }
    // ---- 11. total loss, summed in the order of fitting.py:457-460 ----------------------
    T total = data_loss;
    if (body) {
        // S.red[4..8]: the prior sums of squares taken together with the data term above
        const T pp = use_gmm ? gmm_val : S.red[4];
        total += pp * bpw2;
        total += S.red[5] * sw2;
        T ang = 0;
        {
            const int idx[4] = {52, 55, 9, 12};
            for (int a = 0; a < 4; ++a) {
                T e4 = sfx_exp(S.fp[3 + idx[a]] * (a == 0 ? (T)1 : (T)-1));
                ang += e4 * e4;
            }
        }
        total += ang * bendw;
        if (coll) total += (T)st.coll_loss_weight * S.coll_loss;      // pen_loss (fitting.py:453-458)
        T jaw = 0;
        for (int e = 0; e < 3; ++e) {
            T v = S.x[L.off_jaw + e] * (T)st.jaw_prior_weight[e];
            jaw += v * v;
        }
        total += jaw;
        total += S.red[6] * ew2;
        total += S.red[7] * hw2;
        total += S.red[8] * hw2;
    } else if (st.depth_loss_weight > 0) {
        T dlw = (T)st.depth_loss_weight;
        T dz = ct[2] - cam[SFX_CAM_TZ];
        total += dlw * dlw * (dz * dz);
    }
    SFX_SYNC();
    if (SFX_TID == 0) {
        S.loss = total;
        S.n_evals += 1;
        // rows of the blend matrix streamed by this evaluation (forward + adjoint)
        S.n_passes += (coll ? 3 * M.V : S.n_rows) +
                      (st.need_blend_grad ? S.n_rows + (coll ? 3 * S.n_touch : 0) : 0);
    }
    SFX_SYNC();
    SFX_LAP(S, 15);
    if (!coll) SFX_PROF_END(S, 10, tail);
    SFX_PROF_END(S, 0, eval);
#ifdef SFX_TRACE
    SFX_TRACE((double)total);
#endif
}

// Per-stage effective joint weights (fit_single_frame.py:569-574): body block keeps the base
// weights, the 42 hand keypoints take hand_joint_weight, the face block face_joint_weight, and
// low-confidence keypoints are zeroed again.
template <typename T>
SFX_FN void stage_joint_weights_only(const SfxStage& st, const T* jw_base, const unsigned char* lowconf,
                                     int K, Scratch<T>& S) {
    SFX_FOR(k, K) {
        T w = jw_base[k];
        if (k >= st.n_body_kpts + 42) w = (T)st.face_joint_weight;
        else if (k >= st.n_body_kpts) w = (T)st.hand_joint_weight;
        if (lowconf[k]) w = 0;
        S.jw[k] = w;
    }
    SFX_SYNC();
}

// camera stage with use_conf_for_camera_init: sum_k conf_k^2 over the init joints (constant
// during the stage; fitting.py:509-511)
template <typename T>
SFX_FN void stage_confidence_sum(const T* conf, const unsigned char* init_mask, int K, Scratch<T>& S) {
    T v = block_reduce<T>(K, [=](int k) { return init_mask[k] ? conf[k] * conf[k] : (T)0; },
                          OpAdd<T>(), (T)0, &S.red[0]);
    if (SFX_TID == 0) S.confsq = v;
    SFX_SYNC();
}

// Support rows the stage has to stream: the three rows of every slot whose model joint feeds at
// least one keypoint with a non-zero weight (hand / face weights are zero in the early stages,
// undetected keypoints have zero confidence, the camera stage looks at a dozen body joints).
// Exact: the skipped rows only ever enter terms multiplied by a zero weight.
template <typename T>
SFX_FN void stage_live_rows(const ModelView<T>& M, const SfxStage& st, const T* conf,
                            const unsigned char* init_mask, Scratch<T>& S, bool all_rows) {
    SFX_SYNC();
    SFX_FOR(s, SFX_NSLOT) {
        const int j = s < SFX_NEXTRA ? SFX_NJ + s : SFX_NJ + SFX_NEXTRA + (s - SFX_NEXTRA) / 3;
        bool live = all_rows;
        if (!live && j < M.NJOUT)
            for (int e = M.inv_ptr[j]; e < M.inv_ptr[j + 1]; ++e) {
                const int k = M.inv_idx[e];
                if (st.loss_kind == SFX_LOSS_CAMERA_INIT) live = live || init_mask[k] != 0;
                else live = live || (st.use_joints_conf ? S.jw[k] * conf[k] : S.jw[k]) != (T)0;
            }
        S.slot_live[s] = live ? 1 : 0;
    }
    SFX_SYNC();
    // Row r stays with streaming warp r % 15 and the warps keep ascending order, exactly as when
    // all 675 rows are streamed: the skipped rows would have added exact zeros to the adjoint's
    // per-warp partial sums, so the result is bit-identical to streaming everything.
    if (SFX_TID == 0) {
        int n = 0;
        for (int w = 0; w < SFX_NSTREAM; ++w) {
            S.wptr[w] = (unsigned short)n;
            for (int r = w; r < SFX_NSLOT * 3; r += SFX_NSTREAM)
                if (S.slot_live[r / 3]) S.rows[n++] = (unsigned short)r;
        }
        S.wptr[SFX_NSTREAM] = (unsigned short)n;
        S.n_rows = n;
        S.rows_dirty = 1;
    }
    SFX_SYNC();
}

// everything a stage needs before its first evaluation
template <typename T>
SFX_FN void stage_setup(const ModelView<T>& M, const SfxStage& st, const T* jw_base,
                        const unsigned char* lowconf, const T* conf, const unsigned char* init_mask,
                        int K, Scratch<T>& S, bool all_rows = false) {
    stage_joint_weights_only(st, jw_base, lowconf, K, S);
    if (st.loss_kind == SFX_LOSS_CAMERA_INIT && st.use_conf_camera)
        stage_confidence_sum(conf, init_mask, K, S);
    stage_live_rows(M, st, conf, init_mask, S, all_rows);
}

// ------------------------------------------------------- optimiser plumbing (compact <-> full)
template <typename T>
SFX_FN void scatter_active(Scratch<T>& S, int D) {        // xa -> x
    SFX_FOR(i, D) S.x[S.act[i]] = S.xa[i];
    SFX_SYNC();
}
template <typename T>
SFX_FN void gather_grad(Scratch<T>& S, int D, T* dst) {   // gfull -> dst (compact)
    SFX_FOR(i, D) dst[i] = S.gfull[S.act[i]];
    SFX_SYNC();
}

template <typename T>
struct EvalCtx {
    const ModelView<T>* M;
    const SfxLayout* L;
    const SfxStage* st;
    const T* gt;
    const T* conf;
    const unsigned char* init_mask;
    const T* cam;
    const T* reg_pose;
    void* stream_ws;
    const CollWS<T>* coll;      // workspace of the interpenetration term, or nullptr
    T* gram;                    // [2][SFX_HIST][SFX_HIST] s_i.y_j and y_i.y_j by history slot (Gram two-loop), or nullptr
};

// closure(): evaluate at S.xa, leave loss in S.loss and the compact gradient in S.gl
// ("last evaluated gradient" -- what run_fitting's gtol test reads, fitting.py:191-193).
template <typename T>
SFX_FN double closure(const EvalCtx<T>& E, Scratch<T>& S) {
    const int D = E.st->n_active;
    scatter_active(S, D);
    eval_frame(*E.M, *E.L, *E.st, E.gt, E.conf, E.init_mask, E.cam, E.reg_pose, S, E.stream_ws, E.coll);
    gather_grad(S, D, S.gl);
    return (double)S.loss;
}

// lswolfe cubic interpolation (lbfgs_ls.py:11-36); all scalars in double
SFX_FN double cubic_min(double x1, double f1, double g1, double x2, double f2, double g2,
                        bool has_bounds, double lo, double hi) {
    if (!has_bounds) {
        lo = x1 <= x2 ? x1 : x2;
        hi = x1 <= x2 ? x2 : x1;
    }
    double d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2);
    double disc = d1 * d1 - g1 * g2;
    if (disc >= 0) {
        double d2 = sqrt(disc);
        double pos = x1 <= x2 ? x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))
                              : x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2 * d2));
        return fmin(fmax(pos, lo), hi);
    }
    return (lo + hi) / 2.;
}

SFX_FN double dmax(double a, double b) { return a > b ? a : b; }
SFX_FN double dmin(double a, double b) { return a < b ? a : b; }

// Evaluate the objective at x0 + t d (lbfgs_ls.py:249-254); x is restored by the caller's
// next write of xa, so only xa is touched here.
template <typename T>
SFX_FN double probe(const EvalCtx<T>& E, Scratch<T>& S, double t, T* gdst, double* gtd_out) {
    const int D = E.st->n_active;
    const T tt = (T)t;
#ifdef SFX_TRACE_STEP
    SFX_TRACE_STEP(t);
#endif
    SFX_LAP(S, 20);                       // line-search logic since the previous probe
    SFX_FOR(i, D) S.xa[i] = S.x0[i] + tt * S.d[i];
    SFX_SYNC();
    double f = closure(E, S);
    SFX_LAP(S, 21);                       // gather_grad after the evaluation
    SFX_FOR(i, D) gdst[i] = S.gl[i];
    *gtd_out = (double)block_dot(S.gl, S.d, D, &S.red[2]);
    SFX_LAP(S, 22);                       // copy + directional derivative
    return f;
}

template <typename T>
SFX_FN void vcopy(T* dst, const T* src, int n) {
    SFX_FOR(i, n) dst[i] = src[i];
    SFX_SYNC();
}

// Strong-Wolfe line search (lbfgs_ls.py:39-167).  On entry S.g = gradient at x0, on exit
// S.g = gradient at the accepted point; returns f there and the step in *t_io.
// Gradient slots: g_prev (bracket phase "previous"), bg0 / bg1 (bracket ends), q (new probe).
template <typename T>
SFX_FN double strong_wolfe(const EvalCtx<T>& E, Scratch<T>& S, double* t_io, double f, double gtd,
                           double d_norm, int* n_evals_out) {
    const SfxStage& st = *E.st;
    const int D = st.n_active;
    const double c1 = 1e-4, c2 = 0.9;
    const int max_ls = 25;
    const int max_iter = st.max_iter;
    double t = *t_io;
    double gtd_new;
    double f_new = probe(E, S, t, S.q, &gtd_new);
    int evals = 1;
    double t_prev = 0, f_prev = f, gtd_prev = gtd;
    vcopy(S.g_prev, S.g, D);
    bool done = false;
    int it = 0;
    // bracket: positions bt, values bf, directional derivatives bd; gradients in bg0 / bg1
    double bt[2] = {0, 0}, bf[2] = {0, 0}, bd[2] = {0, 0};
    int nb = 0;
    while (it < max_ls) {
        if (f_new > (f + c1 * t * gtd) || (it > 1 && f_new >= f_prev)) {
            bt[0] = t_prev; bt[1] = t; bf[0] = f_prev; bf[1] = f_new; bd[0] = gtd_prev; bd[1] = gtd_new;
            vcopy(S.bg0, S.g_prev, D); vcopy(S.bg1, S.q, D);
            nb = 2;
            break;
        }
        if (fabs(gtd_new) <= -c2 * gtd) {
            bt[0] = t; bf[0] = f_new; bd[0] = gtd_new;
            vcopy(S.bg0, S.q, D);
            nb = 1;
            done = true;
            break;
        }
        if (gtd_new >= 0) {
            bt[0] = t_prev; bt[1] = t; bf[0] = f_prev; bf[1] = f_new; bd[0] = gtd_prev; bd[1] = gtd_new;
            vcopy(S.bg0, S.g_prev, D); vcopy(S.bg1, S.q, D);
            nb = 2;
            break;
        }
        double lo = t + 0.01 * (t - t_prev), hi = t * 10;
        double tmp = t;
        t = cubic_min(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, true, lo, hi);
        t_prev = tmp; f_prev = f_new; gtd_prev = gtd_new;
        vcopy(S.g_prev, S.q, D);
        f_new = probe(E, S, t, S.q, &gtd_new);
        evals += 1;
        it += 1;
    }
    if (it == max_ls) {
        bt[0] = 0; bt[1] = t; bf[0] = f; bf[1] = f_new; bd[0] = gtd; bd[1] = gtd_new;
        vcopy(S.bg0, S.g, D); vcopy(S.bg1, S.q, D);
        nb = 2;
    }
    bool stalled = false;
    int lo_i = 0, hi_i = 1;
    if (nb == 2 && !(bf[0] <= bf[1])) { lo_i = 1; hi_i = 0; }
    if (nb == 1) { lo_i = 0; hi_i = 0; }
    while (!done && it < max_iter) {
        t = cubic_min(bt[0], bf[0], bd[0], bt[1], bf[1], bd[1], false, 0, 0);
        double bmax = dmax(bt[0], bt[1]), bmin = dmin(bt[0], bt[1]);
        double eps = 0.1 * (bmax - bmin);
        if (dmin(bmax - t, t - bmin) < eps) {
            if (stalled || t >= bmax || t <= bmin) {
                t = fabs(t - bmax) < fabs(t - bmin) ? bmax - eps : bmin + eps;
                stalled = false;
            } else {
                stalled = true;
            }
        } else {
            stalled = false;
        }
        f_new = probe(E, S, t, S.q, &gtd_new);
        evals += 1;
        it += 1;
        if (f_new > (f + c1 * t * gtd) || f_new >= bf[lo_i]) {
            bt[hi_i] = t; bf[hi_i] = f_new; bd[hi_i] = gtd_new;
            vcopy(hi_i ? S.bg1 : S.bg0, S.q, D);
            if (bf[0] <= bf[1]) { lo_i = 0; hi_i = 1; } else { lo_i = 1; hi_i = 0; }
        } else {
            if (fabs(gtd_new) <= -c2 * gtd) {
                done = true;
            } else if (gtd_new * (bt[hi_i] - bt[lo_i]) >= 0) {
                bt[hi_i] = bt[lo_i]; bf[hi_i] = bf[lo_i]; bd[hi_i] = bd[lo_i];
                vcopy(hi_i ? S.bg1 : S.bg0, lo_i ? S.bg1 : S.bg0, D);
            }
            bt[lo_i] = t; bf[lo_i] = f_new; bd[lo_i] = gtd_new;
            vcopy(lo_i ? S.bg1 : S.bg0, S.q, D);
        }
        if (fabs(bt[1] - bt[0]) * d_norm < st.tol_change) break;
    }
    vcopy(S.g, lo_i ? S.bg1 : S.bg0, D);
    *t_io = bt[lo_i];
    *n_evals_out = evals;
    return bf[lo_i];
}

// persistent optimiser state of one frame within a stage (lbfgs_ls.py:292-300, :436-443)
template <typename T>
struct LbfgsState {
    int n_iter;          // global iteration counter
    int num_old;         // stored (s, y) pairs
    int head;            // ring start
    double t;
    T H_diag;
    double prev_loss;
};


#ifdef __CUDACC__
// Sum of the 32 per-lane partials in exactly the association of the xor butterfly
// (p + shfl_xor 16, 8, 4, 2, 1), but through shared memory: one store, one warp barrier, eight
// broadcast 16-byte loads and a 5-deep add tree -- about half the latency of five dependent
// shuffles, which matters because every step of the two-loop recursion waits on this value.
template <typename T>
__device__ __forceinline__ T warp_tree_sum(T p, T* slot, int lane) {
    slot[lane] = p;
    __syncwarp();
    T a[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = slot[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int l = 0; l < o; ++l) a[l] = a[l] + a[l + o];
    return a[0];
}
#endif

#ifdef __CUDACC__
// Two-loop recursion of one L-BFGS iteration (lbfgs_ls.py:336-358) by warp 0 alone: the
// direction lives in registers (element e <-> lane e % 32, register e / 32), the (s, y) pairs
// stream from global memory one pair ahead of use, and nothing synchronises the block.  The
// arithmetic -- per-lane partial sums over e = lane, lane + 32, ... followed by an xor
// butterfly -- is the one block_reduce() performs, so both paths produce the same bits.
template <typename T, int NR>
__device__ __noinline__ void two_loop_warp_n(Scratch<T>& S, int k, int head, int H, T hd,
                                            const T* __restrict__ hist_s,
                                            const T* __restrict__ hist_y, int D) {
    SFX_ASSUME_SHARED(S);
    constexpr int PF = 3;                 // (s, y) pairs in flight ahead of the one in use
    const int lane = threadIdx.x;
    // only the last register of a lane can lie beyond D
    const bool tail_live = 32 * (NR - 1) + lane < D;
    T q[NR], sb[PF][NR], yb[PF][NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int e = 32 * r + lane;
        q[r] = e < D ? -S.g[e] : (T)0;
    }
    // ring position of pair i is (head + i) % H; the loops walk it incrementally
#define SFX_TL_FETCH(slot, pos)                                                    \
    do {                                                                           \
        const T* ps_ = hist_s + (pos) * SFX_NP_MAX + lane;                         \
        const T* py_ = hist_y + (pos) * SFX_NP_MAX + lane;                         \
        _Pragma("unroll") for (int r = 0; r < NR - 1; ++r) {                       \
            sb[slot][r] = ps_[32 * r];                                             \
            yb[slot][r] = py_[32 * r];                                             \
        }                                                                          \
        sb[slot][NR - 1] = tail_live ? ps_[32 * (NR - 1)] : (T)0;                  \
        yb[slot][NR - 1] = tail_live ? py_[32 * (NR - 1)] : (T)0;                  \
    } while (0)
    // ---- first loop: i = k-1 .. 0 ----
    int fpos = (head + k - 1) % H;        // position of the next pair to fetch
#pragma unroll
    for (int u = 0; u < PF; ++u)
        if (k - 1 - u >= 0) {
            SFX_TL_FETCH(u, fpos);
            fpos = fpos == 0 ? H - 1 : fpos - 1;
        }
    for (int i0 = k - 1; i0 >= 0; i0 -= PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int i = i0 - u;
            if (i >= 0) {
                T p = 0;
#pragma unroll
                for (int r = 0; r < NR; ++r) p = p + sb[u][r] * q[r];
                p = warp_tree_sum(p, S.tl_red + 32 * (i & 1), lane);
                const T a = p * S.ro[i];
                if (lane == 0) S.al[i] = a;
#pragma unroll
                for (int r = 0; r < NR; ++r) q[r] += -a * yb[u][r];
                if (i - PF >= 0) {
                    SFX_TL_FETCH(u, fpos);
                    fpos = fpos == 0 ? H - 1 : fpos - 1;
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) q[r] = q[r] * hd;       // q now holds the direction d
    __syncwarp();
    // ---- second loop: i = 0 .. k-1 ----
    fpos = head % H;
#pragma unroll
    for (int u = 0; u < PF; ++u)
        if (u < k) {
            SFX_TL_FETCH(u, fpos);
            fpos = fpos + 1 == H ? 0 : fpos + 1;
        }
    for (int i0 = 0; i0 < k; i0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int i = i0 + u;
            if (i < k) {
                T p = 0;
#pragma unroll
                for (int r = 0; r < NR; ++r) p = p + yb[u][r] * q[r];
                p = warp_tree_sum(p, S.tl_red + 32 * (i & 1), lane);
                const T co = S.al[i] - p * S.ro[i];
#pragma unroll
                for (int r = 0; r < NR; ++r) q[r] += co * sb[u][r];
                if (i + PF < k) {
                    SFX_TL_FETCH(u, fpos);
                    fpos = fpos + 1 == H ? 0 : fpos + 1;
                }
            }
        }
    }
#undef SFX_TL_FETCH
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int e = 32 * r + lane;
        if (e < D) S.d[e] = q[r];
    }
}

template <typename T>
__device__ __forceinline__ void two_loop_warp(Scratch<T>& S, int k, int head, int H, T hd,
                                              const T* hist_s, const T* hist_y, int D) {
    // elements beyond D are zeros and add nothing, so the register count follows D
    if (D <= 32) two_loop_warp_n<T, 1>(S, k, head, H, hd, hist_s, hist_y, D);
    else if (D <= 128) two_loop_warp_n<T, 4>(S, k, head, H, hd, hist_s, hist_y, D);
    else two_loop_warp_n<T, SFX_NP_MAX / 32>(S, k, head, H, hd, hist_s, hist_y, D);
}
#endif

// ---- two-loop recursion in coefficient space ("Gram" two-loop, SfxStage::generic_two_loop == 2) ----
// The recursion of lbfgs_ls.py:336-358 only ever combines the vectors g, s_i, y_i, so it can be
// run on their inner products: with SY[i][j] = s_i.y_j, YY[i][j] = y_i.y_j (kept per frame and
// extended by one row / column whenever a pair enters the history), sg[i] = s_i.g and
// yg[i] = y_i.g, the 2k dependent steps become scalar recurrences
//     first loop   j = k-1..0 :  al_j = ro_j b_j ;  b_i -= al_j SY[i][j] (i < j) ;  e_i -= al_j YY[i][j]
//     second loop  j = 0..k-1 :  c_j = al_j - ro_j f_j ;  f_i += c_j SY[j][i] (i > j),  f = H_diag e
// (b_i = s_i.q, e_i = y_i.q, f_i = y_i.r of the textbook recursion) and the direction is one
// combination  d = -H_diag g + sum_j (-H_diag al_j) y_j + c_j s_j.  A step of the chain is a
// register broadcast and one multiply-add (80 cycles measured, gram_chain_f32) instead of a 32-lane
// floating-point sum over the parameter vector (~250); the inner products and the final combination are spread
// over all warps.  Same mathematics as the reference's recursion, different rounding: it is an
// explicit option (the default recursion reproduces the reference's operation order).
#ifdef __CUDACC__
#define SFX_GRAM_NL 32
#define SFX_GRAM_NRK ((SFX_HIST + 31) / 32)
#define SFX_GRAM_NRD (SFX_NP_MAX / 32)
#else
#define SFX_GRAM_NL 1
#define SFX_GRAM_NRK SFX_HIST
#define SFX_GRAM_NRD SFX_NP_MAX
#endif

template <typename T>
SFX_FN T gram_warp_sum(T p) {
#ifdef __CUDACC__
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
#endif
    return p;
}
// value `idx` of a lane-distributed array (element i lives in lane i % NL, register i / NL)
template <typename T>
SFX_FN T gram_pick(const T* regs, int idx) {
#ifdef __CUDACC__
    static_assert(SFX_GRAM_NRK <= 4, "history longer than 128 pairs");
    const int r = idx >> 5;
    T v = regs[0];
#pragma unroll
    for (int q = 1; q < SFX_GRAM_NRK; ++q) v = r == q ? regs[q] : v;
    return __shfl_sync(0xffffffffu, v, idx & 31);
#else
    return regs[idx];
#endif
}

// Packed view of the two Gram matrices in recursion order (pair 0 = oldest): element (i, j) is
// s_i.y_j above the diagonal and y_i.y_j on / below it -- all the recursion reads.
template <typename T>
struct GramStaged {            // k x k copy in shared memory, odd row stride (conflict-free by row and by column)
    const T* G;
    int ld;
    SFX_MFN T operator()(int i, int j) const { return G[i * ld + j]; }
};
template <typename T>
struct GramInPlace {           // the per-frame global arrays, indexed by history slot
    const T* GSY;
    const T* GYY;
    int head, H;
    SFX_MFN T operator()(int i, int j) const {
        const int pi = (head + i) % H, pj = (head + j) % H;
        return i < j ? GSY[pi * SFX_HIST + pj] : GYY[pi * SFX_HIST + pj];
    }
};

// the two scalar recurrences, run by one warp with the state in registers; the loop bodies are
// branch-free (clamped loads + selects) so that the warp never diverges inside the chain
template <typename T, typename ACC>
SFX_FN void gram_chain(Scratch<T>& S, int k, T hd, ACC G, int lane) {
    T b[SFX_GRAM_NRK], e[SFX_GRAM_NRK];
    int ic[SFX_GRAM_NRK];
#pragma unroll
    for (int r = 0; r < SFX_GRAM_NRK; ++r) {
        const int i = lane + SFX_GRAM_NL * r;
        ic[r] = i < k ? i : k - 1;                       // clamped: always a valid element
        b[r] = i < k ? -S.sg[i] : (T)0;
        e[r] = i < k ? -S.yg[i] : (T)0;
    }
    const T* ro = S.ro;
    T* al = S.al;
    T* cf = S.cf;
    for (int j = k - 1; j >= 0; --j) {
        // column j (s_i.y_j for i < j, y_i.y_j for i >= j) and row j (y_j.y_i for i < j): no
        // dependence on the chain, so the loads overlap the broadcast below
        T colv[SFX_GRAM_NRK], rowv[SFX_GRAM_NRK];
#pragma unroll
        for (int r = 0; r < SFX_GRAM_NRK; ++r) {
            colv[r] = G(ic[r], j);
            rowv[r] = G(j, ic[r]);
        }
        const T a = gram_pick(b, j) * ro[j];
        if (lane == 0) al[j] = a;
#pragma unroll
        for (int r = 0; r < SFX_GRAM_NRK; ++r) {
            const int i = lane + SFX_GRAM_NL * r;
            const bool lo = i < j, in = i < k;
            const T cb = lo ? colv[r] : (T)0;
            const T ce = lo ? rowv[r] : (in ? colv[r] : (T)0);
            b[r] -= a * cb;
            e[r] -= a * ce;
        }
    }
    SFX_SYNCWARP();
#pragma unroll
    for (int r = 0; r < SFX_GRAM_NRK; ++r) e[r] = e[r] * hd;      // f_i = y_i . (H_diag q)
    for (int j = 0; j < k; ++j) {
        T rowv[SFX_GRAM_NRK];
#pragma unroll
        for (int r = 0; r < SFX_GRAM_NRK; ++r) rowv[r] = G(j, ic[r]);
        const T c = al[j] - gram_pick(e, j) * ro[j];
        if (lane == 0) cf[j] = c;
#pragma unroll
        for (int r = 0; r < SFX_GRAM_NRK; ++r) {
            const int i = lane + SFX_GRAM_NL * r;
            const T cv = (i > j && i < k) ? rowv[r] : (T)0;
            e[r] += c * cv;
        }
    }
}

// Inner products of every pair with the gradient and, when a pair was just stored, with that pair
// (S.x0 = s_new, S.q = y_new: lbfgs_step).  Warp w takes pairs w, w + nw, ...; P pairs (2 P NR
// loads per lane) are in flight at once because every row costs an L2 round trip.
template <typename T, int NR, int P>
SFX_FN void gram_dots(Scratch<T>& S, int k, int head, int H, const T* hist_s, const T* hist_y, int D,
                      bool has_new, T* GSY, T* GYY, T* sG, int lds, int warp, int nw, int lane) {
    constexpr int LD = SFX_HIST;
    const int pn = (head + k - 1) % H;
    const int n = k - 1;                              // recursion index of the new pair
    T gv[NR], sn[NR], yn[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int el = lane + SFX_GRAM_NL * r;
        gv[r] = el < D ? S.g[el] : (T)0;
        sn[r] = (has_new && el < D) ? S.x0[el] : (T)0;
        yn[r] = (has_new && el < D) ? S.q[el] : (T)0;
    }
    for (int i0 = warp; i0 < k; i0 += P * nw) {
        T sv[P][NR], yv[P][NR];
        int pi[P];
#pragma unroll
        for (int h = 0; h < P; ++h) {
            const int i = i0 + h * nw;
            pi[h] = (head + (i < k ? i : i0)) % H;    // past the end: repeat a valid pair (unused)
            const T* srow = hist_s + (long)pi[h] * SFX_NP_MAX;
            const T* yrow = hist_y + (long)pi[h] * SFX_NP_MAX;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int el = lane + SFX_GRAM_NL * r;
                sv[h][r] = el < D ? srow[el] : (T)0;
                yv[h][r] = el < D ? yrow[el] : (T)0;
            }
        }
#pragma unroll
        for (int h = 0; h < P; ++h) {
            T q[5];
#pragma unroll
            for (int u = 0; u < 5; ++u) q[u] = 0;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                q[0] += sv[h][r] * gv[r];
                q[1] += yv[h][r] * gv[r];
                q[2] += sv[h][r] * yn[r];
                q[3] += sn[r] * yv[h][r];
                q[4] += yv[h][r] * yn[r];
            }
#pragma unroll
            for (int u = 0; u < 5; ++u) q[u] = gram_warp_sum(q[u]);
            const int i = i0 + h * nw;
            if (lane == 0 && i < k) {
                S.sg[i] = q[0];
                S.yg[i] = q[1];
                if (has_new) {
                    GSY[pi[h] * LD + pn] = q[2];
                    GSY[pn * LD + pi[h]] = q[3];
                    GYY[pi[h] * LD + pn] = q[4];
                    GYY[pn * LD + pi[h]] = q[4];
                    if (sG) {
                        if (i < n) sG[i * lds + n] = q[2];      // s_i . y_new
                        sG[n * lds + i] = q[4];                 // y_new . y_i  (i <= n)
                    }
                }
            }
        }
    }
}

// d = -H_diag g + sum_j (-H_diag al_j) y_j + c_j s_j : pairs spread over the warps (P in flight),
// per-warp partial vectors in `part`
template <typename T, int NR, int P>
SFX_FN void gram_combine(Scratch<T>& S, int k, int head, int H, T hd, const T* hist_s, const T* hist_y,
                         int D, T* part, int warp, int nw, int lane) {
    T acc[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[r] = 0;
    for (int j0 = warp; j0 < k; j0 += P * nw) {
        T sv[P][NR], yv[P][NR], cy[P], cs[P];
#pragma unroll
        for (int h = 0; h < P; ++h) {
            const int j = j0 + h * nw;
            const bool live = j < k;
            const int pj = (head + (live ? j : j0)) % H;
            cy[h] = live ? -hd * S.al[j] : (T)0;
            cs[h] = live ? S.cf[j] : (T)0;
            const T* srow = hist_s + (long)pj * SFX_NP_MAX;
            const T* yrow = hist_y + (long)pj * SFX_NP_MAX;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int el = lane + SFX_GRAM_NL * r;
                sv[h][r] = el < D ? srow[el] : (T)0;
                yv[h][r] = el < D ? yrow[el] : (T)0;
            }
        }
#pragma unroll
        for (int h = 0; h < P; ++h)
#pragma unroll
            for (int r = 0; r < NR; ++r) acc[r] += cy[h] * yv[h][r] + cs[h] * sv[h][r];
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int el = lane + SFX_GRAM_NL * r;
        if (el < D) part[warp * SFX_NP_MAX + el] = acc[r];
    }
}

template <typename T>
SFX_FN_NOINLINE void gram_two_loop(const EvalCtx<T>& E, Scratch<T>& S, int k, int head, int H, T hd,
                          const T* hist_s, const T* hist_y, int D, bool has_new) {
    SFX_ASSUME_SHARED(E);
    SFX_ASSUME_SHARED(S);
    constexpr int LD = SFX_HIST;
    T* GSY = E.gram;
    T* GYY = E.gram + (size_t)LD * LD;
#ifdef __CUDACC__
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31;
#else
    const int warp = 0, nw = 1, lane = 0;
#endif
    SFX_SYNC();
    SFX_PROF_BEGIN(gd);
    // ---- the live k x k blocks, packed, into the idle blend ring when they fit (otherwise --
    //      float64 with a long history -- they are read in place).  The row / column of a pair
    //      stored in this iteration is written by gram_dots, straight from its inner products. ----
    T* part = S.q;                     // per-warp partial directions (host: one "warp")
    T* sG = nullptr;
    int lds = 0;
    bool fast = false;
#ifdef __CUDACC__
    {
        size_t area_bytes;
        unsigned char* area = idle_area<T>(E.stream_ws, &area_bytes);
        part = reinterpret_cast<T*>(area);
        const size_t part_bytes = (size_t)nw * SFX_NP_MAX * sizeof(T);
        const int ko = has_new ? k - 1 : k;              // pairs whose products are already stored
        if constexpr (sizeof(T) == 4)
            fast = part_bytes + (size_t)SFX_GRAM_ROWS * SFX_GRAM_LDF * sizeof(T) <= area_bytes;
        if (fast) {
            // float32 with the ring: fixed-stride [128][129] block, zero outside the live part,
            // for gram_chain_f32.  A warp copies four rows at a time (16 loads in flight per lane).
            static_assert(SFX_GRAM_ROWS == 128, "four 32-column registers per row");
            sG = reinterpret_cast<T*>(area + part_bytes);
            lds = SFX_GRAM_LDF;
            int pj[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int j = lane + 32 * r;
                pj[r] = head + (j < ko ? j : 0);
                if (pj[r] >= H) pj[r] -= H;
            }
            for (int i0 = warp; i0 < SFX_GRAM_ROWS; i0 += 4 * nw) {
                T v[4][4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int i = i0 + h * nw;
                    int pi = head + (i < ko ? i : 0);
                    if (pi >= H) pi -= H;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int j = lane + 32 * r;
                        v[h][r] = (i < ko && j < ko) ? (i < j ? GSY[pi * LD + pj[r]] : GYY[pi * LD + pj[r]])
                                                     : (T)0;
                    }
                }
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int i = i0 + h * nw;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int j = lane + 32 * r;
                        const bool new_pair = has_new && (i == ko || j == ko) && i < k && j < k;
                        if (i < SFX_GRAM_ROWS && !new_pair) sG[i * SFX_GRAM_LDF + j] = v[h][r];
                    }
                }
            }
        } else {
            // tight k x k block (odd row stride) for the generic chain
            lds = k | 1;
            if (part_bytes + (size_t)k * lds * sizeof(T) <= area_bytes) {
                sG = reinterpret_cast<T*>(area + part_bytes);
                const int n = ko * ko;
                constexpr int U = 4;                     // loads in flight per thread
                for (int idx0 = SFX_TID; idx0 < n; idx0 += U * SFX_NT) {
                    T v[U];
                    int dst[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int idx = idx0 + u * SFX_NT;
                        const int i = idx / ko, j = idx - i * ko;
                        int pi = head + i, pj = head + j;
                        if (pi >= H) pi -= H;
                        if (pj >= H) pj -= H;
                        dst[u] = i * lds + j;
                        v[u] = idx < n ? (i < j ? GSY[pi * LD + pj] : GYY[pi * LD + pj]) : (T)0;
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (idx0 + u * SFX_NT < n) sG[dst[u]] = v[u];
                }
            }
        }
    }
    if (D <= 128) gram_dots<T, 4, 4>(S, k, head, H, hist_s, hist_y, D, has_new, GSY, GYY, sG, lds, warp, nw, lane);
    else gram_dots<T, SFX_GRAM_NRD, 2>(S, k, head, H, hist_s, hist_y, D, has_new, GSY, GYY, sG, lds, warp, nw, lane);
#else
    gram_dots<T, SFX_GRAM_NRD, 2>(S, k, head, H, hist_s, hist_y, D, has_new, GSY, GYY, sG, lds, warp, nw, lane);
#endif
    SFX_SYNC();
    SFX_PROF_END(S, 11, gd);
    SFX_PROF_BEGIN(gc);
    if (SFX_IS_WARP0) {
        bool done = false;
#ifdef __CUDACC__
        if constexpr (sizeof(T) == 4) {
            if (fast) {
                gram_chain_f32(S, k, hd, sG, lane);
                done = true;
            }
        }
#endif
        if (!done) {
            if (sG) gram_chain(S, k, hd, GramStaged<T>{sG, lds}, lane);
            else gram_chain(S, k, hd, GramInPlace<T>{GSY, GYY, head, H}, lane);
        }
    }
    SFX_SYNC();
    SFX_PROF_END(S, 12, gc);
    SFX_PROF_BEGIN(gm);
#ifdef __CUDACC__
    if (D <= 128) gram_combine<T, 4, 4>(S, k, head, H, hd, hist_s, hist_y, D, part, warp, nw, lane);
    else gram_combine<T, SFX_GRAM_NRD, 2>(S, k, head, H, hd, hist_s, hist_y, D, part, warp, nw, lane);
#else
    gram_combine<T, SFX_GRAM_NRD, 2>(S, k, head, H, hd, hist_s, hist_y, D, part, warp, nw, lane);
#endif
    SFX_SYNC();
    SFX_FOR(el, D) {
        T v = -hd * S.g[el];
        for (int w = 0; w < nw; ++w) v += part[w * SFX_NP_MAX + el];
        S.d[el] = v;
    }
    SFX_SYNC();
    SFX_PROF_END(S, 13, gm);
#ifdef __CUDACC__
    idle_area_release(E.stream_ws);
#endif
}

// One LBFGS.step (lbfgs_ls.py:256-445).  Returns the loss at entry ("orig_loss").
template <typename T>
SFX_FN double lbfgs_step(const EvalCtx<T>& E, Scratch<T>& S, LbfgsState<T>& ls, T* hist_s, T* hist_y) {
    const SfxStage& st = *E.st;
    const int D = st.n_active;
    const int H = st.history;
    double orig_loss = closure(E, S);
    double loss = orig_loss;
    int current_evals = 1;
    vcopy(S.g, S.gl, D);
    if ((double)block_absmax(S.g, D, &S.red[2]) <= st.tol_grad) return orig_loss;
    int n_iter = 0;
    while (n_iter < st.max_iter) {
        n_iter += 1;
        ls.n_iter += 1;
        if (ls.n_iter == 1) {
            SFX_FOR(i, D) S.d[i] = -S.g[i];
            SFX_SYNC();
            ls.num_old = 0;
            ls.head = 0;
            ls.H_diag = 1;
        } else {
            // y = g - prev_g ; s = d * t   (staged in q / x0 until accepted)
            SFX_LAP(S, 23);               // end-of-iteration tests of the previous iteration
            const T tt = (T)ls.t;
            SFX_FOR(i, D) {
                S.q[i] = S.g[i] - S.prev_g[i];
                S.x0[i] = S.d[i] * tt;
            }
            // y.s and y.y in one pass (y.y is only used when the pair is accepted)
            {
                const T* qv = S.q;
                const T* sv = S.x0;
                block_reduce2<T>(D, [=](int i) { return qv[i] * sv[i]; }, OpAdd<T>(), (T)0,
                                 [=](int i) { return qv[i] * qv[i]; }, OpAdd<T>(), (T)0, &S.red[2]);
            }
            const T ys = S.red[2], yy_pair = S.red[3];
            bool has_new = false;
            if (ys > (T)1e-10) {
                has_new = true;
                int slot;
                if (ls.num_old == H) {
                    slot = ls.head;
                    ls.head = (ls.head + 1) % H;
                    // shift of the per-pair scalars (history pops the oldest pair)
                    SFX_SYNC();
                    if (SFX_TID == 0)
                        for (int i = 0; i + 1 < H; ++i) S.ro[i] = S.ro[i + 1];
                    SFX_SYNC();
                } else {
                    slot = (ls.head + ls.num_old) % H;
                    ls.num_old += 1;
                }
                T* ys_row = hist_y + (long)slot * SFX_NP_MAX;
                T* ss_row = hist_s + (long)slot * SFX_NP_MAX;
                SFX_FOR(i, D) {
                    ys_row[i] = S.q[i];
                    ss_row[i] = S.x0[i];
                }
#ifdef __CUDACC__
                // the rows are read back by TMA bulk copies (async proxy) in two_loop_staged
                asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
                if (SFX_TID == 0) S.ro[ls.num_old - 1] = (T)1 / ys;
                ls.H_diag = ys / yy_pair;
            }
            // two-loop recursion
            SFX_LAP(S, 24);               // history update (y, s, ys, yy)
            const int k = ls.num_old;
            if (E.st->generic_two_loop == 2 && E.gram != nullptr) {
                SFX_PROF_BEGIN(tlg);
                gram_two_loop(E, S, k, ls.head, H, ls.H_diag, hist_s, hist_y, D, has_new);
                SFX_PROF_END(S, 1, tlg);
            } else
#ifdef __CUDACC__
            if (!E.st->generic_two_loop) {
                SFX_SYNC();
                SFX_PROF_BEGIN(tl);
                if (threadIdx.x < 32) {
                    if (!two_loop_staged(S, k, ls.head, H, ls.H_diag, hist_s, hist_y, D, E.stream_ws))
                        two_loop_warp(S, k, ls.head, H, ls.H_diag, hist_s, hist_y, D);
                }
                SFX_SYNC();
                SFX_PROF_END(S, 1, tl);
            } else
#endif
            {
            SFX_FOR(i, D) S.q[i] = -S.g[i];
            SFX_SYNC();
            for (int i = k - 1; i >= 0; --i) {
                const T* srow = hist_s + (long)((ls.head + i) % H) * SFX_NP_MAX;
                const T* yrow = hist_y + (long)((ls.head + i) % H) * SFX_NP_MAX;
                T a = block_dot(srow, S.q, D, &S.red[2]) * S.ro[i];
                if (SFX_TID == 0) S.al[i] = a;
                SFX_FOR(e, D) S.q[e] += -a * yrow[e];
            }
            SFX_SYNC();
            const T hd = ls.H_diag;
            SFX_FOR(i, D) S.d[i] = S.q[i] * hd;
            SFX_SYNC();
            for (int i = 0; i < k; ++i) {
                const T* srow = hist_s + (long)((ls.head + i) % H) * SFX_NP_MAX;
                const T* yrow = hist_y + (long)((ls.head + i) % H) * SFX_NP_MAX;
                T be = block_dot(yrow, S.d, D, &S.red[2]) * S.ro[i];
                T co = S.al[i] - be;
                SFX_FOR(e, D) S.d[e] += co * srow[e];
            }
            SFX_SYNC();
            }
        }
        SFX_LAP(S, 25);                   // two-loop recursion
        vcopy(S.prev_g, S.g, D);
        ls.prev_loss = loss;
        double t;
        if (ls.n_iter == 1) {
            T gs = block_abssum(S.g, D, &S.red[2]);
            T inv = (T)1 / gs;
            t = (inv < (T)1 ? (double)inv : 1.0) * st.lr;
        } else {
            t = st.lr;
        }
        // g.d and max |d| (the line search's d_norm) in one pass
        {
            const T* gv = S.g;
            const T* dv = S.d;
            block_reduce2<T>(D, [=](int i) { return gv[i] * dv[i]; }, OpAdd<T>(), (T)0,
                             [=](int i) { return sfx_abs(dv[i]); }, OpMax<T>(), (T)0, &S.red[2]);
        }
        const double gtd = (double)S.red[2], d_norm = (double)S.red[3];
        if (gtd > -st.tol_change) {
            ls.t = t;
            break;
        }
        vcopy(S.x0, S.xa, D);      // x_init
        SFX_LAP(S, 26);                   // step length, gtd, copies before the line search
        int ls_evals = 0;
        loss = strong_wolfe(E, S, &t, loss, gtd, d_norm, &ls_evals);
        SFX_LAP(S, 27);                   // line-search epilogue (after the last probe)
        {
            const T tt = (T)t;
            SFX_FOR(i, D) S.xa[i] = S.x0[i] + tt * S.d[i];
            SFX_SYNC();
        }
        ls.t = t;
        // max |g| (optimality) and max |d t| (step size test) in one pass
        {
            const T tt = (T)t;
            const T* gv = S.g;
            const T* dd = S.d;
            block_reduce2<T>(D, [=](int i) { return sfx_abs(gv[i]); }, OpMax<T>(), (T)0,
                             [=](int i) { return sfx_abs(dd[i] * tt); }, OpMax<T>(), (T)0, &S.red[2]);
        }
        const bool opt_cond = (double)S.red[2] <= st.tol_grad;
        const double step_max = (double)S.red[3];
        current_evals += ls_evals;
        if (n_iter == st.max_iter) break;
        if (current_evals >= st.max_eval) break;
        if (opt_cond) break;
        if (step_max <= st.tol_change) break;
        if (fabs(loss - ls.prev_loss) < st.tol_change) break;
    }
    scatter_active(S, D);
    return orig_loss;
}

// torch.optim.Adam step (optim_factory.py:45-48; no weight decay, no amsgrad)
template <typename T>
SFX_FN double adam_step(const EvalCtx<T>& E, Scratch<T>& S, int* step_count) {
    const SfxStage& st = *E.st;
    const int D = st.n_active;
    double loss = closure(E, S);
    *step_count += 1;
    const double b1 = st.adam_beta1, b2 = st.adam_beta2;
    const double bc1 = 1.0 - pow(b1, (double)*step_count);
    const double bc2 = 1.0 - pow(b2, (double)*step_count);
    const T step_size = (T)(st.lr / bc1);
    const T bc2s = (T)sqrt(bc2);
    SFX_FOR(i, D) {
        T gi = S.gl[i];
        T m = S.m1[i] = (T)b1 * S.m1[i] + (T)(1.0 - b1) * gi;
        T v = S.m2[i] = (T)b2 * S.m2[i] + (T)(1.0 - b2) * gi * gi;
        T denom = sfx_sqrt(v) / bc2s + (T)st.adam_eps;
        S.xa[i] = S.xa[i] - step_size * (m / denom);
    }
    SFX_SYNC();
    scatter_active(S, D);
    return loss;
}

SFX_FN double rel_change(double prev, double cur) {
    double m = fabs(prev);
    if (fabs(cur) > m) m = fabs(cur);
    if (1.0 > m) m = 1.0;
    return (prev - cur) / m;
}

// FittingMonitor.run_fitting (fitting.py:147-217) for one frame.  S.x holds the frame's
// parameters on entry and on exit.  Returns the reference's return value (prev_loss; NaN when
// the reference would return None).
template <typename T>
SFX_FN_NOINLINE double run_fitting(const EvalCtx<T>& E, Scratch<T>& S, T* hist_s, T* hist_y, int* flags) {
    SFX_ASSUME_SHARED(E);
    SFX_ASSUME_SHARED(S);
    SFX_ASSUME_SHARED_PTR(E.st);
    SFX_ASSUME_SHARED_PTR(E.M);
    SFX_ASSUME_SHARED_PTR(E.L);
    SFX_ASSUME_SHARED_PTR(E.stream_ws);
    SFX_ASSUME_SHARED_PTR(flags);
    const SfxStage& st = *E.st;
    const int D = st.n_active;
    // compact index list
    SFX_FOR(b, st.n_blocks)
        for (int i = 0; i < st.block_len[b]; ++i) S.act[st.block_start[b] + i] = st.block_off[b] + i;
    SFX_SYNC();
    SFX_FOR(i, D) {
        S.xa[i] = S.x[S.act[i]];
        S.m1[i] = 0;
        S.m2[i] = 0;
    }
    SFX_SYNC();
    LbfgsState<T> ls;
    ls.n_iter = 0; ls.num_old = 0; ls.head = 0; ls.t = 0; ls.H_diag = 1; ls.prev_loss = 0;
    int adam_steps = 0;
    double prev_loss = NAN;
    bool have_prev = false;
    for (int n = 0; n < st.maxiters; ++n) {
        double loss = st.opt_kind == SFX_OPT_ADAM ? adam_step(E, S, &adam_steps)
                                                  : lbfgs_step(E, S, ls, hist_s, hist_y);
        if (loss != loss) { if (SFX_TID == 0) *flags |= SFX_FLAG_NAN; break; }
        if (isinf(loss)) { if (SFX_TID == 0) *flags |= SFX_FLAG_INF; break; }
        if (n > 0 && have_prev && st.ftol > 0) {
            if (rel_change(prev_loss, loss) <= st.ftol) break;
        }
        // all(|max(grad of block)| < gtol) over the live parameter blocks (signed max!)
        bool all_small = true;
        for (int b = 0; b < st.n_blocks; ++b) {
            const T* gb = S.gl + st.block_start[b];
            T mx = block_reduce<T>(st.block_len[b], [=](int i) { return gb[i]; }, OpMax<T>(),
                                   (T)-INFINITY, &S.red[2]);
            if (!(fabs((double)mx) < st.gtol)) all_small = false;
        }
        if (all_small) break;
        prev_loss = loss;
        have_prev = true;
    }
    return prev_loss;
}

}  // namespace sfx
