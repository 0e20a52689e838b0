// libsfx.so -- kernels and C ABI (include/sfx.h) of the B200 SMPL-X fitting engine.
//
//   fit_stage_kernel   one block per frame; the whole FittingMonitor.run_fitting stage
//                      (closure evaluations + strong-Wolfe L-BFGS or Adam) in one launch
//   eval_kernel        one closure evaluation per frame (loss, gradient, mapped joints)
//   mesh path          full-mesh vertices for output (see sfx_mesh.cuh)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (see __graft_entry__.build)
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/sfx.h"
#include "sfx_core.cuh"
#include "sfx_model_prep.h"
#include "sfx_stream.cuh"
#include "sfx_mesh.cuh"
#include "sfx_mesh_tc.cuh"
#include "sfx_mesh_fused.cuh"
#include "sfx_metrics.cuh"
#include "sfx_ingest.cuh"

using namespace sfx;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(SFX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    } while (0)

// ------------------------------------------------------------------------------ kernels
#define SFX_THREADS 512
// -DSFX_DEV_F32_ONLY: development builds that leave the float64 instantiations of the four large
// kernels out (halves the compile time while iterating on the float32 path); never the product build
#ifdef SFX_DEV_F32_ONLY
#define SFX_NO_F64() return fail(SFX_ERR_UNSUPPORTED, "float64 kernels are left out of this development build")
#endif

template <typename T>
__device__ __forceinline__ StreamWS carve_stream(unsigned char* base, int ring_mode) {
    StreamWS ws;
    size_t ring = ring_mode ? (size_t)SFX_NWARP * SFX_NBUF * SFX_KPAD * sizeof(T)
                            : (size_t)SFX_NWARP * SFX_KPAD * sizeof(T);
    ws.ring = base;
    ws.bars = reinterpret_cast<uint64_t*>(base + ring);
    ws.fills = reinterpret_cast<unsigned int*>(base + ring + SFX_NWARP * SFX_NBUF * sizeof(uint64_t));
    ws.tl_bars = reinterpret_cast<uint64_t*>(base + ring + SFX_NWARP * SFX_NBUF * sizeof(uint64_t) +
                                             SFX_NWARP * sizeof(unsigned int));
    ws.tl_calls = reinterpret_cast<unsigned int*>(ws.tl_bars + SFX_TL_GROUPS);
    ws.ring_mode = ring_mode & 1;
    ws.tl_generic = (ring_mode >> 1) & 1;
    ws.stream_regs = (ring_mode >> 2) & 1;
    ws.tl_shfl = (ring_mode >> 3) & 1;
    ws.wide = 0;
    ws.wc = nullptr;
    return ws;
}

// Per-block copies, in shared memory, of everything the evaluation reads through references:
// the model view (its kinematic tables are indexed per lane on the critical path of the chain),
// the layout, the stage, the evaluation context and the workspace descriptors.  As kernel
// parameters they are only reachable through generic loads from the parameter window once a
// reference crosses into an out-of-line function.
template <typename T>
struct BlockCtx {
    ModelView<T> M;
    SfxLayout L;
    SfxStage st;
    EvalCtx<T> E;
    CollWS<T> CW;
    StreamWS ws;
    int flags, s_idx;
};

template <typename T>
__device__ __forceinline__ void stage_block_ctx(BlockCtx<T>& C, const ModelView<T>& M, const SfxLayout& L) {
    const int* src = reinterpret_cast<const int*>(&M);
    int* dst = reinterpret_cast<int*>(&C.M);
    for (int i = threadIdx.x; i < (int)(sizeof(ModelView<T>) / 4); i += blockDim.x) dst[i] = src[i];
    const int* ls = reinterpret_cast<const int*>(&L);
    int* ld = reinterpret_cast<int*>(&C.L);
    for (int i = threadIdx.x; i < (int)(sizeof(SfxLayout) / 4); i += blockDim.x) ld[i] = ls[i];
    __syncthreads();
}

template <typename T>
__host__ __device__ constexpr size_t scratch_bytes() {
    return (sizeof(Scratch<T>) + 1023) / 1024 * 1024;
}

template <typename T>
__device__ __forceinline__ void load_frame(const BatchView<T>& Bv, int f, Scratch<T>& S) {
    const int np = Bv.lay.np;
    for (int i = threadIdx.x; i < SFX_NP_MAX; i += blockDim.x)
        S.x[i] = i < np ? Bv.params[(size_t)f * np + i] : (T)0;
    if (threadIdx.x == 0) {
        S.n_evals = S.n_passes = 0;
        S.n_touch = 0;
        S.coll_overflow = 0;
        S.coll_max_cand = S.coll_max_touch = S.coll_max_iters = S.coll_max_hits = 0;
        for (int i = 0; i < 16; ++i) S.prof[i] = 0;
#ifdef SFX_CYCLE_PROF
        for (int i = 0; i < SFX_NLAP; ++i) S.lap[i] = 0;
        S.lap_t = clock64();
#endif
    }
    __syncthreads();
}

// workspace of the interpenetration term for this block: global slot blockIdx.x + the blend
// ring as shared work area (idle while the search runs)
template <typename T>
__device__ __forceinline__ bool block_coll_ws(const ModelView<T>& M, const BatchView<T>& Bv,
                                              const StreamWS& ws, CollWS<T>& W) {
    if (!M.coll_ready || !Bv.coll_vals) return false;
    const int ring_bytes = (int)((ws.ring_mode ? (size_t)SFX_NWARP * SFX_NBUF : (size_t)SFX_NWARP) *
                                 SFX_KPAD * sizeof(T));
    W = coll_block_ws<T>(M.V, M.F, Bv.coll_vals + (size_t)blockIdx.x * Bv.coll_vals_stride,
                         Bv.coll_idx + (size_t)blockIdx.x * Bv.coll_idx_stride, ws.ring, ring_bytes);
    return true;
}

template <typename T>
__global__ void __launch_bounds__(SFX_THREADS, 1)
fit_stage_kernel(const __grid_constant__ ModelView<T> Mp, const __grid_constant__ BatchView<T> Bv,
                 const __grid_constant__ SfxStage stp, int ring_mode, T* final_loss_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    Scratch<T>& S = *reinterpret_cast<Scratch<T>*>(smem);
    __shared__ BlockCtx<T> C;
    const ModelView<T>& M = C.M;
    StreamWS& ws = C.ws;
    const int f = Bv.frame_ids ? Bv.frame_ids[blockIdx.x] : (int)blockIdx.x;
    if (threadIdx.x == 0) {
        ws = carve_stream<T>(smem + scratch_bytes<T>(), ring_mode);
        C.flags = 0;
    }
    {
        const int* s32 = reinterpret_cast<const int*>(&stp);
        int* d32 = reinterpret_cast<int*>(&C.st);
        for (int i = threadIdx.x; i < (int)(sizeof(SfxStage) / 4); i += blockDim.x) d32[i] = s32[i];
    }
    stage_block_ctx(C, Mp, Bv.lay);
    const SfxStage& st = C.st;
    const int K = M.K;
    stream_init<T>(ws);
    load_frame(Bv, f, S);
    support_begin_frame(M, S);
    stage_setup(M, st, Bv.jw_base + (size_t)f * K, Bv.lowconf + (size_t)f * K, Bv.conf + (size_t)f * K,
                Bv.init_mask + (size_t)f * K, K, S);
    if (threadIdx.x == 0) {
        EvalCtx<T>& E = C.E;
        E.M = &C.M; E.L = &C.L; E.st = &C.st;
        E.gt = Bv.gt + (size_t)f * K * 2;
        E.conf = Bv.conf + (size_t)f * K;
        E.init_mask = Bv.init_mask + (size_t)f * K;
        E.cam = Bv.cam + (size_t)f * SFX_CAM_STRIDE;
        E.reg_pose = Bv.reg_pose ? Bv.reg_pose + (size_t)f * Bv.lay.n_pose : nullptr;
        E.stream_ws = &C.ws;
        E.gram = Bv.gram ? Bv.gram + (size_t)f * 2 * SFX_HIST * SFX_HIST : nullptr;
        E.coll = block_coll_ws(M, Bv, ws, C.CW) ? &C.CW : nullptr;
    }
    __syncthreads();
    double r = run_fitting(C.E, S, Bv.hist_s + (size_t)f * SFX_HIST * SFX_NP_MAX,
                           Bv.hist_y + (size_t)f * SFX_HIST * SFX_NP_MAX, &C.flags);
    __syncthreads();
    const int flags = C.flags | (S.coll_overflow ? SFX_FLAG_COLL_OVERFLOW : 0);
    if (threadIdx.x == 0 && Bv.coll_stat) {
        Bv.coll_stat[4 * f] = max(Bv.coll_stat[4 * f], S.coll_max_cand);
        Bv.coll_stat[4 * f + 1] = max(Bv.coll_stat[4 * f + 1], S.coll_max_touch);
        Bv.coll_stat[4 * f + 2] = max(Bv.coll_stat[4 * f + 2], S.coll_max_iters);
        Bv.coll_stat[4 * f + 3] = max(Bv.coll_stat[4 * f + 3], S.coll_max_hits);
    }
    const int np = Bv.lay.np;
    for (int i = threadIdx.x; i < np; i += blockDim.x) Bv.params[(size_t)f * np + i] = S.x[i];
    if (threadIdx.x == 0) {
        Bv.final_loss[f] = (T)r;
        if (final_loss_out) final_loss_out[f] = (T)r;
        Bv.n_evals[f] += S.n_evals;
        Bv.n_passes[f] += S.n_passes;
        Bv.flags[f] |= flags;
    }
}

template <typename T>
__global__ void __launch_bounds__(SFX_THREADS, 1)
eval_kernel(const __grid_constant__ ModelView<T> Mp, const __grid_constant__ BatchView<T> Bv,
            const __grid_constant__ SfxStage stp, int ring_mode, T* loss_out, T* grad_out,
            T* joints_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    Scratch<T>& S = *reinterpret_cast<Scratch<T>*>(smem);
    __shared__ BlockCtx<T> C;
    const ModelView<T>& M = C.M;
    StreamWS& ws = C.ws;
    const int f = blockIdx.x;
    if (threadIdx.x == 0) ws = carve_stream<T>(smem + scratch_bytes<T>(), ring_mode);
    {
        const int* s32 = reinterpret_cast<const int*>(&stp);
        int* d32 = reinterpret_cast<int*>(&C.st);
        for (int i = threadIdx.x; i < (int)(sizeof(SfxStage) / 4); i += blockDim.x) d32[i] = s32[i];
    }
    stage_block_ctx(C, Mp, Bv.lay);
    const SfxStage& st = C.st;
    const int K = M.K;
    stream_init<T>(ws);
    load_frame(Bv, f, S);
    support_begin_frame(M, S);
    // the mapped joints of EVERY keypoint are an output: no row may be skipped then
    stage_setup(M, st, Bv.jw_base + (size_t)f * K, Bv.lowconf + (size_t)f * K, Bv.conf + (size_t)f * K,
                Bv.init_mask + (size_t)f * K, K, S, joints_out != nullptr);
    if (threadIdx.x == 0) C.s_idx = block_coll_ws(M, Bv, ws, C.CW) ? 1 : 0;
    __syncthreads();
    eval_frame(M, C.L, st, Bv.gt + (size_t)f * K * 2, Bv.conf + (size_t)f * K,
               Bv.init_mask + (size_t)f * K, Bv.cam + (size_t)f * SFX_CAM_STRIDE,
               Bv.reg_pose ? Bv.reg_pose + (size_t)f * Bv.lay.n_pose : nullptr, S, &ws,
               C.s_idx ? &C.CW : nullptr);
    if (threadIdx.x == 0 && S.coll_overflow) Bv.flags[f] |= SFX_FLAG_COLL_OVERFLOW;
    if (threadIdx.x == 0 && Bv.coll_stat) {
        Bv.coll_stat[4 * f] = max(Bv.coll_stat[4 * f], S.coll_max_cand);
        Bv.coll_stat[4 * f + 1] = max(Bv.coll_stat[4 * f + 1], S.coll_max_touch);
        Bv.coll_stat[4 * f + 2] = max(Bv.coll_stat[4 * f + 2], S.coll_max_iters);
        Bv.coll_stat[4 * f + 3] = max(Bv.coll_stat[4 * f + 3], S.coll_max_hits);
    }
    const int np = Bv.lay.np;
    if (loss_out && threadIdx.x == 0) loss_out[f] = S.loss;
    if (grad_out)
        for (int i = threadIdx.x; i < np; i += blockDim.x) grad_out[(size_t)f * np + i] = S.gfull[i];
    if (joints_out)
        for (int i = threadIdx.x; i < K * 3; i += blockDim.x)
            joints_out[(size_t)f * K * 3 + i] = S.X[3 * M.joint_map[i / 3] + (i % 3)];
    if (threadIdx.x == 0) Bv.n_evals[f] += 1;
}

// Pose prologue only: skinning transforms A [B][55*12] and blend coefficients c [B][512] for the
// full-mesh path.
template <typename T>
__global__ void __launch_bounds__(SFX_THREADS, 1)
mesh_coef_kernel(const __grid_constant__ ModelView<T> Mp, const __grid_constant__ BatchView<T> Bv,
                 int use_vposer, T* Aout, T* Cout) {
    extern __shared__ __align__(1024) unsigned char smem[];
    Scratch<T>& S = *reinterpret_cast<Scratch<T>*>(smem);
    __shared__ BlockCtx<T> C;
    const ModelView<T>& M = C.M;
    if (threadIdx.x == 0) C.ws = carve_stream<T>(smem + scratch_bytes<T>(), 0);    // plain loads
    stage_block_ctx(C, Mp, Bv.lay);
    const int f = blockIdx.x;
    load_frame(Bv, f, S);
    frame_constants(M, S);                       // (no support tables: only A and c leave this kernel)
    __syncthreads();
    pose_prologue(M, C.L, S, use_vposer != 0, &C.ws, false);
    if (threadIdx.x < 32) chain_forward(M, S);
    __syncthreads();
    for (int i = threadIdx.x; i < SFX_NJ * 12; i += blockDim.x) Aout[(size_t)f * SFX_NJ * 12 + i] = S.A[i];
    for (int i = threadIdx.x; i < SFX_KPAD; i += blockDim.x) Cout[(size_t)f * SFX_KPAD + i] = S.c[i];
}


// ---- orientation bookkeeping between the camera stage and the body stages -------------------
// cv2.Rodrigues both ways, in double (reference fit_single_frame.py:527-538 runs it on the host)
__device__ void cv_rodrigues(const double* r, double* R) {
    double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < 2.220446049250313e-16) {
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    double x = r[0] / th, y = r[1] / th, z = r[2] / th, c = cos(th), s = sin(th), c1 = 1.0 - c;
    R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}
__device__ void cv_inv_rodrigues(const double* R, double* r) {
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    double th = acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
        double t = (R[0] + 1) * 0.5;
        double x = sqrt(t > 0 ? t : 0.0);
        t = (R[4] + 1) * 0.5;
        double y = sqrt(t > 0 ? t : 0.0) * (R[1] < 0 ? -1.0 : 1.0);
        t = (R[8] + 1) * 0.5;
        double z = sqrt(t > 0 ? t : 0.0) * (R[2] < 0 ? -1.0 : 1.0);
        if (fabs(x) < fabs(y) && fabs(x) < fabs(z) && ((R[5] > 0) != (y * z > 0))) z = -z;
        double k = th / sqrt(x * x + y * y + z * z);
        r[0] = x * k; r[1] = y * k; r[2] = z * k;
        return;
    }
    double k = th / (2 * s);
    r[0] = rx * k; r[1] = ry * k; r[2] = rz * k;
}

// body_model.reset_params(global_orient=orient, body_pose=pose_embedding)
// (fit_single_frame.py:546-551): every block except the orientation, the pose embedding and the
// camera translation restarts from zero.  flip = 0 remembers the camera-stage orientation;
// flip = 1 snapshots the first orientation's result and starts from Rodrigues(saved).R_y(pi).
template <typename T>
__global__ void begin_orientation_kernel(BatchView<T> Bv, int flip, T* go_saved, T* params_alt,
                                         T* loss_alt) {
    const int f = Bv.frame_ids ? Bv.frame_ids[blockIdx.x] : (int)blockIdx.x;
    const SfxLayout& L = Bv.lay;
    T* x = Bv.params + (size_t)f * L.np;
    if (flip) {
        for (int i = threadIdx.x; i < L.np; i += blockDim.x) params_alt[(size_t)f * L.np + i] = x[i];
        if (threadIdx.x == 0) loss_alt[f] = Bv.final_loss[f];
    } else if (threadIdx.x < 3) {
        go_saved[3 * f + threadIdx.x] = x[L.off_go + threadIdx.x];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L.np; i += blockDim.x) {
        const bool keep = (i >= L.off_go && i < L.off_go + 3) ||
                          (i >= L.off_pose && i < L.off_pose + L.n_pose) ||
                          (i >= L.off_camt && i < L.off_camt + 3);
        if (!keep) x[i] = 0;
    }
    if (threadIdx.x == 0) {
        if (flip) {
            double r[3] = {(double)go_saved[3 * f], (double)go_saved[3 * f + 1], (double)go_saved[3 * f + 2]};
            const double ry[3] = {0.0, 3.14159265358979323846, 0.0};
            double Ra[9], Rb[9], Rc[9], out[3];
            cv_rodrigues(r, Ra);
            cv_rodrigues(ry, Rb);
            mat3_mul(Ra, Rb, Rc);
            cv_inv_rodrigues(Rc, out);
            for (int k = 0; k < 3; ++k) x[L.off_go + k] = (T)(float)out[k];   // torch.tensor(.., float32)
        } else {
            for (int k = 0; k < 3; ++k) x[L.off_go + k] = go_saved[3 * f + k];
        }
    }
}

// results[argmin loss] (fit_single_frame.py:662-668): first orientation wins only if its loss
// is strictly lower.
template <typename T>
__global__ void select_orientation_kernel(BatchView<T> Bv, const T* params_alt, const T* loss_alt) {
    const int f = Bv.frame_ids ? Bv.frame_ids[blockIdx.x] : (int)blockIdx.x;
    const SfxLayout& L = Bv.lay;
    if (loss_alt[f] < Bv.final_loss[f]) {
        for (int i = threadIdx.x; i < L.np; i += blockDim.x)
            Bv.params[(size_t)f * L.np + i] = params_alt[(size_t)f * L.np + i];
        __syncthreads();
        if (threadIdx.x == 0) Bv.final_loss[f] = loss_alt[f];
    }
}

// fitting.guess_init (fitting.py:36-110): similar-triangles depth from the mean 3-D and 2-D edge
// lengths, one thread per frame, double arithmetic without contraction in the order of the host
// mirror (fit_frames.guess_init_depth: numpy float64), result rounded to the batch dtype.
#define SFX_GUESS_EDGES_MAX 16
struct GuessEdges { int n; int idx[2 * SFX_GUESS_EDGES_MAX]; };
template <typename T>
__global__ void guess_init_kernel(BatchView<T> Bv, const T* __restrict__ joints, int K, GuessEdges E,
                                  const double* __restrict__ focal, const double* __restrict__ len2d,
                                  const unsigned char* __restrict__ need, T* cam) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= Bv.B || !need[f]) return;
    const T* J = joints + (size_t)f * K * 3;
    double sum = 0.0;
    for (int e = 0; e < E.n; ++e) {
        const T* a = J + 3 * E.idx[2 * e];
        const T* b = J + 3 * E.idx[2 * e + 1];
        const double dx = __dsub_rn((double)a[0], (double)b[0]);
        const double dy = __dsub_rn((double)a[1], (double)b[1]);
        const double dz = __dsub_rn((double)a[2], (double)b[2]);
        const double l = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        sum = e == 0 ? l : __dadd_rn(sum, l);
    }
    const double mean3 = __ddiv_rn(sum, (double)E.n);
    const T est = (T)__dmul_rn(focal[f], __ddiv_rn(mean3, len2d[f]));
    T* x = Bv.params + (size_t)f * Bv.lay.np + Bv.lay.off_camt;
    x[0] = 0; x[1] = 0; x[2] = est;
    cam[(size_t)f * SFX_CAM_STRIDE + SFX_CAM_TZ] = est;
}

// ---- the whole per-frame pipeline in one launch -----------------------------------------------
// Persistent blocks pull frames from a counter (frames that need the second orientation are
// listed first by the host); a frame never waits for the slowest frame of a stage, and the
// second orientation of a side-view frame overlaps other frames' work.
template <typename T>
__device__ void reset_for_orientation(const SfxLayout& L, Scratch<T>& S) {
    for (int i = threadIdx.x; i < L.np; i += blockDim.x) {
        const bool keep = (i >= L.off_go && i < L.off_go + 3) ||
                          (i >= L.off_pose && i < L.off_pose + L.n_pose) ||
                          (i >= L.off_camt && i < L.off_camt + 3);
        if (!keep) S.x[i] = 0;
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(SFX_THREADS, 1)
fit_pipeline_kernel(const __grid_constant__ ModelView<T> Mp, const __grid_constant__ BatchView<T> Bv,
                    const SfxPipeline* __restrict__ P, const unsigned char* __restrict__ flip,
                    int n_frames, int* counter, int ring_mode, T* cam_loss_out, T* params_last,
                    int wide_cl, int dyn_bytes, T* alt_rows) {
    extern __shared__ __align__(1024) unsigned char smem[];
    Scratch<T>& S = *reinterpret_cast<Scratch<T>*>(smem);
    __shared__ BlockCtx<T> C;
    __shared__ WideCtl wctl;
    const ModelView<T>& M = C.M;
    StreamWS& ws = C.ws;
    const SfxStage& st = C.st;
    const int np = Bv.lay.np;
    const SfxLayout& L = C.L;
    if (threadIdx.x == 0) ws = carve_stream<T>(smem + scratch_bytes<T>(), ring_mode);
    if constexpr (sizeof(T) == 4) {
        // wide frames: this block belongs to a cluster of wide_cl CTAs (launched as such); rank 0
        // leads, the others serve it until the leader has run out of frames
        if (wide_cl > 1) {
            // programmatic dependent launch: the one-block frames' kernel, queued behind this one
            // in the same stream, may start as soon as every CTA of this grid is resident and got
            // here -- so the clusters hold their GPC-aligned SMs before the other grid spreads out
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
            const int rank = (int)cluster_ctarank();
            if (threadIdx.x == 0) {
                mbar_init(&wctl.go_bar, 1);
                mbar_init(&wctl.done_bar, wide_cl - 1);
                mbar_init(&wctl.load_bar, 1);
                wctl.cmd = 0;
                wctl.n_rows = 0;
                wctl.done_phase = 0;
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            cluster_sync_all();                  // every CTA's barriers exist before anyone signals
            if (rank != 0) {
                wide_helper_main(reinterpret_cast<const float*>(Mp.PK), smem, (size_t)dyn_bytes, &wctl, wide_cl);
                cluster_sync_all();
                return;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                ws.wide = wide_cl;
                ws.wc = &wctl;
            }
        }
    }
    stage_block_ctx(C, Mp, Bv.lay);
    const int K = M.K;
    stream_init<T>(ws);
    auto load_stage = [&](const SfxStage* src) {
        __syncthreads();
        const int* s32 = reinterpret_cast<const int*>(src);
        int* d32 = reinterpret_cast<int*>(&C.st);
        for (int i = threadIdx.x; i < (int)(sizeof(SfxStage) / 4); i += blockDim.x) d32[i] = s32[i];
        __syncthreads();
    };
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            C.s_idx = atomicAdd(counter, 1);
            C.flags = 0;
        }
        __syncthreads();
        const int idx = C.s_idx;
        if (idx >= n_frames) break;
        const int f = Bv.frame_ids ? Bv.frame_ids[idx] : idx;
        load_frame(Bv, f, S);
        T* alt = alt_rows + (size_t)f * np;      // first orientation's result (global: one row per frame)
        const long long t_frame = clock64();     // SM cycles this block spends on the frame (slot 4 of prof)
        SFX_PROF_BEGIN(total);
        support_begin_frame(M, S);
        if (threadIdx.x == 0) {
            EvalCtx<T>& E = C.E;
            E.M = &C.M; E.L = &C.L; E.st = &C.st;
            E.gt = Bv.gt + (size_t)f * K * 2;
            E.conf = Bv.conf + (size_t)f * K;
            E.init_mask = Bv.init_mask + (size_t)f * K;
            E.cam = Bv.cam + (size_t)f * SFX_CAM_STRIDE;
            E.reg_pose = Bv.reg_pose ? Bv.reg_pose + (size_t)f * Bv.lay.n_pose : nullptr;
            E.stream_ws = &C.ws;
            E.gram = Bv.gram ? Bv.gram + (size_t)f * 2 * SFX_HIST * SFX_HIST : nullptr;
            E.coll = block_coll_ws(M, Bv, ws, C.CW) ? &C.CW : nullptr;
        }
        __syncthreads();
        const EvalCtx<T>& E = C.E;
        T* hs = Bv.hist_s + (size_t)f * SFX_HIST * SFX_NP_MAX;
        T* hy = Bv.hist_y + (size_t)f * SFX_HIST * SFX_NP_MAX;
        // stage C: camera translation + global orientation (fit_single_frame.py:473-496)
        load_stage(&P->cam);
        stage_setup(M, st, Bv.jw_base + (size_t)f * K, Bv.lowconf + (size_t)f * K, Bv.conf + (size_t)f * K,
                Bv.init_mask + (size_t)f * K, K, S);
        double r = run_fitting(E, S, hs, hy, &C.flags);
        __syncthreads();
        if (threadIdx.x == 0 && cam_loss_out) cam_loss_out[f] = (T)r;
        T go0[3] = {S.x[L.off_go], S.x[L.off_go + 1], S.x[L.off_go + 2]};
        const int n_orient = flip && flip[f] ? 2 : 1;
        double loss0 = 0;
        for (int o = 0; o < n_orient; ++o) {
            __syncthreads();
            if (o == 1) {
                for (int i = threadIdx.x; i < np; i += blockDim.x) alt[i] = S.x[i];
                loss0 = r;
                __syncthreads();
                if (threadIdx.x == 0) {
                    double rv[3] = {(double)go0[0], (double)go0[1], (double)go0[2]};
                    const double ry[3] = {0.0, 3.14159265358979323846, 0.0};
                    double Ra[9], Rb[9], Rc[9], out[3];
                    cv_rodrigues(rv, Ra);
                    cv_rodrigues(ry, Rb);
                    mat3_mul(Ra, Rb, Rc);
                    cv_inv_rodrigues(Rc, out);
                    for (int k = 0; k < 3; ++k) S.x[L.off_go + k] = (T)(float)out[k];
                }
                __syncthreads();
            }
            reset_for_orientation(L, S);
            for (int si = 0; si < P->n_stages; ++si) {
                load_stage(&P->body[si]);
                stage_setup(M, st, Bv.jw_base + (size_t)f * K, Bv.lowconf + (size_t)f * K, Bv.conf + (size_t)f * K,
                Bv.init_mask + (size_t)f * K, K, S);
                r = run_fitting(E, S, hs, hy, &C.flags);
            }
        }
        __syncthreads();
        // the mesh written to vertices.ply is the LAST orientation's (fit_single_frame.py:671-676)
        if (params_last)
            for (int i = threadIdx.x; i < np; i += blockDim.x) params_last[(size_t)f * np + i] = S.x[i];
        // results[argmin loss] (fit_single_frame.py:662-668)
        const bool restore = n_orient == 2 && (loss0 < r);
        for (int i = threadIdx.x; i < np; i += blockDim.x)
            Bv.params[(size_t)f * np + i] = restore ? alt[i] : S.x[i];
        if (threadIdx.x == 0) {
            Bv.final_loss[f] = (T)(restore ? loss0 : r);
            Bv.n_evals[f] += S.n_evals;
            Bv.n_passes[f] += S.n_passes;
            Bv.flags[f] |= C.flags | (S.coll_overflow ? SFX_FLAG_COLL_OVERFLOW : 0);
            if (Bv.coll_stat) {
                Bv.coll_stat[4 * f] = max(Bv.coll_stat[4 * f], S.coll_max_cand);
                Bv.coll_stat[4 * f + 1] = max(Bv.coll_stat[4 * f + 1], S.coll_max_touch);
                Bv.coll_stat[4 * f + 2] = max(Bv.coll_stat[4 * f + 2], S.coll_max_iters);
                Bv.coll_stat[4 * f + 3] = max(Bv.coll_stat[4 * f + 3], S.coll_max_hits);
            }
#ifdef SFX_CYCLE_PROF
            S.prof[4] += clock64() - _t_total;
            for (int i = 0; i < 16; ++i) Bv.prof[(size_t)f * 64 + i] = S.prof[i];
            for (int i = 0; i < SFX_NLAP; ++i) Bv.prof[(size_t)f * 64 + 16 + i] = S.lap[i];
#else
            if (Bv.prof) Bv.prof[(size_t)f * 64 + 4] = clock64() - t_frame;
#endif
        }
    }
    if constexpr (sizeof(T) == 4) {
        if (wide_cl > 1) {                       // leader: release the helpers, leave together
            __syncthreads();
            wide_post(ws, SFX_WIDE_EXIT, 0);
            cluster_sync_all();
        }
    }
}

// ---- diagnostics: L2 -> SM read bandwidth of this device ------------------------------------
// Every block sweeps the same L2-resident buffer with 16-byte loads that bypass L1 (ld.global.cg);
// bench.py reports the blend-row traffic of the fit kernel against this measured figure.
__global__ void __launch_bounds__(512)
l2_read_kernel(const uint4* __restrict__ buf, size_t n16, int iters, unsigned int* sink) {
    unsigned int acc = 0;
    const size_t start = ((size_t)blockIdx.x * 7919u * blockDim.x) % n16;     // blocks start apart
    for (int it = 0; it < iters; ++it)
        for (size_t i0 = threadIdx.x; i0 < n16; i0 += (size_t)blockDim.x * 4) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                size_t i = start + i0 + (size_t)u * blockDim.x;
                if (i >= n16) i -= n16;
                if (i >= n16) i -= n16;
                v[u] = __ldcg(buf + i);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        }
    if (acc == 0x9e3779b9u) *sink = acc;         // keeps the loads alive
}

// ------------------------------------------------------------------------------ handles
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = n;
        return n ? cudaMalloc(&p, n) : cudaSuccess;
    }
    template <typename U>
    cudaError_t upload(const std::vector<U>& v) {
        cudaError_t e = alloc(v.size() * sizeof(U));
        if (e != cudaSuccess) return e;
        return v.empty() ? cudaSuccess : cudaMemcpy(p, v.data(), bytes, cudaMemcpyHostToDevice);
    }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

struct sfx_model {
    int use_double = 0;
    int V = 0, F = 0, K = 0, NB = 0, NE = 0, NH = 0;
    ModelView<float> vf;
    ModelView<double> vd;
    DevBuf PK, vt, J0, JS, Wd, hand_l, hand_r, pose_mean, sv_vid, lmk_bary, dyn_vid, dyn_bary,
        joint_map, inv_ptr, inv_idx, faces, dyn_pack, gmm_means, gmm_prec, gmm_logw, vp_w1, vp_b1, vp_w2,
        vp_b2, vp_w3, vp_b3, part_ptr, part_faces, face_part, part_allow, vf_ptr, vf_idx, sk_ptr,
        sk_j, sk_w, cl_ptr, part_cl_ptr;
    std::vector<int> faces_host;
    std::vector<double> vt_host;
    int device = 0;
    int num_sms = 0;
    MeshPlan mesh;        // TMA descriptor of the blend matrix for the tensor-core mesh kernel
    FusedPlan fused;      // descriptors of the fused blend + skinning kernel (sfx_mesh_fused.cuh)
    DevBuf whi, wlo;      // skinning weights [Vpad][64], float16 hi / lo split
    DevBuf pk16;          // float16 copy of PK: K1 operand of the fused mesh kernel (32 MB)
};

template <typename T>
static int upload_model(const sfx_model_desc& d, sfx_model* m, ModelView<T>& view) {
    HostModel<T> h;
    std::string e = prepare_model(d, h);
    if (!e.empty()) return fail(SFX_ERR_ARG, e);
    m->V = h.V; m->F = h.F; m->K = h.K; m->NB = h.NB; m->NE = h.NE; m->NH = h.NH;
    h.fill_scalars(view);
    CUDA_TRY(m->PK.upload(h.PK));
    CUDA_TRY(m->vt.upload(h.vt));
    CUDA_TRY(m->J0.upload(h.J0));
    CUDA_TRY(m->JS.upload(h.JS));
    CUDA_TRY(m->Wd.upload(h.Wd));
    CUDA_TRY(m->hand_l.upload(h.hand_l));
    CUDA_TRY(m->hand_r.upload(h.hand_r));
    CUDA_TRY(m->pose_mean.upload(h.pose_mean));
    CUDA_TRY(m->sv_vid.upload(h.sv_vid));
    CUDA_TRY(m->lmk_bary.upload(h.lmk_bary));
    CUDA_TRY(m->dyn_vid.upload(h.dyn_vid));
    CUDA_TRY(m->dyn_bary.upload(h.dyn_bary));
    CUDA_TRY(m->joint_map.upload(h.joint_map));
    CUDA_TRY(m->inv_ptr.upload(h.inv_ptr));
    CUDA_TRY(m->inv_idx.upload(h.inv_idx));
    CUDA_TRY(m->faces.upload(h.faces));
    CUDA_TRY(m->sk_ptr.upload(h.sk_ptr));
    CUDA_TRY(m->sk_j.upload(h.sk_j));
    CUDA_TRY(m->sk_w.upload(h.sk_w));
    CUDA_TRY(m->dyn_pack.upload(h.dyn_pack));
    view.dyn_pack = (const DynRowPack<T>*)m->dyn_pack.p;
    view.sk_ptr = (const int*)m->sk_ptr.p; view.sk_j = (const unsigned char*)m->sk_j.p;
    view.sk_w = (const T*)m->sk_w.p;
    m->faces_host = h.faces;
    m->vt_host.assign(h.vt.begin(), h.vt.end());
    view.PK = (const T*)m->PK.p; view.vt = (const T*)m->vt.p; view.J0 = (const T*)m->J0.p;
    view.JS = (const T*)m->JS.p; view.Wd = (const T*)m->Wd.p;
    view.hand_l = (const T*)m->hand_l.p; view.hand_r = (const T*)m->hand_r.p;
    view.pose_mean = (const T*)m->pose_mean.p; view.sv_vid = (const int*)m->sv_vid.p;
    view.lmk_bary = (const T*)m->lmk_bary.p; view.dyn_vid = (const int*)m->dyn_vid.p;
    view.dyn_bary = (const T*)m->dyn_bary.p; view.joint_map = (const int*)m->joint_map.p;
    view.inv_ptr = (const int*)m->inv_ptr.p; view.inv_idx = (const int*)m->inv_idx.p;
    return SFX_OK;
}

struct sfx_batch {
    const sfx_model* m = nullptr;
    int B = 0, use_vposer = 0;
    SfxLayout lay;
    size_t es = 4;        // element size of the batch dtype
    DevBuf params, gt, conf, jw, lowconf, init_mask, cam, reg_pose, hist_s, hist_y, gram, final_loss,
        n_evals, n_passes, flags, Acoef, Ccoef, c16, vposed, ahi, alo, go_saved, params_alt, loss_alt, pipe, counter,
        cam_loss, params_last, prof, coll_vals, coll_idx, coll_stat, guess;
    bool last_valid = false;
    bool has_reg = false;

    std::vector<unsigned char> stage_host;     // host staging for set_targets
    template <typename T>
    BatchView<T> view(const int* frame_ids) const {
        BatchView<T> v;
        v.B = B; v.lay = lay;
        v.params = (T*)params.p; v.gt = (const T*)gt.p; v.conf = (const T*)conf.p;
        v.jw_base = (const T*)jw.p; v.lowconf = (const unsigned char*)lowconf.p;
        v.init_mask = (const unsigned char*)init_mask.p; v.cam = (const T*)cam.p;
        v.reg_pose = has_reg ? (const T*)reg_pose.p : nullptr;
        v.hist_s = (T*)hist_s.p; v.hist_y = (T*)hist_y.p; v.gram = (T*)gram.p; v.final_loss = (T*)final_loss.p;
        v.coll_vals = (T*)coll_vals.p; v.coll_idx = (unsigned short*)coll_idx.p;
        v.coll_stat = (int*)coll_stat.p;
        v.coll_vals_stride = coll_vals_per_block(m->V, m->F);
        v.coll_idx_stride = coll_idx_per_block(m->V, m->F);
        v.prof = (long long*)prof.p; v.n_evals = (int*)n_evals.p; v.n_passes = (int*)n_passes.p; v.flags = (int*)flags.p; v.frame_ids = frame_ids;
        return v;
    }
};

template <typename T>
static size_t fit_smem(int ring_mode) {
    return scratch_bytes<T>() + stream_smem_bytes<T>(ring_mode);
}

// bit 0: TMA ring (float32) vs plain loads; bit 1: SFX_TWO_LOOP_GENERIC=1 selects the
// generic-pointer staged two-loop recursion (A/B check of the shared-address version)
static int ring_mode_for(const sfx_model* m) {
    const char* e = getenv("SFX_STREAM_DIRECT");
    if (e && e[0] == '1') return 0;
    if (m->use_double) return 0;
    const char* g = getenv("SFX_TWO_LOOP_GENERIC");
    const char* r = getenv("SFX_STREAM_REGS");
    const char* h = getenv("SFX_TWO_LOOP_SHFL");
    return 1 | ((g && g[0] == '1') ? 2 : 0) | ((r && r[0] == '1') ? 4 : 0) | ((h && h[0] == '1') ? 8 : 0);
}

template <typename T>
static int upload_vposer(sfx_model* m, ModelView<T>& v, const T* w1, const T* b1, const T* w2,
                         const T* b2, const T* w3, const T* b3) {
    std::vector<T> w3p((size_t)128 * SFX_KPAD, (T)0), b3p(128, (T)0);
    for (size_t i = 0; i < (size_t)126 * SFX_KPAD; ++i) w3p[i] = w3[i];
    for (int i = 0; i < 126; ++i) b3p[i] = b3[i];
    std::vector<T> vw1(w1, w1 + 512 * SFX_NLATENT), vb1(b1, b1 + 512), vw2(w2, w2 + (size_t)512 * 512),
        vb2(b2, b2 + 512);
    CUDA_TRY(m->vp_w1.upload(vw1)); CUDA_TRY(m->vp_b1.upload(vb1));
    CUDA_TRY(m->vp_w2.upload(vw2)); CUDA_TRY(m->vp_b2.upload(vb2));
    CUDA_TRY(m->vp_w3.upload(w3p)); CUDA_TRY(m->vp_b3.upload(b3p));
    v.vp_w1 = (const T*)m->vp_w1.p; v.vp_b1 = (const T*)m->vp_b1.p; v.vp_w2 = (const T*)m->vp_w2.p;
    v.vp_b2 = (const T*)m->vp_b2.p; v.vp_w3 = (const T*)m->vp_w3.p; v.vp_b3 = (const T*)m->vp_b3.p;
    v.vp_ready = 1;
    return SFX_OK;
}


// ------------------------------------------------------------------------------ C ABI
extern "C" {

const char* sfx_last_error(void) { return g_err.c_str(); }

int sfx_aligned_errors(const void* est_dev, const void* gt_dev, const int32_t* idx_dev, int32_t B,
                       int32_t N, int32_t n, int32_t mode, int32_t hip0, int32_t hip1,
                       int32_t use_double, void* err_dev, void* transform_dev, void* stream) {
    if (!est_dev || !gt_dev || !err_dev || B < 0 || N < 1 || n < 1 || (!idx_dev && n != N))
        return fail(SFX_ERR_ARG, "bad argument");
    if (mode < SFX_ALIGN_NONE || mode > SFX_ALIGN_SCALE) return fail(SFX_ERR_ARG, "unknown alignment");
    if (mode == SFX_ALIGN_PELVIS && (hip0 < 0 || hip1 < 0 || hip0 >= n || hip1 >= n))
        return fail(SFX_ERR_ARG, "pelvis alignment: hip index out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(SFX_ERR_CUDA, "no CUDA device visible: libsfx has no CPU path");
    std::string e = use_double
        ? aligned_errors<double>((const double*)est_dev, (const double*)gt_dev, idx_dev, B, N, n, mode,
                                 hip0, hip1, (double*)err_dev, (double*)transform_dev, (cudaStream_t)stream)
        : aligned_errors<float>((const float*)est_dev, (const float*)gt_dev, idx_dev, B, N, n, mode,
                                hip0, hip1, (float*)err_dev, (float*)transform_dev, (cudaStream_t)stream);
    if (!e.empty()) return fail(SFX_ERR_CUDA, e);
    return SFX_OK;
}
int sfx_version(void) { return 100; }

int sfx_model_create(const sfx_model_desc* desc, sfx_model** out) {
    if (!desc || !out) return fail(SFX_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(SFX_ERR_CUDA, "no CUDA device visible: libsfx has no CPU path");
    sfx_model* m = new sfx_model();
    m->use_double = desc->use_double ? 1 : 0;
    cudaGetDevice(&m->device);
    cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, m->device);
    int rc = m->use_double ? upload_model<double>(*desc, m, m->vd) : upload_model<float>(*desc, m, m->vf);
    if (rc == SFX_OK && !m->use_double) {
        std::string e = mesh_plan_create(m->mesh, (const float*)m->PK.p, m->V);
        if (e.empty()) {
            // skinning weights for the tensor-core kernel: 64 padded joints, split x = hi + lo
            const int Vpad = (m->V + FU_TV - 1) / FU_TV * FU_TV;
            std::vector<__half> whi((size_t)Vpad * FU_WJ, __float2half_rn(0.f)), wlo((size_t)Vpad * FU_WJ, __float2half_rn(0.f));
            for (int v = 0; v < m->V; ++v)
                for (int j = 0; j < SFX_NJ; ++j)
                    f16_split(desc->arrays_float64
                                   ? (float)static_cast<const double*>(desc->lbs_weights)[(size_t)v * SFX_NJ + j]
                                   : static_cast<const float*>(desc->lbs_weights)[(size_t)v * SFX_NJ + j],
                               &whi[(size_t)v * FU_WJ + j],
                               &wlo[(size_t)v * FU_WJ + j]);
            cudaError_t ce = m->whi.upload(whi);
            if (ce == cudaSuccess) ce = m->wlo.upload(wlo);
            if (ce == cudaSuccess) ce = m->pk16.alloc((size_t)3 * m->V * SFX_KPAD * sizeof(__half));
            if (ce != cudaSuccess) e = std::string("upload skinning weights: ") + cudaGetErrorString(ce);
            else e = fused_plan_create(m->fused, (const float*)m->PK.p, (__half*)m->pk16.p, m->V, (const __half*)m->whi.p,
                                       (const __half*)m->wlo.p, Vpad);
        }
        if (!e.empty()) rc = fail(SFX_ERR_CUDA, e);
    }
    if (rc != SFX_OK) {
        delete m;
        return rc;
    }
    *out = m;
    return SFX_OK;
}

void sfx_model_destroy(sfx_model* m) { delete m; }

int sfx_model_set_vposer(sfx_model* m, const void* fc1_w, const void* fc1_b, const void* fc2_w,
                         const void* fc2_b, const void* out_w, const void* out_b) {
    if (!m || !fc1_w || !fc1_b || !fc2_w || !fc2_b || !out_w || !out_b)
        return fail(SFX_ERR_ARG, "null argument");
    if (m->use_double)
        return upload_vposer<double>(m, m->vd, (const double*)fc1_w, (const double*)fc1_b,
                                     (const double*)fc2_w, (const double*)fc2_b,
                                     (const double*)out_w, (const double*)out_b);
    return upload_vposer<float>(m, m->vf, (const float*)fc1_w, (const float*)fc1_b,
                                (const float*)fc2_w, (const float*)fc2_b, (const float*)out_w,
                                (const float*)out_b);
}

int sfx_model_set_gmm(sfx_model* m, int32_t num_gaussians, int32_t dim, const void* means,
                      const void* precisions, const void* log_nll_weights) {
    if (!m || !means || !precisions || !log_nll_weights) return fail(SFX_ERR_ARG, "null argument");
    if (num_gaussians < 1 || num_gaussians > 16 || dim < 1 || num_gaussians * dim > SFX_KPAD)
        return fail(SFX_ERR_ARG, "mixture prior: need 1..16 components and components x dim <= 512");
    const size_t es = m->use_double ? 8 : 4;
    CUDA_TRY(m->gmm_means.alloc((size_t)num_gaussians * dim * es));
    CUDA_TRY(m->gmm_prec.alloc((size_t)num_gaussians * dim * dim * es));
    CUDA_TRY(m->gmm_logw.alloc((size_t)num_gaussians * es));
    CUDA_TRY(cudaMemcpy(m->gmm_means.p, means, m->gmm_means.bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(m->gmm_prec.p, precisions, m->gmm_prec.bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(m->gmm_logw.p, log_nll_weights, m->gmm_logw.bytes, cudaMemcpyHostToDevice));
    if (m->use_double) {
        m->vd.gmm_M = num_gaussians; m->vd.gmm_D = dim;
        m->vd.gmm_means = (const double*)m->gmm_means.p; m->vd.gmm_prec = (const double*)m->gmm_prec.p;
        m->vd.gmm_logw = (const double*)m->gmm_logw.p;
    } else {
        m->vf.gmm_M = num_gaussians; m->vf.gmm_D = dim;
        m->vf.gmm_means = (const float*)m->gmm_means.p; m->vf.gmm_prec = (const float*)m->gmm_prec.p;
        m->vf.gmm_logw = (const float*)m->gmm_logw.p;
    }
    return SFX_OK;
}

static int set_collision_impl(sfx_model* m, const int32_t* faces_segm, const int32_t* faces_parents,
                              const int32_t* ign_part_pairs, int32_t n_ign_pairs, bool unfiltered);

int sfx_model_set_collision(sfx_model* m, const int32_t* faces_segm, const int32_t* faces_parents,
                            const int32_t* ign_part_pairs, int32_t n_ign_pairs) {
    if (!m || !faces_segm || !faces_parents || (n_ign_pairs > 0 && !ign_part_pairs) || n_ign_pairs < 0)
        return fail(SFX_ERR_ARG, "bad argument");
    return set_collision_impl(m, faces_segm, faces_parents, ign_part_pairs, n_ign_pairs, false);
}

int sfx_model_set_collision_unfiltered(sfx_model* m, const int32_t* faces_group) {
    if (!m || !faces_group) return fail(SFX_ERR_ARG, "bad argument");
    return set_collision_impl(m, faces_group, nullptr, nullptr, 0, true);
}

static int set_collision_impl(sfx_model* m, const int32_t* faces_segm, const int32_t* faces_parents,
                              const int32_t* ign_part_pairs, int32_t n_ign_pairs, bool unfiltered) {
    HostCollision c;
    std::string e = prepare_collision(m->V, m->F, m->faces_host.data(), m->vt_host.data(), faces_segm,
                                      faces_parents, ign_part_pairs, n_ign_pairs, c, unfiltered);
    if (!e.empty()) return fail(SFX_ERR_ARG, e);
    CUDA_TRY(m->part_ptr.upload(c.part_ptr));
    CUDA_TRY(m->part_faces.upload(c.part_faces));
    CUDA_TRY(m->face_part.upload(c.face_part));
    CUDA_TRY(m->part_allow.upload(c.part_allow));
    CUDA_TRY(m->vf_ptr.upload(c.vf_ptr));
    CUDA_TRY(m->vf_idx.upload(c.vf_idx));
    CUDA_TRY(m->cl_ptr.upload(c.cl_ptr));
    CUDA_TRY(m->part_cl_ptr.upload(c.part_cl_ptr));
    auto fill = [&](auto& v) {
        v.n_clusters = c.n_clusters; v.cl_ptr = (const int*)m->cl_ptr.p;
        v.part_cl_ptr = (const int*)m->part_cl_ptr.p;
        v.coll_ready = 1; v.F = m->F; v.n_parts = c.n_parts; v.faces = (const int*)m->faces.p;
        v.part_ptr = (const int*)m->part_ptr.p; v.part_faces = (const int*)m->part_faces.p;
        v.face_part = (const unsigned char*)m->face_part.p;
        v.part_allow = (const unsigned long long*)m->part_allow.p;
        v.vf_ptr = (const int*)m->vf_ptr.p; v.vf_idx = (const int*)m->vf_idx.p;
    };
    if (m->use_double) fill(m->vd); else fill(m->vf);
    return SFX_OK;
}

int sfx_batch_enable_collisions(sfx_batch* b) {
    if (!b) return fail(SFX_ERR_ARG, "null argument");
    const int ready = b->m->use_double ? b->m->vd.coll_ready : b->m->vf.coll_ready;
    if (!ready) return fail(SFX_ERR_ARG, "interpenetration: call sfx_model_set_collision first");
    if (b->coll_vals.p) return SFX_OK;
    // one slot per block; no launch uses more blocks than frames
    CUDA_TRY(b->coll_vals.alloc((size_t)b->B * coll_vals_per_block(b->m->V, b->m->F) * b->es));
    CUDA_TRY(b->coll_idx.alloc((size_t)b->B * coll_idx_per_block(b->m->V, b->m->F) * sizeof(unsigned short)));
    CUDA_TRY(b->coll_stat.alloc((size_t)b->B * 4 * sizeof(int)));
    CUDA_TRY(cudaMemset(b->coll_stat.p, 0, b->coll_stat.bytes));
    return SFX_OK;
}

int32_t* sfx_batch_coll_stat_dev(sfx_batch* b) { return b ? (int32_t*)b->coll_stat.p : nullptr; }

int sfx_batch_create(const sfx_model* m, int32_t B, int32_t use_vposer, sfx_batch** out) {
    if (!m || !out || B < 1) return fail(SFX_ERR_ARG, "bad argument");
    if (use_vposer && !(m->use_double ? m->vd.vp_ready : m->vf.vp_ready))
        return fail(SFX_ERR_ARG, "use_vposer: call sfx_model_set_vposer first");
    sfx_batch* b = new sfx_batch();
    b->m = m; b->B = B; b->use_vposer = use_vposer;
    b->lay = make_layout(m->NB, m->NE, m->NH, use_vposer);
    if (b->lay.np > SFX_NP_MAX) {
        // e.g. use_pca = False (45 + 45 hand components) with 25+ shape / expression coefficients
        const int np = b->lay.np;
        delete b;
        return fail(SFX_ERR_UNSUPPORTED, "parameter vector of " + std::to_string(np) + " entries exceeds SFX_NP_MAX (" +
                    std::to_string(SFX_NP_MAX) + "): use fewer shape / expression coefficients or hand PCA");
    }
    b->es = m->use_double ? 8 : 4;
    const size_t es = b->es, K = m->K;
#define ALLOC(buf, n)                                                         \
    do {                                                                      \
        cudaError_t _e = b->buf.alloc(n);                                     \
        if (_e == cudaSuccess && (n)) _e = cudaMemset(b->buf.p, 0, n);        \
        if (_e != cudaSuccess) {                                              \
            delete b;                                                         \
            return fail(SFX_ERR_CUDA, std::string("alloc " #buf ": ") + cudaGetErrorString(_e)); \
        }                                                                     \
    } while (0)
    ALLOC(params, (size_t)B * b->lay.np * es);
    ALLOC(gt, (size_t)B * K * 2 * es);
    ALLOC(conf, (size_t)B * K * es);
    ALLOC(jw, (size_t)B * K * es);
    ALLOC(lowconf, (size_t)B * K);
    ALLOC(init_mask, (size_t)B * K);
    ALLOC(cam, (size_t)B * SFX_CAM_STRIDE * es);
    ALLOC(reg_pose, (size_t)B * b->lay.n_pose * es);
    ALLOC(hist_s, (size_t)B * SFX_HIST * SFX_NP_MAX * es);
    ALLOC(hist_y, (size_t)B * SFX_HIST * SFX_NP_MAX * es);
    ALLOC(gram, (size_t)B * 2 * SFX_HIST * SFX_HIST * es);
    ALLOC(final_loss, (size_t)B * es);
    ALLOC(n_evals, (size_t)B * sizeof(int));
    ALLOC(n_passes, (size_t)B * sizeof(int));
    ALLOC(flags, (size_t)B * sizeof(int));
    ALLOC(Acoef, (size_t)B * SFX_NJ * 12 * es);
    ALLOC(Ccoef, (size_t)mesh_padded_frames(B) * SFX_KPAD * es);
    ALLOC(vposed, (size_t)mesh_padded_frames(B) * 3 * m->V * es);
    ALLOC(c16, (size_t)mesh_padded_frames(B) * SFX_KPAD * sizeof(__half));
    ALLOC(ahi, (size_t)mesh_padded_frames(B) * 12 * FU_WJ * sizeof(__half));
    ALLOC(alo, (size_t)mesh_padded_frames(B) * 12 * FU_WJ * sizeof(__half));
    ALLOC(go_saved, (size_t)B * 3 * es);
    ALLOC(params_alt, (size_t)B * b->lay.np * es);
    ALLOC(loss_alt, (size_t)B * es);
    ALLOC(pipe, sizeof(SfxPipeline));
    ALLOC(counter, 64);
    ALLOC(cam_loss, (size_t)B * es);
    ALLOC(prof, (size_t)B * 64 * sizeof(long long));
    ALLOC(params_last, (size_t)B * b->lay.np * es);
    ALLOC(guess, (size_t)B * (2 * sizeof(double) + 8));
#undef ALLOC
    *out = b;
    return SFX_OK;
}

void sfx_batch_destroy(sfx_batch* b) { delete b; }

int sfx_batch_layout(const sfx_batch* b, SfxLayout* out) {
    if (!b || !out) return fail(SFX_ERR_ARG, "null argument");
    *out = b->lay;
    return SFX_OK;
}

int sfx_batch_set_targets(sfx_batch* b, const void* keypoints, const void* joint_weights,
                          const uint8_t* lowconf, const uint8_t* init_mask, const void* cam,
                          const void* reg_pose, void* stream) {
    if (!b || !keypoints || !joint_weights || !lowconf || !init_mask || !cam)
        return fail(SFX_ERR_ARG, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t es = b->es, K = b->m->K, B = b->B;
    // split [B,K,3] keypoints into gt [B,K,2] and conf [B,K] on the host (tiny), then copy
    b->stage_host.resize(B * K * 3 * es);
    unsigned char* gt_h = b->stage_host.data();
    unsigned char* conf_h = gt_h + B * K * 2 * es;
    for (size_t i = 0; i < B * K; ++i) {
        const unsigned char* src = (const unsigned char*)keypoints + i * 3 * es;
        memcpy(gt_h + i * 2 * es, src, 2 * es);
        memcpy(conf_h + i * es, src + 2 * es, es);
    }
    CUDA_TRY(cudaMemcpyAsync(b->gt.p, gt_h, B * K * 2 * es, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->conf.p, conf_h, B * K * es, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->jw.p, joint_weights, B * K * es, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->lowconf.p, lowconf, B * K, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->init_mask.p, init_mask, B * K, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->cam.p, cam, B * SFX_CAM_STRIDE * es, cudaMemcpyHostToDevice, s));
    b->has_reg = reg_pose != nullptr;
    if (reg_pose)
        CUDA_TRY(cudaMemcpyAsync(b->reg_pose.p, reg_pose, B * b->lay.n_pose * es,
                                 cudaMemcpyHostToDevice, s));
    return SFX_OK;
}

int sfx_batch_set_targets_dev(sfx_batch* b, const void* gt, const void* conf,
                              const void* joint_weights, const uint8_t* lowconf,
                              const uint8_t* init_mask, const void* cam, const void* reg_pose) {
    if (!b || !gt || !conf || !joint_weights || !lowconf || !init_mask || !cam)
        return fail(SFX_ERR_ARG, "null argument");
    const size_t es = b->es, K = b->m->K, B = b->B;
    CUDA_TRY(cudaMemcpy(b->gt.p, gt, B * K * 2 * es, cudaMemcpyDeviceToDevice));
    CUDA_TRY(cudaMemcpy(b->conf.p, conf, B * K * es, cudaMemcpyDeviceToDevice));
    CUDA_TRY(cudaMemcpy(b->jw.p, joint_weights, B * K * es, cudaMemcpyDeviceToDevice));
    CUDA_TRY(cudaMemcpy(b->lowconf.p, lowconf, B * K, cudaMemcpyDeviceToDevice));
    CUDA_TRY(cudaMemcpy(b->init_mask.p, init_mask, B * K, cudaMemcpyDeviceToDevice));
    CUDA_TRY(cudaMemcpy(b->cam.p, cam, B * SFX_CAM_STRIDE * es, cudaMemcpyDeviceToDevice));
    b->has_reg = reg_pose != nullptr;
    if (reg_pose)
        CUDA_TRY(cudaMemcpy(b->reg_pose.p, reg_pose, B * b->lay.n_pose * es, cudaMemcpyDeviceToDevice));
    return SFX_OK;
}

int sfx_batch_set_params(sfx_batch* b, const void* host_params, void* stream) {
    if (!b || !host_params) return fail(SFX_ERR_ARG, "null argument");
    CUDA_TRY(cudaMemcpyAsync(b->params.p, host_params, (size_t)b->B * b->lay.np * b->es,
                             cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return SFX_OK;
}

int sfx_batch_get_params(const sfx_batch* b, void* host_params, void* stream) {
    if (!b || !host_params) return fail(SFX_ERR_ARG, "null argument");
    CUDA_TRY(cudaMemcpyAsync(host_params, b->params.p, (size_t)b->B * b->lay.np * b->es,
                             cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return SFX_OK;
}

void* sfx_batch_params_dev(sfx_batch* b) { return b ? b->params.p : nullptr; }
int32_t* sfx_batch_evals_dev(sfx_batch* b) { return b ? (int32_t*)b->n_evals.p : nullptr; }
int32_t* sfx_batch_flags_dev(sfx_batch* b) { return b ? (int32_t*)b->flags.p : nullptr; }
int32_t* sfx_batch_passes_dev(sfx_batch* b) { return b ? (int32_t*)b->n_passes.p : nullptr; }

int sfx_batch_reset_counters(sfx_batch* b, void* stream) {
    if (!b) return fail(SFX_ERR_ARG, "null argument");
    CUDA_TRY(cudaMemsetAsync(b->n_evals.p, 0, (size_t)b->B * sizeof(int), (cudaStream_t)stream));
    CUDA_TRY(cudaMemsetAsync(b->n_passes.p, 0, (size_t)b->B * sizeof(int), (cudaStream_t)stream));
    CUDA_TRY(cudaMemsetAsync(b->flags.p, 0, (size_t)b->B * sizeof(int), (cudaStream_t)stream));
    if (b->coll_stat.p) CUDA_TRY(cudaMemsetAsync(b->coll_stat.p, 0, b->coll_stat.bytes, (cudaStream_t)stream));
    return SFX_OK;
}

static int check_stage(const sfx_batch* b, const SfxStage* st) {
    if (!b || !st) return fail(SFX_ERR_ARG, "null argument");
    if (st->n_active < 1 || st->n_active > SFX_NP_MAX || st->n_blocks < 1 ||
        st->n_blocks > SFX_MAX_BLOCKS)
        return fail(SFX_ERR_ARG, "stage: bad active parameter set");
    int pos = 0;
    for (int i = 0; i < st->n_blocks; ++i) {
        if (st->block_start[i] != pos || st->block_len[i] < 1 || st->block_off[i] < 0 ||
            st->block_off[i] + st->block_len[i] > b->lay.np)
            return fail(SFX_ERR_ARG, "stage: bad parameter block");
        pos += st->block_len[i];
    }
    if (pos != st->n_active) return fail(SFX_ERR_ARG, "stage: blocks do not add up to n_active");
    if (st->history < 1 || st->history > SFX_HIST) return fail(SFX_ERR_ARG, "stage: history out of range");
    if (st->pprior_kind == SFX_PPRIOR_REGRESSION && !b->has_reg)
        return fail(SFX_ERR_ARG, "stage: regression prior requested but no reg_pose was set");
    if (st->pprior_kind == SFX_PPRIOR_GMM) {
        const int gm = b->m->use_double ? b->m->vd.gmm_M : b->m->vf.gmm_M;
        const int gd = b->m->use_double ? b->m->vd.gmm_D : b->m->vf.gmm_D;
        if (gm < 1) return fail(SFX_ERR_ARG, "stage: mixture prior requested but sfx_model_set_gmm was not called");
        if (gd != b->lay.n_pose || b->use_vposer)
            return fail(SFX_ERR_ARG, "stage: mixture prior dimension does not match the body pose");
    }
    if ((st->use_vposer != 0) != (b->use_vposer != 0))
        return fail(SFX_ERR_ARG, "stage: use_vposer does not match the batch (sfx_batch_create)");
    if (st->use_vposer) {
        const int ready = b->m->use_double ? b->m->vd.vp_ready : b->m->vf.vp_ready;
        if (!ready) return fail(SFX_ERR_ARG, "stage: use_vposer but sfx_model_set_vposer was not called");
        if (st->loss_kind == SFX_LOSS_SMPLIFY && st->pprior_kind != SFX_PPRIOR_LATENT)
            return fail(SFX_ERR_ARG, "stage: use_vposer needs the latent pose prior (fitting.py:389-395)");
    } else if (st->pprior_kind == SFX_PPRIOR_LATENT) {
        return fail(SFX_ERR_ARG, "stage: latent pose prior without use_vposer");
    }
    if (st->opt_kind != SFX_OPT_LBFGSLS && st->opt_kind != SFX_OPT_ADAM)
        return fail(SFX_ERR_UNSUPPORTED, "optimiser kind not supported on the device");
    if (st->loss_kind == SFX_LOSS_SMPLIFY && st->coll_loss_weight > 0) {
        if (!b->coll_vals.p)
            return fail(SFX_ERR_ARG, "stage: coll_loss_weight > 0 needs sfx_model_set_collision and "
                                     "sfx_batch_enable_collisions");
        if (!(st->coll_sigma > 0)) return fail(SFX_ERR_ARG, "stage: coll_sigma (df_cone_height) must be positive");
    }
    return SFX_OK;
}

int sfx_batch_guess_init(sfx_batch* b, const void* joints_dev, const int32_t* edge_idxs, int32_t n_edges,
                         const double* focal, const double* mean_len2d, const uint8_t* need, void* stream) {
    if (!b || !joints_dev || !edge_idxs || !focal || !mean_len2d || !need)
        return fail(SFX_ERR_ARG, "null argument");
    if (n_edges < 1 || n_edges > SFX_GUESS_EDGES_MAX) return fail(SFX_ERR_ARG, "guess_init: 1..16 edges");
    GuessEdges E;
    E.n = n_edges;
    for (int i = 0; i < 2 * n_edges; ++i) {
        if (edge_idxs[i] < 0 || edge_idxs[i] >= b->m->K) return fail(SFX_ERR_ARG, "guess_init: edge index out of range");
        E.idx[i] = edge_idxs[i];
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t B = b->B;
    double* d_focal = (double*)b->guess.p;
    double* d_len = d_focal + B;
    unsigned char* d_need = (unsigned char*)(d_len + B);
    CUDA_TRY(cudaMemcpyAsync(d_focal, focal, B * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_len, mean_len2d, B * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_need, need, B, cudaMemcpyHostToDevice, s));
    const int grid = (int)((B + 127) / 128);
    if (b->m->use_double)
        guess_init_kernel<double><<<grid, 128, 0, s>>>(b->view<double>(nullptr), (const double*)joints_dev, b->m->K, E,
                                                       d_focal, d_len, d_need, (double*)b->cam.p);
    else
        guess_init_kernel<float><<<grid, 128, 0, s>>>(b->view<float>(nullptr), (const float*)joints_dev, b->m->K, E,
                                                      d_focal, d_len, d_need, (float*)b->cam.p);
    CUDA_TRY(cudaGetLastError());
    return SFX_OK;
}

int sfx_eval(sfx_batch* b, const SfxStage* st, void* loss_dev, void* grad_dev, void* joints_dev,
             void* stream) {
    int rc = check_stage(b, st);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int rm = ring_mode_for(b->m);
    if (b->m->use_double) {
#ifdef SFX_DEV_F32_ONLY
        SFX_NO_F64();
#else
        size_t smem = fit_smem<double>(rm);
        CUDA_TRY(cudaFuncSetAttribute(eval_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_kernel<double><<<b->B, SFX_THREADS, smem, s>>>(b->m->vd, b->view<double>(nullptr), *st, rm,
                                                            (double*)loss_dev, (double*)grad_dev,
                                                            (double*)joints_dev);
#endif
    } else {
        size_t smem = fit_smem<float>(rm);
        CUDA_TRY(cudaFuncSetAttribute(eval_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_kernel<float><<<b->B, SFX_THREADS, smem, s>>>(b->m->vf, b->view<float>(nullptr), *st, rm,
                                                          (float*)loss_dev, (float*)grad_dev,
                                                          (float*)joints_dev);
    }
    CUDA_TRY(cudaGetLastError());
    return SFX_OK;
}

int sfx_fit_stage(sfx_batch* b, const SfxStage* st, const int32_t* frame_ids_dev, int32_t n_ids,
                  void* final_loss_dev, void* stream) {
    int rc = check_stage(b, st);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = frame_ids_dev ? n_ids : b->B;
    if (grid < 1) return SFX_OK;
    const int rm = ring_mode_for(b->m);
    if (b->m->use_double) {
#ifdef SFX_DEV_F32_ONLY
        SFX_NO_F64();
#else
        size_t smem = fit_smem<double>(rm);
        CUDA_TRY(cudaFuncSetAttribute(fit_stage_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fit_stage_kernel<double><<<grid, SFX_THREADS, smem, s>>>(
            b->m->vd, b->view<double>(frame_ids_dev), *st, rm, (double*)final_loss_dev);
#endif
    } else {
        size_t smem = fit_smem<float>(rm);
        CUDA_TRY(cudaFuncSetAttribute(fit_stage_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fit_stage_kernel<float><<<grid, SFX_THREADS, smem, s>>>(
            b->m->vf, b->view<float>(frame_ids_dev), *st, rm, (float*)final_loss_dev);
    }
    CUDA_TRY(cudaGetLastError());
    return SFX_OK;
}

int sfx_batch_begin_orientation(sfx_batch* b, int32_t flip, const int32_t* frame_ids_dev,
                                int32_t n_ids, void* stream) {
    if (!b) return fail(SFX_ERR_ARG, "null argument");
    const int grid = frame_ids_dev ? n_ids : b->B;
    if (grid < 1) return SFX_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (b->m->use_double)
        begin_orientation_kernel<double><<<grid, 128, 0, s>>>(
            b->view<double>(frame_ids_dev), flip, (double*)b->go_saved.p, (double*)b->params_alt.p,
            (double*)b->loss_alt.p);
    else
        begin_orientation_kernel<float><<<grid, 128, 0, s>>>(
            b->view<float>(frame_ids_dev), flip, (float*)b->go_saved.p, (float*)b->params_alt.p,
            (float*)b->loss_alt.p);
    CUDA_TRY(cudaGetLastError());
    return SFX_OK;
}

int sfx_batch_select_orientation(sfx_batch* b, const int32_t* frame_ids_dev, int32_t n_ids,
                                 void* stream) {
    if (!b) return fail(SFX_ERR_ARG, "null argument");
    const int grid = frame_ids_dev ? n_ids : b->B;
    if (grid < 1) return SFX_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (b->m->use_double)
        select_orientation_kernel<double><<<grid, 128, 0, s>>>(
            b->view<double>(frame_ids_dev), (const double*)b->params_alt.p, (const double*)b->loss_alt.p);
    else
        select_orientation_kernel<float><<<grid, 128, 0, s>>>(
            b->view<float>(frame_ids_dev), (const float*)b->params_alt.p, (const float*)b->loss_alt.p);
    CUDA_TRY(cudaGetLastError());
    return SFX_OK;
}

void* sfx_batch_final_loss_dev(sfx_batch* b) { return b ? b->final_loss.p : nullptr; }

int sfx_fit_pipeline(sfx_batch* b, const SfxPipeline* pipe, const int32_t* order_dev,
                     const uint8_t* flip_dev, void* stream) {
    if (!b || !pipe) return fail(SFX_ERR_ARG, "null argument");
    if (pipe->n_stages < 1 || pipe->n_stages > SFX_MAX_STAGES)
        return fail(SFX_ERR_ARG, "pipeline: number of stages out of range");
    int rc = check_stage(b, &pipe->cam);
    for (int i = 0; i < pipe->n_stages && !rc; ++i) rc = check_stage(b, &pipe->body[i]);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(b->pipe.p, pipe, sizeof(SfxPipeline), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(b->counter.p, 0, 2 * sizeof(int), s));
    const int rm = ring_mode_for(b->m);
    // wide frames (SfxPipeline::n_wide): float32 with the TMA ring, an explicit launch order, and
    // only while the clusters fit next to the one-block frames (everything stays co-resident)
    int n_wide = pipe->n_wide;
    if (n_wide < 0 || n_wide > b->B) return fail(SFX_ERR_ARG, "pipeline: n_wide out of range");
    if (b->m->use_double || !(rm & 1) || (rm & 4) || !order_dev || b->coll_vals.p) n_wide = 0;
    while (n_wide > 0 && n_wide * SFX_WIDE_CLUSTER + (b->B - n_wide) > b->m->num_sms) --n_wide;
    if (n_wide > 0) {
        size_t smem = fit_smem<float>(rm);
        CUDA_TRY(cudaFuncSetAttribute(fit_pipeline_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(n_wide * SFX_WIDE_CLUSTER));
        cfg.blockDim = dim3(SFX_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = SFX_WIDE_CLUSTER;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, fit_pipeline_kernel<float>, b->m->vf, b->view<float>(order_dev),
                                    (const SfxPipeline*)b->pipe.p, flip_dev, n_wide, (int*)b->counter.p + 1, rm,
                                    (float*)b->cam_loss.p, (float*)b->params_last.p, (int)SFX_WIDE_CLUSTER,
                                    (int)smem, (float*)b->params_alt.p));
        const int n_main = b->B - n_wide;
        if (n_main > 0) {
            // same stream, allowed to overlap the wide grid (which triggers at its very start):
            // the frames of the two grids are independent, nothing here waits on the first grid
            cudaLaunchConfig_t cm;
            memset(&cm, 0, sizeof(cm));
            cm.gridDim = dim3((unsigned)(n_main < b->m->num_sms ? n_main : b->m->num_sms));
            cm.blockDim = dim3(SFX_THREADS);
            cm.dynamicSmemBytes = smem;
            cm.stream = s;
            cudaLaunchAttribute am[1];
            am[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            am[0].val.programmaticStreamSerializationAllowed = 1;
            cm.attrs = am;
            cm.numAttrs = 1;
            CUDA_TRY(cudaLaunchKernelEx(&cm, fit_pipeline_kernel<float>, b->m->vf,
                                        b->view<float>(order_dev + n_wide), (const SfxPipeline*)b->pipe.p,
                                        flip_dev, n_main, (int*)b->counter.p, rm, (float*)b->cam_loss.p,
                                        (float*)b->params_last.p, 1, (int)smem, (float*)b->params_alt.p));
        }
        CUDA_TRY(cudaGetLastError());
        b->last_valid = true;
        return SFX_OK;
    }
    const int grid = b->B < b->m->num_sms ? b->B : b->m->num_sms;
    if (b->m->use_double) {
#ifdef SFX_DEV_F32_ONLY
        SFX_NO_F64();
#else
        size_t smem = fit_smem<double>(rm);
        CUDA_TRY(cudaFuncSetAttribute(fit_pipeline_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fit_pipeline_kernel<double><<<grid, SFX_THREADS, smem, s>>>(
            b->m->vd, b->view<double>(order_dev), (const SfxPipeline*)b->pipe.p, flip_dev, b->B,
            (int*)b->counter.p, rm, (double*)b->cam_loss.p, (double*)b->params_last.p, 1, (int)smem,
            (double*)b->params_alt.p);
#endif
    } else {
        size_t smem = fit_smem<float>(rm);
        CUDA_TRY(cudaFuncSetAttribute(fit_pipeline_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fit_pipeline_kernel<float><<<grid, SFX_THREADS, smem, s>>>(
            b->m->vf, b->view<float>(order_dev), (const SfxPipeline*)b->pipe.p, flip_dev, b->B,
            (int*)b->counter.p, rm, (float*)b->cam_loss.p, (float*)b->params_last.p, 1, (int)smem,
            (float*)b->params_alt.p);
    }
    CUDA_TRY(cudaGetLastError());
    b->last_valid = true;
    return SFX_OK;
}

void* sfx_batch_cam_loss_dev(sfx_batch* b) { return b ? b->cam_loss.p : nullptr; }
int sfx_pack_keypoints(const float* body_dev, const float* lhand_dev, const float* rhand_dev,
                       const float* face_dev, int32_t B, int32_t n_body, int32_t n_face,
                       int32_t use_face_contour, float* out_dev, void* stream) {
    if (!body_dev || !lhand_dev || !rhand_dev || !face_dev || !out_dev || B < 1 || n_body < 1 || n_face < 68)
        return fail(SFX_ERR_ARG, "bad argument");
    const int K = n_body + 42 + 51 + (use_face_contour ? 17 : 0);
    const int n = B * K;
    pack_keypoints_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        body_dev, lhand_dev, rhand_dev, face_dev, B, n_body, n_face, use_face_contour, K, out_dev);
    CUDA_TRY(cudaGetLastError());
    return SFX_OK;
}

int sfx_keypoint_masks(const float* keypoints_dev, const float* base_joint_weights_dev,
                       const int32_t* init_joints_idxs_dev, int32_t n_init, int32_t n_body,
                       float confidence_threshold, int32_t B, int32_t K, int32_t use_double, void* gt_dev,
                       void* conf_dev, void* joint_weights_dev, uint8_t* lowconf_dev, uint8_t* init_mask_dev,
                       void* stream) {
    if (!keypoints_dev || !base_joint_weights_dev || (n_init > 0 && !init_joints_idxs_dev) || !gt_dev ||
        !conf_dev || !joint_weights_dev || !lowconf_dev || !init_mask_dev || B < 1 || K < 1 || n_init < 0)
        return fail(SFX_ERR_ARG, "bad argument");
    const int n = B * K;
    cudaStream_t s = (cudaStream_t)stream;
    if (use_double)
        keypoint_masks_kernel<double><<<(n + 255) / 256, 256, 0, s>>>(
            keypoints_dev, base_joint_weights_dev, init_joints_idxs_dev, n_init, n_body, confidence_threshold,
            B, K, (double*)gt_dev, (double*)conf_dev, (double*)joint_weights_dev, lowconf_dev, init_mask_dev);
    else
        keypoint_masks_kernel<float><<<(n + 255) / 256, 256, 0, s>>>(
            keypoints_dev, base_joint_weights_dev, init_joints_idxs_dev, n_init, n_body, confidence_threshold,
            B, K, (float*)gt_dev, (float*)conf_dev, (float*)joint_weights_dev, lowconf_dev, init_mask_dev);
    CUDA_TRY(cudaGetLastError());
    return SFX_OK;
}

int sfx_blend_keypoints(const float* openpose_dev, const float* mmpose_dev, const float* stats_dev,
                        const int32_t* pair_mmpose_dev, const int32_t* pair_openpose_dev, int32_t n_pairs,
                        int32_t B, float* out_dev, void* stream) {
    if (!openpose_dev || !mmpose_dev || !stats_dev || !pair_mmpose_dev || !pair_openpose_dev || !out_dev ||
        B < 1 || n_pairs < 1 || n_pairs > 67)
        return fail(SFX_ERR_ARG, "bad argument");
    const int n = B * (n_pairs + 68);
    blend_keypoints_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        openpose_dev, mmpose_dev, stats_dev, pair_mmpose_dev, pair_openpose_dev, n_pairs, B, out_dev);
    CUDA_TRY(cudaGetLastError());
    return SFX_OK;
}

int sfx_diag_l2_read_gbs(int64_t buffer_bytes, int32_t iters, double* gbs_out) {
    if (!gbs_out || buffer_bytes < (1 << 20) || iters < 1) return fail(SFX_ERR_ARG, "bad argument");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DevBuf buf, sink;
    CUDA_TRY(buf.alloc((size_t)buffer_bytes));
    CUDA_TRY(sink.alloc(64));
    CUDA_TRY(cudaMemset(buf.p, 1, (size_t)buffer_bytes));
    const size_t n16 = (size_t)buffer_bytes / 16;
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    l2_read_kernel<<<sms, 512>>>((const uint4*)buf.p, n16, 1, (unsigned int*)sink.p);   // warm L2
    CUDA_TRY(cudaEventRecord(a));
    l2_read_kernel<<<sms, 512>>>((const uint4*)buf.p, n16, iters, (unsigned int*)sink.p);
    CUDA_TRY(cudaEventRecord(b));
    CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    CUDA_TRY(cudaGetLastError());
    *gbs_out = (double)sms * (double)n16 * 16.0 * iters / (ms * 1e-3) / 1e9;
    return SFX_OK;
}

long long* sfx_batch_prof_dev(sfx_batch* b) { return b ? (long long*)b->prof.p : nullptr; }

static int forward_mesh_impl(sfx_batch* b, void* vertices_dev, void* joints_dev, void* stream);

int sfx_forward_mesh(sfx_batch* b, void* vertices_dev, void* joints_dev, void* stream) {
    return forward_mesh_impl(b, vertices_dev, joints_dev, stream);
}

int sfx_forward_mesh_last(sfx_batch* b, void* vertices_dev, void* joints_dev, void* stream) {
    if (!b) return fail(SFX_ERR_ARG, "null argument");
    if (!b->last_valid) return fail(SFX_ERR_ARG, "no pipeline launch has run on this batch");
    // evaluate at the last orientation's parameters, then put the selected ones back
    std::swap(b->params.p, b->params_last.p);
    int rc = forward_mesh_impl(b, vertices_dev, joints_dev, stream);
    std::swap(b->params.p, b->params_last.p);
    return rc;
}

}  // extern "C"

static int forward_mesh_impl(sfx_batch* b, void* vertices_dev, void* joints_dev, void* stream) {
    if (!b || !vertices_dev) return fail(SFX_ERR_ARG, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const sfx_model* m = b->m;
    if (m->use_double) {
#ifdef SFX_DEV_F32_ONLY
        SFX_NO_F64();
#else
        size_t smem = fit_smem<double>(0);
        CUDA_TRY(cudaFuncSetAttribute(mesh_coef_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mesh_coef_kernel<double><<<b->B, SFX_THREADS, smem, s>>>(m->vd, b->view<double>(nullptr), b->use_vposer,
                                                         (double*)b->Acoef.p, (double*)b->Ccoef.p);
        CUDA_TRY(cudaGetLastError());
        std::string e = mesh_forward_simt<double>(m->vd, b->B, (const double*)b->Acoef.p,
                                                  (const double*)b->Ccoef.p, (double*)b->vposed.p,
                                                  (double*)vertices_dev, s);
        if (!e.empty()) return fail(SFX_ERR_CUDA, e);
#endif
    } else {
        size_t smem = fit_smem<float>(0);
        CUDA_TRY(cudaFuncSetAttribute(mesh_coef_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mesh_coef_kernel<float><<<b->B, SFX_THREADS, smem, s>>>(m->vf, b->view<float>(nullptr), b->use_vposer,
                                                       (float*)b->Acoef.p, (float*)b->Ccoef.p);
        CUDA_TRY(cudaGetLastError());
        // blend contraction: tcgen05 / TMA kernel (SFX_MESH_SIMT=1 selects the fp32 SIMT kernel,
        // the accuracy reference of the tf32 path); skinning: SIMT epilogue kernel
        const char* simt = getenv("SFX_MESH_SIMT");
        const char* unfused = getenv("SFX_MESH_UNFUSED");
        std::string e;
        if (simt && simt[0] == '1') {
            e = mesh_forward_simt<float>(m->vf, b->B, (const float*)b->Acoef.p,
                                         (const float*)b->Ccoef.p, (float*)b->vposed.p,
                                         (float*)vertices_dev, s);
        } else if (!(unfused && unfused[0] == '1')) {
            // one tensor-core kernel: blend (float16 operands; SFX_MESH_K1_TF32=1: tf32) + skinning
            // (3 x f16 split) + transform epilogue
            const char* k1tf32 = getenv("SFX_MESH_K1_TF32");
            e = mesh_fused_tc(m->fused, b->B, (const float*)b->Ccoef.p, (const float*)b->Acoef.p,
                              (__half*)b->ahi.p, (__half*)b->alo.p, (__half*)b->c16.p,
                              !(k1tf32 && k1tf32[0] == '1'), (const float*)m->vt.p,
                              (float*)vertices_dev, s);
        } else {
            e = mesh_blend_tc(m->mesh, b->B, (const float*)b->Ccoef.p, (const float*)m->vt.p,
                              (float*)b->vposed.p, s);
            if (e.empty())
                e = mesh_skin<float>(m->vf, b->B, (const float*)b->Acoef.p, (const float*)b->vposed.p,
                                     (float*)vertices_dev, s);
        }
        if (!e.empty()) return fail(SFX_ERR_CUDA, e);
    }
    if (joints_dev) {
        // mapped joints through the sparse-support evaluation (forward part only matters)
        SfxStage st;
        memset(&st, 0, sizeof(st));
        st.loss_kind = SFX_LOSS_CAMERA_INIT;
        st.rho = 100;
        st.n_active = 3; st.n_blocks = 1; st.block_start[0] = 0; st.block_len[0] = 3;
        st.block_off[0] = b->lay.off_go; st.history = SFX_HIST; st.opt_kind = SFX_OPT_LBFGSLS;
        st.use_vposer = b->use_vposer;
        int rc = sfx_eval(b, &st, nullptr, nullptr, joints_dev, stream);
        if (rc) return rc;
    }
    return SFX_OK;
}
