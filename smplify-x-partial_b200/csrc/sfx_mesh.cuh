// Full-mesh SMPL-X forward for B frames (body_model(return_verts=True), reference
// fit_single_frame.py:611; third-party smplx lbs.lbs):
//
//   v_posed[b][r] = vt[r] + PK[r][:] . c[b][:]          r in [0, 3V)   (blend, K = 512)
//   T[b][v]       = sum_j Wd[v][j] * A[b][j]            (3x4 per vertex) (skinning)
//   verts[b][v]   = T[b][v] . [v_posed[b][v], 1]
//
// Inputs come from mesh_coef_kernel (pose prologue per frame): A [B][55*12], c [B][512].
// Two implementations of the blend contraction: SIMT (any dtype; validation and fp64) and
// the tcgen05 / TMA tensor-core kernel in sfx_mesh_tc.cuh (fp32 product path).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "sfx_core.cuh"

namespace sfx {

// frames padded to the UMMA M tile
inline int mesh_padded_frames(int B) { return (B + 127) / 128 * 128; }

// ---- SIMT blend: one warp owns one PK row (kept in registers) and sweeps the frames --------
template <typename T, int FRAMES_PER_PASS>
__global__ void __launch_bounds__(256)
mesh_blend_simt_kernel(const T* __restrict__ PK, const T* __restrict__ vt,
                       const T* __restrict__ C, int B, long nrows, T* __restrict__ vposed) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int NV = SFX_KPAD / (32 * VEC);
    constexpr int NE = VEC * NV;
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    T vals[NE];
    {
        const uint4* p = reinterpret_cast<const uint4*>(PK + row * SFX_KPAD);
#pragma unroll
        for (int i = 0; i < NV; ++i)
            *reinterpret_cast<uint4*>(vals + i * VEC) = p[i * 32 + lane];
    }
    const T base = vt[row];
    for (int b0 = 0; b0 < B; b0 += FRAMES_PER_PASS) {
        T acc[FRAMES_PER_PASS];
#pragma unroll
        for (int f = 0; f < FRAMES_PER_PASS; ++f) {
            acc[f] = 0;
            const int b = b0 + f;
            if (b < B) {
                const uint4* cp = reinterpret_cast<const uint4*>(C + (long)b * SFX_KPAD);
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    uint4 raw = cp[i * 32 + lane];
                    const T* cv = reinterpret_cast<const T*>(&raw);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[f] += vals[i * VEC + e] * cv[e];
                }
            }
        }
#pragma unroll
        for (int f = 0; f < FRAMES_PER_PASS; ++f) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[f] += __shfl_xor_sync(0xffffffffu, acc[f], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int f = 0; f < FRAMES_PER_PASS; ++f)
                if (b0 + f < B) vposed[(long)(b0 + f) * nrows + row] = base + acc[f];
        }
    }
}

// ---- skinning: block = (vertex tile, frame); A[b] staged in shared memory ------------------
template <typename T>
__global__ void __launch_bounds__(256)
mesh_skin_kernel(const T* __restrict__ Wd, const T* __restrict__ A, const T* __restrict__ vposed,
                 int V, T* __restrict__ verts) {
    __shared__ T sA[SFX_NJ * 12];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < SFX_NJ * 12; i += blockDim.x) sA[i] = A[(long)b * SFX_NJ * 12 + i];
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const T* w = Wd + (long)v * SFX_WROW;
    T Tm[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) Tm[k] = 0;
    for (int j = 0; j < SFX_NJ; ++j) {
        const T wj = w[j];
        if (wj != (T)0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) Tm[k] += wj * sA[12 * j + k];
        }
    }
    const T* p = vposed + ((long)b * V + v) * 3;
    const T x = p[0], y = p[1], z = p[2];
    T* o = verts + ((long)b * V + v) * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) o[r] = Tm[4 * r] * x + Tm[4 * r + 1] * y + Tm[4 * r + 2] * z + Tm[4 * r + 3];
}

template <typename T>
static std::string mesh_skin(const ModelView<T>& M, int B, const T* A, const T* vposed, T* verts,
                             cudaStream_t s) {
    dim3 grid((M.V + 255) / 256, B);
    mesh_skin_kernel<T><<<grid, 256, 0, s>>>(M.Wd, A, vposed, M.V, verts);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? "" : std::string("mesh_skin_kernel: ") + cudaGetErrorString(e);
}

template <typename T>
static std::string mesh_forward_simt(const ModelView<T>& M, int B, const T* A, const T* C, T* vposed,
                                     T* verts, cudaStream_t s) {
    const long nrows = 3L * M.V;
    const int warps = 8;
    mesh_blend_simt_kernel<T, 4><<<(unsigned)((nrows + warps - 1) / warps), warps * 32, 0, s>>>(
        M.PK, M.vt, C, B, nrows, vposed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return std::string("mesh_blend_simt_kernel: ") + cudaGetErrorString(e);
    return mesh_skin(M, B, A, vposed, verts, s);
}

}  // namespace sfx
