// Evaluation metrics of the reference: per-point error between fitted and
// ground-truth point sets after an alignment --
//   Procrustes (similarity transform)  utils.py:540-594 ProcrustesAlignment.__call__
//   pelvis (mean of two hip points)    utils.py:650-671 PelvisAlignment
//   scale only                         utils.py:729-771 ScaleAlignment
//   none
// followed by mpjpe / vertex_to_vertex_error (utils.py:596-615): sqrt(sum((a - b)^2, -1)).
// The reference loops over frames on the host with numpy (eval.py:14-44 compute_v2v); here one
// block owns one frame: a first pass over the two point sets accumulates means, variances and
// the 3x3 cross-covariance, one thread solves the 3x3 problem (Jacobi SVD, double), a second
// pass writes the per-point errors.  HBM-bound: 2 x N x 12 bytes read twice, N x 4 written.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include <string>

namespace sfx {

#define SFX_ALIGN_NONE 0
#define SFX_ALIGN_PROCRUSTES 1
#define SFX_ALIGN_PELVIS 2
#define SFX_ALIGN_SCALE 3

// A = U diag(s) V^T for a 3x3 matrix by one-sided Jacobi (Hestenes) in double
__device__ inline void svd3(const double* A, double* U, double* s, double* V) {
    double B[9], Vm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; ++i) B[i] = A[i];
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double a = 0, b = 0, c = 0;
                for (int i = 0; i < 3; ++i) {
                    a += B[3 * i + p] * B[3 * i + p];
                    b += B[3 * i + q] * B[3 * i + q];
                    c += B[3 * i + p] * B[3 * i + q];
                }
                off += fabs(c);
                if (fabs(c) <= 1e-300 || fabs(c) <= 1e-17 * sqrt(a * b)) continue;
                const double zeta = (b - a) / (2.0 * c);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int i = 0; i < 3; ++i) {
                    const double bp = B[3 * i + p], bq = B[3 * i + q];
                    B[3 * i + p] = cs * bp - sn * bq;
                    B[3 * i + q] = sn * bp + cs * bq;
                    const double vp = Vm[3 * i + p], vq = Vm[3 * i + q];
                    Vm[3 * i + p] = cs * vp - sn * vq;
                    Vm[3 * i + q] = sn * vp + cs * vq;
                }
            }
        if (off == 0) break;
    }
    // columns of B are U * s; order by decreasing singular value
    double n[3];
    int ord[3] = {0, 1, 2};
    for (int j = 0; j < 3; ++j) n[j] = sqrt(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (n[ord[b]] > n[ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    for (int j = 0; j < 3; ++j) {
        const int c = ord[j];
        s[j] = n[c];
        for (int i = 0; i < 3; ++i) {
            V[3 * i + j] = Vm[3 * i + c];
            U[3 * i + j] = n[c] > 0 ? B[3 * i + c] / n[c] : 0.0;
        }
    }
    // a vanishing singular value leaves its left vector undefined: complete the basis
    if (!(s[2] > 1e-300 * (s[0] > 0 ? s[0] : 1.0))) {
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
}

__device__ inline double det3(const double* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
           M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// est, gt: [B][N][3]; idx: optional point subset [n_idx] (eval.py:24-26 `vids`); err: [B][n]
// hips: the two point indices PelvisAlignment averages (after the subset is applied)
template <typename T>
__global__ void __launch_bounds__(256)
aligned_error_kernel(const T* __restrict__ est, const T* __restrict__ gt, const int* __restrict__ idx,
                     int N, int n, int mode, int hip0, int hip1, T* __restrict__ err,
                     T* __restrict__ transform) {
    __shared__ double red[17][8];
    __shared__ double sol[16];          // scale, R[9], t[3] applied to est (and gt offset for pelvis)
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T* E = est + (size_t)b * N * 3;
    const T* G = gt + (size_t)b * N * 3;
    auto at = [&](int i) { return idx ? idx[i] : i; };
    if (mode == SFX_ALIGN_PROCRUSTES || mode == SFX_ALIGN_SCALE) {
        // means first (the reference subtracts them before anything else), then the centred sums
        double acc[17];
        for (int k = 0; k < 17; ++k) acc[k] = 0;
        for (int i = tid; i < n; i += blockDim.x) {
            const int p = at(i);
            for (int d = 0; d < 3; ++d) {
                acc[d] += (double)E[3 * p + d];
                acc[3 + d] += (double)G[3 * p + d];
            }
        }
        for (int k = 0; k < 6; ++k) {
            double v = acc[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[k][warp] = v;
        }
        __syncthreads();
        double mu[6];
        for (int k = 0; k < 6; ++k) {
            double v = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[k][w];
            mu[k] = v / n;
        }
        __syncthreads();
        for (int k = 0; k < 17; ++k) acc[k] = 0;
        for (int i = tid; i < n; i += blockDim.x) {
            const int p = at(i);
            double x1[3], x2[3];
            for (int d = 0; d < 3; ++d) {
                x1[d] = (double)E[3 * p + d] - mu[d];
                x2[d] = (double)G[3 * p + d] - mu[3 + d];
            }
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) acc[3 * r + c] += x1[r] * x2[c];       // K = X1 X2^T
            acc[9] += x1[0] * x1[0] + x1[1] * x1[1] + x1[2] * x1[2];               // var1
            acc[10] += x2[0] * x2[0] + x2[1] * x2[1] + x2[2] * x2[2];              // var2
        }
        for (int k = 0; k < 11; ++k) {
            double v = acc[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[k][warp] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double Kc[11];
            for (int k = 0; k < 11; ++k) {
                double v = 0;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[k][w];
                Kc[k] = v;
            }
            double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, scale;
            if (mode == SFX_ALIGN_PROCRUSTES) {
                double U[9], s[3], V[9];
                svd3(Kc, U, s, V);
                // R = V Z U^T with Z fixing det(R) = +1 (utils.py:574-579)
                double UVt[9];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                        UVt[3 * r + c] = U[3 * r] * V[3 * c] + U[3 * r + 1] * V[3 * c + 1] + U[3 * r + 2] * V[3 * c + 2];
                const double dt = det3(UVt);
                const double z = dt > 0 ? 1.0 : (dt < 0 ? -1.0 : 0.0);
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                        R[3 * r + c] = V[3 * r] * U[3 * c] + V[3 * r + 1] * U[3 * c + 1] + z * V[3 * r + 2] * U[3 * c + 2];
                double tr = 0;                      // trace(R K)
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) tr += R[3 * r + c] * Kc[3 * c + r];
                scale = tr / Kc[9];
            } else {
                scale = sqrt(Kc[10] / Kc[9]);
            }
            sol[0] = scale;
            for (int k = 0; k < 9; ++k) sol[1 + k] = R[k];
            for (int r = 0; r < 3; ++r)             // t = mu2 - scale R mu1
                sol[10 + r] = mu[3 + r] - scale * (R[3 * r] * mu[0] + R[3 * r + 1] * mu[1] + R[3 * r + 2] * mu[2]);
            sol[13] = sol[14] = sol[15] = 0;
        }
    } else if (tid == 0) {
        sol[0] = 1;
        for (int k = 0; k < 9; ++k) sol[1 + k] = (k % 4 == 0) ? 1.0 : 0.0;
        for (int r = 0; r < 3; ++r) sol[10 + r] = sol[13 + r] = 0;
        if (mode == SFX_ALIGN_PELVIS) {
            // both sets move to their own pelvis (utils.py:659-671)
            const int p0 = at(hip0), p1 = at(hip1);
            for (int r = 0; r < 3; ++r) {
                sol[10 + r] = -0.5 * ((double)E[3 * p0 + r] + (double)E[3 * p1 + r]);
                sol[13 + r] = -0.5 * ((double)G[3 * p0 + r] + (double)G[3 * p1 + r]);
            }
        }
    }
    __syncthreads();
    const double sc = sol[0];
    for (int i = tid; i < n; i += blockDim.x) {
        const int p = at(i);
        const double x = E[3 * p], y = E[3 * p + 1], z = E[3 * p + 2];
        double d2 = 0;
        for (int r = 0; r < 3; ++r) {
            const double a = sc * (sol[1 + 3 * r] * x + sol[2 + 3 * r] * y + sol[3 + 3 * r] * z) + sol[10 + r];
            const double g = (double)G[3 * p + r] + sol[13 + r];
            d2 += (a - g) * (a - g);
        }
        err[(size_t)b * n + i] = (T)sqrt(d2);
    }
    if (transform && tid < 13) transform[(size_t)b * 13 + tid] = (T)sol[tid];
}

template <typename T>
static std::string aligned_errors(const T* est, const T* gt, const int* idx, int B, int N, int n,
                                  int mode, int hip0, int hip1, T* err, T* transform, cudaStream_t s) {
    if (B < 1) return "";
    aligned_error_kernel<T><<<B, 256, 0, s>>>(est, gt, idx, N, n, mode, hip0, hip1, err, transform);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? "" : std::string("aligned_error_kernel: ") + cudaGetErrorString(e);
}

}  // namespace sfx
