// Microbenchmark + self-check of the scalar recurrences of the Gram two-loop: the generic chain
// (csrc/sfx_core.cuh gram_chain) against the shared-address float32 one (csrc/sfx_stream.cuh
// gram_chain_f32) on the same zero-padded [128][129] block: prints cycles per step of both
// (2k steps) and whether the coefficients agree bit for bit.  Diagnostics only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --extended-lambda \
//        -I smplify-x-partial_b200/csrc -o profiles/microbench/gram_chain.bin profiles/microbench/gram_chain.cu
#include <cstdio>
#include <cstring>
__device__ long long g_probe;
#define SFX_CHAIN_PROBE g_probe
#include "sfx_core.cuh"
#include "sfx_stream.cuh"
using namespace sfx;
__global__ void chain_kernel(int k, float hd, int which, long long* out, float* coef) {
    extern __shared__ __align__(1024) unsigned char smem[];
    Scratch<float>& S = *reinterpret_cast<Scratch<float>*>(smem);
    float* G = reinterpret_cast<float*>(smem + ((sizeof(Scratch<float>) + 1023) / 1024 * 1024));
    for (int idx = threadIdx.x; idx < SFX_GRAM_ROWS * SFX_GRAM_LDF; idx += blockDim.x) {
        const int i = idx / SFX_GRAM_LDF, j = idx % SFX_GRAM_LDF;
        unsigned h = (unsigned)(i * 131 + j * 71 + 7) * 2654435761u;
        G[idx] = (i < k && j < k) ? ((int)((h >> 8) % 2001) - 1000) * 2e-6f + (i == j ? 1.f : 0.f) : 0.f;
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) { S.sg[i] = 0.01f * (i % 13) - 0.05f; S.yg[i] = 0.02f * (i % 7) - 0.06f; S.ro[i] = 0.9f + 0.001f * i; }
    __syncthreads();
    long long t0 = clock64();
    if (threadIdx.x < 32) {
        if (which == 0) gram_chain(S, k, hd, GramStaged<float>{G, SFX_GRAM_LDF}, (int)threadIdx.x);
        else gram_chain_f32(S, k, hd, G, (int)threadIdx.x);
    }
    __syncthreads();
    if (threadIdx.x == 0) { out[0] = clock64() - t0; out[1] = g_probe - t0; }
    for (int i = threadIdx.x; i < k; i += blockDim.x) { coef[i] = S.al[i]; coef[128 + i] = S.cf[i]; }
}
int main() {
    long long* out; float* coef;
    cudaMalloc(&out, 16); cudaMalloc(&coef, 256 * 4);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int k : {1, 31, 33, 64, 100}) {
        long long h[2] = {0, 0}, h1[2] = {0, 0};
        float c[2][256];
        for (int which = 0; which < 2; ++which) {
            for (int it = 0; it < 3; ++it) {
                cudaMemset(coef, 0, 256 * 4);
                chain_kernel<<<1, 512, smem>>>(k, 0.7f, which, out, coef);
                cudaMemcpy(&h[which], out, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&h1[which], out + 1, 8, cudaMemcpyDeviceToHost);
            }
            cudaMemcpy(c[which], coef, 256 * 4, cudaMemcpyDeviceToHost);
        }
        printf("  f32 first loop %.1f cycles/step, second %.1f\n", (double)h1[1] / k, (double)(h[1] - h1[1]) / k);
        printf("k=%3d  generic %.1f cycles/step   f32 shared-address %.1f cycles/step   coefficients %s  (al[0]=%g cf[k-1]=%g) %s\n",
               k, (double)h[0] / (2 * k), (double)h[1] / (2 * k),
               memcmp(c[0], c[1], sizeof(c[0])) == 0 ? "bit-identical" : "DIFFER", c[1][0], c[1][128 + k - 1],
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
