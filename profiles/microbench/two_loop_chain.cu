// microbenchmark: one warp, two-loop first loop over k pairs staged in shared memory
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
template <int MODE>
__device__ __forceinline__ float tree(float p, uint32_t slot, int lane) {
    if (MODE == 1) { for (int o = 16; o > 0; o >>= 1) p = p + __shfl_xor_sync(0xffffffffu, p, o); return p; }
    if (MODE == 2) return p;     // no reduction at all
    sts_f32(slot + 4u * lane, p);
    __syncwarp();
    float a[32];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[4*i]), "=f"(a[4*i+1]), "=f"(a[4*i+2]), "=f"(a[4*i+3]) : "r"(slot + 16u * i) : "memory");
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int l = 0; l < o; ++l) a[l] = a[l] + a[l + o];
    return a[0];
}
template <int MODE>
__global__ void k(float* out, long long* cyc, int K) {
    extern __shared__ float sm[];
    float* hist = sm;               // [K][2][120]
    float* red = sm + 100 * 240;    // [64]
    float* ro = red + 64;           // [100]
    float* al = ro + 100;
    for (int i = threadIdx.x; i < 100 * 240 + 264; i += blockDim.x) sm[i] = 0.001f * ((i * 37) % 101) - 0.05f;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const uint32_t RB = 480;
    const uint32_t a_base = smem_u32(hist) + 4u * lane, a_red = smem_u32(red), a_ro = smem_u32(ro), a_al = smem_u32(al);
    float q[4], sc[4], yc[4], sn[4], yn[4];
    for (int r = 0; r < 4; ++r) q[r] = 0.01f * (lane + 32 * r);
    const bool tail = 96 + lane < 119;
    long long t0 = clock64();
    for (int rep = 0; rep < 10; ++rep) {
#define LOAD(ds, dy, idx) do { uint32_t as_ = a_base + (uint32_t)(idx) * 2u * RB, ay_ = as_ + RB; \
    _Pragma("unroll") for (int r = 0; r < 3; ++r) { ds[r] = lds_f32(as_ + 128u * r); dy[r] = lds_f32(ay_ + 128u * r); } \
    ds[3] = tail ? lds_f32(as_ + 384u) : 0.f; dy[3] = tail ? lds_f32(ay_ + 384u) : 0.f; } while (0)
        LOAD(sn, yn, 0);
        for (int idx = 0; idx < K; ++idx) {
            const int i = K - 1 - idx;
            const float ro_i = lds_f32(a_ro + 4u * i);
#pragma unroll
            for (int r = 0; r < 4; ++r) { sc[r] = sn[r]; yc[r] = yn[r]; }
            if (idx + 1 < K) LOAD(sn, yn, idx + 1);
            float p = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) p = p + sc[r] * q[r];
            p = tree<MODE>(p, a_red + 128u * (i & 1), lane);
            const float a = p * ro_i;
            if (lane == 0) sts_f32(a_al + 4u * i, a);
#pragma unroll
            for (int r = 0; r < 4; ++r) q[r] += -a * yc[r];
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = t1 - t0;
    out[lane] = q[0] + q[1] + q[2] + q[3];
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    size_t smem = (100 * 240 + 264) * 4;
    long long h;
    for (int mode = 0; mode < 3; ++mode) {
        auto fn = mode == 0 ? k<0> : (mode == 1 ? k<1> : k<2>);
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int it = 0; it < 2; ++it) { fn<<<1, 512, smem>>>(out, cyc, 100); cudaDeviceSynchronize(); }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("mode %d (0 smem tree, 1 shuffles, 2 no reduction): %.1f cycles per step (%s)\n", mode, h / 1000.0, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
