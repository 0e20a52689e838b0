// Microbenchmark: what ONE SM draws from L2 with TMA bulk copies of 2 KiB rows (the blend-row stream of
// the fit kernel, csrc/sfx_stream.cuh) as a function of the rows in flight, alone and with every SM
// doing the same.  One 32-thread warp per ring slot group is not needed: a single thread issues, as
// the copies are asynchronous.  Diagnostics only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/microbench/l2_per_sm.bin profiles/microbench/l2_per_sm.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ROW 2048
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// lane 0 of each of NW = min(16, depth) warps keeps depth / NW row copies in flight (the fit kernel's
// streaming warps do the same); rows are taken with a stride so that successive rows fall into
// different L2 slices
__global__ void k(const unsigned char* buf, size_t nrows, int depth, int rows_total, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)depth * ROW);
    const int nw = depth < 16 ? depth : 16, warp = threadIdx.x >> 5, per = depth / nw;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long t0 = clock64();
    if ((threadIdx.x & 31) == 0 && warp < nw) {
        size_t r = ((size_t)blockIdx.x * 7919u + (size_t)warp * 997u) % nrows;
        const int mine = rows_total / nw;
        for (int i = 0; i < mine + per; ++i) {
            const int s = warp * per + i % per;
            if (i >= per) mbar_wait(bars + s, ((i / per) - 1) & 1);
            if (i < mine) {
                mbar_expect(bars + s, ROW);
                bulk(sm + (size_t)s * ROW, buf + r * ROW, ROW, bars + s);
                r += 37; if (r >= nrows) r -= nrows;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}
int main() {
    const size_t bytes = 8u << 20, nrows = bytes / ROW;      // 8 MiB: L2 resident
    unsigned char* buf; long long* cyc;
    cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes); cudaMalloc(&cyc, 256 * sizeof(long long));
    int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
    const int rows_total = 19200;      // a multiple of every warp count
    printf("SM clock %d kHz (attribute), %d SMs; GB/s per SM from elapsed SM cycles at 1.965 GHz\n", khz, sms);
    for (int grid : {1, 2, 8, 148}) {
        if (grid > sms) grid = sms;
        for (int depth : {4, 8, 16, 32, 64, 96}) {
            size_t smem = (size_t)depth * ROW + depth * 8 + 64;
            k<<<grid, 512, smem>>>(buf, nrows, depth, 2000, cyc);           // warm
            k<<<grid, 512, smem>>>(buf, nrows, depth, rows_total, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[256]; cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            double worst = 0, sum = 0;
            for (int i = 0; i < grid; ++i) { double g = (double)rows_total * ROW / (h[i] / 1.965e9) / 1e9; sum += g; if (i == 0 || g < worst) worst = g; }
            printf("blocks %3d  rows in flight %3d (%3d KB): %.1f GB/s per SM (slowest %.1f), total %.0f GB/s\n", grid, depth, depth * 2, sum / grid, worst, sum);
        }
    }
    return 0;
}
