// Device implementation of the two passes that dominate a closure evaluation: streaming the
// 225*3 support rows of the blend matrix PK (2 KB each in fp32) through the SM.
//
//   blend_forward :  vp[r]  = vt[row_r] + PK[row_r] . c            (warp per row, 512-long dot)
//   blend_adjoint :  dc[k]  = sum_r PK[row_r][k] * dvp[r]          (per-warp partials, then a
//                                                                   fixed-order cross-warp sum)
//
// Ring mode (fp32): every warp owns SFX_NBUF 2 KB shared-memory buffers, each with its own
// mbarrier; lane 0 issues `cp.async.bulk` (TMA bulk copy, UBLKCP in SASS) for the warp's next
// rows while the warp consumes the current one, so ~96 KB are in flight per SM with no
// register staging and no block-wide synchronisation inside a pass.
// Direct mode (fp64 validation build, or SFX_STREAM_DIRECT): plain coalesced 16-byte loads.
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#include <stdint.h>
#include "sfx_core.cuh"

namespace sfx {

#define SFX_NWARP 16
#define SFX_NBUF 4

struct StreamWS {
    unsigned char* ring;        // ring mode: [NWARP][NBUF][512*sizeof(T)]; else partials only
    uint64_t* bars;             // [NWARP][NBUF]
    unsigned int* fills;        // [NWARP] number of buffers filled so far (phase tracking)
    int ring_mode;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <typename T>
__device__ __forceinline__ void stream_init(StreamWS& ws) {
    if (ws.ring_mode) {
        if (threadIdx.x < SFX_NWARP * SFX_NBUF) mbar_init(ws.bars + threadIdx.x, 1);
        if (threadIdx.x < SFX_NWARP) ws.fills[threadIdx.x] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

template <typename T>
struct RowVec {
    static constexpr int VEC = 16 / sizeof(T);            // elements per 16-byte access
    static constexpr int NV = SFX_KPAD / (32 * VEC);      // 16-byte accesses per lane per row
    static constexpr int NE = VEC * NV;                   // elements per lane per row (16)
};

// element index of (access i, lane l, component e)
#define SFX_ELEM(i, l, e) ((((i) * 32 + (l)) * RowVec<T>::VEC) + (e))

template <typename T>
__device__ __forceinline__ void load_row_regs(const T* row, int lane, T* dst) {
    const uint4* p = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int i = 0; i < RowVec<T>::NV; ++i) {
        uint4 v = p[i * 32 + lane];
        *reinterpret_cast<uint4*>(dst + i * RowVec<T>::VEC) = v;
    }
}

// Generic driver: FN(row_index_r, const T* row_values_in_regs[16]) is called by every lane of the
// warp that owns support row r, rows taken in the fixed order r = warp, warp + NWARP, ...
template <typename T, typename RP, typename FN>
__device__ __forceinline__ void stream_rows(int nrows, RP rowptr, StreamWS& ws, FN fn) {
    // warp 0 runs the kinematic chain while warps 1 .. NWARP-1 stream the rows
    const int warp = (threadIdx.x >> 5) - 1, lane = threadIdx.x & 31;
    constexpr int NSTREAM = SFX_NWARP - 1;
    const int count = warp >= 0 ? (nrows - warp + NSTREAM - 1) / NSTREAM : 0;
    constexpr uint32_t ROWB = SFX_KPAD * sizeof(T);
    T vals[RowVec<T>::NE];
    if (count == 0) return;
    if (ws.ring_mode) {
        unsigned char* mybuf = ws.ring + (size_t)warp * SFX_NBUF * ROWB;
        uint64_t* mybar = ws.bars + warp * SFX_NBUF;
        const unsigned int n0 = ws.fills[warp];
        auto issue = [&](int i) {
            int r = warp + i * NSTREAM;
            unsigned int n = n0 + i;
            int b = n % SFX_NBUF;
            mbar_expect_tx(mybar + b, ROWB);
            bulk_g2s(mybuf + (size_t)b * ROWB, rowptr(r), ROWB, mybar + b);
        };
        if (lane == 0) {
            fence_proxy_async();
            for (int i = 0; i < SFX_NBUF && i < count; ++i) issue(i);
        }
        for (int i = 0; i < count; ++i) {
            unsigned int n = n0 + i;
            int b = n % SFX_NBUF;
            mbar_wait(mybar + b, (n / SFX_NBUF) & 1);
            load_row_regs<T>(reinterpret_cast<const T*>(mybuf + (size_t)b * ROWB), lane, vals);
            __syncwarp();
            if (lane == 0 && i + SFX_NBUF < count) {
                fence_proxy_async();
                issue(i + SFX_NBUF);
            }
            fn(warp + i * NSTREAM, vals);
        }
        __syncwarp();
        if (lane == 0) ws.fills[warp] = n0 + count;
    } else {
        for (int i = 0; i < count; ++i) {
            int r = warp + i * NSTREAM;
            load_row_regs<T>(rowptr(r), lane, vals);
            fn(r, vals);
        }
    }
}

template <typename T>
__device__ __forceinline__ void blend_forward(const ModelView<T>& M, Scratch<T>& S, void* wsp) {
    StreamWS& ws = *reinterpret_cast<StreamWS*>(wsp);
    const int lane = threadIdx.x & 31;
    T c[RowVec<T>::NE];
#pragma unroll
    for (int i = 0; i < RowVec<T>::NV; ++i)
#pragma unroll
        for (int e = 0; e < RowVec<T>::VEC; ++e) c[i * RowVec<T>::VEC + e] = S.c[SFX_ELEM(i, lane, e)];
    const T* PK = M.PK;
    const int* vid = S.vid;
    auto rowptr = [=](int r) { return PK + ((long)vid[r / 3] * 3 + (r % 3)) * SFX_KPAD; };
    stream_rows<T>(SFX_NSLOT * 3, rowptr, ws, [&](int r, const T* v) {
        T acc = 0;
#pragma unroll
        for (int i = 0; i < RowVec<T>::NE; ++i) acc += v[i] * c[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) S.vp[r] = S.vt_s[r] + acc;
    });
}

template <typename T>
__device__ __forceinline__ void blend_adjoint(const ModelView<T>& M, Scratch<T>& S, void* wsp) {
    StreamWS& ws = *reinterpret_cast<StreamWS*>(wsp);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T acc[RowVec<T>::NE];
#pragma unroll
    for (int i = 0; i < RowVec<T>::NE; ++i) acc[i] = 0;
    const T* PK = M.PK;
    const int* vid = S.vid;
    auto rowptr = [=](int r) { return PK + ((long)vid[r / 3] * 3 + (r % 3)) * SFX_KPAD; };
    stream_rows<T>(SFX_NSLOT * 3, rowptr, ws, [&](int r, const T* v) {
        const T w = S.dvp[r];
#pragma unroll
        for (int i = 0; i < RowVec<T>::NE; ++i) acc[i] += v[i] * w;
    });
    // cross-warp reduction in a fixed order; the partial sums reuse the ring storage
    __syncthreads();
    T* part = reinterpret_cast<T*>(ws.ring);
#pragma unroll
    for (int i = 0; i < RowVec<T>::NV; ++i)
#pragma unroll
        for (int e = 0; e < RowVec<T>::VEC; ++e)
            part[warp * SFX_KPAD + SFX_ELEM(i, lane, e)] = acc[i * RowVec<T>::VEC + e];
    __syncthreads();
    for (int k = threadIdx.x; k < SFX_KPAD; k += blockDim.x) {
        T s = 0;
#pragma unroll
        for (int w = 0; w < SFX_NWARP; ++w) s += part[w * SFX_KPAD + k];
        S.dc[k] = s;
    }
    __syncthreads();
    if (ws.ring_mode) fence_proxy_async();   // generic writes to the ring precede later bulk copies
}


// ---- dense 512-wide rows (VPoser decoder layers): y = bias + W x  and  out = W^T g -----------
template <typename T>
__device__ __forceinline__ void rows_dot(const T* W, const T* bias, int nrows, const T* x, T* y,
                                         void* wsp) {
    StreamWS& ws = *reinterpret_cast<StreamWS*>(wsp);
    const int lane = threadIdx.x & 31;
    T c[RowVec<T>::NE];
#pragma unroll
    for (int i = 0; i < RowVec<T>::NV; ++i)
#pragma unroll
        for (int e = 0; e < RowVec<T>::VEC; ++e) c[i * RowVec<T>::VEC + e] = x[SFX_ELEM(i, lane, e)];
    auto rowptr = [=](int r) { return W + (long)r * SFX_KPAD; };
    stream_rows<T>(nrows, rowptr, ws, [&](int r, const T* v) {
        T acc = 0;
#pragma unroll
        for (int i = 0; i < RowVec<T>::NE; ++i) acc += v[i] * c[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[r] = bias[r] + acc;
    });
    __syncthreads();
}

template <typename T>
__device__ __forceinline__ void rows_accum(const T* W, int nrows, const T* g, T* out, void* wsp) {
    StreamWS& ws = *reinterpret_cast<StreamWS*>(wsp);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T acc[RowVec<T>::NE];
#pragma unroll
    for (int i = 0; i < RowVec<T>::NE; ++i) acc[i] = 0;
    auto rowptr = [=](int r) { return W + (long)r * SFX_KPAD; };
    stream_rows<T>(nrows, rowptr, ws, [&](int r, const T* v) {
        const T w = g[r];
#pragma unroll
        for (int i = 0; i < RowVec<T>::NE; ++i) acc[i] += v[i] * w;
    });
    __syncthreads();
    T* part = reinterpret_cast<T*>(ws.ring);
#pragma unroll
    for (int i = 0; i < RowVec<T>::NV; ++i)
#pragma unroll
        for (int e = 0; e < RowVec<T>::VEC; ++e)
            part[warp * SFX_KPAD + SFX_ELEM(i, lane, e)] = acc[i * RowVec<T>::VEC + e];
    __syncthreads();
    for (int k = threadIdx.x; k < SFX_KPAD; k += blockDim.x) {
        T s = 0;
#pragma unroll
        for (int w = 0; w < SFX_NWARP; ++w) s += part[w * SFX_KPAD + k];
        out[k] = s;
    }
    __syncthreads();
    if (ws.ring_mode) fence_proxy_async();
}

template <typename T>
__host__ __device__ inline size_t stream_smem_bytes(int ring_mode) {
    size_t ring = ring_mode ? (size_t)SFX_NWARP * SFX_NBUF * SFX_KPAD * sizeof(T)
                            : (size_t)SFX_NWARP * SFX_KPAD * sizeof(T);
    return ring + SFX_NWARP * SFX_NBUF * sizeof(uint64_t) + SFX_NWARP * sizeof(unsigned int) + 64;
}

}  // namespace sfx
#endif
