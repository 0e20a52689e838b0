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
#include <type_traits>
#include "sfx_core.cuh"

namespace sfx {

#define SFX_NWARP 16
#define SFX_NBUF 4
#ifndef SFX_NFLY
#define SFX_NFLY SFX_NBUF      // rows a warp keeps in flight (<= SFX_NBUF; profiles/microbench/l2_per_sm.cu)
#endif

struct StreamWS {
    unsigned char* ring;        // ring mode: [NWARP][NBUF][512*sizeof(T)]; else partials only
    uint64_t* bars;             // [NWARP][NBUF]
    unsigned int* fills;        // [NWARP] number of buffers filled so far (phase tracking)
    uint64_t* tl_bars;          // [SFX_TL_GROUPS] two-loop history staging (see two_loop_staged)
    unsigned int* tl_calls;     // [1] staged two-loop calls so far (phase tracking)
    int ring_mode;
    int tl_generic;             // debug: the generic-pointer staged recursion (A/B test)
    int stream_regs;            // rows travel global -> registers, several in flight per warp (no ring)
    int tl_shfl;                // A/B: shuffle butterfly instead of the shared-memory tree in the two-loop
    int wide;                   // > 1: this block leads a cluster of `wide` CTAs (see "wide frames" below)
    struct WideCtl* wc;
};
#define SFX_TL_GROUPS 8

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
        if (threadIdx.x < SFX_TL_GROUPS) mbar_init(ws.tl_bars + threadIdx.x, 1);
        if (threadIdx.x == 0) ws.tl_calls[0] = 0;
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

// Generic driver: FN(row_index_r, const T* row_values_in_regs[16], aux) is called by every lane
// of the warp that owns support row r, rows taken in the fixed order r = warp, warp + NWARP, ...
// AUX(r) is a per-row scalar (e.g. the adjoint weight of the row) that may live in global
// memory: lane l fetches the value of the warp's row i0 + l once per 32 rows and the warp reads
// it back by shuffle, so its latency is paid once per 32 rows instead of on every row.
struct NoAux {
    __device__ __forceinline__ float operator()(int) const { return 0.f; }
};

// The rows of a pass are numbered 0 .. n-1; streaming warp w (0 .. 14) owns `count` of them, its
// i-th one being row index(w, i).  Strided ownership (w, w + 15, ...) is the default.
struct StridedRows {
    int nrows;
    __device__ __forceinline__ int count(int w) const { return (nrows - w + SFX_NSTREAM - 1) / SFX_NSTREAM; }
    __device__ __forceinline__ int index(int w, int i) const { return w + i * SFX_NSTREAM; }
};
// Live support rows grouped by warp (Scratch::rows / wptr), optionally followed by `extra`
// further rows numbered from `base` on and owned in strided fashion.
struct GroupedRows {
    const unsigned short* wptr;
    int base, extra;
    __device__ __forceinline__ int count(int w) const {
        return (int)wptr[w + 1] - (int)wptr[w] + (extra - w + SFX_NSTREAM - 1) / SFX_NSTREAM;
    }
    __device__ __forceinline__ int index(int w, int i) const {
        const int own = (int)wptr[w + 1] - (int)wptr[w];
        return i < own ? (int)wptr[w] + i : base + w + (i - own) * SFX_NSTREAM;
    }
};

template <typename T, typename OWN, typename RP, typename FN, typename AUX>
__device__ __forceinline__ void stream_rows_own(OWN own, RP rowptr, StreamWS& ws, FN fn, AUX auxf,
                                                bool use_aux) {
    static_assert(SFX_NSTREAM == SFX_NWARP - 1, "one warp runs the kinematic chain");
    // warp 0 runs the kinematic chain while warps 1 .. NWARP-1 stream the rows
    const int warp = (threadIdx.x >> 5) - 1, lane = threadIdx.x & 31;
    const int count = warp >= 0 ? own.count(warp) : 0;
    constexpr uint32_t ROWB = SFX_KPAD * sizeof(T);
    T vals[RowVec<T>::NE];
    if (count == 0) return;
    T auxv = 0;
    auto aux_at = [&](int i) -> T {
        if (!use_aux) return (T)0;
        if ((i & 31) == 0) auxv = i + lane < count ? (T)auxf(own.index(warp, i + lane)) : (T)0;
        return __shfl_sync(0xffffffffu, auxv, i & 31);
    };
    if (ws.stream_regs) {
        // software pipeline in registers: PF rows of the warp in flight as plain 16-byte loads
        constexpr int PF = 3;
        T buf[PF][RowVec<T>::NE];
#pragma unroll
        for (int u = 0; u < PF; ++u)
            if (u < count) load_row_regs<T>(rowptr(own.index(warp, u)), lane, buf[u]);
        for (int i0 = 0; i0 < count; i0 += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int i = i0 + u;
                if (i < count) {
                    const T aux = aux_at(i);
#pragma unroll
                    for (int e = 0; e < RowVec<T>::NE; ++e) vals[e] = buf[u][e];
                    if (i + PF < count) load_row_regs<T>(rowptr(own.index(warp, i + PF)), lane, buf[u]);
                    fn(own.index(warp, i), vals, aux);
                }
            }
        }
        return;
    }
    if (ws.ring_mode) {
        unsigned char* mybuf = ws.ring + (size_t)warp * SFX_NBUF * ROWB;
        uint64_t* mybar = ws.bars + warp * SFX_NBUF;
        SFX_ASSUME_SHARED_PTR(mybuf);
        SFX_ASSUME_SHARED_PTR(mybar);
        SFX_ASSUME_SHARED_PTR(ws.fills);
        const unsigned int n0 = ws.fills[warp];
        auto issue = [&](int i) {
            int r = own.index(warp, i);
            unsigned int n = n0 + i;
            int b = n % SFX_NBUF;
            mbar_expect_tx(mybar + b, ROWB);
            bulk_g2s(mybuf + (size_t)b * ROWB, rowptr(r), ROWB, mybar + b);
        };
        if (lane == 0) {
            fence_proxy_async();
            for (int i = 0; i < SFX_NFLY && i < count; ++i) issue(i);
        }
        for (int i = 0; i < count; ++i) {
            const T aux = aux_at(i);
            unsigned int n = n0 + i;
            int b = n % SFX_NBUF;
            mbar_wait(mybar + b, (n / SFX_NBUF) & 1);
            load_row_regs<T>(reinterpret_cast<const T*>(mybuf + (size_t)b * ROWB), lane, vals);
            __syncwarp();
            if (lane == 0 && i + SFX_NFLY < count) {
                fence_proxy_async();
                issue(i + SFX_NFLY);
            }
            fn(own.index(warp, i), vals, aux);
        }
        __syncwarp();
        if (lane == 0) ws.fills[warp] = n0 + count;
    } else {
        for (int i = 0; i < count; ++i) {
            const T aux = aux_at(i);
            int r = own.index(warp, i);
            load_row_regs<T>(rowptr(r), lane, vals);
            fn(r, vals, aux);
        }
    }
}

template <typename T, typename RP, typename FN, typename AUX>
__device__ __forceinline__ void stream_rows_aux(int nrows, RP rowptr, StreamWS& ws, FN fn, AUX auxf,
                                                bool use_aux) {
    stream_rows_own<T>(StridedRows{nrows}, rowptr, ws, fn, auxf, use_aux);
}

template <typename T, typename RP, typename FN>
__device__ __forceinline__ void stream_rows(int nrows, RP rowptr, StreamWS& ws, FN fn) {
    stream_rows_aux<T>(nrows, rowptr, ws, [&](int r, const T* v, T) { fn(r, v); }, NoAux(), false);
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
    const unsigned short* rows = S.rows;          // the stage's live support rows
    auto rowptr = [=](int i) {
        const int r = rows[i];
        return PK + ((long)vid[r / 3] * 3 + (r % 3)) * SFX_KPAD;
    };
    stream_rows_own<T>(GroupedRows{S.wptr, 0, 0}, rowptr, ws, [&](int i, const T* v, T) {
        T acc = 0;
#pragma unroll
        for (int k = 0; k < RowVec<T>::NE; ++k) acc += v[k] * c[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const int r = rows[i];
        if (lane == 0) S.vp[r] = S.vt_s[r] + acc;
    }, NoAux(), false);
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
    const unsigned short* rows = S.rows;
    auto rowptr = [=](int i) {
        const int r = rows[i];
        return PK + ((long)vid[r / 3] * 3 + (r % 3)) * SFX_KPAD;
    };
    stream_rows_own<T>(GroupedRows{S.wptr, 0, 0}, rowptr, ws, [&](int i, const T* v, T) {
        const T w = S.dvp[rows[i]];
#pragma unroll
        for (int k = 0; k < RowVec<T>::NE; ++k) acc[k] += v[k] * w;
    }, NoAux(), false);
    // cross-warp reduction in a fixed order; the partial sums reuse the ring storage
    __syncthreads();
    T* part = reinterpret_cast<T*>(ws.ring);
    SFX_ASSUME_SHARED_PTR(part);
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


// every row of the blend matrix (interpenetration term): vp_g[r] = vt[r] + PK[r] . c
template <typename T>
__device__ __forceinline__ void blend_forward_full(const ModelView<T>& M, Scratch<T>& S, void* wsp,
                                                   T* vp_g) {
    StreamWS& ws = *reinterpret_cast<StreamWS*>(wsp);
    const int lane = threadIdx.x & 31;
    T c[RowVec<T>::NE];
#pragma unroll
    for (int i = 0; i < RowVec<T>::NV; ++i)
#pragma unroll
        for (int e = 0; e < RowVec<T>::VEC; ++e) c[i * RowVec<T>::VEC + e] = S.c[SFX_ELEM(i, lane, e)];
    const T* PK = M.PK;
    const T* vt = M.vt;
    auto rowptr = [=](int r) { return PK + (long)r * SFX_KPAD; };
    stream_rows_aux<T>(3 * M.V, rowptr, ws, [&](int r, const T* v, T base) {
        T acc = 0;
#pragma unroll
        for (int i = 0; i < RowVec<T>::NE; ++i) acc += v[i] * c[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) vp_g[r] = base + acc;
    }, [=](int r) { return vt[r]; }, true);
}

// support rows followed by the three rows of every touched vertex tv[t]; the adjoint weights of
// those rows are dvpc[3 t + k] (compact order)
template <typename T>
__device__ __forceinline__ void blend_adjoint_ext(const ModelView<T>& M, Scratch<T>& S, void* wsp,
                                                  const unsigned short* tv, const T* dvpc) {
    StreamWS& ws = *reinterpret_cast<StreamWS*>(wsp);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T acc[RowVec<T>::NE];
#pragma unroll
    for (int i = 0; i < RowVec<T>::NE; ++i) acc[i] = 0;
    const T* PK = M.PK;
    const int* vid = S.vid;
    const T* dvp = S.dvp;
    const unsigned short* rows = S.rows;
    const int NS3 = S.n_rows;                     // live support rows first, then the touched vertices'
    auto rowptr = [=](int i) {
        long row;
        if (i < NS3) {
            const int r = rows[i];
            row = (long)vid[r / 3] * 3 + (r % 3);
        } else {
            row = (long)tv[(i - NS3) / 3] * 3 + ((i - NS3) % 3);
        }
        return PK + row * SFX_KPAD;
    };
    stream_rows_own<T>(GroupedRows{S.wptr, NS3, 3 * S.n_touch}, rowptr, ws, [&](int i, const T* v, T w) {
#pragma unroll
        for (int k = 0; k < RowVec<T>::NE; ++k) acc[k] += v[k] * w;
    }, [=](int i) { return i < NS3 ? dvp[rows[i]] : dvpc[i - NS3]; }, true);
    __syncthreads();
    T* part = reinterpret_cast<T*>(ws.ring);
    SFX_ASSUME_SHARED_PTR(part);
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
    if (ws.ring_mode) fence_proxy_async();
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
    SFX_ASSUME_SHARED_PTR(part);
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


// ---- L-BFGS two-loop recursion with the history staged in shared memory --------------------
// The (s, y) pairs live in global memory (154 KB per frame); reading them pair by pair puts an
// L2 round trip on the critical path of every step of the recursion.  Here lane 0 issues one TMA
// bulk copy per history row at the start -- all 2k of them in flight at once, into the blend
// ring that is idle at this point -- grouped on 8 mbarriers in consumption order; the first loop
// starts as soon as the newest group has landed and the second loop re-reads the same shared
// memory, so the history crosses the L2->SM path once per iteration instead of twice.
// Arithmetic and summation order are those of two_loop_warp_n / block_reduce (bit-identical).
template <typename T, int NR>
__device__ __noinline__ void two_loop_staged_n(Scratch<T>& S, int k, int head, int H, T hd,
                                               const T* __restrict__ hist_s,
                                               const T* __restrict__ hist_y, int D, StreamWS& ws) {
    SFX_ASSUME_SHARED(S);
    SFX_ASSUME_SHARED(ws);
    const int lane = threadIdx.x;
    const uint32_t RB = (uint32_t)((D * sizeof(T) + 15) & ~(size_t)15);
    const int per = (k + SFX_TL_GROUPS - 1) / SFX_TL_GROUPS;
    const unsigned int parity = ws.tl_calls[0] & 1;
    unsigned char* base = ws.ring;
    if (lane == 0) {
        fence_proxy_async();
        int pos = (head + k - 1) % H;
        for (int g = 0; g < SFX_TL_GROUPS; ++g) {
            const int first = g * per;
            const int n = first < k ? (k - first < per ? k - first : per) : 0;
            mbar_expect_tx(ws.tl_bars + g, (uint32_t)n * 2 * RB);
            for (int j = 0; j < n; ++j) {
                unsigned char* dst = base + (size_t)(first + j) * 2 * RB;
                bulk_g2s(dst, hist_s + (long)pos * SFX_NP_MAX, RB, ws.tl_bars + g);
                bulk_g2s(dst + RB, hist_y + (long)pos * SFX_NP_MAX, RB, ws.tl_bars + g);
                pos = pos == 0 ? H - 1 : pos - 1;
            }
        }
    }
    __syncwarp();
    const bool tail_live = 32 * (NR - 1) + lane < D;
    T q[NR], sc[NR], yc[NR], sn[NR], yn[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int e = 32 * r + lane;
        q[r] = e < D ? -S.g[e] : (T)0;
    }
#define SFX_TL_LOAD(dsts, dsty, idx)                                                \
    do {                                                                            \
        const T* ps_ = reinterpret_cast<const T*>(base + (size_t)(idx) * 2 * RB) + lane;      \
        const T* py_ = reinterpret_cast<const T*>(base + (size_t)(idx) * 2 * RB + RB) + lane; \
        _Pragma("unroll") for (int r = 0; r < NR - 1; ++r) {                        \
            dsts[r] = ps_[32 * r];                                                  \
            dsty[r] = py_[32 * r];                                                  \
        }                                                                           \
        dsts[NR - 1] = tail_live ? ps_[32 * (NR - 1)] : (T)0;                       \
        dsty[NR - 1] = tail_live ? py_[32 * (NR - 1)] : (T)0;                       \
    } while (0)
    // ---- first loop: consumption index idx = 0 .. k-1  <->  pair i = k-1-idx ----
    mbar_wait(ws.tl_bars, parity);
    SFX_TL_LOAD(sn, yn, 0);
    for (int idx = 0; idx < k; ++idx) {
        const int i = k - 1 - idx;
#pragma unroll
        for (int r = 0; r < NR; ++r) { sc[r] = sn[r]; yc[r] = yn[r]; }
        if (idx + 1 < k) {
            if ((idx + 1) % per == 0) mbar_wait(ws.tl_bars + (idx + 1) / per, parity);
            SFX_TL_LOAD(sn, yn, idx + 1);
        }
        T p = 0;
#pragma unroll
        for (int r = 0; r < NR; ++r) p = p + sc[r] * q[r];
        p = warp_tree_sum(p, S.tl_red + 32 * (i & 1), lane);
        const T a = p * S.ro[i];
        if (lane == 0) S.al[i] = a;
#pragma unroll
        for (int r = 0; r < NR; ++r) q[r] += -a * yc[r];
    }
    // groups that hold no pair still complete their phase: wait so every barrier is observed
    for (int g = (k + per - 1) / per; g < SFX_TL_GROUPS; ++g) mbar_wait(ws.tl_bars + g, parity);
#pragma unroll
    for (int r = 0; r < NR; ++r) q[r] = q[r] * hd;       // q now holds the direction d
    __syncwarp();
    // ---- second loop: pair i = 0 .. k-1  <->  idx = k-1-i (already in shared memory) ----
    SFX_TL_LOAD(sn, yn, k - 1);
    for (int i = 0; i < k; ++i) {
#pragma unroll
        for (int r = 0; r < NR; ++r) { sc[r] = sn[r]; yc[r] = yn[r]; }
        if (i + 1 < k) SFX_TL_LOAD(sn, yn, k - 2 - i);
        T p = 0;
#pragma unroll
        for (int r = 0; r < NR; ++r) p = p + yc[r] * q[r];
        p = warp_tree_sum(p, S.tl_red + 32 * (i & 1), lane);
        const T co = S.al[i] - p * S.ro[i];
#pragma unroll
        for (int r = 0; r < NR; ++r) q[r] += co * sc[r];
    }
#undef SFX_TL_LOAD
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int e = 32 * r + lane;
        if (e < D) S.d[e] = q[r];
    }
    __syncwarp();
    if (lane == 0) ws.tl_calls[0] += 1;
}

// ---- the same recursion, float32 only, written against shared-state-space addresses ---------
// Every step of the recursion is one dependent chain (dot product -> warp sum -> scalar ->
// axpy), executed by a single warp while the other fifteen wait: its length is what the longest
// frames of a batch pay per L-BFGS iteration.  This version computes the shared-memory addresses
// once (the generic-pointer version re-derives the shared window with S2R on every access),
// counts the staging groups down instead of dividing, and fetches the per-pair scalars ahead of
// the chain.  Arithmetic and summation order are unchanged (bit-identical, tested).
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
template <bool SHFL>
__device__ __forceinline__ float warp_tree_sum_sa(float p, uint32_t slot, int lane) {
    if (SHFL) {
        // the same association through five shuffles
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p = p + __shfl_xor_sync(0xffffffffu, p, o);
        return p;
    }
    sts_f32(slot + 4u * lane, p);
    __syncwarp();
    float a[32];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(a[4 * i]), "=f"(a[4 * i + 1]), "=f"(a[4 * i + 2]), "=f"(a[4 * i + 3])
                     : "r"(slot + 16u * i)
                     : "memory");
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int l = 0; l < o; ++l) a[l] = a[l] + a[l + o];
    return a[0];
}

template <int NR, bool SHFL>
__device__ __noinline__ void two_loop_staged_f32(Scratch<float>& S, int k, int head, int H, float hd,
                                                 const float* __restrict__ hist_s,
                                                 const float* __restrict__ hist_y, int D, StreamWS& ws) {
    SFX_ASSUME_SHARED(S);
    SFX_ASSUME_SHARED(ws);
    const int lane = threadIdx.x;
    const uint32_t RB = (uint32_t)((D * sizeof(float) + 15) & ~(size_t)15);
    const int per = (k + SFX_TL_GROUPS - 1) / SFX_TL_GROUPS;
    const unsigned int parity = ws.tl_calls[0] & 1;
    unsigned char* base = ws.ring;
    if (lane == 0) {
        fence_proxy_async();
        int pos = (head + k - 1) % H;
        for (int g = 0; g < SFX_TL_GROUPS; ++g) {
            const int first = g * per;
            const int n = first < k ? (k - first < per ? k - first : per) : 0;
            mbar_expect_tx(ws.tl_bars + g, (uint32_t)n * 2 * RB);
            for (int j = 0; j < n; ++j) {
                unsigned char* dst = base + (size_t)(first + j) * 2 * RB;
                bulk_g2s(dst, hist_s + (long)pos * SFX_NP_MAX, RB, ws.tl_bars + g);
                bulk_g2s(dst + RB, hist_y + (long)pos * SFX_NP_MAX, RB, ws.tl_bars + g);
                pos = pos == 0 ? H - 1 : pos - 1;
            }
        }
    }
    __syncwarp();
    const bool tail_live = 32 * (NR - 1) + lane < D;
    const uint32_t a_base = smem_u32(base) + 4u * lane;      // this lane's column of the staged pairs
    const uint32_t a_red = smem_u32(S.tl_red), a_ro = smem_u32(S.ro), a_al = smem_u32(S.al);
    float q[NR], sc[NR], yc[NR], sn[NR], yn[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int e = 32 * r + lane;
        q[r] = e < D ? -S.g[e] : 0.f;
    }
#define SFX_TL_LOAD(dsts, dsty, idx)                                                   \
    do {                                                                               \
        const uint32_t as_ = a_base + (uint32_t)(idx) * 2u * RB, ay_ = as_ + RB;       \
        _Pragma("unroll") for (int r = 0; r < NR - 1; ++r) {                           \
            dsts[r] = lds_f32(as_ + 128u * r);                                         \
            dsty[r] = lds_f32(ay_ + 128u * r);                                         \
        }                                                                              \
        dsts[NR - 1] = tail_live ? lds_f32(as_ + 128u * (NR - 1)) : 0.f;               \
        dsty[NR - 1] = tail_live ? lds_f32(ay_ + 128u * (NR - 1)) : 0.f;               \
    } while (0)
    // ---- first loop: consumption index idx = 0 .. k-1  <->  pair i = k-1-idx ----
    mbar_wait(ws.tl_bars, parity);
    SFX_TL_LOAD(sn, yn, 0);
    int next_group_at = per, group = 1;
    for (int idx = 0; idx < k; ++idx) {
        const int i = k - 1 - idx;
        const float ro_i = lds_f32(a_ro + 4u * i);
#pragma unroll
        for (int r = 0; r < NR; ++r) { sc[r] = sn[r]; yc[r] = yn[r]; }
        if (idx + 1 < k) {
            if (idx + 1 == next_group_at) {
                mbar_wait(ws.tl_bars + group, parity);
                group += 1;
                next_group_at += per;
            }
            SFX_TL_LOAD(sn, yn, idx + 1);
        }
        float p = 0;
#pragma unroll
        for (int r = 0; r < NR; ++r) p = p + sc[r] * q[r];
        p = warp_tree_sum_sa<SHFL>(p, a_red + 128u * (i & 1), lane);
        const float a = p * ro_i;
        if (lane == 0) sts_f32(a_al + 4u * i, a);
#pragma unroll
        for (int r = 0; r < NR; ++r) q[r] += -a * yc[r];
    }
    // groups that hold no pair still complete their phase: wait so every barrier is observed
    for (int g = (k + per - 1) / per; g < SFX_TL_GROUPS; ++g) mbar_wait(ws.tl_bars + g, parity);
#pragma unroll
    for (int r = 0; r < NR; ++r) q[r] = q[r] * hd;       // q now holds the direction d
    __syncwarp();
    // ---- second loop: pair i = 0 .. k-1  <->  idx = k-1-i (already in shared memory) ----
    SFX_TL_LOAD(sn, yn, k - 1);
    for (int i = 0; i < k; ++i) {
        const float ro_i = lds_f32(a_ro + 4u * i), al_i = lds_f32(a_al + 4u * i);
#pragma unroll
        for (int r = 0; r < NR; ++r) { sc[r] = sn[r]; yc[r] = yn[r]; }
        if (i + 1 < k) SFX_TL_LOAD(sn, yn, k - 2 - i);
        float p = 0;
#pragma unroll
        for (int r = 0; r < NR; ++r) p = p + yc[r] * q[r];
        p = warp_tree_sum_sa<SHFL>(p, a_red + 128u * (i & 1), lane);
        const float co = al_i - p * ro_i;
#pragma unroll
        for (int r = 0; r < NR; ++r) q[r] += co * sc[r];
    }
#undef SFX_TL_LOAD
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int e = 32 * r + lane;
        if (e < D) S.d[e] = q[r];
    }
    __syncwarp();
    if (lane == 0) ws.tl_calls[0] += 1;
}

template <typename T, int NR>
__device__ __forceinline__ void two_loop_staged_pick(Scratch<T>& S, int k, int head, int H, T hd,
                                                     const T* hist_s, const T* hist_y, int D, StreamWS& ws) {
    if constexpr (sizeof(T) == 4) {
        if (!ws.tl_generic) {
            if (ws.tl_shfl) two_loop_staged_f32<NR, true>(S, k, head, H, hd, hist_s, hist_y, D, ws);
            else two_loop_staged_f32<NR, false>(S, k, head, H, hd, hist_s, hist_y, D, ws);
            return;
        }
    }
    two_loop_staged_n<T, NR>(S, k, head, H, hd, hist_s, hist_y, D, ws);
}

// returns false when the staged path does not apply (plain-load mode, or history too large for
// the ring: float64, more than 128 active parameters with a long history)
template <typename T>
__device__ __forceinline__ bool two_loop_staged(Scratch<T>& S, int k, int head, int H, T hd,
                                                const T* hist_s, const T* hist_y, int D, void* wsp) {
    StreamWS& ws = *reinterpret_cast<StreamWS*>(wsp);
    const size_t RB = (D * sizeof(T) + 15) & ~(size_t)15;
    const size_t ring_bytes = (size_t)SFX_NWARP * SFX_NBUF * SFX_KPAD * sizeof(T);
    if (!ws.ring_mode || k < 1 || (size_t)k * 2 * RB > ring_bytes || D > 128) return false;
    if (D <= 32) two_loop_staged_pick<T, 1>(S, k, head, H, hd, hist_s, hist_y, D, ws);
    else two_loop_staged_pick<T, 4>(S, k, head, H, hd, hist_s, hist_y, D, ws);
    return true;
}

// ---- scalar recurrences of the Gram two-loop (sfx_core.cuh gram_chain), float32, written against
// shared-state-space addresses.  G is the packed [SFX_GRAM_ROWS][SFX_GRAM_LDF] block (zero outside
// the live k x k part), so a column is one walking address + immediates and the history is cut
// into four 32-pair segments whose owner register is known at compile time (no selects on the
// dependent chain).
//
// Blocked by four: the recursion's critical path is  b_j -> broadcast -> al_j -> b_(j-1) -> ...,
// one shuffle (~25 cycles) + two dependent floating-point operations per pair when taken pair by
// pair.  Here the four values b_j .. b_(j-3) are broadcast at once and every lane runs the 4 x 4
// triangular piece of the recurrence itself (six multiply-adds on values it loaded ahead of the
// chain), so the chain pays one shuffle latency per FOUR pairs; the rank-4 update of the remaining
// entries follows off the critical path.  Every entry still receives the same multiply-adds in
// the same order as in gram_chain<float> (pair j descending in the first loop, ascending in the
// second): bit-identical to it (tests/test_gpu_parity.py, staging variants).
template <int JR>
__device__ __forceinline__ void gram_first_step(float (&b)[4], float (&e)[4], uint32_t colA, uint32_t rowA,
                                                uint32_t a_ro, uint32_t a_al, int j, int lane) {
    constexpr uint32_t LDB = 4u * SFX_GRAM_LDF;
    float colv[4], rowv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) colv[r] = lds_f32(colA + 32u * LDB * r);
#pragma unroll
    for (int r = 0; r < 4; ++r) rowv[r] = r <= JR ? lds_f32(rowA + 128u * r) : 0.f;
    const float ro_j = lds_f32(a_ro + 4u * j);
    const float a = __shfl_sync(0xffffffffu, b[JR], j & 31) * ro_j;
    if (lane == 0) sts_f32(a_al + 4u * j, a);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (r < JR) {
            b[r] -= a * colv[r];
            e[r] -= a * rowv[r];
        } else if (r == JR) {
            const bool lo = lane < (j & 31);
            b[r] -= a * (lo ? colv[r] : 0.f);
            e[r] -= a * (lo ? rowv[r] : colv[r]);
        } else {
            e[r] -= a * colv[r];
        }
    }
}

// pairs j, j-1, j-2, j-3 (j % 4 == 3) of segment JR
template <int JR>
__device__ __forceinline__ void gram_first_block(float (&b)[4], float (&e)[4], uint32_t gA, uint32_t colA,
                                                 uint32_t rowA, uint32_t a_ro, uint32_t a_al, int j, int lane) {
    constexpr uint32_t LDB = 4u * SFX_GRAM_LDF;
    float col[4][4], row[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int r = 0; r < 4; ++r) col[u][r] = lds_f32(colA - 4u * u + 32u * LDB * r);
#pragma unroll
        for (int r = 0; r < 4; ++r) row[u][r] = r <= JR ? lds_f32(rowA - LDB * u + 128u * r) : 0.f;
    }
    // the 4 x 4 piece: G(j - v, j - u) for u < v (row below the column: s_i . y_j), and 1 / (y.s)
    const uint32_t gjj = gA + LDB * (uint32_t)j + 4u * (uint32_t)j;
    const float g10 = lds_f32(gjj - LDB), g20 = lds_f32(gjj - 2u * LDB), g30 = lds_f32(gjj - 3u * LDB);
    const float g21 = lds_f32(gjj - 2u * LDB - 4u), g31 = lds_f32(gjj - 3u * LDB - 4u);
    const float g32 = lds_f32(gjj - 3u * LDB - 8u);
    const float ro0 = lds_f32(a_ro + 4u * j), ro1 = lds_f32(a_ro + 4u * j - 4u);
    const float ro2 = lds_f32(a_ro + 4u * j - 8u), ro3 = lds_f32(a_ro + 4u * j - 12u);
    const int l0 = j & 31;
    float b0 = __shfl_sync(0xffffffffu, b[JR], l0), b1 = __shfl_sync(0xffffffffu, b[JR], l0 - 1);
    float b2 = __shfl_sync(0xffffffffu, b[JR], l0 - 2), b3 = __shfl_sync(0xffffffffu, b[JR], l0 - 3);
    float a[4];
    a[0] = b0 * ro0;
    b1 -= a[0] * g10;
    b2 -= a[0] * g20;
    b3 -= a[0] * g30;
    a[1] = b1 * ro1;
    b2 -= a[1] * g21;
    b3 -= a[1] * g31;
    a[2] = b2 * ro2;
    b3 -= a[2] * g32;
    a[3] = b3 * ro3;
    if (lane == 0) {
        sts_f32(a_al + 4u * (uint32_t)j, a[0]);
        sts_f32(a_al + 4u * (uint32_t)j - 4u, a[1]);
        sts_f32(a_al + 4u * (uint32_t)j - 8u, a[2]);
        sts_f32(a_al + 4u * (uint32_t)j - 12u, a[3]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (r < JR) {
                b[r] -= a[u] * col[u][r];
                e[r] -= a[u] * row[u][r];
            } else if (r == JR) {
                const bool lo = lane < l0 - u;
                b[r] -= a[u] * (lo ? col[u][r] : 0.f);
                e[r] -= a[u] * (lo ? row[u][r] : col[u][r]);
            } else {
                e[r] -= a[u] * col[u][r];
            }
        }
    }
}

template <int JR>
__device__ __forceinline__ void gram_second_step(float (&e)[4], uint32_t rowA, uint32_t a_ro, uint32_t a_al,
                                                 uint32_t a_cf, int j, int lane) {
    float rowv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) rowv[r] = r >= JR ? lds_f32(rowA + 128u * r) : 0.f;
    const float ro_j = lds_f32(a_ro + 4u * j), al_j = lds_f32(a_al + 4u * j);
    const float c = al_j - __shfl_sync(0xffffffffu, e[JR], j & 31) * ro_j;
    if (lane == 0) sts_f32(a_cf + 4u * j, c);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (r > JR) e[r] += c * rowv[r];
        else if (r == JR) e[r] += c * (lane > (j & 31) ? rowv[r] : 0.f);
    }
}

// pairs j, j+1, j+2, j+3 (j % 4 == 0) of segment JR
template <int JR>
__device__ __forceinline__ void gram_second_block(float (&e)[4], uint32_t gA, uint32_t rowA, uint32_t a_ro,
                                                  uint32_t a_al, uint32_t a_cf, int j, int lane) {
    constexpr uint32_t LDB = 4u * SFX_GRAM_LDF;
    float row[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int r = 0; r < 4; ++r) row[u][r] = r >= JR ? lds_f32(rowA + LDB * u + 128u * r) : 0.f;
    // G(j + u, j + v) for u < v (above the diagonal: s . y)
    const uint32_t gjj = gA + LDB * (uint32_t)j + 4u * (uint32_t)j;
    const float g01 = lds_f32(gjj + 4u), g02 = lds_f32(gjj + 8u), g03 = lds_f32(gjj + 12u);
    const float g12 = lds_f32(gjj + LDB + 8u), g13 = lds_f32(gjj + LDB + 12u);
    const float g23 = lds_f32(gjj + 2u * LDB + 12u);
    float ro[4], al[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        ro[u] = lds_f32(a_ro + 4u * (uint32_t)(j + u));
        al[u] = lds_f32(a_al + 4u * (uint32_t)(j + u));
    }
    const int l0 = j & 31;
    float e0 = __shfl_sync(0xffffffffu, e[JR], l0), e1 = __shfl_sync(0xffffffffu, e[JR], l0 + 1);
    float e2 = __shfl_sync(0xffffffffu, e[JR], l0 + 2), e3 = __shfl_sync(0xffffffffu, e[JR], l0 + 3);
    float c[4];
    c[0] = al[0] - e0 * ro[0];
    e1 += c[0] * g01;
    e2 += c[0] * g02;
    e3 += c[0] * g03;
    c[1] = al[1] - e1 * ro[1];
    e2 += c[1] * g12;
    e3 += c[1] * g13;
    c[2] = al[2] - e2 * ro[2];
    e3 += c[2] * g23;
    c[3] = al[3] - e3 * ro[3];
    if (lane == 0) {
        sts_f32(a_cf + 4u * (uint32_t)j, c[0]);
        sts_f32(a_cf + 4u * (uint32_t)j + 4u, c[1]);
        sts_f32(a_cf + 4u * (uint32_t)j + 8u, c[2]);
        sts_f32(a_cf + 4u * (uint32_t)j + 12u, c[3]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (r > JR) e[r] += c[u] * row[u][r];
            else if (r == JR) e[r] += c[u] * (lane > l0 + u ? row[u][r] : 0.f);
        }
    }
}

__device__ __forceinline__ void gram_chain_f32(Scratch<float>& S, int k, float hd, const float* G, int lane) {
    static_assert(SFX_HIST <= 128, "four 32-pair segments");
    constexpr uint32_t LDB = 4u * SFX_GRAM_LDF;          // bytes per row
    const uint32_t gA = smem_u32(G);
    const uint32_t a_ro = smem_u32(S.ro), a_al = smem_u32(S.al), a_cf = smem_u32(S.cf);
    float b[4], e[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = lane + 32 * r;
        b[r] = i < k ? -S.sg[i] : 0.f;
        e[r] = i < k ? -S.yg[i] : 0.f;
    }
    auto first_segment = [&](auto jr_c) {
        constexpr int JR = decltype(jr_c)::value;
        const int jlo = 32 * JR, jhi = k - 1 < jlo + 31 ? k - 1 : jlo + 31;
        uint32_t colA = gA + LDB * lane + 4u * jhi;      // element (lane, j); register r: + 32 rows
        uint32_t rowA = gA + LDB * jhi + 4u * lane;      // element (j, lane); register r: + 32 columns
        int j = jhi;
        for (; j >= jlo && (j & 3) != 3; --j) {          // ragged top of the segment
            gram_first_step<JR>(b, e, colA, rowA, a_ro, a_al, j, lane);
            colA -= 4u;
            rowA -= LDB;
        }
        for (; j >= jlo; j -= 4) {
            gram_first_block<JR>(b, e, gA, colA, rowA, a_ro, a_al, j, lane);
            colA -= 16u;
            rowA -= 4u * LDB;
        }
    };
    using C0 = std::integral_constant<int, 0>; using C1 = std::integral_constant<int, 1>;
    using C2 = std::integral_constant<int, 2>; using C3 = std::integral_constant<int, 3>;
    first_segment(C3()); first_segment(C2()); first_segment(C1()); first_segment(C0());
    __syncwarp();
#ifdef SFX_CHAIN_PROBE
    if (lane == 0) SFX_CHAIN_PROBE = clock64();
#endif
#pragma unroll
    for (int r = 0; r < 4; ++r) e[r] = e[r] * hd;
    auto second_segment = [&](auto jr_c) {
        constexpr int JR = decltype(jr_c)::value;
        const int jlo = 32 * JR, jhi = k - 1 < jlo + 31 ? k - 1 : jlo + 31;
        uint32_t rowA = gA + LDB * jlo + 4u * lane;
        int j = jlo;
        for (; j + 3 <= jhi; j += 4) {
            gram_second_block<JR>(e, gA, rowA, a_ro, a_al, a_cf, j, lane);
            rowA += 4u * LDB;
        }
        for (; j <= jhi; ++j) {                          // ragged end of the history
            gram_second_step<JR>(e, rowA, a_ro, a_al, a_cf, j, lane);
            rowA += LDB;
        }
    };
    second_segment(C0()); second_segment(C1()); second_segment(C2()); second_segment(C3());
    __syncwarp();
}


// ---- wide frames: one frame, a cluster of CTAs ------------------------------------------------
// A batch smaller than the GPU leaves SMs idle while its longest frames -- those that fit two
// orientations, known before the launch -- run on alone; most of their time is the two passes
// over the live rows of the blend matrix in the last annealing stage (675 rows x 2 KiB, twice per
// evaluation, through one SM's L2 port).  Such a frame is given a cluster of SFX_WIDE_CLUSTER
// CTAs: rank 0 (the leader) runs the whole per-frame flow exactly as a single block does, ranks
// 1.. (the helpers) keep the stage's live rows RESIDENT in their shared memory (7 x 110 rows
// hold all 675) and do the two passes on them when the leader asks, over distributed shared memory:
//
//   LOAD   helper pulls its share of the leader's row list (S.rows / S.vid / S.vt_s) and copies
//          those rows of PK from global memory into its shared memory (TMA bulk copies); once per
//          stage and whenever the contour look-up row changes
//   FWD    helper pulls c[512] from the leader, a warp per row: vp[r] = vt[r] + row . c, written
//          straight into the leader's S.vp
//   ADJ    helper pulls dvp of its rows, a thread per column: partial dc[k] = sum_rows row[k] dvp,
//          written into the leader's ring; the leader adds the helpers' partials in rank order
//
// Commands travel as a word in the helper's shared memory + a remote mbarrier arrive
// (release.cluster); completions as remote arrives on the leader's mbarrier.  The forward pass
// gives bit-identical results to the single-block path (a row's dot product does not depend on
// who computes it); the adjoint sums the rows in a different order (by helper, then row, instead
// of by streaming warp), so a wide frame's rounding differs from the same frame run by one
// block: an explicit option (SfxPipeline::n_wide), deterministic run to run.
#define SFX_WIDE_CLUSTER 8
#define SFX_WIDE_ROWS_MAX 110        // resident rows per helper (dynamic shared memory / 2 KiB)

struct WideCtl {                     // static shared memory: same offset in every CTA of the cluster
    uint64_t go_bar;                 // helper: the leader's commands arrive here
    uint64_t done_bar;               // leader: one arrival per helper and command
    uint64_t load_bar;               // helper: TMA completion of a LOAD
    int cmd, n_rows;                 // written by the leader (remote stores)
    unsigned int done_phase;         // leader: commands completed so far
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int ld_cluster_s32(uint32_t a) {
    int v;
    asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ unsigned short ld_cluster_u16(uint32_t a) {
    unsigned short v;
    asm volatile("ld.shared::cluster.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_cluster_v4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t a, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_s32(uint32_t a, int v) {
    asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Polls without ordering semantics and orders once after the phase has completed: an acquire at
// cluster scope invalidates the SM's L1 (CCTL.IVALL in SASS), which inside the polling loop cost the
// leader its cached model tables on every poll (20 % of the wide grid's stall samples,
// profiles/r02l_ncu_source_hotspots_pipeline_kernel.txt).
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ bool wide_active(void* wsp) {
    return reinterpret_cast<StreamWS*>(wsp)->wide > 1;
}

// leader: post `cmd` to every helper (threads 32 .. 32 + helpers - 1, one helper each).  The data
// the command refers to was written before the caller's last block barrier.
__device__ __forceinline__ void wide_post(StreamWS& ws, int cmd, int n_rows) {
    const int h = (int)threadIdx.x - 32;
    if (h >= 0 && h < ws.wide - 1) {
        WideCtl* wc = ws.wc;
        st_cluster_s32(mapa_u32(smem_u32(&wc->n_rows), h + 1), n_rows);
        st_cluster_s32(mapa_u32(smem_u32(&wc->cmd), h + 1), cmd);
        mbar_arrive_remote(mapa_u32(smem_u32(&wc->go_bar), h + 1));
    }
}
// leader, every thread: wait for the helpers' completion of the outstanding command
__device__ __forceinline__ void wide_wait(StreamWS& ws) {
    SFX_ASSUME_SHARED_PTR(ws.wc);
    const unsigned int p = ws.wc->done_phase;
    mbar_wait_cluster(&ws.wc->done_bar, p & 1);
}

template <typename T>
__device__ __forceinline__ void wide_begin(Scratch<T>& S, void* wsp, int cmd) {
    StreamWS& ws = *reinterpret_cast<StreamWS*>(wsp);
    // callers sit right behind a block barrier: S.c / S.dvp, S.rows_dirty and done_phase are settled
    if (cmd == SFX_WIDE_FWD && S.rows_dirty) {
        wide_post(ws, SFX_WIDE_LOAD, S.n_rows);
        wide_wait(ws);
        __syncthreads();
        if (threadIdx.x == 0) {
            ws.wc->done_phase += 1;
            S.rows_dirty = 0;
        }
        __syncthreads();
    }
    wide_post(ws, cmd, S.n_rows);
}

template <typename T>
__device__ __forceinline__ void wide_end_forward(Scratch<T>& S, void* wsp) {
    StreamWS& ws = *reinterpret_cast<StreamWS*>(wsp);
    wide_wait(ws);                    // S.vp of every live row was written by the helpers
    __syncthreads();
    if (threadIdx.x == 0) ws.wc->done_phase += 1;
}

template <typename T>
__device__ __forceinline__ void wide_end_adjoint(Scratch<T>& S, void* wsp) {
    StreamWS& ws = *reinterpret_cast<StreamWS*>(wsp);
    wide_wait(ws);
    __syncthreads();
    if (threadIdx.x == 0) ws.wc->done_phase += 1;
    const T* part = reinterpret_cast<const T*>(ws.ring);
    SFX_ASSUME_SHARED_PTR(part);
    const int nh = ws.wide - 1;
    for (int k = threadIdx.x; k < SFX_KPAD; k += blockDim.x) {
        T s = 0;
        for (int h = 0; h < nh; ++h) s += part[h * SFX_KPAD + k];
        S.dc[k] = s;
    }
    __syncthreads();
    if (ws.ring_mode) fence_proxy_async();   // generic reads of the ring precede later bulk copies
}

// helper CTA (cluster rank >= 1, float32): serves the leader until SFX_WIDE_EXIT.  `dyn` is the
// block's dynamic shared memory (rows, then the per-row tables); the leader's Scratch sits at the
// same offset of its own dynamic shared memory.
__device__ __noinline__ void wide_helper_main(const float* __restrict__ PK, unsigned char* dyn,
                                              size_t dyn_bytes, WideCtl* wc, int cl) {
    SFX_ASSUME_SHARED_PTR(dyn);
    SFX_ASSUME_SHARED_PTR(wc);
    const int rank = (int)cluster_ctarank();
    const int nh = cl - 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* rowbuf = reinterpret_cast<float*>(dyn);
    unsigned char* tab = dyn + (size_t)SFX_WIDE_ROWS_MAX * SFX_KPAD * sizeof(float);
    int* t_grow = reinterpret_cast<int*>(tab);                               // global row of PK
    float* t_vt = reinterpret_cast<float*>(tab + 4 * SFX_WIDE_ROWS_MAX);     // template coordinate
    float* t_dvp = reinterpret_cast<float*>(tab + 8 * SFX_WIDE_ROWS_MAX);    // adjoint weight (per pass)
    unsigned short* t_r = reinterpret_cast<unsigned short*>(tab + 12 * SFX_WIDE_ROWS_MAX);  // slot row
    float* t_c = rowbuf + (size_t)(SFX_WIDE_ROWS_MAX - 1) * SFX_KPAD;        // last row slot: copy of c
    // the leader's Scratch<float>, seen through the cluster window
    Scratch<float>* Sl = reinterpret_cast<Scratch<float>*>(dyn);
    const uint32_t a_rows = mapa_u32(smem_u32(Sl->rows), 0), a_vid = mapa_u32(smem_u32(Sl->vid), 0);
    const uint32_t a_vts = mapa_u32(smem_u32(Sl->vt_s), 0), a_c = mapa_u32(smem_u32(Sl->c), 0);
    const uint32_t a_vp = mapa_u32(smem_u32(Sl->vp), 0), a_dvp = mapa_u32(smem_u32(Sl->dvp), 0);
    // the leader's ring follows its Scratch (carve_stream): partial dc of helper `rank`
    const uint32_t a_part = mapa_u32(smem_u32(dyn + ((sizeof(Scratch<float>) + 1023) / 1024 * 1024)), 0) +
                            (uint32_t)(rank - 1) * SFX_KPAD * 4u;
    const uint32_t a_done = mapa_u32(smem_u32(&wc->done_bar), 0);
    unsigned int go_phase = 0, load_phase = 0;
    int cnt = 0;
    for (;;) {
        mbar_wait_cluster(&wc->go_bar, go_phase & 1);
        go_phase += 1;
        const int cmd = *reinterpret_cast<volatile int*>(&wc->cmd);
        if (cmd == SFX_WIDE_EXIT) break;
        if (cmd == SFX_WIDE_LOAD) {
            const int n = *reinterpret_cast<volatile int*>(&wc->n_rows);
            const int per = (n + nh - 1) / nh;
            const int lo = (rank - 1) * per;
            cnt = n - lo < per ? n - lo : per;
            if (cnt < 0) cnt = 0;
            if (cnt > SFX_WIDE_ROWS_MAX - 1) cnt = SFX_WIDE_ROWS_MAX - 1;   // cannot happen: 675 / 7 = 97
            __syncthreads();                       // the previous pass is done with the row buffer
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(&wc->load_bar, (uint32_t)cnt * SFX_KPAD * 4u);
            }
            if (tid < cnt) {
                const int r = (int)ld_cluster_u16(a_rows + 2u * (uint32_t)(lo + tid));
                const int vid = ld_cluster_s32(a_vid + 4u * (uint32_t)(r / 3));
                t_r[tid] = (unsigned short)r;
                t_grow[tid] = vid * 3 + (r % 3);
                t_vt[tid] = ld_cluster_f32(a_vts + 4u * (uint32_t)r);
            }
            __syncthreads();
            if (tid < cnt)
                bulk_g2s(rowbuf + (size_t)tid * SFX_KPAD, PK + (size_t)t_grow[tid] * SFX_KPAD,
                         SFX_KPAD * 4u, &wc->load_bar);
            mbar_wait(&wc->load_bar, load_phase & 1);   // (no rows: the expect_tx arrival completes it)
            load_phase += 1;
        } else if (cmd == SFX_WIDE_FWD) {
            // one copy of c per helper over the cluster network (2 KiB), then this lane's 16
            // coefficients in the element order of the streaming passes
            t_c[tid] = ld_cluster_f32(a_c + 4u * (uint32_t)tid);
            __syncthreads();
            float c[16];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) c[4 * i + e] = t_c[(i * 32 + lane) * 4 + e];
            for (int j = warp; j < cnt; j += SFX_NWARP) {
                const float4* row = reinterpret_cast<const float4*>(rowbuf + (size_t)j * SFX_KPAD);
                float acc = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = row[i * 32 + lane];
                    acc += v.x * c[4 * i];
                    acc += v.y * c[4 * i + 1];
                    acc += v.z * c[4 * i + 2];
                    acc += v.w * c[4 * i + 3];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) st_cluster_f32(a_vp + 4u * (uint32_t)t_r[j], t_vt[j] + acc);
            }
        } else if (cmd == SFX_WIDE_ADJ) {
            if (tid < cnt) t_dvp[tid] = ld_cluster_f32(a_dvp + 4u * (uint32_t)t_r[tid]);
            __syncthreads();
            float acc = 0;
            for (int j = 0; j < cnt; ++j) acc += rowbuf[(size_t)j * SFX_KPAD + tid] * t_dvp[j];
            st_cluster_f32(a_part + 4u * (uint32_t)tid, acc);
        }
        __syncthreads();                           // every thread's remote stores are issued
        if (tid == 0) mbar_arrive_remote(a_done);  // release.cluster: cumulative over the barrier
    }
}

template <typename T>
__device__ __forceinline__ unsigned char* idle_area(void* wsp, size_t* bytes) {
    StreamWS& ws = *reinterpret_cast<StreamWS*>(wsp);
    *bytes = ws.ring_mode ? (size_t)SFX_NWARP * SFX_NBUF * SFX_KPAD * sizeof(T)
                          : (size_t)SFX_NWARP * SFX_KPAD * sizeof(T);
    return ws.ring;
}
// the ring is written by TMA (async proxy) next: order the generic-proxy accesses before it
__device__ __forceinline__ void idle_area_release(void* wsp) {
    if (reinterpret_cast<StreamWS*>(wsp)->ring_mode) fence_proxy_async();
}

template <typename T>
__host__ __device__ inline size_t stream_smem_bytes(int ring_mode) {
    size_t ring = ring_mode ? (size_t)SFX_NWARP * SFX_NBUF * SFX_KPAD * sizeof(T)
                            : (size_t)SFX_NWARP * SFX_KPAD * sizeof(T);
    return ring + SFX_NWARP * SFX_NBUF * sizeof(uint64_t) + SFX_NWARP * sizeof(unsigned int) +
           SFX_TL_GROUPS * sizeof(uint64_t) + 16 + 64;
}

}  // namespace sfx
#endif
