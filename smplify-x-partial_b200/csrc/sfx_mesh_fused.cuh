// Full-mesh SMPL-X forward (body_model(return_verts=True), reference fit_single_frame.py:611;
// third-party smplx lbs.lbs) as ONE tensor-core kernel for sm_100a: both dense contractions on
// tcgen05 with TMA-fed operands, the posed offsets never leaving tensor memory.
//
//   K1  blend     vp[v][c][f]  = sum_k PK[3v+c][k] . C[f][k]          k < 512   (kind::f16 on float16 copies; kind::tf32 optional)
//   K2  skinning  T[v][f][e]   = sum_j W[v][j] . A[f][j][e]           j < 64, e < 12  (3 x f16 split)
//   epilogue      vert[f][v][r] = T[v][f][4r..4r+2] . (vt[v] + vp[v][.][f]) + T[v][f][4r+3]
//
// One CTA owns a tile of 128 vertices x 128 frames.  TMEM (512 columns x 128 lanes, lane =
// vertex): columns [0,384) the three K1 accumulators (component c at 128 c, column = frame),
// columns [384,480) the K2 accumulator of one chunk of 8 frames (column = 12 f + e).
//
//   warp 8 / lane 0   TMA producer.  K1: 16 k-blocks of {3 x [128 v][32 k] of PK through a 3-D
//                     tensor map over PK viewed as [V][3][512], [128 f][32 k] of C}, 3-stage ring.
//                     K2 (re-using the ring's memory once K1's MMAs have drained): the two W tiles
//                     (hi / lo, float16: 64 joints = one 128-byte swizzle row) once, then per
//                     chunk the hi / lo tiles of A^T for 8 frames.
//   warp 9 / lane 0   MMA issuer.  K1: 4 x tcgen05.mma m128 n128 k8 (kind::tf32) per component
//                     and k-block; K2 per chunk: 4 k-steps x {Whi.Ahi, Wlo.Ahi, Whi.Alo} m128 n96
//                     k16 (kind::f16): half the instructions of a tf32 split (k8) for the same
//                     products.
//   warps 0..7        epilogue, a thread per vertex and half chunk (warps w and w + 4 share a TMEM lane
//                     quadrant and take 4 frames each; 16 warps measured no faster): tcgen05.ld of T (48 columns) and of the chunk's
//                     frames of the three K1 accumulators, 4 x (3x4 transform), staged in shared
//                     memory and written as 1536-byte contiguous runs per frame (coalesced 128-byte
//                     warp stores; a TMA tensor store cannot address [B][V][3] float rows of
//                     3 * 10475 * 4 = 125 700 bytes: the row pitch is not a multiple of 16).
//
// tf32 keeps 10 mantissa bits.  K1 multiplies centimetre-sized blend offsets: error ~5e-6 m.  K2
// multiplies metre-sized transforms, so W and A are split x = hi + lo with hi = x rounded to
// float16 (11 significant bits, like tf32; weights lie in [0, 1] and transforms within a few
// metres, far inside the float16 range) and lo = the remainder rounded to float16; the products
// of float16 values are exact in the float32 accumulator and the three significant ones are
// accumulated (the lo.lo term is 2^-22 relative): measured against the float32 SIMT kernel in
// tests/test_gpu_parity.py.
//
// K1 runs on float16 copies of PK (made once per model, 32 MB) and of the coefficients (made by the
// split kernel): 8 k-blocks of 64 instead of 16 of 32, half the bytes and half the MMA
// instructions, and float16 ROUNDS to the 11 significant bits that tf32 TRUNCATES to -- 5.6e-5 m
// against 1.3e-4 m worst-case difference to the float32 SIMT kernel, 26.2 us against 33.8 us
// (SFX_MESH_K1_TF32=1 selects the tf32 variant).  Entries of PK below 6e-5 become float16
// subnormals: an absolute error of at most 3e-8 m per term.
//
// Measured and not adopted (128 frames, with K1 in tf32: whole forward 56 us, this kernel 33.8 us):
//   * frame tiles of 64 or 48 frames (164 / 246 CTAs instead of 82 on 148 SMs): 66 us -- every
//     extra frame tile streams the blend matrix again;
//   * a tiled copy of PK whose TMA boxes are 16 KB of consecutive bytes (instead of 128-byte
//     pieces 6 KB apart): no change -- the K1 phase is bound by what one SM draws (1 MB per CTA at
//     ~88 GB/s), not by the DRAM access pattern;
//   * chunks of 4 frames with two alternating K2 accumulators: 62 us -- the epilogue's chain per
//     chunk (tensor-memory load, wait, staging, barrier, copy-out) is latency, not throughput.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "sfx_mesh_tc.cuh"

namespace sfx {

constexpr int FU_TV = 128;                  // vertices per tile (UMMA M)
constexpr int FU_TF = 128;                  // frames per tile (UMMA N of K1)
constexpr int FU_CH = 8;                    // frames per K2 chunk
constexpr int FU_N2 = 12 * FU_CH;           // UMMA N of K2 (96)
constexpr int FU_KB = 32;                   // tf32 elements per k-block (one 128-byte swizzle row)
constexpr int FU_KB16 = 64;                 // float16 elements per k-block (K1 on float16 operands)
constexpr int FU_STAGES = 3;
constexpr int FU_A_BYTES = FU_TV * FU_KB * 4;                       // 16 KB: one [128][32] tile
constexpr int FU_STAGE_BYTES = 3 * FU_A_BYTES + FU_TF * FU_KB * 4;  // 64 KB
constexpr int FU_RING_BYTES = FU_STAGES * FU_STAGE_BYTES;           // 192 KB
constexpr int FU_WJ = 64;                   // padded joint count (K of K2): 64 float16 = one 128-byte swizzle row
constexpr int FU_K2 = 16;                   // float16 MMA depth (32 bytes)
constexpr int FU_W_BOX = FU_TV * FU_WJ * 2;                         // 16 KB: one [128 v][64 j] float16 tile
constexpr int FU_W_BYTES = 2 * FU_W_BOX;                            // Whi | Wlo
constexpr int FU_A2_BOX = FU_N2 * FU_WJ * 2;                        // 12 KB: one [96][64 j] float16 tile
constexpr int FU_A2_BYTES = 2 * FU_A2_BOX;                          // hi | lo: 24 KB
constexpr int FU_OUT_BYTES = FU_CH * 3 * FU_TV * 4;                 // 12 KB staging per chunk
constexpr int FU_SMEM_BYTES = FU_RING_BYTES + 2 * FU_OUT_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int FU_EPW = 2;                    // epilogue warps per TMEM lane quadrant (each takes FU_CH / FU_EPW frames of a chunk)
constexpr int FU_THREADS = 32 * (4 * FU_EPW + 2);   // epilogue warps, then the TMA warp and the MMA warp
static_assert(FU_W_BYTES + 2 * FU_A2_BYTES <= FU_RING_BYTES, "K2 operands re-use the K1 ring");

struct FusedPlan {
    CUtensorMap map_pk3;      // PK as [V][3][512] fp32, box {32, 1, 128}, SWIZZLE_128B
    CUtensorMap map_pk3h;     // the float16 copy of PK, same view, box {64, 1, 128}
    CUtensorMap map_whi, map_wlo;   // [Vpad][64] float16, box {64, 128}
    bool ready = false;
    int V = 0;
};

static std::string make_map3(CUtensorMap* map, const void* base, uint64_t V, bool f16 = false) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
            return "cuTensorMapEncodeTiled is not available from this driver";
        fn = (PFN_encodeTiled)p;
    }
    const cuuint64_t es = f16 ? 2 : 4;
    cuuint64_t dims[3] = {SFX_KPAD, 3, V};
    cuuint64_t strides[2] = {SFX_KPAD * es, 3ull * SFX_KPAD * es};
    cuuint32_t box[3] = {(cuuint32_t)(f16 ? FU_KB16 : FU_KB), 1, FU_TV};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base,
                    dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled (3-D) failed with code " + std::to_string((int)r);
    return "";
}

// [rows][64] float16, K-major: a row is one 128-byte swizzle row
static std::string make_tile_map_f16(CUtensorMap* map, const __half* base, uint64_t rows, uint32_t box_rows,
                                     uint64_t cols = 64) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
            return "cuTensorMapEncodeTiled is not available from this driver";
        fn = (PFN_encodeTiled)p;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(__half)};
    cuuint32_t box[2] = {64, box_rows};        // 64 float16 = one 128-byte swizzle row
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled (float16) failed with code " + std::to_string((int)r);
    return "";
}

// x -> (hi, lo): hi = x rounded to float16 (the 11 bits one tensor-core product keeps), lo the
// remainder, rounded to float16
__host__ __device__ inline void f16_split(float x, __half* hi, __half* lo) {
    *hi = __float2half_rn(x);
    *lo = __float2half_rn(x - __half2float(*hi));
}

// float16 copy of the blend matrix (K1 operand of the fused kernel), made once per model
__global__ void mesh_pk_half_kernel(const float* __restrict__ PK, long n, __half* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2half_rn(PK[i]);
}

static std::string fused_plan_create(FusedPlan& plan, const float* PK, __half* pk16_dev, int V, const __half* whi_dev,
                                     const __half* wlo_dev, int Vpad) {
    plan.V = V;
    const long n = 3L * V * SFX_KPAD;
    mesh_pk_half_kernel<<<(unsigned)((n + 255) / 256), 256>>>(PK, n, pk16_dev);
    cudaError_t ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) return std::string("mesh_pk_half_kernel: ") + cudaGetErrorString(ce);
    std::string e = make_map3(&plan.map_pk3, PK, (uint64_t)V);
    if (e.empty()) e = make_map3(&plan.map_pk3h, pk16_dev, (uint64_t)V, true);
    if (e.empty()) e = make_tile_map_f16(&plan.map_whi, whi_dev, Vpad, FU_TV);
    if (e.empty()) e = make_tile_map_f16(&plan.map_wlo, wlo_dev, Vpad, FU_TV);
    plan.ready = e.empty();
    return e;
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ uint32_t umma_idesc_tf32_n(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// kind::f16 with float16 operands, fp32 accumulate, A and B K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc_f16_n(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#define SFX_TMEM_LD(N, taddr, v, off)                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"   \
                 : "=r"(v[off]), "=r"(v[off + 1]), "=r"(v[off + 2]), "=r"(v[off + 3]),              \
                   "=r"(v[off + 4]), "=r"(v[off + 5]), "=r"(v[off + 6]), "=r"(v[off + 7])           \
                 : "r"(taddr)                                                                       \
                 : "memory")

// K1F16: K1 on float16 operands (the float16 copy of PK and of the coefficients, kind::f16 k16
// steps, 8 k-blocks of 64) instead of tf32 (fp32 operands, k8 steps, 16 k-blocks of 32): half the
// bytes and half the MMA instructions; float16 rounds to the 11 significant bits tf32 truncates to
template <bool K1F16>
__global__ void __launch_bounds__(FU_THREADS, 1)
mesh_fused_tc_kernel(const __grid_constant__ CUtensorMap map_pk3, const __grid_constant__ CUtensorMap map_c,
                     const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
                     const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                     const float* __restrict__ vt, float* __restrict__ verts, int B, int V) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* outbuf = smem + FU_RING_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(outbuf + 2 * FU_OUT_BYTES);
    uint64_t* empty = full + FU_STAGES;
    uint64_t* d1_full = empty + FU_STAGES;
    uint64_t* w_full = d1_full + 1;
    uint64_t* a2_full = w_full + 1;          // [2]
    uint64_t* a2_empty = a2_full + 2;        // [2]
    uint64_t* d2_full = a2_empty + 2;
    uint64_t* d2_empty = d2_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_empty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int v0 = blockIdx.x * FU_TV;       // first vertex of the tile
    const int m0 = blockIdx.y * FU_TF;       // first frame of the tile
    constexpr int KBE = K1F16 ? FU_KB16 : FU_KB;       // elements per k-block (128 bytes either way)
    constexpr int NUM_KB = SFX_KPAD / KBE;
    const int nfr = B - m0 < FU_TF ? B - m0 : FU_TF;
    const int nchunk = (nfr + FU_CH - 1) / FU_CH;

    if (threadIdx.x == 0) {
        for (int s = 0; s < FU_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(d1_full, 1);
        mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(a2_full + s, 1);
            mbar_init(a2_empty + s, 1);
        }
        mbar_init(d2_full, 1);
        mbar_init(d2_empty, 128 * FU_EPW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d2 = tmem_base + 3 * FU_TF;

    // K2 operand layout inside the (re-used) ring
    unsigned char* wbuf = smem;                          // Whi | Wlo, [128 v][64 j] float16 each
    unsigned char* a2buf = smem + FU_W_BYTES;            // two slots of {Ahi, Alo}, [96][64 j] float16 each

    if (warp == 4 * FU_EPW && lane == 0) {
        // ===== TMA producer =====
        for (int kb = 0; kb < NUM_KB; ++kb) {
            const int s = kb % FU_STAGES;
            if (kb >= FU_STAGES) mbar_wait(empty + s, ((kb / FU_STAGES) - 1) & 1);
            unsigned char* a = smem + s * FU_STAGE_BYTES;
            mbar_expect_tx(full + s, FU_STAGE_BYTES);
            for (int c = 0; c < 3; ++c) tma_load_3d(a + c * FU_A_BYTES, &map_pk3, kb * KBE, c, v0, full + s);
            tma_load_2d(a + 3 * FU_A_BYTES, &map_c, kb * KBE, m0, full + s);
        }
        // K2 operands go where the K1 stages were: wait until K1's MMAs have read them all
        mbar_wait(d1_full, 0);
        mbar_expect_tx(w_full, FU_W_BYTES);
        tma_load_2d(wbuf, &map_whi, 0, v0, w_full);
        tma_load_2d(wbuf + FU_W_BOX, &map_wlo, 0, v0, w_full);
        for (int q = 0; q < nchunk; ++q) {
            const int s = q & 1;
            if (q >= 2) mbar_wait(a2_empty + s, ((q >> 1) - 1) & 1);
            unsigned char* a = a2buf + s * FU_A2_BYTES;
            const int row0 = (m0 + q * FU_CH) * 12;
            mbar_expect_tx(a2_full + s, FU_A2_BYTES);
            tma_load_2d(a, &map_ahi, 0, row0, a2_full + s);
            tma_load_2d(a + FU_A2_BOX, &map_alo, 0, row0, a2_full + s);
        }
    } else if (warp == 4 * FU_EPW + 1 && lane == 0) {
        // ===== MMA issuer =====
        const uint32_t idesc1 = K1F16 ? umma_idesc_f16_n(FU_TF) : umma_idesc_tf32_n(FU_TF);
        for (int kb = 0; kb < NUM_KB; ++kb) {
            const int s = kb % FU_STAGES;
            mbar_wait(full + s, (kb / FU_STAGES) & 1);
            tc_fence_after();
            unsigned char* a = smem + s * FU_STAGE_BYTES;
            const uint64_t db = umma_desc_sw128(a + 3 * FU_A_BYTES);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const uint64_t da = umma_desc_sw128(a + c * FU_A_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k) {           // four 32-byte k-steps per 128-byte row
                    if (K1F16) umma_f16(tmem_base + c * FU_TF, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0);
                    else umma_tf32(tmem_base + c * FU_TF, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0);
                }
            }
            tc_commit(empty + s);
        }
        tc_commit(d1_full);                  // K1 accumulators complete, K1 operands consumed
        const uint32_t idesc2 = umma_idesc_f16_n(FU_N2);
        const uint64_t whi = umma_desc_sw128(wbuf), wlo = umma_desc_sw128(wbuf + FU_W_BOX);
        mbar_wait(w_full, 0);
        for (int q = 0; q < nchunk; ++q) {
            const int s = q & 1;
            mbar_wait(a2_full + s, (q >> 1) & 1);
            if (q >= 1) mbar_wait(d2_empty, (q - 1) & 1);      // the epilogue has read the previous chunk
            tc_fence_after();
            unsigned char* a = a2buf + s * FU_A2_BYTES;
            const uint64_t ahi = umma_desc_sw128(a), alo = umma_desc_sw128(a + FU_A2_BOX);
            // 55 joints: the k-step of joints 48 .. 63 still holds seven of them
#pragma unroll
            for (int k = 0; k < FU_WJ / FU_K2; ++k) {
                umma_f16(tmem_d2, whi + 2 * k, ahi + 2 * k, idesc2, k == 0 ? 0u : 1u);
                umma_f16(tmem_d2, wlo + 2 * k, ahi + 2 * k, idesc2, 1u);
                umma_f16(tmem_d2, whi + 2 * k, alo + 2 * k, idesc2, 1u);
            }
            tc_commit(a2_empty + s);
            tc_commit(d2_full);
        }
    } else if (warp < 4 * FU_EPW) {
        // ===== epilogue: thread = vertex (TMEM lane 32 (warp % 4) + lane) x part of the chunk's frames =====
        const int quad = warp & 3, half = warp >> 2;       // `half`: which part of the chunk (0 .. FU_EPW-1)
        constexpr int HF = FU_CH / FU_EPW;                 // frames per thread and chunk
        static_assert(FU_CH % FU_EPW == 0 && (12 * HF) % 8 == 0, "whole 8-column loads per thread");
        const int vloc = quad * 32 + lane, v = v0 + vloc;
        float t0 = 0, t1 = 0, t2 = 0;
        if (v < V) { t0 = vt[3 * v]; t1 = vt[3 * v + 1]; t2 = vt[3 * v + 2]; }
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const int tile_floats = (V - v0 < FU_TV ? V - v0 : FU_TV) * 3;     // valid floats of a frame run
        mbar_wait(d1_full, 0);
        tc_fence_after();
        for (int q = 0; q < nchunk; ++q) {
            // the chunk's frames of the K1 accumulators do not depend on K2: their loads are issued
            // before the wait, so only the read of T stands between two chunks' MMAs
            uint32_t T[12 * HF], P[3 * FU_CH];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                SFX_TMEM_LD(8, tmem_base + lane_addr + c * FU_TF + q * FU_CH, P, 8 * c);
            mbar_wait(d2_full, q & 1);
            tc_fence_after();
#pragma unroll
            for (int c8 = 0; c8 < 12 * HF / 8; ++c8)
                SFX_TMEM_LD(8, tmem_d2 + lane_addr + 12 * HF * half + 8 * c8, T, 8 * c8);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(d2_empty);                         // D2 may be overwritten by the next chunk
            float* ob = reinterpret_cast<float*>(outbuf + (q & 1) * FU_OUT_BYTES);
#pragma unroll
            for (int fh = 0; fh < HF; ++fh) {
                const int f = half * HF + fh;
                // (selects over compile-time indices: a register array cannot be indexed by `half`)
                float px = 0, py = 0, pz = 0;
#pragma unroll
                for (int h2 = 0; h2 < FU_EPW; ++h2)
                    if (half == h2) {
                        px = __uint_as_float(P[h2 * HF + fh]);
                        py = __uint_as_float(P[8 + h2 * HF + fh]);
                        pz = __uint_as_float(P[16 + h2 * HF + fh]);
                    }
                const float x = t0 + px, y = t1 + py, z = t2 + pz;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float* Tr = reinterpret_cast<const float*>(T) + 12 * fh + 4 * r;
                    ob[f * 3 * FU_TV + 3 * vloc + r] = Tr[0] * x + Tr[1] * y + Tr[2] * z + Tr[3];
                }
            }
            // the 128 threads of this part wrote its frames: named barrier 1 + part
            asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
            // coalesced copy-out: per frame one contiguous run of up to 384 floats
#pragma unroll
            for (int fh = 0; fh < HF; ++fh) {
                const int f = half * HF + fh;
                const int fr = m0 + q * FU_CH + f;
                if (fr < B) {
                    float* dst = verts + ((size_t)fr * V + v0) * 3;
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int e = u * FU_TV + vloc;
                        if (e < tile_floats) dst[e] = ob[f * 3 * FU_TV + e];
                    }
                }
            }
            // the staging buffer is double buffered; the barrier of chunk q + 1 orders re-use
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512)
                     : "memory");
    }
}

// A^T tiles of K2: per frame 12 rows x 64 joints (hi and lo), row 12 f + e, from the skinning
// transforms A [B][55][12] of the pose prologue; behind them (c16 != nullptr) the float16 copy of the
// blend coefficients C [Bpad][512] for K1
__global__ void mesh_at_split_kernel(const float* __restrict__ A, int B, int Bpad, __half* __restrict__ ahi,
                                     __half* __restrict__ alo, const float* __restrict__ C,
                                     __half* __restrict__ c16) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_split = Bpad * 12 * FU_WJ;
    if (idx < n_split) {
        const int j = idx % FU_WJ, row = idx / FU_WJ, f = row / 12, e = row % 12;
        float x = 0.f;
        if (f < B && j < SFX_NJ) x = A[((size_t)f * SFX_NJ + j) * 12 + e];
        f16_split(x, ahi + idx, alo + idx);
        return;
    }
    idx -= n_split;
    if (c16 && idx < Bpad * SFX_KPAD) c16[idx] = __float2half_rn(idx / SFX_KPAD < B ? C[idx] : 0.f);
}

// k1_f16: K1 on the float16 copies (default); false: K1 in tf32 on the fp32 arrays (SFX_MESH_K1_TF32=1)
static std::string mesh_fused_tc(const FusedPlan& plan, int B, const float* C, const float* A, __half* ahi,
                                 __half* alo, __half* c16, bool k1_f16, const float* vt, float* verts,
                                 cudaStream_t s) {
    if (!plan.ready) return "fused tensor-core mesh plan was not created";
    const int Bpad = (B + FU_TF - 1) / FU_TF * FU_TF;
    CUtensorMap map_c, map_ahi, map_alo;
    std::string e = k1_f16 ? make_tile_map_f16(&map_c, c16, Bpad, FU_TF, SFX_KPAD)
                           : make_tile_map(&map_c, C, Bpad, SFX_KPAD, FU_TF, FU_KB);
    if (e.empty()) e = make_tile_map_f16(&map_ahi, ahi, (uint64_t)Bpad * 12, FU_N2);
    if (e.empty()) e = make_tile_map_f16(&map_alo, alo, (uint64_t)Bpad * 12, FU_N2);
    if (!e.empty()) return e;
    const int n = Bpad * 12 * FU_WJ + (k1_f16 ? Bpad * SFX_KPAD : 0);
    mesh_at_split_kernel<<<(n + 255) / 256, 256, 0, s>>>(A, B, Bpad, ahi, alo, C, k1_f16 ? c16 : nullptr);
    auto kern = k1_f16 ? mesh_fused_tc_kernel<true> : mesh_fused_tc_kernel<false>;
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_SMEM_BYTES);
    if (ce != cudaSuccess) return std::string("mesh_fused_tc_kernel attr: ") + cudaGetErrorString(ce);
    dim3 grid((plan.V + FU_TV - 1) / FU_TV, Bpad / FU_TF);
    kern<<<grid, FU_THREADS, FU_SMEM_BYTES, s>>>(k1_f16 ? plan.map_pk3h : plan.map_pk3, map_c, plan.map_whi,
                                                 plan.map_wlo, map_ahi, map_alo, vt, verts, B, plan.V);
    ce = cudaGetLastError();
    return ce == cudaSuccess ? "" : std::string("mesh_fused_tc_kernel: ") + cudaGetErrorString(ce);
}

}  // namespace sfx
