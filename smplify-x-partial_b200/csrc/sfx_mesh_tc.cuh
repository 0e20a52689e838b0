// Full-mesh blendshape contraction on the 5th-generation tensor cores (sm_100a):
//
//     v_posed[b][r] = vt[r] + sum_k C[b][k] * PK[r][k]        b < B (frames), r < 3V, k < 512
//
// i.e. D[M = frames][N = mesh rows] = A[M][K] . B[N][K]^T with both operands K-major exactly as
// they lie in HBM (C[B][512] coefficient rows from the pose prologue, PK[3V][512] blend matrix),
// kind::tf32, fp32 accumulation in TMEM.  One CTA per (128 mesh rows) x (128 frames) tile:
//
//   warp 0 / lane 0   TMA producer: cp.async.bulk.tensor.2d of A and B k-blocks (128 rows x 32
//                     tf32 = 128 B, SWIZZLE_128B) into a 4-stage shared-memory ring
//   warp 1 / lane 0   MMA issuer: 4 x tcgen05.mma (UMMA 128x128x8) per k-block, tcgen05.commit
//                     releases the stage back to the producer and finally signals the epilogue
//   warps 0..3        epilogue: tcgen05.ld 32 lanes x 32 columns at a time, + template vertex,
//                     128-byte row segments to v_posed[b][r0 .. r0+31]
//
// The kernel is HBM-bound: it streams the 64 MB blend matrix once per 128 frames (the A tiles,
// 256 KB per frame tile, are re-read from L2 by every CTA).  tf32 keeps 10 mantissa bits of each
// input: the worst-case vertex error is ~2e-5 m (shape displacement of centimetres x 2^-11), see
// DESIGN.md.  The float64 build and the parity checks use the SIMT kernel in sfx_mesh.cuh.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "sfx_core.cuh"
#include "sfx_stream.cuh"

namespace sfx {

constexpr int TC_BLOCK_M = 128;      // frames per tile (UMMA M)
constexpr int TC_BLOCK_N = 128;      // mesh rows per tile (UMMA N)
constexpr int TC_BLOCK_K = 32;       // tf32 elements per k-block = one 128-byte swizzle row
constexpr int TC_UMMA_K = 8;         // tf32 MMA depth (32 bytes)
constexpr int TC_STAGES = 4;
constexpr int TC_STAGE_BYTES = (TC_BLOCK_M + TC_BLOCK_N) * TC_BLOCK_K * 4;     // 32 KB
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_TMEM_COLS = 128;

struct MeshPlan {
    CUtensorMap map_pk;       // [3V][512] fp32, box 32 x 128, SWIZZLE_128B
    bool ready = false;
    int V = 0;
};

// ---- host: tensor maps through the driver entry point (no link-time libcuda dependency) ------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static std::string make_tile_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols,
                                 uint32_t box_rows, uint32_t box_cols) {
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
    cuuint64_t strides[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r);
    return "";
}

static std::string mesh_plan_create(MeshPlan& plan, const float* PK, int V) {
    plan.V = V;
    std::string e = make_tile_map(&plan.map_pk, PK, 3ull * V, SFX_KPAD, TC_BLOCK_N, TC_BLOCK_K);
    plan.ready = e.empty();
    return e;
}

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// K-major operand tile of 128-byte rows, SWIZZLE_128B: 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);          // start address       [0,14)
    d |= (uint64_t)1 << 16;                                     // leading byte offset [16,30) (unused)
    d |= (uint64_t)(1024 >> 4) << 32;                           // stride byte offset  [32,46)
    d |= (uint64_t)1 << 46;                                     // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                     // SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = TC_BLOCK_N
__device__ __forceinline__ uint32_t umma_idesc_tf32() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BLOCK_N >> 3) << 17) |
           ((uint32_t)(TC_BLOCK_M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128, 1)
mesh_blend_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const float* __restrict__ vt, float* __restrict__ vposed, int B, int nrows) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tmem_full = empty + TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * TC_BLOCK_N;      // first mesh row of the tile
    const int m0 = blockIdx.y * TC_BLOCK_M;      // first frame of the tile
    constexpr int NUM_KB = SFX_KPAD / TC_BLOCK_K;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "n"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        for (int kb = 0; kb < NUM_KB; ++kb) {
            const int s = kb % TC_STAGES;
            if (kb >= TC_STAGES) mbar_wait(empty + s, ((kb / TC_STAGES) - 1) & 1);
            unsigned char* a = smem + s * TC_STAGE_BYTES;
            unsigned char* b = a + TC_BLOCK_M * TC_BLOCK_K * 4;
            mbar_expect_tx(full + s, TC_STAGE_BYTES);
            tma_load_2d(a, &map_a, kb * TC_BLOCK_K, m0, full + s);
            tma_load_2d(b, &map_b, kb * TC_BLOCK_K, n0, full + s);
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        const uint32_t idesc = umma_idesc_tf32();
        for (int kb = 0; kb < NUM_KB; ++kb) {
            const int s = kb % TC_STAGES;
            mbar_wait(full + s, (kb / TC_STAGES) & 1);
            tc_fence_after();
            unsigned char* a = smem + s * TC_STAGE_BYTES;
            unsigned char* b = a + TC_BLOCK_M * TC_BLOCK_K * 4;
            const uint64_t da = umma_desc_sw128(a), db = umma_desc_sw128(b);
#pragma unroll
            for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                // advance inside the 128-byte swizzle row: 32 bytes = 2 sixteen-byte units
                umma_tf32(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            }
            tc_commit(empty + s);               // frees the stage once these MMAs have read it
        }
        tc_commit(tmem_full);                   // accumulator complete
    }
    __syncwarp();
    // ===== epilogue: all four warps, warp w owns TMEM lanes 32w .. 32w+31 (frames) =====
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int frame = m0 + warp * 32 + lane;
    float* out = vposed + (size_t)frame * nrows;
#pragma unroll 1
    for (int c = 0; c < TC_BLOCK_N; c += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
              "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
              "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
              "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
              "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (frame < B) {
            const int r0 = n0 + c;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int r = r0 + j;
                if (r < nrows) out[r] = __uint_as_float(v[j]) + __ldg(vt + r);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "n"(TC_TMEM_COLS)
                     : "memory");
    }
}

// C [B_pad][512] changes address per batch: its tensor map is rebuilt per call (cheap, host side)
static std::string mesh_blend_tc(const MeshPlan& plan, int B, const float* C, const float* vt,
                                 float* vposed, cudaStream_t s) {
    if (!plan.ready) return "tensor-core mesh plan was not created";
    const int Bpad = (B + TC_BLOCK_M - 1) / TC_BLOCK_M * TC_BLOCK_M;
    CUtensorMap map_a;
    std::string e = make_tile_map(&map_a, C, Bpad, SFX_KPAD, TC_BLOCK_M, TC_BLOCK_K);
    if (!e.empty()) return e;
    const int nrows = 3 * plan.V;
    cudaError_t ce = cudaFuncSetAttribute(mesh_blend_tc_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (ce != cudaSuccess) return std::string("mesh_blend_tc_kernel attr: ") + cudaGetErrorString(ce);
    dim3 grid((nrows + TC_BLOCK_N - 1) / TC_BLOCK_N, Bpad / TC_BLOCK_M);
    mesh_blend_tc_kernel<<<grid, 128, TC_SMEM_BYTES, s>>>(map_a, plan.map_pk, vt, vposed, B, nrows);
    ce = cudaGetLastError();
    return ce == cudaSuccess ? "" : std::string("mesh_blend_tc_kernel: ") + cudaGetErrorString(ce);
}

}  // namespace sfx
