// TMA-fed tcgen05 / TMEM GEMMs on pre-split bf16 operands (sm_100a).
//
// The fp32-parity arithmetic of this code base is "bf16x3": x ~= hi + lo (two bf16), products hi*hi + hi*lo + lo*hi with
// fp32 accumulation.  Converting fp32 operands inside every GEMM costs instruction issue and keeps the loads on the LSU
// path, so here the PRODUCER of an activation writes it once as two bf16 planes (hi, lo: the same 4 bytes per element as
// fp32) and every consumer GEMM streams those planes straight into shared memory with TMA (cp.async.bulk.tensor,
// 128-byte swizzle) and feeds them to tcgen05.mma without touching a register:
//
//   warp 0   TMA producer   : waits "slot empty", arms the slot's mbarrier with the byte count, issues the tile loads
//   warp 1   MMA issuer     : waits "slot full", one elected thread issues tcgen05.mma (M = 128, N = BN, K = 16) for the
//                             split products, tcgen05.commit -> "slot empty"; after the last k-block -> "accumulator full"
//   warps 2-9 epilogue      : wait "accumulator full", tcgen05.ld the fp32 tile out of TMEM (thread = row), bias /
//                             activation / residual, store fp32 and/or hi-lo planes for the next GEMM, -> "accumulator empty"
//   accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1; CTAs are persistent.
//
//   gemm_tma_nt : C[M,N] = act(A[M,K] W[N,K]^T + bias) (+ C)         both operands K-major (K contiguous)
//   gemm_tma_tn : dW[Mo,No] += sum_p A[p,Mo]^T B[p,No]               both operands MN-major (positions strided): the
//                 weight gradients; the position range is split over CTAs, fp32 atomics combine the partial sums.
//
// Shared-memory operand tiles use the canonical UMMA SWIZZLE_128B layouts (cute/atom/mma_traits_sm100.hpp):
//   K-major : rows of 64 bf16 (128 B), 8-row atoms of 1024 B, SBO = 1024 B, k-step of 16 elements = +32 B on the start address
//   MN-major: K-rows of 64 MN-elements (128 B), 8-K-row atoms of 1024 B (SBO = 1024 B), next 64 MN-elements at LBO,
//             k-step of 16 = +2048 B on the start address
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "kernels.h"
#include "tc5_common.cuh"

namespace dp {
namespace {

constexpr int BM = 128, BK = 64;

__device__ __forceinline__ void sts128f(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
constexpr int A_TILE = BM * BK * 2;  // 16 KB

template <int BN, bool SPLIT>
struct NtCfg {
    static constexpr int W_TILE = BN * BK * 2;
    static constexpr int STAGE = (A_TILE + W_TILE) * (SPLIT ? 2 : 1);
    // [128 rows x 32 fp32] swizzled staging tiles for the TMA stores: one per warp set, two where two operand stages still fit beside
    // them (the store of chunk i then drains while chunk i + 1 is converted; with one tile the epilogue waited ~40 % of its time there)
    static constexpr int OUT_TILE = 128 * 128;
    static constexpr int OUT_BUFS = 2 * STAGE + 4 * OUT_TILE + 1280 <= 227 * 1024 ? 2 : 1;
    static constexpr int OUT_STAGE = 2 * OUT_BUFS * OUT_TILE;
    static constexpr int STAGES = (227 * 1024 - 1280 - OUT_STAGE) / STAGE > 6 ? 6 : (227 * 1024 - 1280 - OUT_STAGE) / STAGE;
    static constexpr int SMEM = STAGES * STAGE + OUT_STAGE + 1024 /*alignment slack*/ + 256 /*barriers*/;
    static_assert(STAGES >= 2 && SMEM <= 227 * 1024, "operand ring and staging tiles must fit");
    static constexpr uint32_t TMEM_COLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
};

template <int BN, bool SPLIT, bool DROP>
__global__ void __launch_bounds__(320, 1)
gemm_tma_nt_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                   const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmCh, const __grid_constant__ CUtensorMap tmCl,
                   const TmaGemmArgs p, const int tma_store, const int tma_planes) {
    using Cfg = NtCfg<BN, SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* out_stage = smem + Cfg::STAGES * Cfg::STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(out_stage + Cfg::OUT_STAGE);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tfull = empty + Cfg::STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        prefetch_tmap(&tmAh); prefetch_tmap(&tmWh);
        if (SPLIT) { prefetch_tmap(&tmAl); prefetch_tmap(&tmWl); }
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int n_tiles = p.N / BN, m_tiles = ceil_div(p.M, BM), tiles = m_tiles * n_tiles, nkb = p.K / BK;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* st = smem + stage * Cfg::STAGE;
                    mbar_expect_tx(full + stage, Cfg::STAGE);
                    tma_load_2d(st, &tmAh, full + stage, kb * BK, m0);
                    if (SPLIT) tma_load_2d(st + A_TILE, &tmAl, full + stage, kb * BK, m0);
                    uint8_t* sw = st + A_TILE * (SPLIT ? 2 : 1);
                    tma_load_2d(sw, &tmWh, full + stage, kb * BK, n0);
                    if (SPLIT) tma_load_2d(sw + Cfg::W_TILE, &tmWl, full + stage, kb * BK, n0);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC = idesc_bf16(BM, BN, 0, 0);
        uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            mbar_wait(tempty + as, aphase ^ 1);  // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t d = tmem + as * BN;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full + stage, phase);
                tc_fence_after();
                {   // whole warp, uniform operands; one elected lane issues (see umma_w)
                    const uint32_t a_hi = smem_u32(smem + stage * Cfg::STAGE);
                    const uint32_t a_lo = a_hi + A_TILE;
                    const uint32_t w_hi = a_hi + A_TILE * (SPLIT ? 2 : 1);
                    const uint32_t w_lo = w_hi + Cfg::W_TILE;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ah = desc_sw128(a_hi + k * 32, 16, 1024), wh = desc_sw128(w_hi + k * 32, 16, 1024);
                        umma_w(d, ah, wh, IDESC, (kb | k) != 0);
                        if (SPLIT) {
                            const uint64_t al = desc_sw128(a_lo + k * 32, 16, 1024), wl = desc_sw128(w_lo + k * 32, 16, 1024);
                            umma_w(d, ah, wl, IDESC, 1);
                            umma_w(d, al, wh, IDESC, 1);
                        }
                    }
                    umma_commit_w(empty + stage);                    // smem slot reusable once these MMAs retire
                    if (kb == nkb - 1) umma_commit_w(tfull + as);    // accumulator complete
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        // two warps per TMEM lane quarter; they take alternate 32-column chunks so twice as many stores are in flight
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;
        uint32_t as = 0, aphase = 0;
        int ob = 0;   // staging tile of this warp set in use
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
            mbar_wait(tfull + as, aphase);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            const bool ok = row < p.M;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int c0 = half * 32; c0 < BN; c0 += 64, ob ^= (Cfg::OUT_BUFS - 1)) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + as * BN + c0, v);
                if (ok) {
                    const int col = n0 + c0;
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = *reinterpret_cast<const float4*>(p.bias + col + j);
                            v[j] = fmaf(p.bias_scale, b.x, v[j]); v[j + 1] = fmaf(p.bias_scale, b.y, v[j + 1]);
                            v[j + 2] = fmaf(p.bias_scale, b.z, v[j + 2]); v[j + 3] = fmaf(p.bias_scale, b.w, v[j + 3]);
                        }
                    }
                    float* dst = p.C ? p.C + (size_t)row * p.ldc + col : nullptr;
                    if (p.accumulate && !(DROP && p.drop_thr)) {
                        const float* rsrc = p.Cin ? p.Cin + (size_t)row * p.ldc + col : dst;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 o = *reinterpret_cast<const float4*>(rsrc + j);
                            v[j] += o.x; v[j + 1] += o.y; v[j + 2] += o.z; v[j + 3] += o.w;
                        }
                    }
                    if (p.act == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                    } else if (p.act == 2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
                    } else if (p.act == 3) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 1.0f / (1.0f + expf(-v[j]));
                    }
                    if (DROP && p.drop_thr) {  // dropout on the sub-layer output, then (if any) the residual
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            v[j] = drop_keep(p.drop_key, (uint32_t)row, (uint32_t)(col + j), p.drop_thr) ? v[j] * p.drop_scale : 0.f;
                        if (p.accumulate) {
                            const float* rsrc = p.Cin ? p.Cin + (size_t)row * p.ldc + col : dst;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 o = *reinterpret_cast<const float4*>(rsrc + j);
                                v[j] += o.x; v[j + 1] += o.y; v[j + 2] += o.z; v[j + 3] += o.w;
                            }
                        }
                    }
                    if (p.mul_c) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 o = *reinterpret_cast<const float4*>(dst + j);
                            v[j] *= o.x; v[j + 1] *= o.y; v[j + 2] *= o.z; v[j + 3] *= o.w;
                        }
                    }
                    if (p.mask) {
                        const float* mk = p.mask + (size_t)row * p.ldmask + col;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 o = *reinterpret_cast<const float4*>(mk + j);
                            v[j] = o.x > 0.f ? v[j] : 0.f; v[j + 1] = o.y > 0.f ? v[j + 1] : 0.f;
                            v[j + 2] = o.z > 0.f ? v[j + 2] : 0.f; v[j + 3] = o.w > 0.f ? v[j + 3] : 0.f;
                        }
                    }
                    if (p.mask_hi) {
                        const uint4* mk = reinterpret_cast<const uint4*>(p.mask_hi + (size_t)row * p.ldmask_hi + col);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 o = mk[j];
                            const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                            for (int t = 0; t < 4; ++t) {   // bf16 > 0  <=>  sign bit clear and not (+-)zero
                                const uint32_t lo16 = w[t] & 0xffffu, hi16 = w[t] >> 16;
                                if (lo16 == 0u || lo16 >= 0x8000u) v[8 * j + 2 * t] = 0.f;
                                if (hi16 == 0u || hi16 >= 0x8000u) v[8 * j + 2 * t + 1] = 0.f;
                            }
                        }
                    }
                    if (p.out_scale != 0.f) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
                    }
                    if (dst && !tma_store) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                    if (p.C_hi && !tma_planes) {
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) split_pair(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
                        uint4* dh = reinterpret_cast<uint4*>(p.C_hi + (size_t)row * p.ldch + col);
#pragma unroll
                        for (int j = 0; j < 4; ++j) dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        if (p.C_lo) {
                            uint4* dl = reinterpret_cast<uint4*>(p.C_lo + (size_t)row * p.ldch + col);
#pragma unroll
                            for (int j = 0; j < 4; ++j) dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                        }
                    }
                    if (p.stats) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
                    }
                }
                if (tma_store) {
                    // fp32 output through shared memory and a TMA store: full 128-byte rows leave the SM as bulk writes instead
                    // of 16-byte pieces of 32 different lines per store instruction.  One staging tile per warp set (4 warps).
                    const uint32_t stg = smem_u32(out_stage) + (half * Cfg::OUT_BUFS + ob) * Cfg::OUT_TILE;
                    const int bar_id = 1 + half;
                    if ((warp & 3) == 2 && lane == 0) {   // the store that last read this staging tile has drained
                        if (Cfg::OUT_BUFS == 2) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
                    }
                    asm volatile("bar.sync %0, 128;\n" ::"r"(bar_id) : "memory");
                    const int rl = q * 32 + lane;
#pragma unroll
                    for (int j = 0; j < 8; ++j) sts128f(stg + rl * 128 + ((j ^ (rl & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    proxy_fence_async();
                    asm volatile("bar.sync %0, 128;\n" ::"r"(bar_id) : "memory");
                    if ((warp & 3) == 2 && lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(&tmC),
                                     "r"(stg), "r"(n0 + c0), "r"(m0)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                    }
                }
                if (tma_planes) {
                    // plane outputs through shared memory and TMA stores as well: [128 rows x 32 bf16] tiles (64-byte rows, 64-byte
                    // swizzle), hi at +0 and lo at +8 KB of this warp set's staging area (never used together with the fp32 staging)
                    const uint32_t stg = smem_u32(out_stage) + (half * Cfg::OUT_BUFS + ob) * Cfg::OUT_TILE;
                    const int bar_id = 1 + half;
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) split_pair(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
                    if ((warp & 3) == 2 && lane == 0) {   // the stores that last read this staging tile have drained
                        if (Cfg::OUT_BUFS == 2) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
                    }
                    asm volatile("bar.sync %0, 128;\n" ::"r"(bar_id) : "memory");
                    const int rl = q * 32 + lane, sw = (rl >> 1) & 3;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        sts128u(stg + rl * 64 + ((j ^ sw) << 4), hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        if (p.C_lo) sts128u(stg + 8192 + rl * 64 + ((j ^ sw) << 4), lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                    }
                    proxy_fence_async();
                    asm volatile("bar.sync %0, 128;\n" ::"r"(bar_id) : "memory");
                    if ((warp & 3) == 2 && lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(&tmCh),
                                     "r"(stg), "r"(n0 + c0), "r"(m0)
                                     : "memory");
                        if (p.C_lo)
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(&tmCl),
                                         "r"(stg + 8192), "r"(n0 + c0), "r"(m0)
                                         : "memory");
                        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + as);  // this warp's quarter of the accumulator is drained
            if (p.stats) {
                const int g_lo = (m0 + q * 32) / p.rows_per_group;
                const int g_hi = (min(m0 + q * 32 + 31, p.M - 1)) / p.rows_per_group;
                if (g_lo == g_hi) {
                    double a = warp_sum_d((double)s1), b = warp_sum_d((double)s2);
                    if (lane == 0 && m0 + q * 32 < p.M) { atomicAdd(p.stats + 2 * g_lo, a); atomicAdd(p.stats + 2 * g_lo + 1, b); }
                } else if (ok) {
                    const int g = row / p.rows_per_group;
                    atomicAdd(p.stats + 2 * g, (double)s1);
                    atomicAdd(p.stats + 2 * g + 1, (double)s2);
                }
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        if ((tma_store || tma_planes) && (warp & 3) == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// 2-D bf16 tensor [rows, inner] with row stride ld (elements); box = [box_rows, 64 inner elements], 128-byte swizzle
bool make_map(CUtensorMap* map, const void* base, long long rows, long long inner, long long ld, int box_rows) {
    auto fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// fp32 output [M, N] (row stride ldc): box = 128 rows x 32 columns (128 bytes), 128-byte swizzle
bool make_map_c(CUtensorMap* map, const void* base, long long M, long long N, long long ldc) {
    auto fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t gstr[1] = {(cuuint64_t)ldc * 4};
    cuuint32_t box[2] = {32u, 128u};
    cuuint32_t estr[2] = {1u, 1u};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// bf16 plane output [M, N] (row stride ldch): box = 128 rows x 32 columns (64 bytes), 64-byte swizzle
bool make_map_p(CUtensorMap* map, const void* base, long long M, long long N, long long ldch) {
    auto fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t gstr[1] = {(cuuint64_t)ldch * 2};
    cuuint32_t box[2] = {32u, 128u};
    cuuint32_t estr[2] = {1u, 1u};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, bool SPLIT>
cudaError_t launch_nt(const TmaGemmArgs& a, cudaStream_t st) {
    using Cfg = NtCfg<BN, SPLIT>;
    CUtensorMap mAh, mAl, mWh, mWl, mC, mCh, mCl;
    if (!make_map(&mAh, a.A_hi, a.M, a.K, a.lda, BM) || !make_map(&mWh, a.W_hi, a.N, a.K, a.ldw, BN)) return cudaErrorInvalidValue;
    if (SPLIT) {
        if (!make_map(&mAl, a.A_lo, a.M, a.K, a.lda, BM) || !make_map(&mWl, a.W_lo, a.N, a.K, a.ldw, BN)) return cudaErrorInvalidValue;
    } else {
        mAl = mAh; mWl = mWh;
    }
    // plain fp32 outputs leave through TMA stores; read-modify-write epilogues keep the direct path
    int tma_store = a.C != nullptr && !a.accumulate && !a.mul_c && !a.mask && (a.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(a.C) & 15) == 0;
    mC = mAh;
    if (tma_store && !make_map_c(&mC, a.C, a.M, a.N, a.ldc)) tma_store = 0;
    // plane outputs share the staging area with the fp32 stores: staged when the fp32 output is absent or takes the direct path
    int tma_planes = a.C_hi != nullptr && !tma_store && (a.ldch & 7) == 0 && (reinterpret_cast<uintptr_t>(a.C_hi) & 15) == 0 &&
                     (a.C_lo == nullptr || (reinterpret_cast<uintptr_t>(a.C_lo) & 15) == 0);
    mCh = mAh; mCl = mAh;
    if (tma_planes && !make_map_p(&mCh, a.C_hi, a.M, a.N, a.ldch)) tma_planes = 0;
    if (tma_planes && a.C_lo && !make_map_p(&mCl, a.C_lo, a.M, a.N, a.ldch)) tma_planes = 0;
    const int tiles = ceil_div(a.M, BM) * (a.N / BN);
    const int grid = tiles < 148 ? tiles : 148;
    cudaError_t e;
    if (a.drop_thr) {   // the dropout epilogue is a separate instantiation: the plain one keeps its register budget
        e = cudaFuncSetAttribute(gemm_tma_nt_kernel<BN, SPLIT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        gemm_tma_nt_kernel<BN, SPLIT, true><<<grid, 320, Cfg::SMEM, st>>>(mAh, mAl, mWh, mWl, mC, mCh, mCl, a, tma_store, tma_planes);
    } else {
        e = cudaFuncSetAttribute(gemm_tma_nt_kernel<BN, SPLIT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        gemm_tma_nt_kernel<BN, SPLIT, false><<<grid, 320, Cfg::SMEM, st>>>(mAh, mAl, mWh, mWl, mC, mCh, mCl, a, tma_store, tma_planes);
    }
    return cudaGetLastError();
}

}  // namespace

bool gemm_tma_nt_supported(const TmaGemmArgs& a) {
    if (a.M <= 0 || a.K <= 0 || a.K % BK || a.N % 64) return false;
    if ((a.lda & 7) || (a.ldw & 7)) return false;
    if (a.C && (a.ldc & 3)) return false;
    if (a.C_hi && (a.ldch & 7)) return false;
    if (a.mask_hi && (a.ldmask_hi & 7)) return false;
    if ((a.accumulate || a.mul_c) && !a.C) return false;
    if (!a.C && !a.C_hi) return false;
    return encode_fn() != nullptr;
}

cudaError_t launch_gemm_tma_nt(const TmaGemmArgs& a, bool split, cudaStream_t st) {
    if (!gemm_tma_nt_supported(a)) return cudaErrorInvalidValue;
    if (split && (!a.A_lo || !a.W_lo)) return cudaErrorInvalidValue;
#define DP_NT(BNV) (split ? launch_nt<BNV, true>(a, st) : launch_nt<BNV, false>(a, st))
    if (a.N % 256 == 0) return DP_NT(256);
    if (a.N % 128 == 0) return DP_NT(128);
    return DP_NT(64);
#undef DP_NT
}

// (weight-gradient kernel)
namespace {
// ================================================================================================
// Weight gradients: C_s[Mo, nb_s] += scale * sum_p A[p, Mo]^T B_s[p, nb_s] for up to two B segments that share the A
// operand (e.g. dW_ih = dG^T X and dW_hh = dG^T h_prev in ONE pass over dG).  A and B are [P, cols] planes, i.e. MN-major
// operands (the contraction index p is the slow one).  One CTA = one 128-row slice of A^T x all N = nb0 + nb1 columns x a
// contiguous range of positions; accumulators stay in TMEM for the CTA's whole range, fp32 atomics combine the ranges.
constexpr int KP = 32;                      // positions per pipeline stage (2 MMAs of K = 16)
constexpr int BOX = KP * 128;               // one TMA box: KP rows of 64 bf16 = 4 KB (4 swizzle atoms of 8 rows)

// Cluster variant (MC): the four CTAs of a cluster take the four 128-row slices of the output for the SAME position range, so they need the same
// B tiles: every CTA loads a quarter of the stage's B boxes and multicasts them to all four (cp.async.bulk.tensor ... .multicast::cluster); a
// stage is free again when all four CTAs have consumed it (tcgen05.commit with a cluster multicast onto every CTA's `empty` barrier).  Without
// it every B tile crosses L2 -> SM four times and the kernel runs into the L2 bandwidth (measured 7.5 TB/s L2 -> SM at 88 us per launch).
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc_w(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "{\n .reg .pred q;\n elect.sync _|q, 0xffffffff;\n"
        " @q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}\n" ::"r"(smem_u32(bar)),
        "h"(mask)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

template <bool SPLIT, bool MC>
__global__ void __launch_bounds__(192, 1)
gemm_tma_tn_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmB0h, const __grid_constant__ CUtensorMap tmB0l,
                   const __grid_constant__ CUtensorMap tmB1h, const __grid_constant__ CUtensorMap tmB1l, const TmaWgradArgs p,
                   int rows_per_cta, int stages, uint32_t tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int N = p.nb0 + p.nb1, nbox_b = N / 64;
    const int a_bytes = 2 * BOX, b_bytes = nbox_b * BOX;
    const int stage_bytes = (a_bytes + b_bytes) * (SPLIT ? 2 : 1);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
    uint64_t* empty = full + stages;
    uint64_t* done = empty + stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, MC ? 4 : 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync_all();   // every CTA's barriers exist before a peer multicasts into them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    uint32_t crank = 0;
    if (MC) asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(crank));

    const int m0 = blockIdx.y * 128;
    const int pbeg = blockIdx.x * rows_per_cta, pend = min(p.P, pbeg + rows_per_cta);
    const int nst = pbeg < pend ? ceil_div(pend - pbeg, KP) : 0;  // rows beyond pend inside the last stage belong to the next CTA:
                                                                  // rows_per_cta is a multiple of KP, so only the tensor end is ragged
                                                                  // and TMA zero-fills it
    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int it = 0; it < nst; ++it) {
                const int p0 = pbeg + it * KP;
                mbar_wait(empty + stage, phase ^ 1);
                uint8_t* st = smem + stage * stage_bytes;
                mbar_expect_tx(full + stage, stage_bytes);
                for (int pl = 0; pl < (SPLIT ? 2 : 1); ++pl) {
                    uint8_t* sa = st + pl * a_bytes;
                    const CUtensorMap* ma = pl ? &tmAl : &tmAh;
                    tma_load_2d(sa, ma, full + stage, m0, p0);
                    tma_load_2d(sa + BOX, ma, full + stage, m0 + 64, p0);
                    uint8_t* sb = st + (SPLIT ? 2 : 1) * a_bytes + pl * b_bytes;
                    const CUtensorMap* mb0 = pl ? &tmB0l : &tmB0h;
                    const CUtensorMap* mb1 = pl ? &tmB1l : &tmB1h;
                    int bx = 0;
                    if (MC) {   // box b of the stage is loaded by CTA (b + plane) % 4 of the cluster and lands in all four
                        for (int c = 0; c < p.nb0; c += 64, ++bx)
                            if (((bx + pl) & 3) == (int)crank) tma_load_2d_mc(sb + bx * BOX, mb0, full + stage, c, p0, (uint16_t)0xF);
                        for (int c = 0; c < p.nb1; c += 64, ++bx)
                            if (((bx + pl) & 3) == (int)crank) tma_load_2d_mc(sb + bx * BOX, mb1, full + stage, c, p0, (uint16_t)0xF);
                    } else {
                        for (int c = 0; c < p.nb0; c += 64, ++bx) tma_load_2d(sb + bx * BOX, mb0, full + stage, c, p0);
                        for (int c = 0; c < p.nb1; c += 64, ++bx) tma_load_2d(sb + bx * BOX, mb1, full + stage, c, p0);
                    }
                }
                if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = idesc_bf16(128, N, 1, 1);
        uint32_t stage = 0, phase = 0;
        for (int it = 0; it < nst; ++it) {
            mbar_wait(full + stage, phase);
            tc_fence_after();
            {   // whole warp, uniform operands; one elected lane issues
                const uint32_t a_hi = smem_u32(smem + stage * stage_bytes);
                const uint32_t a_lo = a_hi + a_bytes;
                const uint32_t b_hi = a_hi + (SPLIT ? 2 : 1) * a_bytes;
                const uint32_t b_lo = b_hi + b_bytes;
#pragma unroll
                for (int k = 0; k < KP / 16; ++k) {
                    // MN-major: LBO = next 64 MN elements (one box), SBO = next 8 K-rows (1024 B); 16 K-rows per MMA = 2048 B
                    const uint64_t ah = desc_sw128(a_hi + k * 2048, BOX, 1024), bh = desc_sw128(b_hi + k * 2048, BOX, 1024);
                    umma_w(tmem, ah, bh, idesc, (it | k) != 0);
                    if (SPLIT) {
                        const uint64_t al = desc_sw128(a_lo + k * 2048, BOX, 1024), bl = desc_sw128(b_lo + k * 2048, BOX, 1024);
                        umma_w(tmem, ah, bl, idesc, 1);
                        umma_w(tmem, al, bh, idesc, 1);
                    }
                }
                if (MC) umma_commit_mc_w(empty + stage, (uint16_t)0xF);
                else umma_commit_w(empty + stage);
                if (it == nst - 1) umma_commit_w(done);
            }
            __syncwarp();
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
        }
    } else if (nst > 0) {
        const int q = warp & 3;
        mbar_wait(done, 0);
        tc_fence_after();
        const int row = m0 + q * 32 + lane;  // row of dW = column of A
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
            if (row < p.Mo) {
                const bool seg1 = c0 >= p.nb0;
                float* C = seg1 ? p.C1 : p.C0;
                const int ldc = seg1 ? p.ldc1 : p.ldc0, cb = seg1 ? c0 - p.nb0 : c0;
                const int tr = seg1 ? p.transpose1 : p.transpose0;
                if (!tr && (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) {
                    // 32 consecutive floats of one output row: eight 16-byte vector reductions instead of 32 scalar ones
                    float* dst = C + (size_t)row * ldc + cb;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(dst + j), "f"(v[j] * p.scale), "f"(v[j + 1] * p.scale),
                                     "f"(v[j + 2] * p.scale), "f"(v[j + 3] * p.scale)
                                     : "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float* dst = tr ? C + (size_t)(cb + j) * ldc + row : C + (size_t)row * ldc + cb + j;
                        atomicAdd(dst, v[j] * p.scale);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync_all();   // no CTA leaves while a peer may still multicast into its shared memory or arrive on its barriers
    if (warp == 1) tmem_dealloc(tmem, tmem_cols);
}

// planes [P, cols] -> box of KP positions x 64 columns
bool make_map_mn(CUtensorMap* map, const void* base, long long P, long long cols, long long ld) {
    auto fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)P};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)KP};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace

// Measured on B200 (tests/tools/time_wgrad.py, dW_ih | dW_hh of one direction at B = 16): 142 us plain, 144 us with the multicast -- for clusters
// of <= 4 CTAs the L2 already serves the four unicast requests of a B tile about as cheaply, so the variant is kept (tested) but off by default.
int g_wgrad_multicast = 0;   // gemm_tma_set_wgrad_multicast
int gemm_tma_set_wgrad_multicast(int on) { const int prev = g_wgrad_multicast; g_wgrad_multicast = on ? 1 : 0; return prev; }

bool gemm_tma_tn_supported(const TmaWgradArgs& a) {
    if (a.P <= 0 || a.Mo <= 0 || a.Mo % 128) return false;
    const int N = a.nb0 + a.nb1;
    if (a.nb0 <= 0 || a.nb0 % 64 || a.nb1 % 64 || N > 256 || N % 16) return false;
    if ((a.lda & 7) || (a.ldb0 & 7) || (a.nb1 && (a.ldb1 & 7))) return false;
    return encode_fn() != nullptr;
}

cudaError_t launch_gemm_tma_tn(const TmaWgradArgs& a, bool split, cudaStream_t st) {
    if (!gemm_tma_tn_supported(a)) return cudaErrorInvalidValue;
    if (split && (!a.A_lo || !a.B0_lo || (a.nb1 && !a.B1_lo))) return cudaErrorInvalidValue;
    const int N = a.nb0 + a.nb1;
    CUtensorMap mAh, mAl, mB0h, mB0l, mB1h, mB1l;
    if (!make_map_mn(&mAh, a.A_hi, a.P, a.Mo, a.lda) || !make_map_mn(&mB0h, a.B0_hi, a.P, a.nb0, a.ldb0)) return cudaErrorInvalidValue;
    mAl = mAh; mB0l = mB0h;
    if (split && (!make_map_mn(&mAl, a.A_lo, a.P, a.Mo, a.lda) || !make_map_mn(&mB0l, a.B0_lo, a.P, a.nb0, a.ldb0))) return cudaErrorInvalidValue;
    mB1h = mB0h; mB1l = mB0l;
    if (a.nb1) {
        if (!make_map_mn(&mB1h, a.B1_hi, a.P, a.nb1, a.ldb1)) return cudaErrorInvalidValue;
        mB1l = mB1h;
        if (split && !make_map_mn(&mB1l, a.B1_lo, a.P, a.nb1, a.ldb1)) return cudaErrorInvalidValue;
    }
    const int stage_bytes = (2 * BOX + (N / 64) * BOX) * (split ? 2 : 1);
    int stages = (200 * 1024) / stage_bytes;
    if (stages > 8) stages = 8;
    const int smem = stages * stage_bytes + 1024 + 256;
    const int mtiles = a.Mo / 128;
    // cluster / multicast variant: the output has a multiple of four 128-row slices.  33 clusters of four CTAs are co-resident on a B200
    // (cudaOccupancyMaxActiveClusters, tests/tools/ubench_cluster.cu): split the positions over 132 / mtiles ranges so that one wave covers the launch
    const bool mc = g_wgrad_multicast && (mtiles % 4 == 0);
    int ksplit = (mc ? 132 : 148) / mtiles;
    if (ksplit < 1) ksplit = 1;
    int rows = ceil_div(ceil_div(a.P, ksplit), KP) * KP;
    if (rows < 4 * KP) rows = 4 * KP;
    dim3 grid(ceil_div(a.P, rows), mtiles);
    const uint32_t tcols = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    cudaError_t e;
    if (mc) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(192);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 4; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (split) {
            e = cudaFuncSetAttribute(gemm_tma_tn_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            return cudaLaunchKernelEx(&cfg, gemm_tma_tn_kernel<true, true>, mAh, mAl, mB0h, mB0l, mB1h, mB1l, a, rows, stages, tcols);
        }
        e = cudaFuncSetAttribute(gemm_tma_tn_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        return cudaLaunchKernelEx(&cfg, gemm_tma_tn_kernel<false, true>, mAh, mAl, mB0h, mB0l, mB1h, mB1l, a, rows, stages, tcols);
    }
    if (split) {
        e = cudaFuncSetAttribute(gemm_tma_tn_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        gemm_tma_tn_kernel<true, false><<<grid, 192, smem, st>>>(mAh, mAl, mB0h, mB0l, mB1h, mB1l, a, rows, stages, tcols);
    } else {
        e = cudaFuncSetAttribute(gemm_tma_tn_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        gemm_tma_tn_kernel<false, false><<<grid, 192, smem, st>>>(mAh, mAl, mB0h, mB0l, mB1h, mB1l, a, rows, stages, tcols);
    }
    return cudaGetLastError();
}

namespace {

// ------------------------------------------------------------------------------------------------ fp32 -> planes
__global__ void __launch_bounds__(256) split_rows_kernel(const float4* __restrict__ src, long long ld4, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                                         long long rows, int C4, int relu) {
    const long long total = rows * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / C4;
        const int c = (int)(i % C4);
        float4 v = ldg_stream(src + r * ld4 + c);
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        uint2 h, l;
        split_pair(v.x, v.y, h.x, l.x);
        split_pair(v.z, v.w, h.y, l.y);
        hi[i] = h;
        if (lo) lo[i] = l;
    }
}
// column sums of a pair of planes: thread = (8-column group, row lane)
__global__ void __launch_bounds__(256) colsum_planes_kernel(const uint4* __restrict__ hi, const uint4* __restrict__ lo, long long rows, int C8,
                                                            float* __restrict__ out, int rows_per_cta) {
    __shared__ float sh[256][9];
    const int lanes = 256 / C8;
    const int c = threadIdx.x % C8, rl = threadIdx.x / C8;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    auto add = [&](const uint4& v) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            acc[2 * t] += __uint_as_float(w[t] << 16);
            acc[2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u);
        }
    };
    if (rl < lanes) {
        long long r = r0 + rl;
        if (lo) {
            for (; r + 3 * lanes < r1; r += 4 * lanes) {   // eight independent 128-bit loads in flight per thread
                const uint4 a0 = hi[r * C8 + c], a1 = hi[(r + lanes) * C8 + c], a2 = hi[(r + 2 * lanes) * C8 + c], a3 = hi[(r + 3 * lanes) * C8 + c];
                const uint4 b0 = lo[r * C8 + c], b1 = lo[(r + lanes) * C8 + c], b2 = lo[(r + 2 * lanes) * C8 + c], b3 = lo[(r + 3 * lanes) * C8 + c];
                add(a0); add(a1); add(a2); add(a3); add(b0); add(b1); add(b2); add(b3);
            }
        } else {
            for (; r + 7 * lanes < r1; r += 8 * lanes) {
                uint4 a[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] = hi[(r + u * lanes) * C8 + c];
#pragma unroll
                for (int u = 0; u < 8; ++u) add(a[u]);
            }
        }
        for (; r < r1; r += lanes) {
            add(hi[r * C8 + c]);
            if (lo) add(lo[r * C8 + c]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sh[threadIdx.x][i] = acc[i];
    __syncthreads();
    if (threadIdx.x < C8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float t = 0.f;
            for (int l = 0; l < lanes; ++l) t += sh[l * C8 + threadIdx.x][i];
            atomicAdd(out + (size_t)threadIdx.x * 8 + i, t);
        }
    }
}

// planes + column sums of the same rows in one pass (bias gradients of the layers whose output gradient is split anyway):
// thread = (float4 column, row lane); a CTA walks a contiguous row range and adds its partial sums with one atomic per column
__global__ void __launch_bounds__(256) split_rows_colsum_kernel(const float4* __restrict__ src, long long ld4, uint2* __restrict__ hi,
                                                                uint2* __restrict__ lo, long long rows, int C4, float* __restrict__ colsum,
                                                                int rows_per_cta, unsigned drop_thr, unsigned drop_key, float drop_scale) {
    __shared__ float4 sh[256];
    const int lanes = 256 / C4;
    const int c = threadIdx.x % C4, rl = threadIdx.x / C4;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl < lanes) {
        auto emit = [&](long long r, float4 v) {
            if (drop_thr) {  // gradient of a dropped sub-layer output: the forward's mask of element (row, column)
                const uint32_t c0 = 4u * (uint32_t)c;
                v.x = drop_keep(drop_key, (uint32_t)r, c0, drop_thr) ? v.x * drop_scale : 0.f;
                v.y = drop_keep(drop_key, (uint32_t)r, c0 + 1, drop_thr) ? v.y * drop_scale : 0.f;
                v.z = drop_keep(drop_key, (uint32_t)r, c0 + 2, drop_thr) ? v.z * drop_scale : 0.f;
                v.w = drop_keep(drop_key, (uint32_t)r, c0 + 3, drop_thr) ? v.w * drop_scale : 0.f;
            }
            uint2 h, l;
            split_pair(v.x, v.y, h.x, l.x);
            split_pair(v.z, v.w, h.y, l.y);
            hi[r * C4 + c] = h;
            if (lo) lo[r * C4 + c] = l;
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        };
        long long r = r0 + rl;
        for (; r + 3 * lanes < r1; r += 4 * lanes) {  // four independent 128-bit loads in flight per thread
            const float4 v0 = ldg_stream(src + r * ld4 + c), v1 = ldg_stream(src + (r + lanes) * ld4 + c);
            const float4 v2 = ldg_stream(src + (r + 2 * lanes) * ld4 + c), v3 = ldg_stream(src + (r + 3 * lanes) * ld4 + c);
            emit(r, v0); emit(r + lanes, v1); emit(r + 2 * lanes, v2); emit(r + 3 * lanes, v3);
        }
        for (; r < r1; r += lanes) emit(r, ldg_stream(src + r * ld4 + c));
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < C4) {
        float4 t = sh[threadIdx.x];
        for (int l = 1; l < lanes; ++l) {
            const float4 v = sh[l * C4 + threadIdx.x];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        float* o = colsum + (size_t)threadIdx.x * 4;
        atomicAdd(o, t.x); atomicAdd(o + 1, t.y); atomicAdd(o + 2, t.z); atomicAdd(o + 3, t.w);
    }
}
}  // namespace

cudaError_t launch_colsum_planes(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long rows, int C, float* out, cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
    if ((C & 7) || C > 2048) return cudaErrorInvalidValue;
    const int lanes = 256 / (C / 8) < 1 ? 1 : 256 / (C / 8);
    int rows_per_cta = (int)ceil_div_ll(rows, 148LL * 2);   // few CTAs (measured: 4 / 2 / 1 CTAs per SM -> 39.9 / 38.7 / 39.0 ms SepFormer step): every CTA ends with one atomic per column on the same C addresses
    if (rows_per_cta < 16 * lanes) rows_per_cta = 16 * lanes;
    colsum_planes_kernel<<<(unsigned)ceil_div_ll(rows, rows_per_cta), 256, 0, st>>>(reinterpret_cast<const uint4*>(hi), reinterpret_cast<const uint4*>(lo),
                                                                                   rows, C / 8, out, rows_per_cta);
    return cudaGetLastError();
}

cudaError_t launch_split_rows_colsum(const float* src, long long ld, __nv_bfloat16* hi, __nv_bfloat16* lo, long long rows, int C,
                                     float* colsum, cudaStream_t st, unsigned drop_thr, unsigned drop_key, float drop_scale) {
    if (rows <= 0) return cudaSuccess;
    if ((C & 3) || (ld & 3) || C > 1024) return cudaErrorInvalidValue;
    const int lanes = 256 / (C / 4) < 1 ? 1 : 256 / (C / 4);
    int rows_per_cta = (int)ceil_div_ll(rows, 148LL * 2);   // few CTAs (measured: 4 / 2 / 1 CTAs per SM -> 39.9 / 38.7 / 39.0 ms SepFormer step): every CTA ends with one atomic per column on the same C addresses
    if (rows_per_cta < 16 * lanes) rows_per_cta = 16 * lanes;
    split_rows_colsum_kernel<<<(unsigned)ceil_div_ll(rows, rows_per_cta), 256, 0, st>>>(
        reinterpret_cast<const float4*>(src), ld / 4, reinterpret_cast<uint2*>(hi), reinterpret_cast<uint2*>(lo), rows, C / 4, colsum, rows_per_cta,
        drop_thr, drop_key, drop_scale);
    return cudaGetLastError();
}

cudaError_t launch_split_rows(const float* src, long long ld, __nv_bfloat16* hi, __nv_bfloat16* lo, long long rows, int C, int relu,
                              cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
    if ((C & 3) || (ld & 3)) return cudaErrorInvalidValue;
    long long total = rows * (C / 4);
    long long blocks = ceil_div_ll(total, 256);
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    split_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), ld / 4, reinterpret_cast<uint2*>(hi),
                                                         reinterpret_cast<uint2*>(lo), rows, C / 4, relu);
    return cudaGetLastError();
}

}  // namespace dp
