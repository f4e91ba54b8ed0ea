// Self-attention core on the 5th-generation tensor cores (sm_100a): softmax(q k^T / sqrt(d)) v for the short sequences of
// the dual-path transformer blocks (DPTNet: L = 82..100, d = 16, dptnet.py:48,75; SepFormer: L = 130..258, d = 32,
// sepformer.py:124-215).  One CTA = one (sequence, head):
//   * Q, K, V tiles of the head arrive by TMA straight from the bf16 hi/lo planes the QKV GEMM wrote ([P, 3E]); a 4-D tensor
//     map turns the strided positions of an inter-chunk sequence into one box, so no permute/contiguous copy exists
//   * S = Q K^T : tcgen05.mma, A = Q (K-major), B = K (K-major), accumulator [128 queries x L keys] fp32 in tensor memory
//   * softmax  : thread = query row, tcgen05.ld the row, exact max / exp2 / sum in fp32; probabilities go back to tensor memory
//     as bf16 hi (and lo) with tcgen05.st -- the [B*S*h, L, L] tensor the reference materialises (sepformer.py:142) never
//     leaves the SM
//   * O = P V  : tcgen05.mma with the A operand (P) read from tensor memory and B = V as an MN-major operand (the key index is
//     the contraction index and the slow index of the V tile); O / rowsum leaves as hi/lo planes for the out-projection GEMM
// fp32-parity mode forms hi*hi + hi*lo + lo*hi for both contractions; bf16 mode a single product.
// Shared-memory tiles use the canonical UMMA layouts for 64-byte (d = 32) or 32-byte (d = 16) rows (SWIZZLE_64B / SWIZZLE_32B).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstring>

#include "common.cuh"
#include "kernels.h"
#include "tc5_common.cuh"

namespace dp {
namespace {

constexpr float kLog2e = 1.4426950408889634f;

// shared-memory matrix descriptor with an explicit swizzle mode (2 = 128 B, 4 = 64 B, 6 = 32 B)
__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr), "r"(r[0]),
        "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
        "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct AttnTcArgs {
    float* O;               // optional fp32 [P, E]
    __nv_bfloat16* O_hi;    // planes [P, E]
    __nv_bfloat16* O_lo;
    float* LSE;             // optional [P, heads] (log2 domain)
    int E, heads;
    int inter, len, nseq_or_K, S, B;
    long long s_t;
    float scale_log2;
    unsigned drop_thr, drop_key;   // attention-probability dropout (0 = off): mask element (query position * heads + head, key index)
    float drop_scale;
};

// LMAX: 128 or 256 keys / queries per sequence at most (tensor-memory columns of S)
template <int D, int LMAX, bool SPLIT, bool DROP>
__global__ void __launch_bounds__(288, (LMAX == 128 ? 2 : 1))   // 256 tensor-memory columns per CTA at LMAX = 128: two CTAs share an SM
attn_tc5_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmL, const AttnTcArgs p) {
    constexpr int ROWB = D * 2;                       // bytes per tile row
    constexpr uint32_t LAYOUT = D == 32 ? 4u : 6u;    // SWIZZLE_64B / SWIZZLE_32B
    constexpr int SBO = 8 * ROWB;
    constexpr int TILE = LMAX * ROWB;                 // one [LMAX rows x D] tile
    constexpr int PL = SPLIT ? 2 : 1;
    constexpr int MT = LMAX / 128;
    constexpr uint32_t COL_S = 0, COL_PH = LMAX, COL_PL = LMAX + LMAX / 2;   // S fp32 | P hi (2 keys per column) | P lo
    constexpr uint32_t TCOLS = 2 * LMAX;              // 256 or 512
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                 // [PL][TILE]
    uint8_t* sK = sQ + PL * TILE;
    uint8_t* sV = sK + PL * TILE;
    uint64_t* ld_full = reinterpret_cast<uint64_t*>(sV + PL * TILE);
    uint64_t* s_full = ld_full + 1;
    uint64_t* p_ready = s_full + 1;
    uint64_t* o_full = p_ready + 1;
    uint64_t* o_read = o_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_read + 1);
    float* xch = reinterpret_cast<float*>(tmem_slot + 2);   // [2 halves][2 (max, sum)][128 rows]: the two threads of a row exchange through it

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x, h = blockIdx.y;
    const int L = p.len;
    const int Lp = (L + 15) & ~15;
    // sequence -> TMA coordinates and position of (sequence, t = 0)
    int c1, c2, c3;          // coordinates of the non-row dims
    long long base;
    if (!p.inter) { c1 = 0; c2 = q; c3 = 0; base = (long long)q * L; }
    else { const int b = q / p.nseq_or_K, k = q % p.nseq_or_K; c1 = k; c2 = 0; c3 = b; base = (long long)b * p.S * p.nseq_or_K + k; }

    if (threadIdx.x == 0) {
        mbar_init(ld_full, 1); mbar_init(s_full, 1); mbar_init(p_ready, 256); mbar_init(o_full, 1); mbar_init(o_read, 128);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        prefetch_tmap(&tmH);
        if (SPLIT) prefetch_tmap(&tmL);
    }
    if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA loads + MMA issue (one warp, uniform control flow) =====================
        if (lane == 0) {
            mbar_expect_tx(ld_full, 3 * PL * TILE);
            for (int pl = 0; pl < PL; ++pl) {
                const CUtensorMap* m = pl ? &tmL : &tmH;
                for (int part = 0; part < 3; ++part) {          // q | k | v thirds of the in_proj output
                    uint8_t* dst = (part == 0 ? sQ : part == 1 ? sK : sV) + pl * TILE;
                    const int col = part * p.E + h * D;
                    for (int mt = 0; mt < MT; ++mt) {
                        // rows = time index: dim 1 (intra) or dim 2 (inter)
                        if (!p.inter) tma_load_4d(dst + mt * 128 * ROWB, m, ld_full, col, mt * 128, c2, c3);
                        else tma_load_4d(dst + mt * 128 * ROWB, m, ld_full, col, c1, mt * 128, c3);
                    }
                }
            }
        }
        mbar_wait(ld_full, 0);
        tc_fence_after();
        const uint32_t idesc_s = idesc_bf16(128, Lp, 0, 0);
        const uint32_t idesc_o = idesc_bf16(128, D, 0, 1);      // B = V is MN-major
        const int nmt = (L + 127) / 128;
        for (int mt = 0; mt < nmt; ++mt) {
            if (mt > 0) { mbar_wait(o_read, (mt - 1) & 1); tc_fence_after(); }   // O (aliasing S) of the previous tile has been read
            // ---- S = Q K^T
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
                const uint64_t qh = desc_sw(smem_u32(sQ) + mt * 128 * ROWB + k * 32, 16, SBO, LAYOUT);
                const uint64_t kh = desc_sw(smem_u32(sK) + k * 32, 16, SBO, LAYOUT);
                umma_w(tmem + COL_S, qh, kh, idesc_s, k != 0);
                if (SPLIT) {
                    const uint64_t ql = desc_sw(smem_u32(sQ) + TILE + mt * 128 * ROWB + k * 32, 16, SBO, LAYOUT);
                    const uint64_t kl = desc_sw(smem_u32(sK) + TILE + k * 32, 16, SBO, LAYOUT);
                    umma_w(tmem + COL_S, qh, kl, idesc_s, 1);
                    umma_w(tmem + COL_S, ql, kh, idesc_s, 1);
                }
            }
            umma_commit_w(s_full);
            // ---- O = P V (P from tensor memory)
            mbar_wait(p_ready, mt & 1);
            tc_fence_after();
            for (int j = 0; j < Lp / 16; ++j) {
                const uint64_t vh = desc_sw(smem_u32(sV) + j * 16 * ROWB, 16, SBO, LAYOUT);
                umma_ts_w(tmem + COL_S, tmem + COL_PH + j * 8, vh, idesc_o, j != 0);
                if (SPLIT) {
                    const uint64_t vl = desc_sw(smem_u32(sV) + TILE + j * 16 * ROWB, 16, SBO, LAYOUT);
                    umma_ts_w(tmem + COL_S, tmem + COL_PH + j * 8, vl, idesc_o, 1);
                    umma_ts_w(tmem + COL_S, tmem + COL_PL + j * 8, vh, idesc_o, 1);
                }
            }
            umma_commit_w(o_full);
        }
    } else {
        // ===================== softmax + output: TWO threads per query row (warps w and w + 4 share a tensor-memory lane quarter and
        // take alternate 32-column chunks of the row; maximum and sum are exchanged through shared memory) =====================
        const int qd = warp & 3, r = qd * 32 + lane;
        const int half = (warp - 1) >> 2;                 // warps 1-4: chunks 0, 2, ...; warps 5-8: chunks 1, 3, ...
        const uint32_t lane_addr = tmem + ((uint32_t)(qd * 32) << 16);
        const int nmt = (L + 127) / 128;
        for (int mt = 0; mt < nmt; ++mt) {
            const int qi = mt * 128 + r;
            mbar_wait(s_full, mt & 1);
            tc_fence_after();
            float mx = -INFINITY;
            for (int c0 = half * 32; c0 < Lp; c0 += 64) {
                float v[32];
                tmem_ld32_nowait(lane_addr + COL_S + c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < L) mx = fmaxf(mx, v[j]);
            }
            xch[(half * 2 + 0) * 128 + r] = mx;
            asm volatile("bar.sync 1, 256;\n" ::: "memory");
            mx = fmaxf(mx, xch[((half ^ 1) * 2 + 0) * 128 + r]);
            const float mxs = mx * p.scale_log2;
            const uint32_t drow = (uint32_t)(base + (long long)qi * p.s_t) * (uint32_t)p.heads + (uint32_t)h;
            float sum = 0.f;
            for (int c0 = half * 32; c0 < Lp; c0 += 64) {
                float v[32];
                tmem_ld32_nowait(lane_addr + COL_S + c0, v);
                uint32_t ph[16], plo[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float e0 = (c0 + j < L) ? exp2f(fmaf(v[j], p.scale_log2, -mxs)) : 0.f;
                    float e1 = (c0 + j + 1 < L) ? exp2f(fmaf(v[j + 1], p.scale_log2, -mxs)) : 0.f;
                    sum += e0 + e1;   // the softmax denominator is that of the un-dropped probabilities
                    if (DROP) {
                        if (!drop_keep(p.drop_key, drow, (uint32_t)(c0 + j), p.drop_thr)) e0 = 0.f;
                        if (!drop_keep(p.drop_key, drow, (uint32_t)(c0 + j + 1), p.drop_thr)) e1 = 0.f;
                    }
                    split_pair(e0, e1, ph[j >> 1], plo[j >> 1]);
                }
                tmem_st16(lane_addr + COL_PH + (c0 >> 1), ph);
                if (SPLIT) tmem_st16(lane_addr + COL_PL + (c0 >> 1), plo);
            }
            xch[(half * 2 + 1) * 128 + r] = sum;
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(p_ready);
            asm volatile("bar.sync 1, 256;\n" ::: "memory");
            if (half) continue;   // the first thread of the pair writes the output row
            sum += xch[(1 * 2 + 1) * 128 + r];
            // ---- O row
            mbar_wait(o_full, mt & 1);
            tc_fence_after();
            float o[32];
            tmem_ld32_nowait(lane_addr + COL_S, o);   // D <= 32 columns are meaningful
            tc_fence_before();
            mbar_arrive(o_read);
            if (qi < L) {
                const float inv = (DROP ? p.drop_scale : 1.0f) / sum;
                const size_t pos = (size_t)(base + (long long)qi * p.s_t);
                if (p.O != nullptr) {
                    float4* dst = reinterpret_cast<float4*>(p.O + pos * p.E + h * D);
#pragma unroll
                    for (int c = 0; c < D / 4; ++c) dst[c] = make_float4(o[4 * c] * inv, o[4 * c + 1] * inv, o[4 * c + 2] * inv, o[4 * c + 3] * inv);
                }
                if (p.O_hi != nullptr) {
                    uint32_t hi[D / 2], lo[D / 2];
#pragma unroll
                    for (int d = 0; d < D / 2; ++d) split_pair(o[2 * d] * inv, o[2 * d + 1] * inv, hi[d], lo[d]);
                    uint4* dh = reinterpret_cast<uint4*>(p.O_hi + pos * p.E + h * D);
#pragma unroll
                    for (int c = 0; c < D / 8; ++c) dh[c] = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                    if (p.O_lo != nullptr) {
                        uint4* dl = reinterpret_cast<uint4*>(p.O_lo + pos * p.E + h * D);
#pragma unroll
                        for (int c = 0; c < D / 8; ++c) dl[c] = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                    }
                }
                if (p.LSE != nullptr) p.LSE[pos * p.heads + h] = mxs + log2f(sum);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn3() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(q);
    }
    return fn;
}

// QKV planes [P, 3E]: intra [nseq][len][3E] (box = 128 time rows of one sequence), inter [B][S][K][3E] (box = 128 s rows of one (b,k))
bool make_qkv_map(CUtensorMap* map, const void* base, int E, int D, const LstmFusedGeom& gm) {
    auto fn = encode_fn3();
    if (!fn) return false;
    cuuint64_t gdim[4], gstr[3];
    cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
    const cuuint64_t rowb = (cuuint64_t)3 * E * 2;
    if (!gm.inter) {
        gdim[0] = 3 * E; gdim[1] = (cuuint64_t)gm.len; gdim[2] = (cuuint64_t)gm.nseq; gdim[3] = 1;
        gstr[0] = rowb; gstr[1] = (cuuint64_t)gm.len * rowb; gstr[2] = (cuuint64_t)gm.len * gm.nseq * rowb;
        box[0] = D; box[1] = 128; box[2] = 1; box[3] = 1;
    } else {
        gdim[0] = 3 * E; gdim[1] = (cuuint64_t)gm.K; gdim[2] = (cuuint64_t)gm.S; gdim[3] = (cuuint64_t)gm.B;
        gstr[0] = rowb; gstr[1] = (cuuint64_t)gm.K * rowb; gstr[2] = (cuuint64_t)gm.K * gm.S * rowb;
        box[0] = D; box[1] = 1; box[2] = 128; box[3] = 1;
    }
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              D == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int D, int LMAX, bool SPLIT>
cudaError_t launch(const CUtensorMap& mh, const CUtensorMap& ml, const AttnTcArgs& a, int nseq, cudaStream_t st) {
    const int smem = 3 * (SPLIT ? 2 : 1) * LMAX * D * 2 + 1024 + 256 + 2 * 2 * 128 * 4;
    dim3 grid(nseq, a.heads);
    cudaError_t e;
    if (a.drop_thr) {
        e = cudaFuncSetAttribute(attn_tc5_kernel<D, LMAX, SPLIT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attn_tc5_kernel<D, LMAX, SPLIT, true><<<grid, 288, smem, st>>>(mh, ml, a);
    } else {
        e = cudaFuncSetAttribute(attn_tc5_kernel<D, LMAX, SPLIT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attn_tc5_kernel<D, LMAX, SPLIT, false><<<grid, 288, smem, st>>>(mh, ml, a);
    }
    return cudaGetLastError();
}

}  // namespace

bool attn_tc5_supported(int E, int heads, const LstmFusedGeom& gm) {
    if (heads <= 0 || E % heads) return false;
    const int D = E / heads;
    if (D != 16 && D != 32) return false;
    if (gm.len < 1 || gm.len > 256) return false;
    return encode_fn3() != nullptr;
}

cudaError_t launch_attn_fwd_tc5(const __nv_bfloat16* qkv_hi, const __nv_bfloat16* qkv_lo, float* O, __nv_bfloat16* O_hi, __nv_bfloat16* O_lo,
                                float* LSE, int E, int heads, const LstmFusedGeom& gm, bool split, cudaStream_t st, unsigned drop_thr,
                                unsigned drop_key, float drop_scale) {
    if (!attn_tc5_supported(E, heads, gm)) return cudaErrorInvalidValue;
    if (split && !qkv_lo) return cudaErrorInvalidValue;
    const int D = E / heads;
    CUtensorMap mh, ml;
    if (!make_qkv_map(&mh, qkv_hi, E, D, gm)) return cudaErrorInvalidValue;
    ml = mh;
    if (split && !make_qkv_map(&ml, qkv_lo, E, D, gm)) return cudaErrorInvalidValue;
    AttnTcArgs a;
    memset(&a, 0, sizeof(a));
    a.O = O; a.O_hi = O_hi; a.O_lo = O_lo; a.LSE = LSE; a.E = E; a.heads = heads;
    a.inter = gm.inter; a.len = gm.len; a.nseq_or_K = gm.inter ? gm.K : gm.nseq; a.S = gm.S; a.B = gm.B;
    a.s_t = gm.inter ? gm.K : 1;
    a.scale_log2 = kLog2e / sqrtf((float)D);
    a.drop_thr = drop_thr; a.drop_key = drop_key; a.drop_scale = drop_scale;
    const bool big = gm.len > 128;
#define DP_ATT(DD)                                                                                                   \
    (big ? (split ? launch<DD, 256, true>(mh, ml, a, gm.nseq, st) : launch<DD, 256, false>(mh, ml, a, gm.nseq, st))  \
         : (split ? launch<DD, 128, true>(mh, ml, a, gm.nseq, st) : launch<DD, 128, false>(mh, ml, a, gm.nseq, st)))
    return D == 32 ? DP_ATT(32) : DP_ATT(16);
#undef DP_ATT
}

}  // namespace dp
