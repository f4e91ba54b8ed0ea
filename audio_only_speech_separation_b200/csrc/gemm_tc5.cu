// tcgen05 / TMEM implementations of the position-major GEMMs (sm_100a 5th-generation tensor cores).
//
//   gemm_nt_tc5 : C[M,N] (=|+=) A[M,K] W^T (+bias)     LSTM input projection, its input gradient, dH = dY Wp
//   gemm_tn_tc5 : dW[Mo,No] += A[P,Mo]^T B[P,No]        weight gradients (contraction over positions)
//
// Operands are fp32 in HBM (activations) or pre-split bf16 (weights).  Threads convert fp32 -> bf16 hi/lo on the fly
// and write shared memory in the UMMA "no-swizzle, K-major" canonical layout: 8x16-byte core matrices, i.e.
// element (row r, k) lives at  (k/8)*LBO + (r/8)*SBO + (r%8)*16 + (k%8)*2  with SBO = 128 B (rows of a core matrix are
// contiguous) and LBO = rows*16 + 16 B (one k-chunk of all rows, padded by 16 B so the 128-bit stores of a
// quarter-warp hit distinct banks).  One elected thread issues tcgen05.mma (M = 128, N = 64..256, K = 16 per
// instruction) for the three split products hi*hi + hi*lo + lo*hi; accumulators live in TMEM (fp32) and come back
// through tcgen05.ld for the epilogue.  tcgen05.commit -> mbarrier signals "operands consumed / accumulator ready".
// Every wait is bounded (trap instead of hang).
#include "common.cuh"
#include "kernels.h"

namespace dp {
namespace {

constexpr int TM = 128;  // UMMA_M: rows per tile, one TMEM lane per row
constexpr int TK = 64;   // K elements per shared-memory chunk (4 MMAs of K = 16)
constexpr uint32_t SPIN_LIMIT = 1u << 27;

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > SPIN_LIMIT) __trap();
    }
}
// shared-memory matrix descriptor, SWIZZLE_NONE, sm_100 version field = 1 (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, dense
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// 32 consecutive fp32 columns of this warp's 32 TMEM lanes (thread = lane = accumulator row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
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

// ------------------------------------------------------------------------------------------------
// C[M,N] (=|+=) A[M,K] W[N,K]^T (+ bias).  W given as bf16 hi/lo, [N,K] (w_kn = 0) or [K,N] (w_kn = 1).
// One CTA (128 threads) computes a 128 x BN tile; 2 CTAs per SM overlap each other's load / MMA / epilogue phases.
template <int BN, bool SPLIT>
__global__ void __launch_bounds__(128) gemm_nt_tc5_kernel(const GemmNtArgs p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr uint32_t LBO_A = TM * 16 + 16, LBO_W = BN * 16 + 16;
    constexpr uint32_t A_BYTES = (TK / 8) * LBO_A, W_BYTES = (TK / 8) * LBO_W;
    unsigned char* a_hi = smem;
    unsigned char* a_lo = a_hi + A_BYTES;
    unsigned char* w_hi = a_lo + A_BYTES;
    unsigned char* w_lo = w_hi + W_BYTES;
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(w_lo + W_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t IDESC = instr_desc(TM, BN);

    const int nchunks = p.N / BN, mtiles = ceil_div(p.M, TM), ntiles = mtiles * nchunks, nk = p.K / TK;
    uint32_t phase = 0;
    bool pending = false;  // an un-waited commit is outstanding (its MMAs may still read shared memory)

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int m0 = (tile / nchunks) * TM, n0 = (tile % nchunks) * BN;
        for (int kc = 0; kc < nk; ++kc) {
            const int k0 = kc * TK;
            // ---- global -> registers (issued before waiting for the previous MMAs) ----
            float4 av[8][2];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = m0 + (tid >> 3) + 16 * i, kq = tid & 7;
                if (row < p.M) {
                    long long r = row;
                    if (p.a_rpb) r += (long long)(row / p.a_rpb) * p.a_skip;
                    const float4* src = reinterpret_cast<const float4*>(p.A + r * p.lda + k0 + kq * 8);
                    av[i][0] = src[0];
                    av[i][1] = src[1];
                } else {
                    av[i][0] = av[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (pending) {  // previous chunk's MMAs must be done with the shared-memory operands
                mbar_wait(mma_bar, phase);
                phase ^= 1;
                pending = false;
            }
            // ---- A: split to bf16 hi/lo, one 16-byte k-chunk per (row, kq) ----
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rl = (tid >> 3) + 16 * i, kq = tid & 7;
                uint4 hi, lo;
                split_pair(av[i][0].x, av[i][0].y, hi.x, lo.x);
                split_pair(av[i][0].z, av[i][0].w, hi.y, lo.y);
                split_pair(av[i][1].x, av[i][1].y, hi.z, lo.z);
                split_pair(av[i][1].z, av[i][1].w, hi.w, lo.w);
                const uint32_t off = kq * LBO_A + rl * 16;
                *reinterpret_cast<uint4*>(a_hi + off) = hi;
                if (SPLIT) *reinterpret_cast<uint4*>(a_lo + off) = lo;
            }
            // ---- W tile [BN rows (n)] x [TK (k)] ----
            if (!p.w_kn) {
#pragma unroll
                for (int i = 0; i < BN / 16; ++i) {
                    const int nl = (tid >> 3) + 16 * i, kq = tid & 7;
                    const size_t g = (size_t)(n0 + nl) * p.ldw + k0 + kq * 8;
                    const uint32_t off = kq * LBO_W + nl * 16;
                    *reinterpret_cast<uint4*>(w_hi + off) = *reinterpret_cast<const uint4*>(p.Whi + g);
                    if (SPLIT) *reinterpret_cast<uint4*>(w_lo + off) = *reinterpret_cast<const uint4*>(p.Wlo + g);
                }
            } else {  // W stored [K,N]: gather 8 k values per (n) chunk -- 2-byte accesses, small tiles only (L2 resident)
                for (int idx = tid; idx < BN * (TK / 8); idx += 128) {
                    const int nl = idx % BN, kq = idx / BN;
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const size_t g0 = (size_t)(k0 + kq * 8 + 2 * j) * p.ldw + n0 + nl;
                        const uint16_t h0 = reinterpret_cast<const uint16_t*>(p.Whi)[g0], h1 = reinterpret_cast<const uint16_t*>(p.Whi)[g0 + p.ldw];
                        h[j] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                        if (SPLIT) {
                            const uint16_t l0 = reinterpret_cast<const uint16_t*>(p.Wlo)[g0], l1 = reinterpret_cast<const uint16_t*>(p.Wlo)[g0 + p.ldw];
                            l[j] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                        }
                    }
                    const uint32_t off = kq * LBO_W + nl * 16;
                    *reinterpret_cast<uint4*>(w_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                    if (SPLIT) *reinterpret_cast<uint4*>(w_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
                }
            }
            proxy_fence();  // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < TK / 16; ++ks) {
                    const uint64_t ah = smem_desc(smem_u32(a_hi) + ks * 2 * LBO_A, LBO_A, 128);
                    const uint64_t wh = smem_desc(smem_u32(w_hi) + ks * 2 * LBO_W, LBO_W, 128);
                    umma_bf16(tmem, ah, wh, IDESC, (kc | ks) != 0);
                    if (SPLIT) {
                        const uint64_t al = smem_desc(smem_u32(a_lo) + ks * 2 * LBO_A, LBO_A, 128);
                        const uint64_t wl = smem_desc(smem_u32(w_lo) + ks * 2 * LBO_W, LBO_W, 128);
                        umma_bf16(tmem, ah, wl, IDESC, 1);
                        umma_bf16(tmem, al, wh, IDESC, 1);
                    }
                }
                umma_commit(mma_bar);
            }
            pending = true;
        }
        // ---- epilogue: TMEM -> registers -> global (thread = row) ----
        mbar_wait(mma_bar, phase);
        phase ^= 1;
        pending = false;
        tc_fence_after();
        const int row = m0 + warp * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            if (row < p.M) {
                float* dst = p.C + (size_t)row * p.ldc + n0 + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    if (p.bias) {
                        const float4 b = *reinterpret_cast<const float4*>(p.bias + n0 + c0 + j);
                        o.x = fmaf(p.bias_scale, b.x, o.x); o.y = fmaf(p.bias_scale, b.y, o.y);
                        o.z = fmaf(p.bias_scale, b.z, o.z); o.w = fmaf(p.bias_scale, b.w, o.w);
                    }
                    if (p.accumulate) {
                        const float4 c = *reinterpret_cast<const float4*>(dst + j);
                        o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
                    }
                    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4*>(dst + j) = o;
                }
            }
        }
        tc_fence_before();
        __syncthreads();  // accumulator drained by every warp before the next tile's first MMA overwrites it
        tc_fence_after();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc<BN>(tmem);
}

template <int BN>
cudaError_t launch_nt(const GemmNtArgs& a, bool split, cudaStream_t st) {
    const size_t smem = 2 * (TK / 8) * (TM * 16 + 16) + 2 * (TK / 8) * (BN * 16 + 16) + 64;
    const int ntiles = ceil_div(a.M, TM) * (a.N / BN);
    const int grid = ntiles < 2 * 148 ? ntiles : 2 * 148;
    cudaError_t e;
    if (split) {
        e = cudaFuncSetAttribute(gemm_nt_tc5_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gemm_nt_tc5_kernel<BN, true><<<grid, 128, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(gemm_nt_tc5_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gemm_nt_tc5_kernel<BN, false><<<grid, 128, smem, st>>>(a);
    }
    return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------
// dW[Mo,No] += scale * sum_p A[p,Mo]^T B[p',No]  (p' = p + shift where the time index allows, see GemmTnArgs).
// The contraction index is the position p (slow index of both operands), so threads transpose 8(p) x 4(col) register
// blocks on the way to shared memory and both operands end up K-major.  One CTA owns a contiguous range of
// positions and the WHOLE output: MT = Mo/128 accumulator tiles of No columns each (MT * No <= 512 TMEM columns), so
// the big operand (dG, [P,1024]) is read from HBM exactly once.  Epilogue: TMEM -> registers -> fp32 atomics.
constexpr int PK = 32;  // positions per shared-memory chunk (2 MMAs of K = 16)

template <int MT, int NO, bool SPLIT>
__global__ void __launch_bounds__(256) gemm_tn_tc5_kernel(const GemmTnArgs p, int rows_per_cta, int transpose_out) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int MO = MT * 128;
    constexpr uint32_t LBO_A = MO * 16 + 16, LBO_B = NO * 16 + 16;
    constexpr uint32_t A_BYTES = (PK / 8) * LBO_A, B_BYTES = (PK / 8) * LBO_B;
    constexpr int NCOLS = MT * NO;  // 512, 256, 128 or 64
    unsigned char* a_hi = smem;
    unsigned char* a_lo = a_hi + A_BYTES;
    unsigned char* b_hi = a_lo + A_BYTES;
    unsigned char* b_lo = b_hi + B_BYTES;
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(b_lo + B_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) tmem_alloc<NCOLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t IDESC = instr_desc(128, NO);

    const int pbeg = blockIdx.x * rows_per_cta, pend = min(p.P, pbeg + rows_per_cta);
    uint32_t phase = 0;
    bool pending = false, any = false;

    // (8 positions) x (4 columns) register blocks: unit u -> column quad cq = u % QUADS, position octet po = u / QUADS
    auto load_block = [&](const float* base, long long ld, int p0, int cq, int po, bool shifted, float4 (&v)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int pr = p0 + po * 8 + j;
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pr < pend) {
                long long q = pr;
                bool ok = true;
                if (shifted && p.tmod) {
                    const int t = (pr / p.tdiv) % p.tmod;
                    ok = p.shift < 0 ? (t > 0) : (t < p.tmod - 1);
                    q += p.shift;
                }
                if (shifted && p.b_rpb) q += (long long)(pr / p.b_rpb) * p.b_skip;
                if (ok) v[j] = *reinterpret_cast<const float4*>(base + q * ld + cq * 4);
            }
        }
    };
    auto store_block = [&](unsigned char* hi, unsigned char* lo, uint32_t lbo, int cq, int po, const float4 (&v)[8]) {
        const float* f = reinterpret_cast<const float*>(v);  // f[j*4 + c]: position j, column c
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4 h, l;
            split_pair(f[0 * 4 + c], f[1 * 4 + c], h.x, l.x);
            split_pair(f[2 * 4 + c], f[3 * 4 + c], h.y, l.y);
            split_pair(f[4 * 4 + c], f[5 * 4 + c], h.z, l.z);
            split_pair(f[6 * 4 + c], f[7 * 4 + c], h.w, l.w);
            const uint32_t off = po * lbo + (cq * 4 + c) * 16;
            *reinterpret_cast<uint4*>(hi + off) = h;
            if (SPLIT) *reinterpret_cast<uint4*>(lo + off) = l;
        }
    };

    constexpr int A_UNITS = (MO / 4) * (PK / 8), B_UNITS = (NO / 4) * (PK / 8);
    for (int p0 = pbeg; p0 < pend; p0 += PK) {
        // A operand (dY / dG), in passes of 256 units so at most one register block per thread is in flight
        for (int u0 = 0; u0 < A_UNITS; u0 += 256) {
            const int u = u0 + tid;
            float4 v[8];
            const int cq = u % (MO / 4), po = u / (MO / 4);
            if (u < A_UNITS) load_block(p.A, p.lda, p0, cq, po, false, v);
            if (u0 == 0 && pending) {
                mbar_wait(mma_bar, phase);
                phase ^= 1;
                pending = false;
            }
            if (u < A_UNITS) store_block(a_hi, a_lo, LBO_A, cq, po, v);
        }
        for (int u0 = 0; u0 < B_UNITS; u0 += 256) {
            const int u = u0 + tid;
            if (u < B_UNITS) {
                float4 v[8];
                const int cq = u % (NO / 4), po = u / (NO / 4);
                load_block(p.B, p.ldb, p0, cq, po, true, v);
                store_block(b_hi, b_lo, LBO_B, cq, po, v);
            }
        }
        proxy_fence();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < PK / 16; ++ks) {
                const uint64_t bh = smem_desc(smem_u32(b_hi) + ks * 2 * LBO_B, LBO_B, 128);
                const uint64_t bl = smem_desc(smem_u32(b_lo) + ks * 2 * LBO_B, LBO_B, 128);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const uint64_t ah = smem_desc(smem_u32(a_hi) + ks * 2 * LBO_A + mt * 128 * 16, LBO_A, 128);
                    const uint32_t d = tmem + mt * NO;
                    umma_bf16(d, ah, bh, IDESC, (any || ks) ? 1u : 0u);
                    if (SPLIT) {
                        const uint64_t al = smem_desc(smem_u32(a_lo) + ks * 2 * LBO_A + mt * 128 * 16, LBO_A, 128);
                        umma_bf16(d, ah, bl, IDESC, 1);
                        umma_bf16(d, al, bh, IDESC, 1);
                    }
                }
            }
            umma_commit(mma_bar);
        }
        pending = true;
        any = true;
    }
    if (pending) {
        mbar_wait(mma_bar, phase);
        phase ^= 1;
    }
    tc_fence_after();
    if (any) {
        // 8 warps: warp w reads lane quarter (w & 3); warps 0-3 take the even 32-column groups, warps 4-7 the odd ones
        const int q = warp & 3, par = warp >> 2;
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll 1
            for (int c0 = par * 32; c0 < NO; c0 += 64) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + mt * NO + c0, v);
                const int row = mt * 128 + q * 32 + lane;
                if (row < p.Mo) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = c0 + j;
                        if (col < p.No) {
                            float* dst = transpose_out ? p.C + (size_t)col * p.ldc + row : p.C + (size_t)row * p.ldc + col;
                            atomicAdd(dst, v[j] * p.scale);
                        }
                    }
                }
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<NCOLS>(tmem);
}

template <int MT, int NO>
cudaError_t launch_tn(const GemmTnArgs& a, bool split, int transpose_out, cudaStream_t st) {
    const size_t smem = 2 * (PK / 8) * (MT * 128 * 16 + 16) + 2 * (PK / 8) * (NO * 16 + 16) + 64;
    int rows = ceil_div(ceil_div(a.P, 148), PK) * PK;
    if (rows < 4 * PK) rows = 4 * PK;
    const int grid = ceil_div(a.P, rows);
    cudaError_t e;
    if (split) {
        e = cudaFuncSetAttribute(gemm_tn_tc5_kernel<MT, NO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gemm_tn_tc5_kernel<MT, NO, true><<<grid, 256, smem, st>>>(a, rows, transpose_out);
    } else {
        e = cudaFuncSetAttribute(gemm_tn_tc5_kernel<MT, NO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gemm_tn_tc5_kernel<MT, NO, false><<<grid, 256, smem, st>>>(a, rows, transpose_out);
    }
    return cudaGetLastError();
}

}  // namespace

bool gemm_nt_tc5_supported(const GemmNtArgs& a) {
    if (a.stats != nullptr || a.M <= 0 || a.relu_a || a.mask || a.relu > 1 || a.mul_c) return false;
    if (a.K % TK || (a.lda & 3) || (a.ldc & 3) || (a.ldw & 7)) return false;
    if (!(a.N == 64 || a.N == 128 || a.N % 256 == 0)) return false;
    return true;
}

cudaError_t launch_gemm_nt_tc5(const GemmNtArgs& a, bool split, cudaStream_t st) {
    if (!gemm_nt_tc5_supported(a)) return cudaErrorInvalidValue;
    if (a.N == 64) return launch_nt<64>(a, split, st);
    if (a.N == 128) return launch_nt<128>(a, split, st);
    return launch_nt<256>(a, split, st);
}

}  // namespace dp

namespace dp {

// Supported weight-gradient shapes: (Mo, No) in {(1024,64), (512,128), (256,64)}; other shapes use the legacy kernel.
bool gemm_tn_tc5_supported(const GemmTnArgs& a) {
    if ((a.lda & 3) || (a.ldb & 3) || a.P <= 0 || a.relu_b) return false;
    return (a.Mo == 1024 && a.No == 64) || (a.Mo == 512 && a.No == 128) || (a.Mo == 256 && a.No == 64);
}

cudaError_t launch_gemm_tn_tc5(const GemmTnArgs& a, bool split, int transpose_out, cudaStream_t st) {
    if (!gemm_tn_tc5_supported(a)) return cudaErrorInvalidValue;
    if (a.Mo == 1024) return launch_tn<8, 64>(a, split, transpose_out, st);
    if (a.Mo == 512) return launch_tn<4, 128>(a, split, transpose_out, st);
    return launch_tn<2, 64>(a, split, transpose_out, st);
}

}  // namespace dp
