// Position-major GEMMs of the dual-path hot path (sm_100a).
//
//   gemm_nt : C[M,N] = A[M,K] * W^T       M = positions (10^5..10^6 rows), N,K in {16..1024}
//             LSTM input projection (gc3_basics.py:22, x_t W_ih^T), ProjRNN.proj (:23), the 1x1 convs of
//             gc3_network.py:55,99 / dprnn.py:85, the encoder/decoder framing matmuls (:49,105) and all of
//             their input-gradients.
//   gemm_tn : dW[Mo,No] += A[P,Mo]^T * B[P,No]     weight gradients (reduction over positions).
//
// Arithmetic: inputs are fp32 in HBM; each operand is split on the fly into bf16 hi + lo and the product is
// formed as hi*hi + hi*lo + lo*hi on the tensor cores with fp32 accumulation ("bf16x3", SPLIT = true), which
// keeps the model inside the fp32 parity gate (rel-L2 1e-4; measured 1.4e-5).  SPLIT = false is the bf16 mode.
// These kernels are HBM-bound at these widths (AI ~ 60 FLOP/B, SURVEY 7 hard part 2): A is read once with
// 128-bit loads, W tiles come from L2, C is written once.
#include "common.cuh"
#include "kernels.h"

namespace dp {

namespace {

constexpr int BM = 128, BN = 64, BK = 32;
constexpr int AST = BK + 8;   // A smem row stride (bf16): 80 B -> conflict-free ldmatrix
constexpr int BST_NK = BK + 8;  // W stored [N,K]
constexpr int BST_KN = BN + 8;  // W stored [K,N]: 144 B rows

template <bool WKN, bool SPLIT>
__global__ void __launch_bounds__(256) gemm_nt_kernel(const GemmNtArgs p) {
    __shared__ __align__(16) __nv_bfloat16 As[2][BM * AST];
    __shared__ __align__(16) __nv_bfloat16 Bs[2][BN * BST_NK > BK * BST_KN ? BN * BST_NK : BK * BST_KN];
    __shared__ double red[8][2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

    float4 areg[4];
    uint4 bhi = make_uint4(0, 0, 0, 0), blo = make_uint4(0, 0, 0, 0);

    auto load_tile = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int row = m0 + (tid >> 3) + 32 * i;
            int k = k0 + (tid & 7) * 4;
            if (row < p.M && k < p.K) {
                long long r = row;
                if (p.a_rpb) r += (long long)(row / p.a_rpb) * p.a_skip;
                areg[i] = *reinterpret_cast<const float4*>(p.A + r * p.lda + k);
            } else {
                areg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (!WKN) {
            int n = n0 + (tid >> 2), k = k0 + (tid & 3) * 8;
            bool ok = (n < p.N) && (k < p.K);
            bhi = ok ? *reinterpret_cast<const uint4*>(p.Whi + (size_t)n * p.ldw + k) : make_uint4(0, 0, 0, 0);
            if (SPLIT) blo = ok ? *reinterpret_cast<const uint4*>(p.Wlo + (size_t)n * p.ldw + k) : make_uint4(0, 0, 0, 0);
        } else {
            int k = k0 + (tid >> 3), n = n0 + (tid & 7) * 8;
            bool ok = (n < p.N) && (k < p.K);
            bhi = ok ? *reinterpret_cast<const uint4*>(p.Whi + (size_t)k * p.ldw + n) : make_uint4(0, 0, 0, 0);
            if (SPLIT) blo = ok ? *reinterpret_cast<const uint4*>(p.Wlo + (size_t)k * p.ldw + n) : make_uint4(0, 0, 0, 0);
        }
    };
    auto store_tile = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int r = (tid >> 3) + 32 * i, kc = (tid & 7) * 4;
            uint2 hi, lo;
            if (p.relu_a) {
                areg[i].x = fmaxf(areg[i].x, 0.f); areg[i].y = fmaxf(areg[i].y, 0.f);
                areg[i].z = fmaxf(areg[i].z, 0.f); areg[i].w = fmaxf(areg[i].w, 0.f);
            }
            split_pair(areg[i].x, areg[i].y, hi.x, lo.x);
            split_pair(areg[i].z, areg[i].w, hi.y, lo.y);
            *reinterpret_cast<uint2*>(&As[0][r * AST + kc]) = hi;
            if (SPLIT) *reinterpret_cast<uint2*>(&As[1][r * AST + kc]) = lo;
        }
        if (!WKN) {
            int off = (tid >> 2) * BST_NK + (tid & 3) * 8;
            *reinterpret_cast<uint4*>(&Bs[0][off]) = bhi;
            if (SPLIT) *reinterpret_cast<uint4*>(&Bs[1][off]) = blo;
        } else {
            int off = (tid >> 3) * BST_KN + (tid & 7) * 8;
            *reinterpret_cast<uint4*>(&Bs[0][off]) = bhi;
            if (SPLIT) *reinterpret_cast<uint4*>(&Bs[1][off]) = blo;
        }
    };

    const int nk = ceil_div(p.K, BK);
    load_tile(0);
    for (int kt = 0; kt < nk; ++kt) {
        store_tile();
        __syncthreads();
        if (kt + 1 < nk) load_tile((kt + 1) * BK);  // in flight during the MMAs below
#pragma unroll
        for (int kk = 0; kk < BK; kk += 16) {
            uint32_t ahi[2][4], alo[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                int off = (wm * 32 + mt * 16 + (lane & 15)) * AST + kk + (lane >> 4) * 8;
                ldmatrix_x4(ahi[mt], smem_u32(&As[0][off]));
                if (SPLIT) ldmatrix_x4(alo[mt], smem_u32(&As[1][off]));
            }
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                uint32_t r[4];
                int off;
                if (!WKN)
                    off = (wn * 32 + np * 16 + (lane & 7) + (lane >> 4) * 8) * BST_NK + kk + ((lane >> 3) & 1) * 8;
                else
                    off = (kk + (lane & 7) + ((lane >> 3) & 1) * 8) * BST_KN + wn * 32 + np * 16 + (lane >> 4) * 8;
                if (!WKN) ldmatrix_x4(r, smem_u32(&Bs[0][off])); else ldmatrix_x4_trans(r, smem_u32(&Bs[0][off]));
                bh[2 * np][0] = r[0]; bh[2 * np][1] = r[1]; bh[2 * np + 1][0] = r[2]; bh[2 * np + 1][1] = r[3];
                if (SPLIT) {
                    if (!WKN) ldmatrix_x4(r, smem_u32(&Bs[1][off])); else ldmatrix_x4_trans(r, smem_u32(&Bs[1][off]));
                    bl[2 * np][0] = r[0]; bl[2 * np][1] = r[1]; bl[2 * np + 1][0] = r[2]; bl[2 * np + 1][1] = r[3];
                }
            }
            // one sweep per split product so that consecutive HMMAs hit different accumulators
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[mt][nt], ahi[mt], bh[nt]);
            if (SPLIT) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[mt][nt], ahi[mt], bl[nt]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[mt][nt], alo[mt], bh[nt]);
            }
        }
        __syncthreads();
    }

    // ---- epilogue ----
    const bool want_stats = p.stats != nullptr;
    int g_first = 0, g_last = 0;
    if (want_stats) {
        g_first = m0 / p.rows_per_group;
        g_last = (min(m0 + BM, p.M) - 1) / p.rows_per_group;
    }
    const bool uniform = (g_first == g_last);
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            int row = m0 + wm * 32 + mt * 16 + (lane >> 2) + half * 8;
            float r1 = 0.f, r2 = 0.f;
            if (row < p.M) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    int col = n0 + wn * 32 + nt * 8 + (lane & 3) * 2;
                    if (col < p.N) {
                        float x0 = acc[mt][nt][half * 2], x1 = acc[mt][nt][half * 2 + 1];
                        if (p.bias) {
                            x0 = fmaf(p.bias_scale, p.bias[col], x0);
                            x1 = fmaf(p.bias_scale, p.bias[col + 1], x1);
                        }
                        float2* dst = reinterpret_cast<float2*>(p.C + (size_t)row * p.ldc + col);
                        if (p.accumulate) {
                            float2 o = *dst;
                            x0 += o.x; x1 += o.y;
                        }
                        if (p.relu == 1) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                        else if (p.relu == 2) { x0 = tanhf(x0); x1 = tanhf(x1); }
                        else if (p.relu == 3) { x0 = 1.0f / (1.0f + expf(-x0)); x1 = 1.0f / (1.0f + expf(-x1)); }
                        if (p.mul_c) {
                            float2 o = *dst;
                            x0 *= o.x; x1 *= o.y;
                        }
                        if (p.mask) {
                            const float2 mk = *reinterpret_cast<const float2*>(p.mask + (size_t)row * p.ldmask + col);
                            x0 = mk.x > 0.f ? x0 : 0.f;
                            x1 = mk.y > 0.f ? x1 : 0.f;
                        }
                        *dst = make_float2(x0, x1);
                        r1 += x0 + x1;
                        r2 = fmaf(x0, x0, fmaf(x1, x1, r2));
                    }
                }
                if (want_stats && !uniform) {  // tile straddles two groups (rare): per-row atomics
                    int g = row / p.rows_per_group;
                    atomicAdd(&p.stats[2 * g], (double)r1);
                    atomicAdd(&p.stats[2 * g + 1], (double)r2);
                }
            }
            s1 += r1; s2 += r2;
        }
    if (want_stats && uniform) {
        s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
        if (lane == 0) { red[warp][0] = s1; red[warp][1] = s2; }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < 8; ++w) { a += red[w][0]; b += red[w][1]; }
            atomicAdd(&p.stats[2 * g_first], a);
            atomicAdd(&p.stats[2 * g_first + 1], b);
        }
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int TM = 64, TN = 64, TK = 32, TST = 64 + 8;
constexpr int TN_PCHUNK = 1024;

template <bool SPLIT>
__global__ void __launch_bounds__(256) gemm_tn_kernel(const GemmTnArgs p) {
    __shared__ __align__(16) __nv_bfloat16 As[2][TK * TST];
    __shared__ __align__(16) __nv_bfloat16 Bs[2][TK * TST];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const int pbeg = blockIdx.x * TN_PCHUNK, pend = min(p.P, pbeg + TN_PCHUNK);
    const int mo0 = blockIdx.y * TM, no0 = blockIdx.z * TN;

    float acc[2][2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

    float4 ar[2], br[2];
    auto load_tile = [&](int p0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int pr = p0 + (tid >> 4) + 16 * i;
            int c = (tid & 15) * 4;
            ar[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            br[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pr < pend) {
                if (mo0 + c < p.Mo) ar[i] = *reinterpret_cast<const float4*>(p.A + (size_t)pr * p.lda + mo0 + c);
                if (no0 + c < p.No) {
                    long long q = pr;
                    bool ok = true;
                    if (p.tmod) {
                        int t = (pr / p.tdiv) % p.tmod;
                        ok = p.shift < 0 ? (t > 0) : (t < p.tmod - 1);
                        q += p.shift;
                    }
                    if (p.b_rpb) q += (long long)(pr / p.b_rpb) * p.b_skip;
                    if (ok) br[i] = *reinterpret_cast<const float4*>(p.B + q * p.ldb + no0 + c);
                    if (p.relu_b) {
                        br[i].x = fmaxf(br[i].x, 0.f); br[i].y = fmaxf(br[i].y, 0.f);
                        br[i].z = fmaxf(br[i].z, 0.f); br[i].w = fmaxf(br[i].w, 0.f);
                    }
                }
            }
        }
    };
    auto store_tile = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int off = ((tid >> 4) + 16 * i) * TST + (tid & 15) * 4;
            uint2 hi, lo;
            split_pair(ar[i].x, ar[i].y, hi.x, lo.x);
            split_pair(ar[i].z, ar[i].w, hi.y, lo.y);
            *reinterpret_cast<uint2*>(&As[0][off]) = hi;
            if (SPLIT) *reinterpret_cast<uint2*>(&As[1][off]) = lo;
            split_pair(br[i].x, br[i].y, hi.x, lo.x);
            split_pair(br[i].z, br[i].w, hi.y, lo.y);
            *reinterpret_cast<uint2*>(&Bs[0][off]) = hi;
            if (SPLIT) *reinterpret_cast<uint2*>(&Bs[1][off]) = lo;
        }
    };

    load_tile(pbeg);
    for (int p0 = pbeg; p0 < pend; p0 += TK) {
        store_tile();
        __syncthreads();
        if (p0 + TK < pend) load_tile(p0 + TK);
#pragma unroll
        for (int kk = 0; kk < TK; kk += 16) {
            uint32_t ahi[2][4], alo[2][4], bh[2][2], bl[2][2];
            const int mi = lane >> 3;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                int off = (kk + (lane & 7) + (mi >> 1) * 8) * TST + wm * 32 + mt * 16 + (mi & 1) * 8;
                ldmatrix_x4_trans(ahi[mt], smem_u32(&As[0][off]));
                if (SPLIT) ldmatrix_x4_trans(alo[mt], smem_u32(&As[1][off]));
            }
            {
                uint32_t r[4];
                int off = (kk + (lane & 7) + ((lane >> 3) & 1) * 8) * TST + wn * 16 + (lane >> 4) * 8;
                ldmatrix_x4_trans(r, smem_u32(&Bs[0][off]));
                bh[0][0] = r[0]; bh[0][1] = r[1]; bh[1][0] = r[2]; bh[1][1] = r[3];
                if (SPLIT) {
                    ldmatrix_x4_trans(r, smem_u32(&Bs[1][off]));
                    bl[0][0] = r[0]; bl[0][1] = r[1]; bl[1][0] = r[2]; bl[1][1] = r[3];
                }
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) mma_bf16(acc[mt][nt], ahi[mt], bh[nt]);
            if (SPLIT) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) mma_bf16(acc[mt][nt], ahi[mt], bl[nt]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) mma_bf16(acc[mt][nt], alo[mt], bh[nt]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int row = mo0 + wm * 32 + mt * 16 + (lane >> 2) + (e >> 1) * 8;
                int col = no0 + wn * 16 + nt * 8 + (lane & 3) * 2 + (e & 1);
                if (row < p.Mo && col < p.No) atomicAdd(p.C + (size_t)row * p.ldc + col, acc[mt][nt][e] * p.scale);
            }
}

constexpr int CS_ROWS = 256;
// out[n] += scale * sum_p A[p, n]: 256 threads = (256 / C4) row lanes x C4 float4 columns; coalesced 128-bit loads,
// shared-memory reduction over the row lanes, one atomic per column per CTA.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int lda, int P, int N, float scale,
                                                     float* out, float* out2) {
    __shared__ float4 sh[256];
    const int C4 = N >> 2, lanes = 256 / C4;
    const int c4 = threadIdx.x % C4, rl = threadIdx.x / C4;
    const int pbeg = blockIdx.x * CS_ROWS, pend = min(P, pbeg + CS_ROWS);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl < lanes)
        for (int r = pbeg + rl; r < pend; r += lanes) {
            float4 v = ldg_stream(reinterpret_cast<const float4*>(A + (size_t)r * lda) + c4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    sh[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < C4) {
        float4 t = sh[threadIdx.x];
        for (int l = 1; l < lanes; ++l) {
            float4 v = sh[l * C4 + threadIdx.x];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        float* o = out + threadIdx.x * 4;
        atomicAdd(o, t.x * scale); atomicAdd(o + 1, t.y * scale); atomicAdd(o + 2, t.z * scale); atomicAdd(o + 3, t.w * scale);
        if (out2) {
            o = out2 + threadIdx.x * 4;
            atomicAdd(o, t.x * scale); atomicAdd(o + 1, t.y * scale); atomicAdd(o + 2, t.z * scale); atomicAdd(o + 3, t.w * scale);
        }
    }
}

}  // namespace

cudaError_t launch_gemm_nt(const GemmNtArgs& a, bool split, cudaStream_t st) {
    if (a.M <= 0) return cudaSuccess;
    if ((a.K & 3) || (a.N & 1) || (a.lda & 3) || (a.ldw & 7) || (a.ldc & 1)) return cudaErrorInvalidValue;
    dim3 grid(ceil_div(a.M, BM), ceil_div(a.N, BN));
    if (a.w_kn) {
        if (split) gemm_nt_kernel<true, true><<<grid, 256, 0, st>>>(a); else gemm_nt_kernel<true, false><<<grid, 256, 0, st>>>(a);
    } else {
        if (split) gemm_nt_kernel<false, true><<<grid, 256, 0, st>>>(a); else gemm_nt_kernel<false, false><<<grid, 256, 0, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_gemm_tn(const GemmTnArgs& a, bool split, cudaStream_t st) {
    if (a.P <= 0) return cudaSuccess;
    if ((a.lda & 3) || (a.ldb & 3) || (a.Mo & 3) || (a.No & 3)) return cudaErrorInvalidValue;
    dim3 grid(ceil_div(a.P, TN_PCHUNK), ceil_div(a.Mo, TM), ceil_div(a.No, TN));
    if (split) gemm_tn_kernel<true><<<grid, 256, 0, st>>>(a); else gemm_tn_kernel<false><<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_colsum(const float* A, int lda, int P, int N, float scale, float* out, float* out2, cudaStream_t st) {
    if (P <= 0) return cudaSuccess;
    if ((N & 3) || N > 1024 || 256 % (N / 4) || (lda & 3)) return cudaErrorInvalidValue;
    colsum_kernel<<<ceil_div(P, CS_ROWS), 256, 0, st>>>(A, lda, P, N, scale, out, out2);
    return cudaGetLastError();
}

}  // namespace dp
