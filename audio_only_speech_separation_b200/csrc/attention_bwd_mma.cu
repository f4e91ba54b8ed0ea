// Self-attention backward on the warp-level tensor cores (sm_100a), for sequences of up to 320 positions and head widths 16 / 32.
//
// Replaces autograd through nn.MultiheadAttention's core (look2hear/models/utils/dptnet.py:48, sepformer.py:124-133: bmm +
// softmax + bmm with the [B*S*h, L, L] probabilities materialised) for the dual-path layouts.  One CTA = one (sequence, head):
// Q, K, V and dO of that head live in shared memory as bf16 hi/lo rows; the probabilities are recomputed from the saved
// log-sum-exp, never stored.  Two sweeps over 16-row tiles distributed over the 8 warps:
//   A (tile = 16 queries):  S = Q K^T, dP = dO V^T per block of 64 keys;  P = 2^(S c - lse);  dS = P (dP - delta);  dQ += dS K
//   B (tile = 16 keys):     S^T = K Q^T, dP^T = V dO^T per block of 64 queries;  dV += P^T dO;  dK += dS^T Q
// S / dP accumulators become the A operand of the second product in registers (m16n8k16 C-fragment pairs = A fragment); K, Q
// and dO serve as "k-major" B operands through ldmatrix.trans.  SPLIT = bf16x3 products (fp32-parity mode), otherwise single bf16.
// The CUDA-core kernel of transformer.cu (exact fp32) remains the C-ABI operator's default and the fallback for longer sequences.
#include "common.cuh"
#include "kernels.h"

namespace dp {
namespace {

constexpr float kLog2eB = 1.4426950408889634f;
constexpr int LMAX = 320;   // shared memory: 8 planes x 320 rows x 80 B (head width 32, fp32-parity mode) = 205 KB

__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

template <bool SPLIT>
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const uint32_t (&bh)[2],
                                     const uint32_t (&bl)[2]) {
    mma_bf16(d, ah, bh);
    if (SPLIT) {
        mma_bf16(d, ah, bl);
        mma_bf16(d, al, bh);
    }
}

// B fragment (k16 x n8) of X^T for X stored [n][k] rows (k contiguous): plain ldmatrix
__device__ __forceinline__ void load_b_rows(const __nv_bfloat16* base, int stride, int n0, int k0, int lane, uint32_t (&b)[2]) {
    ldmatrix_x2(b, smem_u32(base + (n0 + (lane & 7)) * stride + k0 + ((lane >> 3) & 1) * 8));
}
// B fragment (k16 x n8) of X for X stored [k][n] rows (n contiguous): transposing ldmatrix
__device__ __forceinline__ void load_b_cols(const __nv_bfloat16* base, int stride, int k0, int n0, int lane, uint32_t (&b)[2]) {
    ldmatrix_x2_trans(b, smem_u32(base + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * stride + n0));
}
// A fragment (m16 x k16) for X stored [m][k] rows
__device__ __forceinline__ void load_a_rows(const __nv_bfloat16* base, int stride, int m0, int k0, int lane, uint32_t (&a)[4]) {
    ldmatrix_x4(a, smem_u32(base + (m0 + (lane & 7) + ((lane >> 3) & 1) * 8) * stride + k0 + (lane >> 4) * 8));
}

template <int D, bool SPLIT, bool DROP>
__global__ void __launch_bounds__(256, ((D == 16 || !SPLIT) ? 2 : 1)) attn_bwd_mma_kernel(const float* __restrict__ QKV, const float* __restrict__ O,
                                                              const float* __restrict__ LSE, const float* __restrict__ dO,
                                                              float* __restrict__ dQKV, int E, int heads, SeqMap m, float scale,
                                                              const unsigned drop_thr, const unsigned drop_key, const float drop_scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int RS = D + 8;      // bf16 row stride: 16-byte aligned rows, conflict-free ldmatrix
    constexpr int KS = D / 16;     // k-steps over the head width
    constexpr int DN = D / 8;      // n-tiles over the head width
    const int L = m.len;
    const int LP = (L + 63) & ~63;  // rows allocated / visited (multiple of the 64-wide blocks)
    // hi planes of Q, K, V, dO, then (fp32-parity mode only) their lo planes: bf16 mode needs half the shared memory, two CTAs per SM
    __nv_bfloat16* Qh = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* Kh = Qh + LP * RS;
    __nv_bfloat16* Vh = Kh + LP * RS;
    __nv_bfloat16* Gh = Vh + LP * RS;
    __nv_bfloat16* Ql = Gh + LP * RS;
    __nv_bfloat16* Kl = SPLIT ? Ql + LP * RS : Ql;
    __nv_bfloat16* Vl = SPLIT ? Kl + LP * RS : Ql;
    __nv_bfloat16* Gl = SPLIT ? Vl + LP * RS : Ql;
    float* lse = reinterpret_cast<float*>(SPLIT ? Gl + LP * RS : Ql);   // [LP]  +big for padded rows: their probabilities vanish
    float* dlt = lse + LP;                                 // [LP]  delta_i = dO_i . O_i
    float* kbias = dlt + LP;                               // [LP]  0 for real keys, -big for padded ones

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int q = blockIdx.x, h = blockIdx.y;
    const long long base = (long long)(q / m.qdiv) * m.s_hi + (long long)(q % m.qdiv) * m.s_lo;
    const int ld = 3 * E;
    const float c2 = scale * kLog2eB;

    // ---- stage Q, K, V, dO as bf16 hi / lo rows (zeros beyond L), delta and lse
    for (int idx = tid; idx < LP * (D / 4); idx += 256) {
        const int j = idx / (D / 4), cc = idx % (D / 4);
        float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), kv = qv, vv = qv, gv = qv;
        if (j < L) {
            const long long p = base + (long long)j * m.s_t;
            const float* row = QKV + p * ld + h * D + cc * 4;
            qv = *reinterpret_cast<const float4*>(row);
            kv = *reinterpret_cast<const float4*>(row + E);
            vv = *reinterpret_cast<const float4*>(row + 2 * E);
            gv = *reinterpret_cast<const float4*>(dO + p * E + h * D + cc * 4);
        }
        uint2 hi, lo;
        const int o = j * RS + cc * 4;
        split_pair(qv.x, qv.y, hi.x, lo.x); split_pair(qv.z, qv.w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(Qh + o) = hi; if (SPLIT) *reinterpret_cast<uint2*>(Ql + o) = lo;
        split_pair(kv.x, kv.y, hi.x, lo.x); split_pair(kv.z, kv.w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(Kh + o) = hi; if (SPLIT) *reinterpret_cast<uint2*>(Kl + o) = lo;
        split_pair(vv.x, vv.y, hi.x, lo.x); split_pair(vv.z, vv.w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(Vh + o) = hi; if (SPLIT) *reinterpret_cast<uint2*>(Vl + o) = lo;
        split_pair(gv.x, gv.y, hi.x, lo.x); split_pair(gv.z, gv.w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(Gh + o) = hi; if (SPLIT) *reinterpret_cast<uint2*>(Gl + o) = lo;
    }
    for (int i = tid; i < LP; i += 256) {
        float a = 0.f, l = 1e30f;
        if (i < L) {
            const long long p = base + (long long)i * m.s_t;
            const float4* orow = reinterpret_cast<const float4*>(O + p * E + h * D);
            const float4* grow = reinterpret_cast<const float4*>(dO + p * E + h * D);
#pragma unroll
            for (int cc = 0; cc < D / 4; ++cc) {
                const float4 ov = orow[cc], gv = grow[cc];
                a = fmaf(ov.x, gv.x, a); a = fmaf(ov.y, gv.y, a); a = fmaf(ov.z, gv.z, a); a = fmaf(ov.w, gv.w, a);
            }
            l = LSE[p * heads + h];
        }
        dlt[i] = a;
        lse[i] = l;
        kbias[i] = i < L ? 0.f : -1e30f;
    }
    __syncthreads();

    const int ntile = (L + 15) >> 4;
    // ================================================================ sweep A: dQ
    for (int mt = warp; mt < ntile; mt += 8) {
        const int r0 = mt * 16;
        uint32_t qa_h[KS][4], qa_l[KS][4], ga_h[KS][4], ga_l[KS][4];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            load_a_rows(Qh, RS, r0, ks * 16, lane, qa_h[ks]);
            load_a_rows(Gh, RS, r0, ks * 16, lane, ga_h[ks]);
            if (SPLIT) {
                load_a_rows(Ql, RS, r0, ks * 16, lane, qa_l[ks]);
                load_a_rows(Gl, RS, r0, ks * 16, lane, ga_l[ks]);
            }
        }
        const float l0 = lse[r0 + g], l1 = lse[r0 + g + 8], d0 = dlt[r0 + g], d1 = dlt[r0 + g + 8];
        // dropout mask rows of this thread's two queries: (stream position * heads + head)
        const uint32_t drow0 = (uint32_t)(base + (long long)(r0 + g) * m.s_t) * (uint32_t)heads + (uint32_t)h;
        const uint32_t drow1 = (uint32_t)(base + (long long)(r0 + g + 8) * m.s_t) * (uint32_t)heads + (uint32_t)h;
        float dq[DN][4];
#pragma unroll
        for (int n = 0; n < DN; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) dq[n][e] = 0.f;
        for (int j0 = 0; j0 < LP; j0 += 64) {
            float s[8][4], dp[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { s[n][e] = 0.f; dp[n][e] = 0.f; }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    uint32_t bh[2], bl[2];
                    load_b_rows(Kh, RS, j0 + n * 8, ks * 16, lane, bh);
                    if (SPLIT) load_b_rows(Kl, RS, j0 + n * 8, ks * 16, lane, bl);
                    mma3<SPLIT>(s[n], qa_h[ks], qa_l[ks], bh, bl);
                    load_b_rows(Vh, RS, j0 + n * 8, ks * 16, lane, bh);
                    if (SPLIT) load_b_rows(Vl, RS, j0 + n * 8, ks * 16, lane, bl);
                    mma3<SPLIT>(dp[n], ga_h[ks], ga_l[ks], bh, bl);
                }
            }
            // dS = P (dP - delta), in place of S
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float kb0 = kbias[j0 + n * 8 + 2 * c], kb1 = kbias[j0 + n * 8 + 2 * c + 1];
                if (DROP) {  // dP = (dO V^T) masked like the forward's probabilities; the softmax backward uses the un-dropped P
                    const uint32_t jc = (uint32_t)(j0 + n * 8 + 2 * c);
                    dp[n][0] = drop_keep(drop_key, drow0, jc, drop_thr) ? dp[n][0] * drop_scale : 0.f;
                    dp[n][1] = drop_keep(drop_key, drow0, jc + 1, drop_thr) ? dp[n][1] * drop_scale : 0.f;
                    dp[n][2] = drop_keep(drop_key, drow1, jc, drop_thr) ? dp[n][2] * drop_scale : 0.f;
                    dp[n][3] = drop_keep(drop_key, drow1, jc + 1, drop_thr) ? dp[n][3] * drop_scale : 0.f;
                }
                s[n][0] = ex2_approx(fmaf(s[n][0], c2, kb0 - l0)) * (dp[n][0] - d0);
                s[n][1] = ex2_approx(fmaf(s[n][1], c2, kb1 - l0)) * (dp[n][1] - d0);
                s[n][2] = ex2_approx(fmaf(s[n][2], c2, kb0 - l1)) * (dp[n][2] - d1);
                s[n][3] = ex2_approx(fmaf(s[n][3], c2, kb1 - l1)) * (dp[n][3] - d1);
            }
            // dQ += dS K_block  (contraction over the 64 keys of the block)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t ah[4], al[4];
                split_pair(s[2 * kk][0], s[2 * kk][1], ah[0], al[0]);
                split_pair(s[2 * kk][2], s[2 * kk][3], ah[1], al[1]);
                split_pair(s[2 * kk + 1][0], s[2 * kk + 1][1], ah[2], al[2]);
                split_pair(s[2 * kk + 1][2], s[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
                for (int n = 0; n < DN; ++n) {
                    uint32_t bh[2], bl[2];
                    load_b_cols(Kh, RS, j0 + kk * 16, n * 8, lane, bh);
                    if (SPLIT) load_b_cols(Kl, RS, j0 + kk * 16, n * 8, lane, bl);
                    mma3<SPLIT>(dq[n], ah, al, bh, bl);
                }
            }
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int i = r0 + g + 8 * hh;
            if (i < L) {
                float* dst = dQKV + (base + (long long)i * m.s_t) * ld + h * D + 2 * c;
#pragma unroll
                for (int n = 0; n < DN; ++n) *reinterpret_cast<float2*>(dst + n * 8) = make_float2(dq[n][2 * hh] * scale, dq[n][2 * hh + 1] * scale);
            }
        }
    }
    // ================================================================ sweep B: dK, dV (rows = keys, columns = queries)
    for (int mt = warp; mt < ntile; mt += 8) {
        const int r0 = mt * 16;
        uint32_t ka_h[KS][4], ka_l[KS][4], va_h[KS][4], va_l[KS][4];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            load_a_rows(Kh, RS, r0, ks * 16, lane, ka_h[ks]);
            load_a_rows(Vh, RS, r0, ks * 16, lane, va_h[ks]);
            if (SPLIT) {
                load_a_rows(Kl, RS, r0, ks * 16, lane, ka_l[ks]);
                load_a_rows(Vl, RS, r0, ks * 16, lane, va_l[ks]);
            }
        }
        float dk[DN][4], dv[DN][4];
#pragma unroll
        for (int n = 0; n < DN; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) { dk[n][e] = 0.f; dv[n][e] = 0.f; }
        for (int i0 = 0; i0 < LP; i0 += 64) {
            float s[8][4], dp[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { s[n][e] = 0.f; dp[n][e] = 0.f; }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    uint32_t bh[2], bl[2];
                    load_b_rows(Qh, RS, i0 + n * 8, ks * 16, lane, bh);
                    if (SPLIT) load_b_rows(Ql, RS, i0 + n * 8, ks * 16, lane, bl);
                    mma3<SPLIT>(s[n], ka_h[ks], ka_l[ks], bh, bl);
                    load_b_rows(Gh, RS, i0 + n * 8, ks * 16, lane, bh);
                    if (SPLIT) load_b_rows(Gl, RS, i0 + n * 8, ks * 16, lane, bl);
                    mma3<SPLIT>(dp[n], va_h[ks], va_l[ks], bh, bl);
                }
            }
            // P^T in s, dS^T in dp (columns are queries: their lse / delta; padded queries have lse = +big)
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int i = i0 + n * 8 + 2 * c;
                const float la = lse[i], lb = lse[i + 1], da = dlt[i], db = dlt[i + 1];
                float m0 = 1.f, m1 = 1.f, m2 = 1.f, m3 = 1.f;   // mask * 1/(1-p) of (query = column, key = row)
                if (DROP) {
                    const uint32_t qa = (uint32_t)(base + (long long)i * m.s_t) * (uint32_t)heads + (uint32_t)h;
                    const uint32_t qb = (uint32_t)(base + (long long)(i + 1) * m.s_t) * (uint32_t)heads + (uint32_t)h;
                    const uint32_t ka = (uint32_t)(r0 + g), kb = (uint32_t)(r0 + g + 8);
                    m0 = drop_keep(drop_key, qa, ka, drop_thr) ? drop_scale : 0.f;
                    m1 = drop_keep(drop_key, qb, ka, drop_thr) ? drop_scale : 0.f;
                    m2 = drop_keep(drop_key, qa, kb, drop_thr) ? drop_scale : 0.f;
                    m3 = drop_keep(drop_key, qb, kb, drop_thr) ? drop_scale : 0.f;
                }
                float pv;
                pv = ex2_approx(fmaf(s[n][0], c2, -la)); dp[n][0] = pv * (dp[n][0] * m0 - da); s[n][0] = pv * m0;
                pv = ex2_approx(fmaf(s[n][1], c2, -lb)); dp[n][1] = pv * (dp[n][1] * m1 - db); s[n][1] = pv * m1;
                pv = ex2_approx(fmaf(s[n][2], c2, -la)); dp[n][2] = pv * (dp[n][2] * m2 - da); s[n][2] = pv * m2;
                pv = ex2_approx(fmaf(s[n][3], c2, -lb)); dp[n][3] = pv * (dp[n][3] * m3 - db); s[n][3] = pv * m3;
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t ph[4], pl[4], sh[4], sl[4];
                split_pair(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
                split_pair(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
                split_pair(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
                split_pair(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
                split_pair(dp[2 * kk][0], dp[2 * kk][1], sh[0], sl[0]);
                split_pair(dp[2 * kk][2], dp[2 * kk][3], sh[1], sl[1]);
                split_pair(dp[2 * kk + 1][0], dp[2 * kk + 1][1], sh[2], sl[2]);
                split_pair(dp[2 * kk + 1][2], dp[2 * kk + 1][3], sh[3], sl[3]);
#pragma unroll
                for (int n = 0; n < DN; ++n) {
                    uint32_t bh[2], bl[2];
                    load_b_cols(Gh, RS, i0 + kk * 16, n * 8, lane, bh);
                    if (SPLIT) load_b_cols(Gl, RS, i0 + kk * 16, n * 8, lane, bl);
                    mma3<SPLIT>(dv[n], ph, pl, bh, bl);
                    load_b_cols(Qh, RS, i0 + kk * 16, n * 8, lane, bh);
                    if (SPLIT) load_b_cols(Ql, RS, i0 + kk * 16, n * 8, lane, bl);
                    mma3<SPLIT>(dk[n], sh, sl, bh, bl);
                }
            }
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int j = r0 + g + 8 * hh;
            if (j < L) {
                float* dst = dQKV + (base + (long long)j * m.s_t) * ld + h * D + 2 * c;
#pragma unroll
                for (int n = 0; n < DN; ++n) {
                    *reinterpret_cast<float2*>(dst + E + n * 8) = make_float2(dk[n][2 * hh] * scale, dk[n][2 * hh + 1] * scale);
                    *reinterpret_cast<float2*>(dst + 2 * E + n * 8) = make_float2(dv[n][2 * hh], dv[n][2 * hh + 1]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ forward (same machinery)
// softmax(q k^T / sqrt(d)) v for sequences the tcgen05 kernel does not take (257..320 positions; 16 s at 16 kHz gives 258 chunks).
// One CTA = one (sequence, head); K, V (all rows) and Q as bf16 hi / lo rows in shared memory; a warp owns 16-query tiles and walks
// the keys in blocks of 64 with an online softmax (running row maximum / sum per thread quad), P re-used from the accumulator
// fragments as the A operand of P V.  Outputs like the other forward kernels: O fp32 and / or planes, log-sum-exp (log2 domain).
template <int D, bool SPLIT, bool DROP>
__global__ void __launch_bounds__(256, (SPLIT ? 1 : 2)) attn_fwd_mma_kernel(const float* __restrict__ QKV, float* __restrict__ O,
                                                                            __nv_bfloat16* __restrict__ O_hi, __nv_bfloat16* __restrict__ O_lo,
                                                                            float* __restrict__ LSE, int E, int heads, SeqMap m, float scale,
                                                                            const unsigned drop_thr, const unsigned drop_key, const float drop_scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int RS = D + 8, KS = D / 16, DN = D / 8;
    const int L = m.len;
    const int LP = (L + 63) & ~63;
    __nv_bfloat16* Qh = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* Kh = Qh + LP * RS;
    __nv_bfloat16* Vh = Kh + LP * RS;
    __nv_bfloat16* Ql = Vh + LP * RS;
    __nv_bfloat16* Kl = SPLIT ? Ql + LP * RS : Ql;
    __nv_bfloat16* Vl = SPLIT ? Kl + LP * RS : Ql;
    float* kbias = reinterpret_cast<float*>(SPLIT ? Vl + LP * RS : Ql);   // [LP] 0 for real keys, -big for padded ones

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int q = blockIdx.x, h = blockIdx.y;
    const long long base = (long long)(q / m.qdiv) * m.s_hi + (long long)(q % m.qdiv) * m.s_lo;
    const int ld = 3 * E;
    const float c2 = scale * kLog2eB;
    for (int idx = tid; idx < LP * (D / 4); idx += 256) {
        const int j = idx / (D / 4), cc = idx % (D / 4);
        float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), kv = qv, vv = qv;
        if (j < L) {
            const float* row = QKV + (base + (long long)j * m.s_t) * ld + h * D + cc * 4;
            qv = *reinterpret_cast<const float4*>(row);
            kv = *reinterpret_cast<const float4*>(row + E);
            vv = *reinterpret_cast<const float4*>(row + 2 * E);
        }
        uint2 hi, lo;
        const int o = j * RS + cc * 4;
        split_pair(qv.x, qv.y, hi.x, lo.x); split_pair(qv.z, qv.w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(Qh + o) = hi; if (SPLIT) *reinterpret_cast<uint2*>(Ql + o) = lo;
        split_pair(kv.x, kv.y, hi.x, lo.x); split_pair(kv.z, kv.w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(Kh + o) = hi; if (SPLIT) *reinterpret_cast<uint2*>(Kl + o) = lo;
        split_pair(vv.x, vv.y, hi.x, lo.x); split_pair(vv.z, vv.w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(Vh + o) = hi; if (SPLIT) *reinterpret_cast<uint2*>(Vl + o) = lo;
    }
    for (int i = tid; i < LP; i += 256) kbias[i] = i < L ? 0.f : -1e30f;
    __syncthreads();

    const int ntile = (L + 15) >> 4;
    for (int mt = warp; mt < ntile; mt += 8) {
        const int r0 = mt * 16;
        uint32_t qa_h[KS][4], qa_l[KS][4];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            load_a_rows(Qh, RS, r0, ks * 16, lane, qa_h[ks]);
            if (SPLIT) load_a_rows(Ql, RS, r0, ks * 16, lane, qa_l[ks]);
        }
        const uint32_t drow0 = (uint32_t)(base + (long long)(r0 + g) * m.s_t) * (uint32_t)heads + (uint32_t)h;
        const uint32_t drow1 = (uint32_t)(base + (long long)(r0 + g + 8) * m.s_t) * (uint32_t)heads + (uint32_t)h;
        float mx0 = -1e30f, mx1 = -1e30f, l0 = 0.f, l1 = 0.f;   // running maxima (log2 domain) and this thread's partial row sums
        float o[DN][4];
#pragma unroll
        for (int n = 0; n < DN; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
        for (int j0 = 0; j0 < LP; j0 += 64) {
            float s[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
#pragma unroll
                for (int e = 0; e < 4; ++e) s[n][e] = 0.f;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    uint32_t bh[2], bl[2];
                    load_b_rows(Kh, RS, j0 + n * 8, ks * 16, lane, bh);
                    if (SPLIT) load_b_rows(Kl, RS, j0 + n * 8, ks * 16, lane, bl);
                    mma3<SPLIT>(s[n], qa_h[ks], qa_l[ks], bh, bl);
                }
            }
            float b0 = -1e30f, b1 = -1e30f;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float kb0 = kbias[j0 + n * 8 + 2 * c], kb1 = kbias[j0 + n * 8 + 2 * c + 1];
                s[n][0] = fmaf(s[n][0], c2, kb0); s[n][1] = fmaf(s[n][1], c2, kb1);
                s[n][2] = fmaf(s[n][2], c2, kb0); s[n][3] = fmaf(s[n][3], c2, kb1);
                b0 = fmaxf(b0, fmaxf(s[n][0], s[n][1])); b1 = fmaxf(b1, fmaxf(s[n][2], s[n][3]));
            }
            b0 = fmaxf(b0, __shfl_xor_sync(0xffffffffu, b0, 1)); b0 = fmaxf(b0, __shfl_xor_sync(0xffffffffu, b0, 2));
            b1 = fmaxf(b1, __shfl_xor_sync(0xffffffffu, b1, 1)); b1 = fmaxf(b1, __shfl_xor_sync(0xffffffffu, b1, 2));
            const float n0 = fmaxf(mx0, b0), n1 = fmaxf(mx1, b1);
            const float f0 = ex2_approx(mx0 - n0), f1 = ex2_approx(mx1 - n1);
            mx0 = n0; mx1 = n1;
            l0 *= f0; l1 *= f1;
#pragma unroll
            for (int n = 0; n < DN; ++n) { o[n][0] *= f0; o[n][1] *= f0; o[n][2] *= f1; o[n][3] *= f1; }
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                s[n][0] = ex2_approx(s[n][0] - mx0); s[n][1] = ex2_approx(s[n][1] - mx0);
                s[n][2] = ex2_approx(s[n][2] - mx1); s[n][3] = ex2_approx(s[n][3] - mx1);
                l0 += s[n][0] + s[n][1]; l1 += s[n][2] + s[n][3];   // denominators of the un-dropped probabilities
                if (DROP) {
                    const uint32_t jc = (uint32_t)(j0 + n * 8 + 2 * c);
                    if (!drop_keep(drop_key, drow0, jc, drop_thr)) s[n][0] = 0.f;
                    if (!drop_keep(drop_key, drow0, jc + 1, drop_thr)) s[n][1] = 0.f;
                    if (!drop_keep(drop_key, drow1, jc, drop_thr)) s[n][2] = 0.f;
                    if (!drop_keep(drop_key, drow1, jc + 1, drop_thr)) s[n][3] = 0.f;
                }
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t ah[4], al[4];
                split_pair(s[2 * kk][0], s[2 * kk][1], ah[0], al[0]);
                split_pair(s[2 * kk][2], s[2 * kk][3], ah[1], al[1]);
                split_pair(s[2 * kk + 1][0], s[2 * kk + 1][1], ah[2], al[2]);
                split_pair(s[2 * kk + 1][2], s[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
                for (int n = 0; n < DN; ++n) {
                    uint32_t bh[2], bl[2];
                    load_b_cols(Vh, RS, j0 + kk * 16, n * 8, lane, bh);
                    if (SPLIT) load_b_cols(Vl, RS, j0 + kk * 16, n * 8, lane, bl);
                    mma3<SPLIT>(o[n], ah, al, bh, bl);
                }
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int i = r0 + g + 8 * hh;
            if (i >= L) continue;
            const float lsum = hh ? l1 : l0, inv = (DROP ? drop_scale : 1.0f) / lsum;
            const size_t pos = (size_t)(base + (long long)i * m.s_t);
#pragma unroll
            for (int n = 0; n < DN; ++n) {
                const float v0 = o[n][2 * hh] * inv, v1 = o[n][2 * hh + 1] * inv;
                const size_t off = pos * E + h * D + n * 8 + 2 * c;
                if (O != nullptr) *reinterpret_cast<float2*>(O + off) = make_float2(v0, v1);
                if (O_hi != nullptr) {
                    uint32_t ph, pl;
                    split_pair(v0, v1, ph, pl);
                    *reinterpret_cast<uint32_t*>(O_hi + off) = ph;
                    if (O_lo != nullptr) *reinterpret_cast<uint32_t*>(O_lo + off) = pl;
                }
            }
            if (LSE != nullptr && c == 0) LSE[pos * heads + h] = (hh ? mx1 : mx0) + log2f(lsum);
        }
    }
}

template <int D, bool SPLIT>
cudaError_t launch_fwd_one(const float* QKV, float* O, __nv_bfloat16* O_hi, __nv_bfloat16* O_lo, float* LSE, int E, int heads, const SeqMap& m,
                           cudaStream_t st, unsigned drop_thr, unsigned drop_key, float drop_scale) {
    const int LP = (m.len + 63) & ~63;
    const size_t smem = (size_t)(SPLIT ? 6 : 3) * LP * (D + 8) * sizeof(__nv_bfloat16) + (size_t)LP * sizeof(float);
    const float scale = 1.0f / sqrtf((float)D);
    cudaError_t e;
    if (drop_thr) {
        e = cudaFuncSetAttribute(attn_fwd_mma_kernel<D, SPLIT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attn_fwd_mma_kernel<D, SPLIT, true><<<dim3(m.nseq, heads), 256, smem, st>>>(QKV, O, O_hi, O_lo, LSE, E, heads, m, scale, drop_thr, drop_key, drop_scale);
    } else {
        e = cudaFuncSetAttribute(attn_fwd_mma_kernel<D, SPLIT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attn_fwd_mma_kernel<D, SPLIT, false><<<dim3(m.nseq, heads), 256, smem, st>>>(QKV, O, O_hi, O_lo, LSE, E, heads, m, scale, 0u, 0u, 1.f);
    }
    return cudaGetLastError();
}

template <int D, bool SPLIT>
cudaError_t launch_one(const float* QKV, const float* O, const float* LSE, const float* dO, float* dQKV, int E, int heads, const SeqMap& m,
                       cudaStream_t st, unsigned drop_thr, unsigned drop_key, float drop_scale) {
    const int LP = (m.len + 63) & ~63;
    const size_t smem = (size_t)(SPLIT ? 8 : 4) * LP * (D + 8) * sizeof(__nv_bfloat16) + (size_t)3 * LP * sizeof(float);
    const float scale = 1.0f / sqrtf((float)D);
    cudaError_t e;
    if (drop_thr) {
        e = cudaFuncSetAttribute(attn_bwd_mma_kernel<D, SPLIT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attn_bwd_mma_kernel<D, SPLIT, true><<<dim3(m.nseq, heads), 256, smem, st>>>(QKV, O, LSE, dO, dQKV, E, heads, m, scale, drop_thr, drop_key, drop_scale);
    } else {
        e = cudaFuncSetAttribute(attn_bwd_mma_kernel<D, SPLIT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attn_bwd_mma_kernel<D, SPLIT, false><<<dim3(m.nseq, heads), 256, smem, st>>>(QKV, O, LSE, dO, dQKV, E, heads, m, scale, 0u, 0u, 1.f);
    }
    return cudaGetLastError();
}

}  // namespace

namespace {
int g_attn_fwd_mode = 0;   // 0 automatic, 1 tcgen05 kernel wherever it applies, 2 warp-level kernel wherever it applies
}
int attn_fwd_set_mode(int mode) {
    if (mode < 0 || mode > 2) return -1;
    g_attn_fwd_mode = mode;
    return 0;
}
// Measured on B200 (tests/tools/time_attention.py, forward, one layer): bf16 mode 86 / 90 us (warp-level) against 120 / 163 us (tcgen05)
// for SepFormer's 250 / 130-position sequences, 98 / 113 against 112 / 130 us for DPTNet; fp32-parity mode 159 against 136 us at 250
// positions (three split products on the legacy pipe), but 174 / 112 / 131 against 182 / 128 / 152 us for the shorter ones.
bool attn_fwd_prefers_mma(int E, int heads, int len, bool split) {
    SeqMap m{1, len, 1, 0, 0, 1};
    if (!attn_bwd_mma_supported(E, heads, m)) return false;
    if (g_attn_fwd_mode == 1) return false;
    if (g_attn_fwd_mode == 2) return true;
    return !split || len <= 192;
}

bool attn_bwd_mma_supported(int E, int heads, const SeqMap& m) {
    if (heads <= 0 || E % heads) return false;
    const int D = E / heads;
    return (D == 16 || D == 32) && m.len >= 1 && m.len <= LMAX;
}

cudaError_t launch_attn_bwd_mma(const float* QKV, const float* O, const float* LSE, const float* dO, float* dQKV, int E, int heads,
                                const SeqMap& m, bool split, cudaStream_t st, unsigned drop_thr, unsigned drop_key, float drop_scale) {
    if (m.nseq <= 0 || m.len <= 0) return cudaSuccess;
    if (!attn_bwd_mma_supported(E, heads, m)) return cudaErrorInvalidValue;
    const int D = E / heads;
#define DP_AB(DD, SP) launch_one<DD, SP>(QKV, O, LSE, dO, dQKV, E, heads, m, st, drop_thr, drop_key, drop_scale)
    if (D == 16) return split ? DP_AB(16, true) : DP_AB(16, false);
    return split ? DP_AB(32, true) : DP_AB(32, false);
#undef DP_AB
}

cudaError_t launch_attn_fwd_mma(const float* QKV, float* O, __nv_bfloat16* O_hi, __nv_bfloat16* O_lo, float* LSE, int E, int heads,
                                const SeqMap& m, bool split, cudaStream_t st, unsigned drop_thr, unsigned drop_key, float drop_scale) {
    if (m.nseq <= 0 || m.len <= 0) return cudaSuccess;
    if (!attn_bwd_mma_supported(E, heads, m)) return cudaErrorInvalidValue;
    const int D = E / heads;
#define DP_AF(DD, SP) launch_fwd_one<DD, SP>(QKV, O, O_hi, O_lo, LSE, E, heads, m, st, drop_thr, drop_key, drop_scale)
    if (D == 16) return split ? DP_AF(16, true) : DP_AF(16, false);
    return split ? DP_AF(32, true) : DP_AF(32, false);
#undef DP_AF
}

}  // namespace dp
