// Self-attention and LayerNorm kernels of the transformer dual-path blocks (sm_100a).
//
//   DPTNet   : nn.MultiheadAttention(64, 4 heads, d = 16), LayerNorm(64, eps 1e-5)        look2hear/models/utils/dptnet.py:48-82
//   SepFormer: nn.MultiheadAttention(256, 8 heads, d = 32), LayerNorm(256, eps 1e-6)      look2hear/models/sepformer.py:124-215,316-370
//
// Activations are channels-last rows addressed by position p; a sequence (q, t) lives at the position given by the
// SeqMap (intra-chunk sequences walk k, inter-chunk sequences walk s), so attention needs none of the reference's
// permute().contiguous() copies (dptnet.py:147-155, sepformer.py:618-637).
//
// Attention: the sequences are short (L = 82..258) and heads narrow (d = 16/32): K and V of one (sequence, head) sit in
// shared memory, one thread owns one query row and runs an exact fp32 streaming softmax (scores never touch HBM; the
// reference materialises [B*S*h, L, L] probabilities, sepformer.py:142).  The backward recomputes the probabilities
// from the saved log-sum-exp: pass A (thread = query) produces dQ, pass B (thread = key) produces dK and dV, so no
// atomics are needed and the result is deterministic.
#include "common.cuh"
#include "kernels.h"

namespace dp {
namespace {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ long long seq_base(const SeqMap& m, int q) {
    return (long long)(q / m.qdiv) * m.s_hi + (long long)(q % m.qdiv) * m.s_lo;
}

// ------------------------------------------------------------------------------------------------ attention forward
// grid (nseq, heads); QKV [P, 3E] = [q | k | v], head h uses columns [h*D, (h+1)*D) of each third.
template <int D>
__global__ void __launch_bounds__(256) attn_fwd_kernel(const float* __restrict__ QKV, float* __restrict__ O, float* __restrict__ LSE,
                                                       int E, int heads, SeqMap m, float scale_log2, __nv_bfloat16* __restrict__ O_hi,
                                                       __nv_bfloat16* __restrict__ O_lo, const unsigned drop_thr, const unsigned drop_key,
                                                       const float drop_scale) {
    extern __shared__ __align__(16) float sm_att[];
    constexpr int DS = D + 4;  // padded row stride
    const int L = m.len;
    float* Ks = sm_att;
    float* Vs = Ks + (size_t)L * DS;
    const int q = blockIdx.x, h = blockIdx.y;
    const long long base = seq_base(m, q);
    const int ld = 3 * E;
    for (int idx = threadIdx.x; idx < L * (D / 4); idx += blockDim.x) {
        const int j = idx / (D / 4), c = idx % (D / 4);
        const float* row = QKV + (base + (long long)j * m.s_t) * ld + h * D + c * 4;
        *reinterpret_cast<float4*>(Ks + j * DS + c * 4) = *reinterpret_cast<const float4*>(row + E);
        *reinterpret_cast<float4*>(Vs + j * DS + c * 4) = *reinterpret_cast<const float4*>(row + 2 * E);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const long long p = base + (long long)i * m.s_t;
        float qv[D], acc[D];
        {
            const float4* src = reinterpret_cast<const float4*>(QKV + p * ld + h * D);
#pragma unroll
            for (int c = 0; c < D / 4; ++c) {
                float4 v = src[c];
                qv[4 * c] = v.x * scale_log2; qv[4 * c + 1] = v.y * scale_log2; qv[4 * c + 2] = v.z * scale_log2; qv[4 * c + 3] = v.w * scale_log2;
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] = 0.f;
        float mx = -INFINITY, l = 0.f;
        for (int j0 = 0; j0 < L; j0 += 4) {
            float s[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                float a = 0.f;
                if (j < L) {
                    const float4* kr = reinterpret_cast<const float4*>(Ks + j * DS);
#pragma unroll
                    for (int c = 0; c < D / 4; ++c) {
                        float4 kv = kr[c];
                        a = fmaf(qv[4 * c], kv.x, a); a = fmaf(qv[4 * c + 1], kv.y, a); a = fmaf(qv[4 * c + 2], kv.z, a); a = fmaf(qv[4 * c + 3], kv.w, a);
                    }
                } else {
                    a = -INFINITY;
                }
                s[u] = a;
            }
            const float cm = fmaxf(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])), mx);
            const float corr = exp2f(mx - cm);  // first chunk: exp2(-inf) = 0
            l *= corr;
#pragma unroll
            for (int d = 0; d < D; ++d) acc[d] *= corr;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                if (j < L) {
                    float pj = exp2f(s[u] - cm);
                    l += pj;   // the denominator is that of the un-dropped probabilities
                    if (drop_thr && !drop_keep(drop_key, (uint32_t)p * (uint32_t)heads + (uint32_t)h, (uint32_t)j, drop_thr)) pj = 0.f;
                    const float4* vr = reinterpret_cast<const float4*>(Vs + j * DS);
#pragma unroll
                    for (int c = 0; c < D / 4; ++c) {
                        float4 vv = vr[c];
                        acc[4 * c] = fmaf(pj, vv.x, acc[4 * c]); acc[4 * c + 1] = fmaf(pj, vv.y, acc[4 * c + 1]);
                        acc[4 * c + 2] = fmaf(pj, vv.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(pj, vv.w, acc[4 * c + 3]);
                    }
                }
            }
            mx = cm;
        }
        const float inv = (drop_thr ? drop_scale : 1.0f) / l;
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] *= inv;
        if (O != nullptr) {
            float4* dst = reinterpret_cast<float4*>(O + p * E + h * D);
#pragma unroll
            for (int c = 0; c < D / 4; ++c) dst[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
        }
        if (O_hi != nullptr) {  // operand planes for the out-projection GEMM
            uint32_t hi[D / 2], lo[D / 2];
#pragma unroll
            for (int d = 0; d < D / 2; ++d) split_pair(acc[2 * d], acc[2 * d + 1], hi[d], lo[d]);
            uint4* dh = reinterpret_cast<uint4*>(O_hi + p * E + h * D);
#pragma unroll
            for (int c = 0; c < D / 8; ++c) dh[c] = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
            if (O_lo != nullptr) {
                uint4* dl = reinterpret_cast<uint4*>(O_lo + p * E + h * D);
#pragma unroll
                for (int c = 0; c < D / 8; ++c) dl[c] = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
            }
        }
        if (LSE) LSE[p * heads + h] = mx + log2f(l);  // log2 domain
    }
}

// ------------------------------------------------------------------------------------------------ attention backward
template <int D>
__global__ void __launch_bounds__(256) attn_bwd_kernel(const float* __restrict__ QKV, const float* __restrict__ O, const float* __restrict__ LSE,
                                                       const float* __restrict__ dO, float* __restrict__ dQKV, int E, int heads, SeqMap m,
                                                       float scale) {
    extern __shared__ __align__(16) float sm_att[];
    constexpr int DS = D + 4;
    const int L = m.len;
    float* Qs = sm_att;
    float* Ks = Qs + (size_t)L * DS;
    float* Vs = Ks + (size_t)L * DS;
    float* Gs = Vs + (size_t)L * DS;  // dO
    float* lse = Gs + (size_t)L * DS;
    float* dlt = lse + L;             // D_i = dO_i . O_i
    const int q = blockIdx.x, h = blockIdx.y;
    const long long base = seq_base(m, q);
    const int ld = 3 * E;
    const float scale_log2 = scale * kLog2e;
    for (int idx = threadIdx.x; idx < L * (D / 4); idx += blockDim.x) {
        const int j = idx / (D / 4), c = idx % (D / 4);
        const long long p = base + (long long)j * m.s_t;
        const float* row = QKV + p * ld + h * D + c * 4;
        *reinterpret_cast<float4*>(Qs + j * DS + c * 4) = *reinterpret_cast<const float4*>(row);
        *reinterpret_cast<float4*>(Ks + j * DS + c * 4) = *reinterpret_cast<const float4*>(row + E);
        *reinterpret_cast<float4*>(Vs + j * DS + c * 4) = *reinterpret_cast<const float4*>(row + 2 * E);
        *reinterpret_cast<float4*>(Gs + j * DS + c * 4) = *reinterpret_cast<const float4*>(dO + p * E + h * D + c * 4);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const long long p = base + (long long)i * m.s_t;
        const float4* orow = reinterpret_cast<const float4*>(O + p * E + h * D);
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < D / 4; ++c) {
            float4 ov = orow[c];
            float4 gv = *reinterpret_cast<const float4*>(Gs + i * DS + c * 4);
            a = fmaf(ov.x, gv.x, a); a = fmaf(ov.y, gv.y, a); a = fmaf(ov.z, gv.z, a); a = fmaf(ov.w, gv.w, a);
        }
        dlt[i] = a;
        lse[i] = LSE[p * heads + h];
    }
    __syncthreads();
    // ---- pass A: thread = query i -> dQ_i = scale * sum_j dS_ij K_j
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        float qv[D], gv[D], dq[D];
#pragma unroll
        for (int d = 0; d < D; ++d) { qv[d] = Qs[i * DS + d] * scale_log2; gv[d] = Gs[i * DS + d]; dq[d] = 0.f; }
        const float li = lse[i], di = dlt[i];
        for (int j = 0; j < L; ++j) {
            const float4* kr = reinterpret_cast<const float4*>(Ks + j * DS);
            const float4* vr = reinterpret_cast<const float4*>(Vs + j * DS);
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < D / 4; ++c) {
                float4 kv = kr[c], vv = vr[c];
                s = fmaf(qv[4 * c], kv.x, s); s = fmaf(qv[4 * c + 1], kv.y, s); s = fmaf(qv[4 * c + 2], kv.z, s); s = fmaf(qv[4 * c + 3], kv.w, s);
                dp = fmaf(gv[4 * c], vv.x, dp); dp = fmaf(gv[4 * c + 1], vv.y, dp); dp = fmaf(gv[4 * c + 2], vv.z, dp); dp = fmaf(gv[4 * c + 3], vv.w, dp);
            }
            const float ds = exp2f(s - li) * (dp - di);
#pragma unroll
            for (int c = 0; c < D / 4; ++c) {
                float4 kv = kr[c];
                dq[4 * c] = fmaf(ds, kv.x, dq[4 * c]); dq[4 * c + 1] = fmaf(ds, kv.y, dq[4 * c + 1]);
                dq[4 * c + 2] = fmaf(ds, kv.z, dq[4 * c + 2]); dq[4 * c + 3] = fmaf(ds, kv.w, dq[4 * c + 3]);
            }
        }
        float4* dst = reinterpret_cast<float4*>(dQKV + (base + (long long)i * m.s_t) * ld + h * D);
#pragma unroll
        for (int c = 0; c < D / 4; ++c) dst[c] = make_float4(dq[4 * c] * scale, dq[4 * c + 1] * scale, dq[4 * c + 2] * scale, dq[4 * c + 3] * scale);
    }
    // ---- pass B: thread = key j -> dK_j = scale * sum_i dS_ij Q_i ; dV_j = sum_i P_ij dO_i
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        float kv[D], vv[D], dk[D], dv[D];
#pragma unroll
        for (int d = 0; d < D; ++d) { kv[d] = Ks[j * DS + d] * scale_log2; vv[d] = Vs[j * DS + d]; dk[d] = 0.f; dv[d] = 0.f; }
        for (int i = 0; i < L; ++i) {
            const float4* qr = reinterpret_cast<const float4*>(Qs + i * DS);
            const float4* gr = reinterpret_cast<const float4*>(Gs + i * DS);
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < D / 4; ++c) {
                float4 qq = qr[c], gg = gr[c];
                s = fmaf(kv[4 * c], qq.x, s); s = fmaf(kv[4 * c + 1], qq.y, s); s = fmaf(kv[4 * c + 2], qq.z, s); s = fmaf(kv[4 * c + 3], qq.w, s);
                dp = fmaf(vv[4 * c], gg.x, dp); dp = fmaf(vv[4 * c + 1], gg.y, dp); dp = fmaf(vv[4 * c + 2], gg.z, dp); dp = fmaf(vv[4 * c + 3], gg.w, dp);
            }
            const float pij = exp2f(s - lse[i]);
            const float ds = pij * (dp - dlt[i]);
#pragma unroll
            for (int c = 0; c < D / 4; ++c) {
                float4 qq = qr[c], gg = gr[c];
                dk[4 * c] = fmaf(ds, qq.x, dk[4 * c]); dk[4 * c + 1] = fmaf(ds, qq.y, dk[4 * c + 1]);
                dk[4 * c + 2] = fmaf(ds, qq.z, dk[4 * c + 2]); dk[4 * c + 3] = fmaf(ds, qq.w, dk[4 * c + 3]);
                dv[4 * c] = fmaf(pij, gg.x, dv[4 * c]); dv[4 * c + 1] = fmaf(pij, gg.y, dv[4 * c + 1]);
                dv[4 * c + 2] = fmaf(pij, gg.z, dv[4 * c + 2]); dv[4 * c + 3] = fmaf(pij, gg.w, dv[4 * c + 3]);
            }
        }
        float* row = dQKV + (base + (long long)j * m.s_t) * ld + h * D;
        float4* dkd = reinterpret_cast<float4*>(row + E);
        float4* dvd = reinterpret_cast<float4*>(row + 2 * E);
#pragma unroll
        for (int c = 0; c < D / 4; ++c) {
            dkd[c] = make_float4(dk[4 * c] * scale, dk[4 * c + 1] * scale, dk[4 * c + 2] * scale, dk[4 * c + 3] * scale);
            dvd[c] = make_float4(dv[4 * c], dv[4 * c + 1], dv[4 * c + 2], dv[4 * c + 3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// z = a (+ b); y = LN_E(z) * gamma + beta; out = (res ? res + y : y) (+ unfold affine/PReLU); z optionally stored.
// LPR lanes per row (16 for E = 64, 32 otherwise), each lane owns CPL float4 chunks.
template <int E>
struct LnGeom {
    static constexpr int LPR = (E / 4 < 32) ? E / 4 : 32;
    static constexpr int CPL = E / (4 * LPR);
    static constexpr int RPW = 32 / LPR;  // rows per warp
};

template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int E>
__global__ void __launch_bounds__(256) add_ln_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ zout,
                                                     float4* __restrict__ out, const float4* __restrict__ res, const float4* __restrict__ gamma,
                                                     const float4* __restrict__ beta, long long rows, float eps, const float4* __restrict__ cw,
                                                     const float4* __restrict__ cb, const float* __restrict__ slope, uint2* __restrict__ out_hi,
                                                     uint2* __restrict__ out_lo) {
    using G = LnGeom<E>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % G::LPR, rw = lane / G::LPR;
    const long long stride = (long long)gridDim.x * 8 * G::RPW;
    for (long long r0 = ((long long)blockIdx.x * 8 + warp) * G::RPW; r0 < rows; r0 += stride) {  // warp-uniform trip count
        const long long r = r0 + rw;
        const bool valid = r < rows;
        float4 v[G::CPL];
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < G::CPL; ++c) {
            const long long i = r * (E / 4) + sub + c * G::LPR;
            v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                v[c] = ldg_stream(a + i);
                if (b) {
                    float4 w = ldg_stream(b + i);
                    v[c].x += w.x; v[c].y += w.y; v[c].z += w.z; v[c].w += w.w;
                }
                if (zout) zout[i] = v[c];
            }
            s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
        }
        const float mean = row_sum<G::LPR>(s) * (1.0f / E);
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < G::CPL; ++c) {
            v[c].x -= mean; v[c].y -= mean; v[c].z -= mean; v[c].w -= mean;
            ss = fmaf(v[c].x, v[c].x, fmaf(v[c].y, v[c].y, fmaf(v[c].z, v[c].z, fmaf(v[c].w, v[c].w, ss))));
        }
        const float rstd = rsqrtf(row_sum<G::LPR>(ss) * (1.0f / E) + eps);
#pragma unroll
        for (int c = 0; c < G::CPL; ++c) {
            const int c4 = sub + c * G::LPR;
            const long long i = r * (E / 4) + c4;
            const float4 ga = gamma[c4], be = beta[c4];
            float4 o;
            o.x = fmaf(v[c].x * rstd, ga.x, be.x); o.y = fmaf(v[c].y * rstd, ga.y, be.y);
            o.z = fmaf(v[c].z * rstd, ga.z, be.z); o.w = fmaf(v[c].w * rstd, ga.w, be.w);
            if (!valid) continue;
            if (res) {
                float4 w = ldg_stream(res + i);
                o.x += w.x; o.y += w.y; o.z += w.z; o.w += w.w;
            }
            if (cw) {
                const float sl = slope[0];
                const float4 w = cw[c4], bb = cb[c4];
                o.x = fmaf(w.x, o.x, bb.x); o.y = fmaf(w.y, o.y, bb.y); o.z = fmaf(w.z, o.z, bb.z); o.w = fmaf(w.w, o.w, bb.w);
                o.x = o.x >= 0.f ? o.x : sl * o.x; o.y = o.y >= 0.f ? o.y : sl * o.y;
                o.z = o.z >= 0.f ? o.z : sl * o.z; o.w = o.w >= 0.f ? o.w : sl * o.w;
            }
            if (out != nullptr) out[i] = o;
            if (out_hi != nullptr) {  // operand planes for the GEMM that consumes the normalised rows
                uint2 hh, ll;
                split_pair(o.x, o.y, hh.x, ll.x);
                split_pair(o.z, o.w, hh.y, ll.y);
                out_hi[i] = hh;
                if (out_lo != nullptr) out_lo[i] = ll;
            }
        }
    }
}

// dz = rstd * (gamma*dy - mean(gamma*dy) - xhat * mean(gamma*dy*xhat)); dgamma += sum dy*xhat; dbeta += sum dy.
// dz may alias dy; acc (optional) += dz.
template <int E>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ z, float4* __restrict__ dz,
                                                     float4* __restrict__ acc, const float4* __restrict__ gamma, long long rows, float eps,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
    using G = LnGeom<E>;
    __shared__ float sh[2][8][E];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % G::LPR, rw = lane / G::LPR;
    const long long stride = (long long)gridDim.x * 8 * G::RPW;
    float4 dg[G::CPL], db[G::CPL];
#pragma unroll
    for (int c = 0; c < G::CPL; ++c) dg[c] = db[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long r0 = ((long long)blockIdx.x * 8 + warp) * G::RPW; r0 < rows; r0 += stride) {  // warp-uniform trip count
        const long long r = r0 + rw;
        const bool valid = r < rows;
        float4 v[G::CPL], d[G::CPL];
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < G::CPL; ++c) {
            const long long i = r * (E / 4) + sub + c * G::LPR;
            v[c] = d[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                v[c] = ldg_stream(z + i);
                d[c] = dy[i];
            }
            s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
        }
        const float mean = row_sum<G::LPR>(s) * (1.0f / E);
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < G::CPL; ++c) {
            v[c].x -= mean; v[c].y -= mean; v[c].z -= mean; v[c].w -= mean;
            ss = fmaf(v[c].x, v[c].x, fmaf(v[c].y, v[c].y, fmaf(v[c].z, v[c].z, fmaf(v[c].w, v[c].w, ss))));
        }
        const float rstd = rsqrtf(row_sum<G::LPR>(ss) * (1.0f / E) + eps);
        float s1 = 0.f, s2 = 0.f;
        float4 gd[G::CPL];
#pragma unroll
        for (int c = 0; c < G::CPL; ++c) {
            const float4 ga = gamma[sub + c * G::LPR];
            v[c].x *= rstd; v[c].y *= rstd; v[c].z *= rstd; v[c].w *= rstd;  // xhat
            dg[c].x = fmaf(d[c].x, v[c].x, dg[c].x); dg[c].y = fmaf(d[c].y, v[c].y, dg[c].y);
            dg[c].z = fmaf(d[c].z, v[c].z, dg[c].z); dg[c].w = fmaf(d[c].w, v[c].w, dg[c].w);
            db[c].x += d[c].x; db[c].y += d[c].y; db[c].z += d[c].z; db[c].w += d[c].w;
            gd[c] = make_float4(ga.x * d[c].x, ga.y * d[c].y, ga.z * d[c].z, ga.w * d[c].w);
            s1 += (gd[c].x + gd[c].y) + (gd[c].z + gd[c].w);
            s2 = fmaf(gd[c].x, v[c].x, fmaf(gd[c].y, v[c].y, fmaf(gd[c].z, v[c].z, fmaf(gd[c].w, v[c].w, s2))));
        }
        const float m1 = row_sum<G::LPR>(s1) * (1.0f / E), m2 = row_sum<G::LPR>(s2) * (1.0f / E);
#pragma unroll
        for (int c = 0; c < G::CPL; ++c) {
            const long long i = r * (E / 4) + sub + c * G::LPR;
            float4 o;
            o.x = rstd * (gd[c].x - m1 - v[c].x * m2); o.y = rstd * (gd[c].y - m1 - v[c].y * m2);
            o.z = rstd * (gd[c].z - m1 - v[c].z * m2); o.w = rstd * (gd[c].w - m1 - v[c].w * m2);
            if (!valid) continue;
            dz[i] = o;
            if (acc) {
                float4 w = acc[i];
                w.x += o.x; w.y += o.y; w.z += o.z; w.w += o.w;
                acc[i] = w;
            }
        }
    }
    // reduce the parameter gradients: rows-in-warp (shuffle) -> warps (shared) -> global atomics
#pragma unroll
    for (int c = 0; c < G::CPL; ++c) {
        float* g4 = reinterpret_cast<float*>(&dg[c]);
        float* b4 = reinterpret_cast<float*>(&db[c]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float x = g4[k], y = b4[k];
            for (int o = G::LPR; o < 32; o <<= 1) { x += __shfl_xor_sync(0xffffffffu, x, o); y += __shfl_xor_sync(0xffffffffu, y, o); }
            if (rw == 0) {
                sh[0][warp][(sub + c * G::LPR) * 4 + k] = x;
                sh[1][warp][(sub + c * G::LPR) * 4 + k] = y;
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * E; e += blockDim.x) {
        const int which = e / E, col = e % E;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sh[which][w][col];
        atomicAdd((which ? dbeta : dgamma) + col, s);
    }
}

// out[n] += scale * sum_p A[p*lda + n] for any N % 4 == 0, N <= 1024.
__global__ void __launch_bounds__(256) colsum_any_kernel(const float* __restrict__ A, long long lda, int P, int N, float scale,
                                                         float* __restrict__ out, int rows_per_cta) {
    __shared__ float4 sh[256];
    const int C4 = N >> 2;
    const int pbeg = blockIdx.x * rows_per_cta, pend = min(P, pbeg + rows_per_cta);
    for (int c0 = 0; c0 < C4; c0 += 256) {
        const int cw = min(256, C4 - c0);       // columns (float4) handled in this sweep
        const int lanes = 256 / cw;              // row lanes
        const int c4 = threadIdx.x % cw, rl = threadIdx.x / cw;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rl < lanes) {
            int r = pbeg + rl;
            for (; r + 3 * lanes < pend; r += 4 * lanes) {  // four independent 128-bit loads in flight per thread
                const float4 v0 = ldg_stream(reinterpret_cast<const float4*>(A + (size_t)r * lda) + c0 + c4);
                const float4 v1 = ldg_stream(reinterpret_cast<const float4*>(A + (size_t)(r + lanes) * lda) + c0 + c4);
                const float4 v2 = ldg_stream(reinterpret_cast<const float4*>(A + (size_t)(r + 2 * lanes) * lda) + c0 + c4);
                const float4 v3 = ldg_stream(reinterpret_cast<const float4*>(A + (size_t)(r + 3 * lanes) * lda) + c0 + c4);
                s.x += (v0.x + v1.x) + (v2.x + v3.x); s.y += (v0.y + v1.y) + (v2.y + v3.y);
                s.z += (v0.z + v1.z) + (v2.z + v3.z); s.w += (v0.w + v1.w) + (v2.w + v3.w);
            }
            for (; r < pend; r += lanes) {
                const float4 v = ldg_stream(reinterpret_cast<const float4*>(A + (size_t)r * lda) + c0 + c4);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        }
        sh[threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.x < cw) {
            float4 t = sh[threadIdx.x];
            for (int l = 1; l < lanes; ++l) {
                float4 v = sh[l * cw + threadIdx.x];
                t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
            }
            float* o = out + (size_t)(c0 + threadIdx.x) * 4;
            atomicAdd(o, t.x * scale); atomicAdd(o + 1, t.y * scale); atomicAdd(o + 2, t.z * scale); atomicAdd(o + 3, t.w * scale);
        }
        __syncthreads();
    }
}

inline int ln_grid(long long rows, int rpw) {
    long long blocks = ceil_div_ll(rows, 8LL * rpw);
    long long cap = 148LL * 8;
    return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

cudaError_t launch_attn_fwd(const float* QKV, float* O, float* LSE, int E, int heads, const SeqMap& m, cudaStream_t st,
                            __nv_bfloat16* O_hi, __nv_bfloat16* O_lo, unsigned drop_thr, unsigned drop_key, float drop_scale) {
    if (m.nseq <= 0 || m.len <= 0) return cudaSuccess;
    const int D = E / heads;
    if (E % heads || (D != 16 && D != 32)) return cudaErrorInvalidValue;
    const size_t smem = (size_t)2 * m.len * (D + 4) * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    const int threads = m.len >= 256 ? 256 : ((m.len + 31) / 32) * 32;
    const float scale_log2 = kLog2e / sqrtf((float)D);
    dim3 grid(m.nseq, heads);
    cudaError_t e;
    if (D == 16) {
        e = cudaFuncSetAttribute(attn_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attn_fwd_kernel<16><<<grid, threads, smem, st>>>(QKV, O, LSE, E, heads, m, scale_log2, O_hi, O_lo, drop_thr, drop_key, drop_scale);
    } else {
        e = cudaFuncSetAttribute(attn_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attn_fwd_kernel<32><<<grid, threads, smem, st>>>(QKV, O, LSE, E, heads, m, scale_log2, O_hi, O_lo, drop_thr, drop_key, drop_scale);
    }
    return cudaGetLastError();
}

cudaError_t launch_attn_bwd(const float* QKV, const float* O, const float* LSE, const float* dO, float* dQKV, int E, int heads,
                            const SeqMap& m, cudaStream_t st) {
    if (m.nseq <= 0 || m.len <= 0) return cudaSuccess;
    const int D = E / heads;
    if (E % heads || (D != 16 && D != 32)) return cudaErrorInvalidValue;
    const size_t smem = ((size_t)4 * m.len * (D + 4) + 2 * m.len) * sizeof(float);
    if (smem > 220 * 1024) return cudaErrorInvalidValue;
    const int threads = m.len >= 256 ? 256 : ((m.len + 31) / 32) * 32;
    const float scale = 1.0f / sqrtf((float)D);
    dim3 grid(m.nseq, heads);
    cudaError_t e;
    if (D == 16) {
        e = cudaFuncSetAttribute(attn_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attn_bwd_kernel<16><<<grid, threads, smem, st>>>(QKV, O, LSE, dO, dQKV, E, heads, m, scale);
    } else {
        e = cudaFuncSetAttribute(attn_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attn_bwd_kernel<32><<<grid, threads, smem, st>>>(QKV, O, LSE, dO, dQKV, E, heads, m, scale);
    }
    return cudaGetLastError();
}

cudaError_t launch_add_ln(const float* a, const float* b, float* zout, float* out, const float* res, const float* gamma, const float* beta,
                          long long rows, int E, float eps, const float* cw, const float* cb, const float* slope, cudaStream_t st,
                          __nv_bfloat16* out_hi, __nv_bfloat16* out_lo) {
    if (rows <= 0) return cudaSuccess;
#define DP_LN(EE)                                                                                                                   \
    add_ln_kernel<EE><<<ln_grid(rows, LnGeom<EE>::RPW), 256, 0, st>>>(                                                              \
        reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), reinterpret_cast<float4*>(zout),                    \
        reinterpret_cast<float4*>(out), reinterpret_cast<const float4*>(res), reinterpret_cast<const float4*>(gamma),               \
        reinterpret_cast<const float4*>(beta), rows, eps, reinterpret_cast<const float4*>(cw), reinterpret_cast<const float4*>(cb), slope, \
        reinterpret_cast<uint2*>(out_hi), reinterpret_cast<uint2*>(out_lo))
    if (E == 64) DP_LN(64);
    else if (E == 128) DP_LN(128);
    else if (E == 256) DP_LN(256);
    else return cudaErrorInvalidValue;
#undef DP_LN
    return cudaGetLastError();
}

cudaError_t launch_ln_bwd(const float* dy, const float* z, float* dz, float* acc, const float* gamma, long long rows, int E, float eps,
                          float* dgamma, float* dbeta, cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
#define DP_LNB(EE)                                                                                                                  \
    ln_bwd_kernel<EE><<<ln_grid(rows, LnGeom<EE>::RPW * 4), 256, 0, st>>>(                                                          \
        reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(z), reinterpret_cast<float4*>(dz),                     \
        reinterpret_cast<float4*>(acc), reinterpret_cast<const float4*>(gamma), rows, eps, dgamma, dbeta)
    if (E == 64) DP_LNB(64);
    else if (E == 128) DP_LNB(128);
    else if (E == 256) DP_LNB(256);
    else return cudaErrorInvalidValue;
#undef DP_LNB
    return cudaGetLastError();
}

cudaError_t launch_colsum_any(const float* A, long long lda, int P, int N, float scale, float* out, cudaStream_t st) {
    if (P <= 0) return cudaSuccess;
    if ((N & 3) || (lda & 3)) return cudaErrorInvalidValue;
    // ~4 CTAs per SM: a [32500, 1024] operand was 167 us with 64 CTAs of 512 rows (one load in flight per thread)
    int rows_per_cta = ceil_div(P, 4 * 148);
    rows_per_cta = rows_per_cta < 16 ? 16 : rows_per_cta;
    colsum_any_kernel<<<ceil_div(P, rows_per_cta), 256, 0, st>>>(A, lda, P, N, scale, out, rows_per_cta);
    return cudaGetLastError();
}

}  // namespace dp
