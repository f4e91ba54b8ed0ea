// GroupComm TasNet engine (inference): look2hear/models/gc3_network.py:133-184 with group_size > 1 and module "DPRNN" or "DPTNet" --
// context encoder / decoder (GC_RNN, groupcomm.py:10-45), TAC (gc3_basics.py:28-60), the grouped DPRNN stack (dprnn.py:53-88) or the
// grouped DPTNet stack (dptnet.py:133-162: 4-head attention at width n + BiLSTM feed-forward) and the grouped mask head.  The C-ABI is declared in include/dualpath_b200.h (dp_gctasnet_*).
//
// With G groups every per-group operator is tiny (bn_dim / G = 4 or 8 features, hidden_dim / G = 8 or 16 LSTM units, shared by all
// groups), far below a tensor-core tile: the work is HBM / latency bound fp32 CUDA-core arithmetic, one thread per (position, group)
// -- or per (sequence, LSTM unit) in the recurrence -- with the per-group weights in registers or shared memory.
//
// Data layout in HBM (fp32): every activation is "group-channel-last" [positions, G, n] (C = G*n = bn_dim contiguous floats per
// position), so the reference's view / permute / contiguous pairs around TAC, the RNNs and the norms (gc3_basics.py:45-57,
// groupcomm.py:33-43, dprnn.py:64-82) do not exist:
//   frames   enc [B*F, E]   feat / un [B*F, C]   mask [B*F, G, spk*E/G]
//   context  Xc, A, Y [B*Lc, ctx, C]   (position = (b*Lc + l)*ctx + t ; sequences walk t)
//   DPRNN    A, Y [B, S2, K, C]        (position = (b*S2 + s)*K + k ; row sequences walk k, col sequences walk s)
//   LSTM out Hh [positions, G, 2h]
// GroupNorm statistics are double-precision (sum, sum of squares) pairs accumulated with atomics by the producing kernel and consumed
// by the residual kernel that follows; all slots are cleared by one memset at the start of a forward.
#include <new>
#include <vector>

#include "../../include/dualpath_b200.h"
#include "common.cuh"
#include "engine_common.h"
#include "kernels.h"

using namespace dp;

struct dp_gctasnet {
    dp_gctasnet_config cfg;
    std::vector<int64_t> off;
    int64_t n_params;
    int launches;
};

namespace {

constexpr int HEAD = 12, TAC_N = 11, RNN_N = 12, XF_N = 18, GC_LAYER = TAC_N + RNN_N, GC_BLOCK = 2 * GC_LAYER, DP_LAYER = TAC_N + 2 * RNN_N,
              DPT_LAYER = TAC_N + 2 * XF_N;
enum { P_ENC_W, P_BN_G, P_BN_B, P_BN_W, P_OUT_W, P_OUT_B, P_MASK_W, P_MASK_B, P_DEC_W, P_CAT_W, P_CAT_B, P_CAT_A };

struct TacW { const float *w1, *b1, *a1, *w2, *b2, *a2, *w3, *b3, *a3, *gamma, *beta; };
struct RnnW { const float *wih[2], *whh[2], *bih[2], *bhh[2], *pw, *pb, *gamma, *beta; };
// transformer layer of the grouped DPTNet (dptnet.py:45-56): `rnn` = linear1 (the BiLSTM feed-forward) with pw / pb = linear2 and
// gamma / beta = norm2; then the attention projections and norm1
struct XfW { RnnW rnn; const float *win, *bin, *wo, *bo, *g1, *b1; };

__device__ __forceinline__ float prelu1(float v, float a) { return v >= 0.f ? v : a * v; }

// (sum, sum of squares) of one thread's outputs into the statistics slot of its (sample, group).  Lanes of a warp that hold the same
// group of the same sample are combined first (positions of a warp are consecutive, so this is the common case).
__device__ __forceinline__ void stats_add(double* stats, float s, float ss, bool active, int pos, int pos_lane0, int npos, int G, int g,
                                          int pps) {
    const int lane = threadIdx.x & 31;
    int last = pos_lane0 + 32 / G - 1;
    if (last > npos - 1) last = npos - 1;
    const int smp0 = pos_lane0 / pps;
    const bool uniform = smp0 == (last / pps);   // warp-uniform: lane 0's position is the same for the whole warp
    double ds = active ? (double)s : 0.0, dss = active ? (double)ss : 0.0;
    if (uniform) {
        for (int o = G; o < 32; o <<= 1) {
            ds += __shfl_xor_sync(0xffffffffu, ds, o);
            dss += __shfl_xor_sync(0xffffffffu, dss, o);
        }
        if (lane < G && pos_lane0 < npos) {
            double* d = stats + 2 * ((size_t)smp0 * G + g);
            atomicAdd(d, ds);
            atomicAdd(d + 1, dss);
        }
    } else if (active) {
        double* d = stats + 2 * ((size_t)(pos / pps) * G + g);
        atomicAdd(d, ds);
        atomicAdd(d + 1, dss);
    }
}

// ---- TAC: transform, average over the groups, concatenate (gc3_basics.py:38-55); the GroupNorm + residual follow in gn_res ----------
template <int NG, int HG, int GT>
__global__ void __launch_bounds__(256) gc_tac_kernel(const float* __restrict__ X, float* __restrict__ Y, double* __restrict__ stats, TacW w,
                                                     int npos, int pps) {
    constexpr int G = GT;   // compile-time group count: the lane reductions and the gather below unroll completely
    constexpr int TH = 3 * HG, KMAX = (TH + GT - 1) / GT, LD2 = TH + 1;   // odd row pitch: the lanes' column reads hit distinct banks
    __shared__ __align__(16) float s_w1[TH * NG], s_b1[TH], s_w2[TH * LD2], s_b2[TH], s_w3[NG * 2 * TH], s_b3[NG];
    for (int i = threadIdx.x; i < TH * NG; i += blockDim.x) s_w1[i] = w.w1[i];
    for (int i = threadIdx.x; i < TH * TH; i += blockDim.x) s_w2[(i / TH) * LD2 + i % TH] = w.w2[i];
    for (int i = threadIdx.x; i < NG * 2 * TH; i += blockDim.x) s_w3[i] = w.w3[i];
    for (int i = threadIdx.x; i < TH; i += blockDim.x) { s_b1[i] = w.b1[i]; s_b2[i] = w.b2[i]; }
    if (threadIdx.x < NG) s_b3[threadIdx.x] = w.b3[threadIdx.x];
    __syncthreads();
    const float a1 = __ldg(w.a1), a2 = __ldg(w.a2), a3 = __ldg(w.a3);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // element indices fit 31 bits (checked by the host)
    const int pos = i / G, pos0 = (i - (threadIdx.x & 31)) / G;
    const int g = i % G;
    const bool active = pos < npos;
    float x[NG];
#pragma unroll
    for (int j = 0; j < NG; j += 4) {
        float4 v = active ? *reinterpret_cast<const float4*>(X + i * NG + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
    }
    // TAC_mean: its TH rows are spread over the G lanes of a position (row r on lane r % G) and gathered in the output sum below;
    // each group-mean value is folded into those rows as soon as its reduction over the groups is done
    float y[TH], ms[KMAX];
    constexpr float inv_g = 1.f / (float)G;   // G is a power of two: exact
#pragma unroll
    for (int k = 0; k < KMAX; ++k) ms[k] = (g + k * G < TH) ? s_b2[g + k * G] : 0.f;
#pragma unroll
    for (int r = 0; r < TH; ++r) {
        float acc = s_b1[r];
#pragma unroll
        for (int j = 0; j < NG; ++j) acc = fmaf(s_w1[r * NG + j], x[j], acc);
        y[r] = prelu1(acc, a1);
        float v = y[r];
#pragma unroll
        for (int o = G >> 1; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        v *= inv_g;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (g + k * G < TH) ms[k] = fmaf(s_w2[(g + k * G) * LD2 + r], v, ms[k]);
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) ms[k] = prelu1(ms[k], a2);
    float out[NG];
#pragma unroll
    for (int j = 0; j < NG; ++j) {
        float acc = s_b3[j];
#pragma unroll
        for (int r = 0; r < TH; ++r) acc = fmaf(s_w3[j * 2 * TH + r], y[r], acc);
        out[j] = acc;
    }
#pragma unroll
    for (int r = 0; r < TH; ++r) {
        const float v = __shfl_sync(0xffffffffu, ms[r / G], r % G, G);
#pragma unroll
        for (int j = 0; j < NG; ++j) out[j] = fmaf(s_w3[j * 2 * TH + TH + r], v, out[j]);
    }
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < NG; ++j) {
        out[j] = prelu1(out[j], a3);
        s += out[j];
        ss = fmaf(out[j], out[j], ss);
    }
    if (active) {
#pragma unroll
        for (int j = 0; j < NG; j += 4) *reinterpret_cast<float4*>(Y + i * NG + j) = make_float4(out[j], out[j + 1], out[j + 2], out[j + 3]);
    }
    stats_add(stats, s, ss, active, pos, pos0, npos, G, g, pps);
}

// ---- out = res + GroupNorm(1, n)(y): statistics per (sample, group) over n x (positions of the sample) ------------------------------
template <int NG>
__global__ void __launch_bounds__(256) gc_gn_res_kernel(const float* __restrict__ Y, const float* __restrict__ R, float* __restrict__ Out,
                                                        const double* __restrict__ stats, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int total, int G, int pps, double eps,
                                                        const float* __restrict__ cat_w, const float* __restrict__ cat_b,
                                                        const float* __restrict__ cat_a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // element indices fit 31 bits (checked by the host)
    if (i >= total) return;
    const int gshift = 31 - __clz(G);   // G is a power of two
    const int pos = i >> gshift;
    const int g = i & (G - 1);
    const double* d = stats + 2 * ((size_t)(pos / pps) * G + g);
    const double inv = 1.0 / ((double)pps * NG);
    const double mean = d[0] * inv;
    double var = d[1] * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + eps));
#pragma unroll
    for (int j = 0; j < NG; j += 4) {
        const float4 y = *reinterpret_cast<const float4*>(Y + i * NG + j);
        const float4 r = *reinterpret_cast<const float4*>(R + i * NG + j);
        float4 o;
        o.x = r.x + ((y.x - mu) * rstd * __ldg(gamma + j) + __ldg(beta + j));
        o.y = r.y + ((y.y - mu) * rstd * __ldg(gamma + j + 1) + __ldg(beta + j + 1));
        o.z = r.z + ((y.z - mu) * rstd * __ldg(gamma + j + 2) + __ldg(beta + j + 2));
        o.w = r.w + ((y.w - mu) * rstd * __ldg(gamma + j + 3) + __ldg(beta + j + 3));
        if (cat_w) {   // unfold: depthwise 1x1 conv + PReLU of the shared concat_block (dprnn.py:31-34,82)
            const float a = __ldg(cat_a);
            o.x = prelu1(fmaf(o.x, __ldg(cat_w + j), __ldg(cat_b + j)), a);
            o.y = prelu1(fmaf(o.y, __ldg(cat_w + j + 1), __ldg(cat_b + j + 1)), a);
            o.z = prelu1(fmaf(o.z, __ldg(cat_w + j + 2), __ldg(cat_b + j + 2)), a);
            o.w = prelu1(fmaf(o.w, __ldg(cat_w + j + 3), __ldg(cat_b + j + 3)), a);
        }
        *reinterpret_cast<float4*>(Out + i * NG + j) = o;
    }
}

// ---- BiLSTM recurrence of a ProjRNN at group width (gc3_basics.py:19-24): one thread per (sequence, unit), h exchanged by shuffles ---
// sequence (o, g): positions (o / qdiv) * s_hi + (o % qdiv) * s_lo + t * s_t ; blockIdx.y = direction.  Gate rows i, f, g, o.
template <int NG, int HG>
__global__ void __launch_bounds__(128) gc_lstm_kernel(const float* __restrict__ X, float* __restrict__ Hh, RnnW w, long long nouter, int G,
                                                      int len, int qdiv, long long s_hi, long long s_lo, long long s_t, int staged) {
    // staged: the inputs of the CTA's 128 / HG sequences are copied to shared memory first, so that a step never waits for a global
    // load (with few CTAs per SM -- small batches -- the one-step-ahead register prefetch below is bound by the L2 latency)
    extern __shared__ __align__(16) float xsm[];   // [128 / HG][len][NG]
    if (staged) {
        constexpr int SPB = 128 / HG, N4 = NG / 4;
        const long long nseq = nouter * G, qb = (long long)blockIdx.x * SPB;
        for (int idx = threadIdx.x; idx < SPB * len * N4; idx += blockDim.x) {
            const int k4 = idx % N4, s2 = (idx / N4) % SPB, tt = idx / (N4 * SPB);   // adjacent sequences = adjacent groups: contiguous
            long long qq = qb + s2;
            if (qq >= nseq) qq = nseq - 1;
            const long long oo = qq / G, bpp = (oo / qdiv) * s_hi + (oo % qdiv) * s_lo;
            const float4 v = *reinterpret_cast<const float4*>(X + (bpp + (long long)tt * s_t) * (G * NG) + (qq % G) * NG + k4 * 4);
            *reinterpret_cast<float4*>(xsm + ((size_t)s2 * len + tt) * NG + k4 * 4) = v;
        }
        __syncthreads();
    }
    const float* xloc = xsm + (size_t)(threadIdx.x / HG) * len * NG;
    const int dir = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = (int)(i % HG);
    const long long q = i / HG;
    const bool active = q < nouter * G;
    const long long qc = active ? q : nouter * G - 1;
    const int g = (int)(qc % G);
    const long long o = qc / G;
    const long long bp = (o / qdiv) * s_hi + (o % qdiv) * s_lo;
    float wi[4][NG], wh[4][HG], b[4];
#pragma unroll
    for (int gt = 0; gt < 4; ++gt) {
        const int row = gt * HG + j;
#pragma unroll
        for (int k = 0; k < NG; ++k) wi[gt][k] = __ldg(w.wih[dir] + row * NG + k);
#pragma unroll
        for (int k = 0; k < HG; ++k) wh[gt][k] = __ldg(w.whh[dir] + row * HG + k);
        b[gt] = __ldg(w.bih[dir] + row) + __ldg(w.bhh[dir] + row);
    }
    const int C = G * NG, C2 = G * 2 * HG;
    float h = 0.f, c = 0.f;
    int t = dir ? len - 1 : 0;
    const int dt = dir ? -1 : 1;
    float x[NG], xn[NG];
    if (!staged) {
        const float* xp = X + (bp + (long long)t * s_t) * C + g * NG;
#pragma unroll
        for (int k = 0; k < NG; k += 4) { float4 v = *reinterpret_cast<const float4*>(xp + k); xn[k] = v.x; xn[k + 1] = v.y; xn[k + 2] = v.z; xn[k + 3] = v.w; }
    }
    for (int step = 0; step < len; ++step, t += dt) {
        if (staged) {
#pragma unroll
            for (int k = 0; k < NG; k += 4) { float4 v = *reinterpret_cast<const float4*>(xloc + t * NG + k); x[k] = v.x; x[k + 1] = v.y; x[k + 2] = v.z; x[k + 3] = v.w; }
        } else {
#pragma unroll
            for (int k = 0; k < NG; ++k) x[k] = xn[k];
            if (step + 1 < len) {
                const float* xp = X + (bp + (long long)(t + dt) * s_t) * C + g * NG;
#pragma unroll
                for (int k = 0; k < NG; k += 4) { float4 v = *reinterpret_cast<const float4*>(xp + k); xn[k] = v.x; xn[k + 1] = v.y; xn[k + 2] = v.z; xn[k + 3] = v.w; }
            }
        }
        float a[4] = {b[0], b[1], b[2], b[3]};
#pragma unroll
        for (int k = 0; k < NG; ++k) {
#pragma unroll
            for (int gt = 0; gt < 4; ++gt) a[gt] = fmaf(wi[gt][k], x[k], a[gt]);
        }
#pragma unroll
        for (int k = 0; k < HG; ++k) {
            const float hk = __shfl_sync(0xffffffffu, h, k, HG);
#pragma unroll
            for (int gt = 0; gt < 4; ++gt) a[gt] = fmaf(wh[gt][k], hk, a[gt]);
        }
        // ex2.approx / rcp.approx cell functions of the main recurrence engine (common.cuh): ~1e-7 absolute error per evaluation
        const float ig = sigmoid_cell<true>(a[0]), fg = sigmoid_cell<true>(a[1]), gg = tanh_cell<true>(a[2]), og = sigmoid_cell<true>(a[3]);
        c = fmaf(fg, c, ig * gg);
        h = og * tanh_cell<true>(c);
        if (active) Hh[(bp + (long long)t * s_t) * C2 + g * 2 * HG + dir * HG + j] = h;
    }
}

// ---- context stage of GC_RNN fused (groupcomm.py:38-42): BiLSTM + Linear(2h -> n) + GroupNorm(1, n) + residual of one (block, group)
// unit by the 2h lanes of a warp segment.  The norm of this stage spans only ctx x n values of one unit, so nothing has to leave the
// CTA: x and h of the unit sit in shared memory, the norm statistics are shuffle reductions (two-pass variance), A is updated in place.
template <int NG, int HG>
__global__ void __launch_bounds__(128) gc_ctx_rnn_kernel(float* A, RnnW w, long long nunits, int G, int ctx, float eps) {
    constexpr int UL = 2 * HG, UPB = 128 / UL, LDH = UL + 1, MAXO = 8;
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                       // [UPB][ctx][NG]
    float* hs = sm + UPB * ctx * NG;      // [UPB][ctx][LDH]
    __shared__ float s_pw[NG * LDH], s_pb[NG], s_ga[NG], s_be[NG];   // odd row pitch: lanes with different output features hit distinct banks
    for (int i = threadIdx.x; i < NG * UL; i += blockDim.x) s_pw[(i / UL) * LDH + i % UL] = w.pw[i];
    if (threadIdx.x < NG) { s_pb[threadIdx.x] = w.pb[threadIdx.x]; s_ga[threadIdx.x] = w.gamma[threadIdx.x]; s_be[threadIdx.x] = w.beta[threadIdx.x]; }
    const int C = G * NG;
    // stage the units' inputs: for a fixed frame the units of a CTA are adjacent groups, i.e. contiguous floats
    for (int idx = threadIdx.x; idx < UPB * ctx * (NG / 4); idx += blockDim.x) {
        const int k4 = idx % (NG / 4), uu = (idx / (NG / 4)) % UPB, t = idx / ((NG / 4) * UPB);
        long long qq = (long long)blockIdx.x * UPB + uu;
        if (qq >= nunits) qq = nunits - 1;
        const float4 v = *reinterpret_cast<const float4*>(A + ((qq / G) * ctx + t) * C + (qq % G) * NG + k4 * 4);
        *reinterpret_cast<float4*>(xs + (uu * ctx + t) * NG + k4 * 4) = v;
    }
    const int u = threadIdx.x / UL, l = threadIdx.x % UL, dir = l / HG, j = l % HG;
    const long long q = (long long)blockIdx.x * UPB + u;
    const bool active = q < nunits;
    float wi[4][NG], wh[4][HG], b[4];
#pragma unroll
    for (int gt = 0; gt < 4; ++gt) {
        const int row = gt * HG + j;
#pragma unroll
        for (int k = 0; k < NG; ++k) wi[gt][k] = __ldg(w.wih[dir] + row * NG + k);
#pragma unroll
        for (int k = 0; k < HG; ++k) wh[gt][k] = __ldg(w.whh[dir] + row * HG + k);
        b[gt] = __ldg(w.bih[dir] + row) + __ldg(w.bhh[dir] + row);
    }
    __syncthreads();
    float h = 0.f, c = 0.f;
    int t = dir ? ctx - 1 : 0;
    const int dt = dir ? -1 : 1;
    for (int step = 0; step < ctx; ++step, t += dt) {
        float x[NG];
#pragma unroll
        for (int k = 0; k < NG; k += 4) {
            const float4 v = *reinterpret_cast<const float4*>(xs + (u * ctx + t) * NG + k);
            x[k] = v.x; x[k + 1] = v.y; x[k + 2] = v.z; x[k + 3] = v.w;
        }
        float a[4] = {b[0], b[1], b[2], b[3]};
#pragma unroll
        for (int k = 0; k < NG; ++k) {
#pragma unroll
            for (int gt = 0; gt < 4; ++gt) a[gt] = fmaf(wi[gt][k], x[k], a[gt]);
        }
#pragma unroll
        for (int k = 0; k < HG; ++k) {
            const float hk = __shfl_sync(0xffffffffu, h, k, HG);
#pragma unroll
            for (int gt = 0; gt < 4; ++gt) a[gt] = fmaf(wh[gt][k], hk, a[gt]);
        }
        const float ig = sigmoid_cell<true>(a[0]), fg = sigmoid_cell<true>(a[1]), gg = tanh_cell<true>(a[2]), og = sigmoid_cell<true>(a[3]);
        c = fmaf(fg, c, ig * gg);
        h = og * tanh_cell<true>(c);
        hs[(u * ctx + t) * LDH + l] = h;
    }
    __syncwarp();   // a unit's 2h lanes live in one warp
    const int nout = ctx * NG;
    float y[MAXO], s = 0.f;
#pragma unroll
    for (int m = 0; m < MAXO; ++m) {
        const int o = l + m * UL;
        y[m] = 0.f;
        if (o < nout) {
            const int tt = o / NG, k = o % NG;
            float acc = s_pb[k];
#pragma unroll
            for (int cc = 0; cc < UL; ++cc) acc = fmaf(s_pw[k * LDH + cc], hs[(u * ctx + tt) * LDH + cc], acc);
            y[m] = acc;
            s += acc;
        }
    }
#pragma unroll
    for (int o = UL >> 1; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)nout;
    float ss = 0.f;
#pragma unroll
    for (int m = 0; m < MAXO; ++m)
        if (l + m * UL < nout) ss = fmaf(y[m] - mean, y[m] - mean, ss);
#pragma unroll
    for (int o = UL >> 1; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = 1.f / sqrtf(ss / (float)nout + eps);
    if (active) {
        float* base = A + (q / G) * ctx * C + (q % G) * NG;
#pragma unroll
        for (int m = 0; m < MAXO; ++m) {
            const int o = l + m * UL;
            if (o < nout) {
                const int tt = o / NG, k = o % NG;
                base[(long long)tt * C + k] = xs[(u * ctx + tt) * NG + k] + ((y[m] - mean) * rstd * s_ga[k] + s_be[k]);
            }
        }
    }
}

// ---- Linear(2h -> n) of the ProjRNN + statistics of the GroupNorm that follows ------------------------------------------------------
template <int NG, int HG>
__global__ void __launch_bounds__(256) gc_proj_kernel(const float* __restrict__ Hh, float* __restrict__ Y, double* __restrict__ stats, RnnW w,
                                                      int npos, int G, int pps) {
    __shared__ float s_w[NG * 2 * HG], s_b[NG];
    for (int i = threadIdx.x; i < NG * 2 * HG; i += blockDim.x) s_w[i] = w.pw[i];
    if (threadIdx.x < NG) s_b[threadIdx.x] = w.pb[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // element indices fit 31 bits (checked by the host)
    const int gshift = 31 - __clz(G);                      // G is a power of two
    const int pos = i >> gshift, pos0 = (i - (threadIdx.x & 31)) >> gshift;
    const int g = i & (G - 1);
    const bool active = pos < npos;
    float out[NG];
#pragma unroll
    for (int j = 0; j < NG; ++j) out[j] = s_b[j];
#pragma unroll
    for (int k = 0; k < 2 * HG; k += 4) {
        const float4 v = active ? *reinterpret_cast<const float4*>(Hh + (size_t)i * 2 * HG + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NG; ++j) {
            out[j] = fmaf(s_w[j * 2 * HG + k], v.x, out[j]);
            out[j] = fmaf(s_w[j * 2 * HG + k + 1], v.y, out[j]);
            out[j] = fmaf(s_w[j * 2 * HG + k + 2], v.z, out[j]);
            out[j] = fmaf(s_w[j * 2 * HG + k + 3], v.w, out[j]);
        }
    }
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < NG; ++j) { s += out[j]; ss = fmaf(out[j], out[j], ss); }
    if (active) {
#pragma unroll
        for (int j = 0; j < NG; j += 4) *reinterpret_cast<float4*>(Y + i * NG + j) = make_float4(out[j], out[j + 1], out[j + 2], out[j + 3]);
    }
    stats_add(stats, s, ss, active, pos, pos0, npos, G, g, pps);
}

// ---- grouped DPTNet, first half of a transformer layer (dptnet.py:75-77): 4-head self-attention at width n (head width n/4 = 1 or 2),
// out-projection, residual and LayerNorm(n).  One CTA per `spb` sequences; x, q, k, v, o of a sequence live in shared memory.
template <int NG>
__global__ void __launch_bounds__(128) gc_dpt_attn_kernel(const float* __restrict__ A, float* __restrict__ Z, XfW w, long long nouter, int G,
                                                          int len, int spb, int qdiv, long long s_hi, long long s_lo, long long s_t) {
    constexpr int HD = NG / 4;
    extern __shared__ __align__(16) float sm[];
    const int items = spb * len, C = G * NG;
    float *xs = sm, *qs = xs + items * NG, *ks = qs + items * NG, *vs = ks + items * NG, *os = vs + items * NG;
    const long long nseq = nouter * G, q0 = (long long)blockIdx.x * spb;
    const float scale = HD == 1 ? 1.f : 0.70710678118654752f;   // head_dim ** -0.5
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const long long q = q0 + it / len;
        if (q >= nseq) continue;
        const int i = it % len, g = (int)(q % G);
        const long long o = q / G, pos = (o / qdiv) * s_hi + (o % qdiv) * s_lo + (long long)i * s_t;
        float x[NG];
#pragma unroll
        for (int k = 0; k < NG; k += 4) {
            const float4 v = *reinterpret_cast<const float4*>(A + pos * C + g * NG + k);
            x[k] = v.x; x[k + 1] = v.y; x[k + 2] = v.z; x[k + 3] = v.w;
        }
#pragma unroll
        for (int r = 0; r < 3 * NG; ++r) {
            float acc = __ldg(w.bin + r);
#pragma unroll
            for (int k = 0; k < NG; ++k) acc = fmaf(__ldg(w.win + r * NG + k), x[k], acc);
            if (r < NG) qs[it * NG + r] = acc * scale;
            else if (r < 2 * NG) ks[it * NG + r - NG] = acc;
            else vs[it * NG + r - 2 * NG] = acc;
        }
#pragma unroll
        for (int k = 0; k < NG; ++k) xs[it * NG + k] = x[k];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < items * 4; idx += blockDim.x) {
        const int it = idx >> 2, hh = idx & 3, sl = it / len;
        if (q0 + sl >= nseq) continue;
        const float* kb = ks + sl * len * NG + hh * HD;
        const float* vb = vs + sl * len * NG + hh * HD;
        float qv[HD], acc[HD], m = -INFINITY, den = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) { qv[d] = qs[it * NG + hh * HD + d] * 1.4426950408889634f; acc[d] = 0.f; }
        for (int j = 0; j < len; ++j) {
            float sc = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) sc = fmaf(qv[d], kb[j * NG + d], sc);
            m = fmaxf(m, sc);
        }
        for (int j = 0; j < len; ++j) {
            float sc = -m;
#pragma unroll
            for (int d = 0; d < HD; ++d) sc = fmaf(qv[d], kb[j * NG + d], sc);
            const float e = ex2_approx(sc);
            den += e;
#pragma unroll
            for (int d = 0; d < HD; ++d) acc[d] = fmaf(e, vb[j * NG + d], acc[d]);
        }
        const float inv = 1.f / den;
#pragma unroll
        for (int d = 0; d < HD; ++d) os[it * NG + hh * HD + d] = acc[d] * inv;
    }
    __syncthreads();
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const long long q = q0 + it / len;
        if (q >= nseq) continue;
        const int i = it % len, g = (int)(q % G);
        const long long o = q / G, pos = (o / qdiv) * s_hi + (o % qdiv) * s_lo + (long long)i * s_t;
        float z[NG], mean = 0.f;
#pragma unroll
        for (int r = 0; r < NG; ++r) {
            float acc = __ldg(w.bo + r);
#pragma unroll
            for (int k = 0; k < NG; ++k) acc = fmaf(__ldg(w.wo + r * NG + k), os[it * NG + k], acc);
            z[r] = xs[it * NG + r] + acc;
            mean += z[r];
        }
        mean *= 1.f / NG;
        float var = 0.f;
#pragma unroll
        for (int r = 0; r < NG; ++r) var = fmaf(z[r] - mean, z[r] - mean, var);
        const float rstd = 1.f / sqrtf(var * (1.f / NG) + 1e-5f);
#pragma unroll
        for (int r = 0; r < NG; r += 4)
            *reinterpret_cast<float4*>(Z + pos * C + g * NG + r) =
                make_float4((z[r] - mean) * rstd * __ldg(w.g1 + r) + __ldg(w.b1 + r), (z[r + 1] - mean) * rstd * __ldg(w.g1 + r + 1) + __ldg(w.b1 + r + 1),
                            (z[r + 2] - mean) * rstd * __ldg(w.g1 + r + 2) + __ldg(w.b1 + r + 2), (z[r + 3] - mean) * rstd * __ldg(w.g1 + r + 3) + __ldg(w.b1 + r + 3));
    }
}

// ---- second half (dptnet.py:79-82) after the BiLSTM: ReLU, Linear(4n -> n), residual, LayerNorm(n), then the residual of the dual-path
// block (dptnet.py:150,157) and, with unfold after the column path, the shared concat_block
template <int NG, int HG>
__global__ void __launch_bounds__(256) gc_dpt_out_kernel(const float* __restrict__ Hh, const float* __restrict__ Z, const float* Ain, float* A, XfW w,
                                                         long long total, const float* __restrict__ cat_w, const float* __restrict__ cat_b,
                                                         const float* __restrict__ cat_a) {
    __shared__ float s_w[NG * 2 * HG], s_b[NG];
    for (int i = threadIdx.x; i < NG * 2 * HG; i += blockDim.x) s_w[i] = w.rnn.pw[i];
    if (threadIdx.x < NG) s_b[threadIdx.x] = w.rnn.pb[threadIdx.x];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float y[NG];
#pragma unroll
    for (int j = 0; j < NG; ++j) y[j] = s_b[j];
#pragma unroll
    for (int k = 0; k < 2 * HG; k += 4) {
        float4 v = *reinterpret_cast<const float4*>(Hh + i * 2 * HG + k);
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
#pragma unroll
        for (int j = 0; j < NG; ++j) {
            y[j] = fmaf(s_w[j * 2 * HG + k], v.x, y[j]);
            y[j] = fmaf(s_w[j * 2 * HG + k + 1], v.y, y[j]);
            y[j] = fmaf(s_w[j * 2 * HG + k + 2], v.z, y[j]);
            y[j] = fmaf(s_w[j * 2 * HG + k + 3], v.w, y[j]);
        }
    }
    float mean = 0.f;
#pragma unroll
    for (int j = 0; j < NG; ++j) { y[j] += Z[i * NG + j]; mean += y[j]; }
    mean *= 1.f / NG;
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < NG; ++j) var = fmaf(y[j] - mean, y[j] - mean, var);
    const float rstd = 1.f / sqrtf(var * (1.f / NG) + 1e-5f);
#pragma unroll
    for (int j = 0; j < NG; ++j) {
        float o = Ain[i * NG + j] + ((y[j] - mean) * rstd * __ldg(w.rnn.gamma + j) + __ldg(w.rnn.beta + j));
        if (cat_w) o = prelu1(fmaf(o, __ldg(cat_w + j), __ldg(cat_b + j)), __ldg(cat_a));
        A[i * NG + j] = o;
    }
}

// ---- per-group Linear(n -> nout) shared by the groups: the DPRNN output conv (dprnn.py:85) and the mask conv + ReLU -----------------
template <int NG>
__global__ void __launch_bounds__(256) gc_group_linear_kernel(const float* __restrict__ X, float* __restrict__ Y, const float* __restrict__ W,
                                                              const float* __restrict__ bias, long long total, int nout, int relu) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float x[NG];
#pragma unroll
    for (int j = 0; j < NG; j += 4) { float4 v = *reinterpret_cast<const float4*>(X + i * NG + j); x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w; }
    for (int o = 0; o < nout; ++o) {
        float acc = __ldg(bias + o);
#pragma unroll
        for (int j = 0; j < NG; ++j) acc = fmaf(__ldg(W + o * NG + j), x[j], acc);
        Y[i * nout + o] = relu ? fmaxf(acc, 0.f) : acc;
    }
}

// ---- waveform encoder Conv1d(1, E, win, stride = win/2, no bias) on the zero-padded input + statistics of the bottleneck norm ---------
constexpr int FRAMES_PER_CTA = 32;
__global__ void __launch_bounds__(256) gc_encoder_kernel(const float* __restrict__ x, const float* __restrict__ W, float* __restrict__ enc,
                                                         double* __restrict__ stats, int T, int F, int E, int win) {
    __shared__ double red[16];
    extern __shared__ float wsm[];   // [win][E]: tap-major, so the lanes (consecutive channels) read consecutive words
    for (int i = threadIdx.x; i < E * win; i += blockDim.x) wsm[(i % win) * E + i / win] = W[i];
    __syncthreads();
    const int fpb = blockDim.x / E, fl = threadIdx.x / E, e = threadIdx.x % E, b = blockIdx.y;
    const int stride = win / 2, f0 = blockIdx.x * FRAMES_PER_CTA;
    const float* xb = x + (size_t)b * T;
    float sf = 0.f, ssf = 0.f;   // at most FRAMES_PER_CTA / fpb values per thread: fp32 partials, fp64 from the warp reduction on
    for (int f = f0 + fl; f < min(f0 + FRAMES_PER_CTA, F); f += fpb) {
        float acc = 0.f;
        for (int k = 0; k < win; ++k) {
            const int t = f * stride + k - stride;   // the padded signal starts with `stride` zeros (gc3_network.py:128-129)
            if (t >= 0 && t < T) acc = fmaf(wsm[k * E + e], __ldg(xb + t), acc);
        }
        enc[((size_t)b * F + f) * E + e] = acc;
        sf += acc;
        ssf = fmaf(acc, acc, ssf);
    }
    const double s = warp_sum_d((double)sf), ss = warp_sum_d((double)ssf);
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[2 * warp] = s; red[2 * warp + 1] = ss; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[2 * w + threadIdx.x];
        atomicAdd(stats + 2 * b + threadIdx.x, v);
    }
}

// ---- bottleneck: GroupNorm(1, E, eps = fp32 eps) + Conv1d(E, C, 1, no bias) ------------------------------------------------------------
__global__ void __launch_bounds__(256) gc_bottleneck_kernel(const float* __restrict__ enc, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ W,
                                                            const double* __restrict__ stats, float* __restrict__ feat, int F, int E, int C,
                                                            double eps) {
    extern __shared__ float sm[];
    const int LDW = C + 1;         // padded: the transposing stores below hit distinct banks
    float* wt = sm;                // [E][C + 1]: the weight transposed, so the threads of a frame read consecutive words
    float* row = sm + E * LDW;     // [fpb][E]
    const int fpb = blockDim.x / E, fl = threadIdx.x / E, e = threadIdx.x % E, b = blockIdx.y;
    const int f0 = blockIdx.x * FRAMES_PER_CTA;
    for (int i = threadIdx.x; i < E * C; i += blockDim.x) wt[(i % E) * LDW + i / E] = W[i];
    const double inv = 1.0 / ((double)F * E), mean = stats[2 * b] * inv;
    double var = stats[2 * b + 1] * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + eps));
    const float ga = __ldg(gamma + e) * rstd, be = __ldg(beta + e) - mu * rstd * __ldg(gamma + e);
    for (int fb = f0; fb < min(f0 + FRAMES_PER_CTA, F); fb += fpb) {
        const int f = fb + fl;
        __syncthreads();
        row[fl * E + e] = f < F ? fmaf(enc[((size_t)b * F + f) * E + e], ga, be) : 0.f;
        __syncthreads();
        if (f >= F) continue;
        for (int c = e; c < C; c += E) {
            float acc = 0.f;
#pragma unroll 8
            for (int k = 0; k < E; ++k) acc = fmaf(wt[k * LDW + c], row[fl * E + k], acc);
            feat[((size_t)b * F + f) * C + c] = acc;
        }
    }
}

// mean over the context window (gc3_network.py:150)
__global__ void __launch_bounds__(256) gc_ctx_mean_kernel(const float* __restrict__ A, float* __restrict__ out, long long total, int ctx, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long bl = i / C;
    const int c = (int)(i % C);
    float s = 0.f;
    for (int t = 0; t < ctx; ++t) s += A[(bl * ctx + t) * C + c];
    out[i] = s / (float)ctx;
}

// feature_map.unsqueeze(2) + squeeze_block (gc3_network.py:161)
__global__ void __launch_bounds__(256) gc_bcast_add_kernel(const float* __restrict__ fmap, const float* Xc, float* out,
                                                           long long total, int ctx, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long bl = i / ((long long)ctx * C);
    out[i] = Xc[i] + fmap[bl * C + i % C];
}

// mask * encoder output, ConvTranspose1d(E, 1, win, stride) and the trim of gc3_network.py:174-179, one thread per output sample
__global__ void __launch_bounds__(256) gc_decoder_kernel(const float* __restrict__ Mk, const float* __restrict__ enc, const float* __restrict__ Wd,
                                                         float* __restrict__ est, int B, int spk, int T, int F, int E, int G, int win) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * spk * T) return;
    const int t = (int)(i % T), s = (int)((i / T) % spk), b = (int)(i / ((long long)T * spk));
    const int stride = win / 2, eg = E / G, pos = t + stride;
    float acc = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int f = pos / stride - half, k = pos % stride + half * stride;
        if (f < 0 || f >= F) continue;
        const float* er = enc + ((size_t)b * F + f) * E;
        const float* mr = Mk + ((size_t)b * F + f) * (size_t)(spk * E);
        for (int g = 0, e = 0; g < G; ++g) {
            const float* mg = mr + g * (spk * eg) + s * eg;
            for (int j = 0; j < eg; ++j, ++e) acc = fmaf(__ldg(Wd + e * win + k) * mg[j], er[e], acc);
        }
    }
    est[i] = acc;
}

struct GGeo {
    int B, T, F, Lc, S2, K, ctx, G, n, h, C, E, spk;
    long long PC, PD, PM;   // context positions, DPRNN positions, max
};
bool gc_geometry(const dp_gctasnet* h, int B, int T, GGeo& g) {
    const auto& c = h->cfg;
    if (B <= 0 || T <= 0) return false;
    int rest, frames, crest, drest;
    if (dp_wave_geometry(T, c.win, &rest, &frames)) return false;
    g.B = B; g.T = T; g.F = frames; g.ctx = c.context_size; g.K = c.block_size; g.G = c.group_size;
    g.n = c.bn_dim / c.group_size; g.h = c.hidden_dim / c.group_size; g.C = c.bn_dim; g.E = c.enc_dim; g.spk = c.num_spk;
    if (dp_seg_geometry(g.F, g.ctx, &crest, &g.Lc)) return false;
    if (dp_seg_geometry(g.Lc, g.K, &drest, &g.S2)) return false;
    g.PC = (long long)B * g.Lc * g.ctx;
    g.PD = (long long)B * g.S2 * g.K;
    g.PM = g.PC > g.PD ? g.PC : g.PD;
    if (g.PM * g.G * 2 * g.h >= (1LL << 31)) return false;   // the kernels index elements with 31 bits
    return true;
}
struct GLayout { size_t enc, feat, Xc, A, Y, Hh, sqm, fmap, Mk, stats, stats_bytes, total; };
void gc_layout(const dp_gctasnet* h, const GGeo& g, GLayout& l) {
    Carver c;
    const size_t f = sizeof(float);
    l.enc = c.take((size_t)g.B * g.F * g.E * f);
    l.feat = c.take((size_t)g.B * g.F * g.C * f);
    l.Xc = c.take(g.PC * g.C * f);
    l.A = c.take(g.PM * g.C * f);
    l.Y = c.take(g.PM * g.C * f);
    l.Hh = c.take(g.PM * g.G * 2 * g.h * f);
    l.sqm = c.take((size_t)g.B * g.Lc * g.C * f);
    l.fmap = c.take((size_t)g.B * g.Lc * g.C * f);
    l.Mk = c.take((size_t)g.B * g.F * g.spk * g.E * f);
    l.stats_bytes = ((size_t)g.B + 8 * (size_t)g.B * g.Lc * g.G + 3 * (size_t)h->cfg.layer * g.B * g.G) * 2 * sizeof(double);
    l.stats = c.take(l.stats_bytes);
    l.total = c.off;
}

inline unsigned blocks_for(long long total, int per = 256) { return (unsigned)ceil_div_ll(total, per); }

TacW tac_w(const dp_gctasnet* h, const float* p, int base) {
    const float* q[TAC_N];
    for (int i = 0; i < TAC_N; ++i) q[i] = p + h->off[base + i];
    return TacW{q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9], q[10]};
}
RnnW rnn_w(const dp_gctasnet* h, const float* p, int base) {
    RnnW w;
    for (int d = 0; d < 2; ++d) {
        w.wih[d] = p + h->off[base + 4 * d];
        w.whh[d] = p + h->off[base + 4 * d + 1];
        w.bih[d] = p + h->off[base + 4 * d + 2];
        w.bhh[d] = p + h->off[base + 4 * d + 3];
    }
    w.pw = p + h->off[base + 8]; w.pb = p + h->off[base + 9]; w.gamma = p + h->off[base + 10]; w.beta = p + h->off[base + 11];
    return w;
}

XfW xf_w(const dp_gctasnet* h, const float* p, int base) {
    XfW w;
    w.rnn = rnn_w(h, p, base);   // linear1 (8), linear2.weight / bias, norm2.weight / bias
    w.win = p + h->off[base + 12]; w.bin = p + h->off[base + 13]; w.wo = p + h->off[base + 14]; w.bo = p + h->off[base + 15];
    w.g1 = p + h->off[base + 16]; w.b1 = p + h->off[base + 17];
    return w;
}

struct SeqWalk { long long nouter; int len, qdiv; long long s_hi, s_lo, s_t; };
int g_lstm_staging = 1;   // dp_gctasnet_set_lstm_staging

template <int NG, int HG>
struct Ops {
    static cudaError_t tac(dp_gctasnet* h, const float* X, float* Y, float* Out, double* st, const TacW& w, long long npos, int G, int pps,
                           cudaStream_t s) {
        if (G == 8) gc_tac_kernel<NG, HG, 8><<<blocks_for(npos * G), 256, 0, s>>>(X, Y, st, w, (int)npos, pps);
        else if (G == 16) gc_tac_kernel<NG, HG, 16><<<blocks_for(npos * G), 256, 0, s>>>(X, Y, st, w, (int)npos, pps);
        else gc_tac_kernel<NG, HG, 32><<<blocks_for(npos * G), 256, 0, s>>>(X, Y, st, w, (int)npos, pps);
        gc_gn_res_kernel<NG><<<blocks_for(npos * G), 256, 0, s>>>(Y, X, Out, st, w.gamma, w.beta, (int)(npos * G), G, pps, 1e-5, nullptr, nullptr, nullptr);
        h->launches += 2;
        return cudaGetLastError();
    }
    static cudaError_t lstm(const float* X, float* Hh, const RnnW& w, int G, const SeqWalk& q, cudaStream_t s) {
        dim3 grid(blocks_for(q.nouter * G * HG, 128), 2);
        size_t smem = (size_t)(128 / HG) * q.len * NG * sizeof(float);
        int staged = g_lstm_staging && smem <= 200 * 1024;
        if (!staged) smem = 0;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(gc_lstm_kernel<NG, HG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        gc_lstm_kernel<NG, HG><<<grid, 128, smem, s>>>(X, Hh, w, q.nouter, G, q.len, q.qdiv, q.s_hi, q.s_lo, q.s_t, staged);
        return cudaGetLastError();
    }
    static cudaError_t rnn(dp_gctasnet* h, float* A, float* Y, float* Hh, double* st, const RnnW& w, long long npos, int G, int pps,
                           const SeqWalk& q, double eps, cudaStream_t s, const float* cat_w = nullptr, const float* cat_b = nullptr,
                           const float* cat_a = nullptr) {
        cudaError_t e = lstm(A, Hh, w, G, q, s);
        if (e != cudaSuccess) return e;
        gc_proj_kernel<NG, HG><<<blocks_for(npos * G), 256, 0, s>>>(Hh, Y, st, w, (int)npos, G, pps);
        gc_gn_res_kernel<NG><<<blocks_for(npos * G), 256, 0, s>>>(Y, A, A, st, w.gamma, w.beta, (int)(npos * G), G, pps, eps, cat_w, cat_b, cat_a);
        h->launches += 3;
        return cudaGetLastError();
    }
    // fused context stage (ctx * n / 2h <= 8 outputs per lane, i.e. context_size <= 32 for the built widths)
    static bool ctx_fused(int ctx) { return (ctx * NG + 2 * HG - 1) / (2 * HG) <= 8; }
    static cudaError_t ctx_rnn(dp_gctasnet* h, float* A, const RnnW& w, long long nunits, int G, int ctx, cudaStream_t s) {
        constexpr int UPB = 128 / (2 * HG);
        const size_t smem = (size_t)UPB * ctx * (NG + 2 * HG + 1) * sizeof(float);
        gc_ctx_rnn_kernel<NG, HG><<<blocks_for(nunits, UPB), 128, smem, s>>>(A, w, nunits, G, ctx, 1e-5f);
        h->launches += 1;
        return cudaGetLastError();
    }
    // one transformer layer of the grouped DPTNet along the walk `q`, added onto A (Z = scratch for the first half's output)
    static int xfmr(dp_gctasnet* h, float* A, float* Z, float* Hh, const XfW& w, long long npos, int G, const SeqWalk& q, cudaStream_t s,
                    const float* cat_w, const float* cat_b, const float* cat_a) {
        int spb = 32 / q.len;
        if (spb < 1) spb = 1;
        const size_t smem = (size_t)spb * q.len * NG * 5 * sizeof(float);
        if (smem > 200 * 1024) return fail("dp_gctasnet_forward: sequence of %d frames does not fit the attention kernel's shared memory", q.len);
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(gc_dpt_attn_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gc_dpt_attn_kernel<NG><<<blocks_for(q.nouter * G, spb), 128, smem, s>>>(A, Z, w, q.nouter, G, q.len, spb, q.qdiv, q.s_hi, q.s_lo, q.s_t);
        CK(lstm(Z, Hh, w.rnn, G, q, s));
        gc_dpt_out_kernel<NG, HG><<<blocks_for(npos * G), 256, 0, s>>>(Hh, Z, A, A, w, npos * G, cat_w, cat_b, cat_a);
        h->launches += 3;
        CK(cudaGetLastError());
        return 0;
    }
    static cudaError_t group_linear(dp_gctasnet* h, const float* X, float* Y, const float* W, const float* b, long long total, int nout, int relu,
                                    cudaStream_t s) {
        gc_group_linear_kernel<NG><<<blocks_for(total), 256, 0, s>>>(X, Y, W, b, total, nout, relu);
        h->launches += 1;
        return cudaGetLastError();
    }

    // GC_RNN.forward (groupcomm.py:26-45): `in` is read by the first TAC only (the context blocks stay intact for the decoder)
    static int gc_rnn(dp_gctasnet* h, const float* p, int base, const float* in, float* A, float* Y, float* Hh, double*& st, const GGeo& g,
                      cudaStream_t s) {
        const long long npos = g.PC;
        const SeqWalk q{(long long)g.B * g.Lc, g.ctx, 1, (long long)g.ctx, 0, 1};
        const size_t slot = (size_t)g.B * g.Lc * g.G * 2;
        for (int i = 0; i < 2; ++i) {
            CK(tac(h, i == 0 ? in : A, Y, A, st, tac_w(h, p, base + i * GC_LAYER), npos, g.G, g.ctx, s));
            st += slot;
            if (ctx_fused(g.ctx))
                CK(ctx_rnn(h, A, rnn_w(h, p, base + i * GC_LAYER + TAC_N), (long long)g.B * g.Lc * g.G, g.G, g.ctx, s));
            else
                CK(rnn(h, A, Y, Hh, st, rnn_w(h, p, base + i * GC_LAYER + TAC_N), npos, g.G, g.ctx, q, 1e-5, s));
            st += slot;
        }
        return 0;
    }

    static int forward(dp_gctasnet* h, const float* p, const float* mix, float* est, void* ws, const GGeo& g, cudaStream_t s) {
        const auto& c = h->cfg;
        GLayout l;
        gc_layout(h, g, l);
        float *enc = at<float>(ws, l.enc), *feat = at<float>(ws, l.feat), *Xc = at<float>(ws, l.Xc), *A = at<float>(ws, l.A);
        float *Y = at<float>(ws, l.Y), *Hh = at<float>(ws, l.Hh), *sqm = at<float>(ws, l.sqm), *fmap = at<float>(ws, l.fmap);
        float* Mk = at<float>(ws, l.Mk);
        double* st = at<double>(ws, l.stats);
        h->launches = 0;
        CK(cudaMemsetAsync(st, 0, l.stats_bytes, s));
        const int fpb = 256 / g.E;
        dim3 fgrid(ceil_div(g.F, FRAMES_PER_CTA), g.B);
        gc_encoder_kernel<<<fgrid, 256, (size_t)g.E * c.win * sizeof(float), s>>>(mix, p + h->off[P_ENC_W], enc, st, g.T, g.F, g.E, c.win);
        const size_t smem = ((size_t)g.E * (g.C + 1) + (size_t)fpb * g.E) * sizeof(float);
        gc_bottleneck_kernel<<<fgrid, 256, smem, s>>>(enc, p + h->off[P_BN_G], p + h->off[P_BN_B], p + h->off[P_BN_W], st, feat, g.F, g.E, g.C,
                                                      (double)1.1920928955078125e-07f);
        CK(cudaGetLastError());
        st += 2 * (size_t)g.B;
        h->launches += 2;
        // context encoding (gc3_network.py:145-151)
        CK(launch_segment_cl(feat, Xc, g.B, g.F, g.ctx, g.Lc, g.C, s));
        if (gc_rnn(h, p, HEAD, Xc, A, Y, Hh, st, g, s)) return 1;
        gc_ctx_mean_kernel<<<blocks_for((long long)g.B * g.Lc * g.C), 256, 0, s>>>(A, sqm, (long long)g.B * g.Lc * g.C, g.ctx, g.C);
        CK(cudaGetLastError());
        // DP_Wrapper + grouped DPRNN (groupcomm.py:100-114, dprnn.py:53-88)
        CK(launch_segment_cl(sqm, A, g.B, g.Lc, g.K, g.S2, g.C, s));
        h->launches += 3;
        const int pps = g.S2 * g.K;
        const SeqWalk row{(long long)g.B * g.S2, g.K, 1, (long long)g.K, 0, 1};
        const SeqWalk col{(long long)g.B * g.K, g.S2, g.K, (long long)g.S2 * g.K, 1, (long long)g.K};
        const size_t dslot = (size_t)g.B * g.G * 2;
        const float *cw = c.unfold ? p + h->off[P_CAT_W] : nullptr, *cb = c.unfold ? p + h->off[P_CAT_B] : nullptr;
        const float* ca = c.unfold ? p + h->off[P_CAT_A] : nullptr;
        for (int i = 0; i < c.layer && c.module == DP_MODULE_DPTNET; ++i) {   // dptnet.py:138-157
            const int base = HEAD + 2 * GC_BLOCK + i * DPT_LAYER;
            CK(tac(h, A, Y, A, st, tac_w(h, p, base), g.PD, g.G, pps, s));
            st += dslot;
            if (xfmr(h, A, Y, Hh, xf_w(h, p, base + TAC_N), g.PD, g.G, row, s, nullptr, nullptr, nullptr)) return 1;
            if (xfmr(h, A, Y, Hh, xf_w(h, p, base + TAC_N + XF_N), g.PD, g.G, col, s, cw, cb, ca)) return 1;
        }
        for (int i = 0; i < c.layer && c.module != DP_MODULE_DPTNET; ++i) {
            const int base = HEAD + 2 * GC_BLOCK + i * DP_LAYER;
            CK(tac(h, A, Y, A, st, tac_w(h, p, base), g.PD, g.G, pps, s));
            st += dslot;
            CK(rnn(h, A, Y, Hh, st, rnn_w(h, p, base + TAC_N), g.PD, g.G, pps, row, 1e-8, s));
            st += dslot;
            if (c.unfold)
                CK(rnn(h, A, Y, Hh, st, rnn_w(h, p, base + TAC_N + RNN_N), g.PD, g.G, pps, col, 1e-8, s, p + h->off[P_CAT_W], p + h->off[P_CAT_B],
                       p + h->off[P_CAT_A]));
            else
                CK(rnn(h, A, Y, Hh, st, rnn_w(h, p, base + TAC_N + RNN_N), g.PD, g.G, pps, col, 1e-8, s));
            st += dslot;
        }
        CK(group_linear(h, A, Y, p + h->off[P_OUT_W], p + h->off[P_OUT_B], g.PD * g.G, g.n, 0, s));
        CK(launch_overlap_add_cl(Y, fmap, g.B, g.Lc, g.K, g.S2, g.C, s));
        // context decoding (gc3_network.py:160-166)
        gc_bcast_add_kernel<<<blocks_for(g.PC * g.C), 256, 0, s>>>(fmap, Xc, Xc, g.PC * g.C, g.ctx, g.C);   // in place: the blocks are dead after it
        CK(cudaGetLastError());
        float* Din = Xc;
        if (gc_rnn(h, p, HEAD + GC_BLOCK, Din, A, Y, Hh, st, g, s)) return 1;
        CK(launch_overlap_add_cl(A, feat, g.B, g.F, g.ctx, g.Lc, g.C, s));
        // grouped mask, masking and the decoder (gc3_network.py:169-181)
        CK(group_linear(h, feat, Mk, p + h->off[P_MASK_W], p + h->off[P_MASK_B], (long long)g.B * g.F * g.G, g.spk * g.E / g.G, 1, s));
        gc_decoder_kernel<<<blocks_for((long long)g.B * g.spk * g.T), 256, 0, s>>>(Mk, enc, p + h->off[P_DEC_W], est, g.B, g.spk, g.T, g.F, g.E,
                                                                                   g.G, c.win);
        CK(cudaGetLastError());
        h->launches += 5;
        return 0;
    }
};

}  // namespace

extern "C" {

int dp_gctasnet_n_offsets(int layer, int module) { return HEAD + 2 * GC_BLOCK + layer * (module == DP_MODULE_DPTNET ? DPT_LAYER : DP_LAYER); }

int dp_gctasnet_create(const dp_gctasnet_config* cfg, const int64_t* offsets, int n_offsets, int64_t n_params, dp_gctasnet** out) {
    if (!cfg || !offsets || !out) return fail("dp_gctasnet_create: null argument");
    const int G = cfg->group_size;
    if (G != 8 && G != 16 && G != 32) return fail("dp_gctasnet_create: group_size must be 8, 16 or 32 (got %d)", G);
    if (cfg->bn_dim % G || cfg->hidden_dim % G || cfg->enc_dim % G) return fail("dp_gctasnet_create: group_size must divide enc_dim, bn_dim and hidden_dim");
    const int n = cfg->bn_dim / G, hh = cfg->hidden_dim / G;
    if (!((n == 4 && hh == 8) || (n == 8 && hh == 16)))
        return fail("dp_gctasnet_create: per-group widths (bn_dim / G, hidden_dim / G) must be (4, 8) or (8, 16), got (%d, %d)", n, hh);
    if (cfg->enc_dim % 32 || 256 % cfg->enc_dim) return fail("dp_gctasnet_create: enc_dim must be 32, 64, 128 or 256 (got %d)", cfg->enc_dim);
    if ((size_t)(cfg->enc_dim * (cfg->bn_dim + 1) + 256) * sizeof(float) > 48 * 1024) return fail("dp_gctasnet_create: enc_dim * bn_dim too large");
    if (cfg->win <= 0 || (cfg->win & 1) || cfg->win > 64) return fail("dp_gctasnet_create: win must be even, positive and at most 64");
    if (cfg->context_size <= 0 || (cfg->context_size & 1) || cfg->block_size <= 0 || (cfg->block_size & 1))
        return fail("dp_gctasnet_create: context_size and block_size must be even and positive");
    if (cfg->layer < 1 || cfg->num_spk < 1) return fail("dp_gctasnet_create: layer and num_spk must be >= 1");
    if (cfg->module != DP_MODULE_DPRNN && cfg->module != DP_MODULE_DPTNET) return fail("dp_gctasnet_create: module must be DP_MODULE_DPRNN or DP_MODULE_DPTNET");
    if (n_offsets != dp_gctasnet_n_offsets(cfg->layer, cfg->module))
        return fail("dp_gctasnet_create: expected %d parameter offsets, got %d", dp_gctasnet_n_offsets(cfg->layer, cfg->module), n_offsets);
    for (int i = 0; i < n_offsets; ++i) {
        if (!cfg->unfold && i >= P_CAT_W && i <= P_CAT_A && offsets[i] == -1) continue;   // concat_block exists with unfold only
        if (offsets[i] < 0 || offsets[i] >= n_params || (offsets[i] & 3)) return fail("dp_gctasnet_create: offset %d out of range or not 16-byte aligned", i);
    }
    dp_gctasnet* h = new (std::nothrow) dp_gctasnet;
    if (!h) return fail("dp_gctasnet_create: out of memory");
    h->cfg = *cfg;
    h->off.assign(offsets, offsets + n_offsets);
    h->n_params = n_params;
    h->launches = 0;
    *out = h;
    return 0;
}

void dp_gctasnet_destroy(dp_gctasnet* h) { delete h; }

int64_t dp_gctasnet_workspace_bytes(const dp_gctasnet* h, int B, int T) {
    GGeo g;
    if (!h || !gc_geometry(h, B, T, g)) { fail("dp_gctasnet_workspace_bytes: bad arguments"); return -1; }
    GLayout l;
    gc_layout(h, g, l);
    return (int64_t)l.total;
}

int dp_gctasnet_forward(dp_gctasnet* h, const float* params, const float* mixture, float* est, void* workspace, int B, int T, void* stream) {
    if (!h || !params || !mixture || !est || !workspace) return fail("dp_gctasnet_forward: null argument");
    GGeo g;
    if (!gc_geometry(h, B, T, g)) return fail("dp_gctasnet_forward: bad batch / length (B=%d, T=%d)", B, T);
    if (g.n == 4) return Ops<4, 8>::forward(h, params, mixture, est, workspace, g, S(stream));
    return Ops<8, 16>::forward(h, params, mixture, est, workspace, g, S(stream));
}

int dp_gctasnet_last_launches(const dp_gctasnet* h) { return h ? h->launches : -1; }

int dp_gctasnet_set_lstm_staging(int on) {
    const int prev = g_lstm_staging;
    g_lstm_staging = on ? 1 : 0;
    return prev;
}

}  // extern "C"

// =====================================================================================================================================
// Training: forward that keeps every stage's input, pre-norm tensor, LSTM output and activated gates, and the backward of every stage
// (grouped DPRNN stack; the grouped DPTNet stack has no backward yet).  Reference: autograd through gc3_network.py:133-184,
// groupcomm.py:26-45, gc3_basics.py:7-60, dprnn.py:53-88.  The model has ~31 k parameters shared by all groups, so the kernels below are
// written for clarity (one thread per (position, group) or per (sequence, direction, unit), recomputation instead of extra saves);
// parameter gradients are accumulated by one generic reduction kernel (gct_wgrad_kernel) or in registers (the recurrent weights).
// =====================================================================================================================================
namespace {

struct TacG { float *w1, *b1, *a1, *w2, *b2, *a2, *w3, *b3, *a3, *gamma, *beta; };
struct RnnG { float *wih[2], *whh[2], *bih[2], *bhh[2], *pw, *pb, *gamma, *beta; };
TacG tac_g(const dp_gctasnet* h, float* p, int base) {
    float* q[TAC_N];
    for (int i = 0; i < TAC_N; ++i) q[i] = p + h->off[base + i];
    return TacG{q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9], q[10]};
}
RnnG rnn_g(const dp_gctasnet* h, float* p, int base) {
    RnnG w;
    for (int d = 0; d < 2; ++d) {
        w.wih[d] = p + h->off[base + 4 * d];
        w.whh[d] = p + h->off[base + 4 * d + 1];
        w.bih[d] = p + h->off[base + 4 * d + 2];
        w.bhh[d] = p + h->off[base + 4 * d + 3];
    }
    w.pw = p + h->off[base + 8]; w.pb = p + h->off[base + 9]; w.gamma = p + h->off[base + 10]; w.beta = p + h->off[base + 11];
    return w;
}

__device__ __forceinline__ float block_sum_f(float v, float* sh) {   // sh: 32 floats
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
    return s;
}

// 4-byte asynchronous global -> shared copies: the LSTM kernels below keep several time steps of their operands in flight this way (one
// thread advances one time step per ~0.5 us, while a load takes ~1 us under the write traffic of these kernels)
__device__ __forceinline__ void gct_cp4(void* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void gct_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void gct_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- BiLSTM forward that also saves (i, f, g, o, c) per step: S5 [item = pos * G + g][dir][5][HG] ------------------------------------
template <int NG, int HG>
__global__ void __launch_bounds__(128) gct_lstm_kernel(const float* __restrict__ X, float* __restrict__ Hh, float* __restrict__ S5, RnnW w,
                                                       long long nouter, int G, int len, int qdiv, long long s_hi, long long s_lo, long long s_t) {
    static_assert(128 % HG == 0, "a CTA holds whole unit groups");
    const int dir = blockIdx.y;
    const int j = threadIdx.x % HG;
    float wi[4][NG], wh[4][HG], b[4];
#pragma unroll
    for (int gt = 0; gt < 4; ++gt) {
        const int row = gt * HG + j;
#pragma unroll
        for (int k = 0; k < NG; ++k) wi[gt][k] = __ldg(w.wih[dir] + row * NG + k);
#pragma unroll
        for (int k = 0; k < HG; ++k) wh[gt][k] = __ldg(w.whh[dir] + row * HG + k);
        b[gt] = __ldg(w.bih[dir] + row) + __ldg(w.bhh[dir] + row);
    }
    const int C = G * NG;
    const int dt = dir ? -1 : 1;
    // the input of a (sequence, group) is the same for its HG unit lanes: lane j == 0 requests it FD steps ahead into shared memory
    constexpr int FD = 8;
    __shared__ float xring[FD][128 / HG][NG];
    const int grp = threadIdx.x / HG;
    // a CTA walks several blocks of 128 / HG sequences: the weights stay in registers
    for (long long blk = blockIdx.x; blk * (128 / HG) < nouter * G; blk += gridDim.x) {
    const long long q = blk * (128 / HG) + grp;
    const bool active = q < nouter * G;
    const long long qc = active ? q : nouter * G - 1;
    const int g = (int)(qc % G);
    const long long o = qc / G;
    const long long bp = (o / qdiv) * s_hi + (o % qdiv) * s_lo;
    float h = 0.f, c = 0.f;
    int t = dir ? len - 1 : 0;
    auto request = [&](int step) {
        if (step < len && j == 0) {
            const int tt = dir ? len - 1 - step : step;
            const float* src = X + (bp + (long long)tt * s_t) * C + g * NG;
#pragma unroll
            for (int k = 0; k < NG; ++k) gct_cp4(&xring[step % FD][grp][k], src + k);
        }
        gct_cp_commit();
    };
    for (int step = 0; step < FD - 1; ++step) request(step);
    for (int step = 0; step < len; ++step, t += dt) {
        const long long pos = bp + (long long)t * s_t;
        float a[4] = {b[0], b[1], b[2], b[3]};
        request(step + FD - 1);
        gct_cp_wait<FD - 1>();
        __syncwarp();
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            const float xk = xring[step % FD][grp][k];
#pragma unroll
            for (int gt = 0; gt < 4; ++gt) a[gt] = fmaf(wi[gt][k], xk, a[gt]);
        }
#pragma unroll
        for (int k = 0; k < HG; ++k) {
            const float hk = __shfl_sync(0xffffffffu, h, k, HG);
#pragma unroll
            for (int gt = 0; gt < 4; ++gt) a[gt] = fmaf(wh[gt][k], hk, a[gt]);
        }
        const float ig = sigmoid_cell<true>(a[0]), fg = sigmoid_cell<true>(a[1]), gg = tanh_cell<true>(a[2]), og = sigmoid_cell<true>(a[3]);
        c = fmaf(fg, c, ig * gg);
        h = og * tanh_cell<true>(c);
        if (active) {
            Hh[(pos * G + g) * 2 * HG + dir * HG + j] = h;
            float* s5 = S5 + ((pos * G + g) * 2 + dir) * 5 * HG + j;
            s5[0] = ig; s5[HG] = fg; s5[2 * HG] = gg; s5[3 * HG] = og; s5[4 * HG] = c;
        }
        __syncwarp();   // lane 0 requests into this step's ring slot again in the next iteration
    }
    }
}

// ---- BPTT of that BiLSTM: thread = (sequence, direction, unit).  dXd [dir][item][NG]: gradient with respect to the LSTM input, one
// buffer per direction (summed later); the recurrent / input weights and bias gradients are accumulated in registers and reduced per CTA.
template <int NG, int HG>
__global__ void __launch_bounds__(128) gct_lstm_bwd_kernel(const float* __restrict__ X, const float* __restrict__ Hh, const float* __restrict__ S5,
                                                           const float* __restrict__ dHh, float* __restrict__ dXd, RnnW w, RnnG gw,
                                                           long long nouter, int G, int len, int qdiv, long long s_hi, long long s_lo,
                                                           long long s_t, long long nitems) {
    constexpr int NACC = 4 * HG + 4 * NG + 4;
    __shared__ float acc_sm[HG * NACC];
    for (int i = threadIdx.x; i < HG * NACC; i += blockDim.x) acc_sm[i] = 0.f;
    __syncthreads();
    static_assert(128 % HG == 0, "a CTA holds whole unit groups");
    const int dir = blockIdx.y;
    const int j = threadIdx.x % HG;
    float wi[4][NG], whc[4][HG];   // whc[gt][k] = W_hh[gt * HG + k][j]: what unit k's gate gradients send back to h_j
#pragma unroll
    for (int gt = 0; gt < 4; ++gt) {
#pragma unroll
        for (int k = 0; k < NG; ++k) wi[gt][k] = __ldg(w.wih[dir] + (gt * HG + j) * NG + k);
#pragma unroll
        for (int k = 0; k < HG; ++k) whc[gt][k] = __ldg(w.whh[dir] + (gt * HG + k) * HG + j);
    }
    float aWhh[4][HG], aWih[4][NG], ab[4];
#pragma unroll
    for (int gt = 0; gt < 4; ++gt) {
        ab[gt] = 0.f;
#pragma unroll
        for (int k = 0; k < HG; ++k) aWhh[gt][k] = 0.f;
#pragma unroll
        for (int k = 0; k < NG; ++k) aWih[gt][k] = 0.f;
    }
    const int C = G * NG;
    // What one step reads from global memory is requested BD steps ahead into a shared-memory ring (cp.async): per thread the saved gates
    // and cell state, the previous h and the incoming gradient; per (sequence, group) the LSTM input, requested by lane j == 0.
    constexpr int BD = 4;
    __shared__ float ring[BD][8][128];
    __shared__ float xring[BD][128 / HG][NG];
    const int grp = threadIdx.x / HG;
    // a CTA walks several blocks of 128 / HG sequences: weights and gradient accumulators stay in registers, one reduction at the end
    for (long long blk = blockIdx.x; blk * (128 / HG) < nouter * G; blk += gridDim.x) {
    const long long q = blk * (128 / HG) + grp;
    const bool active = q < nouter * G;
    const long long qc = active ? q : nouter * G - 1;
    const int g = (int)(qc % G);
    const long long o = qc / G;
    const long long bp = (o / qdiv) * s_hi + (o % qdiv) * s_lo;
    float dh_rec = 0.f, dc_carry = 0.f;
    auto item_of = [&](int s) { return (bp + (long long)(dir ? s : len - 1 - s) * s_t) * G + g; };   // reverse of the forward visiting order
    auto request = [&](int s) {
        if (s < len) {
            const int t = dir ? s : len - 1 - s;
            const int tp = dir ? t + 1 : t - 1;              // the step the forward visited just before t
            const long long pos = bp + (long long)t * s_t, item = pos * G + g;
            const float* s5 = S5 + (item * 2 + dir) * 5 * HG + j;
            const int st = s % BD;
#pragma unroll
            for (int q5 = 0; q5 < 5; ++q5) gct_cp4(&ring[st][q5][threadIdx.x], s5 + q5 * HG);
            if (s != len - 1) gct_cp4(&ring[st][5][threadIdx.x], Hh + ((bp + (long long)tp * s_t) * G + g) * 2 * HG + dir * HG + j);
            gct_cp4(&ring[st][6][threadIdx.x], dHh + item * 2 * HG + dir * HG + j);
            if (j == 0) {
#pragma unroll
                for (int k = 0; k < NG; ++k) gct_cp4(&xring[st][grp][k], X + pos * C + g * NG + k);
            }
        }
        gct_cp_commit();
    };
    for (int s = 0; s < BD - 1; ++s) request(s);
    for (int s = 0; s < len; ++s) {
        const bool first = (s == len - 1);
        request(s + BD - 1);
        if (first) gct_cp_wait<0>(); else gct_cp_wait<BD - 2>();   // step s and, for its starting cell state, step s + 1
        __syncwarp();
        const int st = s % BD;
        const long long item = item_of(s);
        const float ig = ring[st][0][threadIdx.x], fg = ring[st][1][threadIdx.x], gg = ring[st][2][threadIdx.x], og = ring[st][3][threadIdx.x];
        const float c = ring[st][4][threadIdx.x];
        const float cp = first ? 0.f : ring[(s + 1) % BD][4][threadIdx.x];   // the cell state the forward step started from
        const float hp = first ? 0.f : ring[st][5][threadIdx.x];
        const float dh = ring[st][6][threadIdx.x] + dh_rec;
        const float tc = tanh_cell<true>(c);
        const float dc = fmaf(dh * og, 1.f - tc * tc, dc_carry);
        dc_carry = dc * fg;
        float dg[4];
        dg[0] = dc * gg * ig * (1.f - ig);
        dg[1] = dc * cp * fg * (1.f - fg);
        dg[2] = dc * ig * (1.f - gg * gg);
        dg[3] = dh * tc * og * (1.f - og);
        if (!active) { dg[0] = dg[1] = dg[2] = dg[3] = 0.f; }
        float xk[NG], px[NG];
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            xk[k] = xring[st][grp][k];
            px[k] = 0.f;
        }
        dh_rec = 0.f;
        float hpk[HG];
#pragma unroll
        for (int k = 0; k < HG; ++k) hpk[k] = __shfl_sync(0xffffffffu, hp, k, HG);
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) {
            ab[gt] += dg[gt];
#pragma unroll
            for (int k = 0; k < NG; ++k) {
                aWih[gt][k] = fmaf(dg[gt], xk[k], aWih[gt][k]);
                px[k] = fmaf(wi[gt][k], dg[gt], px[k]);
            }
#pragma unroll
            for (int k = 0; k < HG; ++k) {
                aWhh[gt][k] = fmaf(dg[gt], hpk[k], aWhh[gt][k]);
                dh_rec = fmaf(whc[gt][k], __shfl_sync(0xffffffffu, dg[gt], k, HG), dh_rec);
            }
        }
#pragma unroll
        for (int k = 0; k < NG; ++k) {
#pragma unroll
            for (int m = HG >> 1; m; m >>= 1) px[k] += __shfl_xor_sync(0xffffffffu, px[k], m);
        }
        if (active && j == 0) {
#pragma unroll
            for (int k = 0; k < NG; ++k) dXd[(long long)dir * nitems * NG + item * NG + k] = px[k];
        }
        __syncwarp();   // the ring slot of this step is requested again by lane 0 in the next iteration
    }
    }
    // reduce the weight-gradient accumulators over the CTA's sequences (shared-memory atomics), then one global atomic per element
    float* mine = acc_sm + j * NACC;
#pragma unroll
    for (int gt = 0; gt < 4; ++gt) {
#pragma unroll
        for (int k = 0; k < HG; ++k) atomicAdd(mine + gt * HG + k, aWhh[gt][k]);
#pragma unroll
        for (int k = 0; k < NG; ++k) atomicAdd(mine + 4 * HG + gt * NG + k, aWih[gt][k]);
        atomicAdd(mine + 4 * HG + 4 * NG + gt, ab[gt]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < HG * NACC; e += blockDim.x) {
        const int jj = e / NACC, r = e % NACC;
        const float v = acc_sm[e];
        if (r < 4 * HG) atomicAdd(gw.whh[dir] + ((r / HG) * HG + jj) * HG + r % HG, v);
        else if (r < 4 * HG + 4 * NG) { const int r2 = r - 4 * HG; atomicAdd(gw.wih[dir] + ((r2 / NG) * HG + jj) * NG + r2 % NG, v); }
        else { const int gt = r - 4 * HG - 4 * NG; atomicAdd(gw.bih[dir] + gt * HG + jj, v); atomicAdd(gw.bhh[dir] + gt * HG + jj, v); }
    }
}

// ---- GroupNorm(1, n) per (sample, group) backward, pass 1: domain sums of gamma dO and gamma dO yhat, and dgamma / dbeta ----------------
template <int NG>
__global__ void __launch_bounds__(256) gct_gn_sums_kernel(const float* __restrict__ dO, const float* __restrict__ Y, const double* __restrict__ stats,
                                                          const float* __restrict__ gamma, double* __restrict__ bst, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, int total, int G, int pps, double eps, int tiles_per_cta) {
    // A CTA walks a contiguous run of 256-item tiles.  The (sample, group) sums are gathered in shared memory and leave as one fp64 atomic
    // per group whenever the run crosses into another sample (items of a tile that belong to the next sample go to global memory directly);
    // dgamma / dbeta stay in registers until the end of the run.
    __shared__ float sh[32];
    __shared__ float acc[64];   // [G][2], G <= 32
    const int gshift = 31 - __clz(G);
    if (threadIdx.x < 64) acc[threadIdx.x] = 0.f;
    __syncthreads();
    float dga[NG], dbe[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) dga[k] = dbe[k] = 0.f;
    const int ntiles = (total + 255) >> 8;
    const int tile0 = blockIdx.x * tiles_per_cta, tile1 = min(tile0 + tiles_per_cta, ntiles);
    int held = -1;   // the sample whose sums are in acc
    auto flush = [&]() {
        __syncthreads();
        if (held >= 0 && threadIdx.x < 2 * G) {
            atomicAdd(bst + 2 * (size_t)held * G + threadIdx.x, (double)acc[threadIdx.x]);
            acc[threadIdx.x] = 0.f;
        }
        __syncthreads();
    };
    for (int tile = tile0; tile < tile1; ++tile) {
        const int first_sample = ((tile << 8) >> gshift) / pps;
        if (first_sample != held) {
            flush();
            held = first_sample;
        }
        const int i = (tile << 8) + threadIdx.x;
        if (i < total) {
            const int pos = i >> gshift, g = i & (G - 1), smp = pos / pps;
            const double* d = stats + 2 * ((size_t)smp * G + g);
            const double inv = 1.0 / ((double)pps * NG), mean = d[0] * inv;
            double var = d[1] * inv - mean * mean;
            if (var < 0.0) var = 0.0;
            const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + eps));
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int k = 0; k < NG; ++k) {
                const float go = dO[(size_t)i * NG + k];
                const float yh = (Y[(size_t)i * NG + k] - mu) * rstd;
                const float ga = __ldg(gamma + k);
                s1 = fmaf(ga, go, s1);
                s2 = fmaf(ga * go, yh, s2);
                dga[k] = fmaf(go, yh, dga[k]);
                dbe[k] += go;
            }
            if (smp == held) {
                atomicAdd(&acc[2 * g], s1);
                atomicAdd(&acc[2 * g + 1], s2);
            } else {
                double* bb = bst + 2 * ((size_t)smp * G + g);
                atomicAdd(bb, (double)s1);
                atomicAdd(bb + 1, (double)s2);
            }
        }
    }
    flush();
#pragma unroll
    for (int k = 0; k < NG; ++k) {
        const float a = block_sum_f(dga[k], sh), bsum = block_sum_f(dbe[k], sh);
        if (threadIdx.x == 0) { atomicAdd(dgamma + k, a); atomicAdd(dbeta + k, bsum); }
    }
}

template <int NG>
__global__ void __launch_bounds__(256) gct_gn_apply_kernel(const float* __restrict__ dO, const float* __restrict__ Y, const double* __restrict__ stats,
                                                           const double* __restrict__ bst, const float* __restrict__ gamma, float* __restrict__ dY,
                                                           int total, int G, int pps, double eps, const float* __restrict__ Wp,
                                                           float* __restrict__ dH, int nh) {
    // Wp != nullptr: also dH[i][c] = sum_k dY[i][k] Wp[k][c] (c < nh): the data gradient of the projection that produced the norm's input
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int gshift = 31 - __clz(G);
    const int pos = i >> gshift, g = i & (G - 1);
    const size_t dom = (size_t)(pos / pps) * G + g;
    const double inv = 1.0 / ((double)pps * NG), mean = stats[2 * dom] * inv;
    double var = stats[2 * dom + 1] * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + eps));
    const float m1 = (float)(bst[2 * dom] * inv), m2 = (float)(bst[2 * dom + 1] * inv);
    float dy[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) {
        const float yh = (Y[(size_t)i * NG + k] - mu) * rstd;
        dy[k] = rstd * (__ldg(gamma + k) * dO[(size_t)i * NG + k] - m1 - yh * m2);
        dY[(size_t)i * NG + k] = dy[k];
    }
    if (Wp) {
        if ((nh & 3) == 0) {
            for (int c = 0; c < nh; c += 4) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < NG; ++k) {
                    const float4 wv = __ldg(reinterpret_cast<const float4*>(Wp + k * nh + c));
                    o.x = fmaf(dy[k], wv.x, o.x); o.y = fmaf(dy[k], wv.y, o.y); o.z = fmaf(dy[k], wv.z, o.z); o.w = fmaf(dy[k], wv.w, o.w);
                }
                *reinterpret_cast<float4*>(dH + (size_t)i * nh + c) = o;
            }
        } else {
            for (int c = 0; c < nh; ++c) {
                float o = 0.f;
#pragma unroll
                for (int k = 0; k < NG; ++k) o = fmaf(dy[k], __ldg(Wp + k * nh + c), o);
                dH[(size_t)i * nh + c] = o;
            }
        }
    }
}

// unfold: Out = prelu(cat_w (R + GN(Y)) + cat_b): turns dOut into the gradient of u = R + GN(Y) in place; concat_block gradients
template <int NG>
__global__ void __launch_bounds__(256) gct_cat_bwd_kernel(float* __restrict__ dO, const float* __restrict__ R, const float* __restrict__ Y,
                                                          const double* __restrict__ stats, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ cw, const float* __restrict__ cb,
                                                          const float* __restrict__ ca, float* __restrict__ gcw, float* __restrict__ gcb,
                                                          float* __restrict__ gca, int total, int G, int pps, double eps) {
    __shared__ float sh[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < total;
    const int gshift = 31 - __clz(G);
    const int ic = active ? i : total - 1;
    const int pos = ic >> gshift, g = ic & (G - 1);
    const double* d = stats + 2 * ((size_t)(pos / pps) * G + g);
    const double inv = 1.0 / ((double)pps * NG), mean = d[0] * inv;
    double var = d[1] * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + eps));
    const float a = __ldg(ca);
    float da = 0.f, dw[NG], db[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) {
        const float u = R[(size_t)ic * NG + k] + ((Y[(size_t)ic * NG + k] - mu) * rstd * __ldg(gamma + k) + __ldg(beta + k));
        const float z = fmaf(u, __ldg(cw + k), __ldg(cb + k));
        const float go = active ? dO[(size_t)i * NG + k] : 0.f;
        const float dz = z >= 0.f ? go : a * go;
        da += z >= 0.f ? 0.f : go * z;
        dw[k] = dz * u;
        db[k] = dz;
        if (active) dO[(size_t)i * NG + k] = dz * __ldg(cw + k);
    }
    const float sa = block_sum_f(da, sh);
    if (threadIdx.x == 0) atomicAdd(gca, sa);
#pragma unroll
    for (int k = 0; k < NG; ++k) {
        const float x = block_sum_f(dw[k], sh), y = block_sum_f(db[k], sh);
        if (threadIdx.x == 0) { atomicAdd(gcw + k, x); atomicAdd(gcb + k, y); }
    }
}

// ---- generic small weight gradient: dW[r][c] += sum_i A[i*lda + r] * Bm[i*ldb + c]  (r < R, c < Cc), db[r] += sum_i A[i*lda + r] --------
__global__ void __launch_bounds__(256) gct_wgrad_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, long long n,
                                                        int R, int Cc, float* __restrict__ dW, float* __restrict__ db, int relu_b, int chunk) {
    const long long i0 = (long long)blockIdx.x * chunk, i1 = i0 + chunk < n ? i0 + chunk : n;
    for (int e = threadIdx.x; e < R * Cc + (db ? R : 0); e += blockDim.x) {
        float acc = 0.f;
        if (e < R * Cc) {
            const int r = e / Cc, c = e % Cc;
            for (long long i = i0; i < i1; ++i) {
                float b = Bm[i * ldb + c];
                if (relu_b) b = fmaxf(b, 0.f);
                acc = fmaf(A[i * lda + r], b, acc);
            }
            atomicAdd(dW + e, acc);
        } else {
            const int r = e - R * Cc;
            for (long long i = i0; i < i1; ++i) acc += A[i * lda + r];
            atomicAdd(db + r, acc);
        }
    }
}

// Tiled form for R * Cc <= 1024 with Cc % 4 == 0 and contiguous, 16-byte aligned operands (every TAC / projection / attention weight here):
// 128 rows of A and Bm are staged in shared memory with coalesced 16-byte loads; a thread owns one row r and four columns of dW and every
// (256 / (R Cc / 4))-th staged row; the row slices are summed in shared memory, then one global atomic per element and CTA.
constexpr int WG_ROWS = 128;
__global__ void __launch_bounds__(256) gct_wgrad_tile_kernel(const float* __restrict__ A, const float* __restrict__ Bm, long long n, int R, int Cc,
                                                             float* __restrict__ dW, float* __restrict__ db, int relu_b, int chunk) {
    extern __shared__ __align__(16) float wg_sm[];
    float* As = wg_sm;
    float* Bs = wg_sm + WG_ROWS * R;
    const int c4n = Cc >> 2, nog = R * c4n, nsl = 256 / nog;
    const int og = threadIdx.x % nog, sl = threadIdx.x / nog;
    const int r = og / c4n, c4 = og % c4n;
    const bool worker = sl < nsl;
    const long long i0 = (long long)blockIdx.x * chunk, i1 = i0 + chunk < n ? i0 + chunk : n;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, accb = 0.f;
    for (long long t0 = i0; t0 < i1; t0 += WG_ROWS) {
        const int rows = (int)(i1 - t0 < WG_ROWS ? i1 - t0 : WG_ROWS);
        __syncthreads();
        const float4* a4 = reinterpret_cast<const float4*>(A + t0 * R);
        for (int idx = threadIdx.x; idx < WG_ROWS * R / 4; idx += 256)
            reinterpret_cast<float4*>(As)[idx] = idx * 4 < rows * R ? a4[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* b4 = reinterpret_cast<const float4*>(Bm + t0 * Cc);
        for (int idx = threadIdx.x; idx < WG_ROWS * Cc / 4; idx += 256) {
            float4 v = idx * 4 < rows * Cc ? b4[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (relu_b) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            reinterpret_cast<float4*>(Bs)[idx] = v;
        }
        __syncthreads();
        if (worker) {
#pragma unroll 4
            for (int i = sl; i < WG_ROWS; i += nsl) {
                const float a = As[i * R + r];
                const float4 b = *reinterpret_cast<const float4*>(Bs + i * Cc + c4 * 4);
                acc[0] = fmaf(a, b.x, acc[0]);
                acc[1] = fmaf(a, b.y, acc[1]);
                acc[2] = fmaf(a, b.z, acc[2]);
                acc[3] = fmaf(a, b.w, acc[3]);
                accb += a;
            }
        }
    }
    __syncthreads();
    float* red = wg_sm;                       // [nsl][R * Cc] | [nsl][R]
    float* redb = wg_sm + nsl * R * Cc;
    if (worker) {
#pragma unroll
        for (int u = 0; u < 4; ++u) red[sl * R * Cc + r * Cc + c4 * 4 + u] = acc[u];
        if (c4 == 0) redb[sl * R + r] = accb;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < R * Cc + (db ? R : 0); e += 256) {
        float v = 0.f;
        if (e < R * Cc) {
            for (int s = 0; s < nsl; ++s) v += red[s * R * Cc + e];
            atomicAdd(dW + e, v);
        } else {
            for (int s = 0; s < nsl; ++s) v += redb[s * R + (e - R * Cc)];
            atomicAdd(db + (e - R * Cc), v);
        }
    }
}

// ---- generic small linear backward (data): dX[i][k] (+)= sum_o dY[i][o] W[o][k]; optional mask (dY counted where Yact > 0: ReLU) -----------
__global__ void __launch_bounds__(256) gct_lin_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ W, const float* __restrict__ Yact,
                                                          float* __restrict__ dX, long long n, int nin, int nout, int accumulate) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * nin) return;
    const long long i = e / nin;
    const int k = (int)(e % nin);
    float acc = 0.f;
    for (int o = 0; o < nout; ++o) {
        float g = dY[i * nout + o];
        if (Yact && Yact[i * nout + o] <= 0.f) g = 0.f;
        acc = fmaf(g, __ldg(W + o * nin + k), acc);
    }
    dX[e] = accumulate ? dX[e] + acc : acc;
}

__global__ void __launch_bounds__(256) gct_add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                                       float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i] + (c ? c[i] : 0.f);
}
__global__ void __launch_bounds__(256) gct_relu_mask_kernel(float* __restrict__ d, const float* __restrict__ y, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && y[i] <= 0.f) d[i] = 0.f;
}

// ---- TAC backward (gc3_basics.py:38-55 + the GroupNorm / residual that follow).  dYraw = gradient of the TAC output before the norm.
// Writes dX = dOut + (TAC path) and the operands of the weight-gradient reductions:
//   S3 [item][NG] = d(pre-activation 3), A3 [item][2 TH] = [y1 ; y2];  S2 [pos][TH], Mv [pos][TH] = group mean of y1;  S1 [item][TH]
template <int NG, int HG, int G>
__global__ void __launch_bounds__(128) gct_tac_bwd_kernel(const float* __restrict__ X, const float* __restrict__ dYraw, const float* __restrict__ dOut,
                                                          float* __restrict__ dX, TacW w, float* __restrict__ S3, float* __restrict__ A3,
                                                          float* __restrict__ S2, float* __restrict__ Mv, float* __restrict__ S1,
                                                          float* __restrict__ ga1, float* __restrict__ ga2, float* __restrict__ ga3, int npos) {
    constexpr int TH = 3 * HG, W2S = TH + 4, MINE = (TH + G - 1) / G;   // W2S: padded row of W2 (lanes of a group read different rows)
    __shared__ float sh[32];
    // the ~1.5 k weights are read by every thread for every product: shared memory (16-byte broadcast reads), W2 also transposed
    __shared__ __align__(16) float sW1[TH * NG], sW2[TH * W2S], sW2t[TH * W2S], sW3[NG * 2 * TH], sB1[TH], sB2[TH], sB3[NG];
    for (int e = threadIdx.x; e < TH * TH; e += blockDim.x) {
        const float v = __ldg(w.w2 + e);
        sW2[(e / TH) * W2S + e % TH] = v;
        sW2t[(e % TH) * W2S + e / TH] = v;
    }
    for (int e = threadIdx.x; e < TH * NG; e += blockDim.x) sW1[e] = __ldg(w.w1 + e);
    for (int e = threadIdx.x; e < NG * 2 * TH; e += blockDim.x) sW3[e] = __ldg(w.w3 + e);
    for (int e = threadIdx.x; e < TH; e += blockDim.x) { sB1[e] = __ldg(w.b1 + e); sB2[e] = __ldg(w.b2 + e); }
    if (threadIdx.x < NG) sB3[threadIdx.x] = __ldg(w.b3 + threadIdx.x);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int pos = i / G, g = i % G;
    const bool active = pos < npos;
    const int ic = active ? i : npos * G - 1;
    const float a1 = __ldg(w.a1), a2 = __ldg(w.a2), a3 = __ldg(w.a3);
    float x[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) x[k] = X[(size_t)ic * NG + k];
    float p1[TH], mv[TH];   // y1 = PReLU(p1) and y2 = PReLU(p2) are re-derived where they are used (registers)
#pragma unroll
    for (int r = 0; r < TH; ++r) {
        float acc = sB1[r];
#pragma unroll
        for (int k = 0; k < NG; ++k) acc = fmaf(sW1[r * NG + k], x[k], acc);
        p1[r] = acc;
        float v = active ? prelu1(acc, a1) : 0.f;   // inactive lanes only exist in the last, partial warp and belong to positions >= npos
        for (int o = G >> 1; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 32);
        mv[r] = v / (float)G;
    }
    // y2 = PReLU(W2 mean + b2) is the same for the G lanes of a position: lane g computes rows g, g + G, ... and the group exchanges them
    float p2[TH], part[MINE];
#pragma unroll
    for (int m = 0; m < MINE; ++m) {
        const int r = m * G + g;
        float acc = 0.f;
        if (r < TH) {
            acc = sB2[r];
#pragma unroll
            for (int c = 0; c < TH; ++c) acc = fmaf(sW2[r * W2S + c], mv[c], acc);
        }
        part[m] = acc;
    }
#pragma unroll
    for (int r = 0; r < TH; ++r) {
        p2[r] = __shfl_sync(0xffffffffu, part[r / G], r % G, G);
    }
    float d3[NG], da3 = 0.f;
#pragma unroll
    for (int k = 0; k < NG; ++k) {
        float acc = sB3[k];
#pragma unroll
        for (int r = 0; r < TH; ++r) acc = fmaf(sW3[k * 2 * TH + r], prelu1(p1[r], a1), acc);
#pragma unroll
        for (int r = 0; r < TH; ++r) acc = fmaf(sW3[k * 2 * TH + TH + r], prelu1(p2[r], a2), acc);
        const float go = active ? dYraw[(size_t)i * NG + k] : 0.f;
        d3[k] = acc >= 0.f ? go : a3 * go;
        da3 += acc >= 0.f ? 0.f : go * acc;
    }
    float dy1[TH], d2[TH], da2 = 0.f;
#pragma unroll
    for (int r = 0; r < TH; ++r) {
        float s = 0.f, t = 0.f;
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            s = fmaf(sW3[k * 2 * TH + r], d3[k], s);
            t = fmaf(sW3[k * 2 * TH + TH + r], d3[k], t);
        }
        dy1[r] = s;
        for (int o = G >> 1; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o, 32);   // y2 is shared by the groups of a position
        d2[r] = p2[r] >= 0.f ? t : a2 * t;
        da2 += p2[r] >= 0.f ? 0.f : t * p2[r];
    }
    float da1 = 0.f, dx[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) dx[k] = 0.f;
#pragma unroll
    for (int m = 0; m < MINE; ++m) {   // W2^T d2, again one share of the rows per lane
        const int c = m * G + g;
        float dm = 0.f;
        if (c < TH) {
#pragma unroll
            for (int r = 0; r < TH; ++r) dm = fmaf(sW2t[c * W2S + r], d2[r], dm);
        }
        part[m] = dm;
    }
    constexpr int CV = TH % 4 == 0 ? 4 : 1;   // S1 leaves in 16-byte pieces when the row length allows
#pragma unroll
    for (int c0 = 0; c0 < TH; c0 += CV) {
        float d1v[CV];
#pragma unroll
        for (int u = 0; u < CV; ++u) {
            const int c = c0 + u;
            const float dm = __shfl_sync(0xffffffffu, part[c / G], c % G, G);
            const float dyc = dy1[c] + dm / (float)G;
            const float d1 = p1[c] >= 0.f ? dyc : a1 * dyc;
            da1 += p1[c] >= 0.f ? 0.f : dyc * p1[c];
            d1v[u] = d1;
#pragma unroll
            for (int k = 0; k < NG; ++k) dx[k] = fmaf(sW1[c * NG + k], d1, dx[k]);
        }
        if (active) {
            if constexpr (CV == 4) *reinterpret_cast<float4*>(S1 + (size_t)i * TH + c0) = make_float4(d1v[0], d1v[1], d1v[2], d1v[3]);
            else S1[(size_t)i * TH + c0] = d1v[0];
        }
    }
    if (active) {   // 16-byte stores: a thread's rows are contiguous
        if constexpr (NG % 4 == 0 && TH % 4 == 0) {
#pragma unroll
            for (int k = 0; k < NG; k += 4) {
                const float4 o = *reinterpret_cast<const float4*>(dOut + (size_t)i * NG + k);
                *reinterpret_cast<float4*>(dX + (size_t)i * NG + k) = make_float4(o.x + dx[k], o.y + dx[k + 1], o.z + dx[k + 2], o.w + dx[k + 3]);
                *reinterpret_cast<float4*>(S3 + (size_t)i * NG + k) = make_float4(d3[k], d3[k + 1], d3[k + 2], d3[k + 3]);
            }
#pragma unroll
            for (int r = 0; r < TH; r += 4) {
                *reinterpret_cast<float4*>(A3 + (size_t)i * 2 * TH + r) = make_float4(prelu1(p1[r], a1), prelu1(p1[r + 1], a1), prelu1(p1[r + 2], a1), prelu1(p1[r + 3], a1));
                *reinterpret_cast<float4*>(A3 + (size_t)i * 2 * TH + TH + r) = make_float4(prelu1(p2[r], a2), prelu1(p2[r + 1], a2), prelu1(p2[r + 2], a2), prelu1(p2[r + 3], a2));
            }
            if (g == 0) {
#pragma unroll
                for (int r = 0; r < TH; r += 4) {
                    *reinterpret_cast<float4*>(S2 + (size_t)pos * TH + r) = make_float4(d2[r], d2[r + 1], d2[r + 2], d2[r + 3]);
                    *reinterpret_cast<float4*>(Mv + (size_t)pos * TH + r) = make_float4(mv[r], mv[r + 1], mv[r + 2], mv[r + 3]);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < NG; ++k) {
                dX[(size_t)i * NG + k] = dOut[(size_t)i * NG + k] + dx[k];
                S3[(size_t)i * NG + k] = d3[k];
            }
#pragma unroll
            for (int r = 0; r < TH; ++r) {
                A3[(size_t)i * 2 * TH + r] = prelu1(p1[r], a1);
                A3[(size_t)i * 2 * TH + TH + r] = prelu1(p2[r], a2);
            }
            if (g == 0) {
#pragma unroll
                for (int r = 0; r < TH; ++r) { S2[(size_t)pos * TH + r] = d2[r]; Mv[(size_t)pos * TH + r] = mv[r]; }
            }
        }
    }
    // PReLU slopes: a2's gradient is counted once per position (lane g == 0)
    const float s1 = block_sum_f(active ? da1 : 0.f, sh), s2 = block_sum_f(active && g == 0 ? da2 : 0.f, sh), s3 = block_sum_f(da3, sh);
    if (threadIdx.x == 0) { atomicAdd(ga1, s1); atomicAdd(ga2, s2); atomicAdd(ga3, s3); }
}

// ---- ends of the network -----------------------------------------------------------------------------------------------------------------
// decoder backward: dMk[b,f,g,s,j] = enc[b,f,e] * sum_k Wd[e][k] d_est[b,s,t(f,k)],  denc[b,f,e] = sum_s Mk * (same sum);  e = g * eg + j
__global__ void __launch_bounds__(256) gct_decoder_bwd_kernel(const float* __restrict__ d_est, const float* __restrict__ Mk, const float* __restrict__ enc,
                                                              const float* __restrict__ Wd, float* __restrict__ dMk, float* __restrict__ denc,
                                                              float* __restrict__ Q, int B, int spk, int T, int F, int E, int G, int win) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * F * E) return;
    const int e = (int)(i % E), f = (int)((i / E) % F), b = (int)(i / ((long long)E * F));
    const int stride = win / 2, eg = E / G, g = e / eg, j = e % eg;
    float de = 0.f;
    for (int s = 0; s < spk; ++s) {
        float q = 0.f;
        for (int k = 0; k < win; ++k) {
            const int t = f * stride + k - stride;
            if (t >= 0 && t < T) q = fmaf(__ldg(Wd + e * win + k), d_est[((size_t)b * spk + s) * T + t], q);
        }
        const size_t mi = ((size_t)b * F + f) * (size_t)(spk * E) + g * (spk * eg) + s * eg + j;
        dMk[mi] = q * enc[i];
        de = fmaf(q, Mk[mi], de);
        Q[((size_t)b * F + f) * (size_t)(spk * E) + (size_t)s * E + e] = Mk[mi] * enc[i];   // masked encoder output, [b,f,s,e]
    }
    denc[i] = de;
}
// dWd[e][k] += sum_{b,s,f} masked[b,f,s,e] * d_est[b,s,t(f,k)]
__global__ void __launch_bounds__(256) gct_decoder_wgrad_kernel(const float* __restrict__ d_est, const float* __restrict__ Q, float* __restrict__ dWd,
                                                                int B, int spk, int T, int F, int E, int win, int fchunk) {
    const int e = threadIdx.x % E, kk = threadIdx.x / E, kpb = blockDim.x / E;
    const int b = blockIdx.y, f0 = blockIdx.x * fchunk, f1 = min(F, f0 + fchunk), stride = win / 2;
    for (int k = kk; k < win; k += kpb) {
        float acc = 0.f;
        for (int s = 0; s < spk; ++s)
            for (int f = f0; f < f1; ++f) {
                const int t = f * stride + k - stride;
                if (t >= 0 && t < T) acc = fmaf(Q[((size_t)b * F + f) * (size_t)(spk * E) + (size_t)s * E + e], d_est[((size_t)b * spk + s) * T + t], acc);
            }
        atomicAdd(dWd + e * win + k, acc);
    }
}
// encoder: dW[e][k] += sum_{b,f} denc[b,f,e] * x[b,t(f,k)]
__global__ void __launch_bounds__(256) gct_encoder_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ denc, float* __restrict__ dW,
                                                                int T, int F, int E, int win, int fchunk) {
    const int e = threadIdx.x % E, kk = threadIdx.x / E, kpb = blockDim.x / E;
    const int b = blockIdx.y, f0 = blockIdx.x * fchunk, f1 = min(F, f0 + fchunk), stride = win / 2;
    for (int k = kk; k < win; k += kpb) {
        float acc = 0.f;
        for (int f = f0; f < f1; ++f) {
            const int t = f * stride + k - stride;
            if (t >= 0 && t < T) acc = fmaf(denc[((size_t)b * F + f) * E + e], x[(size_t)b * T + t], acc);
        }
        atomicAdd(dW + e * win + k, acc);
    }
}
// bottleneck GroupNorm(1, E) over (F, E) per sample with per-channel affine: pass 1 sums, pass 2 applies and adds into denc
__global__ void __launch_bounds__(256) gct_bn_sums_kernel(const float* __restrict__ dU, const float* __restrict__ enc, const double* __restrict__ stats,
                                                          const float* __restrict__ gamma, double* __restrict__ bst, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, int F, int E, double eps) {
    // grid (frame chunks, B), block = E threads x (256 / E) frame lanes; one thread = one channel of every (256 / E)-th frame of the chunk
    __shared__ float sh[32];
    const int e = threadIdx.x % E, fl = threadIdx.x / E, fpb = blockDim.x / E, b = blockIdx.y;
    const double inv = 1.0 / ((double)F * E), mean = stats[2 * b] * inv;
    double var = stats[2 * b + 1] * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + eps)), ga = __ldg(gamma + e);
    float s1 = 0.f, s2 = 0.f, dg = 0.f, dbv = 0.f;
    for (int f = blockIdx.x * FRAMES_PER_CTA + fl; f < min(F, (int)(blockIdx.x + 1) * FRAMES_PER_CTA); f += fpb) {
        const size_t i = ((size_t)b * F + f) * E + e;
        const float go = dU[i], yh = (enc[i] - mu) * rstd;
        s1 = fmaf(ga, go, s1);
        s2 = fmaf(ga * go, yh, s2);
        dg = fmaf(go, yh, dg);
        dbv += go;
    }
    atomicAdd(dgamma + e, dg);
    atomicAdd(dbeta + e, dbv);
    const float t1 = block_sum_f(s1, sh), t2 = block_sum_f(s2, sh);
    if (threadIdx.x == 0) { atomicAdd(bst + 2 * b, (double)t1); atomicAdd(bst + 2 * b + 1, (double)t2); }
}
__global__ void __launch_bounds__(256) gct_bn_apply_kernel(const float* __restrict__ dU, const float* __restrict__ enc, const double* __restrict__ stats,
                                                           const double* __restrict__ bst, const float* __restrict__ gamma, float* __restrict__ denc,
                                                           long long total, int F, int E, double eps) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int e = (int)(i % E), b = (int)(i / ((long long)E * F));
    const double inv = 1.0 / ((double)F * E), mean = stats[2 * b] * inv;
    double var = stats[2 * b + 1] * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + eps));
    const float m1 = (float)(bst[2 * b] * inv), m2 = (float)(bst[2 * b + 1] * inv);
    const float yh = (enc[i] - mu) * rstd;
    denc[i] += rstd * (__ldg(gamma + e) * dU[i] - m1 - yh * m2);
}
// normalised encoder output (the operand of the bottleneck conv's weight gradient)
__global__ void __launch_bounds__(256) gct_bn_norm_kernel(const float* __restrict__ enc, const double* __restrict__ stats, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ out, long long total, int F, int E,
                                                          double eps) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int e = (int)(i % E), b = (int)(i / ((long long)E * F));
    const double inv = 1.0 / ((double)F * E), mean = stats[2 * b] * inv;
    double var = stats[2 * b + 1] * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    out[i] = (enc[i] - (float)mean) * (float)(1.0 / sqrt(var + eps)) * __ldg(gamma + e) + __ldg(beta + e);
}
__global__ void __launch_bounds__(256) gct_ctx_mean_bwd_kernel(const float* __restrict__ dsq, float* __restrict__ dA, long long total, int ctx, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long bl = i / ((long long)ctx * C);
    dA[i] = dsq[bl * C + i % C] / (float)ctx;
}
__global__ void __launch_bounds__(256) gct_ctx_sum_kernel(const float* __restrict__ dX, float* __restrict__ dfmap, long long total, int ctx, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long bl = i / C;
    const int c = (int)(i % C);
    float s = 0.f;
    for (int t = 0; t < ctx; ++t) s += dX[(bl * ctx + t) * C + c];
    dfmap[i] = s;
}

}  // namespace


namespace {

// ---- grouped DPTNet, backward of the second half of a transformer layer (gc_dpt_out_kernel): Out = [concat_block](Ain + LN(Z + Linear(relu(Hh)))).
// Writes dRes (gradient reaching Ain through the residual), dYn (gradient of the LayerNorm input = of Z through its residual, and the operand of
// the Linear's weight gradient) and dHh (through the ReLU); LayerNorm / concat_block parameter gradients by block reductions.
template <int NG, int HG>
__global__ void __launch_bounds__(256) gct_dpt_out_bwd_kernel(const float* __restrict__ Hh, const float* __restrict__ Z, const float* __restrict__ Ain,
                                                              const float* __restrict__ dOut, float* __restrict__ dRes, float* __restrict__ dYn,
                                                              float* __restrict__ dHh, XfW w, float* __restrict__ ggam, float* __restrict__ gbet,
                                                              long long total, const float* __restrict__ cw, const float* __restrict__ cb,
                                                              const float* __restrict__ ca, float* __restrict__ gcw, float* __restrict__ gcb,
                                                              float* __restrict__ gca) {
    __shared__ float s_w[NG * 2 * HG], s_b[NG], sh[32];
    for (int i = threadIdx.x; i < NG * 2 * HG; i += blockDim.x) s_w[i] = w.rnn.pw[i];
    if (threadIdx.x < NG) s_b[threadIdx.x] = w.rnn.pb[threadIdx.x];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < total;
    const long long ic = active ? i : total - 1;
    float y[NG], hr[2 * HG];
#pragma unroll
    for (int j = 0; j < NG; ++j) y[j] = s_b[j];
#pragma unroll
    for (int k = 0; k < 2 * HG; ++k) {
        hr[k] = fmaxf(Hh[ic * 2 * HG + k], 0.f);
#pragma unroll
        for (int j = 0; j < NG; ++j) y[j] = fmaf(s_w[j * 2 * HG + k], hr[k], y[j]);
    }
    float mean = 0.f;
#pragma unroll
    for (int j = 0; j < NG; ++j) { y[j] += Z[ic * NG + j]; mean += y[j]; }
    mean *= 1.f / NG;
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < NG; ++j) var = fmaf(y[j] - mean, y[j] - mean, var);
    const float rstd = 1.f / sqrtf(var * (1.f / NG) + 1e-5f);
    float du[NG], yh[NG], dga[NG], dbe[NG], dcw[NG], dcb[NG], dca = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int j = 0; j < NG; ++j) {
        yh[j] = (y[j] - mean) * rstd;
        float go = active ? dOut[i * NG + j] : 0.f;
        dcw[j] = dcb[j] = 0.f;
        if (cw) {
            const float u = Ain[ic * NG + j] + (yh[j] * __ldg(w.rnn.gamma + j) + __ldg(w.rnn.beta + j));
            const float z = fmaf(u, __ldg(cw + j), __ldg(cb + j)), a = __ldg(ca);
            const float dz = z >= 0.f ? go : a * go;
            dca += z >= 0.f ? 0.f : go * z;
            dcw[j] = dz * u;
            dcb[j] = dz;
            go = dz * __ldg(cw + j);
        }
        du[j] = go;
        dga[j] = go * yh[j];
        dbe[j] = go;
        const float dyh = go * __ldg(w.rnn.gamma + j);
        m1 += dyh;
        m2 = fmaf(dyh, yh[j], m2);
    }
    m1 *= 1.f / NG;
    m2 *= 1.f / NG;
    float dy[NG];
#pragma unroll
    for (int j = 0; j < NG; ++j) dy[j] = rstd * (du[j] * __ldg(w.rnn.gamma + j) - m1 - yh[j] * m2);
    if (active) {
#pragma unroll
        for (int j = 0; j < NG; ++j) { dRes[i * NG + j] = du[j]; dYn[i * NG + j] = dy[j]; }
#pragma unroll
        for (int k = 0; k < 2 * HG; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < NG; ++j) acc = fmaf(s_w[j * 2 * HG + k], dy[j], acc);
            dHh[i * 2 * HG + k] = hr[k] > 0.f ? acc : 0.f;
        }
    }
#pragma unroll
    for (int j = 0; j < NG; ++j) {
        const float a = block_sum_f(dga[j], sh), b = block_sum_f(dbe[j], sh);
        if (threadIdx.x == 0) { atomicAdd(ggam + j, a); atomicAdd(gbet + j, b); }
        if (cw) {
            const float c = block_sum_f(dcw[j], sh), d = block_sum_f(dcb[j], sh);
            if (threadIdx.x == 0) { atomicAdd(gcw + j, c); atomicAdd(gcb + j, d); }
        }
    }
    if (cw) {
        const float e = block_sum_f(dca, sh);
        if (threadIdx.x == 0) atomicAdd(gca, e);
    }
}

// ---- backward of the first half (gc_dpt_attn_kernel): Z = LN(x + Wo attention(x) + bo).  One CTA per `spb` sequences, everything of a sequence in
// shared memory; the softmax is recomputed.  dZ = gradient of Z; dA = dRes + gradient through x (written).  Operands of the weight-gradient
// reductions go to global scratch: DQ [item][3 NG] (gradient of the in-projection output), DZ [item][NG] (of the LayerNorm input), OS [item][NG]
// (attention output); LayerNorm parameter gradients by block reductions.
template <int NG>
__global__ void __launch_bounds__(128) gct_dpt_attn_bwd_kernel(const float* __restrict__ A, const float* __restrict__ dZ, const float* __restrict__ dRes,
                                                               float* __restrict__ dA, float* __restrict__ DQ, float* __restrict__ DZ,
                                                               float* __restrict__ OS, XfW w, float* __restrict__ gg1, float* __restrict__ gb1,
                                                               long long nouter, int G, int len, int spb, int qdiv, long long s_hi, long long s_lo,
                                                               long long s_t) {
    constexpr int HD = NG / 4;
    extern __shared__ __align__(16) float sm[];
    __shared__ float sh[32];
    const int items = spb * len, C = G * NG;
    float *xs = sm, *qs = xs + items * NG, *ks = qs + items * NG, *vs = ks + items * NG, *os = vs + items * NG, *dos = os + items * NG;
    float *dqs = dos + items * NG, *dks = dqs + items * NG, *dvs = dks + items * NG, *ms = dvs + items * NG, *dens = ms + items * 4, *Ds = dens + items * 4;
    const long long nseq = nouter * G, q0 = (long long)blockIdx.x * spb;
    const float scale = HD == 1 ? 1.f : 0.70710678118654752f;
    // forward recomputation: x, q (scaled), k, v
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const long long q = q0 + it / len;
        if (q >= nseq) continue;
        const int i = it % len, g = (int)(q % G);
        const long long o = q / G, pos = (o / qdiv) * s_hi + (o % qdiv) * s_lo + (long long)i * s_t;
        float x[NG];
#pragma unroll
        for (int k = 0; k < NG; ++k) x[k] = A[pos * C + g * NG + k];
#pragma unroll
        for (int r = 0; r < 3 * NG; ++r) {
            float acc = __ldg(w.bin + r);
#pragma unroll
            for (int k = 0; k < NG; ++k) acc = fmaf(__ldg(w.win + r * NG + k), x[k], acc);
            if (r < NG) qs[it * NG + r] = acc * scale;
            else if (r < 2 * NG) ks[it * NG + r - NG] = acc;
            else vs[it * NG + r - 2 * NG] = acc;
        }
#pragma unroll
        for (int k = 0; k < NG; ++k) xs[it * NG + k] = x[k];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < items * 4; idx += blockDim.x) {   // attention output, row maximum and denominator per (query, head)
        const int it = idx >> 2, hh = idx & 3, sl = it / len;
        if (q0 + sl >= nseq) continue;
        const float* kb = ks + sl * len * NG + hh * HD;
        const float* vb = vs + sl * len * NG + hh * HD;
        float qv[HD], acc[HD], m = -INFINITY, den = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) { qv[d] = qs[it * NG + hh * HD + d] * 1.4426950408889634f; acc[d] = 0.f; }
        for (int j = 0; j < len; ++j) {
            float sc = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) sc = fmaf(qv[d], kb[j * NG + d], sc);
            m = fmaxf(m, sc);
        }
        for (int j = 0; j < len; ++j) {
            float sc = -m;
#pragma unroll
            for (int d = 0; d < HD; ++d) sc = fmaf(qv[d], kb[j * NG + d], sc);
            const float e = ex2_approx(sc);
            den += e;
#pragma unroll
            for (int d = 0; d < HD; ++d) acc[d] = fmaf(e, vb[j * NG + d], acc[d]);
        }
        ms[idx] = m;
        dens[idx] = den;
#pragma unroll
        for (int d = 0; d < HD; ++d) os[it * NG + hh * HD + d] = acc[d] / den;
    }
    __syncthreads();
    // LayerNorm backward, out-projection backward: dos = Wo^T dz
    float ag1[NG], ab1[NG];
#pragma unroll
    for (int r = 0; r < NG; ++r) ag1[r] = ab1[r] = 0.f;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const long long q = q0 + it / len;
        if (q >= nseq) continue;
        const int i = it % len, g = (int)(q % G);
        const long long o = q / G, pos = (o / qdiv) * s_hi + (o % qdiv) * s_lo + (long long)i * s_t, item = pos * G + g;
        float z[NG], mean = 0.f;
#pragma unroll
        for (int r = 0; r < NG; ++r) {
            float acc = __ldg(w.bo + r);
#pragma unroll
            for (int k = 0; k < NG; ++k) acc = fmaf(__ldg(w.wo + r * NG + k), os[it * NG + k], acc);
            z[r] = xs[it * NG + r] + acc;
            mean += z[r];
        }
        mean *= 1.f / NG;
        float var = 0.f;
#pragma unroll
        for (int r = 0; r < NG; ++r) var = fmaf(z[r] - mean, z[r] - mean, var);
        const float rstd = 1.f / sqrtf(var * (1.f / NG) + 1e-5f);
        float dzv[NG], m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int r = 0; r < NG; ++r) {
            const float zh = (z[r] - mean) * rstd, go = dZ[item * NG + r];
            ag1[r] = fmaf(go, zh, ag1[r]);
            ab1[r] += go;
            const float dzh = go * __ldg(w.g1 + r);
            m1 += dzh;
            m2 = fmaf(dzh, zh, m2);
            dzv[r] = zh;   // keep zhat; finished below
        }
        m1 *= 1.f / NG;
        m2 *= 1.f / NG;
#pragma unroll
        for (int r = 0; r < NG; ++r) {
            dzv[r] = rstd * (dZ[item * NG + r] * __ldg(w.g1 + r) - m1 - dzv[r] * m2);
            DZ[item * NG + r] = dzv[r];
            OS[item * NG + r] = os[it * NG + r];
        }
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < NG; ++r) acc = fmaf(__ldg(w.wo + r * NG + k), dzv[r], acc);
            dos[it * NG + k] = acc;
        }
#pragma unroll
        for (int r = 0; r < NG; ++r) xs[it * NG + r] = dzv[r];   // x is not needed any more: keep dz (the residual path's gradient) in its place
    }
    __syncthreads();
    // attention backward, queries: D_i = sum_j p_ij dP_ij, dq_i = sum_j p_ij (dP_ij - D_i) k_j
    for (int idx = threadIdx.x; idx < items * 4; idx += blockDim.x) {
        const int it = idx >> 2, hh = idx & 3, sl = it / len;
        if (q0 + sl >= nseq) continue;
        const float* kb = ks + sl * len * NG + hh * HD;
        const float* vb = vs + sl * len * NG + hh * HD;
        float qv[HD], dov[HD], dq[HD], D = 0.f;
        const float m = ms[idx], inv = 1.f / dens[idx];
#pragma unroll
        for (int d = 0; d < HD; ++d) { qv[d] = qs[it * NG + hh * HD + d] * 1.4426950408889634f; dov[d] = dos[it * NG + hh * HD + d]; dq[d] = 0.f; }
        for (int j = 0; j < len; ++j) {
            float sc = -m, dp = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) { sc = fmaf(qv[d], kb[j * NG + d], sc); dp = fmaf(dov[d], vb[j * NG + d], dp); }
            D = fmaf(ex2_approx(sc) * inv, dp, D);
        }
        for (int j = 0; j < len; ++j) {
            float sc = -m, dp = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) { sc = fmaf(qv[d], kb[j * NG + d], sc); dp = fmaf(dov[d], vb[j * NG + d], dp); }
            const float ds = ex2_approx(sc) * inv * (dp - D);
#pragma unroll
            for (int d = 0; d < HD; ++d) dq[d] = fmaf(ds, kb[j * NG + d], dq[d]);
        }
        Ds[idx] = D;
#pragma unroll
        for (int d = 0; d < HD; ++d) dqs[it * NG + hh * HD + d] = dq[d] * scale;   // gradient of the UNSCALED in-projection output
    }
    __syncthreads();
    // keys / values: dk_j = sum_i dS_ij q_i, dv_j = sum_i p_ij dO_i
    for (int idx = threadIdx.x; idx < items * 4; idx += blockDim.x) {
        const int jt = idx >> 2, hh = idx & 3, sl = jt / len;
        if (q0 + sl >= nseq) continue;
        float kv[HD], vv[HD], dk[HD], dv[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) { kv[d] = ks[jt * NG + hh * HD + d]; vv[d] = vs[jt * NG + hh * HD + d]; dk[d] = dv[d] = 0.f; }
        for (int i = 0; i < len; ++i) {
            const int it = sl * len + i, qi = it * 4 + hh;
            float sc = -ms[qi], dp = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) { sc = fmaf(qs[it * NG + hh * HD + d] * 1.4426950408889634f, kv[d], sc); dp = fmaf(dos[it * NG + hh * HD + d], vv[d], dp); }
            const float pij = ex2_approx(sc) / dens[qi], ds = pij * (dp - Ds[qi]);
#pragma unroll
            for (int d = 0; d < HD; ++d) { dk[d] = fmaf(ds, qs[it * NG + hh * HD + d], dk[d]); dv[d] = fmaf(pij, dos[it * NG + hh * HD + d], dv[d]); }
        }
#pragma unroll
        for (int d = 0; d < HD; ++d) { dks[jt * NG + hh * HD + d] = dk[d]; dvs[jt * NG + hh * HD + d] = dv[d]; }
    }
    __syncthreads();
    // in-projection backward: dx = dz + Win^T [dq ; dk ; dv]
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const long long q = q0 + it / len;
        if (q >= nseq) continue;
        const int i = it % len, g = (int)(q % G);
        const long long o = q / G, pos = (o / qdiv) * s_hi + (o % qdiv) * s_lo + (long long)i * s_t, item = pos * G + g;
        float d3[3 * NG];
#pragma unroll
        for (int r = 0; r < NG; ++r) { d3[r] = dqs[it * NG + r]; d3[NG + r] = dks[it * NG + r]; d3[2 * NG + r] = dvs[it * NG + r]; }
#pragma unroll
        for (int r = 0; r < 3 * NG; ++r) DQ[item * 3 * NG + r] = d3[r];
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            float acc = xs[it * NG + k];
#pragma unroll
            for (int r = 0; r < 3 * NG; ++r) acc = fmaf(__ldg(w.win + r * NG + k), d3[r], acc);
            dA[item * NG + k] = dRes[item * NG + k] + acc;
        }
    }
#pragma unroll
    for (int r = 0; r < NG; ++r) {
        const float a = block_sum_f(ag1[r], sh), b = block_sum_f(ab1[r], sh);
        if (threadIdx.x == 0) { atomicAdd(gg1 + r, a); atomicAdd(gb1 + r, b); }
    }
}

}  // namespace

namespace {

struct TLayout {
    size_t enc, feat, feat2, Mk, Q, xn, sqm, fmap, yout;
    size_t ce[5], cd[5], yce[4], ycd[4], hce[2], hcd[2], sce[2], scd[2];
    std::vector<size_t> dp, ydp, hdp, sdp, zdp;   // zdp: first-half outputs Z of the grouped DPTNet's transformer layers
    size_t DQ, OS;
    size_t stats, bst, stats_bytes;
    size_t gA, gB, T1, T2, T3, S3, A3, S2, Mv, S1, dMk, denc, dframe, dsq;
    size_t s_ce[4], s_cd[4];      // statistics slots (in doubles) of the context stages: TAC 0, RNN 0, TAC 1, RNN 1
    std::vector<size_t> s_dp;     // of the DPRNN stack: per layer TAC, row RNN, column RNN
    size_t total;
};
void t_layout(const dp_gctasnet* h, const GGeo& g, TLayout& l) {
    Carver c;
    const size_t f = sizeof(float), L = (size_t)h->cfg.layer;
    const size_t bfe = (size_t)g.B * g.F * g.E * f, bfc = (size_t)g.B * g.F * g.C * f, pc = (size_t)g.PC * g.C * f, pd = (size_t)g.PD * g.C * f;
    const size_t blc = (size_t)g.B * g.Lc * g.C * f, pm = (size_t)g.PM * g.C * f;
    const size_t th = 3 * (size_t)g.h;
    l.enc = c.take(bfe); l.feat = c.take(bfc); l.feat2 = c.take(bfc); l.xn = c.take(bfe);
    l.Mk = c.take((size_t)g.B * g.F * g.spk * g.E * f); l.Q = c.take((size_t)g.B * g.F * g.spk * g.E * f);
    l.sqm = c.take(blc); l.fmap = c.take(blc); l.yout = c.take(pd);
    for (int i = 0; i < 5; ++i) { l.ce[i] = c.take(pc); l.cd[i] = c.take(pc); }
    for (int i = 0; i < 4; ++i) { l.yce[i] = c.take(pc); l.ycd[i] = c.take(pc); }
    for (int i = 0; i < 2; ++i) {
        l.hce[i] = c.take((size_t)g.PC * g.G * 2 * g.h * f); l.hcd[i] = c.take((size_t)g.PC * g.G * 2 * g.h * f);
        l.sce[i] = c.take((size_t)g.PC * g.G * 2 * g.h * 5 * f); l.scd[i] = c.take((size_t)g.PC * g.G * 2 * g.h * 5 * f);
    }
    l.dp.resize(3 * L + 1); l.ydp.resize(3 * L); l.hdp.resize(2 * L); l.sdp.resize(2 * L); l.s_dp.resize(3 * L); l.zdp.resize(2 * L);
    for (size_t i = 0; i <= 3 * L; ++i) l.dp[i] = c.take(pd);
    for (size_t i = 0; i < 3 * L; ++i) l.ydp[i] = c.take(pd);
    for (size_t i = 0; i < 2 * L; ++i) { l.hdp[i] = c.take((size_t)g.PD * g.G * 2 * g.h * f); l.sdp[i] = c.take((size_t)g.PD * g.G * 2 * g.h * 5 * f); }
    const bool dpt = h->cfg.module == DP_MODULE_DPTNET;
    for (size_t i = 0; i < 2 * L; ++i) l.zdp[i] = dpt ? c.take(pd) : 0;
    l.DQ = dpt ? c.take(3 * pd) : 0;
    l.OS = dpt ? c.take(pd) : 0;
    const size_t slot = (size_t)g.B * g.Lc * g.G * 2, dslot = (size_t)g.B * g.G * 2;
    l.stats_bytes = ((size_t)g.B + 8 * (size_t)g.B * g.Lc * g.G + 3 * L * g.B * g.G) * 2 * sizeof(double);
    l.stats = c.take(l.stats_bytes);
    l.bst = c.take(l.stats_bytes);
    size_t o = 2 * (size_t)g.B;
    for (int i = 0; i < 4; ++i) { l.s_ce[i] = o; o += slot; }
    for (size_t i = 0; i < 3 * L; ++i) { l.s_dp[i] = o; o += dslot; }
    for (int i = 0; i < 4; ++i) { l.s_cd[i] = o; o += slot; }
    l.gA = c.take(pm); l.gB = c.take(pm); l.T1 = c.take(pm); l.T2 = c.take((size_t)g.PM * g.G * 2 * g.h * f); l.T3 = c.take(2 * pm);
    l.S3 = c.take(pm); l.A3 = c.take((size_t)g.PM * g.G * 2 * th * f); l.S2 = c.take((size_t)g.PM * th * f); l.Mv = c.take((size_t)g.PM * th * f);
    l.S1 = c.take((size_t)g.PM * g.G * th * f);
    l.dMk = c.take((size_t)g.B * g.F * g.spk * g.E * f); l.denc = c.take(bfe); l.dframe = c.take(bfc); l.dsq = c.take(blc);
    l.total = c.off;
}

inline int wgrad_chunk(long long n, int blocks) {
    long long c = ceil_div_ll(n, blocks);
    return (int)(c < 64 ? 64 : c);
}
inline cudaError_t wgrad(const float* A, int lda, const float* Bm, int ldb, long long n, int R, int Cc, float* dW, float* db, int relu_b, cudaStream_t s) {
    const bool tiled = lda == R && ldb == Cc && (Cc & 3) == 0 && (R & 3) == 0 && R * Cc <= 1024 && R + Cc <= 128 &&
                       ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Bm)) & 15) == 0;
    if (tiled) {
        int chunk = wgrad_chunk(n, 148 * 8);
        chunk = ceil_div(chunk, WG_ROWS) * WG_ROWS;   // whole tiles: every CTA's first row stays 16-byte aligned
        size_t fl = (size_t)WG_ROWS * (R + Cc);
        if (fl < 1280) fl = 1280;
        static const cudaError_t attr = cudaFuncSetAttribute(gct_wgrad_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (attr != cudaSuccess) return attr;
        gct_wgrad_tile_kernel<<<blocks_for(n, chunk), 256, fl * sizeof(float), s>>>(A, Bm, n, R, Cc, dW, db, relu_b, chunk);
        return cudaGetLastError();
    }
    const int chunk = wgrad_chunk(n, 592);
    gct_wgrad_kernel<<<blocks_for(n, chunk), 256, 0, s>>>(A, lda, Bm, ldb, n, R, Cc, dW, db, relu_b, chunk);
    return cudaGetLastError();
}

// 256-item tiles one CTA of gct_gn_sums_kernel walks: about eight CTAs per SM
inline int gn_tiles_per_cta(long long total) { return (int)ceil_div_ll(ceil_div_ll(total, 256), 148 * 8); }

// grid of the persistent LSTM kernels: x = blocks of sequences (balanced over the rounds a resident set of CTAs needs), y = direction
inline dim3 lstm_grid(long long nblk, int resident_per_dir) {
    const long long rounds = ceil_div_ll(nblk, resident_per_dir);
    return dim3((unsigned)ceil_div_ll(nblk, rounds), 2);
}

template <int NG, int HG>
struct TrainOps {
    using F = Ops<NG, HG>;
    static constexpr int TH = 3 * HG;

    static int rnn_fwd(dp_gctasnet* h, const float* Ain, float* Y, float* Hh, float* S5, float* Out, double* st, const RnnW& w, long long npos, int G,
                       int pps, const SeqWalk& q, double eps, cudaStream_t s, const float* cw, const float* cb, const float* ca) {
        const dim3 grid = lstm_grid(blocks_for(q.nouter * G * HG, 128), 1 << 30);   // one block of sequences per CTA measured faster here
        gct_lstm_kernel<NG, HG><<<grid, 128, 0, s>>>(Ain, Hh, S5, w, q.nouter, G, q.len, q.qdiv, q.s_hi, q.s_lo, q.s_t);
        gc_proj_kernel<NG, HG><<<blocks_for(npos * G), 256, 0, s>>>(Hh, Y, st, w, (int)npos, G, pps);
        gc_gn_res_kernel<NG><<<blocks_for(npos * G), 256, 0, s>>>(Y, Ain, Out, st, w.gamma, w.beta, (int)(npos * G), G, pps, eps, cw, cb, ca);
        h->launches += 3;
        CK(cudaGetLastError());
        return 0;
    }
    // one transformer layer of the grouped DPTNet with everything kept: Z = first half's output, Hh / S5 of its BiLSTM feed-forward
    static int xf_fwd(dp_gctasnet* h, const float* Ain, float* Z, float* Hh, float* S5, float* Out, const XfW& w, long long npos, int G,
                      const SeqWalk& q, cudaStream_t s, const float* cw, const float* cb, const float* ca) {
        int spb = 32 / q.len;
        if (spb < 1) spb = 1;
        const size_t smem = (size_t)spb * q.len * NG * 5 * sizeof(float);
        if (smem > 200 * 1024) return fail("dp_gctasnet_forward_train: sequence of %d frames does not fit the attention kernel's shared memory", q.len);
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(gc_dpt_attn_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gc_dpt_attn_kernel<NG><<<blocks_for(q.nouter * G, spb), 128, smem, s>>>(Ain, Z, w, q.nouter, G, q.len, spb, q.qdiv, q.s_hi, q.s_lo, q.s_t);
        const dim3 grid = lstm_grid(blocks_for(q.nouter * G * HG, 128), 1 << 30);   // one block of sequences per CTA measured faster here
        gct_lstm_kernel<NG, HG><<<grid, 128, 0, s>>>(Z, Hh, S5, w.rnn, q.nouter, G, q.len, q.qdiv, q.s_hi, q.s_lo, q.s_t);
        gc_dpt_out_kernel<NG, HG><<<blocks_for(npos * G), 256, 0, s>>>(Hh, Z, Ain, Out, w, npos * G, cw, cb, ca);
        h->launches += 3;
        CK(cudaGetLastError());
        return 0;
    }
    // GC_RNN with everything kept: X[0] -> X[1] (TAC 0) -> X[2] (RNN 0) -> X[3] (TAC 1) -> X[4] (RNN 1)
    static int gc_rnn_fwd(dp_gctasnet* h, const float* p, int base, void* ws, const size_t* X, const size_t* Y, const size_t* Hh, const size_t* S5,
                          const size_t* slots, double* st0, const GGeo& g, cudaStream_t s) {
        const SeqWalk q{(long long)g.B * g.Lc, g.ctx, 1, (long long)g.ctx, 0, 1};
        for (int i = 0; i < 2; ++i) {
            CK(F::tac(h, at<float>(ws, X[2 * i]), at<float>(ws, Y[2 * i]), at<float>(ws, X[2 * i + 1]), st0 + slots[2 * i],
                      tac_w(h, p, base + i * GC_LAYER), g.PC, g.G, g.ctx, s));
            if (rnn_fwd(h, at<float>(ws, X[2 * i + 1]), at<float>(ws, Y[2 * i + 1]), at<float>(ws, Hh[i]), at<float>(ws, S5[i]),
                        at<float>(ws, X[2 * i + 2]), st0 + slots[2 * i + 1], rnn_w(h, p, base + i * GC_LAYER + TAC_N), g.PC, g.G, g.ctx, q, 1e-5, s,
                        nullptr, nullptr, nullptr))
                return 1;
        }
        return 0;
    }

    static int forward(dp_gctasnet* h, const float* p, const float* mix, float* est, void* ws, const GGeo& g, cudaStream_t s) {
        const auto& c = h->cfg;
        TLayout l;
        t_layout(h, g, l);
        float *enc = at<float>(ws, l.enc), *feat = at<float>(ws, l.feat);
        double* st = at<double>(ws, l.stats);
        h->launches = 0;
        CK(cudaMemsetAsync(st, 0, l.stats_bytes, s));
        const int fpb = 256 / g.E;
        dim3 fgrid(ceil_div(g.F, FRAMES_PER_CTA), g.B);
        gc_encoder_kernel<<<fgrid, 256, (size_t)g.E * c.win * sizeof(float), s>>>(mix, p + h->off[P_ENC_W], enc, st, g.T, g.F, g.E, c.win);
        const size_t smem = ((size_t)g.E * (g.C + 1) + (size_t)fpb * g.E) * sizeof(float);
        gc_bottleneck_kernel<<<fgrid, 256, smem, s>>>(enc, p + h->off[P_BN_G], p + h->off[P_BN_B], p + h->off[P_BN_W], st, feat, g.F, g.E, g.C,
                                                      (double)1.1920928955078125e-07f);
        CK(cudaGetLastError());
        CK(launch_segment_cl(feat, at<float>(ws, l.ce[0]), g.B, g.F, g.ctx, g.Lc, g.C, s));
        if (gc_rnn_fwd(h, p, HEAD, ws, l.ce, l.yce, l.hce, l.sce, l.s_ce, st, g, s)) return 1;
        gc_ctx_mean_kernel<<<blocks_for((long long)g.B * g.Lc * g.C), 256, 0, s>>>(at<float>(ws, l.ce[4]), at<float>(ws, l.sqm),
                                                                                  (long long)g.B * g.Lc * g.C, g.ctx, g.C);
        CK(cudaGetLastError());
        CK(launch_segment_cl(at<float>(ws, l.sqm), at<float>(ws, l.dp[0]), g.B, g.Lc, g.K, g.S2, g.C, s));
        const int pps = g.S2 * g.K;
        const SeqWalk row{(long long)g.B * g.S2, g.K, 1, (long long)g.K, 0, 1};
        const SeqWalk col{(long long)g.B * g.K, g.S2, g.K, (long long)g.S2 * g.K, 1, (long long)g.K};
        const float *cw = c.unfold ? p + h->off[P_CAT_W] : nullptr, *cb = c.unfold ? p + h->off[P_CAT_B] : nullptr;
        const float* ca = c.unfold ? p + h->off[P_CAT_A] : nullptr;
        for (int i = 0; i < c.layer && c.module != DP_MODULE_DPTNET; ++i) {
            const int base = HEAD + 2 * GC_BLOCK + i * DP_LAYER;
            CK(F::tac(h, at<float>(ws, l.dp[3 * i]), at<float>(ws, l.ydp[3 * i]), at<float>(ws, l.dp[3 * i + 1]), st + l.s_dp[3 * i], tac_w(h, p, base),
                      g.PD, g.G, pps, s));
            if (rnn_fwd(h, at<float>(ws, l.dp[3 * i + 1]), at<float>(ws, l.ydp[3 * i + 1]), at<float>(ws, l.hdp[2 * i]), at<float>(ws, l.sdp[2 * i]),
                        at<float>(ws, l.dp[3 * i + 2]), st + l.s_dp[3 * i + 1], rnn_w(h, p, base + TAC_N), g.PD, g.G, pps, row, 1e-8, s, nullptr, nullptr,
                        nullptr))
                return 1;
            if (rnn_fwd(h, at<float>(ws, l.dp[3 * i + 2]), at<float>(ws, l.ydp[3 * i + 2]), at<float>(ws, l.hdp[2 * i + 1]),
                        at<float>(ws, l.sdp[2 * i + 1]), at<float>(ws, l.dp[3 * i + 3]), st + l.s_dp[3 * i + 2], rnn_w(h, p, base + TAC_N + RNN_N), g.PD,
                        g.G, pps, col, 1e-8, s, cw, cb, ca))
                return 1;
        }
        for (int i = 0; i < c.layer && c.module == DP_MODULE_DPTNET; ++i) {   // dptnet.py:138-157
            const int base = HEAD + 2 * GC_BLOCK + i * DPT_LAYER;
            CK(F::tac(h, at<float>(ws, l.dp[3 * i]), at<float>(ws, l.ydp[3 * i]), at<float>(ws, l.dp[3 * i + 1]), st + l.s_dp[3 * i], tac_w(h, p, base),
                      g.PD, g.G, pps, s));
            if (xf_fwd(h, at<float>(ws, l.dp[3 * i + 1]), at<float>(ws, l.zdp[2 * i]), at<float>(ws, l.hdp[2 * i]), at<float>(ws, l.sdp[2 * i]),
                       at<float>(ws, l.dp[3 * i + 2]), xf_w(h, p, base + TAC_N), g.PD, g.G, row, s, nullptr, nullptr, nullptr))
                return 1;
            if (xf_fwd(h, at<float>(ws, l.dp[3 * i + 2]), at<float>(ws, l.zdp[2 * i + 1]), at<float>(ws, l.hdp[2 * i + 1]),
                       at<float>(ws, l.sdp[2 * i + 1]), at<float>(ws, l.dp[3 * i + 3]), xf_w(h, p, base + TAC_N + XF_N), g.PD, g.G, col, s, cw, cb, ca))
                return 1;
        }
        CK(F::group_linear(h, at<float>(ws, l.dp[3 * c.layer]), at<float>(ws, l.yout), p + h->off[P_OUT_W], p + h->off[P_OUT_B], g.PD * g.G, g.n, 0, s));
        CK(launch_overlap_add_cl(at<float>(ws, l.yout), at<float>(ws, l.fmap), g.B, g.Lc, g.K, g.S2, g.C, s));
        gc_bcast_add_kernel<<<blocks_for(g.PC * g.C), 256, 0, s>>>(at<float>(ws, l.fmap), at<float>(ws, l.ce[0]), at<float>(ws, l.cd[0]), g.PC * g.C,
                                                                   g.ctx, g.C);
        CK(cudaGetLastError());
        if (gc_rnn_fwd(h, p, HEAD + GC_BLOCK, ws, l.cd, l.ycd, l.hcd, l.scd, l.s_cd, st, g, s)) return 1;
        CK(launch_overlap_add_cl(at<float>(ws, l.cd[4]), at<float>(ws, l.feat2), g.B, g.F, g.ctx, g.Lc, g.C, s));
        CK(F::group_linear(h, at<float>(ws, l.feat2), at<float>(ws, l.Mk), p + h->off[P_MASK_W], p + h->off[P_MASK_B], (long long)g.B * g.F * g.G,
                           g.spk * g.E / g.G, 1, s));
        gc_decoder_kernel<<<blocks_for((long long)g.B * g.spk * g.T), 256, 0, s>>>(at<float>(ws, l.Mk), enc, p + h->off[P_DEC_W], est, g.B, g.spk, g.T,
                                                                                   g.F, g.E, g.G, c.win);
        CK(cudaGetLastError());
        h->launches += 9;
        return 0;
    }

    // ---------------------------------------------------------------------------------------------------------------- backward of one stage
    struct Scratch { float *T1, *T2, *T3, *S3, *A3, *S2, *Mv, *S1, *DQ, *OS; };

    // Out = X + GN(TAC(X)):  dOut -> dX
    static int tac_bwd(dp_gctasnet* h, const float* p, float* gp, int base, const float* X, const float* Y, const double* st, double* bst,
                       const float* dOut, float* dX, long long npos, int G, int pps, const Scratch& k, cudaStream_t s) {
        const TacW w = tac_w(h, p, base);
        const TacG gw = tac_g(h, gp, base);
        const int total = (int)(npos * G);
        const int tpc = gn_tiles_per_cta(total);
        gct_gn_sums_kernel<NG><<<blocks_for(blocks_for(total), tpc), 256, 0, s>>>(dOut, Y, st, w.gamma, bst, gw.gamma, gw.beta, total, G, pps, 1e-5, tpc);
        gct_gn_apply_kernel<NG><<<blocks_for(total), 256, 0, s>>>(dOut, Y, st, bst, w.gamma, k.T1, total, G, pps, 1e-5, nullptr, nullptr, 0);
        if (G == 8)
            gct_tac_bwd_kernel<NG, HG, 8><<<blocks_for(total, 128), 128, 0, s>>>(X, k.T1, dOut, dX, w, k.S3, k.A3, k.S2, k.Mv, k.S1, gw.a1, gw.a2,
                                                                                 gw.a3, (int)npos);
        else if (G == 16)
            gct_tac_bwd_kernel<NG, HG, 16><<<blocks_for(total, 128), 128, 0, s>>>(X, k.T1, dOut, dX, w, k.S3, k.A3, k.S2, k.Mv, k.S1, gw.a1, gw.a2,
                                                                                  gw.a3, (int)npos);
        else
            gct_tac_bwd_kernel<NG, HG, 32><<<blocks_for(total, 128), 128, 0, s>>>(X, k.T1, dOut, dX, w, k.S3, k.A3, k.S2, k.Mv, k.S1, gw.a1, gw.a2,
                                                                                  gw.a3, (int)npos);
        CK(cudaGetLastError());
        CK(wgrad(k.S3, NG, k.A3, 2 * TH, total, NG, 2 * TH, gw.w3, gw.b3, 0, s));
        CK(wgrad(k.S2, TH, k.Mv, TH, npos, TH, TH, gw.w2, gw.b2, 0, s));
        CK(wgrad(k.S1, TH, X, NG, total, TH, NG, gw.w1, gw.b1, 0, s));
        h->launches += 6;
        return 0;
    }
    // Out = [concat_block] (Ain + GN(Linear(BiLSTM(Ain)))):  dOut (clobbered) -> dAin
    static int rnn_bwd(dp_gctasnet* h, const float* p, float* gp, int base, const float* Ain, const float* Y, const float* Hh, const float* S5,
                       const double* st, double* bst, float* dOut, float* dAin, long long npos, int G, int pps, const SeqWalk& q, double eps,
                       bool cat, const Scratch& k, cudaStream_t s) {
        const RnnW w = rnn_w(h, p, base);
        const RnnG gw = rnn_g(h, gp, base);
        const int total = (int)(npos * G);
        if (cat)
            gct_cat_bwd_kernel<NG><<<blocks_for(total), 256, 0, s>>>(dOut, Ain, Y, st, w.gamma, w.beta, p + h->off[P_CAT_W], p + h->off[P_CAT_B],
                                                                     p + h->off[P_CAT_A], gp + h->off[P_CAT_W], gp + h->off[P_CAT_B],
                                                                     gp + h->off[P_CAT_A], total, G, pps, eps);
        const int tpc = gn_tiles_per_cta(total);
        gct_gn_sums_kernel<NG><<<blocks_for(blocks_for(total), tpc), 256, 0, s>>>(dOut, Y, st, w.gamma, bst, gw.gamma, gw.beta, total, G, pps, eps, tpc);
        // k.T1 = gradient of the projection output, k.T2 = gradient of the BiLSTM output (projection weight [NG][2 HG])
        gct_gn_apply_kernel<NG><<<blocks_for(total), 256, 0, s>>>(dOut, Y, st, bst, w.gamma, k.T1, total, G, pps, eps, w.pw, k.T2, 2 * HG);
        CK(cudaGetLastError());
        CK(wgrad(k.T1, NG, Hh, 2 * HG, total, NG, 2 * HG, gw.pw, gw.pb, 0, s));
        const dim3 grid = lstm_grid(blocks_for(q.nouter * G * HG, 128), 222);
        gct_lstm_bwd_kernel<NG, HG><<<grid, 128, 0, s>>>(Ain, Hh, S5, k.T2, k.T3, w, gw, q.nouter, G, q.len, q.qdiv, q.s_hi, q.s_lo, q.s_t,
                                                         (long long)total);
        gct_add3_kernel<<<blocks_for((long long)total * NG), 256, 0, s>>>(dOut, k.T3, k.T3 + (size_t)total * NG, dAin, (long long)total * NG);
        CK(cudaGetLastError());
        h->launches += 5 + (cat ? 1 : 0);
        return 0;
    }
    // Out = [concat_block](Ain + LN2(Z + Linear(relu(BiLSTM(Z))))), Z = LN1(Ain + out_proj(attention(in_proj(Ain)))):  dOut -> dAin
    static int xf_bwd(dp_gctasnet* h, const float* p, float* gp, int base, const float* Ain, const float* Z, const float* Hh, const float* S5,
                      const float* dOut, float* dAin, long long npos, int G, const SeqWalk& q, bool cat, const Scratch& k, cudaStream_t s) {
        const XfW w = xf_w(h, p, base);
        const RnnG gw = rnn_g(h, gp, base);
        const long long total = npos * G;
        // second half: k.S3 = gradient through the block residual, k.T1 = gradient of the LayerNorm input, k.T2 = gradient of the BiLSTM output
        gct_dpt_out_bwd_kernel<NG, HG><<<blocks_for(total), 256, 0, s>>>(Hh, Z, Ain, dOut, k.S3, k.T1, k.T2, w, gw.gamma, gw.beta, total,
                                                                        cat ? p + h->off[P_CAT_W] : nullptr, cat ? p + h->off[P_CAT_B] : nullptr,
                                                                        cat ? p + h->off[P_CAT_A] : nullptr, gp + (cat ? h->off[P_CAT_W] : 0),
                                                                        gp + (cat ? h->off[P_CAT_B] : 0), gp + (cat ? h->off[P_CAT_A] : 0));
        CK(cudaGetLastError());
        CK(wgrad(k.T1, NG, Hh, 2 * HG, total, NG, 2 * HG, gw.pw, gw.pb, 1, s));   // linear2: operand relu(Hh)
        const dim3 grid = lstm_grid(blocks_for(q.nouter * G * HG, 128), 222);
        gct_lstm_bwd_kernel<NG, HG><<<grid, 128, 0, s>>>(Z, Hh, S5, k.T2, k.T3, w.rnn, gw, q.nouter, G, q.len, q.qdiv, q.s_hi, q.s_lo, q.s_t, total);
        // gradient of Z: LayerNorm-2 residual + both LSTM directions (into k.S1)
        gct_add3_kernel<<<blocks_for(total * NG), 256, 0, s>>>(k.T1, k.T3, k.T3 + (size_t)total * NG, k.S1, total * NG);
        int spb = 32 / q.len;
        if (spb < 1) spb = 1;
        const size_t smem = (size_t)spb * q.len * (9 * NG + 12) * sizeof(float);
        if (smem > 200 * 1024) return fail("dp_gctasnet_backward: sequence of %d frames does not fit the attention kernel's shared memory", q.len);
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(gct_dpt_attn_bwd_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gct_dpt_attn_bwd_kernel<NG><<<blocks_for(q.nouter * G, spb), 128, smem, s>>>(Ain, k.S1, k.S3, dAin, k.DQ, k.T1, k.OS, w, gp + h->off[base + 16],
                                                                                     gp + h->off[base + 17], q.nouter, G, q.len, spb, q.qdiv, q.s_hi,
                                                                                     q.s_lo, q.s_t);
        CK(cudaGetLastError());
        CK(wgrad(k.DQ, 3 * NG, Ain, NG, total, 3 * NG, NG, gp + h->off[base + 12], gp + h->off[base + 13], 0, s));   // in_proj
        CK(wgrad(k.T1, NG, k.OS, NG, total, NG, NG, gp + h->off[base + 14], gp + h->off[base + 15], 0, s));           // out_proj (T1 = dz now)
        h->launches += 7;
        return 0;
    }
    // GC_RNN backward: gradient of X[4] in `gin` -> gradient of X[0] (returned pointer is gin or gout)
    static float* gc_rnn_bwd(dp_gctasnet* h, const float* p, float* gp, int base, void* ws, const size_t* X, const size_t* Y, const size_t* Hh,
                             const size_t* S5, const size_t* slots, const double* st0, double* bst0, float* gin, float* gout, const GGeo& g,
                             const Scratch& k, cudaStream_t s, int* err) {
        const SeqWalk q{(long long)g.B * g.Lc, g.ctx, 1, (long long)g.ctx, 0, 1};
        for (int i = 1; i >= 0; --i) {
            if (rnn_bwd(h, p, gp, base + i * GC_LAYER + TAC_N, at<float>(ws, X[2 * i + 1]), at<float>(ws, Y[2 * i + 1]), at<float>(ws, Hh[i]),
                        at<float>(ws, S5[i]), st0 + slots[2 * i + 1], bst0 + slots[2 * i + 1], gin, gout, g.PC, g.G, g.ctx, q, 1e-5, false, k, s)) {
                *err = 1;
                return nullptr;
            }
            if (tac_bwd(h, p, gp, base + i * GC_LAYER, at<float>(ws, X[2 * i]), at<float>(ws, Y[2 * i]), st0 + slots[2 * i], bst0 + slots[2 * i], gout,
                        gin, g.PC, g.G, g.ctx, k, s)) {
                *err = 1;
                return nullptr;
            }
        }
        return gin;   // two ping-pongs per layer: the result is back in gin
    }

    static int backward(dp_gctasnet* h, const float* p, float* gp, const float* mix, const float* d_est, void* ws, const GGeo& g, cudaStream_t s) {
        const auto& c = h->cfg;
        TLayout l;
        t_layout(h, g, l);
        const double* st = at<double>(ws, l.stats);
        double* bst = at<double>(ws, l.bst);
        float *gA = at<float>(ws, l.gA), *gB = at<float>(ws, l.gB), *dMk = at<float>(ws, l.dMk), *denc = at<float>(ws, l.denc);
        float *dframe = at<float>(ws, l.dframe), *dsq = at<float>(ws, l.dsq), *enc = at<float>(ws, l.enc), *Mk = at<float>(ws, l.Mk), *Q = at<float>(ws, l.Q);
        const Scratch k{at<float>(ws, l.T1), at<float>(ws, l.T2), at<float>(ws, l.T3), at<float>(ws, l.S3), at<float>(ws, l.A3), at<float>(ws, l.S2),
                        at<float>(ws, l.Mv), at<float>(ws, l.S1), l.DQ ? at<float>(ws, l.DQ) : nullptr, l.OS ? at<float>(ws, l.OS) : nullptr};
        h->launches = 0;
        CK(cudaMemsetAsync(bst, 0, l.stats_bytes, s));
        const long long bfe = (long long)g.B * g.F * g.E, bfg = (long long)g.B * g.F * g.G;
        const int mo = g.spk * g.E / g.G;   // mask conv outputs per group
        // decoder, masking, mask conv + ReLU
        gct_decoder_bwd_kernel<<<blocks_for(bfe), 256, 0, s>>>(d_est, Mk, enc, p + h->off[P_DEC_W], dMk, denc, Q, g.B, g.spk, g.T, g.F, g.E, g.G, c.win);
        {
            const int fchunk = 128;
            dim3 grid(ceil_div(g.F, fchunk), g.B);
            gct_decoder_wgrad_kernel<<<grid, 256, 0, s>>>(d_est, Q, gp + h->off[P_DEC_W], g.B, g.spk, g.T, g.F, g.E, c.win, fchunk);
        }
        gct_relu_mask_kernel<<<blocks_for(bfg * mo), 256, 0, s>>>(dMk, Mk, bfg * mo);
        gct_lin_bwd_kernel<<<blocks_for(bfg * NG), 256, 0, s>>>(dMk, p + h->off[P_MASK_W], nullptr, dframe, bfg, NG, mo, 0);
        CK(cudaGetLastError());
        CK(wgrad(dMk, mo, at<float>(ws, l.feat2), NG, bfg, mo, NG, gp + h->off[P_MASK_W], gp + h->off[P_MASK_B], 0, s));
        // overlap-add of the decoded context blocks -> its adjoint is the segmentation
        CK(launch_segment_cl(dframe, gA, g.B, g.F, g.ctx, g.Lc, g.C, s));
        int err = 0;
        float* gcd0 = gc_rnn_bwd(h, p, gp, HEAD + GC_BLOCK, ws, l.cd, l.ycd, l.hcd, l.scd, l.s_cd, st, bst, gA, gB, g, k, s, &err);
        if (err) return 1;
        // cd[0] = ce[0] + broadcast(fmap): gradient of ce[0] (kept in T3's second half until the encoder stage is done) and of fmap
        float* gce0_extra = at<float>(ws, l.yout);   // [PC * C] does not fit yout in general: use dedicated space below
        (void)gce0_extra;
        float* keep = at<float>(ws, l.cd[0]);        // cd[0] itself is dead now (its only reader, the first decoder TAC, is done): reuse it
        CK(cudaMemcpyAsync(keep, gcd0, (size_t)g.PC * g.C * sizeof(float), cudaMemcpyDeviceToDevice, s));
        gct_ctx_sum_kernel<<<blocks_for((long long)g.B * g.Lc * g.C), 256, 0, s>>>(gcd0, dsq, (long long)g.B * g.Lc * g.C, g.ctx, g.C);
        CK(cudaGetLastError());
        // fmap = overlap_add(yout) -> adjoint segmentation; DPRNN output conv
        float* gy = gA == gcd0 ? gB : gA;
        CK(launch_segment_cl(dsq, gy, g.B, g.Lc, g.K, g.S2, g.C, s));
        float* gx = gy == gA ? gB : gA;
        gct_lin_bwd_kernel<<<blocks_for(g.PD * g.G * NG), 256, 0, s>>>(gy, p + h->off[P_OUT_W], nullptr, gx, g.PD * g.G, NG, NG, 0);
        CK(cudaGetLastError());
        CK(wgrad(gy, NG, at<float>(ws, l.dp[3 * c.layer]), NG, g.PD * g.G, NG, NG, gp + h->off[P_OUT_W], gp + h->off[P_OUT_B], 0, s));
        const int pps = g.S2 * g.K;
        const SeqWalk row{(long long)g.B * g.S2, g.K, 1, (long long)g.K, 0, 1};
        const SeqWalk col{(long long)g.B * g.K, g.S2, g.K, (long long)g.S2 * g.K, 1, (long long)g.K};
        float *cur = gx, *oth = gy;
        for (int i = c.layer - 1; i >= 0 && c.module == DP_MODULE_DPTNET; --i) {
            const int base = HEAD + 2 * GC_BLOCK + i * DPT_LAYER;
            if (xf_bwd(h, p, gp, base + TAC_N + XF_N, at<float>(ws, l.dp[3 * i + 2]), at<float>(ws, l.zdp[2 * i + 1]), at<float>(ws, l.hdp[2 * i + 1]),
                       at<float>(ws, l.sdp[2 * i + 1]), cur, oth, g.PD, g.G, col, c.unfold != 0, k, s))
                return 1;
            if (xf_bwd(h, p, gp, base + TAC_N, at<float>(ws, l.dp[3 * i + 1]), at<float>(ws, l.zdp[2 * i]), at<float>(ws, l.hdp[2 * i]),
                       at<float>(ws, l.sdp[2 * i]), oth, cur, g.PD, g.G, row, false, k, s))
                return 1;
            if (tac_bwd(h, p, gp, base, at<float>(ws, l.dp[3 * i]), at<float>(ws, l.ydp[3 * i]), st + l.s_dp[3 * i], bst + l.s_dp[3 * i], cur, oth, g.PD,
                        g.G, pps, k, s))
                return 1;
            float* t = cur; cur = oth; oth = t;
        }
        for (int i = c.layer - 1; i >= 0 && c.module != DP_MODULE_DPTNET; --i) {
            const int base = HEAD + 2 * GC_BLOCK + i * DP_LAYER;
            if (rnn_bwd(h, p, gp, base + TAC_N + RNN_N, at<float>(ws, l.dp[3 * i + 2]), at<float>(ws, l.ydp[3 * i + 2]), at<float>(ws, l.hdp[2 * i + 1]),
                        at<float>(ws, l.sdp[2 * i + 1]), st + l.s_dp[3 * i + 2], bst + l.s_dp[3 * i + 2], cur, oth, g.PD, g.G, pps, col, 1e-8,
                        c.unfold != 0, k, s))
                return 1;
            if (rnn_bwd(h, p, gp, base + TAC_N, at<float>(ws, l.dp[3 * i + 1]), at<float>(ws, l.ydp[3 * i + 1]), at<float>(ws, l.hdp[2 * i]),
                        at<float>(ws, l.sdp[2 * i]), st + l.s_dp[3 * i + 1], bst + l.s_dp[3 * i + 1], oth, cur, g.PD, g.G, pps, row, 1e-8, false, k, s))
                return 1;
            if (tac_bwd(h, p, gp, base, at<float>(ws, l.dp[3 * i]), at<float>(ws, l.ydp[3 * i]), st + l.s_dp[3 * i], bst + l.s_dp[3 * i], cur, oth, g.PD,
                        g.G, pps, k, s))
                return 1;
            float* t = cur; cur = oth; oth = t;
        }
        // dp[0] = segmentation(sqm) -> adjoint overlap-add; sqm = mean over the context window of ce[4]
        CK(launch_overlap_add_cl(cur, dsq, g.B, g.Lc, g.K, g.S2, g.C, s));
        gct_ctx_mean_bwd_kernel<<<blocks_for(g.PC * g.C), 256, 0, s>>>(dsq, gA, g.PC * g.C, g.ctx, g.C);
        CK(cudaGetLastError());
        float* gce0 = gc_rnn_bwd(h, p, gp, HEAD, ws, l.ce, l.yce, l.hce, l.sce, l.s_ce, st, bst, gA, gB, g, k, s, &err);
        if (err) return 1;
        gct_add3_kernel<<<blocks_for(g.PC * g.C), 256, 0, s>>>(gce0, keep, nullptr, gce0, g.PC * g.C);
        // ce[0] = segmentation(feat) -> adjoint overlap-add; bottleneck; encoder
        CK(launch_overlap_add_cl(gce0, dframe, g.B, g.F, g.ctx, g.Lc, g.C, s));
        float* dU = at<float>(ws, l.Q);   // gradient of the normalised encoder output [B F E] (Q is dead: the decoder weight gradient is done)
        gct_lin_bwd_kernel<<<blocks_for(bfe), 256, 0, s>>>(dframe, p + h->off[P_BN_W], nullptr, dU, (long long)g.B * g.F, g.E, g.C, 0);
        gct_bn_norm_kernel<<<blocks_for(bfe), 256, 0, s>>>(enc, st, p + h->off[P_BN_G], p + h->off[P_BN_B], at<float>(ws, l.xn), bfe, g.F, g.E,
                                                           (double)1.1920928955078125e-07f);
        CK(cudaGetLastError());
        CK(wgrad(dframe, g.C, at<float>(ws, l.xn), g.E, (long long)g.B * g.F, g.C, g.E, gp + h->off[P_BN_W], nullptr, 0, s));
        {
            dim3 grid(ceil_div(g.F, FRAMES_PER_CTA), g.B);
            gct_bn_sums_kernel<<<grid, 256, 0, s>>>(dU, enc, st, p + h->off[P_BN_G], bst, gp + h->off[P_BN_G], gp + h->off[P_BN_B], g.F, g.E,
                                                    (double)1.1920928955078125e-07f);
            gct_bn_apply_kernel<<<blocks_for(bfe), 256, 0, s>>>(dU, enc, st, bst, p + h->off[P_BN_G], denc, bfe, g.F, g.E, (double)1.1920928955078125e-07f);
            const int fchunk = 128;
            dim3 g2(ceil_div(g.F, fchunk), g.B);
            gct_encoder_wgrad_kernel<<<g2, 256, 0, s>>>(mix, denc, gp + h->off[P_ENC_W], g.T, g.F, g.E, c.win, fchunk);
        }
        CK(cudaGetLastError());
        h->launches += 20;
        return 0;
    }
};

}  // namespace

extern "C" {

int64_t dp_gctasnet_train_workspace_bytes(const dp_gctasnet* h, int B, int T) {
    GGeo g;
    if (!h || !gc_geometry(h, B, T, g)) { fail("dp_gctasnet_train_workspace_bytes: bad arguments"); return -1; }
    TLayout l;
    t_layout(h, g, l);
    return (int64_t)l.total;
}

int dp_gctasnet_forward_train(dp_gctasnet* h, const float* params, const float* mixture, float* est, void* workspace, int B, int T, void* stream) {
    if (!h || !params || !mixture || !est || !workspace) return fail("dp_gctasnet_forward_train: null argument");
    GGeo g;
    if (!gc_geometry(h, B, T, g)) return fail("dp_gctasnet_forward_train: bad batch / length (B=%d, T=%d)", B, T);
    if (g.n == 4) return TrainOps<4, 8>::forward(h, params, mixture, est, workspace, g, S(stream));
    return TrainOps<8, 16>::forward(h, params, mixture, est, workspace, g, S(stream));
}

int dp_gctasnet_backward(dp_gctasnet* h, const float* params, float* grads, const float* mixture, const float* d_est, void* workspace, int B, int T,
                         void* stream) {
    if (!h || !params || !grads || !mixture || !d_est || !workspace) return fail("dp_gctasnet_backward: null argument");
    GGeo g;
    if (!gc_geometry(h, B, T, g)) return fail("dp_gctasnet_backward: bad batch / length (B=%d, T=%d)", B, T);
    if (g.n == 4) return TrainOps<4, 8>::backward(h, params, grads, mixture, d_est, workspace, g, S(stream));
    return TrainOps<8, 16>::backward(h, params, grads, mixture, d_est, workspace, g, S(stream));
}

}  // extern "C"
