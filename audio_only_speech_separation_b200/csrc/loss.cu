// Fused pairwise SNR / SI-SDR / SD-SDR + permutation-invariant reduction for n_src = 2 (sm_100a).
//
// Reference semantics: PairwiseNegSDR.forward (look2hear/losses/matrix.py:22-57: zero-mean, pair matrix
// [b, est, tgt], eps 1e-8 added to the target energy, to the noise energy and inside the log) and
// PITLossWrapper.forward / find_best_perm_factorial (look2hear/losses/pit_wrapper.py:30-67,96-131: best of the
// two permutations, ties -> identity, threshold_byloss keeps min_loss > -30 unless that empties the batch, mean).
//
// The reference materialises ~15 [B,2,2,T] temporaries; here the data (4*T floats per utterance) is streamed
// 2-3 times and reduced into fp64 accumulators:
//   pass 1: sums (means)              pass 2: centred dot / squared distance / target energy
//   pass 3 (SI-SDR only): noise energy of e~ - alpha t~ formed elementwise, as the reference does, so a
//          near-perfect estimate does not lose the noise term to cancellation (SURVEY 7 hard part 5).
// A one-block finalize kernel forms the pair matrix, the permutation, the threshold and the batch mean on the
// device (no host sync, unlike pit_wrapper.py:60-61,130) and emits the coefficients of the analytic gradient
//   d loss / d est[b,e,:] = A * (e - mean e) + Bc * (t_j - mean t_j),  j = target assigned to estimate e.
#include "common.cuh"
#include "kernels.h"

namespace dp {
namespace {

constexpr int LCH = 4096;  // samples per CTA

__device__ __forceinline__ void block_add(double v, double* dst, double* sh) {
    v = warp_sum_d(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
        atomicAdd(dst, s);
    }
}

__global__ void __launch_bounds__(256) loss_sums_kernel(const float* __restrict__ est, const float* __restrict__ tgt, int T, double* sums) {
    __shared__ double sh[8];
    const int b = blockIdx.y, t0 = blockIdx.x * LCH, t1 = min(T, t0 + LCH);
    const float* rows[4] = {est + (size_t)b * 2 * T, est + (size_t)b * 2 * T + T, tgt + (size_t)b * 2 * T, tgt + (size_t)b * 2 * T + T};
    for (int r = 0; r < 4; ++r) {
        float s = 0.f;
        for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) s += rows[r][t];
        block_add((double)s, sums + b * 4 + r, sh);
    }
}

__global__ void __launch_bounds__(256) loss_second_kernel(const float* __restrict__ est, const float* __restrict__ tgt, int T,
                                                          const double* __restrict__ sums, double* second) {
    __shared__ double sh[8];
    const int b = blockIdx.y, t0 = blockIdx.x * LCH, t1 = min(T, t0 + LCH);
    const float* e0 = est + (size_t)b * 2 * T;
    const float* e1 = e0 + T;
    const float* g0 = tgt + (size_t)b * 2 * T;
    const float* g1 = g0 + T;
    const float me0 = (float)(sums[b * 4 + 0] / T), me1 = (float)(sums[b * 4 + 1] / T);
    const float mt0 = (float)(sums[b * 4 + 2] / T), mt1 = (float)(sums[b * 4 + 3] / T);
    float a[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) a[i] = 0.f;
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        float x0 = e0[t] - me0, x1 = e1[t] - me1, y0 = g0[t] - mt0, y1 = g1[t] - mt1;
        a[0] = fmaf(x0, y0, a[0]); a[1] = fmaf(x0, y1, a[1]); a[2] = fmaf(x1, y0, a[2]); a[3] = fmaf(x1, y1, a[3]);
        float d00 = x0 - y0, d01 = x0 - y1, d10 = x1 - y0, d11 = x1 - y1;
        a[4] = fmaf(d00, d00, a[4]); a[5] = fmaf(d01, d01, a[5]); a[6] = fmaf(d10, d10, a[6]); a[7] = fmaf(d11, d11, a[7]);
        a[8] = fmaf(y0, y0, a[8]); a[9] = fmaf(y1, y1, a[9]);
    }
    for (int i = 0; i < 10; ++i) block_add((double)a[i], second + b * 10 + i, sh);
}

__global__ void __launch_bounds__(256) loss_noise_kernel(const float* __restrict__ est, const float* __restrict__ tgt, int T,
                                                         const double* __restrict__ sums, const double* __restrict__ second, double* noise) {
    __shared__ double sh[8];
    const int b = blockIdx.y, t0 = blockIdx.x * LCH, t1 = min(T, t0 + LCH);
    const float* e0 = est + (size_t)b * 2 * T;
    const float* e1 = e0 + T;
    const float* g0 = tgt + (size_t)b * 2 * T;
    const float* g1 = g0 + T;
    const float me0 = (float)(sums[b * 4 + 0] / T), me1 = (float)(sums[b * 4 + 1] / T);
    const float mt0 = (float)(sums[b * 4 + 2] / T), mt1 = (float)(sums[b * 4 + 3] / T);
    const double* s = second + b * 10;
    const float al00 = (float)(s[0] / (s[8] + 1e-8)), al01 = (float)(s[1] / (s[9] + 1e-8));
    const float al10 = (float)(s[2] / (s[8] + 1e-8)), al11 = (float)(s[3] / (s[9] + 1e-8));
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        float x0 = e0[t] - me0, x1 = e1[t] - me1, y0 = g0[t] - mt0, y1 = g1[t] - mt1;
        float n00 = x0 - al00 * y0, n01 = x0 - al01 * y1, n10 = x1 - al10 * y0, n11 = x1 - al11 * y1;
        a[0] = fmaf(n00, n00, a[0]); a[1] = fmaf(n01, n01, a[1]); a[2] = fmaf(n10, n10, a[2]); a[3] = fmaf(n11, n11, a[3]);
    }
    for (int i = 0; i < 4; ++i) block_add((double)a[i], noise + b * 4 + i, sh);
}

// one block; thread b handles utterance b (B <= 1024), then thread 0 reduces
__global__ void loss_finalize_kernel(int B, int sdr_type, int threshold, const double* __restrict__ second, const double* __restrict__ noise,
                                     float* __restrict__ pw, float* __restrict__ loss, int* __restrict__ perm, float* __restrict__ coef) {
    extern __shared__ float minl[];  // [B]
    const double EPS = 1e-8, K10 = 10.0 / log(10.0);
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const double* s = second + b * 10;
        double v[2][2], A[2][2], Bc[2][2];
        for (int e = 0; e < 2; ++e)
            for (int j = 0; j < 2; ++j) {
                double dot = s[e * 2 + j], d2 = s[4 + e * 2 + j], tt = s[8 + j];
                double ratio, a_, b_;
                if (sdr_type == 0) {  // snr: proj = t, noise = e - t
                    double den = d2 + EPS;
                    ratio = tt / den;
                    double k = -K10 / (ratio + EPS);
                    double dr_dd2 = -tt / (den * den);
                    a_ = k * dr_dd2 * 2.0;
                    b_ = -a_;
                } else {
                    double te = tt + EPS, alpha = dot / te;
                    double S = alpha * alpha * tt;
                    double dS = 2.0 * dot * tt / (te * te);  // coefficient on t~
                    if (sdr_type == 1) {  // sisdr: noise = e - alpha t
                        double nn = noise[b * 4 + e * 2 + j], den = nn + EPS;
                        ratio = S / den;
                        double k = -K10 / (ratio + EPS);
                        double rem = dot - alpha * tt;
                        a_ = k * (-S / (den * den)) * 2.0;
                        b_ = k * (dS / den + S / (den * den) * 2.0 * (alpha + rem / te));
                    } else {  // sdsdr: noise = e - t
                        double den = d2 + EPS;
                        ratio = S / den;
                        double k = -K10 / (ratio + EPS);
                        a_ = k * (-2.0 * S / (den * den));
                        b_ = k * (dS / den + 2.0 * S / (den * den));
                    }
                }
                v[e][j] = -10.0 * log10(ratio + EPS);
                A[e][j] = a_;
                Bc[e][j] = b_;
                pw[b * 4 + e * 2 + j] = (float)v[e][j];
            }
        // perm 0: est e -> tgt e ; perm 1: est 1 -> tgt 0, est 0 -> tgt 1   (pit_wrapper.py:105-131)
        float l0 = ((float)v[0][0] + (float)v[1][1]) * 0.5f;
        float l1 = ((float)v[1][0] + (float)v[0][1]) * 0.5f;
        int p = l1 < l0 ? 1 : 0;
        perm[b] = p;
        minl[b] = p ? l1 : l0;
        for (int e = 0; e < 2; ++e) {
            int j = p ? 1 - e : e;
            coef[(b * 2 + e) * 3 + 0] = (float)A[e][j];
            coef[(b * 2 + e) * 3 + 1] = (float)Bc[e][j];
            coef[(b * 2 + e) * 3 + 2] = (float)j;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int kept = 0;
        if (threshold)
            for (int b = 0; b < B; ++b) kept += (minl[b] > -30.f);
        const bool filter = threshold && kept > 0;
        double acc = 0.0;
        int cnt = 0;
        for (int b = 0; b < B; ++b) {
            bool use = !filter || (minl[b] > -30.f);
            if (use) { acc += minl[b]; ++cnt; }
        }
        loss[0] = (float)(acc / cnt);
        const float w = 1.0f / (2.0f * (float)cnt);  // mean over kept utterances and over n_src
        for (int b = 0; b < B; ++b) {
            bool use = !filter || (minl[b] > -30.f);
            float s = use ? w : 0.f;
            for (int e = 0; e < 2; ++e) {  // scale A and Bc of both estimates, keep the target index
                coef[b * 6 + e * 3 + 0] *= s;
                coef[b * 6 + e * 3 + 1] *= s;
            }
        }
    }
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(const float* __restrict__ est, const float* __restrict__ tgt, int T,
                                                       const double* __restrict__ sums, const float* __restrict__ coef, float gscale,
                                                       float* __restrict__ d_est) {
    const int be = blockIdx.y;  // b*2 + e
    const int b = be >> 1, e = be & 1;
    const float A = coef[be * 3] * gscale, Bc = coef[be * 3 + 1] * gscale;
    const int j = (int)coef[be * 3 + 2];
    const float me = (float)(sums[b * 4 + e] / T), mt = (float)(sums[b * 4 + 2 + j] / T);
    const float* er = est + (size_t)be * T;
    const float* tr = tgt + ((size_t)b * 2 + j) * T;
    float* dr = d_est + (size_t)be * T;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) dr[t] = fmaf(A, er[t] - me, Bc * (tr[t] - mt));
}

__global__ void reorder_kernel(const float* __restrict__ est, const int* __restrict__ perm, float* __restrict__ out, int T) {
    const int bi = blockIdx.y, b = bi >> 1, i = bi & 1;
    const int src = perm[b] ? 1 - i : i;
    const float* s = est + ((size_t)b * 2 + src) * T;
    float* d = out + (size_t)bi * T;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) d[t] = s[t];
}

}  // namespace

cudaError_t launch_pit_loss_fwd(const float* est, const float* tgt, int B, int T, int sdr_type, int threshold_byloss, const PitLossWs& ws,
                                float* pw, float* loss, int* perm, float* coef, cudaStream_t st) {
    if (B <= 0 || T <= 0 || B > 65535) return cudaErrorInvalidValue;
    cudaError_t e;
    if ((e = cudaMemsetAsync(ws.sums, 0, sizeof(double) * 4 * B, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ws.second, 0, sizeof(double) * 10 * B, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ws.noise, 0, sizeof(double) * 4 * B, st)) != cudaSuccess) return e;
    dim3 grid(ceil_div(T, LCH), B);
    loss_sums_kernel<<<grid, 256, 0, st>>>(est, tgt, T, ws.sums);
    loss_second_kernel<<<grid, 256, 0, st>>>(est, tgt, T, ws.sums, ws.second);
    if (sdr_type == 1) loss_noise_kernel<<<grid, 256, 0, st>>>(est, tgt, T, ws.sums, ws.second, ws.noise);
    loss_finalize_kernel<<<1, 256, sizeof(float) * B, st>>>(B, sdr_type, threshold_byloss, ws.second, ws.noise, pw, loss, perm, coef);
    return cudaGetLastError();
}

cudaError_t launch_pit_loss_bwd(const float* est, const float* tgt, int B, int T, const double* sums, const float* coef, float grad_scale,
                                float* d_est, cudaStream_t st) {
    if (B <= 0 || T <= 0) return cudaErrorInvalidValue;
    dim3 grid(min(ceil_div(T, 256), 64), 2 * B);
    loss_bwd_kernel<<<grid, 256, 0, st>>>(est, tgt, T, sums, coef, grad_scale, d_est);
    return cudaGetLastError();
}

cudaError_t launch_reorder_sources(const float* est, const int* perm, float* out, int B, int T, cudaStream_t st) {
    if (B <= 0 || T <= 0) return cudaErrorInvalidValue;
    dim3 grid(min(ceil_div(T, 256), 64), 2 * B);
    reorder_kernel<<<grid, 256, 0, st>>>(est, perm, out, T);
    return cudaGetLastError();
}

}  // namespace dp
