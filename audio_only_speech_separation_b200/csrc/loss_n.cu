// Fused pairwise SNR / SI-SDR / SD-SDR + permutation-invariant reduction for n_src = 1 .. 4 (sm_100a).
//
// Same arithmetic as loss.cu (the n_src = 2 kernels every config of the reference uses), for the remaining settings of
// PITLossWrapper (look2hear/losses/pit_wrapper.py:30-131):
//   pit_from = "pw_mtx"   PairwiseNegSDR.forward            (matrix.py:22-57)    pair matrix [b, est, tgt]
//   pit_from = "pw_pt"    SingleSrcNegSDR per (est, tgt)    (matrix.py:75-106, pit_wrapper.py:69-77): the same pair matrix
//   pit_from = "perm_avg" MultiSrcNegSDR(ests[:, perm], t)  (matrix.py:119-152, pit_wrapper.py:79-88): mean over targets of the same entries,
//                         no threshold_byloss
// and find_best_perm_factorial (pit_wrapper.py:106-131): permutations in itertools order, loss of a permutation = fp32 sum over targets i of
// pw[perm[i], i], divided by n_src, first minimum wins; batch_indices[b] = that permutation (estimate index per target).
// Data (2 n T floats per utterance) is streamed 2-3 times into fp64 accumulators as in loss.cu; the finalize kernel also emits the
// coefficients of the analytic gradient  d loss / d est[b,e,:] = A (e - mean e) + Bc (t_j - mean t_j),  j = target assigned to estimate e.
#include "common.cuh"
#include "kernels.h"

namespace dp {
namespace {

constexpr int NCH = 4096;  // samples per CTA

__device__ __forceinline__ void block_add_n(double v, double* dst, double* sh) {
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

// sums[b][2N]: estimates then targets
template <int N>
__global__ void __launch_bounds__(256) lossn_sums_kernel(const float* __restrict__ est, const float* __restrict__ tgt, int T, double* sums) {
    __shared__ double sh[8];
    const int b = blockIdx.y, t0 = blockIdx.x * NCH, t1 = min(T, t0 + NCH);
    for (int r = 0; r < 2 * N; ++r) {
        const float* row = (r < N ? est : tgt) + ((size_t)b * N + (r % N)) * T;
        float s = 0.f;
        for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) s += row[t];
        block_add_n((double)s, sums + b * 2 * N + r, sh);
    }
}

// second[b]: dot[e][j] (N*N), squared distance d2[e][j] (N*N), target energy tt[j] (N)
template <int N>
__global__ void __launch_bounds__(256) lossn_second_kernel(const float* __restrict__ est, const float* __restrict__ tgt, int T,
                                                           const double* __restrict__ sums, double* second) {
    __shared__ double sh[8];
    const int b = blockIdx.y, t0 = blockIdx.x * NCH, t1 = min(T, t0 + NCH);
    float me[N], mt[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        me[i] = (float)(sums[b * 2 * N + i] / T);
        mt[i] = (float)(sums[b * 2 * N + N + i] / T);
    }
    float a[2 * N * N + N];
#pragma unroll
    for (int i = 0; i < 2 * N * N + N; ++i) a[i] = 0.f;
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        float x[N], y[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            x[i] = est[((size_t)b * N + i) * T + t] - me[i];
            y[i] = tgt[((size_t)b * N + i) * T + t] - mt[i];
        }
#pragma unroll
        for (int e = 0; e < N; ++e)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                a[e * N + j] = fmaf(x[e], y[j], a[e * N + j]);
                const float d = x[e] - y[j];
                a[N * N + e * N + j] = fmaf(d, d, a[N * N + e * N + j]);
            }
#pragma unroll
        for (int j = 0; j < N; ++j) a[2 * N * N + j] = fmaf(y[j], y[j], a[2 * N * N + j]);
    }
    for (int i = 0; i < 2 * N * N + N; ++i) block_add_n((double)a[i], second + b * (2 * N * N + N) + i, sh);
}

// SI-SDR only: noise energy of e~ - alpha t~ formed elementwise, as the reference does
template <int N>
__global__ void __launch_bounds__(256) lossn_noise_kernel(const float* __restrict__ est, const float* __restrict__ tgt, int T,
                                                          const double* __restrict__ sums, const double* __restrict__ second, double* noise) {
    __shared__ double sh[8];
    const int b = blockIdx.y, t0 = blockIdx.x * NCH, t1 = min(T, t0 + NCH);
    const double* s = second + b * (2 * N * N + N);
    float me[N], mt[N], al[N * N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        me[i] = (float)(sums[b * 2 * N + i] / T);
        mt[i] = (float)(sums[b * 2 * N + N + i] / T);
    }
#pragma unroll
    for (int e = 0; e < N; ++e)
#pragma unroll
        for (int j = 0; j < N; ++j) al[e * N + j] = (float)(s[e * N + j] / (s[2 * N * N + j] + 1e-8));
    float a[N * N];
#pragma unroll
    for (int i = 0; i < N * N; ++i) a[i] = 0.f;
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        float x[N], y[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            x[i] = est[((size_t)b * N + i) * T + t] - me[i];
            y[i] = tgt[((size_t)b * N + i) * T + t] - mt[i];
        }
#pragma unroll
        for (int e = 0; e < N; ++e)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const float n = x[e] - al[e * N + j] * y[j];
                a[e * N + j] = fmaf(n, n, a[e * N + j]);
            }
    }
    for (int i = 0; i < N * N; ++i) block_add_n((double)a[i], noise + b * N * N + i, sh);
}

// k-th permutation of 0..N-1 in itertools.permutations (lexicographic) order
template <int N>
__device__ __forceinline__ void nth_perm(int k, int (&p)[N]) {
    int avail[N];
#pragma unroll
    for (int i = 0; i < N; ++i) avail[i] = i;
    int fact = 1;
#pragma unroll
    for (int i = 2; i < N; ++i) fact *= i;   // (N-1)!
    int n = N;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int idx = k / fact;
        k -= idx * fact;
        p[i] = avail[idx];
        for (int m = idx; m + 1 < n; ++m) avail[m] = avail[m + 1];
        --n;
        if (n > 1) fact /= n;
    }
}

// one block; thread b handles utterance b, then thread 0 reduces.  coef[b][e] = (A, Bc, target index)
template <int N>
__global__ void lossn_finalize_kernel(int B, int sdr_type, int threshold, const double* __restrict__ second, const double* __restrict__ noise,
                                      float* __restrict__ pw, float* __restrict__ loss, int* __restrict__ perm, float* __restrict__ coef) {
    extern __shared__ float minl[];  // [B]
    const double EPS = 1e-8, K10 = 10.0 / log(10.0);
    int nperm = 1;
#pragma unroll
    for (int i = 2; i <= N; ++i) nperm *= i;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const double* s = second + b * (2 * N * N + N);
        float v[N][N];
        double A[N][N], Bc[N][N];
        for (int e = 0; e < N; ++e)
            for (int j = 0; j < N; ++j) {
                const double dot = s[e * N + j], d2 = s[N * N + e * N + j], tt = s[2 * N * N + j];
                double ratio, a_, b_;
                if (sdr_type == 0) {  // snr: proj = t, noise = e - t
                    const double den = d2 + EPS;
                    ratio = tt / den;
                    const double k = -K10 / (ratio + EPS);
                    a_ = k * (-tt / (den * den)) * 2.0;
                    b_ = -a_;
                } else {
                    const double te = tt + EPS, alpha = dot / te;
                    const double S = alpha * alpha * tt;
                    const double dS = 2.0 * dot * tt / (te * te);
                    if (sdr_type == 1) {  // sisdr: noise = e - alpha t
                        const double nn = noise[b * N * N + e * N + j], den = nn + EPS;
                        ratio = S / den;
                        const double k = -K10 / (ratio + EPS);
                        const double rem = dot - alpha * tt;
                        a_ = k * (-S / (den * den)) * 2.0;
                        b_ = k * (dS / den + S / (den * den) * 2.0 * (alpha + rem / te));
                    } else {  // sdsdr: noise = e - t
                        const double den = d2 + EPS;
                        ratio = S / den;
                        const double k = -K10 / (ratio + EPS);
                        a_ = k * (-2.0 * S / (den * den));
                        b_ = k * (dS / den + 2.0 * S / (den * den));
                    }
                }
                v[e][j] = (float)(-10.0 * log10(ratio + EPS));
                A[e][j] = a_;
                Bc[e][j] = b_;
                pw[(b * N + e) * N + j] = v[e][j];
            }
        // find_best_perm_factorial: loss of permutation p = (sum over targets i of pw[p[i]][i]) / N, first minimum wins
        float best = 0.f;
        int bestp[N];
        for (int k = 0; k < nperm; ++k) {
            int p[N];
            nth_perm<N>(k, p);
            float l = 0.f;
            for (int i = 0; i < N; ++i) l += v[p[i]][i];
            l /= (float)N;
            if (k == 0 || l < best) {
                best = l;
                for (int i = 0; i < N; ++i) bestp[i] = p[i];
            }
        }
        minl[b] = best;
        for (int i = 0; i < N; ++i) {
            perm[b * N + i] = bestp[i];
            const int e = bestp[i];   // estimate e is paired with target i
            coef[(b * N + e) * 3 + 0] = (float)A[e][i];
            coef[(b * N + e) * 3 + 1] = (float)Bc[e][i];
            coef[(b * N + e) * 3 + 2] = (float)i;
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
        for (int b = 0; b < B; ++b)
            if (!filter || (minl[b] > -30.f)) { acc += minl[b]; ++cnt; }
        loss[0] = (float)(acc / cnt);
        const float w = 1.0f / ((float)N * (float)cnt);  // mean over kept utterances and over n_src
        for (int b = 0; b < B; ++b) {
            const float sc = (!filter || (minl[b] > -30.f)) ? w : 0.f;
            for (int e = 0; e < N; ++e) {
                coef[(b * N + e) * 3 + 0] *= sc;
                coef[(b * N + e) * 3 + 1] *= sc;
            }
        }
    }
}

__global__ void __launch_bounds__(256) lossn_bwd_kernel(const float* __restrict__ est, const float* __restrict__ tgt, int N, int T,
                                                        const double* __restrict__ sums, const float* __restrict__ coef, float gscale,
                                                        float* __restrict__ d_est) {
    const int be = blockIdx.y;  // b*N + e
    const int b = be / N, e = be % N;
    const float A = coef[be * 3] * gscale, Bc = coef[be * 3 + 1] * gscale;
    const int j = (int)coef[be * 3 + 2];
    const float me = (float)(sums[b * 2 * N + e] / T), mt = (float)(sums[b * 2 * N + N + j] / T);
    const float* er = est + (size_t)be * T;
    const float* tr = tgt + ((size_t)b * N + j) * T;
    float* dr = d_est + (size_t)be * T;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) dr[t] = fmaf(A, er[t] - me, Bc * (tr[t] - mt));
}

__global__ void reorder_n_kernel(const float* __restrict__ est, const int* __restrict__ perm, float* __restrict__ out, int N, int T) {
    const int bi = blockIdx.y, b = bi / N, i = bi % N;
    const float* s = est + ((size_t)b * N + perm[b * N + i]) * T;
    float* d = out + (size_t)bi * T;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) d[t] = s[t];
}

template <int N>
cudaError_t fwd_n(const float* est, const float* tgt, int B, int T, int sdr_type, int threshold_byloss, const PitLossWs& ws, float* pw, float* loss,
                  int* perm, float* coef, cudaStream_t st) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(ws.sums, 0, sizeof(double) * 2 * N * B, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ws.second, 0, sizeof(double) * (2 * N * N + N) * B, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ws.noise, 0, sizeof(double) * N * N * B, st)) != cudaSuccess) return e;
    dim3 grid(ceil_div(T, NCH), B);
    lossn_sums_kernel<N><<<grid, 256, 0, st>>>(est, tgt, T, ws.sums);
    lossn_second_kernel<N><<<grid, 256, 0, st>>>(est, tgt, T, ws.sums, ws.second);
    if (sdr_type == 1) lossn_noise_kernel<N><<<grid, 256, 0, st>>>(est, tgt, T, ws.sums, ws.second, ws.noise);
    lossn_finalize_kernel<N><<<1, 256, sizeof(float) * B, st>>>(B, sdr_type, threshold_byloss, ws.second, ws.noise, pw, loss, perm, coef);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_pitn_loss_fwd(const float* est, const float* tgt, int B, int N, int T, int sdr_type, int threshold_byloss, const PitLossWs& ws,
                                 float* pw, float* loss, int* perm, float* coef, cudaStream_t st) {
    if (B <= 0 || T <= 0 || B > 12000) return cudaErrorInvalidValue;   // minl[B] lives in the finalize kernel's shared memory
    switch (N) {
        case 1: return fwd_n<1>(est, tgt, B, T, sdr_type, threshold_byloss, ws, pw, loss, perm, coef, st);
        case 2: return fwd_n<2>(est, tgt, B, T, sdr_type, threshold_byloss, ws, pw, loss, perm, coef, st);
        case 3: return fwd_n<3>(est, tgt, B, T, sdr_type, threshold_byloss, ws, pw, loss, perm, coef, st);
        case 4: return fwd_n<4>(est, tgt, B, T, sdr_type, threshold_byloss, ws, pw, loss, perm, coef, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_pitn_loss_bwd(const float* est, const float* tgt, int B, int N, int T, const double* sums, const float* coef, float grad_scale,
                                 float* d_est, cudaStream_t st) {
    if (B <= 0 || T <= 0 || N < 1 || N > 4) return cudaErrorInvalidValue;
    dim3 grid(min(ceil_div(T, 256), 64), N * B);
    lossn_bwd_kernel<<<grid, 256, 0, st>>>(est, tgt, N, T, sums, coef, grad_scale, d_est);
    return cudaGetLastError();
}

cudaError_t launch_reorder_sources_n(const float* est, const int* perm, float* out, int B, int N, int T, cudaStream_t st) {
    if (B <= 0 || T <= 0 || N < 1 || N > 4) return cudaErrorInvalidValue;
    dim3 grid(min(ceil_div(T, 256), 64), N * B);
    reorder_n_kernel<<<grid, 256, 0, st>>>(est, perm, out, N, T);
    return cudaGetLastError();
}

}  // namespace dp
