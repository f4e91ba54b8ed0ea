// BSS-eval SDR with a 512-tap distortion filter and permutation solving, on the GPU (sm_100a).
//
// Replaces fast_bss_eval.sdr_pit_loss as called by the reference's MetricsTracker (look2hear/metrics/wrapper.py:38-41).  fast_bss_eval is a
// third-party dependency that is not vendored in the reference; this file restates its published algorithm (R. Scheibler, "SDR -- Medium Rare
// with Fast Computations", ICASSP 2022; fast_bss_eval.sdr with its defaults filter_length = 512, zero_mean = False, direct solve):
//   1. est, ref rows scaled to unit L2 norm
//   2. acf_i[l]    = sum_t ref_i[t] ref_i[t + l]          (what the package gets from |rfft|^2 with n_fft >= T + L: the linear correlation)
//      xcorr_ij[l] = sum_t ref_i[t] est_j[t + l],  l = 0 .. L-1
//   3. Toeplitz(acf_i) h_ij = xcorr_ij                     (the L-tap filter that projects est_j on the shifts of ref_i)
//   4. coh_ij = <xcorr_ij, h_ij>;  SDR_ij = 10 log10(coh / (1 - coh))     (est_j has unit norm: projected energy = coh)
//   5. permutation of the estimates that maximises the mean SDR; output = that mean (= -sdr_pit_loss(est, ref).mean())
// Everything after the input rows is fp64: the Toeplitz systems of speech autocorrelations are badly conditioned.  The solve is the Levinson
// recursion (O(L^2) per system, one CTA per reference with all right-hand sides).
#include "common.cuh"
#include "kernels.h"

namespace dp {
namespace {

constexpr int SDR_MAXSRC = 4;
constexpr int SDR_CH = 2048;   // samples of the reference per CTA of the correlation kernel

__global__ void __launch_bounds__(256) sdr_norm_kernel(const float* __restrict__ x, int T, double* __restrict__ inv_norm) {
    __shared__ double sh[8];
    const float* row = x + (size_t)blockIdx.x * T;
    double s = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) s += (double)row[t] * (double)row[t];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += sh[w];
        inv_norm[blockIdx.x] = 1.0 / fmax(sqrt(a), 1e-12);
    }
}

// corr[b][i][k][l] += sum over this CTA's chunk of ref_i[t] * y_k[t + l];  k = 0: y = ref_i (autocorrelation), k = 1 + j: y = est_j.
// grid (chunks, n * (n + 1), B), block = L threads (one lag each); both rows staged in shared memory.
__global__ void __launch_bounds__(512) sdr_corr_kernel(const float* __restrict__ est, const float* __restrict__ ref, int n, int T, int L,
                                                       const double* __restrict__ inv_ref, const double* __restrict__ inv_est,
                                                       double* __restrict__ corr) {
    extern __shared__ float sm[];   // x[SDR_CH] | y[SDR_CH + L]
    float* xs = sm;
    float* ys = sm + SDR_CH;
    const int b = blockIdx.z, i = blockIdx.y / (n + 1), k = blockIdx.y % (n + 1);
    const int t0 = blockIdx.x * SDR_CH;
    const float* x = ref + ((size_t)b * n + i) * T;
    const float* y = k == 0 ? x : est + ((size_t)b * n + (k - 1)) * T;
    for (int t = threadIdx.x; t < SDR_CH; t += blockDim.x) xs[t] = (t0 + t < T) ? x[t0 + t] : 0.f;
    for (int t = threadIdx.x; t < SDR_CH + L; t += blockDim.x) ys[t] = (t0 + t < T) ? y[t0 + t] : 0.f;
    __syncthreads();
    const int l = threadIdx.x;
    if (l >= L) return;
    double acc = 0.0;
    for (int t = 0; t < SDR_CH; t += 8) {   // fp32 products of 8 samples, fp64 across groups
        float p = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) p = fmaf(xs[t + u], ys[t + u + l], p);
        acc += (double)p;
    }
    const double sc = inv_ref[b * n + i] * (k == 0 ? inv_ref[b * n + i] : inv_est[b * n + (k - 1)]);
    atomicAdd(corr + (((size_t)b * n + i) * (n + 1) + k) * L + l, acc * sc);
}

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
    v = warp_sum_d(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
    return s;
}

// One CTA per (b, i): Levinson recursion for Toeplitz(acf_i) with the n right-hand sides xcorr_ij (Golub & Van Loan, Alg. 4.7.2 with the
// Durbin predictor a, a[0] = 1):   lam = -(r[k] + sum_{m=1}^{k-1} a[m] r[k-m]) / E;  a'[m] = a[m] + lam a[k-m], a'[k] = lam;  E *= 1 - lam^2;
// mu_j = (b_j[k] - sum_{m<k} x_j[m] r[k-m]) / E;  x_j[m] += mu_j a'[k-m], x_j[k] = mu_j.   Output coh[b][i][j] = <xcorr_ij, x_j>.
__global__ void __launch_bounds__(512) sdr_levinson_kernel(const double* __restrict__ corr, int n, int L, double* __restrict__ coh) {
    extern __shared__ double sd[];   // r[L] | a[L] | an[L] | x[n][L] | red[16]
    double* r = sd;
    double* a = r + L;
    double* an = a + L;
    double* x = an + L;
    double* red = x + (size_t)n * L;
    const int bi = blockIdx.x, tid = threadIdx.x;
    const double* base = corr + (size_t)bi * (n + 1) * L;
    for (int l = tid; l < L; l += blockDim.x) {
        r[l] = base[l];
        a[l] = 0.0;
        for (int j = 0; j < n; ++j) x[j * L + l] = 0.0;
    }
    __syncthreads();
    double E = r[0];
    if (tid == 0) {
        a[0] = 1.0;
        for (int j = 0; j < n; ++j) x[j * L] = base[(1 + j) * L] / E;
    }
    __syncthreads();
    for (int k = 1; k < L; ++k) {
        double p = 0.0;
        for (int m = 1 + tid; m < k; m += blockDim.x) p += a[m] * r[k - m];
        const double lam = -(r[k] + block_sum_d(p, red)) / E;
        for (int m = 1 + tid; m < k; m += blockDim.x) an[m] = a[m] + lam * a[k - m];
        if (tid == 0) { an[k] = lam; an[0] = 1.0; }
        E *= (1.0 - lam * lam);
        __syncthreads();
        for (int j = 0; j < n; ++j) {
            double q = 0.0;
            for (int m = tid; m < k; m += blockDim.x) q += x[j * L + m] * r[k - m];
            const double mu = (base[(1 + j) * L + k] - block_sum_d(q, red)) / E;
            for (int m = tid; m < k; m += blockDim.x) x[j * L + m] += mu * an[k - m];
            if (tid == 0) x[j * L + k] = mu;
        }
        for (int m = tid; m <= k; m += blockDim.x) a[m] = an[m];
        __syncthreads();
    }
    for (int j = 0; j < n; ++j) {
        double q = 0.0;
        for (int m = tid; m < L; m += blockDim.x) q += x[j * L + m] * base[(1 + j) * L + m];
        const double c = block_sum_d(q, red);
        if (tid == 0) coh[(size_t)bi * n + j] = c;
    }
}

// thread b: SDR matrix [ref i][est j], best assignment of estimates to references (maximal mean SDR), mean SDR of it
__global__ void sdr_finalize_kernel(int B, int n, const double* __restrict__ coh, float* __restrict__ out, float* __restrict__ sdr_mat) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double eps = 1.1920928955078125e-07;   // float32 machine epsilon: the coherence is kept inside (eps, 1 - eps)
    double s[SDR_MAXSRC][SDR_MAXSRC];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const double c = fmin(fmax(coh[((size_t)b * n + i) * n + j], eps), 1.0 - eps);
            s[i][j] = 10.0 * log10(c / (1.0 - c));
            if (sdr_mat) sdr_mat[((size_t)b * n + i) * n + j] = (float)s[i][j];
        }
    int p[SDR_MAXSRC] = {0, 1, 2, 3};
    double best = -1e300;
    int nperm = 1;
    for (int i = 2; i <= n; ++i) nperm *= i;
    for (int q = 0; q < nperm; ++q) {   // q-th permutation in lexicographic order
        int avail[SDR_MAXSRC] = {0, 1, 2, 3}, rem = q, fact = nperm, cnt = n;
        for (int i = 0; i < n; ++i) {
            fact /= cnt;
            const int idx = rem / fact;
            rem -= idx * fact;
            p[i] = avail[idx];
            for (int m = idx; m + 1 < cnt; ++m) avail[m] = avail[m + 1];
            --cnt;
        }
        double tot = 0.0;
        for (int i = 0; i < n; ++i) tot += s[i][p[i]];
        if (tot > best) best = tot;
    }
    out[b] = (float)(best / n);
}

}  // namespace

size_t bss_sdr_workspace_bytes(int B, int n, int L) { return sizeof(double) * ((size_t)2 * B * n + (size_t)B * n * (n + 1) * L + (size_t)B * n * n) + 64; }

cudaError_t launch_bss_sdr_pit(const float* est, const float* ref, int B, int n, int T, int L, void* ws, float* out, float* sdr_mat, cudaStream_t st) {
    if (B <= 0 || n < 1 || n > SDR_MAXSRC || T <= 0 || L < 1 || L > 512) return cudaErrorInvalidValue;
    double* inv_ref = static_cast<double*>(ws);
    double* inv_est = inv_ref + (size_t)B * n;
    double* corr = inv_est + (size_t)B * n;
    double* coh = corr + (size_t)B * n * (n + 1) * L;
    cudaError_t e = cudaMemsetAsync(corr, 0, sizeof(double) * (size_t)B * n * (n + 1) * L, st);
    if (e != cudaSuccess) return e;
    sdr_norm_kernel<<<B * n, 256, 0, st>>>(ref, T, inv_ref);
    sdr_norm_kernel<<<B * n, 256, 0, st>>>(est, T, inv_est);
    dim3 grid(ceil_div(T, SDR_CH), n * (n + 1), B);
    sdr_corr_kernel<<<grid, 512, sizeof(float) * (2 * SDR_CH + L), st>>>(est, ref, n, T, L, inv_ref, inv_est, corr);
    const size_t lsm = sizeof(double) * ((size_t)(3 + n) * L + 16);
    e = cudaFuncSetAttribute(sdr_levinson_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm);
    if (e != cudaSuccess) return e;
    sdr_levinson_kernel<<<B * n, 512, lsm, st>>>(corr, n, L, coh);
    sdr_finalize_kernel<<<ceil_div(B, 128), 128, 0, st>>>(B, n, coh, out, sdr_mat);
    return cudaGetLastError();
}

}  // namespace dp
