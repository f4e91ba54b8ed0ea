// Gradient-norm clipping + Adam, fused over one flat parameter buffer (sm_100a).
// Semantics of torch.nn.utils.clip_grad_norm_(params, 5.0) followed by torch.optim.Adam(lr, betas, eps, wd)
// as used by the reference training step (audio_train.py:48,128; system/optimizers.py:58-75).
// One pass to reduce sum(g^2) into fp64, one pass that reads (p, g, m, v) and writes (p, m, v): 7 floats of HBM
// traffic per parameter.  The flat buffer is also what the single NCCL gradient all-reduce operates on.
#include "common.cuh"
#include "kernels.h"

namespace dp {
namespace {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* out) {
    __shared__ double sh[8];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = g[i];
        s += (double)v * (double)v;
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        atomicAdd(out, t);
    }
}

__global__ void __launch_bounds__(256) adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, const double* __restrict__ norm2, float gscale,
                                                        float max_norm, float lr, float b1, float b2, float eps, float bc1, float bc2,
                                                        float wd, const float* __restrict__ hyper) {
    if (hyper) {   // (lr, 1 - b1^t, 1 - b2^t) from device memory: a captured step (CUDA graph) must not bake them into the launch
        lr = hyper[0]; bc1 = hyper[1]; bc2 = hyper[2];
    }
    float coef = gscale;
    if (max_norm > 0.f) {
        float total = (float)sqrt(*norm2) * gscale;
        coef *= fminf(1.0f, max_norm / (total + 1e-6f));
    }
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float pi = p[i];
        float gi = fmaf(wd, pi, g[i] * coef);
        float mi = fmaf(1.f - b1, gi - m[i], m[i]);          // lerp(m, g, 1 - b1)
        float vi = fmaf(1.f - b2, gi * gi, b2 * v[i]);
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

__global__ void set_hyper_kernel(float* hyper, float a, float b, float c) {
    hyper[0] = a; hyper[1] = b; hyper[2] = c;
}

}  // namespace

cudaError_t launch_set_hyper(float* hyper, float lr, float bc1, float bc2, cudaStream_t st) {
    set_hyper_kernel<<<1, 1, 0, st>>>(hyper, lr, bc1, bc2);
    return cudaGetLastError();
}

cudaError_t launch_sumsq(const float* g, long long n, double* out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    long long blocks = ceil_div_ll(n, 256 * 8);
    if (blocks > 148 * 4) blocks = 148 * 4;
    sumsq_kernel<<<(unsigned)blocks, 256, 0, st>>>(g, n, out);
    return cudaGetLastError();
}

cudaError_t launch_adam_clip(float* p, const float* g, float* m, float* v, long long n, const double* norm2, float gscale, float max_norm,
                             float lr, float b1, float b2, float eps, float bc1, float bc2, float wd, cudaStream_t st, const float* hyper) {
    if (n <= 0) return cudaSuccess;
    long long blocks = ceil_div_ll(n, 256 * 4);
    if (blocks > 148 * 8) blocks = 148 * 8;
    adam_clip_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, norm2, gscale, max_norm, lr, b1, b2, eps, bc1, bc2, wd, hyper);
    return cudaGetLastError();
}

}  // namespace dp
