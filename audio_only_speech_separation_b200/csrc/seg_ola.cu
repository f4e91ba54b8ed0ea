// 50%-overlap segmentation and overlap-add (gather / scatter-add), sm_100a.
//
// Reference semantics: look2hear/models/utils/gc3_basics.py:63-91 (pad_segment + split_feature) and :94-109
// (merge_feature); identical maps in look2hear/models/sepformer.py:762-846.  Closed forms,
// with P = K/2 (K even):
//     segment : y[r,k,s] = x[r,(s-1)P + k]  if 0 <= (s-1)P+k < L else 0
//     ola     : x[r,t]   = y[r,(t+P)%K, 2*((t+P)/K)] + y[r,t%K, 2*(t/K)+1]
// Pure HBM-bound index work: every byte is read once and written once, one fp32 add per output element in the
// overlap-add (commutative -> bit-exact against the reference).
//
// Two layouts:
//   * "nchw"  [rows=B*N, K, S] <-> [rows, L]: the reference's public layout.  The innermost output index and the
//     innermost input index differ, so tiles are staged through shared memory and both the global reads and
//     the global writes are issued as full 128-byte warp transactions.
//   * "cl" (channels-last) [B,S,K,C] <-> [B,L,C]: the layout the rest of this library keeps activations in;
//     a frame is C contiguous floats, so both sides move 128-bit vectors with no staging at all.
#include "common.cuh"
#include "kernels.h"

namespace dp {
namespace {

constexpr int SW = 32;  // chunks per CTA window

__global__ void __launch_bounds__(256) segment_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int L, int K, int S) {
    extern __shared__ float tile[];
    const int P = K / 2;
    const int row = blockIdx.y, s0 = blockIdx.x * SW;
    const int ns = min(SW, S - s0);
    const int base = (s0 - 1) * P, len = (ns - 1) * P + K;
    const float* xr = x + (size_t)row * L;
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        int src = base + i;
        tile[i] = (src >= 0 && src < L) ? __ldg(xr + src) : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* yr = y + (size_t)row * K * S + s0;
    if (lane < ns)
        for (int k = warp; k < K; k += 8) yr[(size_t)k * S + lane] = tile[lane * P + k];
}

__global__ void __launch_bounds__(256) overlap_add_nchw_kernel(const float* __restrict__ y, float* __restrict__ x, int L, int K, int S,
                                                               int sw) {
    extern __shared__ float tile[];  // [K][sw + 2]
    const int P = K / 2, W = sw + 2;
    const int row = blockIdx.y, s0 = blockIdx.x * sw;
    const int ns = min(sw + 1, S - s0);  // chunks s0 .. s0+sw
    const float* yr = y + (size_t)row * K * S + s0;
    for (int i = threadIdx.x; i < K * ns; i += blockDim.x) {
        int k = i / ns, s = i - k * ns;
        tile[k * W + s] = __ldg(yr + (size_t)k * S + s);
    }
    __syncthreads();
    const int t0 = s0 * P, t1 = min((s0 + sw) * P, L);
    float* xr = x + (size_t)row * L;
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        int k1 = (t + P) % K, s1 = 2 * ((t + P) / K) - s0;
        int k2 = t % K, s2 = 2 * (t / K) + 1 - s0;
        xr[t] = tile[k1 * W + s1] + tile[k2 * W + s2];
    }
}

__global__ void __launch_bounds__(256) segment_cl_kernel(const float4* __restrict__ f, float4* __restrict__ x, int B, int L, int K, int S,
                                                         int C4) {
    const long long total = (long long)B * S * K * C4;
    const int P = K / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c4 = (int)(i % C4);
        long long r = i / C4;
        int k = (int)(r % K);
        r /= K;
        int s = (int)(r % S);
        int b = (int)(r / S);
        int t = (s - 1) * P + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < L) v = ldg_stream(f + ((size_t)b * L + t) * C4 + c4);
        x[i] = v;
    }
}

__global__ void __launch_bounds__(256) overlap_add_cl_kernel(const float4* __restrict__ x, float4* __restrict__ f, int B, int L, int K, int S,
                                                             int C4) {
    const long long total = (long long)B * L * C4;
    const int P = K / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c4 = (int)(i % C4);
        long long r = i / C4;
        int t = (int)(r % L);
        int b = (int)(r / L);
        int k1 = (t + P) % K, s1 = 2 * ((t + P) / K);
        int k2 = t % K, s2 = 2 * (t / K) + 1;
        float4 a = ldg_stream(x + (((size_t)b * S + s1) * K + k1) * C4 + c4);
        float4 c = ldg_stream(x + (((size_t)b * S + s2) * K + k2) * C4 + c4);
        f[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
    }
}

inline int grid_for(long long total, int per_block = 256) {
    long long blocks = ceil_div_ll(total, per_block);
    long long cap = 148LL * 16;  // multiple of the SM count; grid-stride covers the rest
    return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

cudaError_t launch_segment_nchw(const float* x, float* y, int rows, int L, int K, cudaStream_t st) {
    if (K <= 0 || (K & 1) || L <= 0) return cudaErrorInvalidValue;
    if (rows <= 0) return cudaSuccess;
    const int P = K / 2;
    const int rest = K - (P + L % K) % K;
    const int S = 2 * ((L + rest + P) / K);
    size_t smem = (size_t)((SW - 1) * P + K) * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(segment_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(ceil_div(S, SW), rows);
    segment_nchw_kernel<<<grid, 256, smem, st>>>(x, y, L, K, S);
    return cudaGetLastError();
}

cudaError_t launch_overlap_add_nchw(const float* y, float* x, int rows, int K, int S, int L, cudaStream_t st) {
    if (K <= 0 || (K & 1) || S <= 0 || (S & 1) || L <= 0) return cudaErrorInvalidValue;
    if (rows <= 0) return cudaSuccess;
    int sw = SW;
    while (sw > 2 && (size_t)K * (sw + 2) * sizeof(float) > 160 * 1024) sw /= 2;
    size_t smem = (size_t)K * (sw + 2) * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(overlap_add_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int P = K / 2;
    dim3 grid(ceil_div(ceil_div(L, P), sw), rows);
    overlap_add_nchw_kernel<<<grid, 256, smem, st>>>(y, x, L, K, S, sw);
    return cudaGetLastError();
}

cudaError_t launch_segment_cl(const float* f, float* x, int B, int L, int K, int S, int C, cudaStream_t st) {
    if ((C & 3) || (K & 1)) return cudaErrorInvalidValue;
    long long total = (long long)B * S * K * (C / 4);
    if (total <= 0) return cudaSuccess;
    segment_cl_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const float4*>(f), reinterpret_cast<float4*>(x), B, L, K, S, C / 4);
    return cudaGetLastError();
}

cudaError_t launch_overlap_add_cl(const float* x, float* f, int B, int L, int K, int S, int C, cudaStream_t st) {
    if ((C & 3) || (K & 1)) return cudaErrorInvalidValue;
    long long total = (long long)B * L * (C / 4);
    if (total <= 0) return cudaSuccess;
    overlap_add_cl_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(f), B, L, K, S, C / 4);
    return cudaGetLastError();
}

}  // namespace dp
