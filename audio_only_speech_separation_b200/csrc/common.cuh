// Shared device helpers for the dual-path kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dp {

constexpr int kN = 64;    // feature width of the dual-path stream (enc_dim == bn_dim)
constexpr int kH = 128;   // LSTM hidden size per direction
constexpr int kG = 512;   // 4 * kH gate rows per direction

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---- bf16 split helpers -----------------------------------------------------------------
// x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 significand bits in two bf16 values.
// The three products hi*hi + hi*lo + lo*hi on the tensor cores with fp32 accumulation give
// ~2^-16 relative error per product ("bf16x3"); measured model-level rel-L2 1.4e-5.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
    // element with the LOWER index goes into the LOWER 16 bits (mma fragment convention)
    __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    float ah = bf16_round(a), bh = bf16_round(b);
    hi = pack_bf16x2(ah, bh);
    lo = pack_bf16x2(a - ah, b - bh);
}

// ---- warp-level tensor core MMA (m16n8k16, bf16 inputs, fp32 accumulate) -----------------
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint4& a, const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b[0]), "r"(b[1]));
}

// non-volatile form: ptxas may schedule it among independent instructions (used where MMAs are interleaved with other work)
__device__ __forceinline__ void mma_bf16_sched(float (&d)[4], const uint4& a, const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

// streaming 128-bit accesses (data touched once: keep it out of L1)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

// ---- gate nonlinearities ------------------------------------------------------------------
// PRECISE: ex2.approx + rcp based, abs error ~1e-7 (fp32-parity mode).
// fast:    tanh.approx.f32 (one MUFU op), abs error ~5e-4 (bf16 mode only).
template <bool PRECISE>
__device__ __forceinline__ float sigmoid_f(float x) {
    if (PRECISE) {
        return __fdividef(1.0f, 1.0f + __expf(-x));
    } else {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
        return fmaf(0.5f, t, 0.5f);
    }
}
template <bool PRECISE>
__device__ __forceinline__ float tanh_f(float x) {
    if (PRECISE) {
        // 2*sigmoid(2x) - 1; saturates correctly for |x| large (exp -> 0 or inf)
        return fmaf(2.0f, __fdividef(1.0f, 1.0f + __expf(-2.0f * x)), -1.0f);
    } else {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
        return t;
    }
}

// Branch-free recurrent-cell versions: bare MUFU ops (ex2.approx / rcp.approx, ~2 ulp each, same units __expf / __fdividef use)
// without the range fix-ups of __fdividef -- 1 + 2^y never needs them: y -> +inf gives rcp(inf) = 0, y -> -inf gives rcp(1) = 1.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <bool PRECISE>
__device__ __forceinline__ float sigmoid_cell(float x) {
    if (PRECISE) {
        return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
    } else {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
        return fmaf(0.5f, t, 0.5f);
    }
}
template <bool PRECISE>
__device__ __forceinline__ float tanh_cell(float x) {
    if (PRECISE) {
        return fmaf(2.0f, rcp_approx(1.0f + ex2_approx(-2.8853900817779268f * x)), -1.0f);
    } else {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
        return t;
    }
}
// predicated global stores: keep the recurrent cell update one basic block (an `if` around the stores makes ptxas branch per cell)
__device__ __forceinline__ void stg_pred(float* p, float v, bool pred) {
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p st.global.f32 [%0], %1;\n}\n" ::"l"(p), "f"(v), "r"((unsigned)pred) : "memory");
}
__device__ __forceinline__ void stg_pred(float4* p, float x, float y, float z, float w, bool pred) {
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %5, 0;\n @p st.global.v4.f32 [%0], {%1,%2,%3,%4};\n}\n" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w),
                 "r"((unsigned)pred)
                 : "memory");
}

__device__ __forceinline__ void stg_pred(uint2* p, const uint2& v, bool pred) {
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %3, 0;\n @p st.global.v2.u32 [%0], {%1,%2};\n}\n" ::"l"(p), "r"(v.x), "r"(v.y), "r"((unsigned)pred)
                 : "memory");
}

// ---- counter-based dropout masks ---------------------------------------------------------------------------------------
// keep(key, row, col) is a pure function of the element's indices and a per-(step, layer, site) key, so the backward regenerates the
// forward's mask instead of storing it (tests replay the same function in numpy: tests/dropout_ref.py).  thr24 = p * 2^24.
__host__ __device__ inline uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
__host__ __device__ inline uint32_t drop_site_key(uint32_t seed, int layer, int site) {
    return hash32(seed ^ (0x9E3779B1u * (uint32_t)(layer * 4 + site + 1)));
}
__device__ __forceinline__ bool drop_keep(uint32_t key, uint32_t row, uint32_t col, uint32_t thr24) {
    return (hash32(row * 0x9E3779B1u + col * 0x85EBCA77u + key) >> 8) >= thr24;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dp
