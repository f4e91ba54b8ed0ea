// Persistent BPTT kernel of the bidirectional LSTM (see lstm.cu for the forward kernel and the data layout).
//
// Per step (reverse of the forward visiting order), for a tile of NS = 8*NT sequences of one direction:
//   dh_t      = dH_t + W_hh^T dgates_{t+1}                 tensor cores: A = W_hh^T (hi in registers, lo in smem),
//                                                          B = dgates_{t+1} as bf16 hi/lo in shared memory
//   dc_t      = dh_t * o * (1 - tanh^2 c_t) + dc_{t+1} * f_{t+1}
//   dgates_t  = (dc*g*i(1-i), dc*c_{t-1}*f(1-f), dc*i*(1-g^2), dh*tanh(c_t)*o(1-o))   -> G in place, bf16 hi/lo -> smem
// Latency hiding: the step's activated gates (12 x 128-bit per thread) are requested BEFORE the MMA phase and
// consumed after it; the c_{t-1} and dH tiles of the next step are staged into shared memory with cp.async while
// the current step runs; the c tiles ping-pong (this step's c_{t-1} tile is the next step's c_t tile).  The bias gradient
// (column sums of dgates) is accumulated in registers and reduced once at the end: no separate pass over dG.
#include "common.cuh"
#include "kernels.h"

namespace dp {
int lstm_seqs_per_cta(int nseq, int slots);
namespace {

constexpr int DST = kG + 8;    // dgates smem row stride (bf16): 1040 B
constexpr int SST = kH + 4;    // staged c / dH tile row stride (floats): conflict-free 32-bit reads
constexpr int ALO_BYTES = 8 * 32 * 32 * 16;  // 131072

template <int NT>
__device__ __forceinline__ void load_b_frags(const __nv_bfloat16* base, int stride, int kcol, int lane, uint32_t (&b)[NT][2]) {
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
        uint32_t r[4];
        int off = (np * 16 + (lane & 7) + (lane >> 4) * 8) * stride + kcol + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(r, smem_u32(base + off));
        b[2 * np][0] = r[0]; b[2 * np][1] = r[1]; b[2 * np + 1][0] = r[2]; b[2 * np + 1][1] = r[3];
    }
    if (NT & 1) {
        uint32_t r[2];
        int off = ((NT - 1) * 8 + (lane & 7)) * stride + kcol + ((lane >> 3) & 1) * 8;
        ldmatrix_x2(r, smem_u32(base + off));
        b[NT - 1][0] = r[0]; b[NT - 1][1] = r[1];
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ float4 ld_f4_ordered(const float* p) {  // volatile asm: stays where it is written
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int NT, bool SPLIT>
__global__ void __launch_bounds__(256, 1) lstm_bwd_kernel(const LstmPack w, float* __restrict__ G, const float* __restrict__ Cst,
                                                          const float* __restrict__ dH, float* __restrict__ dbias, const SeqMap m,
                                                          __nv_bfloat16* __restrict__ dG_hi, __nv_bfloat16* __restrict__ dG_lo, const int spc) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NS = 8 * NT;
    uint4* alo = reinterpret_cast<uint4*>(smem);
    __nv_bfloat16* dg_hi = reinterpret_cast<__nv_bfloat16*>(smem + (SPLIT ? ALO_BYTES : 0));
    __nv_bfloat16* dg_lo = dg_hi + NS * DST;
    float* st_c = reinterpret_cast<float*>(dg_lo + NS * DST);   // [2][NS][SST] ping-pong: c_t of this step / c_{t-1}
    float* st_dh = st_c + 2 * NS * SST;                         // [NS][SST] dH_t
    float* dcs = st_dh + NS * SST;                              // [4*NT][256] per-thread dc carry slots
    int* sbase = reinterpret_cast<int*>(dcs + 4 * NT * 256);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int dir = blockIdx.y;
    const int q0 = blockIdx.x * spc;           // this CTA's sequences: slots [0, nv) of its NS-slot tile (lstm_seqs_per_cta)
    const int nv = min(spc, m.nseq - q0);

    uint4 ahi[32];
    {
        const uint4* src = w.whh_b_hi + ((size_t)dir * 8 + warp) * (32 * 32);
#pragma unroll
        for (int ks = 0; ks < 32; ++ks) ahi[ks] = src[ks * 32 + lane];
    }
    if (SPLIT) {
        const uint4* src = w.whh_b_lo + (size_t)dir * 8192;
        for (int i = tid; i < 8192; i += 256) alo[i] = src[i];
    }
    for (int i = tid; i < NS; i += 256) {  // invalid sequences alias the tile's first one (loads harmless, stores masked)
        int q = q0 + i;
        if (i >= nv) q = q0;
        sbase[i] = (int)((q / m.qdiv) * m.s_hi + (q % m.qdiv) * m.s_lo);
    }
    for (int i = tid; i < 3 * NS * SST; i += 256) st_c[i] = 0.f;  // c ping-pong tiles and the dH tile (contiguous)
    __syncthreads();

    // stage the c_{t-1} (= Cst at tp, into c buffer `cb`) and dH_t tiles of one step: NS rows x 512 B each
    auto stage = [&](int t, int tp, bool first, int cb) {
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) {
            const int ch = tid + 256 * i;
            const int which = ch >= NS * 32;
            const int rem = which ? ch - NS * 32 : ch;
            const int sq = rem >> 5, col = rem & 31;
            if (sq >= nv) continue;  // unused slots keep their zero-initialised tiles
            if (which) {
                cp_async16(st_dh + sq * SST + col * 4, dH + ((size_t)(sbase[sq] + (unsigned)t * (unsigned)m.s_t) * 256 + dir * kH + col * 4));
            } else if (!first) {
                cp_async16(st_c + (cb * NS + sq) * SST + col * 4, Cst + ((size_t)(sbase[sq] + (unsigned)tp * (unsigned)m.s_t) * 256 + dir * kH + col * 4));
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };

    bool valid[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) valid[n][e] = (n * 8 + 2 * c + e) < nv;
    const unsigned hcol = (unsigned)(dir * kH + 16 * warp + g);  // H / Cst / dH column of (h = 0); the packed gate
    const int ucol = (16 * warp + g) * 4;                        // column is exactly 4x the H column: G offset = 4 * H offset
    float acc[NT][4];
    float bsum[2][4];
    {
        const int t0 = dir ? 0 : m.len - 1;
        stage(t0, t0, false, 0);                               // c_t of the first visited step -> buffer 0 (dH staged too)
        stage(t0, dir ? t0 + 1 : t0 - 1, m.len == 1, 1);       // its c_{t-1} -> buffer 1 (dH re-staged, harmless)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[n][i] = 0.f; dcs[(n * 4 + i) * 256 + tid] = 0.f; }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) bsum[h][i] = 0.f;
    }

    for (int step = 0; step < m.len; ++step) {
        const int t = dir ? step : (m.len - 1 - step);       // reverse of the forward visiting order
        const int tp = dir ? t + 1 : t - 1;                  // the step visited just before t in the forward pass
        const bool first = (step == m.len - 1);              // t is the forward pass's first step: c_{prev} = 0
        const unsigned toff = (unsigned)t * (unsigned)m.s_t * 256u;
        // 1. request this step's activated gates; they are consumed after the MMA phase below
        float4 gt[NT][4];
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned ho = (unsigned)sbase[n * 8 + 2 * c + (i & 1)] * 256u + hcol + toff + (i >> 1) * 8;
                gt[n][i] = ld_f4_ordered(G + (size_t)ho * 4);
            }
        // 2. dh_rec = W_hh^T dgates of the previous step
        if (step > 0) {
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 32; ++ks) {
                uint32_t bh[NT][2], bl[NT][2];
                load_b_frags<NT>(dg_hi, DST, ks * 16, lane, bh);
                uint4 al;
                if (SPLIT) {
                    load_b_frags<NT>(dg_lo, DST, ks * 16, lane, bl);
                    al = alo[(warp * 32 + ks) * 32 + lane];
                }
#pragma unroll
                for (int n = 0; n < NT; ++n) mma_bf16(acc[n], ahi[ks], bh[n]);
                if (SPLIT) {
#pragma unroll
                    for (int n = 0; n < NT; ++n) mma_bf16(acc[n], ahi[ks], bl[n]);
#pragma unroll
                    for (int n = 0; n < NT; ++n) mma_bf16(acc[n], al, bh[n]);
                }
            }
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        __syncthreads();  // staged tiles landed; every warp is done reading the previous dgates tile
        // 3. cell backward
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int h = i >> 1, e = i & 1;
                const int sl = n * 8 + 2 * c + e, un = 16 * warp + g + 8 * h;
                const float4 a = gt[n][i];
                const float cp = first ? 0.f : st_c[(((step + 1) & 1) * NS + sl) * SST + un];
                const float dh = st_dh[sl * SST + un] + acc[n][i];
                const float tc = tanh_f<SPLIT>(st_c[((step & 1) * NS + sl) * SST + un]);
                const float dc = fmaf(dh * a.w, 1.f - tc * tc, dcs[(n * 4 + i) * 256 + tid]);
                dcs[(n * 4 + i) * 256 + tid] = dc * a.y;
                float4 dg;
                dg.x = dc * a.z * a.x * (1.f - a.x);
                dg.y = dc * cp * a.y * (1.f - a.y);
                dg.z = dc * a.x * (1.f - a.z * a.z);
                dg.w = dh * tc * a.w * (1.f - a.w);
                uint2 hi, lo;
                split_pair(dg.x, dg.y, hi.x, lo.x);
                split_pair(dg.z, dg.w, hi.y, lo.y);
                if (valid[n][e]) {
                    const unsigned ho = (unsigned)sbase[sl] * 256u + hcol + toff + h * 8;
                    if (dG_hi != nullptr) {  // operand planes for the TMA-fed GEMMs that consume dG (same 4 bytes per element)
                        *reinterpret_cast<uint2*>(dG_hi + (size_t)ho * 4) = hi;
                        if (SPLIT && dG_lo != nullptr) *reinterpret_cast<uint2*>(dG_lo + (size_t)ho * 4) = lo;
                    } else {
                        *reinterpret_cast<float4*>(G + (size_t)ho * 4) = dg;
                    }
                    bsum[h][0] += dg.x; bsum[h][1] += dg.y; bsum[h][2] += dg.z; bsum[h][3] += dg.w;
                }
                *reinterpret_cast<uint2*>(dg_hi + sl * DST + un * 4) = hi;
                if (SPLIT) *reinterpret_cast<uint2*>(dg_lo + sl * DST + un * 4) = lo;
            }
        __syncthreads();  // dgates tile complete; staged tiles free
        if (!first) {
            const int tn = tp, tnp = dir ? tn + 1 : tn - 1;
            stage(tn, tnp, step + 1 == m.len - 1, step & 1);  // overwrites this step's (now dead) c_t tile
        }
    }
    // d(b_ih + b_hh) in packed order: reduce over the 4 lanes that share a unit (different sequences), one atomic each
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = bsum[h][i];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (c == 0 && dbias != nullptr) atomicAdd(dbias + dir * kG + ucol + h * 32 + i, v);
        }
}

template <int NT, bool SPLIT>
__global__ void __launch_bounds__(256, 1) lstm_bwd_ks_kernel(const LstmPack w, float* __restrict__ G, const float* __restrict__ Cst,
                                                          const float* __restrict__ dH, float* __restrict__ dbias, const SeqMap m,
                                                          __nv_bfloat16* __restrict__ dG_hi, __nv_bfloat16* __restrict__ dG_lo, const int spc) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NS = 8 * NT;
    uint4* alo = reinterpret_cast<uint4*>(smem);
    __nv_bfloat16* dg_hi = reinterpret_cast<__nv_bfloat16*>(smem + (SPLIT ? ALO_BYTES : 0));
    __nv_bfloat16* dg_lo = dg_hi + NS * DST;
    float* st_c = reinterpret_cast<float*>(dg_lo + NS * DST);   // [2][NS][SST] ping-pong: c_t of this step / c_{t-1}
    float* st_dh = st_c + 2 * NS * SST;                         // [NS][SST] dH_t
    float* dcs = st_dh + NS * SST;                              // [4*NT][256] per-thread dc carry slots
    int* sbase = reinterpret_cast<int*>(dcs + 4 * NT * 256);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int dir = blockIdx.y;
    const int q0 = blockIdx.x * spc;           // this CTA's sequences: slots [0, nv) of its NS-slot tile (lstm_seqs_per_cta)
    const int nv = min(spc, m.nseq - q0);

    // warp (mg, kh): hidden units [32 mg, 32 mg + 32) (m-tiles j = 0, 1 = fragments of the 8-warp pack's warps 2 mg + j), gate columns
    // [256 kh, 256 kh + 256) of the contraction; after the products the two warps of a pair swap one partial tile through shared
    // memory and warp kh finishes m-tile j = kh.  Every warp therefore reads only HALF of the dgates tile (the B operand).
    const int mg = warp >> 1, kh = warp & 1;
    uint4 ahi[2][16];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const uint4* src = w.whh_b_hi + ((size_t)dir * 8 + 2 * mg + j) * (32 * 32) + (size_t)kh * 16 * 32;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) ahi[j][ks] = src[ks * 32 + lane];
    }
    if (SPLIT) {
        const uint4* src = w.whh_b_lo + (size_t)dir * 8192;
        for (int i = tid; i < 8192; i += 256) alo[i] = src[i];
    }
    for (int i = tid; i < NS; i += 256) {  // invalid sequences alias the tile's first one (loads harmless, stores masked)
        int q = q0 + i;
        if (i >= nv) q = q0;
        sbase[i] = (int)((q / m.qdiv) * m.s_hi + (q % m.qdiv) * m.s_lo);
    }
    for (int i = tid; i < 3 * NS * SST; i += 256) st_c[i] = 0.f;  // c ping-pong tiles and the dH tile (contiguous)
    __syncthreads();

    // stage the c_{t-1} (= Cst at tp, into c buffer `cb`) and dH_t tiles of one step: NS rows x 512 B each
    auto stage = [&](int t, int tp, bool first, int cb) {
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) {
            const int ch = tid + 256 * i;
            const int which = ch >= NS * 32;
            const int rem = which ? ch - NS * 32 : ch;
            const int sq = rem >> 5, col = rem & 31;
            if (sq >= nv) continue;  // unused slots keep their zero-initialised tiles
            if (which) {
                cp_async16(st_dh + sq * SST + col * 4, dH + ((size_t)(sbase[sq] + (unsigned)t * (unsigned)m.s_t) * 256 + dir * kH + col * 4));
            } else if (!first) {
                cp_async16(st_c + (cb * NS + sq) * SST + col * 4, Cst + ((size_t)(sbase[sq] + (unsigned)tp * (unsigned)m.s_t) * 256 + dir * kH + col * 4));
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };

    bool valid[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) valid[n][e] = (n * 8 + 2 * c + e) < nv;
    const int wu = 2 * mg + kh;                                  // the 16-unit block this warp finishes
    const unsigned hcol = (unsigned)(dir * kH + 16 * wu + g);    // H / Cst / dH column of (h = 0); the packed gate
    const int ucol = (16 * wu + g) * 4;                          // column is exactly 4x the H column: G offset = 4 * H offset
    float acc[NT][4];
    float bsum[2][4];
    {
        const int t0 = dir ? 0 : m.len - 1;
        stage(t0, t0, false, 0);                               // c_t of the first visited step -> buffer 0 (dH staged too)
        stage(t0, dir ? t0 + 1 : t0 - 1, m.len == 1, 1);       // its c_{t-1} -> buffer 1 (dH re-staged, harmless)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[n][i] = 0.f; dcs[(n * 4 + i) * 256 + tid] = 0.f; }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) bsum[h][i] = 0.f;
    }

    for (int step = 0; step < m.len; ++step) {
        const int t = dir ? step : (m.len - 1 - step);       // reverse of the forward visiting order
        const int tp = dir ? t + 1 : t - 1;                  // the step visited just before t in the forward pass
        const bool first = (step == m.len - 1);              // t is the forward pass's first step: c_{prev} = 0
        const unsigned toff = (unsigned)t * (unsigned)m.s_t * 256u;
        // 1. request this step's activated gates; they are consumed after the MMA phase below
        float4 gt[NT][4];
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned ho = (unsigned)sbase[n * 8 + 2 * c + (i & 1)] * 256u + hcol + toff + (i >> 1) * 8;
                gt[n][i] = ld_f4_ordered(G + (size_t)ho * 4);
            }
        // 2. dh_rec = W_hh^T dgates of the previous step
        float part[2][NT][4];  // partial products over this warp's half of the contraction, m-tiles j = 0, 1
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) part[j][n][i] = 0.f;
        if (step > 0) {
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                const int ksg = kh * 16 + ks;
                uint32_t bh[NT][2], bl[NT][2];
                load_b_frags<NT>(dg_hi, DST, ksg * 16, lane, bh);
                uint4 al[2];
                if (SPLIT) {
                    load_b_frags<NT>(dg_lo, DST, ksg * 16, lane, bl);
#pragma unroll
                    for (int j = 0; j < 2; ++j) al[j] = alo[((2 * mg + j) * 32 + ksg) * 32 + lane];
                }
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int n = 0; n < NT; ++n) mma_bf16(part[j][n], ahi[j][ks], bh[n]);
                if (SPLIT) {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int n = 0; n < NT; ++n) mma_bf16(part[j][n], ahi[j][ks], bl[n]);
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int n = 0; n < NT; ++n) mma_bf16(part[j][n], al[j], bh[n]);
                }
            }
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        __syncthreads();  // staged tiles landed; every warp is done reading the previous dgates tile
        // swap partials inside the warp pair through the (now dead) dgates-lo tile: the word of cell (sequence, unit) is written by the
        // partner's lane that holds the same accumulator element, read by this lane, and later overwritten by this lane's own dgates
        {
            float* xw = reinterpret_cast<float*>(dg_lo);
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int sl = n * 8 + 2 * c + (i & 1), un = 16 * (2 * mg + (1 - kh)) + g + 8 * (i >> 1);
                    xw[sl * (DST / 2) + un * 2] = kh ? part[0][n][i] : part[1][n][i];
                }
        }
        __syncthreads();
        {
            const float* xr = reinterpret_cast<const float*>(dg_lo);
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int sl = n * 8 + 2 * c + (i & 1), un = 16 * wu + g + 8 * (i >> 1);
                    acc[n][i] = (kh ? part[1][n][i] : part[0][n][i]) + xr[sl * (DST / 2) + un * 2];
                }
        }
        // 3. cell backward
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int h = i >> 1, e = i & 1;
                const int sl = n * 8 + 2 * c + e, un = 16 * wu + g + 8 * h;
                const float4 a = gt[n][i];
                const float cp = first ? 0.f : st_c[(((step + 1) & 1) * NS + sl) * SST + un];
                const float dh = st_dh[sl * SST + un] + acc[n][i];
                const float tc = tanh_cell<SPLIT>(st_c[((step & 1) * NS + sl) * SST + un]);
                const float dc = fmaf(dh * a.w, 1.f - tc * tc, dcs[(n * 4 + i) * 256 + tid]);
                dcs[(n * 4 + i) * 256 + tid] = dc * a.y;
                float4 dg;
                dg.x = dc * a.z * a.x * (1.f - a.x);
                dg.y = dc * cp * a.y * (1.f - a.y);
                dg.z = dc * a.x * (1.f - a.z * a.z);
                dg.w = dh * tc * a.w * (1.f - a.w);
                uint2 hi, lo;
                split_pair(dg.x, dg.y, hi.x, lo.x);
                split_pair(dg.z, dg.w, hi.y, lo.y);
                {   // predicated stores: the whole cell phase stays one basic block
                    const bool v = valid[n][e];
                    const unsigned ho = (unsigned)sbase[sl] * 256u + hcol + toff + h * 8;
                    const bool planes = dG_hi != nullptr;
                    stg_pred(reinterpret_cast<uint2*>(dG_hi + (size_t)ho * 4), hi, v && planes);
                    if (SPLIT) stg_pred(reinterpret_cast<uint2*>(dG_lo + (size_t)ho * 4), lo, v && planes && dG_lo != nullptr);
                    stg_pred(reinterpret_cast<float4*>(G + (size_t)ho * 4), dg.x, dg.y, dg.z, dg.w, v && !planes);
                    bsum[h][0] += v ? dg.x : 0.f; bsum[h][1] += v ? dg.y : 0.f;   // selects: unused slots may hold NaN
                    bsum[h][2] += v ? dg.z : 0.f; bsum[h][3] += v ? dg.w : 0.f;
                }
                *reinterpret_cast<uint2*>(dg_hi + sl * DST + un * 4) = hi;
                if (SPLIT) *reinterpret_cast<uint2*>(dg_lo + sl * DST + un * 4) = lo;
            }
        __syncthreads();  // dgates tile complete; staged tiles free
        if (!first) {
            const int tn = tp, tnp = dir ? tn + 1 : tn - 1;
            stage(tn, tnp, step + 1 == m.len - 1, step & 1);  // overwrites this step's (now dead) c_t tile
        }
    }
    // d(b_ih + b_hh) in packed order: reduce over the 4 lanes that share a unit (different sequences), one atomic each
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = bsum[h][i];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (c == 0 && dbias != nullptr) atomicAdd(dbias + dir * kG + ucol + h * 32 + i, v);
        }
}

template <int NT>
cudaError_t bwd_launch(const LstmPack& w, float* G, const float* Cst, const float* dH, float* dbias, const SeqMap& m, bool split,
                       __nv_bfloat16* dG_hi, __nv_bfloat16* dG_lo, cudaStream_t st) {
    const int spc = lstm_seqs_per_cta(m.nseq, 8 * NT);
    dim3 grid(ceil_div(m.nseq, spc), 2);
    int smem = (split ? ALO_BYTES : 0) + 2 * 8 * NT * DST * 2 + 3 * 8 * NT * SST * 4 + 4 * NT * 256 * 4 + 8 * NT * 4;
    cudaError_t e;
#define DP_BWD(KERNEL, SP)                                                                                    \
    do {                                                                                                      \
        e = cudaFuncSetAttribute(KERNEL<NT, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);          \
        if (e != cudaSuccess) return e;                                                                       \
        KERNEL<NT, SP><<<grid, 256, smem, st>>>(w, G, Cst, dH, dbias, m, dG_hi, dG_lo, spc);                  \
    } while (0)
    if (lstm_get_pipeline() == 0) { if (split) DP_BWD(lstm_bwd_kernel, true); else DP_BWD(lstm_bwd_kernel, false); }
    else                          { if (split) DP_BWD(lstm_bwd_ks_kernel, true); else DP_BWD(lstm_bwd_ks_kernel, false); }
#undef DP_BWD
    return cudaGetLastError();
}

}  // namespace

int lstm_pick_nt(int nseq);
int lstm_seqs_per_cta(int nseq, int slots);

cudaError_t launch_lstm_bwd(const LstmPack& w, float* G, const float* Cst, const float* dH, float* dbias, const SeqMap& m, bool split,
                            cudaStream_t st, __nv_bfloat16* dG_hi, __nv_bfloat16* dG_lo) {
    if (m.nseq <= 0 || m.len <= 0) return cudaSuccess;
    if (w.rec5 != nullptr && lstm_rec5_wanted(m, split, true)) return launch_lstm_rec5_bwd(w.rec5, G, Cst, dH, dbias, m, split, st, dG_hi, dG_lo);
    switch (lstm_pick_nt(m.nseq)) {
        case 1: return bwd_launch<1>(w, G, Cst, dH, dbias, m, split, dG_hi, dG_lo, st);
        case 2: return bwd_launch<2>(w, G, Cst, dH, dbias, m, split, dG_hi, dG_lo, st);
        default: return bwd_launch<3>(w, G, Cst, dH, dbias, m, split, dG_hi, dG_lo, st);
    }
}

}  // namespace dp
