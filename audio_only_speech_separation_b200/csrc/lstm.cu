// Persistent bidirectional LSTM recurrence (forward and BPTT) for the dual-path blocks (sm_100a).
//
// Replaces the nn.LSTM time loop of ProjRNN (look2hear/models/utils/gc3_basics.py:16,22; cuDNN RNN in the
// reference) for I = 64, H = 128, one layer, zero initial state, equal-length sequences.
//
// Layout / mapping
//   * one CTA = one direction x a tile of NS = 8*NT sequences; 8 warps; warp w owns hidden units [16w,16w+16)
//   * the whole time loop runs inside the kernel; W_hh never leaves the SM: its bf16 "hi" half lives in
//     REGISTERS as ready-made mma A-fragments (128 regs/thread), its bf16 "lo" half (fp32-parity mode only)
//     in 128 KB of shared memory in fragment order (one conflict-free LDS.128 per fragment)
//   * gates^T[512 x NS] = W_hh[512 x 128] * h^T[128 x NS] on the tensor cores (m16n8k16, fp32 accumulate);
//     the four m-tiles of a warp are the i,f,g,o rows of its 16 units, so every thread ends up with all four
//     gates of its (unit, sequence) cells in its own accumulators: the cell update needs no shuffles or smem
//   * h_t is written back to shared memory as bf16 hi/lo (the next step's B operand) and to HBM as fp32
//   * gate pre-activations G = x W_ih^T + b come from the in-projection GEMM in "packed" column order
//     (dir*512 + unit*4 + gate) so a thread fetches (i,f,g,o) of one cell with a single 128-bit load; they are
//     loaded straight into the accumulators, the next step's lines are prefetched into L2
//   * training: the activated gates overwrite G in place and c_t is saved; the backward kernel walks time in
//     reverse, rebuilds d(gates) per cell, feeds them (bf16 hi/lo, shared memory) to dh_{t-1} = W_hh^T dgates
//     (K = 512) with W_hh^T resident the same way, and leaves d(pre-activations) in G for the weight-gradient GEMMs
//
// SPLIT = true: bf16x3 (hi*hi + hi*lo + lo*hi), precise activations  -> fp32 parity mode
// SPLIT = false: single bf16 product, tanh.approx activations           -> bf16 mode
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace dp {
namespace {

constexpr int HST = kH + 8;   // h^T smem row stride (bf16): 272 B
constexpr int DST = kG + 8;   // dgates smem row stride (bf16): 1040 B
constexpr int ALO_BYTES = 8 * 4 * 8 * 32 * 16;  // 131072

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
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

constexpr int GST = kG + 4;  // staged gate-tile row stride (floats): 2064 B keeps the float4 reads conflict-free

// Position bases of the tile's sequences -> shared memory (invalid sequences alias the tile's first one: their
// loads are harmless and their stores are suppressed).
__device__ __forceinline__ void fill_seq_bases(int* sbase, int NS, int q0, int nv, const SeqMap& m) {
    for (int i = threadIdx.x; i < NS; i += blockDim.x) {
        int q = q0 + i;
        if (i >= nv) q = q0;
        sbase[i] = (int)((q / m.qdiv) * m.s_hi + (q % m.qdiv) * m.s_lo);
    }
}

template <int NT>
struct PipeGroup {          // per-thread view of one sequence group
    float cst[NT][4];
    unsigned hoff[NT][2];   // element offset of (sequence, unit g) in H / Cst at t = 0
    bool valid[NT][2];
};

// One cell update.  FAST = true: bare MUFU activations and predicated stores, one basic block (pipelined kernels);
// FAST = false: the original formulation of the plain kernels (branch around the stores).
template <int NT, bool SPLIT, bool SAVE, bool FAST>
__device__ __forceinline__ void pipe_cell(int q, float (&acc)[4][NT][4], PipeGroup<NT>& gr, __nv_bfloat16* h_hi, __nv_bfloat16* h_lo,
                                          unsigned toff256, bool store, float* __restrict__ G, float* __restrict__ H,
                                          float* __restrict__ Cst, int warp, int g, int c) {
    const int n = q >> 2, idx = q & 3, h = idx >> 1, e = idx & 1;
    const float ig = FAST ? sigmoid_cell<SPLIT>(acc[0][n][idx]) : sigmoid_f<SPLIT>(acc[0][n][idx]);
    const float fg = FAST ? sigmoid_cell<SPLIT>(acc[1][n][idx]) : sigmoid_f<SPLIT>(acc[1][n][idx]);
    const float gg = FAST ? tanh_cell<SPLIT>(acc[2][n][idx]) : tanh_f<SPLIT>(acc[2][n][idx]);
    const float og = FAST ? sigmoid_cell<SPLIT>(acc[3][n][idx]) : sigmoid_f<SPLIT>(acc[3][n][idx]);
    const float cc = fmaf(fg, gr.cst[n][idx], ig * gg);
    gr.cst[n][idx] = cc;
    const float hh = og * (FAST ? tanh_cell<SPLIT>(cc) : tanh_f<SPLIT>(cc));
    const bool st = gr.valid[n][e] && store;
    const unsigned ho = gr.hoff[n][e] + toff256 + h * 8;
    if (FAST) {
        stg_pred(H + ho, hh, st && H != nullptr);
        if (SAVE) {
            stg_pred(Cst + ho, cc, st);
            // packed gate column = 4 x hidden column, so the gate offset is exactly 4 * ho
            stg_pred(reinterpret_cast<float4*>(G + (size_t)ho * 4), ig, fg, gg, og, st);
        }
    } else if (st) {
        if (H != nullptr) H[ho] = hh;
        if (SAVE) {
            Cst[ho] = cc;
            *reinterpret_cast<float4*>(G + (size_t)ho * 4) = make_float4(ig, fg, gg, og);
        }
    }
    const int so = (n * 8 + 2 * c + e) * HST + 16 * warp + g + 8 * h;
    const __nv_bfloat16 hb = __float2bfloat16_rn(hh);
    h_hi[so] = hb;
    if (SPLIT) h_lo[so] = __float2bfloat16_rn(hh - __bfloat162float(hb));
}

template <int NT, bool SPLIT, bool SAVE>
__global__ void __launch_bounds__(256, 1) lstm_fwd_kernel(const LstmPack w, float* __restrict__ G, float* __restrict__ H,
                                                          float* __restrict__ Cst, const SeqMap m, const LstmPlanes pl, const int spc) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NS = 8 * NT;
    uint4* alo = reinterpret_cast<uint4*>(smem);
    __nv_bfloat16* hs_hi = reinterpret_cast<__nv_bfloat16*>(smem + (SPLIT ? ALO_BYTES : 0));
    __nv_bfloat16* hs_lo = hs_hi + NS * HST;
    float* gs = reinterpret_cast<float*>(hs_lo + NS * HST);  // [NS][GST] gate pre-activations of the current step
    int* sbase = reinterpret_cast<int*>(gs + NS * GST);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int dir = blockIdx.y;
    const int q0 = blockIdx.x * spc;           // this CTA's sequences: slots [0, nv) of its NS-slot tile
    const int nv = min(spc, m.nseq - q0);

    uint4 ahi[4][8];
    {
        const uint4* src = w.whh_f_hi + ((size_t)dir * 8 + warp) * (4 * 8 * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) ahi[j][ks] = src[(j * 8 + ks) * 32 + lane];
    }
    if (SPLIT) {
        const uint4* src = w.whh_f_lo + (size_t)dir * 8192;
        for (int i = tid; i < 8192; i += 256) alo[i] = src[i];
    }
    for (int i = tid; i < NS * HST; i += 256) reinterpret_cast<uint32_t*>(hs_hi)[i] = 0u;  // h_{-1} = 0 (hi and lo)
    for (int i = tid; i < NS * GST; i += 256) gs[i] = 0.f;                                   // unused slots are never staged
    fill_seq_bases(sbase, NS, q0, nv, m);
    __syncthreads();

    // stage one step's gate tile: NS rows of 2 KB, 16-byte chunks, fully coalesced, L1-bypassing
    auto stage_gates = [&](int t) {
        const unsigned toff = (unsigned)t * (unsigned)m.s_t;
#pragma unroll
        for (int i = 0; i < NS / 2; ++i) {
            int ch = tid + 256 * i, sq = ch >> 7, col = ch & 127;
            const float* src = G + ((size_t)(sbase[sq] + toff) * 1024 + dir * kG + col * 4);
            if (sq < nv) cp_async16(gs + sq * GST + col * 4, src);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    stage_gates(dir ? m.len - 1 : 0);

    unsigned hoff[NT][2];   // element offset of (sequence, unit g) in H / Cst at t = 0
    bool valid[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int sl = n * 8 + 2 * c + e;
            valid[n][e] = sl < nv;
            hoff[n][e] = (unsigned)sbase[sl] * 256u + (unsigned)(dir * kH + 16 * warp + g);
        }
    float cst[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) cst[n][i] = 0.f;
    const int ucol = (16 * warp + g) * 4;  // packed gate column of (h = 0) inside the direction; h = 1 adds 32

    for (int step = 0; step < m.len; ++step) {
        const int t = dir ? (m.len - 1 - step) : step;
        const unsigned toff = (unsigned)t * (unsigned)m.s_t;
        cp_async_wait_all();
        __syncthreads();  // gate tile of step t landed; h_{t-1} (written by all warps) visible
        if (pl.h_hi != nullptr || pl.hp_hi != nullptr) {
            // h_{t-1} sits in shared memory as bf16 hi/lo rows: copy it out as the operand planes of the GEMMs that consume it
            // (H at the previous position; "h_prev" at this position, zeros at the first step) with coalesced 128-bit stores
            const unsigned tprev = (unsigned)(dir ? t + 1 : t - 1) * (unsigned)m.s_t;
            for (int ch = tid; ch < NS * 16; ch += 256) {
                const int sq = ch >> 4, c16 = ch & 15;
                if (sq >= nv) continue;
                const uint4 vh = *reinterpret_cast<const uint4*>(hs_hi + sq * HST + c16 * 8);
                uint4 vl = make_uint4(0, 0, 0, 0);
                if (SPLIT) vl = *reinterpret_cast<const uint4*>(hs_lo + sq * HST + c16 * 8);
                const size_t col = (size_t)dir * kH + c16 * 8;
                if (pl.hp_hi != nullptr) {
                    const size_t o = (size_t)((unsigned)sbase[sq] + toff) * 256 + col;
                    *reinterpret_cast<uint4*>(pl.hp_hi + o) = vh;
                    if (SPLIT && pl.hp_lo != nullptr) *reinterpret_cast<uint4*>(pl.hp_lo + o) = vl;
                }
                if (pl.h_hi != nullptr && step > 0) {
                    const size_t o = (size_t)((unsigned)sbase[sq] + tprev) * 256 + col;
                    *reinterpret_cast<uint4*>(pl.h_hi + o) = vh;
                    if (SPLIT && pl.h_lo != nullptr) *reinterpret_cast<uint4*>(pl.h_lo + o) = vl;
                }
            }
        }
        float acc[4][NT][4];
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float4 v = *reinterpret_cast<const float4*>(gs + (n * 8 + 2 * c + e) * GST + ucol + h * 32);
                    acc[0][n][h * 2 + e] = v.x; acc[1][n][h * 2 + e] = v.y; acc[2][n][h * 2 + e] = v.z; acc[3][n][h * 2 + e] = v.w;
                }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            uint32_t bh[NT][2], bl[NT][2];
            load_b_frags<NT>(hs_hi, HST, ks * 16, lane, bh);
            uint4 al[4];
            if (SPLIT) {
                load_b_frags<NT>(hs_lo, HST, ks * 16, lane, bl);
#pragma unroll
                for (int j = 0; j < 4; ++j) al[j] = alo[((warp * 4 + j) * 8 + ks) * 32 + lane];
            }
            // the three split products go to the same accumulator: issue them in separate sweeps over the 4*NT
            // accumulators so that consecutive HMMAs are independent
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int n = 0; n < NT; ++n) mma_bf16(acc[j][n], ahi[j][ks], bh[n]);
            if (SPLIT) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int n = 0; n < NT; ++n) mma_bf16(acc[j][n], ahi[j][ks], bl[n]);
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int n = 0; n < NT; ++n) mma_bf16(acc[j][n], al[j], bh[n]);
            }
        }
        __syncthreads();  // every warp is done reading h_{t-1} and the gate tile
        if (step + 1 < m.len) stage_gates(dir ? t - 1 : t + 1);  // lands while the cell update below runs
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int idx = h * 2 + e;
                    float ig = sigmoid_f<SPLIT>(acc[0][n][idx]);
                    float fg = sigmoid_f<SPLIT>(acc[1][n][idx]);
                    float gg = tanh_f<SPLIT>(acc[2][n][idx]);
                    float og = sigmoid_f<SPLIT>(acc[3][n][idx]);
                    float cc = fmaf(fg, cst[n][idx], ig * gg);
                    cst[n][idx] = cc;
                    float hh = og * tanh_f<SPLIT>(cc);
                    if (valid[n][e]) {
                        const unsigned ho = hoff[n][e] + toff * 256u + h * 8;
                        if (H != nullptr) H[ho] = hh;
                        if (SAVE) {
                            Cst[ho] = cc;
                            // packed gate column = 4 x hidden column, so the gate offset is exactly 4 * ho
                            *reinterpret_cast<float4*>(G + (size_t)ho * 4) = make_float4(ig, fg, gg, og);
                        }
                    }
                    const int so = (n * 8 + 2 * c + e) * HST + 16 * warp + g + 8 * h;
                    __nv_bfloat16 hb = __float2bfloat16_rn(hh);
                    hs_hi[so] = hb;
                    if (SPLIT) hs_lo[so] = __float2bfloat16_rn(hh - __bfloat162float(hb));
                }
    }
    if (pl.h_hi != nullptr) {  // the last step's h
        __syncthreads();
        const unsigned tlast = (unsigned)(dir ? 0 : m.len - 1) * (unsigned)m.s_t;
        for (int ch = tid; ch < NS * 16; ch += 256) {
            const int sq = ch >> 4, c16 = ch & 15;
            if (sq >= nv) continue;
            const size_t o = (size_t)((unsigned)sbase[sq] + tlast) * 256 + (size_t)dir * kH + c16 * 8;
            *reinterpret_cast<uint4*>(pl.h_hi + o) = *reinterpret_cast<const uint4*>(hs_hi + sq * HST + c16 * 8);
            if (SPLIT && pl.h_lo != nullptr) *reinterpret_cast<uint4*>(pl.h_lo + o) = *reinterpret_cast<const uint4*>(hs_lo + sq * HST + c16 * 8);
        }
    }
}

// ---- software-pipelined forward recurrence ---------------------------------------------------------------------
// The plain kernel above alternates two phases per step in which every warp does the same thing: tensor-core products
// (tensor pipe busy, MUFU idle) and the cell update (MUFU / stores busy, tensor pipe idle) -- ncu: tensor pipe 45 % active.
// Here a CTA's sequences form two independent groups A (NTA n-tiles) and B (NTB n-tiles) that run half a step apart:
//   block 1 of step t:  W_hh h_A(t-1) on the tensor cores  ||  cell update of group B for step t-1
//   block 2 of step t:  W_hh h_B(t-1) on the tensor cores  ||  cell update of group A for step t
// Both halves of a block sit in ONE basic block of straight-line code (cells spread through the k loop), so the HMMA
// stream of one group and the MUFU / store stream of the other issue from the same warp back to back.  Two CTA barriers
// per step as before; gate tiles are staged per group one block ahead.
template <int NT>
__device__ __forceinline__ void pipe_acc_init(float (&acc)[4][NT][4], const float* gs, int c, int ucol) {
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float4 v = *reinterpret_cast<const float4*>(gs + (n * 8 + 2 * c + e) * GST + ucol + h * 32);
                acc[0][n][h * 2 + e] = v.x; acc[1][n][h * 2 + e] = v.y; acc[2][n][h * 2 + e] = v.z; acc[3][n][h * 2 + e] = v.w;
            }
}

// products of group X (accumulators pre-loaded with the gate pre-activations) interleaved with the cell update of group Y
template <int NTX, int NTY, bool SPLIT, bool SAVE>
__device__ __forceinline__ void pipe_block(float (&accX)[4][NTX][4], const uint4 (&ahi)[4][8], const uint4* alo, const __nv_bfloat16* hX_hi,
                                           const __nv_bfloat16* hX_lo, float (&accY)[4][NTY][4], PipeGroup<NTY>& gy, __nv_bfloat16* hY_hi,
                                           __nv_bfloat16* hY_lo, unsigned toffY256, bool storeY, float* __restrict__ G,
                                           float* __restrict__ H, float* __restrict__ Cst, int warp, int lane) {
    const int g = lane >> 2, c = lane & 3;
    constexpr int CY = 4 * NTY;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        uint32_t bh[NTX][2], bl[NTX][2];
        load_b_frags<NTX>(hX_hi, HST, ks * 16, lane, bh);
        uint4 al[4];
        if (SPLIT) {
            load_b_frags<NTX>(hX_lo, HST, ks * 16, lane, bl);
#pragma unroll
            for (int j = 0; j < 4; ++j) al[j] = alo[((warp * 4 + j) * 8 + ks) * 32 + lane];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int n = 0; n < NTX; ++n) mma_bf16_sched(accX[j][n], ahi[j][ks], bh[n]);
        if (SPLIT) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int n = 0; n < NTX; ++n) mma_bf16_sched(accX[j][n], ahi[j][ks], bl[n]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int n = 0; n < NTX; ++n) mma_bf16_sched(accX[j][n], al[j], bh[n]);
        }
#pragma unroll
        for (int q = 0; q < CY; ++q)
            if ((q * 8) / CY == ks) pipe_cell<NTY, SPLIT, SAVE, true>(q, accY, gy, hY_hi, hY_lo, toffY256, storeY, G, H, Cst, warp, g, c);
    }
}

template <int NTA, int NTB, bool SPLIT, bool SAVE>
__global__ void __launch_bounds__(256, 1) lstm_fwd_pipe_kernel(const LstmPack w, float* __restrict__ G, float* __restrict__ H,
                                                               float* __restrict__ Cst, const SeqMap m, const LstmPlanes pl, const int spc) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NS = 8 * (NTA + NTB), RA = 8 * NTA;
    uint4* alo = reinterpret_cast<uint4*>(smem);
    __nv_bfloat16* hs_hi = reinterpret_cast<__nv_bfloat16*>(smem + (SPLIT ? ALO_BYTES : 0));
    __nv_bfloat16* hs_lo = hs_hi + NS * HST;
    float* gs = reinterpret_cast<float*>(hs_lo + NS * HST);  // [NS][GST]: rows [0, RA) group A, [RA, NS) group B
    int* sbase = reinterpret_cast<int*>(gs + NS * GST);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int dir = blockIdx.y;
    const int q0 = blockIdx.x * spc;           // this CTA's sequences: slots [0, nv) of its NS-slot tile
    const int nv = min(spc, m.nseq - q0);

    uint4 ahi[4][8];
    {
        const uint4* src = w.whh_f_hi + ((size_t)dir * 8 + warp) * (4 * 8 * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) ahi[j][ks] = src[(j * 8 + ks) * 32 + lane];
    }
    if (SPLIT) {
        const uint4* src = w.whh_f_lo + (size_t)dir * 8192;
        for (int i = tid; i < 8192; i += 256) alo[i] = src[i];
    }
    for (int i = tid; i < NS * HST; i += 256) reinterpret_cast<uint32_t*>(hs_hi)[i] = 0u;  // h_{-1} = 0 (hi and lo)
    for (int i = tid; i < NS * GST; i += 256) gs[i] = 0.f;                                   // unused slots are never staged
    fill_seq_bases(sbase, NS, q0, nv, m);
    __syncthreads();

    // stage one step's gate rows [r0, r0 + nr): 2 KB per row in 16-byte chunks, L1-bypassing
    auto stage_gates = [&](int r0, int nr, int t) {
        const unsigned toff = (unsigned)t * (unsigned)m.s_t;
        for (int ch = tid; ch < nr * 128; ch += 256) {
            const int sq = r0 + (ch >> 7), col = ch & 127;
            const float* src = G + ((size_t)(sbase[sq] + toff) * 1024 + dir * kG + col * 4);
            if (sq < nv) cp_async16(gs + sq * GST + col * 4, src);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    // rows [r0, r0 + nr) of the h tile hold h of the previous step: emit them as operand planes ("h_prev" at time t, H at tprev)
    auto copy_planes = [&](int r0, int nr, int t, int tprev, bool has_prev) {
        if (pl.h_hi == nullptr && pl.hp_hi == nullptr) return;
        const unsigned toff = (unsigned)t * (unsigned)m.s_t, tpo = (unsigned)tprev * (unsigned)m.s_t;
        for (int ch = tid; ch < nr * 16; ch += 256) {
            const int sq = r0 + (ch >> 4), c16 = ch & 15;
            if (sq >= nv) continue;
            const uint4 vh = *reinterpret_cast<const uint4*>(hs_hi + sq * HST + c16 * 8);
            uint4 vl = make_uint4(0, 0, 0, 0);
            if (SPLIT) vl = *reinterpret_cast<const uint4*>(hs_lo + sq * HST + c16 * 8);
            const size_t col = (size_t)dir * kH + c16 * 8;
            if (pl.hp_hi != nullptr) {
                const size_t o = (size_t)((unsigned)sbase[sq] + toff) * 256 + col;
                *reinterpret_cast<uint4*>(pl.hp_hi + o) = vh;
                if (SPLIT && pl.hp_lo != nullptr) *reinterpret_cast<uint4*>(pl.hp_lo + o) = vl;
            }
            if (pl.h_hi != nullptr && has_prev) {
                const size_t o = (size_t)((unsigned)sbase[sq] + tpo) * 256 + col;
                *reinterpret_cast<uint4*>(pl.h_hi + o) = vh;
                if (SPLIT && pl.h_lo != nullptr) *reinterpret_cast<uint4*>(pl.h_lo + o) = vl;
            }
        }
    };

    PipeGroup<NTA> ga;
    PipeGroup<NTB> gb;
    float accA[4][NTA][4], accB[4][NTB][4];
#pragma unroll
    for (int n = 0; n < NTA; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            ga.cst[n][e] = 0.f;
            if (e < 2) {
                const int sl = n * 8 + 2 * c + e;
                ga.valid[n][e] = sl < nv;
                ga.hoff[n][e] = (unsigned)sbase[sl] * 256u + (unsigned)(dir * kH + 16 * warp + g);
            }
        }
#pragma unroll
    for (int n = 0; n < NTB; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            gb.cst[n][e] = 0.f;
            // zero pre-activations give h = c = 0: group B's "step -1" cell update in the first block reproduces the initial state
#pragma unroll
            for (int j = 0; j < 4; ++j) accB[j][n][e] = 0.f;
            if (e < 2) {
                const int sl = RA + n * 8 + 2 * c + e;
                gb.valid[n][e] = sl < nv;
                gb.hoff[n][e] = (unsigned)sbase[sl] * 256u + (unsigned)(dir * kH + 16 * warp + g);
            }
        }
    const int ucol = (16 * warp + g) * 4;  // packed gate column of (h = 0) inside the direction; h = 1 adds 32
    const unsigned st256 = (unsigned)m.s_t * 256u;

    stage_gates(0, RA, dir ? m.len - 1 : 0);
    for (int step = 0; step < m.len; ++step) {
        const int t = dir ? (m.len - 1 - step) : step;
        const int tprev = dir ? t + 1 : t - 1;
        cp_async_wait_all();
        __syncthreads();  // A's gate tile of step t landed; h_A(t-1) complete; everyone is done with B's tile and h_B of the last block
        stage_gates(RA, NS - RA, t);
        copy_planes(0, RA, t, tprev, step > 0);
        pipe_acc_init<NTA>(accA, gs, c, ucol);
        pipe_block<NTA, NTB, SPLIT, SAVE>(accA, ahi, alo, hs_hi, hs_lo, accB, gb, hs_hi + RA * HST, hs_lo + RA * HST,
                                          (unsigned)tprev * st256, step > 0, G, H, Cst, warp, lane);
        cp_async_wait_all();
        __syncthreads();  // B's gate tile of step t landed; h_B(t-1) complete; everyone is done with A's tile and h_A(t-1)
        if (step + 1 < m.len) stage_gates(0, RA, dir ? t - 1 : t + 1);
        copy_planes(RA, NS - RA, t, tprev, step > 0);
        pipe_acc_init<NTB>(accB, gs + RA * GST, c, ucol);
        pipe_block<NTB, NTA, SPLIT, SAVE>(accB, ahi, alo, hs_hi + RA * HST, hs_lo + RA * HST, accA, ga, hs_hi, hs_lo, (unsigned)t * st256,
                                          true, G, H, Cst, warp, lane);
    }
    __syncthreads();  // every warp is done reading h_B of the step before (the last block's products) before anyone overwrites it
    {   // group B's last cell update
        const unsigned tl = (unsigned)(dir ? 0 : m.len - 1) * st256;
#pragma unroll
        for (int q = 0; q < 4 * NTB; ++q)
            pipe_cell<NTB, SPLIT, SAVE, true>(q, accB, gb, hs_hi + RA * HST, hs_lo + RA * HST, tl, true, G, H, Cst, warp, g, c);
    }
    if (pl.h_hi != nullptr) {  // the last step's h
        __syncthreads();
        const unsigned tlast = (unsigned)(dir ? 0 : m.len - 1) * (unsigned)m.s_t;
        for (int ch = tid; ch < NS * 16; ch += 256) {
            const int sq = ch >> 4, c16 = ch & 15;
            if (sq >= nv) continue;
            const size_t o = (size_t)((unsigned)sbase[sq] + tlast) * 256 + (size_t)dir * kH + c16 * 8;
            *reinterpret_cast<uint4*>(pl.h_hi + o) = *reinterpret_cast<const uint4*>(hs_hi + sq * HST + c16 * 8);
            if (SPLIT && pl.h_lo != nullptr) *reinterpret_cast<uint4*>(pl.h_lo + o) = *reinterpret_cast<const uint4*>(hs_lo + sq * HST + c16 * 8);
        }
    }
}

// ---- 16-warp forward recurrence -----------------------------------------------------------------------------------
// Same algorithm and shared-memory layout as lstm_fwd_kernel, but 512 threads: warp w owns the 8 hidden units [8w, 8w+8) as two
// m-tiles (rows = [i | f] and [g | o] of its units), so W_hh hi costs 64 registers per thread instead of 128 and the kernel runs
// at <= 128 registers with FOUR warps per scheduler instead of two.  The 8-warp kernels are latency-bound (ncu: 0.26 IPC per
// scheduler, tensor pipe 45 % busy, no dominant stall reason); twice the warps per scheduler hide the MUFU / LDS / store-operand
// latencies of the cell update behind the other warps' work.  The fragments are re-gathered from the 8-warp pack at start-up.
template <int NT, bool SPLIT, bool SAVE>
__global__ void __launch_bounds__(512, 1) lstm_fwd16_kernel(const LstmPack w, float* __restrict__ G, float* __restrict__ H,
                                                            float* __restrict__ Cst, const SeqMap m, const LstmPlanes pl, const int spc) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NS = 8 * NT;
    uint4* alo = reinterpret_cast<uint4*>(smem);
    __nv_bfloat16* hs_hi = reinterpret_cast<__nv_bfloat16*>(smem + (SPLIT ? ALO_BYTES : 0));
    __nv_bfloat16* hs_lo = hs_hi + NS * HST;
    float* gs = reinterpret_cast<float*>(hs_lo + NS * HST);
    int* sbase = reinterpret_cast<int*>(gs + NS * GST);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int dir = blockIdx.y;
    const int q0 = blockIdx.x * spc;           // this CTA's sequences: slots [0, nv) of its NS-slot tile
    const int nv = min(spc, m.nseq - q0);

    // 8-warp pack: [dir][wp][gate][ks][lane][reg], unit = 16 wp + g + 8 (reg & 1), k half = reg >> 1.  This warp's units are
    // 16 (warp >> 1) + 8 (warp & 1) + g: component (warp & 1) (+ 2 for the upper k half) of the words of gates 2 mt and 2 mt + 1.
    uint4 ahi[2][8];
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(w.whh_f_hi) + ((size_t)dir * 8 + (warp >> 1)) * (4 * 8 * 32 * 4);
        const int comp = warp & 1;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int b0 = (((2 * mt) * 8 + ks) * 32 + lane) * 4, b1 = (((2 * mt + 1) * 8 + ks) * 32 + lane) * 4;
                ahi[mt][ks] = make_uint4(src[b0 + comp], src[b1 + comp], src[b0 + comp + 2], src[b1 + comp + 2]);
            }
    }
    if (SPLIT) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(w.whh_f_lo) + (size_t)dir * 8192 * 4;
        for (int i = tid; i < 8192; i += 512) {  // alo[((w16 * 2 + mt) * 8 + ks) * 32 + lane]
            const int ln = i & 31, ks = (i >> 5) & 7, mt = (i >> 8) & 1, w16 = i >> 9, comp = w16 & 1;
            const int wb = (w16 >> 1) * (4 * 8 * 32 * 4);
            const int b0 = wb + (((2 * mt) * 8 + ks) * 32 + ln) * 4, b1 = wb + (((2 * mt + 1) * 8 + ks) * 32 + ln) * 4;
            alo[i] = make_uint4(src[b0 + comp], src[b1 + comp], src[b0 + comp + 2], src[b1 + comp + 2]);
        }
    }
    for (int i = tid; i < NS * HST; i += 512) reinterpret_cast<uint32_t*>(hs_hi)[i] = 0u;  // h_{-1} = 0 (hi and lo)
    for (int i = tid; i < NS * GST; i += 512) gs[i] = 0.f;                                   // unused slots are never staged
    fill_seq_bases(sbase, NS, q0, nv, m);
    __syncthreads();

    auto stage_gates = [&](int t) {
        const unsigned toff = (unsigned)t * (unsigned)m.s_t;
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) {
            int ch = tid + 512 * i, sq = ch >> 7, col = ch & 127;
            const float* src = G + ((size_t)(sbase[sq] + toff) * 1024 + dir * kG + col * 4);
            if (sq < nv) cp_async16(gs + sq * GST + col * 4, src);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    stage_gates(dir ? m.len - 1 : 0);

    unsigned hoff[NT][2];
    bool valid[NT][2];
    float cst[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int sl = n * 8 + 2 * c + e;
            valid[n][e] = sl < nv;
            hoff[n][e] = (unsigned)sbase[sl] * 256u + (unsigned)(dir * kH + 8 * warp + g);
            cst[n][e] = 0.f;
        }
    const int ucol = (8 * warp + g) * 4;  // packed gate column of this thread's unit inside the direction

    for (int step = 0; step < m.len; ++step) {
        const int t = dir ? (m.len - 1 - step) : step;
        const unsigned toff = (unsigned)t * (unsigned)m.s_t;
        cp_async_wait_all();
        __syncthreads();  // gate tile of step t landed; h_{t-1} (written by all warps) visible
        if (pl.h_hi != nullptr || pl.hp_hi != nullptr) {
            const unsigned tprev = (unsigned)(dir ? t + 1 : t - 1) * (unsigned)m.s_t;
            for (int ch = tid; ch < NS * 16; ch += 512) {
                const int sq = ch >> 4, c16 = ch & 15;
                if (sq >= nv) continue;
                const uint4 vh = *reinterpret_cast<const uint4*>(hs_hi + sq * HST + c16 * 8);
                uint4 vl = make_uint4(0, 0, 0, 0);
                if (SPLIT) vl = *reinterpret_cast<const uint4*>(hs_lo + sq * HST + c16 * 8);
                const size_t col = (size_t)dir * kH + c16 * 8;
                if (pl.hp_hi != nullptr) {
                    const size_t o = (size_t)((unsigned)sbase[sq] + toff) * 256 + col;
                    *reinterpret_cast<uint4*>(pl.hp_hi + o) = vh;
                    if (SPLIT && pl.hp_lo != nullptr) *reinterpret_cast<uint4*>(pl.hp_lo + o) = vl;
                }
                if (pl.h_hi != nullptr && step > 0) {
                    const size_t o = (size_t)((unsigned)sbase[sq] + tprev) * 256 + col;
                    *reinterpret_cast<uint4*>(pl.h_hi + o) = vh;
                    if (SPLIT && pl.h_lo != nullptr) *reinterpret_cast<uint4*>(pl.h_lo + o) = vl;
                }
            }
        }
        float acc[2][NT][4];  // [0]: rows g = i, g + 8 = f;  [1]: g-gate, o
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float4 v = *reinterpret_cast<const float4*>(gs + (n * 8 + 2 * c + e) * GST + ucol);
                acc[0][n][e] = v.x; acc[0][n][2 + e] = v.y; acc[1][n][e] = v.z; acc[1][n][2 + e] = v.w;
            }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            uint32_t bh[NT][2];
            load_b_frags<NT>(hs_hi, HST, ks * 16, lane, bh);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int n = 0; n < NT; ++n) mma_bf16(acc[mt][n], ahi[mt][ks], bh[n]);
            if (SPLIT) {
                {
                    uint4 al[2];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) al[mt] = alo[((warp * 2 + mt) * 8 + ks) * 32 + lane];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int n = 0; n < NT; ++n) mma_bf16(acc[mt][n], al[mt], bh[n]);
                }
                uint32_t bl[NT][2];
                load_b_frags<NT>(hs_lo, HST, ks * 16, lane, bl);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int n = 0; n < NT; ++n) mma_bf16(acc[mt][n], ahi[mt][ks], bl[n]);
            }
        }
        __syncthreads();  // every warp is done reading h_{t-1} and the gate tile
        if (step + 1 < m.len) stage_gates(dir ? t - 1 : t + 1);
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float ig = sigmoid_cell<SPLIT>(acc[0][n][e]);
                const float fg = sigmoid_cell<SPLIT>(acc[0][n][2 + e]);
                const float gg = tanh_cell<SPLIT>(acc[1][n][e]);
                const float og = sigmoid_cell<SPLIT>(acc[1][n][2 + e]);
                const float cc = fmaf(fg, cst[n][e], ig * gg);
                cst[n][e] = cc;
                const float hh = og * tanh_cell<SPLIT>(cc);
                const unsigned ho = hoff[n][e] + toff * 256u;
                stg_pred(H + ho, hh, valid[n][e] && H != nullptr);
                if (SAVE) {
                    stg_pred(Cst + ho, cc, valid[n][e]);
                    stg_pred(reinterpret_cast<float4*>(G + (size_t)ho * 4), ig, fg, gg, og, valid[n][e]);
                }
                const int so = (n * 8 + 2 * c + e) * HST + 8 * warp + g;
                const __nv_bfloat16 hb = __float2bfloat16_rn(hh);
                hs_hi[so] = hb;
                if (SPLIT) hs_lo[so] = __float2bfloat16_rn(hh - __bfloat162float(hb));
            }
    }
    if (pl.h_hi != nullptr) {  // the last step's h
        __syncthreads();
        const unsigned tlast = (unsigned)(dir ? 0 : m.len - 1) * (unsigned)m.s_t;
        for (int ch = tid; ch < NS * 16; ch += 512) {
            const int sq = ch >> 4, c16 = ch & 15;
            if (sq >= nv) continue;
            const size_t o = (size_t)((unsigned)sbase[sq] + tlast) * 256 + (size_t)dir * kH + c16 * 8;
            *reinterpret_cast<uint4*>(pl.h_hi + o) = *reinterpret_cast<const uint4*>(hs_hi + sq * HST + c16 * 8);
            if (SPLIT && pl.h_lo != nullptr) *reinterpret_cast<uint4*>(pl.h_lo + o) = *reinterpret_cast<const uint4*>(hs_lo + sq * HST + c16 * 8);
        }
    }
}

// ---- cluster forward recurrence (small passes, inference) --------------------------------------------------------
// At B = 1 a pass has ~80-100 sequences per direction: 11-13 tiles of 8, i.e. two dozen busy SMs, and each of them is bound by the
// legacy tensor pipe -- the products of W_hh (512 x 128, three bf16 pairs) cost 768 mma.m16n8k16 per step whatever the number of
// sequences (measured 1.7-2.0 us per step).  Here the gate rows of a tile are split over a cluster of four CTAs: CTA r owns the hidden
// units [32 r, 32 r + 32) (four warps x 8 units, W_hh hi in registers as in the 16-warp kernel, lo in 32 KB of shared memory), does a
// quarter of the products, updates its cells and writes its slice of h_t (bf16 hi / lo) into the h tile of ALL four CTAs through
// distributed shared memory; one cluster barrier per step.  The h tile is double buffered: step t writes buffer t & 1 while slower
// peers may still read h_{t-1} from the other one.  Same per-warp product order as lstm_fwd16_kernel: identical results.
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_peer(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_peer_b16(uint32_t addr, __nv_bfloat16 v) {
    asm volatile("st.shared::cluster.b16 [%0], %1;\n" ::"r"(addr), "h"(__bfloat16_as_ushort(v)) : "memory");
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

constexpr int GSL = 128 + 4;   // staged gate rows of one CTA: 32 units x 4 gates (+ pad)

template <int NT, bool SPLIT>
__global__ void __launch_bounds__(128, 1) lstm_fwdc_kernel(const LstmPack w, const float* __restrict__ G, float* __restrict__ H, const SeqMap m,
                                                           const LstmPlanes pl, const int spc) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NS = 8 * NT;
    constexpr int ALO = SPLIT ? 4 * 2 * 8 * 32 * 16 : 0;
    constexpr int HPL = NS * HST;                                  // one plane of one h buffer (bf16 elements)
    uint4* alo = reinterpret_cast<uint4*>(smem);
    __nv_bfloat16* hb0 = reinterpret_cast<__nv_bfloat16*>(smem + ALO);   // [buffer][plane hi | lo][NS][HST]
    float* gs = reinterpret_cast<float*>(hb0 + 4 * HPL);               // [buffer][NS][GSL]
    int* sbase = reinterpret_cast<int*>(gs + 2 * NS * GSL);
    __nv_bfloat16* hst = reinterpret_cast<__nv_bfloat16*>(sbase + NS);   // this CTA's slice of h_t: [plane][NS][32 units], sent as 16-byte pieces
    uint64_t* full = reinterpret_cast<uint64_t*>(hst + 2 * NS * 32);      // [2]: every CTA's slice of h_t has landed in buffer b
    uint64_t* empty = full + 2;                                           // [2]: all four CTAs are done reading their buffer b
    constexpr uint32_t SLICES = (SPLIT ? 2u : 1u) * NS * 256u;            // bytes one step delivers into one buffer

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, c = lane & 3;
    const int dir = blockIdx.y;
    const int rank = (int)cluster_rank();
    const int q0 = (blockIdx.x >> 2) * spc;
    const int nv = min(spc, m.nseq - q0);
    const int gw = 4 * rank + warp;                                   // this warp's place among the 16 unit octets

    uint4 ahi[2][8];
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(w.whh_f_hi) + ((size_t)dir * 8 + (gw >> 1)) * (4 * 8 * 32 * 4);
        const int comp = gw & 1;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int b0 = (((2 * mt) * 8 + ks) * 32 + lane) * 4, b1 = (((2 * mt + 1) * 8 + ks) * 32 + lane) * 4;
                ahi[mt][ks] = make_uint4(src[b0 + comp], src[b1 + comp], src[b0 + comp + 2], src[b1 + comp + 2]);
            }
    }
    if (SPLIT) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(w.whh_f_lo) + (size_t)dir * 8192 * 4;
        for (int i = tid; i < 2048; i += 128) {  // alo[((warp * 2 + mt) * 8 + ks) * 32 + lane]
            const int ln = i & 31, ks = (i >> 5) & 7, mt = (i >> 8) & 1, w16 = 4 * rank + (i >> 9), comp = w16 & 1;
            const int wb = (w16 >> 1) * (4 * 8 * 32 * 4);
            const int b0 = wb + (((2 * mt) * 8 + ks) * 32 + ln) * 4, b1 = wb + (((2 * mt + 1) * 8 + ks) * 32 + ln) * 4;
            alo[i] = make_uint4(src[b0 + comp], src[b1 + comp], src[b0 + comp + 2], src[b1 + comp + 2]);
        }
    }
    for (int i = tid; i < 4 * HPL / 2; i += 128) reinterpret_cast<uint32_t*>(hb0)[i] = 0u;   // h_{-1} = 0
    for (int i = tid; i < 2 * NS * GSL; i += 128) gs[i] = 0.f;                                  // unused slots are never staged
    fill_seq_bases(sbase, NS, q0, nv, m);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(full + i)), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(empty + i)), "r"(4));
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    auto stage_gates = [&](int buf, int t) {
        const unsigned toff = (unsigned)t * (unsigned)m.s_t;
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) {
            const int ch = tid + 128 * i, sq = ch >> 5, col = ch & 31;
            const float* src = G + ((size_t)(sbase[sq] + toff) * 1024 + dir * kG + rank * 128 + col * 4);
            if (sq < nv) cp_async16(gs + (buf * NS + sq) * GSL + col * 4, src);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    stage_gates(0, dir ? m.len - 1 : 0);

    unsigned hoff[NT][2];
    bool valid[NT][2];
    float cst[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int sl = n * 8 + 2 * c + e;
            valid[n][e] = sl < nv;
            hoff[n][e] = (unsigned)sbase[sl] * 256u + (unsigned)(dir * kH + 8 * gw + g);
            cst[n][e] = 0.f;
        }
    const int ucol = (8 * warp + g) * 4;   // this thread's unit inside the CTA's staged gate rows
    uint32_t peer[4];   // the peers' h area; their barriers sit at the same distance from it as here
#pragma unroll
    for (int r = 0; r < 4; ++r) peer[r] = map_peer(smem_u32(hb0), (uint32_t)r);
    const uint32_t full_off = smem_u32(full) - smem_u32(hb0), empty_off = smem_u32(empty) - smem_u32(hb0);

    cp_async_wait_all();
    cluster_barrier();   // every CTA of the cluster has zeroed its h tile and initialised its barriers; the first gate rows have landed

    for (int step = 0; step < m.len; ++step) {
        const int t = dir ? (m.len - 1 - step) : step;
        const unsigned toff = (unsigned)t * (unsigned)m.s_t;
        const int cur = step & 1;
        const __nv_bfloat16* hp_hi = hb0 + (cur ^ 1) * 2 * HPL;   // h_{t-1}
        const __nv_bfloat16* hp_lo = hp_hi + HPL;
        if (step + 1 < m.len) stage_gates(cur ^ 1, dir ? t - 1 : t + 1);
        if (tid == 0)   // this step's slices will land in buffer `cur`
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(full + cur)), "r"(SLICES) : "memory");
        float acc[2][NT][4];  // [0]: rows g = i, g + 8 = f;  [1]: g-gate, o
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float4 v = *reinterpret_cast<const float4*>(gs + (cur * NS + n * 8 + 2 * c + e) * GSL + ucol);
                acc[0][n][e] = v.x; acc[0][n][2 + e] = v.y; acc[1][n][e] = v.z; acc[1][n][2 + e] = v.w;
            }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            uint32_t bh[NT][2];
            load_b_frags<NT>(hp_hi, HST, ks * 16, lane, bh);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int n = 0; n < NT; ++n) mma_bf16(acc[mt][n], ahi[mt][ks], bh[n]);
            if (SPLIT) {
                {
                    uint4 al[2];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) al[mt] = alo[((warp * 2 + mt) * 8 + ks) * 32 + lane];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int n = 0; n < NT; ++n) mma_bf16(acc[mt][n], al[mt], bh[n]);
                }
                uint32_t bl[NT][2];
                load_b_frags<NT>(hp_lo, HST, ks * 16, lane, bl);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int n = 0; n < NT; ++n) mma_bf16(acc[mt][n], ahi[mt][ks], bl[n]);
            }
        }
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float ig = sigmoid_cell<SPLIT>(acc[0][n][e]);
                const float fg = sigmoid_cell<SPLIT>(acc[0][n][2 + e]);
                const float gg = tanh_cell<SPLIT>(acc[1][n][e]);
                const float og = sigmoid_cell<SPLIT>(acc[1][n][2 + e]);
                const float cc = fmaf(fg, cst[n][e], ig * gg);
                cst[n][e] = cc;
                const float hh = og * tanh_cell<SPLIT>(cc);
                stg_pred(H + (hoff[n][e] + toff * 256u), hh, valid[n][e] && H != nullptr);
                const int so = (n * 8 + 2 * c + e) * 32 + 8 * warp + g;
                const __nv_bfloat16 hb = __float2bfloat16_rn(hh);
                hst[so] = hb;
                if (SPLIT) hst[NS * 32 + so] = __float2bfloat16_rn(hh - __bfloat162float(hb));
            }
        cp_async_wait_all();
        __syncthreads();   // slice staged; every warp is done reading h_{t-1} (buffer cur ^ 1); the next gate rows are visible
        if (tid < 4)       // tell every CTA that this one no longer needs its buffer cur ^ 1
            asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(peer[tid] + empty_off + (uint32_t)((cur ^ 1) * 8)) : "memory");
        if (step > 0) {    // all four CTAs are done reading buffer `cur` (they were at step - 1): it may be overwritten
            const uint32_t par = (uint32_t)(((step - 1) >> 1) & 1), bar = smem_u32(empty + cur);
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(done) : "r"(bar), "r"(par) : "memory");
        }
        // the slice (64 bytes per sequence and plane) into buffer `cur` of all four CTAs: 16-byte asynchronous stores that count on the
        // destination's `full` barrier (no fence: a release at cluster scope would also wait for the global stores of H)
        for (int i = tid; i < (SPLIT ? 2 : 1) * NS * 4 * 4; i += 128) {
            const int r = i & 3, q = (i >> 2) & 3, row = (i >> 4) % NS, plane = (i >> 4) / NS;
            const uint4 v = *reinterpret_cast<const uint4*>(hst + (plane * NS + row) * 32 + q * 8);
            const uint32_t dst = peer[r] + (uint32_t)(((cur * 2 + plane) * HPL + row * HST + 32 * rank + q * 8) * 2);
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(dst), "r"(v.x),
                         "r"(v.y), "r"(v.z), "r"(v.w), "r"(peer[r] + full_off + (uint32_t)(cur * 8))
                         : "memory");
        }
        {   // h_t complete in this CTA's buffer `cur`
            const uint32_t par = (uint32_t)((step >> 1) & 1), bar = smem_u32(full + cur);
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(done) : "r"(bar), "r"(par) : "memory");
        }
        if (pl.h_hi != nullptr) {   // this CTA's 32 units of h_t -> planes
            const __nv_bfloat16* hc_hi = hb0 + cur * 2 * HPL;
            for (int ch = tid; ch < NS * 4; ch += 128) {
                const int sq = ch >> 2, c16 = 4 * rank + (ch & 3);
                if (sq >= nv) continue;
                const size_t o = (size_t)((unsigned)sbase[sq] + toff) * 256 + (size_t)dir * kH + c16 * 8;
                *reinterpret_cast<uint4*>(pl.h_hi + o) = *reinterpret_cast<const uint4*>(hc_hi + sq * HST + c16 * 8);
                if (SPLIT && pl.h_lo != nullptr)
                    *reinterpret_cast<uint4*>(pl.h_lo + o) = *reinterpret_cast<const uint4*>(hc_hi + HPL + sq * HST + c16 * 8);
            }
        }
    }
    cluster_barrier();   // no CTA leaves while a peer could still address its shared memory
}

// ------------------------------------------------------------------------------------------------
struct PackArgs {
    const float* w_ih[2];
    const float* w_hh[2];
    const float* b_ih[2];
    const float* b_hh[2];
    const float* proj_w;
    LstmPackOut o;
};

__device__ __forceinline__ void store_split(uint32_t* hi, uint32_t* lo, size_t i, float v0, float v1) {
    uint32_t h, l;
    split_pair(v0, v1, h, l);
    hi[i] = h; lo[i] = l;
}

__global__ void pack_lstm_kernel(const PackArgs a) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 65536) {  // W_ih -> [1024,64] packed rows (dir*512 + unit*4 + gate)
        int row = idx >> 6, k = idx & 63;
        int d = row >> 9, u = (row & 511) >> 2, j = row & 3;
        float v = a.w_ih[d][(j * kH + u) * kN + k];
        float vh = bf16_round(v);
        a.o.wih_hi[idx] = __float2bfloat16_rn(vh);
        a.o.wih_lo[idx] = __float2bfloat16_rn(v - vh);
        return;
    }
    idx -= 65536;
    if (idx < 65536) {  // forward-recurrence A fragments of W_hh: [dir][warp][mtile=gate][ks][lane][reg]
        int d = idx >> 15, r = idx & 32767;
        int reg = r & 3, lane = (r >> 2) & 31, ks = (r >> 7) & 7, mt = (r >> 10) & 3, wp = r >> 12;
        int g = lane >> 2, c = lane & 3;
        int unit = 16 * wp + g + (reg & 1) * 8;
        int kk = ks * 16 + 2 * c + (reg >> 1) * 8;
        const float* src = a.w_hh[d] + (size_t)(mt * kH + unit) * kH + kk;
        store_split(reinterpret_cast<uint32_t*>(a.o.whh_f_hi), reinterpret_cast<uint32_t*>(a.o.whh_f_lo), idx, src[0], src[1]);
        return;
    }
    idx -= 65536;
    if (idx < 65536) {  // backward-recurrence A fragments of W_hh^T: [dir][warp][ks][lane][reg]; k' = unit*4 + gate
        int d = idx >> 15, r = idx & 32767;
        int reg = r & 3, lane = (r >> 2) & 31, ks = (r >> 7) & 31, wp = r >> 12;
        int g = lane >> 2, c = lane & 3;
        int mrow = 16 * wp + g + (reg & 1) * 8;
        int k0 = ks * 16 + 2 * c + (reg >> 1) * 8;
        float v[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int kp = k0 + i, u = kp >> 2, j = kp & 3;
            v[i] = a.w_hh[d][(size_t)(j * kH + u) * kH + mrow];
        }
        store_split(reinterpret_cast<uint32_t*>(a.o.whh_b_hi), reinterpret_cast<uint32_t*>(a.o.whh_b_lo), idx, v[0], v[1]);
        return;
    }
    idx -= 65536;
    if (idx < 1024) {
        int d = idx >> 9, u = (idx & 511) >> 2, j = idx & 3;
        a.o.bias[idx] = a.b_ih[d][j * kH + u] + a.b_hh[d][j * kH + u];
        return;
    }
    idx -= 1024;
    if (idx < 65536) {  // packed W_ih transposed: [64 (k)][1024 (packed row)]
        int k = idx >> 10, row = idx & 1023;
        int d = row >> 9, u = (row & 511) >> 2, j = row & 3;
        float v = a.w_ih[d][(j * kH + u) * kN + k];
        float vh = bf16_round(v);
        a.o.wiht_hi[idx] = __float2bfloat16_rn(vh);
        a.o.wiht_lo[idx] = __float2bfloat16_rn(v - vh);
        return;
    }
    idx -= 65536;
    if (idx < 16384 && a.proj_w != nullptr) {  // proj.weight [64,256] -> [256,64]
        int c = idx >> 6, r = idx & 63;
        float v = a.proj_w[r * 256 + c];
        float vh = bf16_round(v);
        a.o.projt_hi[idx] = __float2bfloat16_rn(vh);
        a.o.projt_lo[idx] = __float2bfloat16_rn(v - vh);
    }
}

struct UnpackArgs {
    const float* d_wih;
    const float* d_whh;
    const float* d_bias;
    float* w_ih[2];
    float* w_hh[2];
    float* b_ih[2];
    float* b_hh[2];
};
__global__ void unpack_lstm_grads_kernel(const UnpackArgs a) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 65536) {
        int row = idx >> 6, k = idx & 63;
        int d = row >> 9, u = (row & 511) >> 2, j = row & 3;
        a.w_ih[d][(j * kH + u) * kN + k] += a.d_wih[idx];
        return;
    }
    idx -= 65536;
    if (idx < 131072) {  // [2][512 packed rows][128]
        int d = idx >> 16, r = idx & 65535;
        int row = r >> 7, k = r & 127;
        int u = row >> 2, j = row & 3;
        a.w_hh[d][(size_t)(j * kH + u) * kH + k] += a.d_whh[idx];
        return;
    }
    idx -= 131072;
    if (idx < 1024) {
        int d = idx >> 9, u = (idx & 511) >> 2, j = idx & 3;
        float v = a.d_bias[idx];
        a.b_ih[d][j * kH + u] += v;
        a.b_hh[d][j * kH + u] += v;
    }
}

__global__ void split_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float v = src[i], vh = bf16_round(v);
        hi[i] = __float2bfloat16_rn(vh);
        lo[i] = __float2bfloat16_rn(v - vh);
    }
}

}  // namespace

// Sequences per CTA for a tile of `slots` sequence slots: when the whole pass fits one wave (2 directions x <= 74 CTAs) the
// sequences are spread evenly over all 74 CTAs per direction instead of filling tiles -- the tensor-core work per CTA depends
// only on the slot count, the cell-update / staging / store work on the sequences actually present.
int lstm_seqs_per_cta(int nseq, int slots) {
    if (ceil_div(nseq, slots) > 74) return slots;
    const int spc = ceil_div(nseq, 74);
    return spc < 1 ? 1 : (spc > slots ? slots : spc);
}

namespace {

// fp32 [R, C] row-major -> bf16 hi / lo planes of the TRANSPOSE [C, R] (weights for the input-gradient GEMMs, K-major)
__global__ void transpose_split_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int R, int C) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? src[(size_t)r * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < C && r < R) {
            const float v = tile[threadIdx.x][i], vh = bf16_round(v);
            hi[(size_t)c * R + r] = __float2bfloat16_rn(vh);
            lo[(size_t)c * R + r] = __float2bfloat16_rn(v - vh);
        }
    }
}

template <typename K>
cudaError_t set_smem(K kernel, int bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

template <int NT>
cudaError_t fwd_launch(const LstmPack& w, float* G, float* H, float* Cst, const SeqMap& m, bool split, bool save, const LstmPlanes& pl,
                       cudaStream_t st) {
    const int spc = lstm_seqs_per_cta(m.nseq, 8 * NT);
    dim3 grid(ceil_div(m.nseq, spc), 2);
    int smem = (split ? ALO_BYTES : 0) + 2 * 8 * NT * HST * 2 + 8 * NT * GST * 4 + 8 * NT * 4;
    cudaError_t e;
#define DP_FWD(SP, SV)                                                               \
    do {                                                                             \
        e = set_smem(lstm_fwd_kernel<NT, SP, SV>, smem);                             \
        if (e != cudaSuccess) return e;                                              \
        lstm_fwd_kernel<NT, SP, SV><<<grid, 256, smem, st>>>(w, G, H, Cst, m, pl, spc);   \
    } while (0)
    if (split) { if (save) DP_FWD(true, true); else DP_FWD(true, false); }
    else       { if (save) DP_FWD(false, true); else DP_FWD(false, false); }
#undef DP_FWD
    return cudaGetLastError();
}

template <int NTA, int NTB>
cudaError_t fwd_pipe_launch(const LstmPack& w, float* G, float* H, float* Cst, const SeqMap& m, bool split, bool save, const LstmPlanes& pl,
                            cudaStream_t st) {
    constexpr int NS = 8 * (NTA + NTB);
    const int spc = lstm_seqs_per_cta(m.nseq, NS);
    dim3 grid(ceil_div(m.nseq, spc), 2);
    int smem = (split ? ALO_BYTES : 0) + 2 * NS * HST * 2 + NS * GST * 4 + NS * 4;
    cudaError_t e;
#define DP_FWDP(SP, SV)                                                                          \
    do {                                                                                         \
        e = set_smem(lstm_fwd_pipe_kernel<NTA, NTB, SP, SV>, smem);                              \
        if (e != cudaSuccess) return e;                                                          \
        lstm_fwd_pipe_kernel<NTA, NTB, SP, SV><<<grid, 256, smem, st>>>(w, G, H, Cst, m, pl, spc);    \
    } while (0)
    if (split) { if (save) DP_FWDP(true, true); else DP_FWDP(true, false); }
    else       { if (save) DP_FWDP(false, true); else DP_FWDP(false, false); }
#undef DP_FWDP
    return cudaGetLastError();
}

template <int NT>
cudaError_t fwd16_launch(const LstmPack& w, float* G, float* H, float* Cst, const SeqMap& m, bool split, bool save, const LstmPlanes& pl,
                         cudaStream_t st) {
    const int spc = lstm_seqs_per_cta(m.nseq, 8 * NT);
    dim3 grid(ceil_div(m.nseq, spc), 2);
    int smem = (split ? ALO_BYTES : 0) + 2 * 8 * NT * HST * 2 + 8 * NT * GST * 4 + 8 * NT * 4;
    cudaError_t e;
#define DP_FWD16(SP, SV)                                                               \
    do {                                                                               \
        e = set_smem(lstm_fwd16_kernel<NT, SP, SV>, smem);                             \
        if (e != cudaSuccess) return e;                                                \
        lstm_fwd16_kernel<NT, SP, SV><<<grid, 512, smem, st>>>(w, G, H, Cst, m, pl, spc);   \
    } while (0)
    if (split) { if (save) DP_FWD16(true, true); else DP_FWD16(true, false); }
    else       { if (save) DP_FWD16(false, true); else DP_FWD16(false, false); }
#undef DP_FWD16
    return cudaGetLastError();
}

template <int NT>
cudaError_t fwdc_launch(const LstmPack& w, const float* G, float* H, const SeqMap& m, bool split, const LstmPlanes& pl, int spc, cudaStream_t st) {
    const int smem = (split ? 4 * 2 * 8 * 32 * 16 : 0) + 4 * 8 * NT * HST * 2 + 2 * 8 * NT * GSL * 4 + 8 * NT * 4 + 2 * 8 * NT * 32 * 2 + 64;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * ceil_div(m.nseq, spc), 2);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (split) {
        e = set_smem(lstm_fwdc_kernel<NT, true>, smem);
        if (e != cudaSuccess) return e;
        return cudaLaunchKernelEx(&cfg, lstm_fwdc_kernel<NT, true>, w, G, H, m, pl, spc);
    }
    e = set_smem(lstm_fwdc_kernel<NT, false>, smem);
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, lstm_fwdc_kernel<NT, false>, w, G, H, m, pl, spc);
}

int g_lstm_cluster = 1;   // 0 off, 1 automatic (inference passes of <= 128 sequences), 2 always (inference)
int g_lstm_pipeline = 1;  // 0 plain kernels, 1 automatic, 2 pipelined (groups of 16 + 8 sequences), 3 16-warp kernel

}  // namespace

int lstm_set_pipeline(int mode) {
    if (mode < 0 || mode > 3) return -1;
    g_lstm_pipeline = mode;
    return 0;
}
int lstm_get_pipeline() { return g_lstm_pipeline; }
int lstm_set_cluster(int mode) {
    if (mode < 0 || mode > 2) return -1;
    g_lstm_cluster = mode;
    return 0;
}
int lstm_get_cluster() { return g_lstm_cluster; }

int lstm_pick_nt(int nseq) {
    // smallest tile that still fits the whole pass in one wave of 148 SMs (2 directions per tile)
    for (int nt = 1; nt <= 3; ++nt)
        if (2 * ceil_div(nseq, 8 * nt) <= 148) return nt;
    return 3;
}

cudaError_t launch_lstm_fwd(const LstmPack& w, float* G, float* H, float* Cst, const SeqMap& m, bool split, bool save,
                            cudaStream_t st, const LstmPlanes* planes) {
    if (m.nseq <= 0 || m.len <= 0) return cudaSuccess;
    LstmPlanes pl;
    memset(&pl, 0, sizeof(pl));
    if (planes) pl = *planes;
    if (w.rec5 != nullptr && lstm_rec5_wanted(m, split, false)) return launch_lstm_rec5_fwd(w.rec5, G, H, Cst, m, split, save, st, pl);
    // inference passes too small to fill the GPU: gate rows split over four-CTA clusters (33 clusters are co-resident on a B200)
    // (16 sequence tiles per direction = 32 clusters = one wave).  Measured per pass (tests/tools/time_cluster_rec.py, fp32 / bf16 mode, us):
    // 82 sequences 123 / 82 against 182 / 91 (16-warp kernel), 100: 104 / 71 against 156 / 78; 164: 186 / 122 against 190 / 117 -> automatic
    // up to 128 sequences (8-sequence tiles)
    if (!save && pl.hp_hi == nullptr && g_lstm_cluster != 0 && (g_lstm_cluster == 2 || m.nseq <= 128)) {
        int spc = ceil_div(m.nseq, 16);
        if (spc > 16) spc = 16;
        if (spc <= 8) return fwdc_launch<1>(w, G, H, m, split, pl, spc, st);
        return fwdc_launch<2>(w, G, H, m, split, pl, spc, st);
    }
    const int nt = lstm_pick_nt(m.nseq);
    int pipe = g_lstm_pipeline;
    // automatic (measured on B200 with tests/tools/time_recurrence.py, training mode, intra / inter pass):
    //   B = 16 fp32: plain 552 / 459 us, pipelined 502 / 420, 16-warp 543 / 464      -> pipelined
    //   B = 16 bf16: pipelined 329 / 309, 16-warp 320 / 296                          -> 16-warp
    //   B = 1  fp32: plain 213 / 172, 16-warp 198 / 165; bf16: 113 / 95 vs 107 / 88  -> 16-warp
    if (pipe == 1) pipe = (nt == 3 && split) ? 2 : 3;
    if (pipe == 2) return fwd_pipe_launch<2, 1>(w, G, H, Cst, m, split, save, pl, st);
    if (pipe == 3) {
        switch (nt) {
            case 1: return fwd16_launch<1>(w, G, H, Cst, m, split, save, pl, st);
            case 2: return fwd16_launch<2>(w, G, H, Cst, m, split, save, pl, st);
            default: return fwd16_launch<3>(w, G, H, Cst, m, split, save, pl, st);
        }
    }
    switch (nt) {
        case 1: return fwd_launch<1>(w, G, H, Cst, m, split, save, pl, st);
        case 2: return fwd_launch<2>(w, G, H, Cst, m, split, save, pl, st);
        default: return fwd_launch<3>(w, G, H, Cst, m, split, save, pl, st);
    }
}

cudaError_t launch_pack_lstm(const float* const w_ih[2], const float* const w_hh[2], const float* const b_ih[2],
                             const float* const b_hh[2], const float* proj_w, const LstmPackOut& o, cudaStream_t st) {
    PackArgs a;
    a.proj_w = proj_w;
    for (int d = 0; d < 2; ++d) { a.w_ih[d] = w_ih[d]; a.w_hh[d] = w_hh[d]; a.b_ih[d] = b_ih[d]; a.b_hh[d] = b_hh[d]; }
    a.o = o;
    int total = 65536 * 3 + 1024 + 65536 + 16384;
    pack_lstm_kernel<<<ceil_div(total, 256), 256, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (o.rec5 != nullptr) return launch_pack_lstm_rec5(w_hh, o.rec5, st);
    return cudaSuccess;
}

cudaError_t launch_unpack_lstm_grads(const float* d_wih_pack, const float* d_whh_pack, const float* d_bias_pack,
                                     float* const d_w_ih[2], float* const d_w_hh[2], float* const d_b_ih[2],
                                     float* const d_b_hh[2], cudaStream_t st) {
    UnpackArgs a;
    a.d_wih = d_wih_pack; a.d_whh = d_whh_pack; a.d_bias = d_bias_pack;
    for (int d = 0; d < 2; ++d) { a.w_ih[d] = d_w_ih[d]; a.w_hh[d] = d_w_hh[d]; a.b_ih[d] = d_b_ih[d]; a.b_hh[d] = d_b_hh[d]; }
    int total = 65536 + 131072 + 1024;
    unpack_lstm_grads_kernel<<<ceil_div(total, 256), 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_transpose_split(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, int R, int C, cudaStream_t st) {
    if (R <= 0 || C <= 0) return cudaSuccess;
    transpose_split_kernel<<<dim3(ceil_div(C, 32), ceil_div(R, 32)), dim3(32, 8), 0, st>>>(src, hi, lo, R, C);
    return cudaGetLastError();
}

cudaError_t launch_split_bf16(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    split_bf16_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, st>>>(src, hi, lo, n);
    return cudaGetLastError();
}

}  // namespace dp
