// Fused input projection + bidirectional LSTM recurrence on the 5th-generation tensor cores (sm_100a).
//
// Replaces, for ProjRNN's nn.LSTM(64, 128, bidirectional) (look2hear/models/utils/gc3_basics.py:16,22), both the
// x_t W_ih^T GEMM and the time loop: per step  gates^T[512 x N] = W_ih[512 x 64] x_t^T + W_hh[512 x 128] h_{t-1}^T  is one
// chain of tcgen05.mma instructions (M = 128 gate rows, N = 16 sequences, K = 16), so the [P, 1024] gate pre-activation
// tensor (16x the activation, SURVEY 7 hard part 2) is never written to or read from HBM.
//
// One CTA = one direction x 32 sequences; the whole time loop runs inside the kernel.
//   * weights never leave the SM: the bf16 "hi" halves of W_hh and W_ih (read by two of the three split products) sit in
//     192 KB of shared memory in the UMMA 128-byte-swizzle K-major layout, their "lo" halves (fp32-parity mode) in TENSOR
//     MEMORY as the A operand (256 + 128 columns, written once with tcgen05.st).  The four M tiles are the i, f, g, o rows
//     of the 128 hidden units, so TMEM lane r holds all four gates of unit r and the cell update needs no exchange
//   * measured on B200 (tests/tools/time_lstm.py): a tcgen05.mma with M = 128, K = 16 and N <= 32 occupies the tensor pipe
//     for ~110 cycles whatever the operand source (shared or tensor memory), the accumulator or the issuing warp, i.e. ~7x its
//     arithmetic floor; a step needs (512/128) x (192/16) = 48 of them per split product.  That makes the kernel a win in bf16
//     mode (48 MMAs per step) and a small loss against the register-stationary mma.sync recurrence in fp32-parity mode (144)
//   * x_t tiles arrive by TMA straight from the producer's bf16 hi/lo planes (4-D tensor map: the strided positions of
//     the 16 sequences at time t are one box), double buffered, one step ahead
//   * warp roles: TMA producer | 4 MMA issuers (one per gate tile) | 4 cell-update warps (tcgen05.ld gates -> activations -> c, h; h goes back to
//     shared memory as the next step's B operand) | 2 copy-out warps (h tile -> H / h_prev planes with 128-bit stores).
//   * training additionally stores the activated gates [P,1024] and c_t for the BPTT kernel.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstring>

#include "common.cuh"
#include "kernels.h"
#include "tc5_common.cuh"

namespace dp {
namespace {

constexpr int GN = 16;                 // sequences per group (UMMA N)
constexpr int TILE = GN * 128;         // one [16 rows x 64 bf16] operand tile: 2 KB
constexpr int HH_SM_BYTES = 4 * 2 * 128 * 128;  // 128 KB swizzled shared-memory image of W_hh (hi half)
constexpr int IH_SM_BYTES = 4 * 128 * 128;      // 64 KB  swizzled shared-memory image of W_ih (hi half)
constexpr uint32_t COL_IH = 256, COL_D = 384;

struct FusedArgs {
    const uint32_t* hh_tm;  // lo half, tensor-memory rows: [dir][4][128][64] u32 (pairs of bf16 along k)
    const uint4* hh_sm;     // hi half, [dir][HH_SM_BYTES / 16] swizzled shared-memory image
    const uint32_t* ih_tm;  // [dir][4][128][32] u32
    const uint4* ih_sm;     // [dir][IH_SM_BYTES / 16]
    const float* bias;      // [1024] packed (dir*512 + unit*4 + gate)
    float* G;               // [P,1024] activated gates (SAVE)
    float* Cst;             // [P,256]
    __nv_bfloat16* h_hi;    // planes [P,256]
    __nv_bfloat16* h_lo;
    __nv_bfloat16* hp_hi;
    __nv_bfloat16* hp_lo;
    int inter;              // 0: sequences (b,s) walk k ; 1: sequences (b,k) walk s
    int len;                // time steps
    int nseq;               // intra: number of sequences ; inter: K (sequences per utterance)
    int S, B;               // inter only
    long long s_t;          // position stride of one time step
};

template <bool SPLIT, bool SAVE>
__global__ void __launch_bounds__(352, 1)
lstm_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl, const FusedArgs p) {
    constexpr int PL = SPLIT ? 2 : 1;
    constexpr int NS = 2 * GN;                      // 32 sequences per CTA = UMMA N (two 16-row TMA boxes)
    constexpr int T32 = NS * 128;                   // one [32 rows x 64 bf16] operand tile: 4 KB
    constexpr int OFF_IH = HH_SM_BYTES;
    constexpr int OFF_HS = HH_SM_BYTES + IH_SM_BYTES;
    constexpr int HS_BYTES = PL * 2 * T32;          // planes x k-blocks
    constexpr int OFF_XS = OFF_HS + HS_BYTES;
    constexpr int XS_B = PL * T32;                  // per buffer
    constexpr int OFF_BAR = OFF_XS + 2 * XS_B;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* x_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);  // [b]
    uint64_t* x_empty = x_full + 2;
    uint64_t* d_full = x_empty + 2;
    uint64_t* h_ready = d_full + 1;
    uint64_t* h_copied = h_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_copied + 1);
    int* sbase = reinterpret_cast<int*>(tmem_slot + 2);  // [NS] position of (sequence, t = 0), -1 = not a sequence

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const int len = p.len;
    const int tpb = (p.nseq + GN - 1) / GN;  // inter: 16-row boxes per utterance

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(x_full + i, 1); mbar_init(x_empty + i, 4); }
        mbar_init(d_full, 4); mbar_init(h_ready, 128); mbar_init(h_copied, 2);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        prefetch_tmap(&tmXh);
        if (SPLIT) prefetch_tmap(&tmXl);
    }
    if (tid < NS) {
        const int box = blockIdx.x * 2 + tid / GN, n = tid % GN;
        int v = -1;
        if (!p.inter) {
            const int q = box * GN + n;
            if (q < p.nseq) v = q * len;
        } else {
            const int b = box / tpb, k = (box % tpb) * GN + n;
            if (b < p.B && k < p.nseq) v = b * p.S * p.nseq + k;
        }
        sbase[tid] = v;
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    {   // hi halves of the weights -> shared memory: ready-made 128-byte-swizzled K-major images, plain 128-bit copies
        const uint4* s0 = p.hh_sm + (size_t)dir * (HH_SM_BYTES / 16);
        uint4* d0 = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < HH_SM_BYTES / 16; i += 352) d0[i] = s0[i];
        const uint4* s1 = p.ih_sm + (size_t)dir * (IH_SM_BYTES / 16);
        uint4* d1 = reinterpret_cast<uint4*>(smem + OFF_IH);
        for (int i = tid; i < IH_SM_BYTES / 16; i += 352) d1[i] = s1[i];
        proxy_fence_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (SPLIT && warp >= 5 && warp < 9) {  // lo halves -> tensor memory (lane = gate row of unit r, one column = two k values)
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const uint32_t* src = p.hh_tm + ((size_t)(dir * 4 + j) * 128 + r) * 64;
#pragma unroll 1
            for (int c = 0; c < 64; c += 32) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const uint4 u = *reinterpret_cast<const uint4*>(src + c + i);
                    v[i] = u.x; v[i + 1] = u.y; v[i + 2] = u.z; v[i + 3] = u.w;
                }
                tmem_st32(lane_addr + j * 64 + c, v);
            }
            const uint32_t* si = p.ih_tm + ((size_t)(dir * 4 + j) * 128 + r) * 32;
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const uint4 u = *reinterpret_cast<const uint4*>(si + i);
                v[i] = u.x; v[i + 1] = u.y; v[i + 2] = u.z; v[i + 3] = u.w;
            }
            tmem_st32(lane_addr + COL_IH + j * 32, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ===================== TMA producer: x_t tiles (two 16-sequence boxes per plane), one step ahead =====================
        if (lane == 0) {
            for (int step = 0; step < len; ++step) {
                const int t = dir ? len - 1 - step : step;
                const int b = step & 1;
                mbar_wait(x_empty + b, ((step >> 1) & 1) ^ 1);
                uint8_t* dst = smem + OFF_XS + b * XS_B;
                mbar_expect_tx(x_full + b, XS_B);
                for (int half = 0; half < 2; ++half) {
                    const int box = blockIdx.x * 2 + half;
                    int c1, c2, c3;
                    if (!p.inter) { c1 = t; c2 = box * GN; c3 = 0; }
                    else { c1 = (box % tpb) * GN; c2 = t; c3 = box / tpb; }
                    tma_load_4d(dst + half * TILE, &tmXh, x_full + b, 0, c1, c2, c3);
                    if (SPLIT) tma_load_4d(dst + T32 + half * TILE, &tmXl, x_full + b, 0, c1, c2, c3);
                }
            }
        }
    } else if (warp < 5) {
        // ===================== MMA issuers: one warp per gate tile =====================
        // The four independent gate tiles (i, f, g, o rows = four accumulators) are issued by four warps with warp-uniform
        // control flow (one elected lane issues; see umma_w).  hi weights are the shared-memory A operand (two of the three split
        // products), lo weights the tensor-memory A operand.
        constexpr uint32_t IDESC = idesc_bf16(128, NS, 0, 0);
        const int j = warp - 1;
        const uint64_t d_hh = desc_sw128(smem_u32(smem) + j * 32768, 16, 1024);            // W_hh hi, tile (j, kb) at +(j*2+kb)*16 KB
        const uint64_t d_ih = desc_sw128(smem_u32(smem + OFF_IH) + j * 16384, 16, 1024);   // W_ih hi, tile j at +j*16 KB
        const uint64_t d_hs = desc_sw128(smem_u32(smem + OFF_HS), 16, 1024);
        const uint64_t d_xs0 = desc_sw128(smem_u32(smem + OFF_XS), 16, 1024);
        const uint32_t d = tmem + COL_D + j * NS;
        const uint32_t a_ih = tmem + COL_IH + j * 32, a_hh = tmem + j * 64;
        for (int step = 0; step < len; ++step) {
            const int b = step & 1;
            if (step > 0) mbar_wait(h_ready, (step - 1) & 1);   // h_{t-1} is in shared memory
            mbar_wait(x_full + b, (step >> 1) & 1);
            tc_fence_after();
            const uint64_t d_xs = d_xs0 + (uint64_t)(b * (XS_B >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // x_t W_ih^T, K = 64
                const uint64_t bxh = d_xs + (uint64_t)(k * 2);
                umma_w(d, d_ih + (uint64_t)(k * 2), bxh, IDESC, k != 0);
                if (SPLIT) {
                    umma_w(d, d_ih + (uint64_t)(k * 2), bxh + (uint64_t)(T32 >> 4), IDESC, 1);
                    umma_ts_w(d, a_ih + k * 8, bxh, IDESC, 1);
                }
            }
            if (step > 0) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {  // h_{t-1} W_hh^T, K = 128 (two 64-wide k blocks)
                    const int kb = kk >> 2, k = kk & 3;
                    const uint64_t bhh = d_hs + (uint64_t)(kb * (T32 >> 4) + k * 2);
                    umma_w(d, d_hh + (uint64_t)(kb * 1024 + k * 2), bhh, IDESC, 1);
                    if (SPLIT) {
                        umma_w(d, d_hh + (uint64_t)(kb * 1024 + k * 2), bhh + (uint64_t)((2 * T32) >> 4), IDESC, 1);
                        umma_ts_w(d, a_hh + kk * 8, bhh, IDESC, 1);
                    }
                }
            }
            umma_commit_w(x_empty + b);
            umma_commit_w(d_full);
            __syncwarp();
        }
    } else if (warp < 9) {
        // ===================== cell update: thread = hidden unit, 32 sequences =====================
        const int q = warp & 3, r = q * 32 + lane;
        const float4 bias = *reinterpret_cast<const float4*>(p.bias + dir * kG + r * 4);
        const uint32_t d_addr = tmem + ((uint32_t)(q * 32) << 16) + COL_D;
        uint8_t* hs = smem + OFF_HS;
        const int kb = r >> 6, cch = (r & 63) >> 3, e2 = (r & 7) * 2;
        float cst[NS];
#pragma unroll
        for (int n = 0; n < NS; ++n) cst[n] = 0.f;
        for (int step = 0; step < len; ++step) {
            const int t = dir ? len - 1 - step : step;
            const long long toff = (long long)t * p.s_t;
            mbar_wait(d_full, step & 1);
            tc_fence_after();
            if (step > 0) mbar_wait(h_copied, (step - 1) & 1);  // the copy-out warps are done with the previous h tile
#pragma unroll
            for (int part = 0; part < NS / 8; ++part) {
                float gi[8], gf[8], gg[8], go[8];
                tmem_ld8_nowait(d_addr + 0 * NS + part * 8, gi);
                tmem_ld8_nowait(d_addr + 1 * NS + part * 8, gf);
                tmem_ld8_nowait(d_addr + 2 * NS + part * 8, gg);
                tmem_ld8_nowait(d_addr + 3 * NS + part * 8, go);
                tmem_ld_wait();
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int sl = part * 8 + n;
                    const float ig = sigmoid_f<SPLIT>(gi[n] + bias.x);
                    const float fg = sigmoid_f<SPLIT>(gf[n] + bias.y);
                    const float g2 = tanh_f<SPLIT>(gg[n] + bias.z);
                    const float og = sigmoid_f<SPLIT>(go[n] + bias.w);
                    const float cc = fmaf(fg, cst[sl], ig * g2);
                    cst[sl] = cc;
                    const float hh = og * tanh_f<SPLIT>(cc);
                    if (SAVE) {
                        const int sb = sbase[sl];
                        if (sb >= 0) {
                            const size_t pos = (size_t)(sb + toff);
                            *reinterpret_cast<float4*>(p.G + pos * 1024 + dir * kG + r * 4) = make_float4(ig, fg, g2, og);
                            p.Cst[pos * 256 + dir * kH + r] = cc;
                        }
                    }
                    const __nv_bfloat16 hb = __float2bfloat16_rn(hh);
                    const uint32_t off = kb * T32 + sl * 128 + ((cch ^ (sl & 7)) << 4) + e2;
                    *reinterpret_cast<__nv_bfloat16*>(hs + off) = hb;
                    if (SPLIT) *reinterpret_cast<__nv_bfloat16*>(hs + 2 * T32 + off) = __float2bfloat16_rn(hh - __bfloat162float(hb));
                }
            }
            proxy_fence_async();  // h tile (generic-proxy stores) -> visible to the tensor core's async proxy
            tc_fence_before();
            mbar_arrive(h_ready);
        }
    } else {
        // ===================== copy-out (2 warps, 16 sequences each): h tile -> H (own position) and h_prev (next position) planes =====================
        const int g = warp - 9;
        const uint8_t* hs = smem + OFF_HS;
        if (SAVE && p.hp_hi != nullptr) {  // h_prev of the first visited step is zero
            const long long t0 = (long long)(dir ? len - 1 : 0) * p.s_t;
            for (int ch = lane; ch < GN * 16; ch += 32) {
                const int n = g * GN + (ch >> 4), u = ch & 15, sb = sbase[n];
                if (sb < 0) continue;
                const size_t o = (size_t)(sb + t0) * 256 + dir * kH + u * 8;
                *reinterpret_cast<uint4*>(p.hp_hi + o) = make_uint4(0, 0, 0, 0);
                if (SPLIT && p.hp_lo != nullptr) *reinterpret_cast<uint4*>(p.hp_lo + o) = make_uint4(0, 0, 0, 0);
            }
        }
        for (int step = 0; step < len; ++step) {
            const int t = dir ? len - 1 - step : step;
            const long long toff = (long long)t * p.s_t;
            const long long tnext = (long long)(dir ? t - 1 : t + 1) * p.s_t;
            const bool has_next = step + 1 < len;
            mbar_wait(h_ready, step & 1);
            for (int ch = lane; ch < GN * 16; ch += 32) {
                const int n = g * GN + (ch >> 4), u = ch & 15, sb = sbase[n];
                if (sb < 0) continue;
                const uint32_t off = (u >> 3) * T32 + n * 128 + (((u & 7) ^ (n & 7)) << 4);
                const uint4 vh = *reinterpret_cast<const uint4*>(hs + off);
                uint4 vl = make_uint4(0, 0, 0, 0);
                if (SPLIT) vl = *reinterpret_cast<const uint4*>(hs + 2 * T32 + off);
                const size_t col = (size_t)dir * kH + u * 8;
                if (p.h_hi != nullptr) {
                    const size_t o = (size_t)(sb + toff) * 256 + col;
                    *reinterpret_cast<uint4*>(p.h_hi + o) = vh;
                    if (SPLIT && p.h_lo != nullptr) *reinterpret_cast<uint4*>(p.h_lo + o) = vl;
                }
                if (SAVE && has_next && p.hp_hi != nullptr) {
                    const size_t o = (size_t)(sb + tnext) * 256 + col;
                    *reinterpret_cast<uint4*>(p.hp_hi + o) = vh;
                    if (SPLIT && p.hp_lo != nullptr) *reinterpret_cast<uint4*>(p.hp_lo + o) = vl;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(h_copied);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ weight images
struct Tc5PackArgs {
    const float* w_ih[2];
    const float* w_hh[2];
    uint16_t* hh_tm;  // lo half rows [dir][4][128][128]
    uint8_t* hh_sm;   // hi half image [dir][HH_SM_BYTES]
    uint16_t* ih_tm;  // [dir][4][128][64]
    uint8_t* ih_sm;   // [dir][IH_SM_BYTES]
};
__device__ __forceinline__ uint16_t bf16_bits(float v) {
    __nv_bfloat16 b = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&b);
}
__global__ void pack_tc5_kernel(const Tc5PackArgs a) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 2 * 4 * 128 * 128) {  // W_hh: gate j, unit r, k
        const int k = idx & 127, r = (idx >> 7) & 127, j = (idx >> 14) & 3, d = idx >> 16;
        const float v = a.w_hh[d][(size_t)(j * kH + r) * kH + k];
        const float vh = bf16_round(v);
        a.hh_tm[idx] = bf16_bits(v - vh);
        const int kb = k >> 6, c = (k & 63) >> 3, e = k & 7;
        const size_t off = (size_t)d * HH_SM_BYTES + ((size_t)((j * 2 + kb) * 128 + r)) * 128 + ((c ^ (r & 7)) << 4) + e * 2;
        *reinterpret_cast<uint16_t*>(a.hh_sm + off) = bf16_bits(vh);
        return;
    }
    idx -= 2 * 4 * 128 * 128;
    if (idx < 2 * 4 * 128 * 64) {  // W_ih
        const int k = idx & 63, r = (idx >> 6) & 127, j = (idx >> 13) & 3, d = idx >> 15;
        const float v = a.w_ih[d][(size_t)(j * kH + r) * kN + k];
        const float vh = bf16_round(v);
        a.ih_tm[idx] = bf16_bits(v - vh);
        const int c = k >> 3, e = k & 7;
        const size_t off = (size_t)d * IH_SM_BYTES + ((size_t)(j * 128 + r)) * 128 + ((c ^ (r & 7)) << 4) + e * 2;
        *reinterpret_cast<uint16_t*>(a.ih_sm + off) = bf16_bits(vh);
    }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn2() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(q);
    }
    return fn;
}

// X planes [P,64] viewed as intra: [nseq][len][64] (box = 16 sequences at one time) or inter: [B][S][K][64] (box = 16 k at one s)
bool make_x_map(CUtensorMap* map, const void* base, const LstmFusedGeom& gm) {
    auto fn = encode_fn2();
    if (!fn) return false;
    cuuint64_t gdim[4], gstr[3];
    cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
    if (!gm.inter) {
        gdim[0] = 64; gdim[1] = (cuuint64_t)gm.len; gdim[2] = (cuuint64_t)gm.nseq; gdim[3] = 1;
        gstr[0] = 128; gstr[1] = (cuuint64_t)gm.len * 128; gstr[2] = (cuuint64_t)gm.len * gm.nseq * 128;
        box[0] = 64; box[1] = 1; box[2] = GN; box[3] = 1;
    } else {
        gdim[0] = 64; gdim[1] = (cuuint64_t)gm.K; gdim[2] = (cuuint64_t)gm.S; gdim[3] = (cuuint64_t)gm.B;
        gstr[0] = 128; gstr[1] = (cuuint64_t)gm.K * 128; gstr[2] = (cuuint64_t)gm.K * gm.S * 128;
        box[0] = 64; box[1] = GN; box[2] = 1; box[3] = 1;
    }
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

size_t lstm_tc5_pack_bytes() { return (size_t)2 * 4 * 128 * 128 * 2 + 2 * HH_SM_BYTES + (size_t)2 * 4 * 128 * 64 * 2 + 2 * IH_SM_BYTES; }

cudaError_t launch_pack_lstm_tc5(const float* const w_ih[2], const float* const w_hh[2], void* pack, cudaStream_t st) {
    Tc5PackArgs a;
    uint8_t* b = static_cast<uint8_t*>(pack);
    for (int d = 0; d < 2; ++d) { a.w_ih[d] = w_ih[d]; a.w_hh[d] = w_hh[d]; }
    a.hh_tm = reinterpret_cast<uint16_t*>(b);
    a.hh_sm = b + (size_t)2 * 4 * 128 * 128 * 2;
    a.ih_tm = reinterpret_cast<uint16_t*>(a.hh_sm + 2 * HH_SM_BYTES);
    a.ih_sm = reinterpret_cast<uint8_t*>(a.ih_tm) + (size_t)2 * 4 * 128 * 64 * 2;
    const int total = 2 * 4 * 128 * 128 + 2 * 4 * 128 * 64;
    pack_tc5_kernel<<<ceil_div(total, 256), 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_lstm_fused_fwd(const void* pack, const float* bias, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, float* G, float* Cst,
                                  const LstmPlanes& pl, const LstmFusedGeom& gm, bool split, bool save, cudaStream_t st) {
    if (gm.len <= 0 || gm.nseq <= 0) return cudaSuccess;
    if (split && !x_lo) return cudaErrorInvalidValue;
    CUtensorMap mh, ml;
    if (!make_x_map(&mh, x_hi, gm)) return cudaErrorInvalidValue;
    ml = mh;
    if (split && !make_x_map(&ml, x_lo, gm)) return cudaErrorInvalidValue;
    const uint8_t* b = static_cast<const uint8_t*>(pack);
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.hh_tm = reinterpret_cast<const uint32_t*>(b);
    a.hh_sm = reinterpret_cast<const uint4*>(b + (size_t)2 * 4 * 128 * 128 * 2);
    a.ih_tm = reinterpret_cast<const uint32_t*>(b + (size_t)2 * 4 * 128 * 128 * 2 + 2 * HH_SM_BYTES);
    a.ih_sm = reinterpret_cast<const uint4*>(b + (size_t)2 * 4 * 128 * 128 * 2 + 2 * HH_SM_BYTES + (size_t)2 * 4 * 128 * 64 * 2);
    a.bias = bias; a.G = G; a.Cst = Cst;
    a.h_hi = pl.h_hi; a.h_lo = pl.h_lo; a.hp_hi = pl.hp_hi; a.hp_lo = pl.hp_lo;
    a.inter = gm.inter; a.len = gm.len;
    a.nseq = gm.inter ? gm.K : gm.nseq;
    a.S = gm.S; a.B = gm.B;
    a.s_t = gm.inter ? gm.K : 1;
    const int tiles = gm.inter ? gm.B * ceil_div(gm.K, GN) : ceil_div(gm.nseq, GN);
    dim3 grid(ceil_div(tiles, 2), 2);
    const int pl_n = split ? 2 : 1;
    const int smem = HH_SM_BYTES + IH_SM_BYTES + pl_n * 2 * (2 * TILE) + 2 * pl_n * (2 * TILE) + 512 + 1024;
    cudaError_t e;
#define DP_FUSED(SP, SV)                                                                                     \
    do {                                                                                                     \
        e = cudaFuncSetAttribute(lstm_fused_fwd_kernel<SP, SV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e != cudaSuccess) return e;                                                                      \
        lstm_fused_fwd_kernel<SP, SV><<<grid, 352, smem, st>>>(mh, ml, a);                                   \
    } while (0)
    if (split) { if (save) DP_FUSED(true, true); else DP_FUSED(true, false); }
    else       { if (save) DP_FUSED(false, true); else DP_FUSED(false, false); }
#undef DP_FUSED
    return cudaGetLastError();
}

}  // namespace dp
