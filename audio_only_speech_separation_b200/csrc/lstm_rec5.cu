// Bidirectional LSTM recurrence (forward and BPTT) on the 5th-generation tensor cores (tcgen05 / TMEM, sm_100a).
//
// Replaces the nn.LSTM time loop of ProjRNN (look2hear/models/utils/gc3_basics.py:16,22) and its autograd BPTT for
// I = 64, H = 128, one layer, zero initial state, equal-length sequences; same interface as the mma.sync kernels of lstm.cu /
// lstm_bwd.cu (gate pre-activations G = x W_ih^T + b from the in-projection GEMM in packed column order dir*512 + unit*4 + gate).
//
// One CTA = one direction x up to 32 sequences (UMMA N = 32); the whole time loop runs inside the kernel.
//   forward   gates^T[512 x 32] = W_hh[512 x 128] h^T[128 x 32]: four M = 128 row tiles; tile j holds the i, f, g, o rows of units
//             32j .. 32j+31 ordered so that ONE tcgen05.ld.16x256b pair hands every thread all four gates of its 8 (unit, sequence)
//             cells (lane l of quadrant q: unit 32j + 8q + (l >> 2), sequences 8k + 2(l & 3) + {0, 1}) -- no shuffles, no exchange
//   BPTT      dh^T[128 x 32] = W_hh^T[128 x 512] dgates^T[512 x 32]: one M = 128 tile, K = 512; thread = (unit, 8 sequences)
//   weights   never leave the SM: bf16 "hi" half in TENSOR MEMORY (256 columns, the A operand of two of the three split products),
//             "lo" half (fp32-parity mode) in 128 KB of shared memory (128-byte-swizzled K-major image)
//   issue     one warp issues all tcgen05.mma; six instructions per elect.sync with the descriptors advanced inside the asm block:
//             measured on B200 (tests/tools/ubench_cluster.cu) ~40 cycles per M = 128, N <= 64, K = 16 instruction in such a block
//             against ~108 with one elect + predicate per instruction
//   overlap   forward: the K blocks of step t+1 are issued per 32-unit slice as soon as the cell update of that slice's tile is
//             done (accumulators double buffered in TMEM), so the tensor pipe runs through while the other tiles are updated
//
// SPLIT = true: bf16x3 (hi*hi + hi*lo + lo*hi), ex2/rcp activations -> fp32 parity mode ; false: single product, tanh.approx -> bf16 mode
#include <cuda.h>

#include <cstring>

#include "common.cuh"
#include "kernels.h"
#include "tc5_common.cuh"

namespace dp {
int lstm_seqs_per_cta(int nseq, int slots);
namespace {

constexpr int R5_NS = 32;             // sequence slots per CTA = UMMA N
constexpr int R5_THREADS = 640;       // warp 0: MMA issuer, warps 1..3: copy-out (forward), warps 4..19: cell update
constexpr int R5_IMG = 131072;        // bytes of one direction's weight image (tensor-memory rows or shared-memory image)

struct R5Args {
    const uint32_t* w_tm;   // hi half, tensor-memory rows [dir][128 lanes x 256 columns], forward: [dir][tile][lane][64]
    const uint4* w_sm;      // lo half, [dir][R5_IMG] swizzled shared-memory image
    float* G;
    float* H;
    float* Cst;
    const float* dH;
    float* dbias;
    __nv_bfloat16* dG_hi;
    __nv_bfloat16* dG_lo;
    LstmPlanes pl;
    SeqMap m;
    int spc;
};

__device__ __forceinline__ float4 ld_f4_ordered(const float* p) {  // volatile asm: stays where it is written
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// shared-memory accesses by 32-bit shared address (keeps the generic-address arithmetic out of the cell loops)
__device__ __forceinline__ void cp_async16_s(uint32_t saddr, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts16(uint32_t saddr, __nv_bfloat16 v) {
    asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(saddr), "h"(*reinterpret_cast<const uint16_t*>(&v)) : "memory");
}
__device__ __forceinline__ unsigned lds32(uint32_t saddr) {
    unsigned v;
    asm("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t saddr, const uint2& v) {
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};\n" ::"r"(saddr), "r"(v.x), "r"(v.y) : "memory");
}
// bf16 hi / lo split of two values with PACKED conversions only (cvt.rn.bf16x2.f32 runs on the FMA-side pipe; the scalar
// cvt.rn.bf16.f32 is an XU instruction and competes with ex2 / rcp): hi = bf16(a) | bf16(b) << 16, lo = bf16(a - hi_a) | bf16(b - hi_b) << 16
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo_elem, float hi_elem) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
    return d;
}
__device__ __forceinline__ void split_pair_packed(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = cvt_bf16x2(a, b);
    lo = cvt_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
__device__ __forceinline__ void sts16r(uint32_t saddr, uint32_t v) {   // low 16 bits of v
    asm volatile("{\n .reg .b16 lo, hi;\n mov.b32 {lo, hi}, %1;\n st.shared.b16 [%0], lo;\n}\n" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_f_ordered(const float* p) {
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];\n" : "=f"(v) : "l"(p));
    return v;
}
// 16 lanes x 32 columns of this warp's quadrant half: register i holds row (lane >> 2) + 8 * ((i >> 1) & 1),
// column 8 * (i >> 2) + 2 * (lane & 3) + (i & 1)   (checked on the GPU: tests/tools/ubench_cluster.cu, section I)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// mbarrier wait bounded by the clock (about a second): a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void r5_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)   // suspend-time hint (ns): the warp sleeps in hardware instead of spinning
            : "memory");
        if (done) break;
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 2000000000ll) __trap();
    }
}

// Two K = 16 blocks of one accumulator tile in one elected asm block: W_hi * x_hi, and in fp32-parity mode W_hi * x_lo and W_lo * x_hi.
// Operands are base values plus compile-time offsets (accumulator column DOFF, tensor-memory A column AOFF, descriptor offsets WOFF / XOFF in
// 16-byte units) so that the issuing warp keeps only the bases in registers; the second block is 8 columns / 32 bytes further.
// FIRST: the very first instruction overwrites the accumulator.
#define R5_MMA(D, A, B, ACC) " @q tcgen05.mma.cta_group::1.kind::f16 [" D "], " A ", " B ", %5, " ACC ";\n"
template <bool SPLIT, bool FIRST, int DOFF, int AOFF, int WOFF, int XOFF>
__device__ __forceinline__ void r5_issue(uint32_t d, uint32_t a_tm, uint64_t d_wlo, uint64_t d_xhi, uint64_t d_xlo, uint32_t idesc) {
    if (SPLIT) {
        if (FIRST) {
            asm volatile(
                "{\n .reg .pred q;\n .reg .b64 wl, xh, xl;\n .reg .b32 at, dd;\n elect.sync _|q, 0xffffffff;\n"
                " add.u32 dd, %0, %6;\n add.u32 at, %1, %7;\n add.u64 wl, %2, %8;\n add.u64 xh, %3, %9;\n add.u64 xl, %4, %9;\n"
                R5_MMA("dd", "[at]", "xh", "0") R5_MMA("dd", "[at]", "xl", "1") R5_MMA("dd", "wl", "xh", "1")
                " add.u32 at, at, 8;\n add.u64 wl, wl, 2;\n add.u64 xh, xh, 2;\n add.u64 xl, xl, 2;\n"
                R5_MMA("dd", "[at]", "xh", "1") R5_MMA("dd", "[at]", "xl", "1") R5_MMA("dd", "wl", "xh", "1") "}\n" ::"r"(d),
                "r"(a_tm), "l"(d_wlo), "l"(d_xhi), "l"(d_xlo), "r"(idesc), "n"(DOFF), "n"(AOFF), "n"(WOFF), "n"(XOFF)
                : "memory");
        } else {
            asm volatile(
                "{\n .reg .pred q;\n .reg .b64 wl, xh, xl;\n .reg .b32 at, dd;\n elect.sync _|q, 0xffffffff;\n"
                " add.u32 dd, %0, %6;\n add.u32 at, %1, %7;\n add.u64 wl, %2, %8;\n add.u64 xh, %3, %9;\n add.u64 xl, %4, %9;\n"
                R5_MMA("dd", "[at]", "xh", "1") R5_MMA("dd", "[at]", "xl", "1") R5_MMA("dd", "wl", "xh", "1")
                " add.u32 at, at, 8;\n add.u64 wl, wl, 2;\n add.u64 xh, xh, 2;\n add.u64 xl, xl, 2;\n"
                R5_MMA("dd", "[at]", "xh", "1") R5_MMA("dd", "[at]", "xl", "1") R5_MMA("dd", "wl", "xh", "1") "}\n" ::"r"(d),
                "r"(a_tm), "l"(d_wlo), "l"(d_xhi), "l"(d_xlo), "r"(idesc), "n"(DOFF), "n"(AOFF), "n"(WOFF), "n"(XOFF)
                : "memory");
        }
    } else {
        if (FIRST) {
            asm volatile(
                "{\n .reg .pred q;\n .reg .b64 xh;\n .reg .b32 at, dd;\n elect.sync _|q, 0xffffffff;\n"
                " add.u32 dd, %0, %6;\n add.u32 at, %1, %7;\n add.u64 xh, %3, %9;\n"
                R5_MMA("dd", "[at]", "xh", "0")
                " add.u32 at, at, 8;\n add.u64 xh, xh, 2;\n"
                R5_MMA("dd", "[at]", "xh", "1") "}\n" ::"r"(d),
                "r"(a_tm), "l"(d_wlo), "l"(d_xhi), "l"(d_xlo), "r"(idesc), "n"(DOFF), "n"(AOFF), "n"(WOFF), "n"(XOFF)
                : "memory");
        } else {
            asm volatile(
                "{\n .reg .pred q;\n .reg .b64 xh;\n .reg .b32 at, dd;\n elect.sync _|q, 0xffffffff;\n"
                " add.u32 dd, %0, %6;\n add.u32 at, %1, %7;\n add.u64 xh, %3, %9;\n"
                R5_MMA("dd", "[at]", "xh", "1")
                " add.u32 at, at, 8;\n add.u64 xh, xh, 2;\n"
                R5_MMA("dd", "[at]", "xh", "1") "}\n" ::"r"(d),
                "r"(a_tm), "l"(d_wlo), "l"(d_xhi), "l"(d_xlo), "r"(idesc), "n"(DOFF), "n"(AOFF), "n"(WOFF), "n"(XOFF)
                : "memory");
        }
    }
}

// forward step s, K blocks 2 PQ and 2 PQ + 1 (hidden units 32 PQ .. 32 PQ + 31 = the h slice tile PQ's warps write) for all four gate tiles of one
// 32-sequence group.  d_w / d_xh / d_xl: descriptors of the W_lo image, the h hi plane and the h lo plane at offset 0.  Every product is
// M = 128, N = 32: a small-N tcgen05.mma costs ~40 cycles whatever N <= 64 is (it streams its 4 KB A tile), so a CTA with 64 sequences runs
// TWO independent groups half a step apart -- one group's products overlap the other group's cell update -- instead of one N = 64 product
// whose cell update (twice as long) cannot overlap anything.  NS = 64: group GRP owns accumulator columns 128 GRP .. 128 GRP + 127 and rows
// 32 GRP .. 32 GRP + 31 of the h tile (+4 096 bytes inside every K block).
template <bool SPLIT, int PQ, int J, int NS, int GRP>
__device__ __forceinline__ void r5_fwd_tile(uint32_t dcol, uint32_t tmem, uint64_t d_w, uint64_t d_xh, uint64_t d_xl) {
    constexpr uint32_t IDESC = idesc_bf16(128, 32, 0, 0);
    constexpr int KB = NS * 128;
    r5_issue<SPLIT, PQ == 0, GRP * 128 + J * 32, J * 64 + PQ * 16, ((J * 2 + (PQ >> 1)) * 16384) / 16 + (PQ & 1) * 4,
             ((PQ >> 1) * KB) / 16 + (PQ & 1) * 4 + GRP * 256>(dcol, tmem, d_w, d_xh, d_xl, IDESC);
}
// NS = 32: one group, accumulators double buffered by step parity (dcol), block PQ starts as soon as h slice PQ of the previous step is
// written.  NS = 64: two groups with one accumulator set each; a group's first block waits for all of the group's cell-update warps (they
// have read the accumulators before they publish h) -- the other group's products fill that time.
template <bool SPLIT, int PQ, int NS, int GRP>
__device__ __forceinline__ void r5_fwd_block(uint64_t* h_ready, uint64_t* d_full, int s, uint32_t dcol, uint32_t tmem, uint64_t d_w, uint64_t d_xh,
                                             uint64_t d_xl) {
    if (NS == 32) {
        r5_wait(h_ready + PQ, (s - 1) & 1);
    } else if (PQ == 0) {   // the block overwrites all four accumulators of the group: every tile's warps must have read the previous step's
#pragma unroll
        for (int j = 0; j < 4; ++j) r5_wait(h_ready + GRP * 4 + j, (s - 1) & 1);
    }
    tc_fence_after();
    uint64_t* df = NS == 32 ? d_full + (s & 1) * 4 : d_full + GRP * 4;
    r5_fwd_tile<SPLIT, PQ, 0, NS, GRP>(dcol, tmem, d_w, d_xh, d_xl);
    if (PQ == 3) umma_commit_w(df + 0);
    r5_fwd_tile<SPLIT, PQ, 1, NS, GRP>(dcol, tmem, d_w, d_xh, d_xl);
    if (PQ == 3) umma_commit_w(df + 1);
    r5_fwd_tile<SPLIT, PQ, 2, NS, GRP>(dcol, tmem, d_w, d_xh, d_xl);
    if (PQ == 3) umma_commit_w(df + 2);
    r5_fwd_tile<SPLIT, PQ, 3, NS, GRP>(dcol, tmem, d_w, d_xh, d_xl);
    if (PQ == 3) umma_commit_w(df + 3);
}

// BPTT step, unit quadrant Q: packed gate columns 128 Q .. 128 Q + 127 = K blocks 8 Q .. 8 Q + 7 = 64-wide blocks 2 Q, 2 Q + 1
template <bool SPLIT, int Q, int H>
__device__ __forceinline__ void r5_bwd_pair(uint32_t dacc, uint32_t tmem, uint64_t d_w, uint64_t d_xh, uint64_t d_xl) {
    constexpr uint32_t IDESC = idesc_bf16(128, R5_NS, 0, 0);
    constexpr int KB = R5_NS * 128;
    constexpr int kb = 2 * Q + (H >> 1);
    r5_issue<SPLIT, (Q == 0 && H == 0), 256, (8 * Q + 2 * H) * 8, (kb * 16384) / 16 + (H & 1) * 4, (kb * KB) / 16 + (H & 1) * 4>(dacc, tmem, d_w, d_xh,
                                                                                                                               d_xl, IDESC);
}
template <bool SPLIT, int Q>
__device__ __forceinline__ void r5_bwd_block(uint64_t* g_ready, int s, uint32_t tmem, uint64_t d_w, uint64_t d_xh, uint64_t d_xl) {
    // The accumulator is double buffered (columns 256 + 32 (s & 1)): the first product of a step overwrites all 128 lanes, while the
    // cell-update warps of the other unit quadrants may still be reading the previous step's accumulator.
    const uint32_t dacc = tmem + (uint32_t)((s & 1) * 32);
    r5_wait(g_ready + Q, (s - 1) & 1);
    tc_fence_after();
    r5_bwd_pair<SPLIT, Q, 0>(dacc, tmem, d_w, d_xh, d_xl);
    r5_bwd_pair<SPLIT, Q, 1>(dacc, tmem, d_w, d_xh, d_xl);
    r5_bwd_pair<SPLIT, Q, 2>(dacc, tmem, d_w, d_xh, d_xl);
    r5_bwd_pair<SPLIT, Q, 3>(dacc, tmem, d_w, d_xh, d_xl);
}

// register budgets per warpgroup: the issuer / copy-out group hands registers to the four cell-update groups.  The pool is the CTA's own
// launch allocation (640 threads x 96 registers): 128 x 40 + 512 x 104 = 58 368 <= 61 440 (an increase beyond the pool would block for ever)
__device__ __forceinline__ void r5_regs_small() { asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n" ::: "memory"); }
__device__ __forceinline__ void r5_regs_large() { asm volatile("setmaxnreg.inc.sync.aligned.u32 104;\n" ::: "memory"); }
static_assert(128 * 40 + 512 * 104 <= 640 * 96, "setmaxnreg budget exceeds the launch allocation");

// position bases of the tile's sequences (invalid slots alias the tile's first sequence: loads harmless, stores suppressed).
// SPREAD: the v-th sequence sits in slot (v & 3) * 8 + (v >> 2) (BPTT: equal work for the four sequence octets), else in slot v.
template <bool SPREAD, int NS = R5_NS>
__device__ __forceinline__ void r5_fill_bases(int* sbase, int q0, int nv, const SeqMap& m) {
    if (threadIdx.x < NS) {
        const int n = (int)threadIdx.x;
        const int v = SPREAD ? (n & 7) * 4 + (n >> 3) : n;
        const int q = v < nv ? q0 + v : q0;
        sbase[n] = (int)((q / m.qdiv) * m.s_hi + (q % m.qdiv) * m.s_lo);
    }
}

// ================================================================================================ forward
template <bool SPLIT, bool SAVE, int NS>
__global__ void __launch_bounds__(R5_THREADS, 1) lstm_rec5_fwd_kernel(const R5Args p) {
    constexpr int PL = SPLIT ? 2 : 1;
    constexpr int HALVES = NS / 32;                 // 32-sequence column blocks of the accumulators (NS = 32 or 64 sequence slots)
    constexpr int KB = NS * 128;                    // one 64-wide K block of the h tile: [NS rows][128 B]
    constexpr int PLANE = 2 * KB;                   // one plane (hi or lo) of the h tile
    constexpr int OFF_H = SPLIT ? R5_IMG : 0;
    constexpr int OFF_G = OFF_H + PL * PLANE;       // staged gate pre-activations: [16 warps][8 cells][32 lanes] x 16 B = 64 KB
    constexpr int OFF_BAR = OFF_G + 16 * 8 * 32 * 16;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* d_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);  // [2 buffers (NS = 32) or 2 sequence groups (NS = 64)][4 tiles]
    uint64_t* h_ready = d_full + 8;                                   // [group][4 tiles]
    uint64_t* h_copied = h_ready + 8;                                 // [group]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_copied + 2);
    int* sbase = reinterpret_cast<int*>(tmem_slot + 2);               // [NS]
    uint8_t* hs = smem + OFF_H;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const int len = p.m.len;
    const int q0 = blockIdx.x * p.spc;
    const int nv = min(p.spc, p.m.nseq - q0);
    const long long s_t = p.m.s_t;

    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(d_full + i, 1);
        for (int i = 0; i < 8; ++i) mbar_init(h_ready + i, 128);
        for (int i = 0; i < 2; ++i) mbar_init(h_copied + i, 3);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    r5_fill_bases<false, NS>(sbase, q0, nv, p.m);
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (SPLIT) {
        const uint4* src = p.w_sm + (size_t)dir * (R5_IMG / 16);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < R5_IMG / 16; i += R5_THREADS) dst[i] = src[i];
        proxy_fence_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp >= 4 && warp < 8) {  // hi half -> tensor memory columns 0..255 (tile j at column 64 j; one column = two k values)
        const int q = warp & 3;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const uint32_t* src = p.w_tm + ((size_t)(dir * 4 + j) * 128 + q * 32 + lane) * 64;
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
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < 4) r5_regs_small(); else r5_regs_large();
    if (warp == 0) {
        // ===================== MMA issuer =====================
        const uint64_t d_w = desc_sw128(smem_u32(smem), 16, 1024), d_xh = desc_sw128(smem_u32(hs), 16, 1024);
        const uint64_t d_xl = d_xh + (uint64_t)(PLANE >> 4);
        for (int s = 1; s < len; ++s) {
            const uint32_t dcol = tmem + 256 + (NS == 32 ? (uint32_t)((s & 1) * 128) : 0u);
            r5_fwd_block<SPLIT, 0, NS, 0>(h_ready, d_full, s, dcol, tmem, d_w, d_xh, d_xl);
            r5_fwd_block<SPLIT, 1, NS, 0>(h_ready, d_full, s, dcol, tmem, d_w, d_xh, d_xl);
            r5_fwd_block<SPLIT, 2, NS, 0>(h_ready, d_full, s, dcol, tmem, d_w, d_xh, d_xl);
            r5_fwd_block<SPLIT, 3, NS, 0>(h_ready, d_full, s, dcol, tmem, d_w, d_xh, d_xl);
            if (NS == 64) {   // the second group's products run while the first group's cells are updated, and vice versa
                r5_fwd_block<SPLIT, 0, NS, 1>(h_ready, d_full, s, dcol, tmem, d_w, d_xh, d_xl);
                r5_fwd_block<SPLIT, 1, NS, 1>(h_ready, d_full, s, dcol, tmem, d_w, d_xh, d_xl);
                r5_fwd_block<SPLIT, 2, NS, 1>(h_ready, d_full, s, dcol, tmem, d_w, d_xh, d_xl);
                r5_fwd_block<SPLIT, 3, NS, 1>(h_ready, d_full, s, dcol, tmem, d_w, d_xh, d_xl);
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        // ===================== copy-out (3 warps): h tile -> H planes (own position) and h_prev planes (next position) =====================
        const int cl = (warp - 1) * 32 + lane;
        if (SAVE && p.pl.hp_hi != nullptr) {  // h_prev of the first visited step is zero
            const long long t0 = (long long)(dir ? len - 1 : 0) * s_t;
            for (int ch = cl; ch < NS * 16; ch += 96) {
                const int n = ch >> 4, u = ch & 15;
                if (n >= nv) continue;
                const size_t o = (size_t)(sbase[n] + t0) * 256 + dir * kH + u * 8;
                *reinterpret_cast<uint4*>(p.pl.hp_hi + o) = make_uint4(0, 0, 0, 0);
                if (SPLIT && p.pl.hp_lo != nullptr) *reinterpret_cast<uint4*>(p.pl.hp_lo + o) = make_uint4(0, 0, 0, 0);
            }
        }
        const bool any = p.pl.h_hi != nullptr || p.pl.hp_hi != nullptr;
        for (int s = 0; s < len; ++s) {
            const int t = dir ? len - 1 - s : s;
            const long long toff = (long long)t * s_t;
            const long long tnext = (long long)(dir ? t - 1 : t + 1) * s_t;
            const bool has_next = s + 1 < len;
#pragma unroll
            for (int grp = 0; grp < HALVES; ++grp) {
#pragma unroll
                for (int pq = 0; pq < 4; ++pq) r5_wait(h_ready + grp * 4 + pq, s & 1);
                if (any) {
                    for (int ch = grp * 512 + cl; ch < (grp + 1) * 512; ch += 96) {
                        const int n = ch >> 4, u = ch & 15;
                        if (n >= nv) continue;
                        const uint32_t off = (u >> 3) * KB + n * 128 + (((u & 7) ^ (n & 7)) << 4);
                        const uint4 vh = *reinterpret_cast<const uint4*>(hs + off);
                        uint4 vl = make_uint4(0, 0, 0, 0);
                        if (SPLIT) vl = *reinterpret_cast<const uint4*>(hs + PLANE + off);
                        const size_t col = (size_t)dir * kH + u * 8;
                        if (p.pl.h_hi != nullptr) {
                            const size_t o = (size_t)(sbase[n] + toff) * 256 + col;
                            *reinterpret_cast<uint4*>(p.pl.h_hi + o) = vh;
                            if (SPLIT && p.pl.h_lo != nullptr) *reinterpret_cast<uint4*>(p.pl.h_lo + o) = vl;
                        }
                        if (SAVE && has_next && p.pl.hp_hi != nullptr) {
                            const size_t o = (size_t)(sbase[n] + tnext) * 256 + col;
                            *reinterpret_cast<uint4*>(p.pl.hp_hi + o) = vh;
                            if (SPLIT && p.pl.hp_lo != nullptr) *reinterpret_cast<uint4*>(p.pl.hp_lo + o) = vl;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(h_copied + grp);
            }
        }
    } else {
        // ===================== cell update: warp = (tile j, quadrant q), thread = unit 32j + 8q + (lane >> 2), 8 sequence slots =====================
        // tile 0 goes to the highest warp ids: the scheduler favours them, so the slice the next step's first K blocks wait for is done first
        const int j = 3 - ((warp - 4) >> 2), q = warp & 3, u8 = lane >> 2, c = lane & 3;
        const int unit = 32 * j + 8 * q + u8;
        const uint32_t acc_addr = tmem + ((uint32_t)(q * 32) << 16) + 256 + (uint32_t)(j * 32);   // + 128 per buffer (NS = 32) or sequence group (NS = 64)
        float* const Gc = p.G + (size_t)dir * kG + unit * 4;
        float* const Cc = p.Cst + (size_t)dir * kH + unit;
        float* const Hc = p.H + (size_t)dir * kH + unit;
        const bool has_h = p.H != nullptr;
        // even unit of a pair stores sequence 2c of both units, odd unit sequence 2c + 1: 4 bytes at the even unit's column
        // (row n = 8k + 2c + odd -> 1024 k + rowe * 128, swizzle term (chunk ^ rowe) << 4, independent of k)
        const int chunk = (unit & 63) >> 3;
        const int odd = u8 & 1, rowe = 2 * c + odd;
        const uint32_t hrow = smem_u32(hs) + (uint32_t)((unit >> 6) * KB + (u8 & 6) * 2 + rowe * 128 + ((chunk ^ rowe) << 4));
        const uint32_t sel_send = odd ? 0x5410u : 0x7632u, sel_hi = odd ? 0x3254u : 0x5410u, sel_lo = odd ? 0x3276u : 0x7610u;
        // the step's gate pre-activations (one 128-bit word per cell) are staged ahead into this thread's own shared-memory slots: eight slots =
        // one 32-sequence group.  NS = 64: as soon as a pair of slots has been consumed it is requested again for the cells the thread updates
        // next (the other group's of this step, then the first group's of the next step)
        const uint32_t gst_s = smem_u32(smem + OFF_G) + (uint32_t)(((warp - 4) * 256 + lane) * 16);
        const int kmax = (nv + 7) >> 3;   // sequence octets in use (warp-uniform)
        unsigned sb[8 * HALVES];
        unsigned valid = 0;
        float cst[8 * HALVES];
#pragma unroll
        for (int i = 0; i < 8 * HALVES; ++i) {
            const int n = 8 * (i >> 1) + 2 * c + (i & 1);
            sb[i] = (unsigned)sbase[n];
            valid |= (n < nv ? 1u : 0u) << i;
            cst[i] = 0.f;
        }
        const int dstep = dir ? -(int)s_t : (int)s_t;
        unsigned toff = dir ? (unsigned)((len - 1) * (int)s_t) : 0u;
        auto stage = [&](int hf, unsigned to) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (4 * hf + (i >> 1) < kmax) cp_async16_s(gst_s + i * 512, Gc + (size_t)(sb[8 * hf + i] + to) * 1024);
            asm volatile("cp.async.commit_group;\n" ::);
        };
        stage(0, toff);
        for (int s = 0; s < len; ++s) {
            if (NS == 32 && s > 0) {
                r5_wait(d_full + (s & 1) * 4 + j, (uint32_t)(((s - 2 + (s & 1)) >> 1) & 1));
                tc_fence_after();
            }
#pragma unroll
            for (int hf = 0; hf < HALVES; ++hf) {
                if (NS == 64 && s > 0) {   // this group's accumulators (the other group's products are in flight meanwhile)
                    r5_wait(d_full + hf * 4 + j, (uint32_t)((s - 1) & 1));
                    tc_fence_after();
                }
                uint32_t ra[16], rb[16];
                if (s > 0) {
                    const uint32_t col = NS == 32 ? (uint32_t)((s & 1) * 128) : (uint32_t)(hf * 128);
                    tmem_ld_16x256b_x4(acc_addr + col, ra);                  // lanes 0..15: i rows, f rows
                    tmem_ld_16x256b_x4(acc_addr + col + (16u << 16), rb);    // lanes 16..31: g rows, o rows
                    tmem_ld_wait();
                    r5_wait(h_copied + hf, (s - 1) & 1);   // the copy-out warps are done with the group's previous h rows
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { ra[i] = 0u; rb[i] = 0u; }
                }
                asm volatile("cp.async.wait_all;\n" ::: "memory");   // this thread's own staged words
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (4 * hf + k < kmax) {
                        float hh[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int i = 8 * hf + 2 * k + e;
                            const float4 gp = lds128(gst_s + (2 * k + e) * 512);
                            const bool vld = (valid >> i) & 1u;
                            const float ig = sigmoid_cell<SPLIT>(__uint_as_float(ra[4 * k + e]) + gp.x);
                            const float fg = sigmoid_cell<SPLIT>(__uint_as_float(ra[4 * k + 2 + e]) + gp.y);
                            const float gg = tanh_cell<SPLIT>(__uint_as_float(rb[4 * k + e]) + gp.z);
                            const float og = sigmoid_cell<SPLIT>(__uint_as_float(rb[4 * k + 2 + e]) + gp.w);
                            const float cc = fmaf(fg, cst[i], ig * gg);
                            cst[i] = cc;
                            hh[e] = og * tanh_cell<SPLIT>(cc);
                            const size_t pos = (size_t)(sb[i] + toff);
                            if (has_h) stg_pred(Hc + pos * 256, hh[e], vld);
                            if (SAVE) {
                                stg_pred(Cc + pos * 256, cc, vld);
                                stg_pred(reinterpret_cast<float4*>(Gc + pos * 1024), ig, fg, gg, og, vld);
                            }
                        }
                        // h -> bf16 hi / lo with packed conversions; the two lanes of a unit pair (lane ^ 4) swap one sequence each, so that every
                        // lane stores ONE 32-bit word (two neighbouring units of one sequence) per plane: no 16-bit stores, no bank conflicts
                        uint32_t Hw, Lw = 0u;
                        if (SPLIT) split_pair_packed(hh[0], hh[1], Hw, Lw);
                        else Hw = cvt_bf16x2(hh[0], hh[1]);
                        const uint32_t recv = __shfl_xor_sync(0xffffffffu, __byte_perm(Hw, Lw, sel_send), 4);
                        sts32(hrow + (uint32_t)((4 * hf + k) * 1024), __byte_perm(Hw, recv, sel_hi));
                        if (SPLIT) sts32(hrow + (uint32_t)((4 * hf + k) * 1024 + PLANE), __byte_perm(Lw, recv, sel_lo));
                    }
                    if (NS == 64) {   // slots 2k, 2k + 1 are free: the words of the cells this thread updates next
                        const int nh = hf ^ 1;
                        const unsigned to = hf == 0 ? toff : toff + (unsigned)dstep;
                        if ((hf == 0 || s + 1 < len) && 4 * nh + k < kmax) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) cp_async16_s(gst_s + (2 * k + e) * 512, Gc + (size_t)(sb[8 * nh + 2 * k + e] + to) * 1024);
                        }
                    }
                }
                if (NS == 64) {
                    asm volatile("cp.async.commit_group;\n" ::);
                    proxy_fence_async();  // the group's h rows (generic-proxy stores) -> visible to the tensor core's async proxy
                    tc_fence_before();
                    mbar_arrive(h_ready + hf * 4 + j);
                }
            }
            if (NS == 32) {
                proxy_fence_async();  // h slice (generic-proxy stores) -> visible to the tensor core's async proxy
                tc_fence_before();
                mbar_arrive(h_ready + j);
            }
            toff += (unsigned)dstep;
            if (NS == 32 && s + 1 < len) stage(0, toff);   // next step: the slots were read above
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ================================================================================================ BPTT
// Per step (reverse of the forward visiting order):  dh_t = dH_t + W_hh^T dgates_{t+1};  dc_t = dh_t o (1 - tanh^2 c_t) + dc_{t+1} f_{t+1};
// dgates_t = (dc g i(1-i), dc c_{t-1} f(1-f), dc i (1-g^2), dh tanh(c_t) o(1-o)) -> dG planes (or G in place) and, as bf16 hi/lo, the
// shared-memory B operand of the next step's product.  Bias gradient accumulated in registers, one atomic per thread and gate at the end.
template <bool SPLIT>
__global__ void __launch_bounds__(R5_THREADS, 1) lstm_rec5_bwd_kernel(const R5Args p) {
    constexpr int PL = SPLIT ? 2 : 1;
    constexpr int KB = R5_NS * 128;                 // one 64-wide K block of the dgates tile: [32 rows][128 B]
    constexpr int PLANE = 8 * KB;                   // K = 512: 32 KB per plane
    constexpr int OFF_D = SPLIT ? R5_IMG : 0;
    constexpr int OFF_BAR = OFF_D + PL * PLANE;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* d_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* g_ready = d_full + 1;                                   // [4 unit quadrants]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_ready + 4);
    int* sbase = reinterpret_cast<int*>(tmem_slot + 2);
    uint8_t* ds = smem + OFF_D;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y;
    const int len = p.m.len;
    const int q0 = blockIdx.x * p.spc;
    const int nv = min(p.spc, p.m.nseq - q0);
    const long long s_t = p.m.s_t;

    if (tid == 0) {
        mbar_init(d_full, 1);
        for (int i = 0; i < 4; ++i) mbar_init(g_ready + i, 128);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    r5_fill_bases<true>(sbase, q0, nv, p.m);
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (SPLIT) {
        const uint4* src = p.w_sm + (size_t)dir * (R5_IMG / 16);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < R5_IMG / 16; i += R5_THREADS) dst[i] = src[i];
        proxy_fence_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp >= 4 && warp < 8) {  // W_hh^T hi -> tensor memory columns 0..255 (lane = unit, column = two packed gate columns)
        const int q = warp & 3;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t* src = p.w_tm + ((size_t)dir * 128 + q * 32 + lane) * 256;
#pragma unroll 1
        for (int c = 0; c < 256; c += 32) {
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const uint4 u = *reinterpret_cast<const uint4*>(src + c + i);
                v[i] = u.x; v[i + 1] = u.y; v[i + 2] = u.z; v[i + 3] = u.w;
            }
            tmem_st32(lane_addr + c, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < 4) r5_regs_small(); else r5_regs_large();
    if (warp == 0) {
        // ===================== MMA issuer: 32 K blocks (packed gate columns 16 kk ..) x split products into one accumulator =====================
        const uint64_t d_w = desc_sw128(smem_u32(smem), 16, 1024), d_xh = desc_sw128(smem_u32(ds), 16, 1024);
        const uint64_t d_xl = d_xh + (uint64_t)(PLANE >> 4);
        for (int s = 1; s < len; ++s) {
            r5_bwd_block<SPLIT, 0>(g_ready, s, tmem, d_w, d_xh, d_xl);
            r5_bwd_block<SPLIT, 1>(g_ready, s, tmem, d_w, d_xh, d_xl);
            r5_bwd_block<SPLIT, 2>(g_ready, s, tmem, d_w, d_xh, d_xl);
            r5_bwd_block<SPLIT, 3>(g_ready, s, tmem, d_w, d_xh, d_xl);
            umma_commit_w(d_full);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===================== cell backward: warp = (unit quadrant q, sequence octet part), thread = unit 32q + lane =====================
        // the CTA's v-th sequence sits in slot (v & 3) * 8 + (v >> 2): every octet holds the same number of sequences (+- 1)
        const int q = warp & 3, part = (warp - 4) >> 2;
        const int unit = 32 * q + lane;
        const uint32_t acc_addr = tmem + ((uint32_t)(q * 32) << 16) + 256 + (uint32_t)(part * 8);
        float* const Gc = p.G + (size_t)dir * kG + unit * 4;
        const float* const Cc = p.Cst + (size_t)dir * kH + unit;
        const float* const Dc = p.dH + (size_t)dir * kH + unit;
        __nv_bfloat16* const Ph = p.dG_hi + (size_t)dir * kG + unit * 4;
        __nv_bfloat16* const Pl = p.dG_lo + (size_t)dir * kG + unit * 4;
        const bool planes = p.dG_hi != nullptr, planes_lo = p.dG_lo != nullptr;
        // dgates tile offsets: row n = 8 part + i, 8 bytes at packed columns 4 unit ..: block unit >> 4, chunk (unit & 15) >> 1; n & 7 = i
        const int chunk = (unit & 15) >> 1;
        const uint32_t ds_s = smem_u32(ds) + (uint32_t)((unit >> 4) * KB + (unit & 1) * 8 + part * 1024);
        const int cnt = nv > part ? (nv - part + 3) >> 2 : 0;   // sequences of this octet (warp-uniform)
        const uint32_t sb_s = smem_u32(sbase) + (uint32_t)(part * 32);   // position bases stay in shared memory (register budget)
        float dcs[8], ct[8], bsum[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) dcs[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) bsum[i] = 0.f;
        const int dstep = dir ? (int)s_t : -(int)s_t;               // reverse of the forward visiting order
        unsigned toff = dir ? 0u : (unsigned)((len - 1) * (int)s_t);
#pragma unroll
        for (int i = 0; i < 8; ++i) ct[i] = i < cnt ? ld_f_ordered(Cc + (size_t)(lds32(sb_s + i * 4) + toff) * 256) : 0.f;
        for (int s = 0; s < len; ++s) {
            const bool first = (s == len - 1);       // the forward pass's first step: c_prev = 0
            float4 a[8];
            float cp[8], dh[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < cnt) {
                    const size_t pos = (size_t)(lds32(sb_s + i * 4) + toff);
                    a[i] = ld_f4_ordered(Gc + pos * 1024);
                    dh[i] = ld_f_ordered(Dc + pos * 256);
                    cp[i] = first ? 0.f : ld_f_ordered(Cc + (size_t)(lds32(sb_s + i * 4) + toff + (unsigned)dstep) * 256);
                }
            }
            if (!first) {   // next step's lines -> L2
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < cnt) {
                        const size_t pos = (size_t)(lds32(sb_s + i * 4) + toff + (unsigned)dstep);
                        prefetch_l2(Gc + pos * 1024);
                        prefetch_l2(Dc + pos * 256);
                    }
                }
            }
            float acc[8];
            if (s > 0) {
                r5_wait(d_full, (s - 1) & 1);
                tc_fence_after();
                tmem_ld8_nowait(acc_addr + (uint32_t)((s & 1) * 32), acc);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t off = ds_s + (uint32_t)(i * 128 + ((chunk ^ i) << 4));
                if (i < cnt) {
                    const float dhh = dh[i] + acc[i];
                    const float tc = tanh_cell<SPLIT>(ct[i]);
                    const float dc = fmaf(dhh * a[i].w, 1.f - tc * tc, dcs[i]);
                    dcs[i] = dc * a[i].y;
                    float4 dg;
                    dg.x = dc * a[i].z * a[i].x * (1.f - a[i].x);
                    dg.y = dc * cp[i] * a[i].y * (1.f - a[i].y);
                    dg.z = dc * a[i].x * (1.f - a[i].z * a[i].z);
                    dg.w = dhh * tc * a[i].w * (1.f - a[i].w);
                    ct[i] = cp[i];
                    uint2 hi, lo;
                    split_pair_packed(dg.x, dg.y, hi.x, lo.x);
                    split_pair_packed(dg.z, dg.w, hi.y, lo.y);
                    const size_t pos = (size_t)(lds32(sb_s + i * 4) + toff);
                    if (planes) {
                        *reinterpret_cast<uint2*>(Ph + pos * 1024) = hi;
                        if (SPLIT && planes_lo) *reinterpret_cast<uint2*>(Pl + pos * 1024) = lo;
                    } else {
                        *reinterpret_cast<float4*>(Gc + pos * 1024) = dg;
                    }
                    bsum[0] += dg.x; bsum[1] += dg.y; bsum[2] += dg.z; bsum[3] += dg.w;
                    sts64(off, hi);
                    if (SPLIT) sts64(off + PLANE, lo);
                } else if (s == 0) {   // unused slots: zero once, so that the products never read uninitialised shared memory
                    sts64(off, make_uint2(0u, 0u));
                    if (SPLIT) sts64(off + PLANE, make_uint2(0u, 0u));
                }
            }
            proxy_fence_async();
            tc_fence_before();
            mbar_arrive(g_ready + q);
            toff += (unsigned)dstep;
        }
        if (p.dbias != nullptr && cnt > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) atomicAdd(p.dbias + dir * kG + unit * 4 + i, bsum[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ weight images
struct R5PackArgs {
    const float* w_hh[2];
    uint16_t* f_tm;   // forward hi: [dir][tile j][lane][128 k]  (lane L of the tile: quadrant L >> 5, gate (L >> 3) & 3, unit 32j + 8(L >> 5) + (L & 7))
    uint8_t* f_sm;    // forward lo: [dir][(j*2 + kb)*16 KB + L*128 + swizzled chunk]
    uint16_t* b_tm;   // BPTT hi:    [dir][unit][512 packed gate columns]
    uint8_t* b_sm;    // BPTT lo:    [dir][kb*16 KB + unit*128 + swizzled chunk], kb = column >> 6
};
__device__ __forceinline__ uint16_t r5_bits(float v) {
    __nv_bfloat16 b = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&b);
}
__global__ void r5_pack_kernel(const R5PackArgs a) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 2 * 4 * 128 * 128) {  // forward
        const int k = idx & 127, L = (idx >> 7) & 127, j = (idx >> 14) & 3, d = idx >> 16;
        const int gate = (L >> 3) & 3, unit = 32 * j + 8 * (L >> 5) + (L & 7);
        const float v = a.w_hh[d][(size_t)(gate * kH + unit) * kH + k];
        const float vh = bf16_round(v);
        a.f_tm[idx] = r5_bits(vh);
        const int kb = k >> 6, c = (k & 63) >> 3, e = k & 7;
        const size_t off = (size_t)d * R5_IMG + ((size_t)((j * 2 + kb) * 128 + L)) * 128 + ((c ^ (L & 7)) << 4) + e * 2;
        *reinterpret_cast<uint16_t*>(a.f_sm + off) = r5_bits(v - vh);
        return;
    }
    idx -= 2 * 4 * 128 * 128;
    if (idx < 2 * 128 * 512) {  // BPTT: A[unit u][K = u'*4 + gate] = W_hh[gate*128 + u'][u]
        const int K = idx & 511, u = (idx >> 9) & 127, d = idx >> 16;
        const int up = K >> 2, gate = K & 3;
        const float v = a.w_hh[d][(size_t)(gate * kH + up) * kH + u];
        const float vh = bf16_round(v);
        a.b_tm[idx] = r5_bits(vh);
        const int kb = K >> 6, c = (K & 63) >> 3, e = K & 7;
        const size_t off = (size_t)d * R5_IMG + ((size_t)(kb * 128 + u)) * 128 + ((c ^ (u & 7)) << 4) + e * 2;
        *reinterpret_cast<uint16_t*>(a.b_sm + off) = r5_bits(v - vh);
    }
}

int g_rec5 = 1;  // 0: never, 1: automatic, 2: always (when the shape is supported)

}  // namespace

size_t lstm_rec5_pack_bytes() { return (size_t)4 * 2 * R5_IMG; }

cudaError_t launch_pack_lstm_rec5(const float* const w_hh[2], void* pack, cudaStream_t st) {
    R5PackArgs a;
    uint8_t* b = static_cast<uint8_t*>(pack);
    for (int d = 0; d < 2; ++d) a.w_hh[d] = w_hh[d];
    a.f_tm = reinterpret_cast<uint16_t*>(b);
    a.f_sm = b + (size_t)2 * R5_IMG;
    a.b_tm = reinterpret_cast<uint16_t*>(b + (size_t)4 * R5_IMG);
    a.b_sm = b + (size_t)6 * R5_IMG;
    const int total = 2 * 4 * 128 * 128 + 2 * 128 * 512;
    r5_pack_kernel<<<ceil_div(total, 256), 256, 0, st>>>(a);
    return cudaGetLastError();
}

int lstm_set_rec5(int mode) {
    if (mode < 0 || mode > 2) return -1;
    g_rec5 = mode;
    return 0;
}
int lstm_get_rec5() { return g_rec5; }

// The tcgen05 recurrence handles one wave of 32-sequence tiles per direction; larger passes run several waves.
// Automatic choice, measured on B200 (tests/tools/time_rec5.py, intra / inter pass, training mode, us):
//   B = 16 fp32: forward 395 / 357 against 503 / 420 (mma.sync), BPTT 348 / 310 against 645 / 545
//   B = 16 bf16: forward 378 / 379 against 320 / 297 -> mma.sync, BPTT 290 / 275 against 448 / 393
//   B = 32 fp32: forward with 64-sequence tiles (two groups of 32 half a step apart) 737 / 750 against 1032 / 849 (one N = 64 group:
//                846 / 880), BPTT 880 / 732 against 1322 / 1090;  B = 40: forward 977 / 959 against 1030 / 1263
bool lstm_rec5_wanted(const SeqMap& m, bool split, bool backward) {
    if (g_rec5 == 0) return false;
    if (g_rec5 == 2) return true;
    if (lstm_get_pipeline() != 1) return false;   // an explicitly selected mma.sync variant (dp_set_lstm_pipeline) is honoured
    if (m.nseq < 256) return false;               // small passes: the evenly spread mma.sync kernels keep more SMs busy
    if (backward) return true;
    if (!split) return false;                      // forward: fp32-parity mode only
    if (ceil_div(m.nseq, R5_NS) <= 74) return true;   // one wave of 32-sequence tiles
    return ceil_div(m.nseq, 64) <= 74;                // one wave of 64-sequence tiles (two groups of 32 per CTA)
}

cudaError_t launch_lstm_rec5_fwd(const void* pack, float* G, float* H, float* Cst, const SeqMap& m, bool split, bool save, cudaStream_t st,
                                 const LstmPlanes& pl) {
    if (m.nseq <= 0 || m.len <= 0) return cudaSuccess;
    const uint8_t* b = static_cast<const uint8_t*>(pack);
    R5Args a;
    memset(&a, 0, sizeof(a));
    a.w_tm = reinterpret_cast<const uint32_t*>(b);
    a.w_sm = reinterpret_cast<const uint4*>(b + (size_t)2 * R5_IMG);
    a.G = G; a.H = H; a.Cst = Cst; a.pl = pl; a.m = m;
    // 32-sequence tiles while one wave of them covers the pass (B <= 28 at the bench geometry), 64-sequence tiles beyond
    const int ns = ceil_div(m.nseq, R5_NS) <= 74 ? R5_NS : 64;
    a.spc = lstm_seqs_per_cta(m.nseq, ns);
    dim3 grid(ceil_div(m.nseq, a.spc), 2);
    const int smem = (split ? R5_IMG : 0) + (split ? 2 : 1) * 2 * ns * 128 + 16 * 8 * 32 * 16 + 512 + 1024;
    cudaError_t e;
#define DP_R5F(SP, SV, NSV)                                                                                              \
    do {                                                                                                                 \
        e = cudaFuncSetAttribute(lstm_rec5_fwd_kernel<SP, SV, NSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);  \
        if (e != cudaSuccess) return e;                                                                                  \
        lstm_rec5_fwd_kernel<SP, SV, NSV><<<grid, R5_THREADS, smem, st>>>(a);                                            \
    } while (0)
    if (ns == R5_NS) {
        if (split) { if (save) DP_R5F(true, true, 32); else DP_R5F(true, false, 32); }
        else       { if (save) DP_R5F(false, true, 32); else DP_R5F(false, false, 32); }
    } else {
        if (split) { if (save) DP_R5F(true, true, 64); else DP_R5F(true, false, 64); }
        else       { if (save) DP_R5F(false, true, 64); else DP_R5F(false, false, 64); }
    }
#undef DP_R5F
    return cudaGetLastError();
}

cudaError_t launch_lstm_rec5_bwd(const void* pack, float* G, const float* Cst, const float* dH, float* dbias, const SeqMap& m, bool split,
                                 cudaStream_t st, __nv_bfloat16* dG_hi, __nv_bfloat16* dG_lo) {
    if (m.nseq <= 0 || m.len <= 0) return cudaSuccess;
    const uint8_t* b = static_cast<const uint8_t*>(pack);
    R5Args a;
    memset(&a, 0, sizeof(a));
    a.w_tm = reinterpret_cast<const uint32_t*>(b + (size_t)4 * R5_IMG);
    a.w_sm = reinterpret_cast<const uint4*>(b + (size_t)6 * R5_IMG);
    a.G = G; a.Cst = const_cast<float*>(Cst); a.dH = dH; a.dbias = dbias; a.dG_hi = dG_hi; a.dG_lo = dG_lo; a.m = m;
    a.spc = lstm_seqs_per_cta(m.nseq, R5_NS);
    dim3 grid(ceil_div(m.nseq, a.spc), 2);
    const int smem = (split ? R5_IMG : 0) + (split ? 2 : 1) * 8 * R5_NS * 128 + 512 + 1024;
    cudaError_t e;
    if (split) {
        e = cudaFuncSetAttribute(lstm_rec5_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        lstm_rec5_bwd_kernel<true><<<grid, R5_THREADS, smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(lstm_rec5_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        lstm_rec5_bwd_kernel<false><<<grid, R5_THREADS, smem, st>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace dp
