// Internal launcher declarations (host side) for the dual-path kernels.
// Everything here takes raw device pointers; the public C-ABI lives in include/dualpath_b200.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dp {

// ---------------- GEMMs (gemm.cu) ----------------
// C[M,N] (=|+=) A[M,K] * W^T (+ bias_scale*bias) ; optional ReLU ; optional per-group sum/sumsq.
// W is given pre-split into bf16 hi/lo.  w_kn == 0: W stored [N,K] (K contiguous, ldw = row stride)
//                                        w_kn == 1: W stored [K,N] (N contiguous)
struct GemmNtArgs {
    const float* A;
    long long lda;
    int a_rpb, a_skip;  // row m lives at A + (m + (m / a_rpb) * a_skip) * lda   (a_rpb == 0: plain)
    const __nv_bfloat16* Whi;
    const __nv_bfloat16* Wlo;
    int ldw;
    int w_kn;
    float* C;
    int ldc;
    int M, N, K;
    const float* bias;
    float bias_scale;
    int relu;        // output activation: 0 none, 1 ReLU, 2 tanh, 3 sigmoid
    int accumulate;  // C += ...
    int mul_c;       // after the activation: C = act(...) * C_old  (gated tanh * sigmoid pair, sepformer.py:747)
    double* stats;  // [groups][2] (sum, sumsq) of the stored values, group = row / rows_per_group
    int rows_per_group;
    int relu_a;         // A is replaced by max(A, 0) on load (the ReLU between dptnet.py:79's LSTM and linear2)
    const float* mask;  // optional [M, ldmask]: C is zeroed where mask <= 0 (ReLU backward fused into the producing GEMM)
    int ldmask;
};
cudaError_t launch_gemm_nt(const GemmNtArgs& a, bool split, cudaStream_t st);
// tcgen05 / TMEM version (gemm_tc5.cu); supported(): K % 64 == 0, N in {64, 128, k*256}, w_kn == 0 preferred, no stats
bool gemm_nt_tc5_supported(const GemmNtArgs& a);
cudaError_t launch_gemm_nt_tc5(const GemmNtArgs& a, bool split, cudaStream_t st);

// ---- TMA-fed tcgen05 GEMMs on pre-split bf16 planes (gemm_tma.cu) ----
// C[M,N] = act(A[M,K] W[N,K]^T + bias_scale*bias) (+ C); A and W are bf16 hi (and lo) planes, K contiguous.
// Outputs: fp32 C and/or hi/lo planes (the operand format of the next GEMM).  K % 64 == 0, N % 64 == 0.
struct TmaGemmArgs {
    const __nv_bfloat16* A_hi;
    const __nv_bfloat16* A_lo;  // unused in bf16 mode
    long long lda;              // elements
    const __nv_bfloat16* W_hi;
    const __nv_bfloat16* W_lo;
    int ldw;
    int M, N, K;
    float* C;                   // optional
    int ldc;
    __nv_bfloat16* C_hi;        // optional
    __nv_bfloat16* C_lo;        // optional (with C_hi)
    int ldch;
    const float* bias;
    float bias_scale;
    int act;         // 0 none, 1 ReLU, 2 tanh, 3 sigmoid
    int accumulate;  // C += (needs C)
    const float* Cin;  // with accumulate: read the residual from here instead of C (C = Cin + ...), row stride ldc
    int mul_c;       // C = act(..) * C_old
    double* stats;   // [groups][2] (sum, sumsq) of the stored values
    int rows_per_group;
    const float* mask;  // optional [M, ldmask]: zero where mask <= 0
    int ldmask;
    const __nv_bfloat16* mask_hi;  // the same mask given as a bf16 plane (the hi plane of a ReLU output): zero where <= 0
    int ldmask_hi;
    // dropout on act(A W^T + bias) BEFORE the residual is added (nn.Dropout on a sub-layer output): kept elements * drop_scale.
    // drop_thr = p * 2^24 (0 = off); element (row, col) of the output is kept iff drop_keep(drop_key, row, col, drop_thr)
    unsigned drop_thr, drop_key;
    float drop_scale;
    float out_scale;   // 0 = none: multiplies the final value (after the masks)
};
bool gemm_tma_nt_supported(const TmaGemmArgs& a);
cudaError_t launch_gemm_tma_nt(const TmaGemmArgs& a, bool split, cudaStream_t st);
// Weight gradients on planes: C0[Mo,nb0] += scale * A^T B0 and (optionally, same pass over A) C1[Mo,nb1] += scale * A^T B1.
// A [P, >= Mo] and B [P, >= nb] are planes with the position as the slow index; Mo % 128 == 0, nb % 64 == 0, nb0 + nb1 <= 256.
struct TmaWgradArgs {
    const __nv_bfloat16* A_hi;
    const __nv_bfloat16* A_lo;
    long long lda;
    int Mo;
    const __nv_bfloat16* B0_hi;
    const __nv_bfloat16* B0_lo;
    long long ldb0;
    int nb0;
    const __nv_bfloat16* B1_hi;
    const __nv_bfloat16* B1_lo;
    long long ldb1;
    int nb1;
    float* C0;
    int ldc0;
    int transpose0;  // store C0[col][row]
    float* C1;
    int ldc1;
    int transpose1;
    int P;
    float scale;
};
bool gemm_tma_tn_supported(const TmaWgradArgs& a);
// 1: outputs with a multiple of four 128-row slices run as 4-CTA clusters that multicast the shared B tiles; 0 (default): never.  Returns the previous setting.
int gemm_tma_set_wgrad_multicast(int on);
cudaError_t launch_gemm_tma_tn(const TmaWgradArgs& a, bool split, cudaStream_t st);
// fp32 rows [rows, C] (row stride ld) -> bf16 hi / lo planes [rows, C] (lo optional); relu applies max(x, 0) first
cudaError_t launch_split_rows(const float* src, long long ld, __nv_bfloat16* hi, __nv_bfloat16* lo, long long rows, int C, int relu,
                              cudaStream_t st);

// out[c] += sum over rows of (hi + lo)[row, c] for a pair of planes (lo optional): bias gradients of tensors that exist as planes only
cudaError_t launch_colsum_planes(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long rows, int C, float* out, cudaStream_t st);
// launch_split_rows (no ReLU) that also accumulates the column sums of the rows into colsum[C] (atomics): bias gradients
// drop_thr != 0: the rows pass through the dropout mask of (drop_key, row, col) first (gradient of a dropped sub-layer output)
cudaError_t launch_split_rows_colsum(const float* src, long long ld, __nv_bfloat16* hi, __nv_bfloat16* lo, long long rows, int C,
                                     float* colsum, cudaStream_t st, unsigned drop_thr = 0, unsigned drop_key = 0, float drop_scale = 1.f);

// C[Mo,No] += scale * sum_p A[p,Mo]^T * B[p',No]   (fp32 atomics), p' = p + shift when the time index allows.
struct GemmTnArgs {
    const float* A;
    int lda;
    const float* B;
    long long ldb;
    int b_rpb, b_skip;
    int shift, tdiv, tmod;  // tmod == 0: no shift
    float* C;
    int ldc;
    int P, Mo, No;
    float scale;
    int relu_b;  // B is replaced by max(B, 0) on load
};
cudaError_t launch_gemm_tn(const GemmTnArgs& a, bool split, cudaStream_t st);
// tcgen05 / TMEM version: whole [Mo,No] output resident in TMEM per CTA; transpose_out writes C[col][row]
bool gemm_tn_tc5_supported(const GemmTnArgs& a);
cudaError_t launch_gemm_tn_tc5(const GemmTnArgs& a, bool split, int transpose_out, cudaStream_t st);

// out[n] += scale * sum_p A[p*lda + n]  (and the same into out2 when non-null)
cudaError_t launch_colsum(const float* A, int lda, int P, int N, float scale, float* out, float* out2, cudaStream_t st);

// geometry of the sequences of one dual-path pass over the stream [B,S,K,C] (shared by the TMA-fed sequence kernels)
struct LstmFusedGeom {
    int inter;  // 0: sequences (b,s) of the stream [B,S,K,64] walk k ; 1: sequences (b,k) walk s
    int len;    // time steps (K intra, S inter)
    int nseq;   // number of sequences (B*S intra, B*K inter)
    int K, S, B;
};

// ---------------- LSTM (lstm.cu) ----------------
struct SeqMap {  // position of (sequence q, time t):  (q / qdiv) * s_hi + (q % qdiv) * s_lo + t * s_t
    int nseq, len;
    int qdiv;
    long long s_hi, s_lo, s_t;
};
struct LstmPack {  // per ProjRNN, both directions
    const uint4* whh_f_hi;  // [2][8][4][8][32] fragment-ordered W_hh (forward recurrence A operand)
    const uint4* whh_f_lo;
    const uint4* whh_b_hi;  // [2][8][32][32]   fragment-ordered W_hh^T (backward recurrence A operand)
    const uint4* whh_b_lo;
    const void* rec5;       // weight images of the tcgen05 recurrence (lstm_rec5.cu), null: mma.sync kernels only
};
// G: [P,1024] gate pre-activations (packed column order dir*512 + unit*4 + gate); overwritten with the
// activated gates when save != 0.  H: [P,256] = [h_fwd | h_bwd].  Cst: [P,256] cell states (save only).
// Optional bf16 hi/lo operand planes written by the forward kernel (all [P,256] = [fwd | bwd] like H; lo unused in bf16 mode):
// h_* : h_t at its own position (operand of the out-projection and of its weight gradient)
// hp_*: h of the previously visited step at each position, zeros at a sequence's first step (operand of dW_hh = dG^T h_prev,
//       so that gradient needs neither a shifted read nor boundary masking)
struct LstmPlanes {
    __nv_bfloat16* h_hi;
    __nv_bfloat16* h_lo;
    __nv_bfloat16* hp_hi;
    __nv_bfloat16* hp_lo;
};
// forward recurrence kernel choice: 0 plain 8-warp kernels, 1 automatic, 2 software-pipelined sequence groups, 3 16-warp kernel
int lstm_set_pipeline(int mode);
int lstm_get_pipeline();
int lstm_set_cluster(int mode);   // 0 off, 1 automatic, 2 always: four-CTA-cluster forward recurrence for small inference passes
int lstm_get_cluster();
// H may be null when only the planes are wanted.
cudaError_t launch_lstm_fwd(const LstmPack& w, float* G, float* H, float* Cst, const SeqMap& m, bool split, bool save,
                            cudaStream_t st, const LstmPlanes* planes = nullptr);
// dH: [P,256] incoming gradient of H.  G holds activated gates on entry and d(pre-activations) on exit.
// dbias (optional, [1024] packed order) accumulates sum_p dG[p,:] (the bias gradient) inside the same kernel.
// ---- fused input projection + recurrence on tcgen05 (lstm_tc5.cu) ----
size_t lstm_tc5_pack_bytes();
// weight images for the fused kernel: bf16 hi rows (tensor-memory A operand) and swizzled lo images (shared-memory A operand)
cudaError_t launch_pack_lstm_tc5(const float* const w_ih[2], const float* const w_hh[2], void* pack, cudaStream_t st);
// x planes [P,64] -> H planes (+ h_prev planes, activated gates G [P,1024] and c_t Cst [P,256] when save); bias = packed b_ih + b_hh
cudaError_t launch_lstm_fused_fwd(const void* pack, const float* bias, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, float* G, float* Cst,
                                  const LstmPlanes& pl, const LstmFusedGeom& gm, bool split, bool save, cudaStream_t st);

// dG_hi / dG_lo (optional, [P,1024] planes): d(pre-activations) go there instead of overwriting G.
cudaError_t launch_lstm_bwd(const LstmPack& w, float* G, const float* Cst, const float* dH, float* dbias, const SeqMap& m, bool split,
                            cudaStream_t st, __nv_bfloat16* dG_hi = nullptr, __nv_bfloat16* dG_lo = nullptr);

// ---- tcgen05 recurrence (lstm_rec5.cu): same interface, W_hh hi in tensor memory, lo in shared memory ----
size_t lstm_rec5_pack_bytes();
cudaError_t launch_pack_lstm_rec5(const float* const w_hh[2], void* pack, cudaStream_t st);
// 0: never, 1: automatic (passes with enough sequences), 2: always
int lstm_set_rec5(int mode);
int lstm_get_rec5();
bool lstm_rec5_wanted(const SeqMap& m, bool split, bool backward);
cudaError_t launch_lstm_rec5_fwd(const void* pack, float* G, float* H, float* Cst, const SeqMap& m, bool split, bool save, cudaStream_t st,
                                 const LstmPlanes& pl);
cudaError_t launch_lstm_rec5_bwd(const void* pack, float* G, const float* Cst, const float* dH, float* dbias, const SeqMap& m, bool split,
                                 cudaStream_t st, __nv_bfloat16* dG_hi, __nv_bfloat16* dG_lo);

// Build all packed forms of one ProjRNN's LSTM weights from the natural fp32 parameters.
struct LstmPackOut {
    __nv_bfloat16* wih_hi;  // [1024,64] packed row order
    __nv_bfloat16* wih_lo;
    float* bias;  // [1024] = b_ih + b_hh, packed order
    uint4* whh_f_hi;
    uint4* whh_f_lo;
    uint4* whh_b_hi;
    uint4* whh_b_lo;
    __nv_bfloat16* wiht_hi;   // [64,1024] = packed W_ih transposed (weights of the input-gradient GEMM, K contiguous)
    __nv_bfloat16* wiht_lo;
    __nv_bfloat16* projt_hi;  // [256,64] = proj.weight transposed (optional, needs proj_w)
    __nv_bfloat16* projt_lo;
    void* rec5;               // lstm_rec5_pack_bytes() bytes (optional)
};
cudaError_t launch_pack_lstm(const float* const w_ih[2], const float* const w_hh[2], const float* const b_ih[2],
                             const float* const b_hh[2], const float* proj_w, const LstmPackOut& o, cudaStream_t st);
// grads of the packed forms -> natural parameter gradients (+=)
cudaError_t launch_unpack_lstm_grads(const float* d_wih_pack /*[1024,64]*/, const float* d_whh_pack /*[2][512,128]*/,
                                     const float* d_bias_pack /*[1024]*/, float* const d_w_ih[2], float* const d_w_hh[2],
                                     float* const d_b_ih[2], float* const d_b_hh[2], cudaStream_t st);
// flat fp32 -> bf16 hi/lo (elementwise)
// fp32 [R, C] -> bf16 hi / lo planes of the transpose [C, R]
cudaError_t launch_transpose_split(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, int R, int C, cudaStream_t st);
cudaError_t launch_split_bf16(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long n, cudaStream_t st);

// ---------------- segmentation / overlap-add (seg_ola.cu) ----------------
cudaError_t launch_segment_nchw(const float* x, float* y, int rows, int L, int K, cudaStream_t st);
cudaError_t launch_overlap_add_nchw(const float* y, float* x, int rows, int K, int S, int L, cudaStream_t st);
// channels-last: F[B,L,C] <-> X[B,S,K,C]   (C % 4 == 0)
cudaError_t launch_segment_cl(const float* f, float* x, int B, int L, int K, int S, int C, cudaStream_t st);
cudaError_t launch_overlap_add_cl(const float* x, float* f, int B, int L, int K, int S, int C, cudaStream_t st);

// ---------------- norms / elementwise (elementwise.cu) ----------------
// (sum, sumsq) fp64 per group -> (mean, rstd) fp32 per group; cnt = elements per group
cudaError_t launch_gn_finalize(const double* stats, float* mr, int groups, double cnt, double eps, cudaStream_t st);
// out = res + (y - mean_g) * rstd_g * gamma + beta   (res may be null); group = row / rows_per_group; C channels.
// unfold (cw != null): out = prelu(cw * out + cb, slope)
// out_hi / out_lo (optional): the result also as bf16 hi/lo operand planes for the GEMM that consumes it
cudaError_t launch_gn_apply(const float* y, const float* res, float* out, const float* mr, const float* gamma, const float* beta,
                            long long rows, int rows_per_group, int C, const float* cw, const float* cb, const float* slope,
                            cudaStream_t st, __nv_bfloat16* out_hi = nullptr, __nv_bfloat16* out_lo = nullptr);
// reductions for the GroupNorm backward: red[g] += (sum gamma*d, sum gamma*d*xhat); dgamma += sum d*xhat; dbeta += sum d
cudaError_t launch_gn_bwd_reduce(const float* d, const float* y, const float* mr, const float* gamma, long long rows, int rows_per_group,
                                 int C, double* red, float* dgamma, float* dbeta, cudaStream_t st);
// dy = rstd * (gamma*d - s1/cnt - xhat*s2/cnt)
cudaError_t launch_gn_bwd_apply(const float* d, const float* y, float* dy, const float* mr, const double* red, const float* gamma,
                                long long rows, int rows_per_group, int C, cudaStream_t st);
// unfold backward: d holds the gradient of prelu(cw*s+cb) with s = res + GN(y) (recomputed); on exit d holds d_s;
// accumulates the concat_block parameter gradients
cudaError_t launch_concat_bwd(float* d, const float* y, const float* res, const float* mr, const float* gamma, const float* beta,
                              long long rows, int rows_per_group, int C, const float* cw, const float* cb, const float* slope, float* dcw,
                              float* dcb, float* dslope, cudaStream_t st);
cudaError_t launch_pad_rows(const float* x, float* xp, int rows, int T, int Tp, int front, cudaStream_t st);
// Mx[b,c,t,n] = Mk[b,t,c*C+n] * E[b,t,n]
cudaError_t launch_mask_apply(const float* Mk, const float* E, float* Mx, int B, int L, int nspk, int C, cudaStream_t st);
// dMk = dMx * E * (Mk > 0) ; dE (=|+=) sum_c dMx * Mk
cudaError_t launch_mask_bwd(const float* dMx, const float* Mk, const float* E, float* dMk, float* dE, int accumulate_dE, int B,
                            int L, int nspk, int C, cudaStream_t st);
// out[r, tau] = D[r, t1, j1] + D[r, t2, j2]  (stride = win/2 overlap-add of decoder frames + trim)
cudaError_t launch_dec_ola(const float* D, float* out, int rows, int L, int win, int T, cudaStream_t st);
cudaError_t launch_axpy(float* y, const float* x, float a, long long n, cudaStream_t st);
// SepFormer helpers (sepformer.py): out[p,:] = x[p,:] + pe[t(p),:] with t = k (intra) or s (inter) of position p = (b*S+s)*K+k
cudaError_t launch_add_pe(const float* x, const float* pe, float* out, long long rows, int E, int K, int S, int inter, cudaStream_t st);
// stats[g] += (sum, sumsq) over the rows of group g = row / rows_per_group (gLN statistics, normalizations.py:17-47)
cudaError_t launch_group_stats(const float* y, long long rows, int rows_per_group, int C, double* stats, cudaStream_t st);
cudaError_t launch_prelu(const float* x, float* out, long long n, const float* slope, cudaStream_t st);
// decoder overlap-add without front padding; row (b,c) of D goes to output row c*B+b (spk_major) or b*nspk+c; zero beyond the frames
cudaError_t launch_dec_ola_general(const float* D, float* out, int B, int nspk, int L, int win, int T, int spk_major, cudaStream_t st);

// ---------------- transformer blocks (transformer.cu) ----------------
// Self-attention of nn.MultiheadAttention on channels-last rows: QKV [P,3E] = [q|k|v] -> O [P,E]; sequences via SeqMap.
// LSE (optional, [P,heads], log2 domain) is what the backward needs.  Head width E/heads must be 16 or 32.
// O may be null when only the bf16 hi/lo planes (O_hi, O_lo: operands of the out-projection GEMM) are wanted.
cudaError_t launch_attn_fwd(const float* QKV, float* O, float* LSE, int E, int heads, const SeqMap& m, cudaStream_t st,
                            __nv_bfloat16* O_hi = nullptr, __nv_bfloat16* O_lo = nullptr, unsigned drop_thr = 0, unsigned drop_key = 0,
                            float drop_scale = 1.f);
cudaError_t launch_attn_bwd(const float* QKV, const float* O, const float* LSE, const float* dO, float* dQKV, int E, int heads,
                            const SeqMap& m, cudaStream_t st);
// The same attention on tcgen05 (attention_tc5.cu): QKV given as bf16 hi/lo planes [P,3E] (what the QKV GEMM writes), TMA-fed;
// S and P live in tensor memory.  Sequences up to 256 long; geometry as for the fused LSTM kernel (gm.nseq sequences).
bool attn_tc5_supported(int E, int heads, const LstmFusedGeom& gm);
// attention backward on the warp-level tensor cores (attention_bwd_mma.cu): sequences <= 256, head width 16 / 32;
// split = bf16x3 products (fp32-parity mode), otherwise single bf16 products.  Same arguments as launch_attn_bwd.
bool attn_bwd_mma_supported(int E, int heads, const SeqMap& m);
// which forward kernel the engines use where both apply (0 automatic from measurements, 1 tcgen05, 2 warp-level)
int attn_fwd_set_mode(int mode);
bool attn_fwd_prefers_mma(int E, int heads, int len, bool split);
// forward on the same machinery (online softmax): O fp32 and / or planes, log-sum-exp; same shape limits as the backward
cudaError_t launch_attn_fwd_mma(const float* QKV, float* O, __nv_bfloat16* O_hi, __nv_bfloat16* O_lo, float* LSE, int E, int heads,
                                const SeqMap& m, bool split, cudaStream_t st, unsigned drop_thr = 0, unsigned drop_key = 0,
                                float drop_scale = 1.f);
// drop_thr != 0: attention-probability dropout, mask element (position of the query * heads + head, key index)
cudaError_t launch_attn_bwd_mma(const float* QKV, const float* O, const float* LSE, const float* dO, float* dQKV, int E, int heads,
                                const SeqMap& m, bool split, cudaStream_t st, unsigned drop_thr = 0, unsigned drop_key = 0,
                                float drop_scale = 1.f);
cudaError_t launch_attn_fwd_tc5(const __nv_bfloat16* qkv_hi, const __nv_bfloat16* qkv_lo, float* O, __nv_bfloat16* O_hi, __nv_bfloat16* O_lo,
                                float* LSE, int E, int heads, const LstmFusedGeom& gm, bool split, cudaStream_t st, unsigned drop_thr = 0,
                                unsigned drop_key = 0, float drop_scale = 1.f);
// z = a (+ b) [-> zout]; out = (res ? res : 0) + LayerNorm_E(z) * gamma + beta; then optional unfold affine + PReLU.
cudaError_t launch_add_ln(const float* a, const float* b, float* zout, float* out, const float* res, const float* gamma, const float* beta,
                          long long rows, int E, float eps, const float* cw, const float* cb, const float* slope, cudaStream_t st,
                          __nv_bfloat16* out_hi = nullptr, __nv_bfloat16* out_lo = nullptr);  // out may be null with planes
// LayerNorm backward from the saved pre-norm rows z: dz (may alias dy), acc (optional) += dz, dgamma/dbeta accumulated.
cudaError_t launch_ln_bwd(const float* dy, const float* z, float* dz, float* acc, const float* gamma, long long rows, int E, float eps,
                          float* dgamma, float* dbeta, cudaStream_t st);
// out[n] += scale * sum_p A[p*lda + n], any N % 4 == 0
cudaError_t launch_colsum_any(const float* A, long long lda, int P, int N, float scale, float* out, cudaStream_t st);

// out = prelu(cw * s + cb) (unfold concat_block applied to a stored sum; launch_concat_bwd with mr == nullptr is its backward,
// y = s, res / gamma / beta unused)
cudaError_t launch_affine_prelu(const float* s, float* out, long long rows, int C, const float* cw, const float* cb, const float* slope,
                                __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, cudaStream_t st);
// SepFormer training helpers (elementwise.cu)
cudaError_t launch_gn_bwd_reduce_any(const float* d, const float* y, const float* mr, const float* gamma, long long rows, int rows_per_group,
                                     int C, double* red, float* dgamma, float* dbeta, cudaStream_t st);
cudaError_t launch_dec_ola_general_bwd(const float* d_est, float* dD, int B, int nspk, int L, int win, int T, int spk_major, cudaStream_t st);
cudaError_t launch_mul(const float* a, const float* b, float* out, long long n, cudaStream_t st);
cudaError_t launch_gate_bwd(const float* dg, const float* t1, const float* t2, float* da, float* db, long long n, cudaStream_t st);
cudaError_t launch_prelu_bwd(const float* du, const float* x, float* dx, long long n, const float* slope, float* dslope, cudaStream_t st);
cudaError_t launch_relu_bwd_add(const float* a, const float* b, const float* m, float* out, long long n, cudaStream_t st);

// ---------------- BSS-eval SDR metric (bss_sdr.cu) ----------------
size_t bss_sdr_workspace_bytes(int B, int n, int L);
// mean SDR (dB) per utterance under the best permutation, 512-tap distortion filter; sdr_mat [B,n,n] (ref, est) optional
cudaError_t launch_bss_sdr_pit(const float* est, const float* ref, int B, int n, int T, int L, void* ws, float* out, float* sdr_mat, cudaStream_t st);

// ---------------- loss (loss.cu: n_src = 2; loss_n.cu: n_src = 1 .. 4) ----------------
struct PitLossWs {  // device scratch, all double unless noted
    double* sums;    // [B][4]   sum e0,e1,t0,t1
    double* second;  // [B][10]  dot[2][2], dist2[2][2], tt[2]
    double* noise;   // [B][4]   sisdr noise energy [est][tgt]
};
// sdr_type: 0 snr, 1 sisdr, 2 sdsdr.  Outputs: pw [B,2,2] (est,tgt), loss (1), perm [B] (0: identity, 1: swapped),
// coef [B][2][3] backward coefficients (per estimate: a on (e - mean e), b on (t_j - mean t_j), j index as float)
cudaError_t launch_pit_loss_fwd(const float* est, const float* tgt, int B, int T, int sdr_type, int threshold_byloss,
                                const PitLossWs& ws, float* pw, float* loss, int* perm, float* coef, cudaStream_t st);
// General n_src = N (1..4): sums [B][2N], second [B][2N*N + N], noise [B][N*N]; pw [B,N,N]; perm [B,N] = estimate index per target
// (find_best_perm_factorial's batch_indices); coef [B][N][3]
cudaError_t launch_pitn_loss_fwd(const float* est, const float* tgt, int B, int N, int T, int sdr_type, int threshold_byloss, const PitLossWs& ws,
                                 float* pw, float* loss, int* perm, float* coef, cudaStream_t st);
cudaError_t launch_pitn_loss_bwd(const float* est, const float* tgt, int B, int N, int T, const double* sums, const float* coef, float grad_scale,
                                 float* d_est, cudaStream_t st);
cudaError_t launch_reorder_sources_n(const float* est, const int* perm, float* out, int B, int N, int T, cudaStream_t st);
cudaError_t launch_pit_loss_bwd(const float* est, const float* tgt, int B, int T, const double* sums, const float* coef,
                                float grad_scale, float* d_est, cudaStream_t st);
cudaError_t launch_reorder_sources(const float* est, const int* perm, float* out, int B, int T, cudaStream_t st);

// ---------------- optimizer (optim.cu) ----------------
cudaError_t launch_set_hyper(float* hyper, float lr, float bc1, float bc2, cudaStream_t st);
cudaError_t launch_sumsq(const float* g, long long n, double* out /*+=*/, cudaStream_t st);
// torch clip_grad_norm_ + Adam, fused.  norm2 holds sum of squares of the *unscaled* grads; gscale is applied first
// (1/world_size after an all-reduce SUM).
cudaError_t launch_adam_clip(float* p, const float* g, float* m, float* v, long long n, const double* norm2, float gscale,
                             float max_norm, float lr, float b1, float b2, float eps, float bc1, float bc2, float wd,
                             cudaStream_t st, const float* hyper = nullptr);

}  // namespace dp
