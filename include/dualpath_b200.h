/* dualpath_b200 -- C-ABI of the B200-native look2hear dual-path separation hot path.
 *
 * The reference (spkgyk/audio-only-speech-separation) is pure Python/PyTorch and has no FFI of its own; the seam a
 * maintainer binds against is "the call each torch.nn module on the path makes".  Every entry point below names
 * the reference code it replaces (paths relative to the reference repo).  All pointers are DEVICE pointers
 * (fp32 unless said otherwise), `stream` is a cudaStream_t passed as void*, every call is asynchronous on that
 * stream, allocates nothing, and returns 0 on success or a non-zero code with a message in dp_last_error().
 * There is no CPU implementation behind these symbols: without a CUDA device they fail.
 *
 * precision: DP_PREC_FP32 -> bf16x3 split products on the tensor cores + precise gate activations
 *                            (model output within rel-L2 1e-4 of the fp32 reference; measured ~1.5e-5)
 *            DP_PREC_BF16 -> single bf16 products + tanh.approx activations (within 0.05 dB SI-SNR)
 */
#ifndef DUALPATH_B200_H
#define DUALPATH_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DP_PREC_FP32 0
#define DP_PREC_BF16 1

int dp_version(void);
const char* dp_last_error(void);
/* GEMM backend of the engines: 2 (default) = TMA-fed tcgen05/TMEM kernels on bf16 hi/lo operand planes written by the
 * producers; 1 = first-generation tcgen05 kernels (threads convert fp32 operands); 0 = warp-level mma.sync kernels
 * everywhere.  All are covered by the parity tests. */
int dp_set_gemm_backend(int backend);
/* backend 2 only: the LSTM input projection and recurrence as ONE tcgen05 kernel (weights resident in tensor memory / shared
 * memory, gate pre-activations never in HBM).  mode 1 (default) = automatic (used in bf16 mode, where it is faster than the TMA
 * GEMM + register-stationary mma.sync recurrence), 2 = always, 0 = never. */
int dp_set_fused_lstm(int mode);
/* Forward recurrence kernel of the register-stationary path: 0 = plain 8-warp kernels; 2 = software-pipelined (two sequence
 * groups per CTA half a step apart: one group's tensor-core products are interleaved with the other group's cell update);
 * 3 = 16-warp kernel (8 hidden units per warp, 128 registers, four warps per scheduler); 1 (default) = automatic by pass size and
 * precision.  All variants perform the same arithmetic in the same order per cell. */
int dp_set_lstm_pipeline(int mode);
/* LSTM recurrence (forward and BPTT, gc3_basics.py:16,22) on tcgen05 / tensor memory (csrc/lstm_rec5.cu: W_hh hi half resident in
 * tensor memory, lo half in shared memory, 32 sequences per CTA): 1 (default) = automatic (passes with >= 256 sequences per direction),
 * 2 = always, 0 = never (the register-stationary mma.sync kernels selected by dp_set_lstm_pipeline). */
int dp_set_lstm_tcgen05(int mode);
/* Forward LSTM recurrence of small inference passes (B = 1: ~80-100 sequences per direction) with the gate rows of a sequence tile split
 * over a cluster of four CTAs that exchange h_t through distributed shared memory (csrc/lstm.cu, lstm_fwdc_kernel; same arithmetic and
 * order per cell as the 16-warp kernel): 1 (default) = automatic (inference passes of <= 128 sequences), 2 = every inference pass, 0 = off. */
int dp_set_lstm_cluster(int mode);
/* Weight-gradient GEMM (autograd's dW = dY^T X) as 4-CTA clusters that multicast the shared operand tiles (outputs with a multiple of four
 * 128-row slices): 1 on, 0 (default) off -- measured equal on B200 (the L2 already shares the four unicast requests).  Returns the previous setting. */
int dp_set_wgrad_multicast(int on);
/* Attention forward kernel of the DPTNet / SepFormer engines where both apply (sequences <= 256): 1 = tcgen05 kernel (TMA-fed, scores in
 * tensor memory), 2 = warp-level tensor-core kernel (online softmax, fp32 QKV in), 0 (default) = the faster one per shape and precision
 * as measured (tests/tools/time_attention.py). */
int dp_set_attention_forward(int mode);

/* ---- geometry (integer index maps) --------------------------------------------------------------------- */
/* gc3_basics.py:63-76 pad_segment: rest and chunk count S for L frames and chunk size K (K even). */
int dp_seg_geometry(int L, int K, int* rest, int* S);
/* gc3_network.py:108-131 pad_input: samples appended (rest) and number of encoder frames for T samples. */
int dp_wave_geometry(int T, int win, int* rest, int* frames);

/* ---- (a) segmentation / overlap-add ---------------------------------------------------------------------- */
/* split_feature, gc3_basics.py:79-91 (== sepformer.py:788-814): x[B,N,L] -> y[B,N,K,S], bit-exact. */
int dp_segment_f32(const float* x, float* y, int B, int N, int L, int K, void* stream);
/* merge_feature, gc3_basics.py:94-109 (== sepformer.py:816-846): y[B,N,K,S] -> x[B,N,L], bit-exact. */
int dp_overlap_add_f32(const float* y, float* x, int B, int N, int K, int S, int L, void* stream);
/* the same maps on channels-last tensors f[B,L,C] <-> x[B,S,K,C] (layout used inside the engine), C % 4 == 0 */
int dp_segment_cl_f32(const float* f, float* x, int B, int L, int K, int C, void* stream);
int dp_overlap_add_cl_f32(const float* x, float* f, int B, int L, int K, int C, void* stream);

/* ---- (b) dense contractions: nn.Linear / 1x1 nn.Conv1d / nn.Conv2d / LSTM input projection -------------- */
/* C[M,N] (=|+=) A[M,K] W^T (+ bias_scale*bias) (ReLU).  gc3_basics.py:22-23, gc3_network.py:55,99, dprnn.py:85.
 * W is given as bf16 hi / lo halves (w_lo unused for DP_PREC_BF16); w_kn = 0: W is [N,K], 1: W is [K,N].
 * stats (optional, fp64 [groups][2]) accumulates (sum, sumsq) of the stored C per group of rows_per_group rows:
 * the GroupNorm(1,C) statistics of dprnn.py:71,80 come out of the producing GEMM. */
int dp_linear_f32(const float* A, int64_t lda, const void* w_hi, const void* w_lo, int ldw, int w_kn, const float* bias,
                  float bias_scale, float* C, int ldc, int M, int N, int K, int relu, int accumulate, double* stats,
                  int rows_per_group, int precision, void* stream);
/* The same contraction on operands already stored as bf16 hi/lo planes (what the engine's producers write): TMA-fed
 * tcgen05/TMEM kernel, fp32 and/or plane output.  act: 0 none, 1 ReLU, 2 tanh, 3 sigmoid.  K % 64 == 0, N % 64 == 0. */
int dp_linear_planes_f32(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo, int ldw, const float* bias,
                         float bias_scale, float* C, int ldc, void* c_hi, void* c_lo, int ldch, int M, int N, int K, int act,
                         int accumulate, int precision, void* stream);
/* Weight gradients on planes, one pass over A for up to two outputs: C0[Mo,nb0] += scale * A^T B0, C1[Mo,nb1] += scale * A^T B1
 * (A [P,>=Mo], B [P,>=nb]: the position is the slow index; Mo % 128 == 0, nb % 64 == 0, nb0 + nb1 <= 256; tr != 0 stores C[col][row]).
 * This is how dW_ih = dG^T x and dW_hh = dG^T h_prev of nn.LSTM come out of a single read of dG. */
int dp_linear_wgrad_planes_f32(const void* a_hi, const void* a_lo, int64_t lda, int Mo, const void* b0_hi, const void* b0_lo, int64_t ldb0,
                               int nb0, const void* b1_hi, const void* b1_lo, int64_t ldb1, int nb1, float* C0, int ldc0, int tr0, float* C1,
                               int ldc1, int tr1, int P, float scale, int precision, void* stream);
/* fp32 rows -> bf16 hi/lo planes (lo may be NULL); relu != 0 applies max(x, 0) first */
int dp_split_rows_f32(const float* src, int64_t ld, void* hi, void* lo, int64_t rows, int C, int relu, void* stream);
/* dW[Mo,No] += scale * A[P,Mo]^T B[P,No]  (weight gradients; fp32 atomics). */
int dp_linear_wgrad_f32(const float* A, int lda, const float* B, int64_t ldb, float* dW, int ldc, int P, int Mo, int No,
                        float scale, int precision, void* stream);
/* elementwise fp32 -> bf16 hi / lo */
int dp_split_bf16(const float* src, void* hi, void* lo, int64_t n, void* stream);

/* ---- (c) persistent BiLSTM recurrence: nn.LSTM(64,128,1,bidirectional) of ProjRNN, gc3_basics.py:16,22 -- */
int64_t dp_lstm_pack_bytes(void);
/* natural nn.LSTM parameters (weight_ih_l0[512,64], weight_hh_l0[512,128], bias_*[512], and *_reverse) ->
 * packed forms consumed by the kernels (see csrc/lstm.cu). */
int dp_lstm_pack(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, const float* w_ih_r,
                 const float* w_hh_r, const float* b_ih_r, const float* b_hh_r, void* pack, void* stream);
/* x[P,64] rows at position p; sequence q, time t lives at row (q/qdiv)*s_hi + (q%qdiv)*s_lo + t*s_t.
 * Computes gates = x W_ih^T + b (into G[P,1024], scratch) and the recurrence; H[P,256] = [h_fwd | h_bwd].
 * save != 0 keeps activated gates in G and cell states in Cst[P,256] for dp_bilstm_backward_f32. */
int dp_bilstm_forward_f32(const void* pack, const float* x, float* G, float* H, float* Cst, int64_t P, int nseq, int len,
                          int qdiv, int64_t s_hi, int64_t s_lo, int64_t s_t, int save, int precision, void* stream);
/* the recurrence alone on precomputed gate pre-activations G (what dp_bilstm_forward_f32 runs after its GEMM) */
int dp_lstm_recurrence_f32(const void* pack, float* G, float* H, float* Cst, int nseq, int len, int qdiv, int64_t s_hi,
                           int64_t s_lo, int64_t s_t, int save, int precision, void* stream);
/* dp_lstm_recurrence_f32 with the operand planes the engines' GEMMs consume (any of them may be NULL, and so may H): h_hi / h_lo =
 * bf16 hi / lo of h_t at every position [P,256]; hp_hi / hp_lo = "h_prev": the previous step's h of the same sequence at every
 * position (zeros at a sequence's first step), the B operand of the dW_hh gradient GEMM. */
int dp_lstm_recurrence_planes_f32(const void* pack, float* G, float* H, float* Cst, void* h_hi, void* h_lo, void* hp_hi, void* hp_lo, int nseq,
                                  int len, int qdiv, int64_t s_hi, int64_t s_lo, int64_t s_t, int save, int precision, void* stream);
/* dH[P,256] -> G becomes d(pre-activations) [P,1024] (packed column order); dx[P,64] (=|+=) dG W_ih; dbias (optional,
 * [1024] packed order) += column sums of dG (= d b_ih = d b_hh). Weight grads via dp_linear_wgrad_f32 on G. */
int dp_bilstm_backward_f32(const void* pack, float* G, const float* Cst, const float* dH, float* dx, int accumulate_dx,
                           float* dbias, int64_t P, int nseq, int len, int qdiv, int64_t s_hi, int64_t s_lo, int64_t s_t,
                           int precision, void* stream);

/* ---- (d) fused GroupNorm(1,C) + residual (+ unfold depthwise affine + PReLU), dprnn.py:71-73,80-82,31-34 - */
int dp_groupnorm_finalize(const double* stats, float* mean_rstd, int groups, double count, double eps, void* stream);
int dp_groupnorm_residual_f32(const float* y, const float* res, float* out, const float* mean_rstd, const float* gamma,
                              const float* beta, int64_t rows, int rows_per_group, int C, const float* cw, const float* cb,
                              const float* prelu_slope, void* stream);

/* ---- (b') transformer blocks: nn.MultiheadAttention core + nn.LayerNorm, dptnet.py:48-82, sepformer.py:124-215,316-370 - */
/* Self-attention core on channels-last rows.  qkv[P,3E] = [q | k | v] (the in_proj output), head h uses columns
 * [h*d,(h+1)*d) of each third, d = E/heads in {16, 32}; softmax(q k^T / sqrt(d)) v -> o[P,E].  Sequence q, time t lives
 * at row (q/qdiv)*s_hi + (q%qdiv)*s_lo + t*s_t (same map as the LSTM entry points), so neither the reference's
 * permute().contiguous() copies nor its [B*S*h, L, L] probability tensor exist.  lse (optional, [P,heads]) is saved
 * for dp_attention_backward_f32. */
int dp_attention_forward_f32(const float* qkv, float* o, float* lse, int E, int heads, int nseq, int len, int qdiv, int64_t s_hi,
                             int64_t s_lo, int64_t s_t, void* stream);
int dp_attention_backward_f32(const float* qkv, const float* o, const float* lse, const float* d_o, float* d_qkv, int E, int heads,
                              int nseq, int len, int qdiv, int64_t s_hi, int64_t s_lo, int64_t s_t, void* stream);
/* dp_attention_forward_f32 / dp_attention_backward_f32 on the warp-level tensor cores (forward: online softmax over 64-key blocks;
 * o fp32 and / or o_hi / o_lo planes, lse optional).  Backward: (probabilities recomputed from lse, never stored; bf16x3 products in
 * fp32 mode, single bf16 products in bf16 mode).  Sequence length <= 320.  What the DPTNet / SepFormer engines use (TMA backend). */
int dp_attention_forward_tc_f32(const float* qkv, float* o, void* o_hi, void* o_lo, float* lse, int E, int heads, int nseq, int len, int qdiv,
                                int64_t s_hi, int64_t s_lo, int64_t s_t, int precision, void* stream);
int dp_attention_backward_tc_f32(const float* qkv, const float* o, const float* lse, const float* d_o, float* d_qkv, int E, int heads,
                                 int nseq, int len, int qdiv, int64_t s_hi, int64_t s_lo, int64_t s_t, int precision, void* stream);
/* The same attention on the 5th-generation tensor cores: qkv given as bf16 hi/lo planes [P,3E] on a dual-path stream [B,S,K,.]
 * (inter = 0: sequences (b,s) along k; 1: sequences (b,k) along s), TMA-fed, scores and probabilities in tensor memory.
 * Outputs: o fp32 and/or o_hi/o_lo planes, lse (all optional).  Sequence length <= 256. */
int dp_attention_forward_planes_f32(const void* qkv_hi, const void* qkv_lo, float* o, void* o_hi, void* o_lo, float* lse, int E, int heads,
                                    int inter, int B, int S, int K, int precision, void* stream);
/* z = a (+ b) (stored to z_out when non-null); out = (res ? res : 0) + LayerNorm_E(z) * gamma + beta.  E in {64,128,256}. */
int dp_add_layernorm_f32(const float* a, const float* b, float* z_out, float* out, const float* res, const float* gamma,
                         const float* beta, int64_t rows, int E, float eps, void* stream);
/* LayerNorm backward from the saved pre-norm rows z: dz (may alias dy); acc (optional) += dz; dgamma/dbeta accumulated. */
int dp_layernorm_backward_f32(const float* dy, const float* z, float* dz, float* acc, const float* gamma, int64_t rows, int E,
                              float eps, float* dgamma, float* dbeta, void* stream);

/* ---- (e) fused pairwise SNR / SI-SDR + PIT, losses/matrix.py:13-57, losses/pit_wrapper.py:30-131 ---------- */
int64_t dp_pit_loss_workspace_bytes(int B);
/* sdr_type 0 snr, 1 sisdr, 2 sdsdr; n_src = 2.  Outputs: pw[B,2,2] (est,tgt), loss[1], perm[B] (0 identity, 1
 * swapped).  ws keeps what the backward needs. */
int dp_pit_loss_forward(const float* est, const float* tgt, int B, int T, int sdr_type, int threshold_byloss, void* ws,
                        float* pw, float* loss, int32_t* perm, void* stream);
int dp_pit_loss_backward(const float* est, const float* tgt, int B, int T, const void* ws, float grad_scale, float* d_est,
                         void* stream);
/* pit_wrapper.py:90-94 reordered_sources */
int dp_pit_reorder(const float* est, const int32_t* perm, float* out, int B, int T, void* stream);

/* The same for n_src = N in 1 .. 4 and every pit_from of pit_wrapper.py:30-88 (pw_mtx / pw_pt build this pair matrix, perm_avg averages
 * its entries per permutation: matrix.py:22-57,75-106,119-152); permutations in itertools order, first minimum wins
 * (find_best_perm_factorial, pit_wrapper.py:106-131).  perm[B,N] = estimate index per target (batch_indices). */
int64_t dp_pitn_loss_workspace_bytes(int B, int N);
int dp_pitn_loss_forward(const float* est, const float* tgt, int B, int N, int T, int sdr_type, int threshold_byloss, void* ws,
                         float* pw, float* loss, int32_t* perm, void* stream);
int dp_pitn_loss_backward(const float* est, const float* tgt, int B, int N, int T, const void* ws, float grad_scale, float* d_est,
                          void* stream);
int dp_pitn_reorder(const float* est, const int32_t* perm, float* out, int B, int N, int T, void* stream);

/* ---- SDR metric of the evaluation loop: -fast_bss_eval.sdr_pit_loss(est, ref).mean() per utterance (metrics/wrapper.py:38-41; the
 * package is a third-party dependency, its published algorithm is restated: unit-norm rows, 512-lag correlations, Toeplitz solve,
 * coherence -> dB, best permutation).  mean_sdr[B]; sdr_mat[B,n,n] (reference, estimate) optional. */
int64_t dp_bss_sdr_workspace_bytes(int B, int n_src, int filter_len);
int dp_bss_sdr_pit(const float* est, const float* ref, int B, int n_src, int T, int filter_len, void* ws, float* mean_sdr, float* sdr_mat,
                   void* stream);

/* ---- optimizer step: clip_grad_norm_(max_norm) + Adam, audio_train.py:48,128 ------------------------------ */
/* norm2: device fp64 scalar (scratch).  grad_scale is applied to g first (1/world_size after all-reduce SUM). */
int dp_adam_clip_step(float* p, const float* g, float* m, float* v, int64_t n, double* norm2, float grad_scale,
                      float max_norm, float lr, float beta1, float beta2, float eps, int step, float weight_decay,
                      void* stream);
/* The same step with (lr, 1 - beta1^t, 1 - beta2^t) read from three floats of device memory by the kernel: a training step captured in a CUDA
 * graph replays with the current learning rate and step number: dp_adam_set_hyper writes the three floats on the stream before each replay. */
int dp_adam_set_hyper(float* hyper, float lr, float beta1, float beta2, int step, void* stream);
int dp_adam_clip_step_dev(float* p, const float* g, float* m, float* v, int64_t n, double* norm2, float grad_scale, float max_norm,
                          const float* hyper, float beta1, float beta2, float eps, float weight_decay, void* stream);

/* ---- whole-model engine: TasNet(module="DPRNN").forward, gc3_network.py:133-184 --------------------------- */
typedef struct dp_tasnet dp_tasnet;
#define DP_MODULE_DPRNN 0  /* look2hear/models/utils/dprnn.py */
#define DP_MODULE_DPTNET 1 /* look2hear/models/utils/dptnet.py */
typedef struct {
    int enc_dim, bn_dim, hidden_dim, win, layer, num_spk, block_size, unfold;
    int module; /* DP_MODULE_* : TasNet(module=...) of gc3_network.py:8-22 */
} dp_tasnet_config;

/* Parameter table: element offsets into one flat fp32 parameter buffer, in this order:
 *   0 encoder.weight  1 bottleneck.0.weight  2 bottleneck.0.bias  3 bottleneck.1.weight
 *   4 seq.output.weight  5 seq.output.bias  6 mask.0.weight  7 mask.0.bias  8 decoder.weight
 *   9 concat_block.0.weight  10 concat_block.0.bias  11 concat_block.1.weight   (-1 unless unfold)
 *   then for path pp = 2*layer + (0 row | 1 col), 12 entries at 12 + 12*pp:
 *   weight_ih, weight_hh, bias_ih, bias_hh, weight_ih_reverse, weight_hh_reverse, bias_ih_reverse,
 *   bias_hh_reverse, proj.weight, proj.bias, norm.weight, norm.bias
 * Shared (unfold) parameters simply repeat the same offsets.
 * DP_MODULE_DPTNET (dptnet.py:45-56): 18 entries per path at 12 + 18*pp, the transformer layer of row_xfmr/col_xfmr:
 *   linear1.{weight_ih, weight_hh, bias_ih, bias_hh, and _reverse} (the BiLSTM "feed-forward"), linear2.weight,
 *   linear2.bias, norm2.weight, norm2.bias, self_attn.in_proj_weight, self_attn.in_proj_bias,
 *   self_attn.out_proj.weight, self_attn.out_proj.bias, norm1.weight, norm1.bias */
#define DP_TASNET_HEAD_PARAMS 12
#define DP_TASNET_PATH_PARAMS 12
#define DP_TASNET_PATH_PARAMS_DPTNET 18
int dp_tasnet_create(const dp_tasnet_config* cfg, const int64_t* offsets, int n_offsets, int64_t n_params, dp_tasnet** out);
void dp_tasnet_destroy(dp_tasnet* h);
int64_t dp_tasnet_pack_bytes(const dp_tasnet* h);
int64_t dp_tasnet_workspace_bytes(const dp_tasnet* h, int B, int T, int train);
/* rebuild the packed / split weights after the parameters changed */
int dp_tasnet_pack(dp_tasnet* h, const float* params, void* pack, void* stream);
/* mixture[B,T] -> est[B,num_spk,T].  train != 0 keeps activations in the workspace for dp_tasnet_backward. */
int dp_tasnet_forward(dp_tasnet* h, const float* params, const void* pack, const float* mixture, float* est, void* workspace,
                      int B, int T, int train, int precision, void* stream);
/* d_est[B,num_spk,T] -> grads (flat, same layout as params, ACCUMULATED into). Needs the workspace of the forward. */
int dp_tasnet_backward(dp_tasnet* h, const float* params, const void* pack, const float* d_est, float* grads, void* workspace,
                       int B, int T, int precision, void* stream);
/* number of kernels the last forward / backward call of this handle launched */
int dp_tasnet_last_launches(const dp_tasnet* h);

/* ---- whole-model engine: Sepformer.forward, look2hear/models/sepformer.py:986-1016 (inference) ---------------- */
typedef struct dp_sepformer dp_sepformer;
typedef struct {
    int enc_dim;     /* encoder_out_nchannels (64, 128 or 256)          sepformer.py:906 */
    int win;         /* encoder_kernel_size (stride = win / 2)           :904 */
    int chunk;       /* masknet_chunksize K                              :907 */
    int num_blocks;  /* masknet_numlayers                                :908 */
    int num_spk;     /* masknet_numspks                                  :910 */
    int intra_layers, inter_layers, intra_heads, inter_heads, intra_dffn, inter_dffn;
    int intra_pe, inter_pe;                    /* *_use_positional */
    int intra_norm_before, inter_norm_before;  /* pre-norm (True in configs/sepformer_base.yml) or post-norm layers */
} dp_sepformer_config;

/* Parameter table: element offsets into one flat fp32 buffer (parameters AND the pos_enc.pe buffers), in this order:
 *   0 encoder.conv1d.weight  1 masknet.norm.weight  2 masknet.norm.bias  3 masknet.conv1d.weight  4 masknet.prelu.weight
 *   5 masknet.conv2d.weight  6 masknet.conv2d.bias  7 masknet.output.0.weight  8 masknet.output.0.bias
 *   9 masknet.output_gate.0.weight  10 masknet.output_gate.0.bias  11 masknet.end_conv1x1.weight  12 decoder.weight
 * then for every block j and path (intra_mdl, inter_mdl) a run of 1 + 12*layers + 4 entries:
 *   pos_enc.pe (-1 when positional encoding is off), then per layer: self_att.att.in_proj_weight, in_proj_bias,
 *   out_proj.weight, out_proj.bias, pos_ffn.ffn.0.weight, ffn.0.bias, ffn.3.weight, ffn.3.bias, norm1.weight, norm1.bias,
 *   norm2.weight, norm2.bias; then mdl.norm.weight, mdl.norm.bias, {intra,inter}_norm.gamma, {intra,inter}_norm.beta */
#define DP_SEPFORMER_HEAD_PARAMS 13
#define DP_SEPFORMER_LAYER_PARAMS 12
#define DP_SEPFORMER_PE_LEN 2500 /* PositionalEncoding max_len, sepformer.py:61 */
int dp_sepformer_create(const dp_sepformer_config* cfg, const int64_t* offsets, int n_offsets, int64_t n_params, dp_sepformer** out);
void dp_sepformer_destroy(dp_sepformer* h);
int64_t dp_sepformer_pack_bytes(const dp_sepformer* h);
int64_t dp_sepformer_workspace_bytes(const dp_sepformer* h, int B, int T);
/* bf16 hi/lo split of the flat buffer; call again whenever the parameters changed */
int dp_sepformer_pack(dp_sepformer* h, const float* params, void* pack, void* stream);
/* mixture[B,T] -> est[B,num_spk,T] (rows laid out exactly like the reference's reshape, sepformer.py:1004) */
int dp_sepformer_forward(dp_sepformer* h, const float* params, const void* pack, const float* mixture, float* est, void* workspace,
                         int B, int T, int precision, void* stream);
int dp_sepformer_last_launches(const dp_sepformer* h);
/* Training (pre-norm layers): a forward that keeps in its workspace what the backward needs, and the backward into a flat gradient
 * buffer laid out like params (ACCUMULATED into).  Same pack as the inference engine.
 * dp_sepformer_set_dropout: the transformer layers' training-time dropout (sepformer.py:124-128,261,318-319: attention
 * probabilities, attention output, FFN hidden, FFN output; the reference default is 0.1).  Masks are a counter-based function of
 * (seed, layer, site, element): the backward regenerates them, so call it with the SAME p and seed before the forward and the
 * matching backward.  p = 0 switches the sites off.  Needs the TMA backend. */
int dp_sepformer_set_dropout(dp_sepformer* h, float p, uint32_t seed);
int64_t dp_sepformer_train_workspace_bytes(const dp_sepformer* h, int B, int T);
int dp_sepformer_forward_train(dp_sepformer* h, const float* params, const void* pack, const float* mixture, float* est, void* workspace,
                               int B, int T, int precision, void* stream);
int dp_sepformer_backward(dp_sepformer* h, const float* params, const void* pack, const float* d_est, float* grads, void* workspace,
                          int B, int T, int precision, void* stream);

/* ---- whole-model engine: TasNet.forward with group_size > 1 (GroupComm), module "DPRNN" or "DPTNet" -- look2hear/models/gc3_network.py:133-184
 * with the context encoder / decoder (GC_RNN, utils/groupcomm.py:10-45), TAC (utils/gc3_basics.py:28-60) and the grouped DPRNN stack
 * (utils/dprnn.py:53-88) or DPTNet stack (utils/dptnet.py:133-162).  Inference (forward) only; fp32 CUDA-core arithmetic (the per-group operators are 4..16 wide). ---- */
typedef struct dp_gctasnet dp_gctasnet;
typedef struct {
    int enc_dim, bn_dim, hidden_dim, win, layer, num_spk, context_size, group_size, block_size, unfold; /* gc3_network.py:8-22 */
    int module; /* DP_MODULE_DPRNN or DP_MODULE_DPTNET */
} dp_gctasnet_config;
/* Parameter table (element offsets into one flat fp32 buffer, each a multiple of 4):
 *   0 encoder.weight  1 bottleneck.0.weight  2 bottleneck.0.bias  3 bottleneck.1.weight  4 seq.output.weight  5 seq.output.bias
 *   6 mask.0.weight  7 mask.0.bias  8 decoder.weight  9 concat_block.0.weight  10 concat_block.0.bias  11 concat_block.1.weight
 *   (9..11: -1 unless unfold; with unfold the shared row / col RNNs and norms simply repeat their offsets in every layer)
 *   then context_enc and context_dec (GC_RNN), each 2 layers x 23 entries:
 *     TAC: TAC_input.0.weight, .0.bias, .1.weight (PReLU), TAC_mean.0.weight, .0.bias, .1.weight, TAC_output.0.weight, .0.bias, .1.weight,
 *          TAC_norm.weight, TAC_norm.bias
 *     rnn: weight_ih, weight_hh, bias_ih, bias_hh, the four _reverse, proj.weight, proj.bias, then LN.weight, LN.bias
 *   then per DPRNN layer 35 entries: TAC (11), row_rnn + row_norm (12, as above), col_rnn + col_norm (12);
 *   DP_MODULE_DPTNET: per layer 47 entries: TAC (11), then row_xfmr and col_xfmr, 18 each in the order of DP_TASNET_PATH_PARAMS_DPTNET
 *   above (linear1 BiLSTM 8, linear2.weight/.bias, norm2, self_attn.in_proj_weight/_bias, out_proj.weight/.bias, norm1). */
int dp_gctasnet_n_offsets(int layer, int module);
int dp_gctasnet_create(const dp_gctasnet_config* cfg, const int64_t* offsets, int n_offsets, int64_t n_params, dp_gctasnet** out);
void dp_gctasnet_destroy(dp_gctasnet* h);
int64_t dp_gctasnet_workspace_bytes(const dp_gctasnet* h, int B, int T);
/* mixture[B,T] -> est[B,num_spk,T] */
int dp_gctasnet_forward(dp_gctasnet* h, const float* params, const float* mixture, float* est, void* workspace, int B, int T, void* stream);
int dp_gctasnet_last_launches(const dp_gctasnet* h);
/* Training (module DPRNN): the forward that keeps every stage's input, pre-norm tensor, LSTM output and activated gates in the workspace,
 * and the backward of the whole network (autograd through gc3_network.py:133-184, groupcomm.py:26-45, gc3_basics.py:7-60,
 * dprnn.py:53-88): grads (same offsets as params) is ACCUMULATED into; d_est = gradient of est [B, num_spk, T]. */
int64_t dp_gctasnet_train_workspace_bytes(const dp_gctasnet* h, int B, int T);
int dp_gctasnet_forward_train(dp_gctasnet* h, const float* params, const float* mixture, float* est, void* workspace, int B, int T, void* stream);
int dp_gctasnet_backward(dp_gctasnet* h, const float* params, float* grads, const float* mixture, const float* d_est, void* workspace, int B,
                         int T, void* stream);
/* the narrow recurrence stages a CTA's input sequences in shared memory (default, sequences up to ~800 steps) or prefetches them from
 * global memory one step ahead (longer sequences; 0 forces this path, for cross-checks).  Returns the previous setting. */
int dp_gctasnet_set_lstm_staging(int on);

#ifdef __cplusplus
}
#endif
#endif
