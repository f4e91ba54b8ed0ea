// C-ABI (include/dualpath_b200.h) and the TasNet/DPRNN engine: host-side orchestration of the kernels.
//
// Data layout in HBM (all fp32, "channels-last"):
//   waveform   xp [B, Tp]            zero padded (gc3_network.py:123-129)
//   frames     E, En, Fb, F2, Z [B*L, 64], Mk [B*L, 128], Mx [B*spk*L, 64], D [B*spk*L, 16]
//   stream     X_l [B, S, K, 64]     position p = (b*S + s)*K + k ; intra sequences walk k, inter sequences walk s,
//                                    both read rows of 64 contiguous floats -> none of the reference's
//                                    permute().contiguous() copies (dprnn.py:67,69,76,78) exist here
//   per path   G [P,1024] gates, H [P,256], Cst [P,256], Y [P,64]
// Forward of one path: in-proj GEMM -> persistent BiLSTM -> out-proj GEMM (+GroupNorm statistics in its epilogue)
// -> fused normalise + residual (+ unfold affine/PReLU).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/dualpath_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "engine_common.h"

using namespace dp;

namespace dp {
thread_local char g_err[512] = "";
int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
int cuda_fail(cudaError_t e, const char* what) { return fail("%s: %s", what, cudaGetErrorString(e)); }
}  // namespace dp

namespace {

// ---- packed LSTM buffer layout (bytes) ----
constexpr size_t PK_WIH_HI = 0;
constexpr size_t PK_WIH_LO = PK_WIH_HI + 65536 * 2;
constexpr size_t PK_BIAS = PK_WIH_LO + 65536 * 2;
constexpr size_t PK_WF_HI = PK_BIAS + 1024 * 4;
constexpr size_t PK_WF_LO = PK_WF_HI + 2 * 8192 * 16;
constexpr size_t PK_WB_HI = PK_WF_LO + 2 * 8192 * 16;
constexpr size_t PK_WB_LO = PK_WB_HI + 2 * 8192 * 16;
constexpr size_t PK_WIHT_HI = PK_WB_LO + 2 * 8192 * 16;
constexpr size_t PK_WIHT_LO = PK_WIHT_HI + 65536 * 2;
constexpr size_t PK_PROJT_HI = PK_WIHT_LO + 65536 * 2;
constexpr size_t PK_PROJT_LO = PK_PROJT_HI + 16384 * 2;
constexpr size_t PK_REC5 = PK_PROJT_LO + 16384 * 2;          // tcgen05 recurrence images (lstm_rec5.cu): 1 MB
constexpr size_t PK_BYTES = PK_REC5 + 8 * 131072;

struct LstmPackView {
    const __nv_bfloat16* wih_hi;
    const __nv_bfloat16* wih_lo;
    const float* bias;
    const __nv_bfloat16* wiht_hi;
    const __nv_bfloat16* wiht_lo;
    const __nv_bfloat16* projt_hi;
    const __nv_bfloat16* projt_lo;
    LstmPack rec;
};
LstmPackView view_pack(const void* pack) {
    const char* b = static_cast<const char*>(pack);
    LstmPackView v;
    v.wih_hi = reinterpret_cast<const __nv_bfloat16*>(b + PK_WIH_HI);
    v.wih_lo = reinterpret_cast<const __nv_bfloat16*>(b + PK_WIH_LO);
    v.bias = reinterpret_cast<const float*>(b + PK_BIAS);
    v.rec.whh_f_hi = reinterpret_cast<const uint4*>(b + PK_WF_HI);
    v.rec.whh_f_lo = reinterpret_cast<const uint4*>(b + PK_WF_LO);
    v.rec.whh_b_hi = reinterpret_cast<const uint4*>(b + PK_WB_HI);
    v.rec.whh_b_lo = reinterpret_cast<const uint4*>(b + PK_WB_LO);
    v.wiht_hi = reinterpret_cast<const __nv_bfloat16*>(b + PK_WIHT_HI);
    v.wiht_lo = reinterpret_cast<const __nv_bfloat16*>(b + PK_WIHT_LO);
    v.projt_hi = reinterpret_cast<const __nv_bfloat16*>(b + PK_PROJT_HI);
    v.projt_lo = reinterpret_cast<const __nv_bfloat16*>(b + PK_PROJT_LO);
    v.rec.rec5 = b + PK_REC5;
    return v;
}
LstmPackOut out_pack(void* pack) {
    char* b = static_cast<char*>(pack);
    LstmPackOut o;
    o.wih_hi = reinterpret_cast<__nv_bfloat16*>(b + PK_WIH_HI);
    o.wih_lo = reinterpret_cast<__nv_bfloat16*>(b + PK_WIH_LO);
    o.bias = reinterpret_cast<float*>(b + PK_BIAS);
    o.whh_f_hi = reinterpret_cast<uint4*>(b + PK_WF_HI);
    o.whh_f_lo = reinterpret_cast<uint4*>(b + PK_WF_LO);
    o.whh_b_hi = reinterpret_cast<uint4*>(b + PK_WB_HI);
    o.whh_b_lo = reinterpret_cast<uint4*>(b + PK_WB_LO);
    o.wiht_hi = reinterpret_cast<__nv_bfloat16*>(b + PK_WIHT_HI);
    o.wiht_lo = reinterpret_cast<__nv_bfloat16*>(b + PK_WIHT_LO);
    o.projt_hi = reinterpret_cast<__nv_bfloat16*>(b + PK_PROJT_HI);
    o.projt_lo = reinterpret_cast<__nv_bfloat16*>(b + PK_PROJT_LO);
    o.rec5 = b + PK_REC5;
    return o;
}

int g_backend = 2;  // 0: mma.sync GEMMs everywhere; 1: first-generation tcgen05 kernels (threads convert fp32 operands);
                    // 2 (default): TMA-fed tcgen05 kernels on pre-split bf16 operand planes (gemm_tma.cu) in the engines

// backend 2 only: fused in-projection + recurrence tcgen05 kernel (lstm_tc5.cu) instead of GEMM + mma.sync recurrence.
// 1 = automatic: the fused kernel issues 48 small tcgen05.mma per split product and step (measured ~110 cycles each for
// N <= 32), which wins in bf16 mode but not with the three split products of the fp32-parity mode, where the
// register-stationary mma.sync recurrence is faster.
// 2 = always, 0 = never.
int g_fused_lstm = 1;

}  // namespace

namespace dp {
int gemm_backend() { return g_backend; }
cudaError_t gemm_nt(const GemmNtArgs& a, bool split, cudaStream_t st) {
    if (g_backend == 1 && !a.w_kn && gemm_nt_tc5_supported(a)) return launch_gemm_nt_tc5(a, split, st);
    return launch_gemm_nt(a, split, st);
}
cudaError_t gemm_tn(const GemmTnArgs& a, bool split, cudaStream_t st) {
    if (g_backend == 1 && !a.b_rpb && gemm_tn_tc5_supported(a)) return launch_gemm_tn_tc5(a, split, 0, st);
    return launch_gemm_tn(a, split, st);
}
}  // namespace dp

namespace {

struct PitWsView {
    double* sums;
    double* second;
    double* noise;
    float* coef;
};
PitWsView pit_view(void* ws, int B) {
    PitWsView v;
    double* d = static_cast<double*>(ws);
    v.sums = d; v.second = d + 4 * B; v.noise = d + 14 * B;
    v.coef = reinterpret_cast<float*>(d + 18 * B);
    return v;
}

}  // namespace

// ================================================================================================
extern "C" {

int dp_version(void) { return 101; }
int dp_set_gemm_backend(int backend) {
    if (backend < 0 || backend > 2) return fail("dp_set_gemm_backend: 0 (mma.sync), 1 (tcgen05, converting threads) or 2 (TMA-fed tcgen05 on planes)");
    g_backend = backend;
    return 0;
}
int dp_set_fused_lstm(int mode) {
    if (mode < 0 || mode > 2) return fail("dp_set_fused_lstm: 0 (never), 1 (automatic) or 2 (always)");
    g_fused_lstm = mode;
    return 0;
}
int dp_set_attention_forward(int mode) {
    if (attn_fwd_set_mode(mode) != 0) return fail("dp_set_attention_forward: 0 (automatic), 1 (tcgen05 kernel) or 2 (warp-level tensor-core kernel)");
    return 0;
}
int dp_set_lstm_pipeline(int mode) {
    if (lstm_set_pipeline(mode) != 0) return fail("dp_set_lstm_pipeline: 0 (plain 8-warp kernels), 1 (automatic), 2 (pipelined sequence groups) or 3 (16-warp kernel)");
    return 0;
}
int dp_set_lstm_cluster(int mode) {
    if (lstm_set_cluster(mode) != 0) return fail("dp_set_lstm_cluster: 0 (off), 1 (automatic: inference passes of <= 128 sequences) or 2 (every inference pass)");
    return 0;
}
int dp_set_wgrad_multicast(int on) { return gemm_tma_set_wgrad_multicast(on); }
int dp_set_lstm_tcgen05(int mode) {
    if (lstm_set_rec5(mode) != 0) return fail("dp_set_lstm_tcgen05: 0 (mma.sync recurrence kernels), 1 (automatic) or 2 (tcgen05 recurrence kernels always)");
    return 0;
}
const char* dp_last_error(void) { return g_err; }

int dp_seg_geometry(int L, int K, int* rest, int* Sout) {
    if (L <= 0 || K <= 0 || (K & 1)) return fail("dp_seg_geometry: need L > 0 and even K > 0 (got L=%d K=%d)", L, K);
    int P = K / 2, r = K - (P + L % K) % K;
    if (rest) *rest = r;
    if (Sout) *Sout = 2 * ((L + r + P) / K);
    return 0;
}
int dp_wave_geometry(int T, int win, int* rest, int* frames) {
    if (T <= 0 || win <= 0 || (win & 1)) return fail("dp_wave_geometry: need T > 0 and even win > 0");
    int st = win / 2, r = win - (st + T % win) % win;
    if (rest) *rest = r;
    if (frames) *frames = (T + r + 2 * st - win) / st + 1;
    return 0;
}

int dp_segment_f32(const float* x, float* y, int B, int N, int L, int K, void* stream) {
    if (B < 0 || N < 0) return fail("dp_segment_f32: negative size");
    if (L <= 0 || K <= 0 || (K & 1)) return fail("dp_segment_f32: need L > 0 and even K > 0 (got L=%d K=%d)", L, K);
    CK(launch_segment_nchw(x, y, B * N, L, K, S(stream)));
    return 0;
}
int dp_overlap_add_f32(const float* y, float* x, int B, int N, int K, int Sc, int L, void* stream) {
    int rest, S2;
    if (dp_seg_geometry(L, K, &rest, &S2)) return 1;
    if (S2 != Sc) return fail("dp_overlap_add_f32: S=%d does not match L=%d K=%d (expected %d)", Sc, L, K, S2);
    CK(launch_overlap_add_nchw(y, x, B * N, K, Sc, L, S(stream)));
    return 0;
}
int dp_segment_cl_f32(const float* f, float* x, int B, int L, int K, int C, void* stream) {
    int rest, S2;
    if (dp_seg_geometry(L, K, &rest, &S2)) return 1;
    CK(launch_segment_cl(f, x, B, L, K, S2, C, S(stream)));
    return 0;
}
int dp_overlap_add_cl_f32(const float* x, float* f, int B, int L, int K, int C, void* stream) {
    int rest, S2;
    if (dp_seg_geometry(L, K, &rest, &S2)) return 1;
    CK(launch_overlap_add_cl(x, f, B, L, K, S2, C, S(stream)));
    return 0;
}

int dp_linear_f32(const float* A, int64_t lda, const void* w_hi, const void* w_lo, int ldw, int w_kn, const float* bias, float bias_scale,
                  float* C, int ldc, int M, int N, int K, int relu, int accumulate, double* stats, int rows_per_group, int precision,
                  void* stream) {
    GemmNtArgs a = nt_args(A, lda, (const __nv_bfloat16*)w_hi, (const __nv_bfloat16*)w_lo, ldw, w_kn, C, ldc, M, N, K);
    a.bias = bias; a.bias_scale = bias_scale; a.relu = relu; a.accumulate = accumulate; a.stats = stats;
    a.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
    CK(gemm_nt(a, is_split(precision), S(stream)));
    return 0;
}
int dp_linear_planes_f32(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo, int ldw, const float* bias,
                         float bias_scale, float* C, int ldc, void* c_hi, void* c_lo, int ldch, int M, int N, int K, int act, int accumulate,
                         int precision, void* stream) {
    TmaGemmArgs a;
    memset(&a, 0, sizeof(a));
    a.A_hi = (const __nv_bfloat16*)a_hi; a.A_lo = (const __nv_bfloat16*)a_lo; a.lda = lda;
    a.W_hi = (const __nv_bfloat16*)w_hi; a.W_lo = (const __nv_bfloat16*)w_lo; a.ldw = ldw;
    a.M = M; a.N = N; a.K = K; a.C = C; a.ldc = ldc; a.C_hi = (__nv_bfloat16*)c_hi; a.C_lo = (__nv_bfloat16*)c_lo; a.ldch = ldch;
    a.bias = bias; a.bias_scale = bias_scale; a.act = act; a.accumulate = accumulate;
    if (!gemm_tma_nt_supported(a)) return fail("dp_linear_planes_f32: unsupported shape (need K %% 64 == 0, N %% 64 == 0, aligned strides)");
    CK(launch_gemm_tma_nt(a, is_split(precision), S(stream)));
    return 0;
}
int dp_linear_wgrad_planes_f32(const void* a_hi, const void* a_lo, int64_t lda, int Mo, const void* b0_hi, const void* b0_lo, int64_t ldb0, int nb0,
                               const void* b1_hi, const void* b1_lo, int64_t ldb1, int nb1, float* C0, int ldc0, int tr0, float* C1, int ldc1,
                               int tr1, int P, float scale, int precision, void* stream) {
    TmaWgradArgs a;
    memset(&a, 0, sizeof(a));
    a.A_hi = (const __nv_bfloat16*)a_hi; a.A_lo = (const __nv_bfloat16*)a_lo; a.lda = lda; a.Mo = Mo;
    a.B0_hi = (const __nv_bfloat16*)b0_hi; a.B0_lo = (const __nv_bfloat16*)b0_lo; a.ldb0 = ldb0; a.nb0 = nb0;
    a.B1_hi = (const __nv_bfloat16*)b1_hi; a.B1_lo = (const __nv_bfloat16*)b1_lo; a.ldb1 = ldb1; a.nb1 = nb1;
    a.C0 = C0; a.ldc0 = ldc0; a.transpose0 = tr0; a.C1 = C1; a.ldc1 = ldc1; a.transpose1 = tr1; a.P = P; a.scale = scale;
    if (!gemm_tma_tn_supported(a)) return fail("dp_linear_wgrad_planes_f32: unsupported shape (Mo %% 128, nb %% 64, nb0 + nb1 <= 256, aligned strides)");
    CK(launch_gemm_tma_tn(a, is_split(precision), S(stream)));
    return 0;
}
int dp_split_rows_f32(const float* src, int64_t ld, void* hi, void* lo, int64_t rows, int C, int relu, void* stream) {
    CK(launch_split_rows(src, ld, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, rows, C, relu, S(stream)));
    return 0;
}
int dp_linear_wgrad_f32(const float* A, int lda, const float* B, int64_t ldb, float* dW, int ldc, int P, int Mo, int No, float scale,
                        int precision, void* stream) {
    GemmTnArgs a = tn_args(A, lda, B, ldb, dW, ldc, P, Mo, No);
    a.scale = scale;
    CK(gemm_tn(a, is_split(precision), S(stream)));
    return 0;
}
int dp_split_bf16(const float* src, void* hi, void* lo, int64_t n, void* stream) {
    CK(launch_split_bf16(src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n, S(stream)));
    return 0;
}

int64_t dp_lstm_pack_bytes(void) { return (int64_t)PK_BYTES; }
int dp_lstm_pack(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, const float* w_ih_r, const float* w_hh_r,
                 const float* b_ih_r, const float* b_hh_r, void* pack, void* stream) {
    const float* wi[2] = {w_ih, w_ih_r};
    const float* wh[2] = {w_hh, w_hh_r};
    const float* bi[2] = {b_ih, b_ih_r};
    const float* bh[2] = {b_hh, b_hh_r};
    CK(launch_pack_lstm(wi, wh, bi, bh, nullptr, out_pack(pack), S(stream)));
    return 0;
}

int dp_bilstm_forward_f32(const void* pack, const float* x, float* G, float* H, float* Cst, int64_t P, int nseq, int len, int qdiv,
                          int64_t s_hi, int64_t s_lo, int64_t s_t, int save, int precision, void* stream) {
    if (P > 0x7fffffffLL / 1024) return fail("dp_bilstm_forward_f32: too many positions");
    LstmPackView v = view_pack(pack);
    GemmNtArgs a = nt_args(x, kN, v.wih_hi, v.wih_lo, kN, 0, G, 2 * kG, (int)P, 2 * kG, kN);
    a.bias = v.bias;
    CK(gemm_nt(a, is_split(precision), S(stream)));
    SeqMap m{nseq, len, qdiv, s_hi, s_lo, s_t};
    CK(launch_lstm_fwd(v.rec, G, H, Cst, m, is_split(precision), save != 0, S(stream)));
    return 0;
}
int dp_lstm_recurrence_f32(const void* pack, float* G, float* H, float* Cst, int nseq, int len, int qdiv, int64_t s_hi, int64_t s_lo,
                           int64_t s_t, int save, int precision, void* stream) {
    LstmPackView v = view_pack(pack);
    SeqMap m{nseq, len, qdiv, s_hi, s_lo, s_t};
    CK(launch_lstm_fwd(v.rec, G, H, Cst, m, is_split(precision), save != 0, S(stream)));
    return 0;
}
int dp_lstm_recurrence_planes_f32(const void* pack, float* G, float* H, float* Cst, void* h_hi, void* h_lo, void* hp_hi, void* hp_lo, int nseq,
                                  int len, int qdiv, int64_t s_hi, int64_t s_lo, int64_t s_t, int save, int precision, void* stream) {
    LstmPackView v = view_pack(pack);
    SeqMap m{nseq, len, qdiv, s_hi, s_lo, s_t};
    LstmPlanes pl;
    pl.h_hi = static_cast<__nv_bfloat16*>(h_hi); pl.h_lo = static_cast<__nv_bfloat16*>(h_lo);
    pl.hp_hi = static_cast<__nv_bfloat16*>(hp_hi); pl.hp_lo = static_cast<__nv_bfloat16*>(hp_lo);
    CK(launch_lstm_fwd(v.rec, G, H, Cst, m, is_split(precision), save != 0, S(stream), &pl));
    return 0;
}
int dp_bilstm_backward_f32(const void* pack, float* G, const float* Cst, const float* dH, float* dx, int accumulate_dx, float* dbias,
                           int64_t P, int nseq, int len, int qdiv, int64_t s_hi, int64_t s_lo, int64_t s_t, int precision, void* stream) {
    LstmPackView v = view_pack(pack);
    SeqMap m{nseq, len, qdiv, s_hi, s_lo, s_t};
    CK(launch_lstm_bwd(v.rec, G, Cst, dH, dbias, m, is_split(precision), S(stream)));
    if (dx) {
        GemmNtArgs a = nt_args(G, 2 * kG, v.wiht_hi, v.wiht_lo, 2 * kG, 0, dx, kN, (int)P, kN, 2 * kG);
        a.accumulate = accumulate_dx;
        CK(gemm_nt(a, is_split(precision), S(stream)));
    }
    return 0;
}

int dp_groupnorm_finalize(const double* stats, float* mean_rstd, int groups, double count, double eps, void* stream) {
    CK(launch_gn_finalize(stats, mean_rstd, groups, count, eps, S(stream)));
    return 0;
}
int dp_groupnorm_residual_f32(const float* y, const float* res, float* out, const float* mean_rstd, const float* gamma, const float* beta,
                              int64_t rows, int rows_per_group, int C, const float* cw, const float* cb, const float* prelu_slope,
                              void* stream) {
    CK(launch_gn_apply(y, res, out, mean_rstd, gamma, beta, rows, rows_per_group, C, cw, cb, prelu_slope, S(stream)));
    return 0;
}

int dp_attention_forward_f32(const float* qkv, float* o, float* lse, int E, int heads, int nseq, int len, int qdiv, int64_t s_hi, int64_t s_lo,
                             int64_t s_t, void* stream) {
    if (heads <= 0 || E % heads || (E / heads != 16 && E / heads != 32)) return fail("dp_attention_forward_f32: head width E/heads must be 16 or 32");
    SeqMap m{nseq, len, qdiv, s_hi, s_lo, s_t};
    CK(launch_attn_fwd(qkv, o, lse, E, heads, m, S(stream)));
    return 0;
}
int dp_attention_forward_tc_f32(const float* qkv, float* o, void* o_hi, void* o_lo, float* lse, int E, int heads, int nseq, int len, int qdiv,
                                int64_t s_hi, int64_t s_lo, int64_t s_t, int precision, void* stream) {
    SeqMap m{nseq, len, qdiv, s_hi, s_lo, s_t};
    if (!attn_bwd_mma_supported(E, heads, m)) return fail("dp_attention_forward_tc_f32: head width E/heads must be 16 or 32 and the sequence length <= 320");
    CK(launch_attn_fwd_mma(qkv, o, static_cast<__nv_bfloat16*>(o_hi), static_cast<__nv_bfloat16*>(o_lo), lse, E, heads, m, is_split(precision), S(stream)));
    return 0;
}
int dp_attention_backward_tc_f32(const float* qkv, const float* o, const float* lse, const float* d_o, float* d_qkv, int E, int heads, int nseq,
                                 int len, int qdiv, int64_t s_hi, int64_t s_lo, int64_t s_t, int precision, void* stream) {
    SeqMap m{nseq, len, qdiv, s_hi, s_lo, s_t};
    if (!attn_bwd_mma_supported(E, heads, m)) return fail("dp_attention_backward_tc_f32: head width E/heads must be 16 or 32 and the sequence length <= 320");
    CK(launch_attn_bwd_mma(qkv, o, lse, d_o, d_qkv, E, heads, m, is_split(precision), S(stream)));
    return 0;
}
int dp_attention_backward_f32(const float* qkv, const float* o, const float* lse, const float* d_o, float* d_qkv, int E, int heads, int nseq,
                              int len, int qdiv, int64_t s_hi, int64_t s_lo, int64_t s_t, void* stream) {
    if (heads <= 0 || E % heads || (E / heads != 16 && E / heads != 32)) return fail("dp_attention_backward_f32: head width E/heads must be 16 or 32");
    SeqMap m{nseq, len, qdiv, s_hi, s_lo, s_t};
    CK(launch_attn_bwd(qkv, o, lse, d_o, d_qkv, E, heads, m, S(stream)));
    return 0;
}
int dp_attention_forward_planes_f32(const void* qkv_hi, const void* qkv_lo, float* o, void* o_hi, void* o_lo, float* lse, int E, int heads,
                                    int inter, int B, int Sc, int K, int precision, void* stream) {
    LstmFusedGeom gm;
    gm.inter = inter; gm.len = inter ? Sc : K; gm.nseq = inter ? B * K : B * Sc; gm.K = K; gm.S = Sc; gm.B = B;
    if (!attn_tc5_supported(E, heads, gm)) return fail("dp_attention_forward_planes_f32: need head width 16 or 32 and sequence length <= 256");
    CK(launch_attn_fwd_tc5((const __nv_bfloat16*)qkv_hi, (const __nv_bfloat16*)qkv_lo, o, (__nv_bfloat16*)o_hi, (__nv_bfloat16*)o_lo, lse, E, heads,
                           gm, is_split(precision), S(stream)));
    return 0;
}
int dp_add_layernorm_f32(const float* a, const float* b, float* z_out, float* out, const float* res, const float* gamma, const float* beta,
                         int64_t rows, int E, float eps, void* stream) {
    if (E != 64 && E != 128 && E != 256) return fail("dp_add_layernorm_f32: E must be 64, 128 or 256 (got %d)", E);
    CK(launch_add_ln(a, b, z_out, out, res, gamma, beta, rows, E, eps, nullptr, nullptr, nullptr, S(stream)));
    return 0;
}
int dp_layernorm_backward_f32(const float* dy, const float* z, float* dz, float* acc, const float* gamma, int64_t rows, int E, float eps,
                              float* dgamma, float* dbeta, void* stream) {
    if (E != 64 && E != 128 && E != 256) return fail("dp_layernorm_backward_f32: E must be 64, 128 or 256 (got %d)", E);
    CK(launch_ln_bwd(dy, z, dz, acc, gamma, rows, E, eps, dgamma, dbeta, S(stream)));
    return 0;
}

int64_t dp_pit_loss_workspace_bytes(int B) { return (int64_t)B * (18 * 8 + 6 * 4) + 64; }
int dp_pit_loss_forward(const float* est, const float* tgt, int B, int T, int sdr_type, int threshold_byloss, void* ws, float* pw,
                        float* loss, int32_t* perm, void* stream) {
    if (sdr_type < 0 || sdr_type > 2) return fail("dp_pit_loss_forward: sdr_type must be 0 (snr), 1 (sisdr) or 2 (sdsdr)");
    PitWsView v = pit_view(ws, B);
    PitLossWs w{v.sums, v.second, v.noise};
    CK(launch_pit_loss_fwd(est, tgt, B, T, sdr_type, threshold_byloss, w, pw, loss, perm, v.coef, S(stream)));
    return 0;
}
int dp_pit_loss_backward(const float* est, const float* tgt, int B, int T, const void* ws, float grad_scale, float* d_est, void* stream) {
    PitWsView v = pit_view(const_cast<void*>(ws), B);
    CK(launch_pit_loss_bwd(est, tgt, B, T, v.sums, v.coef, grad_scale, d_est, S(stream)));
    return 0;
}
int dp_pit_reorder(const float* est, const int32_t* perm, float* out, int B, int T, void* stream) {
    CK(launch_reorder_sources(est, perm, out, B, T, S(stream)));
    return 0;
}

// general n_src: workspace = sums [B][2N] | second [B][2N*N+N] | noise [B][N*N] (fp64) | coef [B][N][3] (fp32)
namespace {
struct PitNView { double* sums; double* second; double* noise; float* coef; };
PitNView pitn_view(void* ws, int B, int N) {
    char* b = static_cast<char*>(ws);
    PitNView v;
    v.sums = reinterpret_cast<double*>(b);
    v.second = v.sums + (size_t)B * 2 * N;
    v.noise = v.second + (size_t)B * (2 * N * N + N);
    v.coef = reinterpret_cast<float*>(v.noise + (size_t)B * N * N);
    return v;
}
}  // namespace
int64_t dp_pitn_loss_workspace_bytes(int B, int N) { return (int64_t)B * ((3 * N * N + 3 * N) * 8 + 3 * N * 4) + 64; }
int dp_pitn_loss_forward(const float* est, const float* tgt, int B, int N, int T, int sdr_type, int threshold_byloss, void* ws, float* pw,
                         float* loss, int32_t* perm, void* stream) {
    if (sdr_type < 0 || sdr_type > 2) return fail("dp_pitn_loss_forward: sdr_type must be 0 (snr), 1 (sisdr) or 2 (sdsdr)");
    if (N < 1 || N > 4) return fail("dp_pitn_loss_forward: n_src must be 1 .. 4 (got %d)", N);
    PitNView v = pitn_view(ws, B, N);
    PitLossWs w{v.sums, v.second, v.noise};
    CK(launch_pitn_loss_fwd(est, tgt, B, N, T, sdr_type, threshold_byloss, w, pw, loss, perm, v.coef, S(stream)));
    return 0;
}
int dp_pitn_loss_backward(const float* est, const float* tgt, int B, int N, int T, const void* ws, float grad_scale, float* d_est, void* stream) {
    if (N < 1 || N > 4) return fail("dp_pitn_loss_backward: n_src must be 1 .. 4 (got %d)", N);
    PitNView v = pitn_view(const_cast<void*>(ws), B, N);
    CK(launch_pitn_loss_bwd(est, tgt, B, N, T, v.sums, v.coef, grad_scale, d_est, S(stream)));
    return 0;
}
int dp_pitn_reorder(const float* est, const int32_t* perm, float* out, int B, int N, int T, void* stream) {
    CK(launch_reorder_sources_n(est, perm, out, B, N, T, S(stream)));
    return 0;
}

int64_t dp_bss_sdr_workspace_bytes(int B, int n_src, int filter_len) { return (int64_t)bss_sdr_workspace_bytes(B, n_src, filter_len); }
int dp_bss_sdr_pit(const float* est, const float* ref, int B, int n_src, int T, int filter_len, void* ws, float* mean_sdr, float* sdr_mat,
                   void* stream) {
    if (n_src < 1 || n_src > 4) return fail("dp_bss_sdr_pit: n_src must be 1 .. 4 (got %d)", n_src);
    if (filter_len < 1 || filter_len > 512) return fail("dp_bss_sdr_pit: filter_len must be 1 .. 512 (got %d)", filter_len);
    CK(launch_bss_sdr_pit(est, ref, B, n_src, T, filter_len, ws, mean_sdr, sdr_mat, S(stream)));
    return 0;
}

int dp_adam_clip_step(float* p, const float* g, float* m, float* v, int64_t n, double* norm2, float grad_scale, float max_norm, float lr,
                      float beta1, float beta2, float eps, int step, float weight_decay, void* stream) {
    if (step < 1) return fail("dp_adam_clip_step: step counts from 1");
    CK(cudaMemsetAsync(norm2, 0, sizeof(double), S(stream)));
    if (max_norm > 0.f) CK(launch_sumsq(g, n, norm2, S(stream)));
    double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    CK(launch_adam_clip(p, g, m, v, n, norm2, grad_scale, max_norm, lr, beta1, beta2, eps, (float)bc1, (float)bc2, weight_decay, S(stream)));
    return 0;
}

int dp_adam_set_hyper(float* hyper, float lr, float beta1, float beta2, int step, void* stream) {
    if (!hyper || step < 1) return fail("dp_adam_set_hyper: hyper = device pointer to three floats, step counts from 1");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    CK(launch_set_hyper(hyper, lr, (float)bc1, (float)bc2, S(stream)));
    return 0;
}
int dp_adam_clip_step_dev(float* p, const float* g, float* m, float* v, int64_t n, double* norm2, float grad_scale, float max_norm,
                          const float* hyper, float beta1, float beta2, float eps, float weight_decay, void* stream) {
    if (!hyper) return fail("dp_adam_clip_step_dev: hyper = device pointer to (lr, 1 - beta1^t, 1 - beta2^t)");
    CK(cudaMemsetAsync(norm2, 0, sizeof(double), S(stream)));
    if (max_norm > 0.f) CK(launch_sumsq(g, n, norm2, S(stream)));
    CK(launch_adam_clip(p, g, m, v, n, norm2, grad_scale, max_norm, 0.f, beta1, beta2, eps, 1.f, 1.f, weight_decay, S(stream), hyper));
    return 0;
}

}  // extern "C"

// ================================================================================================
// TasNet / DPRNN engine
// ================================================================================================
struct dp_tasnet {
    dp_tasnet_config cfg;
    std::vector<int64_t> off;
    int64_t n_params;
    int npath;
    int ppath;  // parameter-table entries per path: 12 (DPRNN) or 18 (DPTNet)
    int launches;
};

namespace {

struct Geo {
    int B, T, rest_w, Tp, L, rest_s, Sc, K, P;  // P = Sc*K positions per utterance
    long long PT, BL;
};
int make_geo(const dp_tasnet* h, int B, int T, Geo& g) {
    if (B <= 0 || T <= 0) return fail("need B > 0 and T > 0 (got B=%d T=%d)", B, T);
    g.B = B; g.T = T; g.K = h->cfg.block_size;
    if (dp_wave_geometry(T, h->cfg.win, &g.rest_w, &g.L)) return 1;
    g.Tp = T + g.rest_w + h->cfg.win;
    if (dp_seg_geometry(g.L, g.K, &g.rest_s, &g.Sc)) return 1;
    g.P = g.Sc * g.K;
    g.PT = (long long)B * g.P;
    g.BL = (long long)B * g.L;
    if (g.PT * 256 >= 0xffffffffLL) return fail("batch too large for 32-bit position indexing");
    return 0;
}

struct Layout {
    size_t xp, E, En, Fb, F2, Z, Mk, Mx, D;
    size_t small, small_bytes;  // GroupNorm statistics (zeroed by every forward)
    size_t redz, redz_bytes;    // GroupNorm-backward reductions (zeroed by every backward)
    size_t statsE, mrE, redE;
    std::vector<size_t> X, G, H, Cst, Y, stats, mr, red, dpack;
    std::vector<size_t> QKV, Oa, LSE, Z1, S1, Z2;  // DPTNet only
    std::vector<size_t> Xhl, Hhl, Hphl;            // bf16 hi/lo operand planes (TMA backend): [hi | lo]
    size_t dGhl, dYhl;
    size_t tXhl, tQKVhl, tOhl, tS1hl, tHrhl;       // DPTNet forward operand planes (shared by all paths, nothing is saved)
    std::vector<size_t> Spre;                      // DPTNet + unfold, training: x + norm2(..) before the concat_block
    std::vector<size_t> S1hl;                      // DPTNet training: planes of the LSTM input (norm1 output) per path
    // backward temporaries
    size_t dXs, dY, dH, dpad, dMx, dMk, dE, dZ, dF2, dtmp, dQKV, dOa;
    size_t total;
};

void make_layout(const dp_tasnet* h, const Geo& g, bool train, Layout& l) {
    Carver c;
    const int np = h->npath, nspk = h->cfg.num_spk;
    const size_t f = sizeof(float);
    l.xp = c.take((size_t)g.B * g.Tp * f);
    l.E = c.take(g.BL * 64 * f);
    l.En = c.take(g.BL * 64 * f);
    l.Fb = c.take(g.BL * 64 * f);
    l.F2 = c.take(g.BL * 64 * f);
    l.Z = c.take(g.BL * 64 * f);
    l.Mk = c.take(g.BL * 64 * nspk * f);
    l.Mx = c.take(g.BL * 64 * nspk * f);
    l.D = c.take(g.BL * nspk * h->cfg.win * f);
    // small region
    size_t s0 = c.off;
    l.statsE = c.take(2 * g.B * sizeof(double));
    l.stats.resize(np); l.mr.resize(np); l.red.resize(np);
    for (int p = 0; p < np; ++p) l.stats[p] = c.take(2 * g.B * sizeof(double));
    l.small = s0;
    l.small_bytes = c.off - s0;
    l.mrE = c.take(2 * g.B * f);
    for (int p = 0; p < np; ++p) l.mr[p] = c.take(2 * g.B * f);
    s0 = c.off;
    l.redE = c.take(2 * g.B * sizeof(double));
    for (int p = 0; p < np; ++p) l.red[p] = c.take(2 * g.B * sizeof(double));
    l.redz = s0;
    l.redz_bytes = c.off - s0;
    const int nbuf = train ? np : 1;
    l.X.resize(np + 1); l.G.resize(np); l.H.resize(np); l.Cst.resize(np); l.Y.resize(np); l.dpack.resize(np);
    if (train) {
        for (int p = 0; p <= np; ++p) l.X[p] = c.take(g.PT * 64 * f);
    } else {
        size_t x = c.take(g.PT * 64 * f);
        for (int p = 0; p <= np; ++p) l.X[p] = x;  // residual update happens in place
    }
    for (int p = 0; p < nbuf; ++p) {
        l.G[p] = c.take(g.PT * 1024 * f);
        l.H[p] = c.take(g.PT * 256 * f);
        l.Cst[p] = train ? c.take(g.PT * 256 * f) : 0;
        l.Y[p] = c.take(g.PT * 64 * f);
    }
    for (int p = nbuf; p < np; ++p) { l.G[p] = l.G[0]; l.H[p] = l.H[0]; l.Cst[p] = l.Cst[0]; l.Y[p] = l.Y[0]; }
    // operand planes of the TMA-fed GEMMs (hi plane followed by lo plane)
    l.Xhl.assign(np + 1, 0); l.Hhl.assign(np, 0); l.Hphl.assign(np, 0);
    l.dGhl = l.dYhl = 0;
    if (h->cfg.module == DP_MODULE_DPRNN) {
        if (train) {
            for (int p = 0; p <= np; ++p) l.Xhl[p] = c.take(g.PT * 64 * 2 * 2);
        } else {
            size_t x = c.take(g.PT * 64 * 2 * 2);
            for (int p = 0; p <= np; ++p) l.Xhl[p] = x;
        }
        for (int p = 0; p < nbuf; ++p) {
            l.Hhl[p] = c.take(g.PT * 256 * 2 * 2);
            l.Hphl[p] = train ? c.take(g.PT * 256 * 2 * 2) : 0;
        }
        for (int p = nbuf; p < np; ++p) { l.Hhl[p] = l.Hhl[0]; l.Hphl[p] = l.Hphl[0]; }
        if (train) {
            l.dGhl = c.take(g.PT * 1024 * 2 * 2);
            l.dYhl = c.take(g.PT * 64 * 2 * 2);
        }
    }
    const bool xf = h->cfg.module == DP_MODULE_DPTNET;
    l.QKV.assign(np, 0); l.Oa.assign(np, 0); l.LSE.assign(np, 0); l.Z1.assign(np, 0); l.S1.assign(np, 0); l.Z2.assign(np, 0);
    l.dQKV = l.dOa = 0;
    l.tXhl = l.tQKVhl = l.tOhl = l.tS1hl = l.tHrhl = 0;
    l.Spre.assign(np, 0);
    if (xf && train && h->cfg.unfold)
        for (int p = 1; p < np; p += 2) l.Spre[p] = c.take(g.PT * 64 * f);
    l.S1hl.assign(np, 0);
    if (xf && train) {  // what the TMA-fed weight-gradient GEMMs of the BiLSTM read: x planes, h_prev planes, dG / dY planes
        for (int p = 0; p < np; ++p) {
            l.S1hl[p] = c.take(g.PT * 64 * 2 * 2);
            l.Hphl[p] = c.take(g.PT * 256 * 2 * 2);
        }
        l.dGhl = c.take(g.PT * 1024 * 2 * 2);
        l.dYhl = c.take(g.PT * 64 * 2 * 2);
    }
    if (xf) {
        l.tXhl = c.take(g.PT * 64 * 2 * 2);
        l.tQKVhl = c.take(g.PT * 192 * 2 * 2);
        l.tOhl = c.take(g.PT * 64 * 2 * 2);
        l.tS1hl = c.take(g.PT * 64 * 2 * 2);
        l.tHrhl = c.take(g.PT * 256 * 2 * 2);
        for (int p = 0; p < nbuf; ++p) {
            l.QKV[p] = c.take(g.PT * 192 * f);
            l.Oa[p] = c.take(g.PT * 64 * f);
            l.S1[p] = c.take(g.PT * 64 * f);
            if (train) {
                l.LSE[p] = c.take(g.PT * 4 * f);
                l.Z1[p] = c.take(g.PT * 64 * f);
                l.Z2[p] = c.take(g.PT * 64 * f);
            }
        }
        for (int p = nbuf; p < np; ++p) { l.QKV[p] = l.QKV[0]; l.Oa[p] = l.Oa[0]; l.S1[p] = l.S1[0]; }
        if (train) {
            l.dQKV = c.take(g.PT * 192 * f);
            l.dOa = c.take(g.PT * 64 * f);
        }
    }
    if (train) {
        for (int p = 0; p < np; ++p) l.dpack[p] = c.take((65536 + 131072 + 1024) * f);
        l.dpad = c.take((size_t)g.B * nspk * g.Tp * f);
        l.dXs = c.take(g.PT * 64 * f);
        l.dY = c.take(g.PT * 64 * f);
        l.dH = c.take(g.PT * 256 * f);
        l.dMx = c.take(g.BL * 64 * nspk * f);
        l.dMk = c.take(g.BL * 64 * nspk * f);
        l.dE = c.take(g.BL * 64 * f);
        l.dZ = c.take(g.BL * 64 * f);
        l.dF2 = c.take(g.BL * 64 * f);
        l.dtmp = c.take(g.BL * 64 * f);
    }
    l.total = c.off;
}

TmaGemmArgs tma_args(const __nv_bfloat16* ahl, long long plane, long long lda, const __nv_bfloat16* whi, const __nv_bfloat16* wlo, int ldw,
                     float* C, int ldc, int M, int N, int K) {
    TmaGemmArgs a;
    memset(&a, 0, sizeof(a));
    a.A_hi = ahl; a.A_lo = ahl + plane; a.lda = lda; a.W_hi = whi; a.W_lo = wlo; a.ldw = ldw;
    a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K; a.bias_scale = 1.f;
    return a;
}

SeqMap path_map(const Geo& g, int pp) {
    SeqMap m;
    if ((pp & 1) == 0) {  // intra-chunk (row): sequence (b,s), time k
        m.nseq = g.B * g.Sc; m.len = g.K; m.qdiv = 1 << 30; m.s_hi = 0; m.s_lo = g.K; m.s_t = 1;
    } else {              // inter-chunk (col): sequence (b,k), time s
        m.nseq = g.B * g.K; m.len = g.Sc; m.qdiv = g.K; m.s_hi = (long long)g.Sc * g.K; m.s_lo = 1; m.s_t = g.K;
    }
    return m;
}

}  // namespace

extern "C" {

int dp_tasnet_create(const dp_tasnet_config* cfg, const int64_t* offsets, int n_offsets, int64_t n_params, dp_tasnet** out) {
    if (!cfg || !offsets || !out) return fail("dp_tasnet_create: null argument");
    if (cfg->enc_dim != kN || cfg->bn_dim != kN || cfg->hidden_dim != kH)
        return fail("dp_tasnet_create: this build is specialised for enc_dim = bn_dim = %d, hidden_dim = %d (got %d/%d/%d)", kN, kH,
                    cfg->enc_dim, cfg->bn_dim, cfg->hidden_dim);
    if (cfg->win != 16) return fail("dp_tasnet_create: win must be 16 (got %d)", cfg->win);
    if (cfg->num_spk < 1 || cfg->num_spk > 4) return fail("dp_tasnet_create: num_spk must be in 1..4");
    if (cfg->block_size <= 0 || (cfg->block_size & 1)) return fail("dp_tasnet_create: block_size must be even and positive");
    if (cfg->layer < 1) return fail("dp_tasnet_create: layer must be >= 1");
    if (cfg->module != DP_MODULE_DPRNN && cfg->module != DP_MODULE_DPTNET) return fail("dp_tasnet_create: module must be DP_MODULE_DPRNN or DP_MODULE_DPTNET");
    const int ppath = cfg->module == DP_MODULE_DPTNET ? DP_TASNET_PATH_PARAMS_DPTNET : DP_TASNET_PATH_PARAMS;
    int need = DP_TASNET_HEAD_PARAMS + ppath * 2 * cfg->layer;
    if (n_offsets != need) return fail("dp_tasnet_create: expected %d parameter offsets, got %d", need, n_offsets);
    for (int i = 0; i < n_offsets; ++i) {
        bool optional = (i >= 9 && i <= 11) && !cfg->unfold;
        if (!optional && (offsets[i] < 0 || offsets[i] >= n_params)) return fail("dp_tasnet_create: offset %d out of range", i);
        if (!optional && (offsets[i] & 7)) return fail("dp_tasnet_create: parameter %d must start at a multiple of 8 elements in the flat buffer", i);
    }
    dp_tasnet* h = new (std::nothrow) dp_tasnet();
    if (!h) return fail("dp_tasnet_create: out of host memory");
    h->cfg = *cfg;
    h->off.assign(offsets, offsets + n_offsets);
    h->n_params = n_params;
    h->npath = 2 * cfg->layer;
    h->ppath = ppath;
    h->launches = 0;
    *out = h;
    return 0;
}
void dp_tasnet_destroy(dp_tasnet* h) { delete h; }

inline size_t tc5_pack_stride() { return (lstm_tc5_pack_bytes() + 255) & ~(size_t)255; }
int64_t dp_tasnet_pack_bytes(const dp_tasnet* h) {
    size_t flat = ((size_t)h->n_params * 2 + 255) & ~(size_t)255;
    return (int64_t)(2 * flat + (size_t)h->npath * PK_BYTES + (size_t)h->npath * tc5_pack_stride());
}
int64_t dp_tasnet_workspace_bytes(const dp_tasnet* h, int B, int T, int train) {
    Geo g;
    if (make_geo(h, B, T, g)) return -1;
    Layout l;
    make_layout(h, g, train != 0, l);
    return (int64_t)l.total;
}
int dp_tasnet_last_launches(const dp_tasnet* h) { return h->launches; }

int dp_tasnet_pack(dp_tasnet* h, const float* params, void* pack, void* stream) {
    size_t flat = ((size_t)h->n_params * 2 + 255) & ~(size_t)255;
    char* b = static_cast<char*>(pack);
    CK(launch_split_bf16(params, (__nv_bfloat16*)b, (__nv_bfloat16*)(b + flat), h->n_params, S(stream)));
    for (int pp = 0; pp < h->npath; ++pp) {
        const int64_t* o = h->off.data() + DP_TASNET_HEAD_PARAMS + h->ppath * pp;
        const float* wi[2] = {params + o[0], params + o[4]};
        const float* wh[2] = {params + o[1], params + o[5]};
        const float* bi[2] = {params + o[2], params + o[6]};
        const float* bh[2] = {params + o[3], params + o[7]};
        CK(launch_pack_lstm(wi, wh, bi, bh, params + o[8], out_pack(b + 2 * flat + (size_t)pp * PK_BYTES), S(stream)));
        CK(launch_pack_lstm_tc5(wi, wh, b + 2 * flat + (size_t)h->npath * PK_BYTES + (size_t)pp * tc5_pack_stride(), S(stream)));
    }
    return 0;
}

int dp_tasnet_forward(dp_tasnet* h, const float* params, const void* pack, const float* mixture, float* est, void* ws, int B, int T,
                      int train, int precision, void* stream) {
    Geo g;
    if (make_geo(h, B, T, g)) return 1;
    Layout l;
    make_layout(h, g, train != 0, l);
    cudaStream_t st = S(stream);
    const bool sp = is_split(precision);
    const int nspk = h->cfg.num_spk, win = h->cfg.win, stride = win / 2;
    const int64_t* o = h->off.data();
    const size_t flat = ((size_t)h->n_params * 2 + 255) & ~(size_t)255;
    const __nv_bfloat16* whi = reinterpret_cast<const __nv_bfloat16*>(pack);
    const __nv_bfloat16* wlo = reinterpret_cast<const __nv_bfloat16*>(static_cast<const char*>(pack) + flat);
    const char* lpack = static_cast<const char*>(pack) + 2 * flat;
    int nl = 0;

    CK(cudaMemsetAsync(at<char>(ws, l.small), 0, l.small_bytes, st));
    // pad + encoder (Conv1d 1->64, k16, s8, no bias) as a GEMM over overlapping frames          gc3_network.py:123-140
    float* xp = at<float>(ws, l.xp);
    CK(launch_pad_rows(mixture, xp, B, T, g.Tp, stride, st)); ++nl;
    {
        GemmNtArgs a = nt_args(xp, stride, whi + o[0], wlo + o[0], win, 0, at<float>(ws, l.E), 64, (int)g.BL, 64, win);
        a.a_rpb = g.L; a.a_skip = 1;
        a.stats = at<double>(ws, l.statsE); a.rows_per_group = g.L;
        CK(launch_gemm_nt(a, sp, st)); ++nl;
    }
    // bottleneck: GroupNorm(1,64,eps=2^-23) + 1x1 conv                                            gc3_network.py:53-56,142
    CK(launch_gn_finalize(at<double>(ws, l.statsE), at<float>(ws, l.mrE), B, (double)g.L * 64, 1.1920928955078125e-07, st)); ++nl;
    CK(launch_gn_apply(at<float>(ws, l.E), nullptr, at<float>(ws, l.En), at<float>(ws, l.mrE), params + o[1], params + o[2], g.BL, g.L, 64,
                       nullptr, nullptr, nullptr, st)); ++nl;
    {
        GemmNtArgs a = nt_args(at<float>(ws, l.En), 64, whi + o[3], wlo + o[3], 64, 0, at<float>(ws, l.Fb), 64, (int)g.BL, 64, 64);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
    }
    // segmentation into 50%-overlapping chunks, channels-last                                     gc3_basics.py:79-91
    CK(launch_segment_cl(at<float>(ws, l.Fb), at<float>(ws, l.X[0]), B, g.L, g.K, g.Sc, 64, st)); ++nl;
    const bool tma = g_backend == 2 && h->cfg.module == DP_MODULE_DPRNN;
    const long long plX = g.PT * 64, plH = g.PT * 256;  // elements per plane
    if (tma) {
        __nv_bfloat16* x0 = at<__nv_bfloat16>(ws, l.Xhl[0]);
        CK(launch_split_rows(at<float>(ws, l.X[0]), 64, x0, sp ? x0 + plX : nullptr, g.PT, 64, 0, st)); ++nl;
    }
    const bool xtma = g_backend == 2 && h->cfg.module == DP_MODULE_DPTNET;
    if (xtma) {
        __nv_bfloat16* x0 = at<__nv_bfloat16>(ws, l.tXhl);
        CK(launch_split_rows(at<float>(ws, l.X[0]), 64, x0, sp ? x0 + plX : nullptr, g.PT, 64, 0, st)); ++nl;
    }

    for (int pp = 0; pp < h->npath; ++pp) {                                                     // dprnn.py:62-82
        const int64_t* po = o + DP_TASNET_HEAD_PARAMS + h->ppath * pp;
        LstmPackView v = view_pack(lpack + (size_t)pp * PK_BYTES);
        float* X = at<float>(ws, l.X[pp]);
        float* G = at<float>(ws, l.G[pp]);
        float* H = at<float>(ws, l.H[pp]);
        float* Y = at<float>(ws, l.Y[pp]);
        if (h->cfg.module == DP_MODULE_DPTNET) {                                                // dptnet.py:66-82,147-157
            const SeqMap m = path_map(g, pp);
            float* QKV = at<float>(ws, l.QKV[pp]);
            float* Oa = at<float>(ws, l.Oa[pp]);
            float* S1 = at<float>(ws, l.S1[pp]);
            if (xtma) {
                // TMA-fed tcgen05 GEMMs on operand planes + tcgen05 attention; fp32 copies only where the backward reads them
                __nv_bfloat16* Xh = at<__nv_bfloat16>(ws, l.tXhl);
                __nv_bfloat16* Qh = at<__nv_bfloat16>(ws, l.tQKVhl);
                __nv_bfloat16* Oh = at<__nv_bfloat16>(ws, l.tOhl);
                __nv_bfloat16* S1h = at<__nv_bfloat16>(ws, train ? l.S1hl[pp] : l.tS1hl);
                __nv_bfloat16* Hrh = at<__nv_bfloat16>(ws, l.tHrhl);
                const long long plQ = g.PT * 192;
                LstmFusedGeom gm;
                gm.inter = pp & 1; gm.len = m.len; gm.nseq = m.nseq; gm.K = g.K; gm.S = g.Sc; gm.B = B;
                const bool tc_attn = attn_tc5_supported(64, 4, gm) && !attn_fwd_prefers_mma(64, 4, m.len, sp);
                {   // in_proj: [q|k|v] = x W_in^T + b_in
                    TmaGemmArgs a = tma_args(Xh, plX, 64, whi + po[12], wlo + po[12], 64, (train || !tc_attn) ? QKV : nullptr, 192, (int)g.PT, 192, 64);
                    if (tc_attn) { a.C_hi = Qh; a.C_lo = sp ? Qh + plQ : nullptr; a.ldch = 192; }
                    a.bias = params + po[13];
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                if (tc_attn) {
                    CK(launch_attn_fwd_tc5(Qh, sp ? Qh + plQ : nullptr, train ? Oa : nullptr, Oh, sp ? Oh + plX : nullptr,
                                           train ? at<float>(ws, l.LSE[pp]) : nullptr, 64, 4, gm, sp, st)); ++nl;
                } else {
                    if (attn_bwd_mma_supported(64, 4, m)) {   // warp-level tensor cores (online softmax) on the fp32 QKV
                        CK(launch_attn_fwd_mma(QKV, train ? Oa : nullptr, Oh, sp ? Oh + plX : nullptr, train ? at<float>(ws, l.LSE[pp]) : nullptr, 64, 4, m,
                                               sp, st)); ++nl;
                    } else {
                        CK(launch_attn_fwd(QKV, train ? Oa : nullptr, train ? at<float>(ws, l.LSE[pp]) : nullptr, 64, 4, m, st, Oh, sp ? Oh + plX : nullptr)); ++nl;
                    }
                }
                {   // out_proj
                    TmaGemmArgs a = tma_args(Oh, plX, 64, whi + po[14], wlo + po[14], 64, Y, 64, (int)g.PT, 64, 64);
                    a.bias = params + po[15];
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                // src = norm1(src + attn)
                CK(launch_add_ln(Y, X, train ? at<float>(ws, l.Z1[pp]) : nullptr, S1, nullptr, params + po[16], params + po[17], g.PT, 64, 1e-5f,
                                 nullptr, nullptr, nullptr, st, S1h, sp ? S1h + plX : nullptr)); ++nl;
                {   // BiLSTM "feed-forward": in-projection, recurrence, ReLU, Linear(256 -> 64)
                    TmaGemmArgs a = tma_args(S1h, plX, 64, v.wih_hi, v.wih_lo, 64, G, 1024, (int)g.PT, 1024, 64);
                    a.bias = v.bias;
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                if (train) {  // h_prev planes for the dW_hh gradient GEMM
                    LstmPlanes hp;
                    memset(&hp, 0, sizeof(hp));
                    hp.hp_hi = at<__nv_bfloat16>(ws, l.Hphl[pp]);
                    hp.hp_lo = sp ? hp.hp_hi + plH : nullptr;
                    CK(launch_lstm_fwd(v.rec, G, H, at<float>(ws, l.Cst[pp]), m, sp, true, st, &hp)); ++nl;
                } else {
                    CK(launch_lstm_fwd(v.rec, G, H, nullptr, m, sp, false, st)); ++nl;
                }
                CK(launch_split_rows(H, 256, Hrh, sp ? Hrh + plH : nullptr, g.PT, 256, 1, st)); ++nl;   // relu(H) as operand planes
                {
                    TmaGemmArgs a = tma_args(Hrh, plH, 256, whi + po[8], wlo + po[8], 256, Y, 64, (int)g.PT, 64, 256);
                    a.bias = params + po[9];
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                // out = x + norm2(src + ff)  (+ concat_block after the inter-chunk path when unfold); planes of the result feed the next in_proj
                const bool cat = h->cfg.unfold && (pp & 1);
                if (cat && train) {  // keep the sum before the concat_block for its backward
                    float* Spre = at<float>(ws, l.Spre[pp]);
                    CK(launch_add_ln(Y, S1, at<float>(ws, l.Z2[pp]), Spre, X, params + po[10], params + po[11], g.PT, 64, 1e-5f, nullptr, nullptr,
                                     nullptr, st)); ++nl;
                    CK(launch_affine_prelu(Spre, at<float>(ws, l.X[pp + 1]), g.PT, 64, params + o[9], params + o[10], params + o[11], Xh,
                                           sp ? Xh + plX : nullptr, st)); ++nl;
                    continue;
                }
                CK(launch_add_ln(Y, S1, train ? at<float>(ws, l.Z2[pp]) : nullptr, at<float>(ws, l.X[pp + 1]), X, params + po[10], params + po[11],
                                 g.PT, 64, 1e-5f, cat ? params + o[9] : nullptr, cat ? params + o[10] : nullptr, cat ? params + o[11] : nullptr,
                                 st, Xh, sp ? Xh + plX : nullptr)); ++nl;
                continue;
            }
            {   // in_proj: [q|k|v] = x W_in^T + b_in
                GemmNtArgs a = nt_args(X, 64, whi + po[12], wlo + po[12], 64, 0, QKV, 192, (int)g.PT, 192, 64);
                a.bias = params + po[13];
                CK(launch_gemm_nt(a, sp, st)); ++nl;
            }
            CK(launch_attn_fwd(QKV, Oa, train ? at<float>(ws, l.LSE[pp]) : nullptr, 64, 4, m, st)); ++nl;
            {   // out_proj
                GemmNtArgs a = nt_args(Oa, 64, whi + po[14], wlo + po[14], 64, 0, Y, 64, (int)g.PT, 64, 64);
                a.bias = params + po[15];
                CK(launch_gemm_nt(a, sp, st)); ++nl;
            }
            // src = norm1(src + attn)
            CK(launch_add_ln(Y, X, train ? at<float>(ws, l.Z1[pp]) : nullptr, S1, nullptr, params + po[16], params + po[17], g.PT, 64, 1e-5f,
                             nullptr, nullptr, nullptr, st)); ++nl;
            {   // the "feed-forward" is a BiLSTM over the sequence axis + ReLU + Linear(256 -> 64)
                GemmNtArgs a = nt_args(S1, 64, v.wih_hi, v.wih_lo, 64, 0, G, 1024, (int)g.PT, 1024, 64);
                a.bias = v.bias;
                CK(gemm_nt(a, sp, st)); ++nl;
            }
            CK(launch_lstm_fwd(v.rec, G, H, train ? at<float>(ws, l.Cst[pp]) : nullptr, m, sp, train != 0, st)); ++nl;
            {
                GemmNtArgs a = nt_args(H, 256, whi + po[8], wlo + po[8], 256, 0, Y, 64, (int)g.PT, 64, 256);
                a.bias = params + po[9];
                a.relu_a = 1;
                CK(launch_gemm_nt(a, sp, st)); ++nl;
            }
            // out = x + norm2(src + ff)  (+ concat_block after the inter-chunk path when unfold)
            const bool cat = h->cfg.unfold && (pp & 1);
            if (cat && train) {
                float* Spre = at<float>(ws, l.Spre[pp]);
                CK(launch_add_ln(Y, S1, at<float>(ws, l.Z2[pp]), Spre, X, params + po[10], params + po[11], g.PT, 64, 1e-5f, nullptr, nullptr, nullptr,
                                 st)); ++nl;
                CK(launch_affine_prelu(Spre, at<float>(ws, l.X[pp + 1]), g.PT, 64, params + o[9], params + o[10], params + o[11], nullptr, nullptr,
                                       st)); ++nl;
                continue;
            }
            CK(launch_add_ln(Y, S1, train ? at<float>(ws, l.Z2[pp]) : nullptr, at<float>(ws, l.X[pp + 1]), X, params + po[10], params + po[11],
                             g.PT, 64, 1e-5f, cat ? params + o[9] : nullptr, cat ? params + o[10] : nullptr, cat ? params + o[11] : nullptr,
                             st)); ++nl;
            continue;
        }
        if (tma) {
            // every GEMM operand is a pair of bf16 planes written by its producer; TMA -> tcgen05 (gemm_tma.cu)
            __nv_bfloat16* Xhl = at<__nv_bfloat16>(ws, l.Xhl[pp]);
            __nv_bfloat16* Hhl = at<__nv_bfloat16>(ws, l.Hhl[pp]);
            LstmPlanes pl;
            pl.h_hi = Hhl; pl.h_lo = sp ? Hhl + plH : nullptr;
            pl.hp_hi = train ? at<__nv_bfloat16>(ws, l.Hphl[pp]) : nullptr;
            pl.hp_lo = (train && sp) ? pl.hp_hi + plH : nullptr;
            // automatic: bf16 mode and at least ~one wave of sequences.  Measured (tests/tools/time_lstm.py, DPRNN forward): B = 16 bf16
            // 5.92 ms fused vs 5.95 ms unfused; B = 1 bf16 5.30 vs 1.42 ms (the fused kernel's 64-sequence tiles leave most SMs idle)
            const int nseq_path = (pp & 1) ? B * g.K : B * g.Sc;
            if (g_fused_lstm == 2 || (g_fused_lstm == 1 && !sp && nseq_path >= 1312)) {
                // input projection + recurrence in one tcgen05 kernel: the [P,1024] gate pre-activations never exist in HBM
                LstmFusedGeom gm;
                gm.inter = pp & 1; gm.len = (pp & 1) ? g.Sc : g.K; gm.nseq = (pp & 1) ? B * g.K : B * g.Sc; gm.K = g.K; gm.S = g.Sc; gm.B = B;
                const char* t5 = static_cast<const char*>(pack) + 2 * flat + (size_t)h->npath * PK_BYTES + (size_t)pp * tc5_pack_stride();
                CK(launch_lstm_fused_fwd(t5, v.bias, Xhl, sp ? Xhl + plX : nullptr, train ? G : nullptr, train ? at<float>(ws, l.Cst[pp]) : nullptr,
                                         pl, gm, sp, train != 0, st)); ++nl;
            } else {
                TmaGemmArgs a = tma_args(Xhl, plX, 64, v.wih_hi, v.wih_lo, 64, G, 1024, (int)g.PT, 1024, 64);
                a.bias = v.bias;
                CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                CK(launch_lstm_fwd(v.rec, G, nullptr, train ? at<float>(ws, l.Cst[pp]) : nullptr, path_map(g, pp), sp, train != 0, st, &pl)); ++nl;
            }
            {
                TmaGemmArgs a = tma_args(Hhl, plH, 256, whi + po[8], wlo + po[8], 256, Y, 64, (int)g.PT, 64, 256);
                a.bias = params + po[9];
                a.stats = at<double>(ws, l.stats[pp]); a.rows_per_group = g.P;
                CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
            }
            CK(launch_gn_finalize(at<double>(ws, l.stats[pp]), at<float>(ws, l.mr[pp]), B, (double)g.P * 64, 1e-8, st)); ++nl;
            const bool cat = h->cfg.unfold && (pp & 1);
            __nv_bfloat16* xn = pp + 1 < h->npath ? at<__nv_bfloat16>(ws, l.Xhl[pp + 1]) : nullptr;
            CK(launch_gn_apply(Y, X, at<float>(ws, l.X[pp + 1]), at<float>(ws, l.mr[pp]), params + po[10], params + po[11], g.PT, g.P, 64,
                               cat ? params + o[9] : nullptr, cat ? params + o[10] : nullptr, cat ? params + o[11] : nullptr, st, xn,
                               (xn && sp) ? xn + plX : nullptr)); ++nl;
            continue;
        }
        {
            GemmNtArgs a = nt_args(X, 64, v.wih_hi, v.wih_lo, 64, 0, G, 1024, (int)g.PT, 1024, 64);
            a.bias = v.bias;
            CK(gemm_nt(a, sp, st)); ++nl;
        }
        CK(launch_lstm_fwd(v.rec, G, H, train ? at<float>(ws, l.Cst[pp]) : nullptr, path_map(g, pp), sp, train != 0, st)); ++nl;
        {
            GemmNtArgs a = nt_args(H, 256, whi + po[8], wlo + po[8], 256, 0, Y, 64, (int)g.PT, 64, 256);
            a.bias = params + po[9];
            a.stats = at<double>(ws, l.stats[pp]); a.rows_per_group = g.P;
            CK(launch_gemm_nt(a, sp, st)); ++nl;
        }
        CK(launch_gn_finalize(at<double>(ws, l.stats[pp]), at<float>(ws, l.mr[pp]), B, (double)g.P * 64, 1e-8, st)); ++nl;
        const bool cat = h->cfg.unfold && (pp & 1);
        CK(launch_gn_apply(Y, X, at<float>(ws, l.X[pp + 1]), at<float>(ws, l.mr[pp]), params + po[10], params + po[11], g.PT, g.P, 64,
                           cat ? params + o[9] : nullptr, cat ? params + o[10] : nullptr, cat ? params + o[11] : nullptr, st)); ++nl;
    }
    // overlap-add, then the DPRNN output 1x1 conv (linear => conv(a)+conv(b) = W(a+b) + 2 bias)    dprnn.py:85, gc3_basics.py:94-109
    CK(launch_overlap_add_cl(at<float>(ws, l.X[h->npath]), at<float>(ws, l.F2), B, g.L, g.K, g.Sc, 64, st)); ++nl;
    {
        GemmNtArgs a = nt_args(at<float>(ws, l.F2), 64, whi + o[4], wlo + o[4], 64, 0, at<float>(ws, l.Z), 64, (int)g.BL, 64, 64);
        a.bias = params + o[5]; a.bias_scale = 2.f;
        CK(launch_gemm_nt(a, sp, st)); ++nl;
    }
    // mask = ReLU(Conv1d 64 -> 64*spk), applied to the raw encoder output                         gc3_network.py:169-174
    {
        GemmNtArgs a = nt_args(at<float>(ws, l.Z), 64, whi + o[6], wlo + o[6], 64, 0, at<float>(ws, l.Mk), 64 * nspk, (int)g.BL, 64 * nspk, 64);
        a.bias = params + o[7]; a.relu = 1;
        CK(launch_gemm_nt(a, sp, st)); ++nl;
    }
    CK(launch_mask_apply(at<float>(ws, l.Mk), at<float>(ws, l.E), at<float>(ws, l.Mx), B, g.L, nspk, 64, st)); ++nl;
    // decoder ConvTranspose1d(64 -> 1, k16, s8): per-frame 64 -> 16 product, stride-8 overlap-add, trim   gc3_network.py:177-179
    {
        GemmNtArgs a = nt_args(at<float>(ws, l.Mx), 64, whi + o[8], wlo + o[8], win, 1, at<float>(ws, l.D), win, (int)(g.BL * nspk), win, 64);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
    }
    CK(launch_dec_ola(at<float>(ws, l.D), est, B * nspk, g.L, win, T, st)); ++nl;
    h->launches = nl;
    return 0;
}

int dp_tasnet_backward(dp_tasnet* h, const float* params, const void* pack, const float* d_est, float* grads, void* ws, int B, int T,
                       int precision, void* stream) {
    Geo g;
    if (make_geo(h, B, T, g)) return 1;
    Layout l;
    make_layout(h, g, true, l);
    cudaStream_t st = S(stream);
    const bool sp = is_split(precision);
    const int nspk = h->cfg.num_spk, win = h->cfg.win, stride = win / 2;
    const int64_t* o = h->off.data();
    const size_t flat = ((size_t)h->n_params * 2 + 255) & ~(size_t)255;
    const __nv_bfloat16* whi = reinterpret_cast<const __nv_bfloat16*>(pack);
    const __nv_bfloat16* wlo = reinterpret_cast<const __nv_bfloat16*>(static_cast<const char*>(pack) + flat);
    const char* lpack = static_cast<const char*>(pack) + 2 * flat;
    const int BLi = (int)g.BL, PTi = (int)g.PT;
    int nl = 0;

    CK(cudaMemsetAsync(at<char>(ws, l.dpack[0]), 0, (size_t)h->npath * (((65536 + 131072 + 1024) * sizeof(float) + 255) & ~(size_t)255), st));
    // ---- decoder: d_out -> padded rows (overlapping 16-sample frames = dD), dMx = dD Wdec^T, dWdec += Mx^T dD
    CK(cudaMemsetAsync(at<char>(ws, l.redz), 0, l.redz_bytes, st));
    float* dpad = at<float>(ws, l.dpad);
    CK(launch_pad_rows(d_est, dpad, B * nspk, T, g.Tp, stride, st)); ++nl;
    {
        GemmNtArgs a = nt_args(dpad, stride, whi + o[8], wlo + o[8], win, 0, at<float>(ws, l.dMx), 64, BLi * nspk, 64, win);
        a.a_rpb = g.L; a.a_skip = 1;
        CK(launch_gemm_nt(a, sp, st)); ++nl;
        GemmTnArgs t = tn_args(at<float>(ws, l.Mx), 64, dpad, stride, grads + o[8], win, BLi * nspk, 64, win);
        t.b_rpb = g.L; t.b_skip = 1;
        CK(launch_gemm_tn(t, sp, st)); ++nl;
    }
    // ---- mask: dMk = dMx * E * [Mk > 0], dE = sum_c dMx * Mk
    CK(launch_mask_bwd(at<float>(ws, l.dMx), at<float>(ws, l.Mk), at<float>(ws, l.E), at<float>(ws, l.dMk), at<float>(ws, l.dE), 0, B, g.L, nspk,
                       64, st)); ++nl;
    {
        GemmNtArgs a = nt_args(at<float>(ws, l.dMk), 64 * nspk, whi + o[6], wlo + o[6], 64, 1, at<float>(ws, l.dZ), 64, BLi, 64, 64 * nspk);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
        GemmTnArgs t = tn_args(at<float>(ws, l.dMk), 64 * nspk, at<float>(ws, l.Z), 64, grads + o[6], 64, BLi, 64 * nspk, 64);
        CK(launch_gemm_tn(t, sp, st)); ++nl;
        CK(launch_colsum(at<float>(ws, l.dMk), 64 * nspk, BLi, 64 * nspk, 1.f, grads + o[7], nullptr, st)); ++nl;
    }
    // ---- DPRNN output conv (applied after the overlap-add, bias counted twice)
    {
        GemmNtArgs a = nt_args(at<float>(ws, l.dZ), 64, whi + o[4], wlo + o[4], 64, 1, at<float>(ws, l.dF2), 64, BLi, 64, 64);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
        GemmTnArgs t = tn_args(at<float>(ws, l.dZ), 64, at<float>(ws, l.F2), 64, grads + o[4], 64, BLi, 64, 64);
        CK(launch_gemm_tn(t, sp, st)); ++nl;
        CK(launch_colsum(at<float>(ws, l.dZ), 64, BLi, 64, 2.f, grads + o[5], nullptr, st)); ++nl;
    }
    // ---- overlap-add backward = segmentation of the frame gradient
    float* dXs = at<float>(ws, l.dXs);
    CK(launch_segment_cl(at<float>(ws, l.dF2), dXs, B, g.L, g.K, g.Sc, 64, st)); ++nl;

    for (int pp = h->npath - 1; pp >= 0; --pp) {
        const int64_t* po = o + DP_TASNET_HEAD_PARAMS + h->ppath * pp;
        LstmPackView v = view_pack(lpack + (size_t)pp * PK_BYTES);
        float* X = at<float>(ws, l.X[pp]);
        float* G = at<float>(ws, l.G[pp]);
        float* H = at<float>(ws, l.H[pp]);
        float* Y = at<float>(ws, l.Y[pp]);
        float* mr = at<float>(ws, l.mr[pp]);
        float* dY = at<float>(ws, l.dY);
        float* dH = at<float>(ws, l.dH);
        float* dpk = at<float>(ws, l.dpack[pp]);
        const SeqMap m = path_map(g, pp);
        if (h->cfg.module == DP_MODULE_DPTNET) {
            float* QKV = at<float>(ws, l.QKV[pp]);
            float* Oa = at<float>(ws, l.Oa[pp]);
            float* S1 = at<float>(ws, l.S1[pp]);
            float* dQKV = at<float>(ws, l.dQKV);
            float* dOa = at<float>(ws, l.dOa);
            if (h->cfg.unfold && (pp & 1)) {  // concat_block (depthwise 1x1 + PReLU) on the stored sum: dXs <- d(sum), parameter gradients
                CK(launch_concat_bwd(dXs, at<float>(ws, l.Spre[pp]), nullptr, nullptr, nullptr, nullptr, g.PT, g.P, 64, params + o[9], params + o[10],
                                     params + o[11], grads + o[9], grads + o[10], grads + o[11], st)); ++nl;
            }
            // norm2 backward: dXs is d(x + norm2(z2)); dY <- d z2 (= d ff = the residual part of d src)
            CK(launch_ln_bwd(dXs, at<float>(ws, l.Z2[pp]), dY, nullptr, params + po[10], g.PT, 64, 1e-5f, grads + po[10], grads + po[11], st)); ++nl;
            if (g_backend == 2) {
                // TMA-fed tcgen05 GEMMs on operand planes for the BiLSTM "feed-forward" (the same kernels as the DPRNN path)
                const long long plX = g.PT * 64, plH = g.PT * 256, plG = g.PT * 1024;
                __nv_bfloat16* dYhl = at<__nv_bfloat16>(ws, l.dYhl);
                __nv_bfloat16* dGhl = at<__nv_bfloat16>(ws, l.dGhl);
                __nv_bfloat16* Hrh = at<__nv_bfloat16>(ws, l.tHrhl);
                const __nv_bfloat16* S1hl = at<__nv_bfloat16>(ws, l.S1hl[pp]);
                const __nv_bfloat16* Hphl = at<__nv_bfloat16>(ws, l.Hphl[pp]);
                CK(launch_split_rows_colsum(dY, 64, dYhl, sp ? dYhl + plX : nullptr, g.PT, 64, grads + po[9], st)); ++nl;   // planes + bias gradient
                CK(launch_split_rows(H, 256, Hrh, sp ? Hrh + plH : nullptr, g.PT, 256, 1, st)); ++nl;   // relu(H), recomputed
                {   // linear2 + ReLU: dH = (dY Wp) where H > 0 ; dWp = dY^T relu(H) (as relu(H)^T dY, stored transposed) ; dbp
                    TmaGemmArgs a = tma_args(dYhl, plX, 64, v.projt_hi, v.projt_lo, 64, dH, 256, PTi, 256, 64);
                    a.mask = H; a.ldmask = 256;
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                    TmaWgradArgs t;
                    memset(&t, 0, sizeof(t));
                    t.A_hi = Hrh; t.A_lo = Hrh + plH; t.lda = 256; t.Mo = 256;
                    t.B0_hi = dYhl; t.B0_lo = dYhl + plX; t.ldb0 = 64; t.nb0 = 64;
                    t.C0 = grads + po[8]; t.ldc0 = 256; t.transpose0 = 1; t.P = PTi; t.scale = 1.f;
                    CK(launch_gemm_tma_tn(t, sp, st)); ++nl;
                }
                CK(launch_lstm_bwd(v.rec, G, at<float>(ws, l.Cst[pp]), dH, dpk + 65536 + 131072, m, sp, st, dGhl, sp ? dGhl + plG : nullptr)); ++nl;
                {   // d src (after norm1) = d z2 + dG W_ih (K = 1024)
                    TmaGemmArgs a = tma_args(dGhl, plG, 1024, v.wiht_hi, v.wiht_lo, 1024, dY, 64, PTi, 64, 1024);
                    a.accumulate = 1;
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                for (int d = 0; d < 2; ++d) {  // dW_ih = dG^T src and dW_hh = dG^T h_prev from ONE pass over this direction's dG
                    TmaWgradArgs t;
                    memset(&t, 0, sizeof(t));
                    t.A_hi = dGhl + d * 512; t.A_lo = dGhl + plG + d * 512; t.lda = 1024; t.Mo = 512;
                    t.B0_hi = S1hl; t.B0_lo = S1hl + plX; t.ldb0 = 64; t.nb0 = 64;
                    t.B1_hi = Hphl + d * 128; t.B1_lo = Hphl + plH + d * 128; t.ldb1 = 256; t.nb1 = 128;
                    t.C0 = dpk + d * 512 * 64; t.ldc0 = 64;
                    t.C1 = dpk + 65536 + d * 65536; t.ldc1 = 128;
                    t.P = PTi; t.scale = 1.f;
                    CK(launch_gemm_tma_tn(t, sp, st)); ++nl;
                }
            } else {
                {   // linear2 + ReLU
                    GemmNtArgs a = nt_args(dY, 64, v.projt_hi, v.projt_lo, 64, 0, dH, 256, PTi, 256, 64);
                    a.mask = H; a.ldmask = 256;
                    CK(launch_gemm_nt(a, sp, st)); ++nl;
                    GemmTnArgs t = tn_args(dY, 64, H, 256, grads + po[8], 256, PTi, 64, 256);
                    t.relu_b = 1;
                    CK(launch_gemm_tn(t, sp, st)); ++nl;
                    CK(launch_colsum(dY, 64, PTi, 64, 1.f, grads + po[9], nullptr, st)); ++nl;
                }
                CK(launch_lstm_bwd(v.rec, G, at<float>(ws, l.Cst[pp]), dH, dpk + 65536 + 131072, m, sp, st)); ++nl;
                {   // d src (after norm1) = d z2 + dG W_ih ; LSTM weight gradients
                    GemmNtArgs a = nt_args(G, 1024, v.wiht_hi, v.wiht_lo, 1024, 0, dY, 64, PTi, 64, 1024);
                    a.accumulate = 1;
                    CK(gemm_nt(a, sp, st)); ++nl;
                    GemmTnArgs t = tn_args(G, 1024, S1, 64, dpk, 64, PTi, 1024, 64);
                    CK(gemm_tn(t, sp, st)); ++nl;
                    for (int d = 0; d < 2; ++d) {
                        GemmTnArgs r = tn_args(G + d * 512, 1024, H + d * 128, 256, dpk + 65536 + d * 65536, 128, PTi, 512, 128);
                        r.shift = (d == 0 ? -1 : 1) * (int)m.s_t;
                        r.tdiv = (pp & 1) ? g.K : 1;
                        r.tmod = m.len;
                        CK(gemm_tn(r, sp, st)); ++nl;
                    }
                }
            }
            // norm1 backward: dY <- d z1 (in place), and the residual branch dXs += d z1
            CK(launch_ln_bwd(dY, at<float>(ws, l.Z1[pp]), dY, dXs, params + po[16], g.PT, 64, 1e-5f, grads + po[16], grads + po[17], st)); ++nl;
            {   // out_proj
                GemmNtArgs a = nt_args(dY, 64, whi + po[14], wlo + po[14], 64, 1, dOa, 64, PTi, 64, 64);
                CK(launch_gemm_nt(a, sp, st)); ++nl;
                GemmTnArgs t = tn_args(dY, 64, Oa, 64, grads + po[14], 64, PTi, 64, 64);
                CK(launch_gemm_tn(t, sp, st)); ++nl;
                CK(launch_colsum(dY, 64, PTi, 64, 1.f, grads + po[15], nullptr, st)); ++nl;
            }
            if (g_backend == 2 && attn_bwd_mma_supported(64, 4, m)) {
                CK(launch_attn_bwd_mma(QKV, Oa, at<float>(ws, l.LSE[pp]), dOa, dQKV, 64, 4, m, sp, st)); ++nl;
            } else {
                CK(launch_attn_bwd(QKV, Oa, at<float>(ws, l.LSE[pp]), dOa, dQKV, 64, 4, m, st)); ++nl;
            }
            {   // in_proj
                GemmNtArgs a = nt_args(dQKV, 192, whi + po[12], wlo + po[12], 64, 1, dXs, 64, PTi, 64, 192);
                a.accumulate = 1;
                CK(launch_gemm_nt(a, sp, st)); ++nl;
                GemmTnArgs t = tn_args(dQKV, 192, X, 64, grads + po[12], 64, PTi, 192, 64);
                CK(launch_gemm_tn(t, sp, st)); ++nl;
                CK(launch_colsum_any(dQKV, 192, PTi, 192, 1.f, grads + po[13], st)); ++nl;
            }
            continue;
        }
        if (h->cfg.unfold && (pp & 1)) {
            CK(launch_concat_bwd(dXs, Y, X, mr, params + po[10], params + po[11], g.PT, g.P, 64, params + o[9], params + o[10], params + o[11],
                                 grads + o[9], grads + o[10], grads + o[11], st)); ++nl;
        }
        // GroupNorm backward (dXs itself is the residual branch of the gradient)
        CK(launch_gn_bwd_reduce(dXs, Y, mr, params + po[10], g.PT, g.P, 64, at<double>(ws, l.red[pp]), grads + po[10], grads + po[11], st)); ++nl;
        CK(launch_gn_bwd_apply(dXs, Y, dY, mr, at<double>(ws, l.red[pp]), params + po[10], g.PT, g.P, 64, st)); ++nl;
        if (g_backend == 2) {
            const long long plX = g.PT * 64, plH = g.PT * 256, plG = g.PT * 1024;
            __nv_bfloat16* dYhl = at<__nv_bfloat16>(ws, l.dYhl);
            __nv_bfloat16* dGhl = at<__nv_bfloat16>(ws, l.dGhl);
            const __nv_bfloat16* Xhl = at<__nv_bfloat16>(ws, l.Xhl[pp]);
            const __nv_bfloat16* Hhl = at<__nv_bfloat16>(ws, l.Hhl[pp]);
            const __nv_bfloat16* Hphl = at<__nv_bfloat16>(ws, l.Hphl[pp]);
            // operand planes of dY and, in the same pass, its column sums = the bias gradient of the out-projection
            CK(launch_split_rows_colsum(dY, 64, dYhl, sp ? dYhl + plX : nullptr, g.PT, 64, grads + po[9], st)); ++nl;
            {   // out-projection Linear(256 -> 64): dH = dY Wp ; dWp = dY^T H (computed as H^T dY, stored transposed)
                TmaGemmArgs a = tma_args(dYhl, plX, 64, v.projt_hi, v.projt_lo, 64, dH, 256, PTi, 256, 64);
                CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                TmaWgradArgs t;
                memset(&t, 0, sizeof(t));
                t.A_hi = Hhl; t.A_lo = Hhl + plH; t.lda = 256; t.Mo = 256;
                t.B0_hi = dYhl; t.B0_lo = dYhl + plX; t.ldb0 = 64; t.nb0 = 64;
                t.C0 = grads + po[8]; t.ldc0 = 256; t.transpose0 = 1; t.P = PTi; t.scale = 1.f;
                CK(launch_gemm_tma_tn(t, sp, st)); ++nl;
            }
            // BPTT: activated gates (G) + dH -> d(pre-activations) written as operand planes
            CK(launch_lstm_bwd(v.rec, G, at<float>(ws, l.Cst[pp]), dH, dpk + 65536 + 131072, m, sp, st, dGhl, sp ? dGhl + plG : nullptr)); ++nl;
            {   // dX += dG W_ih (K = 1024)
                TmaGemmArgs a = tma_args(dGhl, plG, 1024, v.wiht_hi, v.wiht_lo, 1024, dXs, 64, PTi, 64, 1024);
                a.accumulate = 1;
                CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
            }
            for (int d = 0; d < 2; ++d) {  // dW_ih = dG^T x and dW_hh = dG^T h_prev from ONE pass over this direction's dG
                TmaWgradArgs t;
                memset(&t, 0, sizeof(t));
                t.A_hi = dGhl + d * 512; t.A_lo = dGhl + plG + d * 512; t.lda = 1024; t.Mo = 512;
                t.B0_hi = Xhl; t.B0_lo = Xhl + plX; t.ldb0 = 64; t.nb0 = 64;
                t.B1_hi = Hphl + d * 128; t.B1_lo = Hphl + plH + d * 128; t.ldb1 = 256; t.nb1 = 128;
                t.C0 = dpk + d * 512 * 64; t.ldc0 = 64;
                t.C1 = dpk + 65536 + d * 65536; t.ldc1 = 128;
                t.P = PTi; t.scale = 1.f;
                CK(launch_gemm_tma_tn(t, sp, st)); ++nl;
            }
            continue;
        }
        // out-projection Linear(256 -> 64)
        {
            GemmNtArgs a = nt_args(dY, 64, v.projt_hi, v.projt_lo, 64, 0, dH, 256, PTi, 256, 64);
            CK(gemm_nt(a, sp, st)); ++nl;
            GemmTnArgs t = tn_args(H, 256, dY, 64, grads + po[8], 256, PTi, 256, 64);  // dWp^T = H^T dY, stored transposed
            if (g_backend == 1 && gemm_tn_tc5_supported(t)) {
                CK(launch_gemm_tn_tc5(t, sp, 1, st)); ++nl;
            } else {
                GemmTnArgs t2 = tn_args(dY, 64, H, 256, grads + po[8], 256, PTi, 64, 256);
                CK(launch_gemm_tn(t2, sp, st)); ++nl;
            }
            CK(launch_colsum(dY, 64, PTi, 64, 1.f, grads + po[9], nullptr, st)); ++nl;
        }
        // BPTT: G (activated gates) -> d(pre-activations)
        CK(launch_lstm_bwd(v.rec, G, at<float>(ws, l.Cst[pp]), dH, dpk + 65536 + 131072, m, sp, st)); ++nl;
        // input projection: dX += dG W_ih ; dW_ih += dG^T X ; dW_hh += dG^T h_prev ; db += colsum(dG)   (packed row order)
        {
            GemmNtArgs a = nt_args(G, 1024, v.wiht_hi, v.wiht_lo, 1024, 0, dXs, 64, PTi, 64, 1024);
            a.accumulate = 1;
            CK(gemm_nt(a, sp, st)); ++nl;
            GemmTnArgs t = tn_args(G, 1024, X, 64, dpk, 64, PTi, 1024, 64);
            CK(gemm_tn(t, sp, st)); ++nl;
            for (int d = 0; d < 2; ++d) {
                GemmTnArgs r = tn_args(G + d * 512, 1024, H + d * 128, 256, dpk + 65536 + d * 65536, 128, PTi, 512, 128);
                r.shift = (d == 0 ? -1 : 1) * (int)m.s_t;
                r.tdiv = (pp & 1) ? g.K : 1;
                r.tmod = m.len;
                CK(gemm_tn(r, sp, st)); ++nl;
            }
        }
    }
    // ---- segmentation backward = overlap-add ; bottleneck conv ; bottleneck GroupNorm ; encoder
    float* dFb = at<float>(ws, l.dF2);
    CK(launch_overlap_add_cl(dXs, dFb, B, g.L, g.K, g.Sc, 64, st)); ++nl;
    float* dEn = at<float>(ws, l.dZ);
    {
        GemmNtArgs a = nt_args(dFb, 64, whi + o[3], wlo + o[3], 64, 1, dEn, 64, BLi, 64, 64);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
        GemmTnArgs t = tn_args(dFb, 64, at<float>(ws, l.En), 64, grads + o[3], 64, BLi, 64, 64);
        CK(launch_gemm_tn(t, sp, st)); ++nl;
    }
    CK(launch_gn_bwd_reduce(dEn, at<float>(ws, l.E), at<float>(ws, l.mrE), params + o[1], g.BL, g.L, 64, at<double>(ws, l.redE), grads + o[1],
                            grads + o[2], st)); ++nl;
    CK(launch_gn_bwd_apply(dEn, at<float>(ws, l.E), at<float>(ws, l.dtmp), at<float>(ws, l.mrE), at<double>(ws, l.redE), params + o[1], g.BL,
                           g.L, 64, st)); ++nl;
    CK(launch_axpy(at<float>(ws, l.dE), at<float>(ws, l.dtmp), 1.f, g.BL * 64, st)); ++nl;
    {   // encoder weight gradient: dW_enc += dE^T frames(xp)
        GemmTnArgs t = tn_args(at<float>(ws, l.dE), 64, at<float>(ws, l.xp), stride, grads + o[0], win, BLi, 64, win);
        t.b_rpb = g.L; t.b_skip = 1;
        CK(launch_gemm_tn(t, sp, st)); ++nl;
    }
    // packed LSTM gradients -> natural parameter layout
    for (int pp = 0; pp < h->npath; ++pp) {
        const int64_t* po = o + DP_TASNET_HEAD_PARAMS + h->ppath * pp;
        const float* dpk = at<float>(ws, l.dpack[pp]);
        float* wi[2] = {grads + po[0], grads + po[4]};
        float* wh[2] = {grads + po[1], grads + po[5]};
        float* bi[2] = {grads + po[2], grads + po[6]};
        float* bh[2] = {grads + po[3], grads + po[7]};
        CK(launch_unpack_lstm_grads(dpk, dpk + 65536, dpk + 65536 + 131072, wi, wh, bi, bh, st)); ++nl;
    }
    h->launches = nl;
    return 0;
}

}  // extern "C"
