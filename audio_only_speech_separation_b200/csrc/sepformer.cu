// SepFormer engine: host-side orchestration of the kernels for look2hear/models/sepformer.py (Sepformer.forward
// :986-1016, Dual_Path_Model.forward :706-760, Dual_Computation_Block.forward :600-642, TransformerBlock :541-556,
// TransformerEncoderLayer :320-370).  Inference path (forward); the C-ABI is declared in include/dualpath_b200.h.
//
// Data layout in HBM (fp32, channels-last, N = encoder_out_nchannels):
//   frames   E, En, Fb, F2 [B*L, N] ; Zs [B*L, N*spk] == [B*L*spk, N] ; Gt, Mk, Mx [B*L*spk, N] ; D [B*spk*L, win]
//   stream   X, R, U [B, S, K, N]  position p = (b*S + s)*K + k ; intra sequences walk k, inter sequences walk s, so the
//            reference's permute().contiguous() pairs around every transformer block (sepformer.py:618-637) do not exist
//   per layer scratch QKV [P, 3N], O [P, N], Hf [P, d_ffn]
#include <new>
#include <vector>

#include "../../include/dualpath_b200.h"
#include "common.cuh"
#include "engine_common.h"
#include "kernels.h"

using namespace dp;

#include "sepformer_common.h"

namespace {

struct SLayout {
    size_t xp, E, En, Fb, X, R, U, QKV, O, Hf, F2, Zs, Gt, Mk, Mx, D, stats, mr, total;
    size_t Uhl, Ohl, Hfhl;  // bf16 hi/lo operand planes of the TMA-fed GEMMs ([hi | lo])
};
void sep_layout(const dp_sepformer* h, const SGeo& g, SLayout& l) {
    Carver c;
    const size_t f = sizeof(float);
    const int N = h->cfg.enc_dim, spk = h->cfg.num_spk;
    const int dmax = h->cfg.intra_dffn > h->cfg.inter_dffn ? h->cfg.intra_dffn : h->cfg.inter_dffn;
    l.xp = c.take((size_t)g.B * g.Tp8 * f);
    l.E = c.take(g.BL * N * f);
    l.En = c.take(g.BL * N * f);
    l.Fb = c.take(g.BL * N * f);
    l.X = c.take(g.PT * N * f);
    l.R = c.take(g.PT * N * f);
    l.U = c.take(g.PT * N * f);
    l.QKV = c.take(g.PT * 3 * N * f);
    l.O = c.take(g.PT * N * f);
    l.Hf = c.take(g.PT * dmax * f);
    l.F2 = c.take(g.BL * N * f);
    l.Zs = c.take(g.BL * N * spk * f);
    l.Gt = c.take(g.BL * N * spk * f);
    l.Mk = c.take(g.BL * N * spk * f);
    l.Mx = c.take(g.BL * N * spk * f);
    l.D = c.take(g.BL * spk * h->cfg.win * f);
    l.stats = c.take(2 * g.B * sizeof(double));
    l.mr = c.take(2 * g.B * f);
    l.Uhl = c.take(g.PT * N * 2 * 2);
    l.Ohl = c.take(g.PT * N * 2 * 2);
    l.Hfhl = c.take(g.PT * dmax * 2 * 2);
    l.total = c.off;
}

}  // namespace

extern "C" {

int dp_sepformer_create(const dp_sepformer_config* cfg, const int64_t* offsets, int n_offsets, int64_t n_params, dp_sepformer** out) {
    if (!cfg || !offsets || !out) return fail("dp_sepformer_create: null argument");
    if (cfg->enc_dim != 64 && cfg->enc_dim != 128 && cfg->enc_dim != 256)
        return fail("dp_sepformer_create: encoder_out_nchannels must be 64, 128 or 256 (got %d)", cfg->enc_dim);
    if (cfg->win <= 0 || cfg->win % 8) return fail("dp_sepformer_create: encoder_kernel_size must be a positive multiple of 8 (got %d)", cfg->win);
    if (cfg->chunk <= 0 || (cfg->chunk & 1)) return fail("dp_sepformer_create: masknet_chunksize must be even and positive");
    if (cfg->num_blocks < 1 || cfg->intra_layers < 1 || cfg->inter_layers < 1) return fail("dp_sepformer_create: layer counts must be >= 1");
    if (cfg->num_spk < 1 || cfg->num_spk > 4) return fail("dp_sepformer_create: masknet_numspks must be in 1..4");
    for (int hd : {cfg->intra_heads, cfg->inter_heads})
        if (hd <= 0 || cfg->enc_dim % hd || (cfg->enc_dim / hd != 16 && cfg->enc_dim / hd != 32))
            return fail("dp_sepformer_create: head width (channels / nhead) must be 16 or 32 (channels %d, nhead %d)", cfg->enc_dim, hd);
    for (int df : {cfg->intra_dffn, cfg->inter_dffn})
        if (df <= 0 || df % 64) return fail("dp_sepformer_create: d_ffn must be a positive multiple of 64 (got %d)", df);
    const int need = HEAD + cfg->num_blocks * (path_entries(cfg->intra_layers) + path_entries(cfg->inter_layers));
    if (n_offsets != need) return fail("dp_sepformer_create: expected %d parameter offsets, got %d", need, n_offsets);
    for (int i = 0; i < n_offsets; ++i) {
        if (offsets[i] == -1) continue;  // optional entries (positional-encoding table when unused)
        if (offsets[i] < 0 || offsets[i] >= n_params) return fail("dp_sepformer_create: offset %d out of range", i);
        if (offsets[i] & 7) return fail("dp_sepformer_create: parameter %d must start at a multiple of 8 elements in the flat buffer", i);
    }
    dp_sepformer* h = new (std::nothrow) dp_sepformer();
    if (!h) return fail("dp_sepformer_create: out of host memory");
    h->cfg = *cfg;
    h->off.assign(offsets, offsets + n_offsets);
    h->n_params = n_params;
    h->launches = 0;
    *out = h;
    return 0;
}
void dp_sepformer_destroy(dp_sepformer* h) { delete h; }
int dp_sepformer_last_launches(const dp_sepformer* h) { return h->launches; }

int64_t dp_sepformer_pack_bytes(const dp_sepformer* h) {
    size_t flat = ((size_t)h->n_params * 2 + 255) & ~(size_t)255;
    return (int64_t)(4 * flat);   // [hi | lo | hi of the transposed layer weights | lo of those], each at the parameter's own offset
}
int64_t dp_sepformer_workspace_bytes(const dp_sepformer* h, int B, int T) {
    SGeo g;
    if (sep_geo(h, B, T, g)) return -1;
    SLayout l;
    sep_layout(h, g, l);
    return (int64_t)l.total;
}
int dp_sepformer_pack(dp_sepformer* h, const float* params, void* pack, void* stream) {
    size_t flat = ((size_t)h->n_params * 2 + 255) & ~(size_t)255;
    char* b = static_cast<char*>(pack);
    CK(launch_split_bf16(params, (__nv_bfloat16*)b, (__nv_bfloat16*)(b + flat), h->n_params, S(stream)));
    // transposed copies of every layer's four weights: the K-major W operand of the input-gradient GEMMs (training, TMA backend)
    __nv_bfloat16* thi = (__nv_bfloat16*)(b + 2 * flat);
    __nv_bfloat16* tlo = (__nv_bfloat16*)(b + 3 * flat);
    const dp_sepformer_config& c = h->cfg;
    const int N = c.enc_dim;
    int idx = HEAD;
    for (int pi = 0; pi < 2 * c.num_blocks; ++pi) {
        const int path = pi & 1;
        const int layers = path ? c.inter_layers : c.intra_layers;
        const int dffn = path ? c.inter_dffn : c.intra_dffn;
        const int64_t* po = h->off.data() + idx;
        idx += path_entries(layers);
        for (int ly = 0; ly < layers; ++ly) {
            const int64_t* lo = po + 1 + PER_LAYER * ly;
            CK(launch_transpose_split(params + lo[0], thi + lo[0], tlo + lo[0], 3 * N, N, S(stream)));   // in_proj  [3N, N]
            CK(launch_transpose_split(params + lo[2], thi + lo[2], tlo + lo[2], N, N, S(stream)));       // out_proj [N, N]
            CK(launch_transpose_split(params + lo[4], thi + lo[4], tlo + lo[4], dffn, N, S(stream)));    // FFN 1    [dffn, N]
            CK(launch_transpose_split(params + lo[6], thi + lo[6], tlo + lo[6], N, dffn, S(stream)));    // FFN 2    [N, dffn]
        }
    }
    return 0;
}

int dp_sepformer_forward(dp_sepformer* h, const float* params, const void* pack, const float* mixture, float* est, void* ws, int B, int T,
                         int precision, void* stream) {
    SGeo g;
    if (sep_geo(h, B, T, g)) return 1;
    SLayout l;
    sep_layout(h, g, l);
    cudaStream_t st = S(stream);
    const bool sp = is_split(precision);
    const dp_sepformer_config& c = h->cfg;
    const int N = c.enc_dim, spk = c.num_spk, win = c.win, stride = win / 2;
    const int64_t* o = h->off.data();
    const size_t flat = ((size_t)h->n_params * 2 + 255) & ~(size_t)255;
    const __nv_bfloat16* whi = reinterpret_cast<const __nv_bfloat16*>(pack);
    const __nv_bfloat16* wlo = reinterpret_cast<const __nv_bfloat16*>(static_cast<const char*>(pack) + flat);
    const int PTi = (int)g.PT, BLi = (int)g.BL;
    int nl = 0;

    float* E = at<float>(ws, l.E);
    float* X = at<float>(ws, l.X);
    float* R = at<float>(ws, l.R);
    float* U = at<float>(ws, l.U);
    float* QKV = at<float>(ws, l.QKV);
    float* Oa = at<float>(ws, l.O);
    float* Hf = at<float>(ws, l.Hf);
    double* stats = at<double>(ws, l.stats);
    float* mr = at<float>(ws, l.mr);

    // ---- encoder: Conv1d(1 -> N, win, stride, no bias, no padding) + ReLU as a GEMM over overlapping frames   sepformer.py:23-40
    float* xp = at<float>(ws, l.xp);
    CK(launch_pad_rows(mixture, xp, B, T, g.Tp8, 0, st)); ++nl;
    CK(cudaMemsetAsync(stats, 0, 2 * B * sizeof(double), st));
    {
        GemmNtArgs a = nt_args(xp, stride, whi + o[0], wlo + o[0], win, 0, E, N, BLi, N, win);
        a.a_rpb = g.L; a.a_skip = g.Tp8 / stride - g.L;
        a.relu = 1;
        a.stats = stats; a.rows_per_group = g.L;
        CK(launch_gemm_nt(a, sp, st)); ++nl;
    }
    // ---- masknet.norm (GroupNorm(1, N, 1e-8)) + masknet.conv1d (1x1, no bias) + segmentation                  :725-731
    CK(launch_gn_finalize(stats, mr, B, (double)g.L * N, 1e-8, st)); ++nl;
    CK(launch_gn_apply(E, nullptr, at<float>(ws, l.En), mr, params + o[1], params + o[2], g.BL, g.L, N, nullptr, nullptr, nullptr, st)); ++nl;
    {
        GemmNtArgs a = nt_args(at<float>(ws, l.En), N, whi + o[3], wlo + o[3], N, 0, at<float>(ws, l.Fb), N, BLi, N, N);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
    }
    CK(launch_segment_cl(at<float>(ws, l.Fb), X, B, g.L, g.K, g.Sc, N, st)); ++nl;

    // ---- dual-path blocks                                                                                      :600-642
    int idx = HEAD;
    for (int j = 0; j < c.num_blocks; ++j) {
        for (int path = 0; path < 2; ++path) {
            const int layers = path ? c.inter_layers : c.intra_layers;
            const int heads = path ? c.inter_heads : c.intra_heads;
            const int dffn = path ? c.inter_dffn : c.intra_dffn;
            const bool pre = (path ? c.inter_norm_before : c.intra_norm_before) != 0;
            const bool use_pe = (path ? c.inter_pe : c.intra_pe) != 0;
            const int64_t* po = o + idx;
            idx += path_entries(layers);
            SeqMap m;
            if (!path) { m.nseq = B * g.Sc; m.len = g.K; m.qdiv = 1 << 30; m.s_hi = 0; m.s_lo = g.K; m.s_t = 1; }
            else       { m.nseq = B * g.K; m.len = g.Sc; m.qdiv = g.K; m.s_hi = (long long)g.Sc * g.K; m.s_lo = 1; m.s_t = g.K; }
            // x + positional encoding (indexed by position along the sequence)                                   :549-553
            if (use_pe) {
                if (po[0] < 0) return fail("dp_sepformer_forward: positional encoding enabled but no pe table given");
                if (m.len > DP_SEPFORMER_PE_LEN) return fail("dp_sepformer_forward: sequence length %d exceeds the positional-encoding table (%d)", m.len, DP_SEPFORMER_PE_LEN);
                CK(launch_add_pe(X, params + po[0], R, g.PT, N, g.K, g.Sc, path, st)); ++nl;
            } else {
                CK(cudaMemcpyAsync(R, X, g.PT * N * sizeof(float), cudaMemcpyDeviceToDevice, st));
            }
            const bool tma = gemm_backend() == 2 && dffn % 64 == 0;
            LstmFusedGeom gm;
            gm.inter = path; gm.len = m.len; gm.nseq = m.nseq; gm.K = g.K; gm.S = g.Sc; gm.B = B;
            const bool tc_attn = tma && attn_tc5_supported(N, heads, gm) && !attn_fwd_prefers_mma(N, heads, m.len, sp);
            __nv_bfloat16* Uh = at<__nv_bfloat16>(ws, l.Uhl);
            __nv_bfloat16* Ul = sp ? Uh + g.PT * N : nullptr;
            __nv_bfloat16* Oh = at<__nv_bfloat16>(ws, l.Ohl);
            __nv_bfloat16* Ol = sp ? Oh + g.PT * N : nullptr;
            __nv_bfloat16* Hh = at<__nv_bfloat16>(ws, l.Hfhl);
            __nv_bfloat16* Hl = sp ? Hh + g.PT * dffn : nullptr;
            if (tma && !pre) { CK(launch_split_rows(R, N, Uh, Ul, g.PT, N, 0, st)); ++nl; }  // post-norm: the first GEMM reads the stream itself
            for (int ly = 0; tma && ly < layers; ++ly) {
                // TMA-fed tcgen05 GEMMs: every operand is a pair of bf16 planes written by its producer (gemm_tma.cu)
                const int64_t* lo = po + 1 + PER_LAYER * ly;
                if (pre) { CK(launch_add_ln(R, nullptr, nullptr, nullptr, nullptr, params + lo[8], params + lo[9], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st, Uh, Ul)); ++nl; }
                if (tc_attn) {
                    // QKV leaves the GEMM as operand planes only; attention runs on tcgen05 with S and P in tensor memory
                    __nv_bfloat16* Qh = reinterpret_cast<__nv_bfloat16*>(QKV);
                    __nv_bfloat16* Ql = sp ? Qh + g.PT * 3 * N : nullptr;
                    TmaGemmArgs a = tma_nt_args(Uh, Ul, N, whi + lo[0], wlo + lo[0], N, nullptr, 0, PTi, 3 * N, N);
                    a.C_hi = Qh; a.C_lo = Ql; a.ldch = 3 * N;
                    a.bias = params + lo[1];
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                    CK(launch_attn_fwd_tc5(Qh, Ql, nullptr, Oh, Ol, nullptr, N, heads, gm, sp, st)); ++nl;
                } else {
                    TmaGemmArgs a = tma_nt_args(Uh, Ul, N, whi + lo[0], wlo + lo[0], N, QKV, 3 * N, PTi, 3 * N, N);
                    a.bias = params + lo[1];
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                    if (attn_bwd_mma_supported(N, heads, m)) {   // 257..320 positions: warp-level tensor cores
                        CK(launch_attn_fwd_mma(QKV, nullptr, Oh, Ol, nullptr, N, heads, m, sp, st)); ++nl;
                    } else {
                        CK(launch_attn_fwd(QKV, nullptr, nullptr, N, heads, m, st, Oh, Ol)); ++nl;
                    }
                }
                {
                    TmaGemmArgs a = tma_nt_args(Oh, Ol, N, whi + lo[2], wlo + lo[2], N, pre ? R : U, N, PTi, N, N);
                    a.bias = params + lo[3];
                    a.accumulate = pre ? 1 : 0;
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                if (pre) { CK(launch_add_ln(R, nullptr, nullptr, nullptr, nullptr, params + lo[10], params + lo[11], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st, Uh, Ul)); ++nl; }
                else     { CK(launch_add_ln(U, R, nullptr, R, nullptr, params + lo[8], params + lo[9], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st, Uh, Ul)); ++nl; }
                {   // FFN hidden goes straight to operand planes: the [P, d_ffn] fp32 tensor never exists
                    TmaGemmArgs a = tma_nt_args(Uh, Ul, N, whi + lo[4], wlo + lo[4], N, nullptr, 0, PTi, dffn, N);
                    a.C_hi = Hh; a.C_lo = Hl; a.ldch = dffn;
                    a.bias = params + lo[5];
                    a.act = 1;
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                {
                    TmaGemmArgs a = tma_nt_args(Hh, Hl, dffn, whi + lo[6], wlo + lo[6], dffn, pre ? R : U, N, PTi, N, dffn);
                    a.bias = params + lo[7];
                    a.accumulate = pre ? 1 : 0;
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                if (!pre) { CK(launch_add_ln(U, R, nullptr, R, nullptr, params + lo[10], params + lo[11], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st, Uh, Ul)); ++nl; }
            }
            for (int ly = 0; !tma && ly < layers; ++ly) {                                                        // :320-370
                const int64_t* lo = po + 1 + PER_LAYER * ly;
                const float* src = R;
                if (pre) {
                    CK(launch_add_ln(R, nullptr, nullptr, U, nullptr, params + lo[8], params + lo[9], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st)); ++nl;
                    src = U;
                }
                {
                    GemmNtArgs a = nt_args(src, N, whi + lo[0], wlo + lo[0], N, 0, QKV, 3 * N, PTi, 3 * N, N);
                    a.bias = params + lo[1];
                    CK(gemm_nt(a, sp, st)); ++nl;
                }
                CK(launch_attn_fwd(QKV, Oa, nullptr, N, heads, m, st)); ++nl;
                {
                    GemmNtArgs a = nt_args(Oa, N, whi + lo[2], wlo + lo[2], N, 0, pre ? R : U, N, PTi, N, N);
                    a.bias = params + lo[3];
                    a.accumulate = pre ? 1 : 0;
                    CK(gemm_nt(a, sp, st)); ++nl;
                }
                if (pre) {
                    CK(launch_add_ln(R, nullptr, nullptr, U, nullptr, params + lo[10], params + lo[11], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st)); ++nl;
                    src = U;
                } else {
                    CK(launch_add_ln(U, R, nullptr, R, nullptr, params + lo[8], params + lo[9], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st)); ++nl;
                    src = R;
                }
                {
                    GemmNtArgs a = nt_args(src, N, whi + lo[4], wlo + lo[4], N, 0, Hf, dffn, PTi, dffn, N);
                    a.bias = params + lo[5];
                    a.relu = 1;
                    CK(gemm_nt(a, sp, st)); ++nl;
                }
                {
                    GemmNtArgs a = nt_args(Hf, dffn, whi + lo[6], wlo + lo[6], dffn, 0, pre ? R : U, N, PTi, N, dffn);
                    a.bias = params + lo[7];
                    a.accumulate = pre ? 1 : 0;
                    CK(gemm_nt(a, sp, st)); ++nl;
                }
                if (!pre) {
                    CK(launch_add_ln(U, R, nullptr, R, nullptr, params + lo[10], params + lo[11], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st)); ++nl;
                }
            }
            // final LayerNorm of the encoder (:465), then gLN over (N, K, S) per utterance + residual            :624-627,637-640
            const int64_t* fo = po + 1 + PER_LAYER * layers;
            CK(launch_add_ln(R, nullptr, nullptr, U, nullptr, params + fo[0], params + fo[1], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st)); ++nl;
            CK(cudaMemsetAsync(stats, 0, 2 * B * sizeof(double), st));
            CK(launch_group_stats(U, g.PT, g.P, N, stats, st)); ++nl;
            CK(launch_gn_finalize(stats, mr, B, (double)g.P * N, 1e-8, st)); ++nl;
            CK(launch_gn_apply(U, X, X, mr, params + fo[2], params + fo[3], g.PT, g.P, N, nullptr, nullptr, nullptr, st)); ++nl;
        }
    }

    // ---- PReLU, conv2d (1x1, N -> N*spk), overlap-add.  The 1x1 conv is linear, so it is applied after the overlap-add
    //      (W(a + b) + 2 bias): half the rows and no [P, N*spk] tensor                                            :736-746
    CK(launch_prelu(X, U, g.PT * N, params + o[4], st)); ++nl;
    CK(launch_overlap_add_cl(U, at<float>(ws, l.F2), B, g.L, g.K, g.Sc, N, st)); ++nl;
    float* Zs = at<float>(ws, l.Zs);
    float* Gt = at<float>(ws, l.Gt);
    float* Mk = at<float>(ws, l.Mk);
    float* Mx = at<float>(ws, l.Mx);
    float* D = at<float>(ws, l.D);
    {
        GemmNtArgs a = nt_args(at<float>(ws, l.F2), N, whi + o[5], wlo + o[5], N, 0, Zs, N * spk, BLi, N * spk, N);
        a.bias = params + o[6]; a.bias_scale = 2.f;
        CK(gemm_nt(a, sp, st)); ++nl;
    }
    // ---- gated output: tanh(conv) * sigmoid(conv), end_conv1x1, ReLU on rows (b, l, spk)                         :747-755
    const int rows = BLi * spk;
    {
        GemmNtArgs a = nt_args(Zs, N, whi + o[7], wlo + o[7], N, 0, Gt, N, rows, N, N);
        a.bias = params + o[8]; a.relu = 2;
        CK(launch_gemm_nt(a, sp, st)); ++nl;
        GemmNtArgs b2 = nt_args(Zs, N, whi + o[9], wlo + o[9], N, 0, Gt, N, rows, N, N);
        b2.bias = params + o[10]; b2.relu = 3; b2.mul_c = 1;
        CK(launch_gemm_nt(b2, sp, st)); ++nl;
        GemmNtArgs e2 = nt_args(Gt, N, whi + o[11], wlo + o[11], N, 0, Mk, N, rows, N, N);
        e2.relu = 1;
        CK(launch_gemm_nt(e2, sp, st)); ++nl;
    }
    // ---- mask * encoder output, decoder ConvTranspose1d(N -> 1, win, stride), pad / trim to T                   :997-1012
    CK(launch_mask_apply(Mk, E, Mx, B, g.L, spk, N, st)); ++nl;
    {
        GemmNtArgs a = nt_args(Mx, N, whi + o[12], wlo + o[12], win, 1, D, win, rows, win, N);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
    }
    // rows of the decoder are ordered (spk, b) in the reference and then reshaped as (b, spk): reproduced
    CK(launch_dec_ola_general(D, est, B, spk, g.L, win, T, 1, st)); ++nl;
    h->launches = nl;
    return 0;
}

}  // extern "C"
