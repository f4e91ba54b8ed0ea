// SepFormer training engine: forward that keeps what the backward needs, and the backward pass into the flat gradient buffer
// (look2hear/models/sepformer.py:986-1016 + autograd; pre-norm layers, dropout inactive -- the reference's dropout sites
// (sepformer.py:124-128,261,318-319) are not applied, see models/sepformer.py).  Same data layout and kernels as the inference
// engine (sepformer.cu); the contractions run on the mma.sync GEMMs on fp32 operands (bf16x3 in fp32-parity mode), attention
// forward/backward on the exact streaming-softmax kernels, so every tensor the backward reads is saved in fp32.
#include <cstring>
#include <new>
#include <vector>

#include "../../include/dualpath_b200.h"
#include "common.cuh"
#include "engine_common.h"
#include "kernels.h"

using namespace dp;

#include "sepformer_common.h"

namespace {

struct TLayer { size_t Rin, U1, QKV, LSE, Oa, Rmid, U2, Hf, U1hl, Oahl, U2hl, Hfhl; };
struct TPath { size_t Rfin, Uf, mr; int first_layer, layers; };
struct TLayout {
    size_t xp, E, En, Fb, statsE, mrE, stats;
    std::vector<size_t> X;       // stream at every path boundary
    std::vector<TLayer> layer;
    std::vector<TPath> path;
    size_t Upre, F2, Zs, T1, T2, Gt, Mk, Mx, D;
    // backward scratch
    size_t dX, dR, dU, dHf, dOa, dQKV, dD, dMx, dMk, dE, dGt, dA, dB, dZs, dF2, dEn, red;
    size_t QKVhl, dRhl, dHfhl, dQKVhl;   // operand planes (hi | lo) shared by all layers (TMA backend)
    size_t total;
};

bool train_tma(int N, int dffn);

void train_layout(const dp_sepformer* h, const SGeo& g, TLayout& l) {
    Carver c;
    const size_t f = sizeof(float);
    const dp_sepformer_config& cf = h->cfg;
    const int N = cf.enc_dim, spk = cf.num_spk;
    const int dmax = cf.intra_dffn > cf.inter_dffn ? cf.intra_dffn : cf.inter_dffn;
    const int hmax = cf.intra_heads > cf.inter_heads ? cf.intra_heads : cf.inter_heads;
    l.xp = c.take((size_t)g.B * g.Tp8 * f);
    l.E = c.take(g.BL * N * f);
    l.En = c.take(g.BL * N * f);
    l.Fb = c.take(g.BL * N * f);
    l.statsE = c.take(2 * g.B * sizeof(double));
    l.mrE = c.take(2 * g.B * f);
    l.stats = c.take(2 * g.B * sizeof(double));
    const int npaths = 2 * cf.num_blocks;
    l.X.resize(npaths + 1);
    for (int p = 0; p <= npaths; ++p) l.X[p] = c.take(g.PT * N * f);
    l.path.resize(npaths);
    int li = 0;
    for (int p = 0; p < npaths; ++p) {
        const int layers = (p & 1) ? cf.inter_layers : cf.intra_layers;
        const int dffn = (p & 1) ? cf.inter_dffn : cf.intra_dffn;
        l.path[p].first_layer = li;
        l.path[p].layers = layers;
        for (int y = 0; y < layers; ++y, ++li) {
            TLayer t;
            t.Rin = c.take(g.PT * N * f);
            t.U1 = c.take(g.PT * N * f);
            t.QKV = c.take(g.PT * 3 * N * f);
            t.LSE = c.take(g.PT * hmax * f);
            t.Oa = c.take(g.PT * N * f);
            t.Rmid = c.take(g.PT * N * f);
            t.U2 = c.take(g.PT * N * f);
            // the fp32 FFN hidden exists on the mma.sync engine only (the TMA engine keeps it as planes): 133 MB per layer at 16 s
            t.Hf = train_tma(N, dffn) ? t.U2 : c.take(g.PT * dffn * f);
            // operand planes (hi | lo) of what the backward's weight-gradient GEMMs read (TMA backend)
            t.U1hl = c.take(g.PT * N * 4);
            t.Oahl = c.take(g.PT * N * 4);
            t.U2hl = c.take(g.PT * N * 4);
            t.Hfhl = c.take(g.PT * dffn * 4);
            l.layer.push_back(t);
        }
        l.path[p].Rfin = c.take(g.PT * N * f);
        l.path[p].Uf = c.take(g.PT * N * f);
        l.path[p].mr = c.take(2 * g.B * f);
    }
    l.Upre = c.take(g.PT * N * f);
    l.F2 = c.take(g.BL * N * f);
    l.Zs = c.take(g.BL * N * spk * f);
    l.T1 = c.take(g.BL * N * spk * f);
    l.T2 = c.take(g.BL * N * spk * f);
    l.Gt = c.take(g.BL * N * spk * f);
    l.Mk = c.take(g.BL * N * spk * f);
    l.Mx = c.take(g.BL * N * spk * f);
    l.D = c.take(g.BL * spk * cf.win * f);
    l.dX = c.take(g.PT * N * f);
    l.dR = c.take(g.PT * N * f);
    l.dU = c.take(g.PT * N * f);
    l.dHf = (train_tma(N, cf.intra_dffn) && train_tma(N, cf.inter_dffn)) ? l.dU : c.take(g.PT * dmax * f);   // planes only on the TMA engine
    l.dOa = c.take(g.PT * N * f);
    l.dQKV = c.take(g.PT * 3 * N * f);
    l.dD = c.take(g.BL * spk * cf.win * f);
    l.dMx = c.take(g.BL * N * spk * f);
    l.dMk = c.take(g.BL * N * spk * f);
    l.dE = c.take(g.BL * N * f);
    l.dGt = c.take(g.BL * N * spk * f);
    l.dA = c.take(g.BL * N * spk * f);
    l.dB = c.take(g.BL * N * spk * f);
    l.dZs = c.take(g.BL * N * spk * f);
    l.dF2 = c.take(g.BL * N * f);
    l.dEn = c.take(g.BL * N * f);
    l.red = c.take(2 * g.B * sizeof(double));
    l.QKVhl = c.take(g.PT * 3 * N * 4);
    l.dRhl = c.take(g.PT * N * 4);
    l.dHfhl = c.take(g.PT * dmax * 4);
    l.dQKVhl = c.take(g.PT * 3 * N * 4);
    l.total = c.off;
}

SeqMap seq_map(const SGeo& g, int B, int path) {
    SeqMap m;
    if (!path) { m.nseq = B * g.Sc; m.len = g.K; m.qdiv = 1 << 30; m.s_hi = 0; m.s_lo = g.K; m.s_t = 1; }
    else       { m.nseq = B * g.K; m.len = g.Sc; m.qdiv = g.K; m.s_hi = (long long)g.Sc * g.K; m.s_lo = 1; m.s_t = g.K; }
    return m;
}

// TMA-fed tcgen05 GEMMs for a path's layers (forward and backward must agree: they exchange operand planes): the weight-gradient
// kernel wants its output rows in blocks of 128 and at most 256 output columns
bool train_tma(int N, int dffn) { return gemm_backend() == 2 && N % 128 == 0 && N <= 256 && dffn % 128 == 0; }

int check_trainable(const dp_sepformer* h) {
    const dp_sepformer_config& c = h->cfg;
    if (c.intra_norm_before && c.inter_norm_before) return 0;
    const bool tma_all = train_tma(c.enc_dim, c.intra_dffn) && train_tma(c.enc_dim, c.inter_dffn);
    if (!tma_all)
        return fail("SepFormer training with post-norm layers needs the TMA backend (enc_dim in {128, 256}, d_ffn %% 128 == 0); the mma.sync "
                    "engine covers pre-norm layers (configs/sepformer_base.yml)");
    return 0;
}

}  // namespace

extern "C" {

int dp_sepformer_set_dropout(dp_sepformer* h, float p, uint32_t seed) {
    if (!(p >= 0.f && p < 1.f)) return fail("dp_sepformer_set_dropout: need 0 <= p < 1 (got %g)", (double)p);
    h->drop_p = p;
    h->drop_seed = seed;
    return 0;
}

int64_t dp_sepformer_train_workspace_bytes(const dp_sepformer* h, int B, int T) {
    SGeo g;
    if (sep_geo(h, B, T, g)) return -1;
    TLayout l;
    train_layout(h, g, l);
    return (int64_t)l.total;
}

int dp_sepformer_forward_train(dp_sepformer* h, const float* params, const void* pack, const float* mixture, float* est, void* ws, int B, int T,
                               int precision, void* stream) {
    if (check_trainable(h)) return 1;
    SGeo g;
    if (sep_geo(h, B, T, g)) return 1;
    TLayout l;
    train_layout(h, g, l);
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
    // dropout of the transformer layers (sepformer.py:124-128,261,318-319): sites 0 attention probabilities, 1 attention output,
    // 2 FFN hidden, 3 FFN output; masks are a pure function of (seed, layer, site, element) and regenerated by the backward
    const unsigned drop_thr = (unsigned)(h->drop_p * 16777216.0f);
    const float drop_scale = 1.0f / (1.0f - h->drop_p);

    float* E = at<float>(ws, l.E);
    double* stats = at<double>(ws, l.stats);
    // ---- encoder + ReLU, masknet.norm, masknet.conv1d, segmentation
    float* xp = at<float>(ws, l.xp);
    CK(launch_pad_rows(mixture, xp, B, T, g.Tp8, 0, st)); ++nl;
    CK(cudaMemsetAsync(at<double>(ws, l.statsE), 0, 2 * B * sizeof(double), st));
    {
        GemmNtArgs a = nt_args(xp, stride, whi + o[0], wlo + o[0], win, 0, E, N, BLi, N, win);
        a.a_rpb = g.L; a.a_skip = g.Tp8 / stride - g.L;
        a.relu = 1;
        a.stats = at<double>(ws, l.statsE); a.rows_per_group = g.L;
        CK(launch_gemm_nt(a, sp, st)); ++nl;
    }
    CK(launch_gn_finalize(at<double>(ws, l.statsE), at<float>(ws, l.mrE), B, (double)g.L * N, 1e-8, st)); ++nl;
    CK(launch_gn_apply(E, nullptr, at<float>(ws, l.En), at<float>(ws, l.mrE), params + o[1], params + o[2], g.BL, g.L, N, nullptr, nullptr, nullptr, st)); ++nl;
    {
        GemmNtArgs a = nt_args(at<float>(ws, l.En), N, whi + o[3], wlo + o[3], N, 0, at<float>(ws, l.Fb), N, BLi, N, N);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
    }
    CK(launch_segment_cl(at<float>(ws, l.Fb), at<float>(ws, l.X[0]), B, g.L, g.K, g.Sc, N, st)); ++nl;

    // ---- dual-path blocks
    int idx = HEAD;
    const int npaths = 2 * c.num_blocks;
    for (int pi = 0; pi < npaths; ++pi) {
        const int path = pi & 1;
        const int layers = path ? c.inter_layers : c.intra_layers;
        const int heads = path ? c.inter_heads : c.intra_heads;
        const int dffn = path ? c.inter_dffn : c.intra_dffn;
        const bool use_pe = (path ? c.inter_pe : c.intra_pe) != 0;
        const int64_t* po = o + idx;
        idx += path_entries(layers);
        const SeqMap m = seq_map(g, B, path);
        const TPath& tp = l.path[pi];
        const bool tma = train_tma(N, dffn);
        const bool pre = (path ? c.inter_norm_before : c.intra_norm_before) != 0;
        LstmFusedGeom gm;
        gm.inter = path; gm.len = m.len; gm.nseq = m.nseq; gm.K = g.K; gm.S = g.Sc; gm.B = B;
        const bool tc_attn = tma && attn_tc5_supported(N, heads, gm) && !attn_fwd_prefers_mma(N, heads, m.len, sp);
        if (drop_thr && !tma)
            return fail("dp_sepformer_forward_train: dropout needs the TMA backend (dp_set_gemm_backend(2)), enc_dim in {128, 256} and d_ffn %% 128 == 0");
        float* Xin = at<float>(ws, l.X[pi]);
        float* R0 = at<float>(ws, l.layer[tp.first_layer].Rin);
        if (use_pe) {
            if (po[0] < 0) return fail("dp_sepformer_forward_train: positional encoding enabled but no pe table given");
            if (m.len > DP_SEPFORMER_PE_LEN) return fail("dp_sepformer_forward_train: sequence length %d exceeds the positional-encoding table", m.len);
            CK(launch_add_pe(Xin, params + po[0], R0, g.PT, N, g.K, g.Sc, path, st)); ++nl;
        } else {
            CK(cudaMemcpyAsync(R0, Xin, g.PT * N * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
        for (int ly = 0; ly < layers; ++ly) {
            const int64_t* lo = po + 1 + PER_LAYER * ly;
            const TLayer& t = l.layer[tp.first_layer + ly];
            float* Rin = at<float>(ws, t.Rin);
            float* U1 = at<float>(ws, t.U1);
            float* QKV = at<float>(ws, t.QKV);
            float* Oa = at<float>(ws, t.Oa);
            float* Rmid = at<float>(ws, t.Rmid);
            float* U2 = at<float>(ws, t.U2);
            float* Hf = at<float>(ws, t.Hf);
            float* Rout = (ly + 1 < layers) ? at<float>(ws, l.layer[tp.first_layer + ly + 1].Rin) : at<float>(ws, tp.Rfin);
            if (tma) {
                // TMA-fed tcgen05 GEMMs on operand planes (the inference engine's layer), keeping in fp32 only what the backward reads in
                // fp32 (Rin, QKV, O, LSE, Rmid, the FFN hidden as the ReLU mask) and as planes what its weight-gradient GEMMs read
                const long long pN = g.PT * N, pD = g.PT * dffn, pQ = g.PT * 3 * N;
                __nv_bfloat16* U1h = at<__nv_bfloat16>(ws, t.U1hl);
                __nv_bfloat16* Oh = at<__nv_bfloat16>(ws, t.Oahl);
                __nv_bfloat16* U2h = at<__nv_bfloat16>(ws, t.U2hl);
                __nv_bfloat16* Hh = at<__nv_bfloat16>(ws, t.Hfhl);
                __nv_bfloat16* Qh = at<__nv_bfloat16>(ws, l.QKVhl);
                if (pre) {
                    CK(launch_add_ln(Rin, nullptr, nullptr, nullptr, nullptr, params + lo[8], params + lo[9], g.PT, N, 1e-6f, nullptr, nullptr, nullptr,
                                     st, U1h, sp ? U1h + pN : nullptr)); ++nl;
                } else if (ly == 0) {   // post-norm: the first GEMM reads the stream itself (later layers get these planes from the previous norm2)
                    CK(launch_split_rows(Rin, N, U1h, sp ? U1h + pN : nullptr, g.PT, N, 0, st)); ++nl;
                }
                {
                    TmaGemmArgs a = tma_nt_args(U1h, sp ? U1h + pN : nullptr, N, whi + lo[0], wlo + lo[0], N, QKV, 3 * N, PTi, 3 * N, N);
                    if (tc_attn) { a.C_hi = Qh; a.C_lo = sp ? Qh + pQ : nullptr; a.ldch = 3 * N; }
                    a.bias = params + lo[1];
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                const int li = tp.first_layer + ly;
                if (tc_attn) {
                    CK(launch_attn_fwd_tc5(Qh, sp ? Qh + pQ : nullptr, Oa, Oh, sp ? Oh + pN : nullptr, at<float>(ws, t.LSE), N, heads, gm, sp, st,
                                           drop_thr, drop_site_key(h->drop_seed, li, 0), drop_scale)); ++nl;
                } else {
                    // sequences beyond the tcgen05 kernel's 256 positions (16 s at 16 kHz: 258 chunks): exact CUDA-core kernel, same dropout masks
                    if (attn_bwd_mma_supported(N, heads, m)) {
                        CK(launch_attn_fwd_mma(QKV, Oa, Oh, sp ? Oh + pN : nullptr, at<float>(ws, t.LSE), N, heads, m, sp, st, drop_thr,
                                               drop_site_key(h->drop_seed, li, 0), drop_scale)); ++nl;
                    } else {
                        CK(launch_attn_fwd(QKV, Oa, at<float>(ws, t.LSE), N, heads, m, st, Oh, sp ? Oh + pN : nullptr, drop_thr,
                                           drop_site_key(h->drop_seed, li, 0), drop_scale)); ++nl;
                    }
                }
                {   // Rmid = Rin + O W_o^T + b_o
                    TmaGemmArgs a = tma_nt_args(Oh, sp ? Oh + pN : nullptr, N, whi + lo[2], wlo + lo[2], N, Rmid, N, PTi, N, N);
                    a.bias = params + lo[3]; a.accumulate = 1; a.Cin = Rin;
                    a.drop_thr = drop_thr; a.drop_key = drop_site_key(h->drop_seed, li, 1); a.drop_scale = drop_scale;
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                if (pre) {
                    CK(launch_add_ln(Rmid, nullptr, nullptr, nullptr, nullptr, params + lo[10], params + lo[11], g.PT, N, 1e-6f, nullptr, nullptr, nullptr,
                                     st, U2h, sp ? U2h + pN : nullptr)); ++nl;
                } else {   // post-norm: Rmid holds z1 = x + dropout(attn); r1 = norm1(z1) as fp32 (residual of the FFN) and planes
                    CK(launch_add_ln(Rmid, nullptr, nullptr, U1, nullptr, params + lo[8], params + lo[9], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st,
                                     U2h, sp ? U2h + pN : nullptr)); ++nl;
                }
                {
                    // the FFN hidden exists as operand planes only: they feed FFN 2, the dW2 GEMM and (hi plane) the ReLU mask of the backward
                    TmaGemmArgs a = tma_nt_args(U2h, sp ? U2h + pN : nullptr, N, whi + lo[4], wlo + lo[4], N, nullptr, 0, PTi, dffn, N);
                    a.C_hi = Hh; a.C_lo = sp ? Hh + pD : nullptr; a.ldch = dffn;
                    a.bias = params + lo[5]; a.act = 1;
                    a.drop_thr = drop_thr; a.drop_key = drop_site_key(h->drop_seed, li, 2); a.drop_scale = drop_scale;
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                {   // pre: Rout = Rmid + Hf W_2^T + b_2 ; post: z2 (U2 buffer) = r1 + Hf W_2^T + b_2, Rout = norm2(z2)
                    TmaGemmArgs a = tma_nt_args(Hh, sp ? Hh + pD : nullptr, dffn, whi + lo[6], wlo + lo[6], dffn, pre ? Rout : U2, N, PTi, N, dffn);
                    a.bias = params + lo[7]; a.accumulate = 1; a.Cin = pre ? Rmid : U1;
                    a.drop_thr = drop_thr; a.drop_key = drop_site_key(h->drop_seed, li, 3); a.drop_scale = drop_scale;
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                if (!pre) {   // the next layer's first GEMM reads these planes
                    __nv_bfloat16* nh = (ly + 1 < layers) ? at<__nv_bfloat16>(ws, l.layer[tp.first_layer + ly + 1].U1hl) : nullptr;
                    CK(launch_add_ln(U2, nullptr, nullptr, Rout, nullptr, params + lo[10], params + lo[11], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st,
                                     nh, (nh && sp) ? nh + pN : nullptr)); ++nl;
                }
                continue;
            }
            CK(launch_add_ln(Rin, nullptr, nullptr, U1, nullptr, params + lo[8], params + lo[9], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st)); ++nl;
            {
                GemmNtArgs a = nt_args(U1, N, whi + lo[0], wlo + lo[0], N, 0, QKV, 3 * N, PTi, 3 * N, N);
                a.bias = params + lo[1];
                CK(launch_gemm_nt(a, sp, st)); ++nl;
            }
            CK(launch_attn_fwd(QKV, Oa, at<float>(ws, t.LSE), N, heads, m, st)); ++nl;
            {   // Rmid = Rin + O W_o^T + b_o
                CK(cudaMemcpyAsync(Rmid, Rin, g.PT * N * sizeof(float), cudaMemcpyDeviceToDevice, st));
                GemmNtArgs a = nt_args(Oa, N, whi + lo[2], wlo + lo[2], N, 0, Rmid, N, PTi, N, N);
                a.bias = params + lo[3]; a.accumulate = 1;
                CK(launch_gemm_nt(a, sp, st)); ++nl;
            }
            CK(launch_add_ln(Rmid, nullptr, nullptr, U2, nullptr, params + lo[10], params + lo[11], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st)); ++nl;
            {
                GemmNtArgs a = nt_args(U2, N, whi + lo[4], wlo + lo[4], N, 0, Hf, dffn, PTi, dffn, N);
                a.bias = params + lo[5]; a.relu = 1;
                CK(launch_gemm_nt(a, sp, st)); ++nl;
            }
            {   // Rout = Rmid + Hf W_2^T + b_2
                CK(cudaMemcpyAsync(Rout, Rmid, g.PT * N * sizeof(float), cudaMemcpyDeviceToDevice, st));
                GemmNtArgs a = nt_args(Hf, dffn, whi + lo[6], wlo + lo[6], dffn, 0, Rout, N, PTi, N, dffn);
                a.bias = params + lo[7]; a.accumulate = 1;
                CK(launch_gemm_nt(a, sp, st)); ++nl;
            }
        }
        const int64_t* fo = po + 1 + PER_LAYER * layers;
        float* Uf = at<float>(ws, tp.Uf);
        CK(launch_add_ln(at<float>(ws, tp.Rfin), nullptr, nullptr, Uf, nullptr, params + fo[0], params + fo[1], g.PT, N, 1e-6f, nullptr, nullptr, nullptr, st)); ++nl;
        CK(cudaMemsetAsync(stats, 0, 2 * B * sizeof(double), st));
        CK(launch_group_stats(Uf, g.PT, g.P, N, stats, st)); ++nl;
        CK(launch_gn_finalize(stats, at<float>(ws, tp.mr), B, (double)g.P * N, 1e-8, st)); ++nl;
        CK(launch_gn_apply(Uf, Xin, at<float>(ws, l.X[pi + 1]), at<float>(ws, tp.mr), params + fo[2], params + fo[3], g.PT, g.P, N, nullptr, nullptr, nullptr, st)); ++nl;
    }

    // ---- tail: PReLU, overlap-add, conv2d (after the overlap-add, bias twice), gate, end conv + ReLU, mask, decoder
    float* Xl = at<float>(ws, l.X[npaths]);
    CK(launch_prelu(Xl, at<float>(ws, l.Upre), g.PT * N, params + o[4], st)); ++nl;
    CK(launch_overlap_add_cl(at<float>(ws, l.Upre), at<float>(ws, l.F2), B, g.L, g.K, g.Sc, N, st)); ++nl;
    float* Zs = at<float>(ws, l.Zs);
    float* T1 = at<float>(ws, l.T1);
    float* T2 = at<float>(ws, l.T2);
    float* Gt = at<float>(ws, l.Gt);
    float* Mk = at<float>(ws, l.Mk);
    float* Mx = at<float>(ws, l.Mx);
    float* D = at<float>(ws, l.D);
    const int rows = BLi * spk;
    {
        GemmNtArgs a = nt_args(at<float>(ws, l.F2), N, whi + o[5], wlo + o[5], N, 0, Zs, N * spk, BLi, N * spk, N);
        a.bias = params + o[6]; a.bias_scale = 2.f;
        CK(launch_gemm_nt(a, sp, st)); ++nl;
        GemmNtArgs t1 = nt_args(Zs, N, whi + o[7], wlo + o[7], N, 0, T1, N, rows, N, N);
        t1.bias = params + o[8]; t1.relu = 2;
        CK(launch_gemm_nt(t1, sp, st)); ++nl;
        GemmNtArgs t2 = nt_args(Zs, N, whi + o[9], wlo + o[9], N, 0, T2, N, rows, N, N);
        t2.bias = params + o[10]; t2.relu = 3;
        CK(launch_gemm_nt(t2, sp, st)); ++nl;
        CK(launch_mul(T1, T2, Gt, (long long)rows * N, st)); ++nl;
        GemmNtArgs e2 = nt_args(Gt, N, whi + o[11], wlo + o[11], N, 0, Mk, N, rows, N, N);
        e2.relu = 1;
        CK(launch_gemm_nt(e2, sp, st)); ++nl;
    }
    CK(launch_mask_apply(Mk, E, Mx, B, g.L, spk, N, st)); ++nl;
    {
        GemmNtArgs a = nt_args(Mx, N, whi + o[12], wlo + o[12], win, 1, D, win, rows, win, N);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
    }
    CK(launch_dec_ola_general(D, est, B, spk, g.L, win, T, 1, st)); ++nl;
    h->launches = nl;
    return 0;
}

int dp_sepformer_backward(dp_sepformer* h, const float* params, const void* pack, const float* d_est, float* grads, void* ws, int B, int T,
                          int precision, void* stream) {
    if (check_trainable(h)) return 1;
    SGeo g;
    if (sep_geo(h, B, T, g)) return 1;
    TLayout l;
    train_layout(h, g, l);
    cudaStream_t st = S(stream);
    const bool sp = is_split(precision);
    const dp_sepformer_config& c = h->cfg;
    const int N = c.enc_dim, spk = c.num_spk, win = c.win, stride = win / 2;
    const int64_t* o = h->off.data();
    const size_t flat = ((size_t)h->n_params * 2 + 255) & ~(size_t)255;
    const __nv_bfloat16* whi = reinterpret_cast<const __nv_bfloat16*>(pack);
    const __nv_bfloat16* wlo = reinterpret_cast<const __nv_bfloat16*>(static_cast<const char*>(pack) + flat);
    const __nv_bfloat16* thi = reinterpret_cast<const __nv_bfloat16*>(static_cast<const char*>(pack) + 2 * flat);   // transposed layer weights
    const __nv_bfloat16* tlo = reinterpret_cast<const __nv_bfloat16*>(static_cast<const char*>(pack) + 3 * flat);
    const int PTi = (int)g.PT, BLi = (int)g.BL, rows = BLi * spk;
    const long long nPN = g.PT * N;
    int nl = 0;
    const unsigned drop_thr = (unsigned)(h->drop_p * 16777216.0f);
    const float drop_scale = 1.0f / (1.0f - h->drop_p);

    float* E = at<float>(ws, l.E);
    float* dX = at<float>(ws, l.dX);
    float* dR = at<float>(ws, l.dR);
    float* dU = at<float>(ws, l.dU);
    float* dHf = at<float>(ws, l.dHf);
    float* dOa = at<float>(ws, l.dOa);
    float* dQKV = at<float>(ws, l.dQKV);
    float* dD = at<float>(ws, l.dD);
    float* dMx = at<float>(ws, l.dMx);
    float* dMk = at<float>(ws, l.dMk);
    float* dE = at<float>(ws, l.dE);
    float* dGt = at<float>(ws, l.dGt);
    float* dA = at<float>(ws, l.dA);
    float* dB = at<float>(ws, l.dB);
    float* dZs = at<float>(ws, l.dZs);
    float* dF2 = at<float>(ws, l.dF2);
    float* dEn = at<float>(ws, l.dEn);
    double* red = at<double>(ws, l.red);

    // ---- decoder: d_est -> dD (frames), dMx = dD Wdec^T, dWdec += Mx^T dD
    CK(launch_dec_ola_general_bwd(d_est, dD, B, spk, g.L, win, T, 1, st)); ++nl;
    {
        GemmNtArgs a = nt_args(dD, win, whi + o[12], wlo + o[12], win, 0, dMx, N, rows, N, win);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
        GemmTnArgs t = tn_args(at<float>(ws, l.Mx), N, dD, win, grads + o[12], win, rows, N, win);
        CK(launch_gemm_tn(t, sp, st)); ++nl;
    }
    // ---- mask * encoder (ReLU of the mask folded in): dMk = dMx E [Mk > 0], dE = sum_c dMx Mk
    CK(launch_mask_bwd(dMx, at<float>(ws, l.Mk), E, dMk, dE, 0, B, g.L, spk, N, st)); ++nl;
    // ---- end_conv1x1 (no bias)
    {
        GemmNtArgs a = nt_args(dMk, N, whi + o[11], wlo + o[11], N, 1, dGt, N, rows, N, N);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
        GemmTnArgs t = tn_args(dMk, N, at<float>(ws, l.Gt), N, grads + o[11], N, rows, N, N);
        CK(launch_gemm_tn(t, sp, st)); ++nl;
    }
    // ---- gate: Gt = tanh(a) * sigmoid(b)
    CK(launch_gate_bwd(dGt, at<float>(ws, l.T1), at<float>(ws, l.T2), dA, dB, (long long)rows * N, st)); ++nl;
    {
        float* Zs = at<float>(ws, l.Zs);
        GemmNtArgs a = nt_args(dA, N, whi + o[7], wlo + o[7], N, 1, dZs, N, rows, N, N);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
        GemmNtArgs b2 = nt_args(dB, N, whi + o[9], wlo + o[9], N, 1, dZs, N, rows, N, N);
        b2.accumulate = 1;
        CK(launch_gemm_nt(b2, sp, st)); ++nl;
        GemmTnArgs t = tn_args(dA, N, Zs, N, grads + o[7], N, rows, N, N);
        CK(launch_gemm_tn(t, sp, st)); ++nl;
        GemmTnArgs t2 = tn_args(dB, N, Zs, N, grads + o[9], N, rows, N, N);
        CK(launch_gemm_tn(t2, sp, st)); ++nl;
        CK(launch_colsum_any(dA, N, rows, N, 1.f, grads + o[8], st)); ++nl;
        CK(launch_colsum_any(dB, N, rows, N, 1.f, grads + o[10], st)); ++nl;
    }
    // ---- conv2d applied after the overlap-add (bias counted twice): dZs viewed as [B*L, N*spk]
    {
        GemmNtArgs a = nt_args(dZs, N * spk, whi + o[5], wlo + o[5], N, 1, dF2, N, BLi, N, N * spk);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
        GemmTnArgs t = tn_args(dZs, N * spk, at<float>(ws, l.F2), N, grads + o[5], N, BLi, N * spk, N);
        CK(launch_gemm_tn(t, sp, st)); ++nl;
        CK(launch_colsum_any(dZs, N * spk, BLi, N * spk, 2.f, grads + o[6], st)); ++nl;
    }
    // ---- overlap-add backward = segmentation; PReLU backward
    const int npaths = 2 * c.num_blocks;
    CK(launch_segment_cl(dF2, dX, B, g.L, g.K, g.Sc, N, st)); ++nl;
    CK(launch_prelu_bwd(dX, at<float>(ws, l.X[npaths]), dX, nPN, params + o[4], grads + o[4], st)); ++nl;

    // ---- dual-path blocks, reversed.  dX = gradient of the stream at the path's output.
    std::vector<int> pidx(npaths);
    {
        int idx = HEAD;
        for (int pi = 0; pi < npaths; ++pi) {
            pidx[pi] = idx;
            idx += path_entries((pi & 1) ? c.inter_layers : c.intra_layers);
        }
    }
    for (int pi = npaths - 1; pi >= 0; --pi) {
        const int path = pi & 1;
        const int layers = path ? c.inter_layers : c.intra_layers;
        const int heads = path ? c.inter_heads : c.intra_heads;
        const int dffn = path ? c.inter_dffn : c.intra_dffn;
        const int64_t* po = o + pidx[pi];
        const int64_t* fo = po + 1 + PER_LAYER * layers;
        const SeqMap m = seq_map(g, B, path);
        const TPath& tp = l.path[pi];
        const bool tma = train_tma(N, dffn);
        const bool pre = (path ? c.inter_norm_before : c.intra_norm_before) != 0;
        if (drop_thr && !(tma && attn_bwd_mma_supported(N, heads, m)))
            return fail("dp_sepformer_backward: dropout needs the TMA backend, enc_dim in {128, 256}, d_ffn %% 128 == 0 and sequences <= 320");
        float* Uf = at<float>(ws, tp.Uf);
        float* mr = at<float>(ws, tp.mr);
        // X_out = Xin + gLN(Uf): gLN backward -> dU (d Uf); the residual branch keeps dX
        CK(cudaMemsetAsync(red, 0, 2 * B * sizeof(double), st));
        CK(launch_gn_bwd_reduce_any(dX, Uf, mr, params + fo[2], g.PT, g.P, N, red, grads + fo[2], grads + fo[3], st)); ++nl;
        CK(launch_gn_bwd_apply(dX, Uf, dU, mr, red, params + fo[2], g.PT, g.P, N, st)); ++nl;
        // final LayerNorm of the encoder
        CK(launch_ln_bwd(dU, at<float>(ws, tp.Rfin), dR, nullptr, params + fo[0], g.PT, N, 1e-6f, grads + fo[0], grads + fo[1], st)); ++nl;
        for (int ly = layers - 1; ly >= 0; --ly) {
            const int64_t* lo = po + 1 + PER_LAYER * ly;
            const TLayer& t = l.layer[tp.first_layer + ly];
            if (tma) {
                const long long pN = g.PT * N, pD = g.PT * dffn, pQ = g.PT * 3 * N;
                const __nv_bfloat16* U1h = at<__nv_bfloat16>(ws, t.U1hl);
                const __nv_bfloat16* Oh = at<__nv_bfloat16>(ws, t.Oahl);
                const __nv_bfloat16* U2h = at<__nv_bfloat16>(ws, t.U2hl);
                const __nv_bfloat16* Hh = at<__nv_bfloat16>(ws, t.Hfhl);
                __nv_bfloat16* dRh = at<__nv_bfloat16>(ws, l.dRhl);
                __nv_bfloat16* dHh = at<__nv_bfloat16>(ws, l.dHfhl);
                __nv_bfloat16* dQh = at<__nv_bfloat16>(ws, l.dQKVhl);
                auto wgrad = [&](const __nv_bfloat16* A, long long plA, int lda, int Mo, const __nv_bfloat16* Bm, long long plB, int ldb, int nb,
                                 float* C, int ldc, int transpose) {
                    TmaWgradArgs w;
                    memset(&w, 0, sizeof(w));
                    w.A_hi = A; w.A_lo = A + plA; w.lda = lda; w.Mo = Mo;
                    w.B0_hi = Bm; w.B0_lo = Bm + plB; w.ldb0 = ldb; w.nb0 = nb;
                    w.C0 = C; w.ldc0 = ldc; w.transpose0 = transpose; w.P = PTi; w.scale = 1.f;
                    return launch_gemm_tma_tn(w, sp, st);
                };
                // FFN: out = Rmid + relu(U2 W1^T + b1) W2^T + b2
                const int li = tp.first_layer + ly;
                // post-norm: dR is the gradient of norm2(z2): first through norm2 (dU <- d z2); pre-norm: dR is already d(Rmid + FFN)
                const float* dF = dR;   // gradient of the (dropped) FFN output
                if (!pre) {
                    CK(launch_ln_bwd(dR, at<float>(ws, t.U2), dU, nullptr, params + lo[10], g.PT, N, 1e-6f, grads + lo[10], grads + lo[11], st)); ++nl;
                    dF = dU;
                }
                // planes of dF masked like the forward + db2
                CK(launch_split_rows_colsum(dF, N, dRh, sp ? dRh + pN : nullptr, g.PT, N, grads + lo[7], st, drop_thr, drop_site_key(h->drop_seed, li, 3),
                                            drop_scale)); ++nl;
                {   // dHf = (dF W2) where Hf > 0, as planes
                    TmaGemmArgs a = tma_nt_args(dRh, sp ? dRh + pN : nullptr, N, thi + lo[6], tlo + lo[6], N, nullptr, 0, PTi, dffn, N);
                    a.C_hi = dHh; a.C_lo = sp ? dHh + pD : nullptr; a.ldch = dffn;
                    a.mask_hi = Hh; a.ldmask_hi = dffn;   // the saved hidden is zero where the ReLU was inactive or the element was dropped
                    if (drop_thr) a.out_scale = drop_scale;
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                CK(wgrad(Hh, pD, dffn, dffn, dRh, pN, N, N, grads + lo[6], dffn, 1)); ++nl;          // dW2 = dF^T Hf, as Hf^T dF stored transposed
                {   // pre: dU = dHf W1 (gradient of norm2's output) ; post: dR = d z2 + dHf W1 (gradient of r1)
                    TmaGemmArgs a = tma_nt_args(dHh, sp ? dHh + pD : nullptr, dffn, thi + lo[4], tlo + lo[4], dffn, pre ? dU : dR, N, PTi, N, dffn);
                    if (!pre) { a.accumulate = 1; a.Cin = dU; }
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                CK(wgrad(dHh, pD, dffn, dffn, U2h, pN, N, N, grads + lo[4], N, 0)); ++nl;             // dW1 = dHf^T (FFN input)
                CK(launch_colsum_planes(dHh, sp ? dHh + pD : nullptr, g.PT, dffn, grads + lo[5], st)); ++nl;
                const float* dA = dR;   // gradient of the (dropped) attention output
                if (pre) {   // LayerNorm 2 (input Rmid): dR += d Rmid
                    CK(launch_ln_bwd(dU, at<float>(ws, t.Rmid), dU, dR, params + lo[10], g.PT, N, 1e-6f, grads + lo[10], grads + lo[11], st)); ++nl;
                } else {     // norm1 (input z1 = Rmid): dU <- d z1
                    CK(launch_ln_bwd(dR, at<float>(ws, t.Rmid), dU, nullptr, params + lo[8], g.PT, N, 1e-6f, grads + lo[8], grads + lo[9], st)); ++nl;
                    dA = dU;
                }
                // attention branch: z1 / Rmid = Rin + dropout(attn(.) W_o^T + b_o)
                CK(launch_split_rows_colsum(dA, N, dRh, sp ? dRh + pN : nullptr, g.PT, N, grads + lo[3], st, drop_thr, drop_site_key(h->drop_seed, li, 1),
                                            drop_scale)); ++nl;   // planes of the (masked) gradient + dbo
                {
                    TmaGemmArgs a = tma_nt_args(dRh, sp ? dRh + pN : nullptr, N, thi + lo[2], tlo + lo[2], N, dOa, N, PTi, N, N);
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                CK(wgrad(dRh, pN, N, N, Oh, pN, N, N, grads + lo[2], N, 0)); ++nl;                      // dWo = dA^T O
                if (attn_bwd_mma_supported(N, heads, m)) {
                    CK(launch_attn_bwd_mma(at<float>(ws, t.QKV), at<float>(ws, t.Oa), at<float>(ws, t.LSE), dOa, dQKV, N, heads, m, sp, st, drop_thr,
                                           drop_site_key(h->drop_seed, li, 0), drop_scale)); ++nl;
                } else {
                    CK(launch_attn_bwd(at<float>(ws, t.QKV), at<float>(ws, t.Oa), at<float>(ws, t.LSE), dOa, dQKV, N, heads, m, st)); ++nl;
                }
                CK(launch_split_rows_colsum(dQKV, 3 * N, dQh, sp ? dQh + pQ : nullptr, g.PT, 3 * N, grads + lo[1], st)); ++nl;  // + db_in
                {   // pre: dU = dQKV Win (gradient of norm1's output) ; post: dR = d z1 + dQKV Win (gradient of the layer input)
                    TmaGemmArgs a = tma_nt_args(dQh, sp ? dQh + pQ : nullptr, 3 * N, thi + lo[0], tlo + lo[0], 3 * N, pre ? dU : dR, N, PTi, N, 3 * N);
                    if (!pre) { a.accumulate = 1; a.Cin = dU; }
                    CK(launch_gemm_tma_nt(a, sp, st)); ++nl;
                }
                CK(wgrad(dQh, pQ, 3 * N, 3 * N, U1h, pN, N, N, grads + lo[0], N, 0)); ++nl;            // dWin = dQKV^T (layer input / its norm)
                if (pre) {   // LayerNorm 1 (input Rin): dR += d Rin
                    CK(launch_ln_bwd(dU, at<float>(ws, t.Rin), dU, dR, params + lo[8], g.PT, N, 1e-6f, grads + lo[8], grads + lo[9], st)); ++nl;
                }
                continue;
            }
            float* Hf = at<float>(ws, t.Hf);
            // FFN: out = Rmid + relu(U2 W1^T + b1) W2^T + b2
            {
                GemmNtArgs a = nt_args(dR, N, whi + lo[6], wlo + lo[6], dffn, 1, dHf, dffn, PTi, dffn, N);
                a.mask = Hf; a.ldmask = dffn;
                CK(launch_gemm_nt(a, sp, st)); ++nl;
                GemmTnArgs w2 = tn_args(dR, N, Hf, dffn, grads + lo[6], dffn, PTi, N, dffn);
                CK(launch_gemm_tn(w2, sp, st)); ++nl;
                CK(launch_colsum_any(dR, N, PTi, N, 1.f, grads + lo[7], st)); ++nl;
                GemmNtArgs b1 = nt_args(dHf, dffn, whi + lo[4], wlo + lo[4], N, 1, dU, N, PTi, N, dffn);
                CK(launch_gemm_nt(b1, sp, st)); ++nl;
                GemmTnArgs w1 = tn_args(dHf, dffn, at<float>(ws, t.U2), N, grads + lo[4], N, PTi, dffn, N);
                CK(launch_gemm_tn(w1, sp, st)); ++nl;
                CK(launch_colsum_any(dHf, dffn, PTi, dffn, 1.f, grads + lo[5], st)); ++nl;
            }
            // LayerNorm 2 (input Rmid): dR += d Rmid
            CK(launch_ln_bwd(dU, at<float>(ws, t.Rmid), dU, dR, params + lo[10], g.PT, N, 1e-6f, grads + lo[10], grads + lo[11], st)); ++nl;
            // attention branch: Rmid = Rin + attn(U1) W_o^T + b_o
            {
                float* Oa = at<float>(ws, t.Oa);
                GemmNtArgs a = nt_args(dR, N, whi + lo[2], wlo + lo[2], N, 1, dOa, N, PTi, N, N);
                CK(launch_gemm_nt(a, sp, st)); ++nl;
                GemmTnArgs wo = tn_args(dR, N, Oa, N, grads + lo[2], N, PTi, N, N);
                CK(launch_gemm_tn(wo, sp, st)); ++nl;
                CK(launch_colsum_any(dR, N, PTi, N, 1.f, grads + lo[3], st)); ++nl;
                CK(launch_attn_bwd(at<float>(ws, t.QKV), Oa, at<float>(ws, t.LSE), dOa, dQKV, N, heads, m, st)); ++nl;
                GemmNtArgs b = nt_args(dQKV, 3 * N, whi + lo[0], wlo + lo[0], N, 1, dU, N, PTi, N, 3 * N);
                CK(launch_gemm_nt(b, sp, st)); ++nl;
                GemmTnArgs wi = tn_args(dQKV, 3 * N, at<float>(ws, t.U1), N, grads + lo[0], N, PTi, 3 * N, N);
                CK(launch_gemm_tn(wi, sp, st)); ++nl;
                CK(launch_colsum_any(dQKV, 3 * N, PTi, 3 * N, 1.f, grads + lo[1], st)); ++nl;
            }
            // LayerNorm 1 (input Rin): dR += d Rin
            CK(launch_ln_bwd(dU, at<float>(ws, t.Rin), dU, dR, params + lo[8], g.PT, N, 1e-6f, grads + lo[8], grads + lo[9], st)); ++nl;
        }
        // R0 = Xin + PE: the gradient of the path input is the residual branch (dX) plus dR
        CK(launch_axpy(dX, dR, 1.f, nPN, st)); ++nl;
    }
    // ---- segmentation backward = overlap-add; masknet.conv1d; masknet.norm; encoder ReLU + weight
    float* dFb = dF2;
    CK(launch_overlap_add_cl(dX, dFb, B, g.L, g.K, g.Sc, N, st)); ++nl;
    {
        GemmNtArgs a = nt_args(dFb, N, whi + o[3], wlo + o[3], N, 1, dEn, N, BLi, N, N);
        CK(launch_gemm_nt(a, sp, st)); ++nl;
        GemmTnArgs t = tn_args(dFb, N, at<float>(ws, l.En), N, grads + o[3], N, BLi, N, N);
        CK(launch_gemm_tn(t, sp, st)); ++nl;
    }
    CK(cudaMemsetAsync(red, 0, 2 * B * sizeof(double), st));
    CK(launch_gn_bwd_reduce_any(dEn, E, at<float>(ws, l.mrE), params + o[1], g.BL, g.L, N, red, grads + o[1], grads + o[2], st)); ++nl;
    float* dEg = at<float>(ws, l.dGt);  // scratch [B*L, N]
    CK(launch_gn_bwd_apply(dEn, E, dEg, at<float>(ws, l.mrE), red, params + o[1], g.BL, g.L, N, st)); ++nl;
    CK(launch_relu_bwd_add(dE, dEg, E, dE, g.BL * N, st)); ++nl;   // both uses of the encoder output, through its ReLU
    {
        GemmTnArgs t = tn_args(dE, N, at<float>(ws, l.xp), stride, grads + o[0], win, BLi, N, win);
        t.b_rpb = g.L; t.b_skip = g.Tp8 / stride - g.L;
        CK(launch_gemm_tn(t, sp, st)); ++nl;
    }
    h->launches = nl;
    return 0;
}

}  // extern "C"
