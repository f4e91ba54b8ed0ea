// Definitions shared by the SepFormer inference engine (sepformer.cu) and training engine (sepformer_train.cu).
#pragma once
#include <vector>

#include "../../include/dualpath_b200.h"
#include "engine_common.h"

struct dp_sepformer {
    dp_sepformer_config cfg;
    std::vector<int64_t> off;
    int64_t n_params;
    int launches;
    float drop_p = 0.f;        // training-time dropout of the transformer layers (dp_sepformer_set_dropout); 0 = off
    uint32_t drop_seed = 0;
};

namespace {

constexpr int HEAD = DP_SEPFORMER_HEAD_PARAMS;
constexpr int PER_LAYER = DP_SEPFORMER_LAYER_PARAMS;

inline int path_entries(int layers) { return 1 + PER_LAYER * layers + 4; }

struct SGeo {
    int B, T, Tp8, L, rest, Sc, K, P;
    long long PT, BL;
};
inline int sep_geo(const dp_sepformer* h, int B, int T, SGeo& g) {
    const int win = h->cfg.win, st = win / 2;
    if (B <= 0) return fail("need B > 0 (got %d)", B);
    if (T < win) return fail("Sepformer needs at least %d samples (got T=%d): the encoder has no padding (sepformer.py:23-40)", win, T);
    g.B = B; g.T = T; g.K = h->cfg.chunk;
    g.L = (T - win) / st + 1;
    g.Tp8 = ((T + st - 1) / st) * st;
    if (dp_seg_geometry(g.L, g.K, &g.rest, &g.Sc)) return 1;
    g.P = g.Sc * g.K;
    g.PT = (long long)B * g.P;
    g.BL = (long long)B * g.L;
    const int dmax = h->cfg.intra_dffn > h->cfg.inter_dffn ? h->cfg.intra_dffn : h->cfg.inter_dffn;
    if (g.PT * (long long)(dmax > 3 * h->cfg.enc_dim ? dmax : 3 * h->cfg.enc_dim) >= 0x7fffffffLL)
        return fail("batch too large for 32-bit row indexing (B=%d T=%d)", B, T);
    return 0;
}

}  // namespace
