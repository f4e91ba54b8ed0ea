/* CPU oracle (TEST INFRASTRUCTURE ONLY -- never linked into the product):
 * plain-C restatement of the reference's integer index maps for the
 * 50%-overlap segmentation and overlap-add of the dual-path wrapper.
 *
 * Follows look2hear/models/utils/gc3_basics.py:63-91 (pad_segment +
 * split_feature) and :94-109 (merge_feature); the SepFormer copies
 * (look2hear/models/sepformer.py:762-846) use the same maps.  It walks the
 * reference's own construction (pad, two half-shifted block views, interleave)
 * rather than the closed form the CUDA kernels use, so the two are independent.
 * Pinned against tests/golden/seg_ola.npz (generated from the reference).
 */
#include <stdlib.h>
#include <string.h>

int oracle_seg_rest(int L, int K) { int P = K / 2; return K - (P + L % K) % K; }

int oracle_num_chunks(int L, int K) {
    int P = K / 2, Lp = L + oracle_seg_rest(L, K) + 2 * P;
    return 2 * ((Lp - P) / K);
}

/* x[B*N][L] -> y[B*N][K][S] */
void oracle_segment_f32(const float* x, float* y, int rows, int L, int K) {
    int P = K / 2, rest = oracle_seg_rest(L, K), Lp = L + rest + 2 * P, S = oracle_num_chunks(L, K);
    int half = S / 2; /* blocks per half-shifted view */
    float* pad = (float*)calloc((size_t)Lp, sizeof(float));
    for (int r = 0; r < rows; ++r) {
        memset(pad, 0, (size_t)Lp * sizeof(float));
        memcpy(pad + P, x + (size_t)r * L, (size_t)L * sizeof(float));
        for (int j = 0; j < half; ++j)
            for (int k = 0; k < K; ++k) {
                /* block1 = pad[:-P] viewed [half][K]; block2 = pad[P:] viewed [half][K];
                 * cat(...,3).view(-1,K) interleaves them: chunk 2j <- block1[j], 2j+1 <- block2[j] */
                y[((size_t)r * K + k) * S + 2 * j] = pad[j * K + k];
                y[((size_t)r * K + k) * S + 2 * j + 1] = pad[P + j * K + k];
            }
    }
    free(pad);
}

/* y[B*N][K][S] -> x[B*N][L],  L = (S/2)*K - P - rest */
void oracle_overlap_add_f32(const float* y, float* x, int rows, int K, int S, int rest) {
    int P = K / 2, half = S / 2, full = half * K, L = full - P - rest;
    float* a = (float*)malloc((size_t)full * sizeof(float));
    float* b = (float*)malloc((size_t)full * sizeof(float));
    for (int r = 0; r < rows; ++r) {
        for (int j = 0; j < half; ++j)
            for (int k = 0; k < K; ++k) {
                a[j * K + k] = y[((size_t)r * K + k) * S + 2 * j];     /* even chunks  */
                b[j * K + k] = y[((size_t)r * K + k) * S + 2 * j + 1]; /* odd chunks   */
            }
        for (int t = 0; t < L; ++t) x[(size_t)r * L + t] = a[P + t] + b[t]; /* input1[P:] + input2[:-P] */
    }
    free(a);
    free(b);
}
