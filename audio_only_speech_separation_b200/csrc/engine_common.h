// Host-side helpers shared by the model engines (api.cu: TasNet DPRNN/DPTNet, sepformer.cu: SepFormer).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "kernels.h"

namespace dp {

// error reporting of the C-ABI (message readable through dp_last_error())
int fail(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define CK(call)                                            \
    do {                                                    \
        cudaError_t _e = (call);                            \
        if (_e != cudaSuccess) return dp::cuda_fail(_e, #call); \
    } while (0)

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool is_split(int precision) { return precision == 0; }  // DP_PREC_FP32

inline GemmNtArgs nt_args(const float* A, long long lda, const __nv_bfloat16* whi, const __nv_bfloat16* wlo, int ldw, int w_kn, float* C,
                          int ldc, int M, int N, int K) {
    GemmNtArgs a;
    memset(&a, 0, sizeof(a));
    a.A = A; a.lda = lda; a.Whi = whi; a.Wlo = wlo; a.ldw = ldw; a.w_kn = w_kn; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
    a.bias_scale = 1.f;
    return a;
}
inline GemmTnArgs tn_args(const float* A, int lda, const float* B, long long ldb, float* C, int ldc, int P, int Mo, int No) {
    GemmTnArgs a;
    memset(&a, 0, sizeof(a));
    a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc; a.P = P; a.Mo = Mo; a.No = No; a.scale = 1.f;
    return a;
}

// carve 256-byte aligned regions out of one caller-provided workspace
struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    }
};

template <typename T>
T* at(void* base, size_t off) { return reinterpret_cast<T*>(static_cast<char*>(base) + off); }

// backend selected with dp_set_gemm_backend: 0 mma.sync, 1 tcgen05 (converting threads), 2 TMA-fed tcgen05 on operand planes
int gemm_backend();
inline TmaGemmArgs tma_nt_args(const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, long long lda, const __nv_bfloat16* whi,
                               const __nv_bfloat16* wlo, int ldw, float* C, int ldc, int M, int N, int K) {
    TmaGemmArgs a;
    memset(&a, 0, sizeof(a));
    a.A_hi = a_hi; a.A_lo = a_lo; a.lda = lda; a.W_hi = whi; a.W_lo = wlo; a.ldw = ldw;
    a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K; a.bias_scale = 1.f;
    return a;
}
// GEMM dispatch honouring dp_set_gemm_backend (defined in api.cu)
cudaError_t gemm_nt(const GemmNtArgs& a, bool split, cudaStream_t st);
cudaError_t gemm_tn(const GemmTnArgs& a, bool split, cudaStream_t st);

}  // namespace dp
