// Micro-benchmarks that size the cluster / tcgen05 recurrence design (round 2): not product code, not a test.
//   A  tcgen05.mma issue cost vs N (M = 128, K = 16, bf16), shared- or tensor-memory A operand, 1..4 accumulators
//   B  the dependent chain of one recurrent step: MMAs -> commit -> tcgen05.ld -> st.shared -> fence -> arrive -> next MMAs
//   C  distributed-shared-memory exchange of an h slice among the 4 CTAs of a cluster + cluster barrier / remote mbarrier
//   D  tcgen05.ld bandwidth
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../audio_only_speech_separation_b200/csrc ubench_tc5.cu -o ubench_tc5
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc5_common.cuh"

using namespace dp;

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

// ------------------------------------------------------------------------------------------------ A: MMA issue cost
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int chain, int rounds, int nacc, int ts, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) tmem_alloc(&slot, 512);
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1) {
        const uint32_t idesc = idesc_bf16(128, N, 0, 0);
        const uint64_t da = desc_sw128(smem_u32(smem), 16, 1024);
        const uint64_t db = desc_sw128(smem_u32(smem + 16384), 16, 1024);
        long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
            const uint32_t d = tmem + 256 + (uint32_t)((r % nacc) * 64);  // accumulators at columns 256.. (N <= 64 when nacc > 1)
            for (int k = 0; k < chain; ++k) {
                if (ts) umma_ts_w(d, tmem + (uint32_t)((k & 3) * 8), db + (uint64_t)((k & 3) * 2), idesc, k != 0);
                else umma_w(d, da + (uint64_t)((k & 3) * 2), db + (uint64_t)((k & 3) * 2), idesc, k != 0);
            }
        }
        umma_commit_w(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if ((tid & 31) == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ B: dependent step chain
// warp 0: MMA issuer ; warps 1..4 (+ 5..8 when epi8): epilogue (TMEM lane quadrant = warp % 4, columns split among the two halves)
__global__ void __launch_bounds__(288, 1) step_chain_kernel(int N, int nmma, int steps, int epi8, int mufu, long long* out, float* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t d_full, h_ready;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nepi = epi8 ? 8 : 4;
    for (int i = tid; i < (16384 + 32768) / 4; i += 288) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) {
        mbar_init(&d_full, 1);
        mbar_init(&h_ready, nepi * 32);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&slot, 512);
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 0) {
        const uint32_t idesc = idesc_bf16(128, N, 0, 0);
        const uint64_t da = desc_sw128(smem_u32(smem), 16, 1024);
        const uint64_t db = desc_sw128(smem_u32(smem + 16384), 16, 1024);
        long long t0 = clock64();
        for (int s = 0; s < steps; ++s) {
            if (s > 0) mbar_wait(&h_ready, (s - 1) & 1);
            tc_fence_after();
            for (int k = 0; k < nmma; ++k) umma_w(tmem + 256, da + (uint64_t)((k & 3) * 2), db + (uint64_t)((k & 3) * 2), idesc, k != 0);
            umma_commit_w(&d_full);
            __syncwarp();
        }
        mbar_wait(&h_ready, (steps - 1) & 1);
        long long t1 = clock64();
        if (lane == 0) out[blockIdx.x] = t1 - t0;
    } else if (warp <= nepi) {
        const int q = warp & 3, half = (warp - 1) >> 2;
        const int c0 = epi8 ? half * (N / 2) : 0, nc = epi8 ? N / 2 : N;
        const uint32_t addr = tmem + ((uint32_t)(q * 32) << 16) + 256 + c0;
        float acc = 0.f;
        for (int s = 0; s < steps; ++s) {
            mbar_wait(&d_full, s & 1);
            tc_fence_after();
            for (int c = 0; c < nc; c += 8) {
                float v[8];
                tmem_ld8_nowait(addr + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float x = v[i] + acc;
                    if (mufu) x = rcp_approx(1.0f + ex2_approx(-1.44f * x));   // one "gate" worth of MUFU per accumulator element
                    acc = x;
                    reinterpret_cast<__nv_bfloat16*>(smem + 16384)[((c0 + c + i) * 64 + (q * 32 + lane) % 64)] = __float2bfloat16_rn(x);
                }
            }
            proxy_fence_async();
            tc_fence_before();
            mbar_arrive(&h_ready);
        }
        sink[blockIdx.x * 288 + tid] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ C: DSMEM exchange
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > SPIN_LIMIT) __trap();
    }
}

// mode 0: st.shared::cluster to the (CS-1) peers + barrier.cluster ; mode 1: same stores + remote mbarrier arrive (one per warp per peer) ;
// mode 2: barrier.cluster only (no data) ; mode 3: remote mbarrier only
template <int CS>
__global__ void __launch_bounds__(256, 1) dsmem_kernel(int bytes, int steps, int mode, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[2];
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    if (tid == 0) {
        mbar_init(&bar[0], (CS - 1) * 8);
        mbar_init(&bar[1], (CS - 1) * 8);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();
    const uint32_t base = smem_u32(smem);
    const int nvec = bytes / 16;  // 16-byte vectors per peer
    long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
        const uint32_t buf = base + (uint32_t)((s & 1) * 32768) + rank * (uint32_t)bytes;
        if (mode <= 1) {
            for (int pr = 1; pr < CS; ++pr) {
                const uint32_t dst = mapa(buf, (rank + pr) % CS);
                for (int i = tid; i < nvec; i += 256) st_cluster_v4(dst + i * 16, make_uint4(s, i, tid, pr));
            }
        }
        if (mode == 0 || mode == 2) {
            cluster_sync_all();
        } else {
            __syncwarp();
            if (lane == 0) {
                for (int pr = 1; pr < CS; ++pr) mbar_arrive_remote(mapa(smem_u32(&bar[s & 1]), (rank + pr) % CS));
            }
            mbar_wait_cluster(&bar[s & 1], (s >> 1) & 1);
        }
    }
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------ D: tcgen05.ld bandwidth
__global__ void __launch_bounds__(256, 1) ldtm_kernel(int nwarps, int iters, long long* out, float* sink) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    float acc = 0.f;
    long long t0 = clock64();
    if (warp < nwarps) {
        const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
        for (int it = 0; it < iters; ++it) {
            float v[32];
#pragma unroll
            for (int c = 0; c < 128; c += 32) {
                tmem_ld32(addr + c, v);
                acc += v[0] + v[31];
            }
        }
    }
    long long t1 = clock64();
    sink[blockIdx.x * 256 + tid] = acc;
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}


// ------------------------------------------------------------------------------------------------ E: tight issue loop, M = 64 / 128
// One elect, 8 MMAs per asm block (descriptors advanced inside the asm), so the issue cost per MMA is a few cycles.
__device__ __forceinline__ void umma_x8(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n .reg .pred q;\n .reg .b64 a, b;\n elect.sync _|q, 0xffffffff;\n mov.b64 a, %1;\n mov.b64 b, %2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n add.u64 a, a, 2;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n add.u64 a, a, 2;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n add.u64 a, a, 2;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n sub.u64 a, a, 6;\n sub.u64 b, b, 6;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n add.u64 a, a, 2;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n add.u64 a, a, 2;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n add.u64 a, a, 2;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc)
        : "memory");
}
__global__ void __launch_bounds__(128, 1) mma_tight_kernel(int M, int N, int rounds, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) tmem_alloc(&slot, 512);
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1) {
        const uint32_t idesc = idesc_bf16(M, N, 0, 0);
        const uint64_t da = desc_sw128(smem_u32(smem), 16, 1024);
        const uint64_t db = desc_sw128(smem_u32(smem + 16384), 16, 1024);
        long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) umma_x8(tmem + 256, da, db, idesc);
        umma_commit_w(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if ((tid & 31) == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ F: accumulator layout for M = 64
// A[r][k=0] = r + 1 (others 0), B[n][k=0] = 1  ->  D[r][n] = r + 1; tensor memory preset to -1; dump which lane holds which row.
__global__ void __launch_bounds__(128, 1) layout_kernel(int M, int N, float* dump) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    __syncthreads();
    for (int r = tid; r < 128; r += 128) {
        *reinterpret_cast<__nv_bfloat16*>(smem + r * 128 + ((0 ^ (r & 7)) << 4)) = __float2bfloat16_rn((float)(r + 1));
    }
    for (int n = tid; n < 256; n += 128) *reinterpret_cast<__nv_bfloat16*>(smem + 16384 + n * 128 + ((0 ^ (n & 7)) << 4)) = __float2bfloat16_rn(1.0f);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) tmem_alloc(&slot, 512);
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    {
        uint32_t v[32];
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-1.0f);
        tmem_st32(tmem + ((uint32_t)(warp * 32) << 16), v);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        umma_w(tmem, desc_sw128(smem_u32(smem), 16, 1024), desc_sw128(smem_u32(smem + 16384), 16, 1024), idesc_bf16(M, N, 0, 0), 0);
        umma_commit_w(&bar);
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
    for (int i = 0; i < 32; ++i) dump[(warp * 32 + lane) * 32 + i] = v[i];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
    const bool only_new = argc > 1;
    long long* out;
    float* sink;
    CK(cudaMalloc(&out, 1024 * sizeof(long long)));
    CK(cudaMalloc(&sink, 148 * 512 * sizeof(float)));
    std::vector<long long> h(1024);
    const int SMEM = 16384 + 32768 + 1024;
    CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    CK(cudaFuncSetAttribute(step_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));

    if (!only_new) {
    printf("== A: tcgen05.mma M=128 K=16 bf16: cycles per MMA (chain=24, rounds=64; issue .. completion of all)\n");
    const int Ns[] = {16, 32, 48, 64, 80, 96, 128, 256};
    for (int ts = 0; ts < 2; ++ts)
        for (int nacc = 1; nacc <= 4; nacc *= 2)
            for (int N : Ns) {
                if (nacc > 1 && N > 64) continue;
                for (int grid : {1, 148}) {
                    mma_rate_kernel<<<grid, 128, SMEM>>>(N, 24, 64, nacc, ts, out);
                    CK(cudaDeviceSynchronize());
                    mma_rate_kernel<<<grid, 128, SMEM>>>(N, 24, 64, nacc, ts, out);
                    CK(cudaDeviceSynchronize());
                    CK(cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                    long long mx = 0;
                    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                    printf("A ts=%d nacc=%d N=%3d grid=%3d : %.1f cyc/MMA (floor 128*N/256 = %d)\n", ts, nacc, N, grid, (double)mx / (24.0 * 64), N / 2);
                }
            }

    printf("== B: dependent step chain (issuer -> commit -> ld -> sts -> fence -> arrive), cycles per step\n");
    for (int mufu = 0; mufu < 2; ++mufu)
        for (int epi8 = 0; epi8 < 2; ++epi8)
            for (int N : {32, 48, 64, 80, 96, 128})
                for (int nmma : {1, 8, 24, 36}) {
                    step_chain_kernel<<<148, 288, SMEM>>>(N, nmma, 200, epi8, mufu, out, sink);
                    CK(cudaDeviceSynchronize());
                    step_chain_kernel<<<148, 288, SMEM>>>(N, nmma, 200, epi8, mufu, out, sink);
                    CK(cudaDeviceSynchronize());
                    CK(cudaMemcpy(h.data(), out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
                    long long mx = 0;
                    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
                    printf("B mufu=%d epi_warps=%d N=%3d nmma=%2d : %.0f cyc/step\n", mufu, epi8 ? 8 : 4, N, nmma, (double)mx / 200);
                }

    printf("== C: DSMEM exchange, cycles per step (256 threads per CTA)\n");
    {
        const int DS = 65536 + 1024;
        CK(cudaFuncSetAttribute(dsmem_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, DS));
        CK(cudaFuncSetAttribute(dsmem_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DS));
        for (int cs : {2, 4})
            for (int mode = 0; mode < 4; ++mode)
                for (int bytes : {2048, 4096, 8192}) {
                    if (mode >= 2 && bytes != 2048) continue;
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(cs == 4 ? 144 : 148);
                    cfg.blockDim = dim3(256);
                    cfg.dynamicSmemBytes = DS;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    for (int rep = 0; rep < 2; ++rep) {
                        if (cs == 4) CK(cudaLaunchKernelEx(&cfg, dsmem_kernel<4>, bytes, 200, mode, out));
                        else CK(cudaLaunchKernelEx(&cfg, dsmem_kernel<2>, bytes, 200, mode, out));
                        CK(cudaDeviceSynchronize());
                    }
                    const int grid = cfg.gridDim.x;
                    CK(cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                    long long mx = 0, mn = 1ll << 60;
                    for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
                    printf("C cluster=%d mode=%d bytes_per_peer=%5d : max %.0f min %.0f cyc/step\n", cs, mode, bytes, (double)mx / 200, (double)mn / 200);
                }
    }

    printf("== D: tcgen05.ld 32x32b.x32, 128 columns per warp per iteration\n");
    for (int nw : {1, 4, 8}) {
        ldtm_kernel<<<148, 256>>>(nw, 1000, out, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("D warps=%d : %.1f cyc per 128-col read per warp -> %.0f B/cyc/SM\n", nw, (double)mx / 1000, nw * 32 * 128 * 4.0 / ((double)mx / 1000));
    }

    }
    printf("== E: tight issue loop (8 MMAs per elect), cycles per MMA\n");
    CK(cudaFuncSetAttribute(mma_tight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    for (int M : {64, 128})
        for (int N : {16, 32, 64, 128, 192, 256}) {
            mma_tight_kernel<<<148, 128, SMEM>>>(M, N, 256, out);
            CK(cudaDeviceSynchronize());
            mma_tight_kernel<<<148, 128, SMEM>>>(M, N, 256, out);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("E M=%3d N=%3d : %.1f cyc/MMA\n", M, N, (double)mx / (256.0 * 8));
        }
    printf("== F: accumulator layout (row value r+1 found at lane, col 0 / col 1 / col 8)\n");
    CK(cudaFuncSetAttribute(layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    {
        float* dump;
        CK(cudaMalloc(&dump, 128 * 32 * sizeof(float)));
        std::vector<float> hd(128 * 32);
        for (int M : {64, 128}) {
            layout_kernel<<<1, 128, SMEM>>>(M, 32, dump);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(hd.data(), dump, hd.size() * sizeof(float), cudaMemcpyDeviceToHost));
            printf("F M=%d N=32: lane:value(col0,col1,col31) ", M);
            for (int l = 0; l < 128; ++l) printf("%d:%g,%g,%g ", l, hd[l * 32], hd[l * 32 + 1], hd[l * 32 + 31]);
            printf("\n");
        }
    }
    printf("done\n");
    return 0;
}
